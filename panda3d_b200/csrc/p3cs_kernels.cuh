// p3cs_kernels.cuh -- launch parameters shared by the kernels and the C-ABI host code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p3cs {

constexpr int kMaxOutputs = 4;      // output arrays per mesh
constexpr int kMaxDynColumns = 12;  // skinned and/or morphed columns per mesh

// Per-blend record written by the blend kernel and gathered by the skin kernel.
//   compact (no 4-component skin column): 12 floats M (rows 0..3 x cols 0..2), 9 floats N
//   (normal matrix 3x3), 1 float flag (normalize) = 22 floats = 88 bytes;
//   full: 16 floats M, 16 floats X, flag, padded to 36 floats = 144 bytes.
//   aligned (strip kernel, align-16 layouts): the compact record + column 3 of M = 26 floats = 104 bytes.
constexpr int kRecCompact = 22;
constexpr int kRecAligned = 26;
constexpr int kRecFull = 36;

// A column that changes per frame: skinned (point/vector/normal) and/or a morph base.
struct DynColumn {
  int16_t start;       // byte offset in the row
  int8_t output;       // output array index
  int8_t num_values;   // 1..4
  int8_t skin;         // 0 none, 1 point, 2 vector, 3 normal
  int8_t homogeneous;  // morph: scale deltas by w (C_point / C_texcoord with 4 values)
  int8_t f64;          // NT_float64 components: read/written through float like GeomVertexRewriter::get_data3/set_data3
  uint8_t typed;       // 0: float32 / float64; else an integer or packed column, p3cs_packer.cuh's typed_code()
};

// A plan of the run kernel (p3cs_run.cuh; built by p3cs_runplan.h at upload): tiles of `tile_rows` rows, per tile a
// range of 32-lane slots, per lane of a slot one item = up to `max_item_rows` consecutive rows selecting one blend.
struct RunPlanDev {
  int tile_rows;                  // 0: no plan
  int tile_bufs;                  // 2 or 3 staged tiles in flight per CTA
  int num_tiles;
  int max_item_rows;
  const int *tile_slot_offsets;   // [num_tiles + 1]
  // per lane of a slot, 2 x 16 bytes in two arrays (coalesced 512-byte loads per warp):
  //   a = { item word, joint0 | joint1 << 16, joint2 | joint3 << 16, weight0 },  b = { weight1, weight2, weight3, blend id }
  // item word: first row in the tile | rows << 13 | skinned << 18 | entries << 19 (0..4; 5: more than four, walk the
  // TransformBlendTable's CSR by blend id); 0 = idle lane.  Padded by 8 slots (the kernel prefetches unconditionally).
  const uint4 *item_a;
  const uint4 *item_b;
  // two palette copies in shared memory (p3cs_runplan.h): entries of the palette area (T for one copy) and first entry
  // of copy B (0: one copy); the items' joint fields hold palette-area INDICES
  int pal_entries, copy_b_base;
};

struct MeshDev {
  int num_rows;
  int num_blends;
  int num_transforms;
  int num_sliders;
  int coordinate_system;
  int num_outputs;
  int num_dyn;
  int full_records;  // 1: kRecFull layout
  int index_bytes;   // 2 or 4 (0: no blend table)
  int has_morphs;
  DynColumn dyn[kMaxDynColumns];
  int stride[kMaxOutputs];
  const uint8_t *src[kMaxOutputs];  // static image of each output array (num_rows*stride)
  const void *index;                // dense transform_blend column, u16 or u32 per row
  const uint32_t *rows_mask;        // bit per row; NULL = every row is in get_rows()
  const int *blend_offsets;         // CSR of the TransformBlendTable
  const int *blend_transform;
  const float *blend_weight;
  const int *morph_row_offsets;     // per-row CSR of morph entries (num_rows+1)
  const uint32_t *morph_keys;       // (dyn column << 16) | slider
  const float4 *morph_deltas;

  // ---- strip kernel tables (static, built at upload).  Rows are cut into tiles of kStripThreads * rows_per_thread rows
  // (the phase-B unit) and consecutive tiles into record groups (the phase-A
  // unit): each group lists the distinct blends its skinned rows select ("group records") and
  // every row stores the index of its blend in that list.
  int num_tiles;
  int rows_per_thread;              // 1 or 2: rows a thread animates per staged tile (tile = 256 or 512 rows)
  int tile_bufs;                    // 2 or 3 staged tiles in flight per CTA
  int group_record_cap;             // most records a record group may hold (sizes the shared-memory record area)
  int aligned;                      // vertex-animation-align-16 layout: one output array, every animated column float32 x 4
                                    // on a 16-byte boundary, no morphs -> the strip kernel's kAligned modes
  int strip_rec_floats;             // floats per shared-memory blend record of the strip kernel (22, 26 or 36)
  const int *group_tile_offsets;    // num_groups+1: group g covers tiles [off[g], off[g+1])
  int num_groups;
  int max_group_records;            // max distinct blends of any group (smem sizing)
  int slice_bytes;                  // smem bytes of one staged 32-row slice (all outputs, 128-aligned)
  int slice_out_offset[kMaxOutputs];// byte offset of output o inside a staged slice
  const uint16_t *row_local;        // per row: index into its group's record list, 0xFFFF = not skinned
  const uint32_t *row_local2;       // rows_per_thread == 2: per (tile, thread) the slots of its two rows (t, t + 256), one load
  const int *group_rec_offsets;     // num_groups+1
  const int *group_rec_blend;       // global blend id of every group record (selection hook)
  const int *grec_entry_offsets;    // CSR of the blend entries, in group-record order
  const uint2 *grec_entries;        // (palette index, float weight bits) per entry
  const uint4 *grec_blocks_hi;      // second uint4 of every group record (the two halves are separate arrays: coalesced loads)
  const uint4 *grec_blocks;         // 2 x uint4 per group record: its first four entries, the
                                    // entry count in the high half of the first index

  // ---- run kernel (p3cs_run.cuh): the row-layout modes kRow24 / 32 / 48 / 56
  int run_mode;                     // StripMode the run kernel handles this mesh in, 0: none
  int run_lockstep_pct;             // lane steps / (32 x warp steps) of the crowd plan, in percent (kernel choice)
  RunPlanDev run_plan[2];           // [0] crowds (large tiles), [1] launches of a few instances (small tiles, many CTAs)
};

constexpr int kStripThreads = 256;  // threads of a strip-kernel CTA; a tile is kStripThreads * rows_per_thread rows

struct GroupDev {
  const float *palettes;  // [n][T][16]
  const float *sliders;   // [n][S]
  float *records;         // [chunk][B][rec]
  uint8_t *out[kMaxOutputs];
  size_t pitch[kMaxOutputs];
  uint8_t *const *out_ptrs[kMaxOutputs];  // optional per-instance destinations (p3cs_group_set_output_pointers); NULL: out + instance * pitch
  int32_t *selection;     // optional [num_rows] debug/parity output (instance 0 of the launch)
  // optional animated tight bounds ("next" row f4): [n][8] = min xyz, max xyz, max |v|^2, found_any
  float *bounds;
  int bounds_output;      // output array of the float32 x 3 column the bounds are taken over
  int bounds_start;       // its byte offset in the row
  // optional: the instances to animate (p3cs_group_animate_instances); NULL = first_instance + blockIdx.y
  const int *instance_list;
  // optional: bit r = row r is referenced by a primitive and counts for the bounds (p3cs_group_set_bounds_rows); NULL = every row
  const uint32_t *bounds_rows;
};

// where instance `inst`'s rows of output `o` go
__device__ __forceinline__ uint8_t *output_base(const GroupDev &grp, int o, int inst) {
  if (grp.out_ptrs[o] != nullptr)
    return reinterpret_cast<uint8_t *>(__ldg(reinterpret_cast<const unsigned long long *>(grp.out_ptrs[o]) + inst));
  return grp.out[o] + (size_t)inst * grp.pitch[o];
}

__device__ __forceinline__ bool bounds_wants(const GroupDev &grp, int row) {
  return grp.bounds_rows == nullptr || ((__ldg(grp.bounds_rows + (row >> 5)) >> (row & 31)) & 1u) != 0;
}

// Joint hierarchy ("next" row f1): static description of a character's skeleton, joints parents-first.
constexpr int kMaxAnimations = 8;  // animations (channel-table sets) one skeleton can hold

// The channel tables of one animation bound to a skeleton (AnimBundle -> AnimChannelMatrixXfmTable per joint).
struct AnimationDev {
  const int *has_channel;    // [J]
  const int *table_length;   // [J][12]
  const int *table_offset;   // [J][12]
  const float *tables;
  const int *slider_has_channel;   // [S] scalar channels of the morph sliders (same `tables`)
  const int *slider_table_length;  // [S]
  const int *slider_table_offset;  // [S]
};

struct SkeletonDev {
  int num_joints;
  int coordinate_system;
  int max_level;
  int num_animations;
  const int *parent;          // -1: top-level joint
  const int *level_order;     // joint indices sorted by depth in the tree
  const int *level_begin;     // [max_level+2] ranges of level_order
  const int *path_offsets;    // [J+1]: joint j's chain of ancestors, outermost first, j itself last, is
  const int *path_joints;     //        path_joints[path_offsets[j] .. path_offsets[j+1])
  AnimationDev anim[kMaxAnimations];
  const float *default_value; // [J][16]
  const float *initial_inverse;
  const float *root_xform;    // [16]
  const int *palette_joint;   // [T]
  int num_sliders;            // morph sliders driven by scalar channels (0: posing does not touch slider values)
  const float *slider_default;
};

// One AnimControl of one instance: which animation it plays, AnimControl::get_frame / get_next_frame / get_frac, and
// its effect in PartBundle's blend map (0 = not in the map).  Layout of p3cs_control_state (include/p3cudaskin.h).
struct ControlState {
  int anim, frame, next_frame;
  float frac, effect;
};

// Forced channels of a group (PartBundle::freeze_joint / control_joint): slot[j] = index of joint j among the forced
// ones or -1; values[instance][count][16] = what the forced channel returns for that instance (absolute instance index).
struct ForcedDev {
  const int *slot;
  const float *values;
  int count;
};

// frames_dev != nullptr: one control of animation 0 per instance, no blending.  Else states_dev[count][num_controls].
void launch_pose(const SkeletonDev &skel, float *palettes, int num_transforms, float *sliders, const int *frames_dev,
                 const ControlState *states_dev, int num_controls, int frame_blend, int blend_type, const ForcedDev &forced,
                 int first_instance, int count, cudaStream_t stream);

void launch_blend(const MeshDev &mesh, const GroupDev &grp, int first_instance, int count, cudaStream_t stream);
void launch_skin(const MeshDev &mesh, const GroupDev &grp, int first_instance, int count, cudaStream_t stream);
// The fused path: palette -> smem, per tile: blend records -> smem, rows via TMA bulk copies.
// Returns the number of kernel launches issued (instances are split in grid.y chunks of 65535).
int launch_strip(const MeshDev &mesh, const GroupDev &grp, int first_instance, int count, int sm_count, cudaStream_t stream);
// CTAs per instance launch_strip will use (1 => bounds are stored directly, no initialisation pass needed)
int strip_ctas_per_instance(const MeshDev &mesh, int count, int sm_count);
// bounds: reset [first, first+count) to (+inf, -inf, 0, 0); scan the finished output arrays (two-kernel path)
void launch_bounds_init(const GroupDev &grp, int first_instance, int count, cudaStream_t stream);  // honours instance_list
void launch_bounds_scan(const MeshDev &mesh, const GroupDev &grp, int first_instance, int count, cudaStream_t stream);
size_t strip_smem_bytes(const MeshDev &mesh);
// StripMode the strip kernel would run this mesh in (p3cs_strip.cu); the kRow modes are the run kernel's
int strip_pick_mode(const MeshDev &mesh);
int strip_mode_row_bytes(int mode);  // row stride of a kRow mode, 0 for the others
int skin_rows_per_tile(const MeshDev &mesh);

}  // namespace p3cs
