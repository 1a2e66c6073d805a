/*
 * p3cudaskin.h -- C-ABI of libp3cudaskin.so, the B200 (sm_100a) backend for
 * Panda3D's CPU vertex-animation hot path.
 *
 * This is the drop-in boundary.  Everything here is
 * extern "C", plain-old-data, pointers and sizes; no C++/torch types cross it.
 * Every entry point cites the reference interface it replaces (paths are
 * relative to the Panda3D source tree, v1.11.0).
 *
 * The reference has no plugin API at this point of the engine; the de-facto
 * boundary is
 *     CPT(GeomVertexData) GeomVertexData::animate_vertices(bool force, Thread *)
 *         panda/src/gobj/geomVertexData.h:160, geomVertexData.cxx:912-975
 *     void GeomVertexData::update_animated_vertices(CData *, Thread *)
 *         panda/src/gobj/geomVertexData.cxx:1433-1752
 * A Panda-side shim flattens GeomVertexData /
 * TransformBlendTable / SliderTable into the descriptors below and calls
 * these functions in place of the CPU loops.  The library is loaded the way
 * Panda loads its other plugins (load_dso + a well-known C symbol, cf.
 * panda/src/audio/audioManager.cxx:63-104): the well-known symbol is
 * p3cs_get_api().
 *
 * Conventions (identical to the reference):
 *   - matrices are LMatrix4f: row-major 16 x float32, row vectors, v' = v * M,
 *     translation in row 3 (panda/src/linmath/lmatrix4_src.I:743-745);
 *   - numeric_type / contents use the integer values of GeomEnums::NumericType
 *     and GeomEnums::Contents (panda/src/gobj/geomEnums.h:177-212);
 *   - row ranges are SparseArray subranges [begin, end)
 *     (panda/src/putil/sparseArray.h:43-163);
 *   - all functions return 0 (P3CS_OK) or a p3cs_status error code; the text
 *     of the last error of the calling thread is p3cs_last_error().  There is
 *     NO CPU fallback: without a usable CUDA device every compute entry point
 *     fails with P3CS_ERR_NO_DEVICE / P3CS_ERR_CUDA.
 */
#ifndef P3CUDASKIN_H
#define P3CUDASKIN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define P3CS_EXPORT __declspec(dllexport)
#else
#define P3CS_EXPORT __attribute__((visibility("default")))
#endif

#define P3CS_API_VERSION 3

typedef enum p3cs_status {
  P3CS_OK = 0,
  P3CS_ERR_INVALID = 1,      /* bad argument / inconsistent descriptor          */
  P3CS_ERR_UNSUPPORTED = 2,  /* legal in Panda but not handled by this backend   */
  P3CS_ERR_CUDA = 3,         /* a CUDA runtime call failed                       */
  P3CS_ERR_NOMEM = 4,        /* host or device allocation failed                 */
  P3CS_ERR_NO_DEVICE = 5     /* no CUDA device (there is no CPU fallback)        */
} p3cs_status;

/* GeomEnums::NumericType, panda/src/gobj/geomEnums.h:177-190 */
enum {
  P3CS_NT_UINT8 = 0, P3CS_NT_UINT16 = 1, P3CS_NT_UINT32 = 2,
  P3CS_NT_PACKED_DCBA = 3, P3CS_NT_PACKED_DABC = 4,
  P3CS_NT_FLOAT32 = 5, P3CS_NT_FLOAT64 = 6, P3CS_NT_STDFLOAT = 7,
  P3CS_NT_INT8 = 8, P3CS_NT_INT16 = 9, P3CS_NT_INT32 = 10, P3CS_NT_PACKED_UFLOAT = 11
};

/* GeomEnums::Contents, panda/src/gobj/geomEnums.h:195-212 */
enum {
  P3CS_C_OTHER = 0, P3CS_C_POINT = 1, P3CS_C_CLIP_POINT = 2, P3CS_C_VECTOR = 3,
  P3CS_C_TEXCOORD = 4, P3CS_C_COLOR = 5, P3CS_C_INDEX = 6, P3CS_C_MORPH_DELTA = 7,
  P3CS_C_MATRIX = 8, P3CS_C_NORMAL = 9
};

/* CoordinateSystem, panda/src/linmath/coordinateSystem.h (PRC `coordinate-system`);
 * selects the hpr convention of decompose_matrix/compose_matrix used for
 * normals under a uniform scale (geomVertexData.cxx:1825-1827). */
enum {
  P3CS_CS_ZUP_RIGHT = 1, P3CS_CS_YUP_RIGHT = 2, P3CS_CS_ZUP_LEFT = 3, P3CS_CS_YUP_LEFT = 4
};

typedef struct p3cs_context_s *p3cs_context;  /* one CUDA device + stream            */
typedef struct p3cs_mesh_s *p3cs_mesh;        /* static, shared data of a vertex data */
typedef struct p3cs_group_s *p3cs_group;      /* n animated instances of one mesh     */

/* One GeomVertexArrayData of the SOURCE GeomVertexData
 * (panda/src/gobj/geomVertexArrayData.h; bytes = get_read_pointer(true)). */
typedef struct p3cs_array_desc {
  const void *data;    /* host pointer, num_rows * stride bytes; not retained     */
  int32_t stride;      /* GeomVertexArrayFormat::get_stride()                     */
  int32_t is_output;   /* 1 if the array survives get_post_animated_format()
                          (geomVertexFormat.cxx:106-135): such arrays are copied
                          row for row (copy_from, geomVertexData.cxx:1460) and
                          then modified in place; 0 for arrays that hold only the
                          transform_blend / morph-delta columns                   */
} p3cs_array_desc;

/* One GeomVertexColumn (panda/src/gobj/geomVertexColumn.h). */
typedef struct p3cs_column_desc {
  int32_t array;         /* index into p3cs_mesh_desc::arrays                     */
  int32_t start;         /* GeomVertexColumn::get_start(), bytes into the row     */
  int32_t num_values;    /* get_num_values(): 4 for a packed_dcba / packed_dabc word, 3 for packed_ufloat */
  int32_t numeric_type;  /* get_numeric_type(), P3CS_NT_*                          */
  int32_t contents;      /* get_contents(), P3CS_C_*                               */
} p3cs_column_desc;

/* One application of a morph: a (C_morph_delta column, slider) pair with the
 * slider's affected rows, in the iteration order of the reference loop
 * geomVertexData.cxx:1467-1543 (morph column outer, slider index inner).  Base and
 * delta may be of any numeric type and contents (colour, texcoord, point, ...):
 * they are read and written like GeomVertexRewriter / GeomVertexReader do. */
typedef struct p3cs_morph_desc {
  int32_t slider;               /* index into the SliderTable (sliderTable.I:52)  */
  p3cs_column_desc base;        /* column that is modified (array must be output) */
  p3cs_column_desc delta;       /* C_morph_delta column read from the source      */
  int32_t num_row_ranges;       /* SliderTable::get_slider_rows(slider)           */
  const int32_t *row_ranges;    /* 2*num_row_ranges ints: begin,end pairs         */
} p3cs_morph_desc;

/* Flattened static description of one animated GeomVertexData. */
typedef struct p3cs_mesh_desc {
  uint32_t struct_size;         /* sizeof(p3cs_mesh_desc), for versioning         */
  int32_t num_rows;             /* GeomVertexData::get_num_rows()                 */
  int32_t coordinate_system;    /* P3CS_CS_*; 0 = P3CS_CS_ZUP_RIGHT               */

  int32_t num_arrays;
  const p3cs_array_desc *arrays;

  /* Columns transformed by the blend matrices: GeomVertexFormat::get_point(i)
   * then get_vector(i) of the post-animated format (geomVertexData.cxx:1582,
   * 1622).  contents is C_POINT, C_VECTOR or C_NORMAL; any numeric type, 1..4 values: float32 x 3 / x 4 are the
   * table_xform_* fast paths, everything else (float64, integers, packed words, narrow columns) takes the
   * GeomVertexRewriter branches (:1782-1799, 1862-1879) with the GeomVertexColumn packers' conversions. */
  int32_t num_skin_columns;
  const p3cs_column_desc *skin_columns;

  /* The `transform_blend` index column (InternalName::get_transform_blend());
   * any of uint8/uint16/uint32, any stride (geomVertexData.cxx:1574-1750).
   * array < 0 when the vertex data has no TransformBlendTable. */
  p3cs_column_desc blend_column;

  /* TransformBlendTable (transformBlendTable.h:45-157) as CSR.  Entries of a
   * blend are listed in stored order (TransformBlend::get_transform(i),
   * i = 0..n-1), which is the accumulation order of
   * TransformBlend::recompute_result (transformBlend.cxx:228-231). */
  int32_t num_transforms;           /* palette size T (rebuild_index order)       */
  int32_t num_blends;
  const int32_t *blend_offsets;     /* num_blends + 1                             */
  const int32_t *blend_transform;   /* palette index per entry                    */
  const float *blend_weight;        /* TransformBlend::get_weight(i)              */

  /* TransformBlendTable::get_rows() (transformBlendTable.I:82-104). */
  int32_t num_row_ranges;
  const int32_t *row_ranges;        /* 2*num_row_ranges ints                      */

  /* SliderTable (sliderTable.h) + GeomVertexFormat morph records. */
  int32_t num_sliders;              /* S = SliderTable::get_num_sliders()         */
  int32_t num_morphs;
  const p3cs_morph_desc *morphs;
} p3cs_mesh_desc;

/* Optional caller-owned device buffers for a group (all may be NULL => the
 * library allocates).  Lets the host keep palettes/outputs in buffers it
 * shares with a renderer or registers with NCCL. */
typedef struct p3cs_group_buffers {
  void *palettes;          /* device, n * T * 64 bytes                            */
  void *sliders;           /* device, n * S * 4 bytes                             */
  void *const *outputs;    /* num output arrays device pointers, each
                              n * p3cs_group_output_pitch() bytes                 */
} p3cs_group_buffers;

/* Timings of the last p3cs_group_animate (CUDA events on the context stream). */
typedef struct p3cs_timings {
  float blend_ms;          /* K1 (blend + normal matrices): sum over its launches; -1 if unknown */
  float skin_ms;           /* K2 (morph + skin + row write): sum over its launches; -1 if unknown */
  float total_ms;          /* first launch start -> last launch end                     */
  int32_t launches;        /* kernels launched by the call                              */
  int32_t blend_launches;
  int32_t skin_launches;
} p3cs_timings;

/* ---- library ---------------------------------------------------------- */

/* Version of this header the library was built against. */
P3CS_EXPORT int p3cs_api_version(void);
/* Number of CUDA devices visible (0 if none / no driver). */
P3CS_EXPORT int p3cs_device_count(void);
/* Text of the calling thread's last error ("" if none).  Mirrors the
 * reference's convention of logging through a NotifyCategory and returning
 * early (geomVertexData.cxx:1565-1570); the shim forwards this text to
 * gobj_cat.error(). */
P3CS_EXPORT const char *p3cs_last_error(void);

/* ---- context ---------------------------------------------------------- */

/* Binds a CUDA device and creates (or adopts) the stream all work of the
 * context is ordered on.  One context per calling (cull) thread keeps the
 * path re-entrant across meshes (SURVEY.md 8b "Threading"). `stream` may be
 * NULL (the library creates a non-blocking stream) or a cudaStream_t. */
P3CS_EXPORT int p3cs_context_create(int device, void *stream, p3cs_context *out);
P3CS_EXPORT int p3cs_context_destroy(p3cs_context ctx);
/* Blocks until all work queued on the context stream is done. */
P3CS_EXPORT int p3cs_context_sync(p3cs_context ctx);
/* Lets kernels of this context store into memory that lives on `peer_device`
 * (NVLink / NVSwitch peer access).  With it, the output buffers handed to
 * p3cs_group_create may be peer-mapped memory of the rendering GPU: every rank
 * then writes its finished tiles straight into the renderer's buffer, the gather
 * overlaps the skinning tile by tile and needs no separate collective. */
P3CS_EXPORT int p3cs_context_enable_peer_access(p3cs_context ctx, int peer_device);
/* A render target: plain device memory on the context's GPU that other processes can map.
 * p3cs_ipc_export fills a 64-byte handle (cudaIpcMemHandle_t) for a pointer returned by
 * p3cs_device_alloc; another rank's context opens it with p3cs_ipc_open (peer access is
 * enabled lazily) and passes the mapped pointer as an output buffer of p3cs_group_create. */
P3CS_EXPORT int p3cs_device_alloc(p3cs_context ctx, size_t bytes, void **out);
P3CS_EXPORT int p3cs_device_free(p3cs_context ctx, void *ptr);
P3CS_EXPORT int p3cs_device_read(p3cs_context ctx, const void *device_src, void *host_dst, size_t bytes);
P3CS_EXPORT int p3cs_ipc_export(p3cs_context ctx, void *ptr, unsigned char handle[64]);
P3CS_EXPORT int p3cs_ipc_open(p3cs_context ctx, const unsigned char handle[64], void **out);
P3CS_EXPORT int p3cs_ipc_close(p3cs_context ctx, void *ptr);

/* ---- "next" row f2: zero-copy hand-off to the renderer ------
 * Today the animated array goes host -> GL every frame (glBufferData / glBufferSubData,
 * glstuff/glGraphicsStateGuardian_src.cxx:7150-7215, driven by the array's _modified stamp,
 * gobj/geomVertexArrayData.cxx:601-606).  p3cs_shareable_alloc returns device memory together with a POSIX
 * file descriptor the renderer imports as its vertex buffer (Vulkan VK_KHR_external_memory_fd, OpenGL
 * GL_EXT_memory_object_fd: glImportMemoryFdEXT + glNamedBufferStorageMemEXT); passed as
 * p3cs_group_buffers::outputs, the animated rows are born inside that buffer.  *allocated_bytes is the size
 * the importer must quote (bytes rounded up to the allocation granularity).  The caller owns the fd
 * (close() it once imported).  p3cs_shareable_import is the CUDA-side import of such an fd (another
 * process or context; also what the tests use to play the renderer). */
P3CS_EXPORT int p3cs_shareable_alloc(p3cs_context ctx, size_t bytes, void **out_ptr, int *out_fd, size_t *allocated_bytes);
P3CS_EXPORT int p3cs_shareable_import(p3cs_context ctx, int fd, size_t allocated_bytes, void **out_ptr);
P3CS_EXPORT int p3cs_shareable_free(p3cs_context ctx, void *ptr);
/* The cudaStream_t of the context (for CUDA-event timing by the caller). */
P3CS_EXPORT void *p3cs_context_stream(p3cs_context ctx);

/* ---- mesh: the static half of a GeomVertexData -------------------------- */

/* Uploads arrays, the transform_blend index column, the blend CSR, the rows
 * mask and the sparse morph deltas to HBM.  Replaces the per-frame
 * new_data->copy_from(this, true) of geomVertexData.cxx:1460 by a one-time
 * upload; call again (destroy + create) when GeomVertexArrayData /
 * TransformBlendTable get_modified() changes. Host pointers are not retained. */
P3CS_EXPORT int p3cs_mesh_create(p3cs_context ctx, const p3cs_mesh_desc *desc, p3cs_mesh *out);
P3CS_EXPORT int p3cs_mesh_destroy(p3cs_mesh mesh);
/* Number of output arrays (arrays with is_output) and their geometry. */
P3CS_EXPORT int p3cs_mesh_num_outputs(p3cs_mesh mesh);
P3CS_EXPORT int p3cs_mesh_output_stride(p3cs_mesh mesh, int output);
P3CS_EXPORT int p3cs_mesh_num_rows(p3cs_mesh mesh);

/* ---- group: n animated instances of a mesh ----------------------------- */

/* A group is n independent instances (a crowd of Characters sharing one
 * GeomVertexData); n = 1 is the plain Actor case.  Each instance has its own
 * palette (T joint matrices = JointVertexTransform::get_matrix in palette
 * order, char/jointVertexTransform.cxx:56-60), its own slider values and its
 * own output arrays (the `_animated_vertices` of geomVertexData.h:313). */
P3CS_EXPORT int p3cs_group_create(p3cs_mesh mesh, int num_instances,
                                  const p3cs_group_buffers *buffers, p3cs_group *out);
P3CS_EXPORT int p3cs_group_destroy(p3cs_group group);

/* Bytes between consecutive instances in output array `output`
 * (num_rows * stride rounded up to 256). */
P3CS_EXPORT size_t p3cs_group_output_pitch(p3cs_group group, int output);
/* Device pointers (valid until the group is destroyed). */
P3CS_EXPORT void *p3cs_group_palettes_device(p3cs_group group);
P3CS_EXPORT void *p3cs_group_sliders_device(p3cs_group group);
P3CS_EXPORT void *p3cs_group_output_device(p3cs_group group, int output);

/* Host -> device: `count` palettes starting at instance `first`;
 * host_matrices = count * T * 16 floats (UserVertexTransform::set_matrix /
 * CharacterJoint::_skinning_matrix values).  Asynchronous on the context
 * stream when the host memory is pinned. */
P3CS_EXPORT int p3cs_group_set_palettes(p3cs_group group, int first, int count,
                                        const float *host_matrices);
/* Host -> device slider values, count * S floats (VertexSlider::get_slider). */
P3CS_EXPORT int p3cs_group_set_sliders(p3cs_group group, int first, int count,
                                       const float *host_values);

/* The hot path: recomputes every blend (update_blend, geomVertexData.cxx:
 * 1550-1556), applies morphs (1462-1543) and skins every point / vector /
 * normal column (1558-1751) of all instances.  Asynchronous. */
P3CS_EXPORT int p3cs_group_animate(p3cs_group group);
/* The same for a subset: only the listed instances (in any order) are recomputed, the others keep their
 * arrays -- the crowd form of the UpdateSeq test in GeomVertexData::animate_vertices
 * (geomVertexData.cxx:937-946), where a character whose joints did not move is not re-animated. */
P3CS_EXPORT int p3cs_group_animate_instances(p3cs_group group, const int32_t *host_instances, int count);
/* A frame's worth of groups (a mixed crowd: one group per character type, what the per-Geom loop of
 * CullableObject::munge_geom, pgraph/cullableObject.cxx:170-175, visits one by one): the groups are independent,
 * so their kernels are forked over side streams of the context and joined back into its stream -- small
 * per-type launches overlap instead of queueing behind each other.  All groups must share one context. */
P3CS_EXPORT int p3cs_animate_groups(p3cs_group const *groups, int num_groups);

/* Device -> host copy of `count` instances of one output array into
 * host_dst (count * num_rows * stride bytes, tightly packed = the byte image
 * of the animated GeomVertexArrayData).  Asynchronous when pinned;
 * follow with p3cs_context_sync. */
P3CS_EXPORT int p3cs_group_read_output(p3cs_group group, int output, int first, int count,
                                       void *host_dst);
/* Per-vertex selected blend index as the kernel saw it (int32 per row, -1 for
 * rows outside TransformBlendTable::get_rows()); parity hook for the
 * bit-exact blend-index / row-selection requirement. */
P3CS_EXPORT int p3cs_group_read_selection(p3cs_group group, int32_t *host_dst);
/* Blended matrices of one instance as computed by K1: num_blends * 16 floats
 * (TransformBlend::get_blend results); parity hook. */
P3CS_EXPORT int p3cs_group_read_blends(p3cs_group group, int instance, float *host_dst);

/* With profiling on, every kernel launch of p3cs_group_animate is bracketed by
 * its own CUDA event pair on the context stream, so blend_ms / skin_ms are the
 * summed per-launch durations (the figures bench.py's roofline uses).  Off by
 * default: then only total_ms is measured (one event pair per call). */
P3CS_EXPORT int p3cs_group_set_profiling(p3cs_group group, int enable);
P3CS_EXPORT int p3cs_group_last_timings(p3cs_group group, p3cs_timings *out);

/* One-call equivalent of GeomVertexData::animate_vertices(force=true) for a
 * single instance with HOST buffers: palette in, sliders in, animated arrays
 * out (host_outputs[i] receives output array i), synchronous. */
P3CS_EXPORT int p3cs_animate_vertices(p3cs_group group, int instance,
                                      const float *host_palette, const float *host_sliders,
                                      void *const *host_outputs);
/* The same for a whole crowd: host palettes (n * T * 16 floats) and sliders (n * S floats, may be NULL) in,
 * host_outputs[i] (n * rows * stride bytes, tightly packed; entries or the array may be NULL) out, for all n
 * instances of the group.  Uploads, kernels and downloads are software-pipelined in chunks of
 * `chunk_instances` instances (<= 0: chosen by the library) over three streams, so after the first chunk the
 * PCIe download never waits.  Asynchronous with respect to the host: follow with p3cs_context_sync.  Host
 * buffers should be page-locked (cudaHostAlloc / p3cs_host_register) or the copies serialise. */
P3CS_EXPORT int p3cs_group_animate_host(p3cs_group group, const float *host_palettes, const float *host_sliders,
                                        void *const *host_outputs, int chunk_instances);
/* Page-locks and device-maps a host buffer (e.g. the result GeomVertexArrayData's VertexDataBuffer, once, as
 * long as it is not reallocated).  When every host_outputs[i] of p3cs_animate_vertices is such a buffer, 16-byte
 * aligned with rows*stride a multiple of 16, the kernel stores the animated rows straight into it over PCIe
 * and no device->host copy follows (the group's device copy of that instance is then left untouched).
 * Unregister before the buffer is freed. */
P3CS_EXPORT int p3cs_host_register(p3cs_context ctx, void *host_ptr, size_t bytes);
P3CS_EXPORT int p3cs_host_unregister(p3cs_context ctx, void *host_ptr);

/* ---- joint hierarchy on the GPU ("next" row f1 of SURVEY.md section 8f) ---- */

typedef struct p3cs_skeleton_s *p3cs_skeleton;

/* What PartBundle::update -> MovingPartBase::do_update -> CharacterJoint::update_internals
 * read, flattened, joints listed parents-first (chan/movingPartBase.cxx:109-166,
 * char/characterJoint.cxx:110-172).  Covers one AnimControl without frame blending
 * (chan/movingPartMatrix.cxx:71-76). */
typedef struct p3cs_skeleton_desc {
  uint32_t struct_size;
  int32_t num_joints;
  int32_t coordinate_system;      /* P3CS_CS_*; convention of compose_matrix */
  int32_t num_palette;            /* T: must equal the mesh's num_transforms */
  const int32_t *parent;          /* [J] index of the parent CharacterJoint, -1 = top-level */
  const int32_t *has_channel;     /* [J] 0: the joint keeps MovingPartMatrix::_default_value */
  const int32_t *table_length;    /* [J][12] AnimChannelMatrixXfmTable tables i j k a b c h p r x y z;
                                     0 = default component (animChannelMatrixXfmTable.I:75-85) */
  const int32_t *table_offset;    /* [J][12] offsets into `tables` */
  int32_t num_table_values;
  const float *tables;
  const float *default_value;     /* [J][16] */
  const float *initial_inverse;   /* [J][16] CharacterJoint::_initial_net_transform_inverse */
  const float *root_xform;        /* [16] PartBundle::get_root_xform() */
  const int32_t *palette_joint;   /* [T] joint behind each palette slot (JointVertexTransform::get_joint) */
  /* Morph sliders driven by scalar channels (CharacterVertexSlider -> CharacterSlider -> AnimChannelScalarTable,
   * char/characterVertexSlider.cxx:56-59, chan/animChannelScalarTable.cxx:88-94), in SliderTable order.  0 sliders:
   * posing leaves the group's slider values alone; otherwise num_sliders must equal the mesh's and every pose call
   * also writes them (MovingPartScalar::get_blend_value, chan/movingPartScalar.cxx:43-105). */
  int32_t num_sliders;
  const int32_t *slider_has_channel;   /* [S] */
  const int32_t *slider_table_length;  /* [S] values in `tables`; 0 = empty table (value 0) */
  const int32_t *slider_table_offset;  /* [S] */
  const float *slider_default;         /* [S] MovingPartScalar::_default_value */
} p3cs_skeleton_desc;

P3CS_EXPORT int p3cs_skeleton_create(p3cs_context ctx, const p3cs_skeleton_desc *desc, p3cs_skeleton *out);
P3CS_EXPORT int p3cs_skeleton_destroy(p3cs_skeleton skeleton);
/* Evaluates the hierarchy of `count` instances at host_frames[i] (AnimControl::get_frame()) and
 * writes their palettes into the group's device buffer -- replaces Character::update +
 * p3cs_group_set_palettes: per frame the host sends one int per instance instead of T matrices. */
P3CS_EXPORT int p3cs_group_pose(p3cs_group group, p3cs_skeleton skeleton, int first, int count, const int32_t *host_frames);
/* The same with PartBundle::set_frame_blend_flag(true): per instance AnimControl::get_frame(), get_next_frame() and
 * (float)get_frac(); every joint takes MovingPartMatrix::get_blend_value's blend of the two frames' values
 * (chan/movingPartMatrix.cxx:78-198) with the bundle's blend type -- PartBundle::BlendType values; the PRC default
 * (anim-blend-type) is normalized-linear.  One AnimControl of effect 1. */
#define P3CS_BT_LINEAR 0
#define P3CS_BT_NORMALIZED_LINEAR 1
#define P3CS_BT_COMPONENTWISE 2
#define P3CS_BT_COMPONENTWISE_QUAT 3  /* rotations blended as quaternions (chan/movingPartMatrix.cxx:273-357) */
P3CS_EXPORT int p3cs_group_pose_blend(p3cs_group group, p3cs_skeleton skeleton, int first, int count, const int32_t *host_frames,
                                      const int32_t *host_next_frames, const float *host_fracs, int blend_type);
/* Several animations on one skeleton and several AnimControls per character (PartBundle::set_anim_blend_flag,
 * set_control_effect): p3cs_skeleton_add_animation uploads the channel tables of one more bound AnimBundle
 * (animation 0 is the one in p3cs_skeleton_desc) and returns its index; p3cs_group_pose_controls evaluates every
 * instance from `num_controls` control states, listed in the order PartBundle iterates its blend map (ascending
 * AnimControl address -- the order MovingPartMatrix::get_blend_value sums in).  Per joint: no control with a channel
 * -> the default value; exactly one and no frame blending -> its value (movingPartBase.cxx:303-333); else the blend
 * weighted by the effects (movingPartMatrix.cxx:78-198). */
typedef struct p3cs_animation_desc {
  uint32_t struct_size;
  int32_t num_table_values;
  const int32_t *has_channel;     /* [J] */
  const int32_t *table_length;    /* [J][12] */
  const int32_t *table_offset;    /* [J][12] */
  const float *tables;
  const int32_t *slider_has_channel;   /* [S] of the skeleton; may be NULL when it has no sliders */
  const int32_t *slider_table_length;  /* [S] */
  const int32_t *slider_table_offset;  /* [S] */
} p3cs_animation_desc;
typedef struct p3cs_control_state {
  int32_t anim;                   /* animation index of this control */
  int32_t frame, next_frame;      /* AnimControl::get_frame(), get_next_frame() */
  float frac;                     /* (float)AnimControl::get_frac(); read only with frame blending */
  float effect;                   /* PartBundle::get_control_effect(); 0 = the control is not in the blend */
} p3cs_control_state;
P3CS_EXPORT int p3cs_skeleton_add_animation(p3cs_skeleton skeleton, const p3cs_animation_desc *desc, int32_t *out_index);
P3CS_EXPORT int p3cs_group_pose_controls(p3cs_group group, p3cs_skeleton skeleton, int first, int count, int num_controls,
                                         const p3cs_control_state *host_states /* [count][num_controls] */,
                                         int frame_blend, int blend_type);
/* Forced channels -- PartBundle::freeze_joint / control_joint (chan/partBundle.h:139-142): a joint with a forced channel
 * takes that channel's value whatever the controls play (MovingPartMatrix::get_blend_value,
 * chan/movingPartMatrix.cxx:53-60).  p3cs_group_set_forced_joints names the joints (indices of the skeleton's joint order;
 * num_forced = 0 releases them all, PartBundle::release_joint) -- every instance starts with the identity;
 * p3cs_group_set_forced_values uploads, for instances [first, first + count), what each forced channel returns
 * (AnimChannelMatrixFixed / AnimChannelMatrixDynamic::get_value: the frozen transform, or the controlling node's
 * transform this frame) as [count][num_forced][16] floats.  Every later p3cs_group_pose* call of the group applies them. */
P3CS_EXPORT int p3cs_group_set_forced_joints(p3cs_group group, p3cs_skeleton skeleton, int num_forced, const int32_t *joints);
P3CS_EXPORT int p3cs_group_set_forced_values(p3cs_group group, int first, int count, const float *host_values);
/* Device -> host read-back of `count` palettes (count * T * 16 floats); parity hook. */
P3CS_EXPORT int p3cs_group_read_palettes(p3cs_group group, int first, int count, float *host_dst);

/* ---- "next" row f4: animated tight bounds ------------------
 * Replaces the CPU walk over the animated vertices behind Character::calc_tight_bounds
 * (char/character.cxx:218-238 -> GeomNode / Geom::calc_tight_bounds ->
 * GeomPrimitive::calc_tight_bounds, gobj/geomPrimitive.cxx:1646-1830, the non-indexed branch without
 * matrix): while a group is animated, the kernel also reduces min / max of `vertex_column` (a float32 x 3
 * column of an output array, normally InternalName::get_vertex()) and max length_squared() over ALL rows of
 * every instance.  NaN components never win a comparison, as with std::min/max there.
 * p3cs_group_enable_bounds(group, NULL) switches it off again. */
P3CS_EXPORT int p3cs_group_enable_bounds(p3cs_group group, const p3cs_column_desc *vertex_column);
/* Restricts the bounds to the rows the Geom's primitives reference -- GeomPrimitive::calc_tight_bounds walks the
 * primitive's vertex indices (gobj/geomPrimitive.cxx:1646-1830, indexed branch), so an unreferenced vertex does not
 * count: bit (r & 31) of host_row_bits[r >> 5] = row r is referenced ((rows + 31) / 32 words, copied).  NULL = every
 * row again. */
P3CS_EXPORT int p3cs_group_set_bounds_rows(p3cs_group group, const uint32_t *host_row_bits);
/* count x 8 floats: min x y z, max x y z, sq_center_dist (max |v|^2), found_any (1.0 / 0.0); of the
 * last p3cs_group_animate.  Synchronous. */
P3CS_EXPORT int p3cs_group_read_bounds(p3cs_group group, int first, int count, float *host_dst);

/* ---- "next" row f3: the AT_hardware upload format (host side, no device work) ----
 * p3cs_hardware_from_blends restates GeomVertexData::copy_from's Panda-table -> hardware-table conversion
 * (gobj/geomVertexData.cxx:574-648, indexed branch): per row (every row, like the reference's loop) the
 * blend's weights and TransformTable slots, <= 4 per vertex -- blends with more entries are cut with
 * TransformBlend::limit_transforms(4) + normalize_weights() (gobj/transformBlend.cxx:95-140); `table`
 * receives, per TransformTable slot in first-use order, the palette index it holds (capacity
 * desc->num_transforms).  Only arrays[blend_column.array] and the blend CSR of `desc` are read. */
P3CS_EXPORT int p3cs_hardware_from_blends(const p3cs_mesh_desc *desc, float *weights /* [num_rows][4] */,
                                          int32_t *indices /* [num_rows][4] */, int32_t *table, int32_t *table_size);
/* The way back, so a mesh stored in the hardware format can feed p3cs_mesh_create: a row's non-zero slots,
 * in slot order, become one blend (identical rows share it).  Capacities: row_blend[num_rows] (write it
 * into a uint32 transform_blend column), blend_offsets[num_rows + 1], blend_transform / blend_weight
 * [4 * num_rows].  For columns produced from <= 4-entry blends the animated result is bit-identical to the
 * original table's. */
P3CS_EXPORT int p3cs_blends_from_hardware(int32_t num_rows, const float *weights, const int32_t *indices, const int32_t *table,
                                          int32_t table_size, int32_t *row_blend, int32_t *blend_offsets,
                                          int32_t *blend_transform, float *blend_weight, int32_t *num_blends);

/* Per-instance destinations of output array `output`: instance i's animated rows (num_rows * stride bytes) go to
 * ptrs[i] instead of the group's own [n][pitch] buffer.  Every pointer is device memory, or host memory page-locked
 * with p3cs_host_register (the kernel's bulk stores then write it over PCIe), 16-byte aligned.  This is how a frame's
 * worth of Characters that share one mesh -- each with its own result GeomVertexData, the `_animated_vertices` the
 * reference reuses in place (panda/src/gobj/geomVertexData.cxx:1448-1453) -- is animated by ONE launch straight into
 * those objects' arrays.  `ptrs` holds n entries and is copied; NULL restores the group's buffer.  The destinations
 * must stay valid until the animate call that uses them has been synchronised. */
P3CS_EXPORT int p3cs_group_set_output_pointers(p3cs_group grp, int output, void *const *ptrs);

/* ---- crowds over several GPUs of one process ----------------------------
 * `cuda-skinning-devices 0 1 2 ...`.  A crowd is num_instances instances of one mesh
 * dealt to the listed devices in consecutive ranges; devices[0] is the rendering GPU and owns the crowd's output
 * buffers ([num_instances][pitch] per output array), which the other devices map over NVLink: their kernels store
 * their finished rows straight into it (no collective, no staging copy).  The deal is sink-aware: devices[0] keeps
 * the share that balances its compute against its inbound link (root_share <= 0: planned from link_gbs, then
 * p3cs_crowd_tune re-splits from measured times).  One context (one stream) per device; animate returns as soon as
 * every device's launch is queued. */
typedef struct p3cs_crowd_s *p3cs_crowd;
typedef struct p3cs_crowd_desc {
  uint32_t struct_size;      /* sizeof(p3cs_crowd_desc) */
  int32_t num_devices;
  const int32_t *devices;    /* CUDA device ordinals; devices[0] = the rendering GPU */
  int32_t num_instances;
  float root_share;          /* share of the instances devices[0] keeps; <= 0: automatic */
  float link_gbs;            /* inbound NVLink rate of devices[0] to plan with; <= 0: 720 */
} p3cs_crowd_desc;

P3CS_EXPORT int p3cs_crowd_create(const p3cs_crowd_desc *desc, const p3cs_mesh_desc *mesh, p3cs_crowd *out);
P3CS_EXPORT int p3cs_crowd_destroy(p3cs_crowd crowd);
P3CS_EXPORT int p3cs_crowd_num_devices(p3cs_crowd crowd);
/* the consecutive range of instances devices[device_index] animates */
P3CS_EXPORT int p3cs_crowd_range(p3cs_crowd crowd, int device_index, int32_t *first, int32_t *count);
/* joint matrices / slider values of instances [first, first + count) (crowd numbering), routed to their devices */
P3CS_EXPORT int p3cs_crowd_set_palettes(p3cs_crowd crowd, int first, int count, const float *host);
P3CS_EXPORT int p3cs_crowd_set_sliders(p3cs_crowd crowd, int first, int count, const float *host);
P3CS_EXPORT int p3cs_crowd_animate(p3cs_crowd crowd);
P3CS_EXPORT int p3cs_crowd_sync(p3cs_crowd crowd);
/* CUDA-event time of the last animate on every device (num_devices floats) */
P3CS_EXPORT int p3cs_crowd_last_ms(p3cs_crowd crowd, float *per_device_ms);
/* measures `steps` passes and re-splits the crowd from the times seen; set palettes / sliders again afterwards */
P3CS_EXPORT int p3cs_crowd_tune(p3cs_crowd crowd, int steps);
/* the gathered buffer of output array `output` on devices[0], and its pitch per instance */
P3CS_EXPORT void *p3cs_crowd_output_device(p3cs_crowd crowd, int output);
P3CS_EXPORT size_t p3cs_crowd_output_pitch(p3cs_crowd crowd, int output);
P3CS_EXPORT int p3cs_crowd_read_output(p3cs_crowd crowd, int output, int first, int count, void *host_dst);

/* ---- plugin table (load_dso + dlsym("p3cs_get_api")) -------------------- */

typedef struct p3cs_api {
  int version;
  int (*device_count)(void);
  const char *(*last_error)(void);
  int (*context_create)(int, void *, p3cs_context *);
  int (*context_destroy)(p3cs_context);
  int (*context_sync)(p3cs_context);
  int (*mesh_create)(p3cs_context, const p3cs_mesh_desc *, p3cs_mesh *);
  int (*mesh_destroy)(p3cs_mesh);
  int (*group_create)(p3cs_mesh, int, const p3cs_group_buffers *, p3cs_group *);
  int (*group_destroy)(p3cs_group);
  int (*group_set_palettes)(p3cs_group, int, int, const float *);
  int (*group_set_sliders)(p3cs_group, int, int, const float *);
  int (*group_animate)(p3cs_group);
  int (*group_read_output)(p3cs_group, int, int, int, void *);
  int (*animate_vertices)(p3cs_group, int, const float *, const float *, void *const *);
  /* version 2 */
  int (*host_register)(p3cs_context, void *, size_t);
  int (*host_unregister)(p3cs_context, void *);
  int (*animate_groups)(p3cs_group const *, int);
  /* version 3 */
  int (*group_set_output_pointers)(p3cs_group, int, void *const *);
  int (*group_animate_instances)(p3cs_group, const int32_t *, int);
} p3cs_api;

P3CS_EXPORT const p3cs_api *p3cs_get_api(void);

#ifdef __cplusplus
}
#endif
#endif /* P3CUDASKIN_H */
