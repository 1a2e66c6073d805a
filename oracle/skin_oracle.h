/*
 * skin_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of Panda3D 1.11.0's CPU vertex-animation path,
 * used ONLY as the parity checker by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Nothing under
 * panda3d_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  The restatement is checked bit-for-bit against the
 * reference's own GeomVertexData::animate_vertices (built from /root/reference
 * by oracle/build_ref.sh, run by oracle/ref_harness.cxx) on the committed
 * fixtures in tests/golden/ (tests/test_oracle_golden.py).  The reference's own
 * test-suite has no golden vectors for this path.
 */
#ifndef SKIN_ORACLE_H
#define SKIN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* GeomEnums values (panda/src/gobj/geomEnums.h:177-212) */
#define P3O_NT_UINT8 0
#define P3O_NT_UINT16 1
#define P3O_NT_UINT32 2
#define P3O_NT_PACKED_DCBA 3
#define P3O_NT_PACKED_DABC 4
#define P3O_NT_FLOAT32 5
#define P3O_NT_FLOAT64 6
#define P3O_NT_STDFLOAT 7
#define P3O_NT_INT8 8
#define P3O_NT_INT16 9
#define P3O_NT_INT32 10
#define P3O_NT_PACKED_UFLOAT 11
#define P3O_C_POINT 1
#define P3O_C_VECTOR 3
#define P3O_C_TEXCOORD 4
#define P3O_C_COLOR 5
#define P3O_C_NORMAL 9
/* CoordinateSystem (panda/src/linmath/coordinateSystem.h) */
#define P3O_CS_ZUP_RIGHT 1
#define P3O_CS_YUP_RIGHT 2
#define P3O_CS_ZUP_LEFT 3
#define P3O_CS_YUP_LEFT 4

typedef struct p3o_column {
  int32_t array, start, num_values, numeric_type, contents;
} p3o_column;

typedef struct p3o_morph {
  int32_t slider;
  p3o_column base, delta;
  int32_t num_row_ranges;
  const int32_t *row_ranges;
} p3o_morph;

typedef struct p3o_mesh {
  int32_t num_rows;
  int32_t coordinate_system;
  int32_t num_arrays;
  const uint8_t *const *array_data; /* source arrays */
  const int32_t *array_stride;
  const int32_t *array_is_output;
  int32_t num_skin_columns;
  const p3o_column *skin_columns;   /* points first, then vectors */
  p3o_column blend_column;          /* array < 0: no blend table */
  int32_t num_transforms, num_blends;
  const int32_t *blend_offsets, *blend_transform;
  const float *blend_weight;
  int32_t num_row_ranges;
  const int32_t *row_ranges;
  int32_t num_sliders, num_morphs;
  const p3o_morph *morphs;
} p3o_mesh;

/* Restates GeomVertexData::update_animated_vertices (geomVertexData.cxx:
 * 1433-1752).  out_arrays[i] receives the i-th OUTPUT array (arrays with
 * is_output, in order), num_rows*stride bytes each.  blends_out (nullable)
 * receives num_blends*16 floats = TransformBlend::get_blend of every blend;
 * selection_out (nullable) receives the blend index used per row (-1 when the
 * row is outside TransformBlendTable::get_rows()).
 * Returns 0, or -1 for an unsupported format. */
int p3o_animate(const p3o_mesh *mesh, const float *palette, const float *sliders,
                uint8_t *const *out_arrays, float *blends_out, int32_t *selection_out);

/* "Next" row f1: the joint hierarchy that produces the palette.  Joints are listed
 * parents-first.  Restates, for one AnimControl without frame blending,
 * AnimChannelMatrixXfmTable::get_value (chan/animChannelMatrixXfmTable.cxx:112-124),
 * MovingPartMatrix::get_blend_value (chan/movingPartMatrix.cxx:52-76) and
 * CharacterJoint::update_internals (char/characterJoint.cxx:110-150). */
typedef struct p3o_skeleton {
  int32_t num_joints;
  int32_t coordinate_system;
  const int32_t *parent;          /* -1: top-level joint (uses the bundle's root transform) */
  const int32_t *has_channel;     /* 0: keeps MovingPartMatrix::_default_value */
  const int32_t *table_length;    /* [J][12] i j k a b c h p r x y z; 0 = default component */
  const int32_t *table_offset;    /* [J][12] into tables */
  const float *tables;
  const float *default_value;     /* [J][16] */
  const float *initial_inverse;   /* [J][16] CharacterJoint::_initial_net_transform_inverse */
  const float *root_xform;        /* [16] PartBundle::get_root_xform() */
  /* morph sliders driven by scalar channels (CharacterSlider -> AnimChannelScalarTable), SliderTable order */
  int32_t num_sliders;
  const int32_t *slider_has_channel;   /* [S] */
  const int32_t *slider_table_length;  /* [S]; 0 = empty table: value 0 (animChannelScalarTable.cxx get_value) */
  const int32_t *slider_table_offset;  /* [S] into tables */
  const float *slider_default;         /* [S] MovingPartScalar::_default_value */
} p3o_skeleton;
/* skinning_out: [J][16] = CharacterJoint::_skinning_matrix of every joint at `frame`. */
int p3o_pose(const p3o_skeleton *skel, int frame, float *skinning_out);
/* With PartBundle::set_frame_blend_flag(true): the values at get_frame() and get_next_frame() blended by
 * get_frac() (MovingPartMatrix::get_blend_value with one control of effect 1; chan/movingPartMatrix.cxx:78-198).
 * blend_type: 0 BT_linear, 1 BT_normalized_linear (the PRC default, anim-blend-type).  frame_blend = 0 is p3o_pose. */
int p3o_pose_blend(const p3o_skeleton *skel, int frame, int next_frame, float frac, int frame_blend, int blend_type,
                   float *skinning_out);

/* Several AnimControls at once (PartBundle::set_anim_blend_flag + set_control_effect): control k poses the
 * animation whose channel tables ctrl_skel[k] carries at frames[k] (/ next_frames[k], fracs[k]) with effects[k];
 * order = PartBundle's _blend map order. */
int p3o_pose_controls(const p3o_skeleton *const *ctrl_skel, int K, const int *frames, const int *next_frames, const float *fracs,
                      const float *effects, int frame_blend, int blend_type, float *skinning_out);
/* ... with forced channels (freeze_joint / control_joint): forced_values[f][16] replaces joint forced_joints[f]'s value */
int p3o_pose_controls_forced(const p3o_skeleton *const *ctrl_skel, int K, const int *frames, const int *next_frames, const float *fracs,
                             const float *effects, int frame_blend, int blend_type, int num_forced, const int *forced_joints,
                             const float *forced_values, float *skinning_out);

/* The slider values of the same controls (MovingPartScalar::get_blend_value, chan/movingPartScalar.cxx:43-105;
 * CharacterSlider::update_internals): no control with a channel -> the default value; exactly one and no frame
 * blending -> table[frame % size]; else sum(v * effect [* (1 - frac) / frac]) / net_effect.  sliders_out: [S]. */
int p3o_pose_sliders(const p3o_skeleton *const *ctrl_skel, int K, const int *frames, const int *next_frames, const float *fracs,
                     const float *effects, int frame_blend, float *sliders_out);

/* Individual pieces, exported for unit tests against the reference's adjacent
 * tests (tests/linmath/test_lmatrix4.py, test_compose_matrix.py). */
void p3o_blend(const float *palette, const int32_t *xf, const float *w, int n, float *out16);
int p3o_invert4(const float *in16, float *out16);
int p3o_invert3(const float *in9, float *out9);
int p3o_decompose3(const float *in9, int cs, float *scale3, float *shear3, float *hpr3);
void p3o_compose3(float *out9, const float *scale3, const float *shear3, const float *hpr3, int cs);
/* The xform/normalize choice of do_transform_vector_column for a C_normal
 * column (geomVertexData.cxx:1811-1840).  Returns the branch taken:
 * 0 identity-scale, 1 uniform scale (decompose/compose), 2 decompose failed,
 * 3 non-uniform (inverse transpose); *normalize is set accordingly. */
int p3o_normal_xform(const float *mat16, int cs, float *xform16, int *normalize);

#ifdef __cplusplus
}
#endif
#endif
