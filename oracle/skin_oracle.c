/*
 * skin_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see skin_oracle.h).
 *
 * Plain-C restatement of Panda3D 1.11.0's CPU vertex animation.  Every
 * function cites the reference lines it follows; arithmetic is written in the
 * reference's exact operation order (scalar, non-Eigen branches) and this file
 * must be compiled without FMA contraction or reassociation
 * (-O2 -ffp-contract=off, no -ffast-math) so that results are bit-identical
 * to the reference built with its default flags on x86-64
 * (-O3 -ffast-math -fno-unsafe-math-optimizations, no FMA in the baseline ISA;
 * dtool/CompilerFlags.cmake:17-18,143).
 *
 * Paths are relative to /root/reference (panda3d/panda3d v1.11.0).
 */
#include "skin_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* dtool/src/dtoolbase/nearly_zero.h:27-76 */
#define NEARLY_ZERO_F 1.0e-6f
#define IS_THRESHOLD_ZERO(v, t) ((v) < (t) && (v) > -(t))
#define IS_THRESHOLD_EQUAL(a, b, t) (IS_THRESHOLD_ZERO((a) - (b), t))
#define IS_NEARLY_ZERO(v) IS_THRESHOLD_ZERO(v, NEARLY_ZERO_F)
#define IS_NEARLY_EQUAL(a, b) IS_THRESHOLD_EQUAL(a, b, NEARLY_ZERO_F)

/* panda/src/linmath/mathNumbers.cxx:24-25 (double expression rounded to float) */
static const float RAD_2_DEG_F = (float)(180.0 / 3.14159265358979323846);
static const float DEG_2_RAD_F = (float)(3.14159265358979323846 / 180.0);

#define M4(m, r, c) ((m)[(r) * 4 + (c)])
#define M3(m, r, c) ((m)[(r) * 3 + (c)])

/* TransformBlend::recompute_result, panda/src/gobj/transformBlend.cxx:226-231:
 * result = zeros_mat; for each entry in stored order: accumulate(M_i, w_i).
 * LMatrix4f::accumulate, panda/src/linmath/lmatrix4_src.I:1311-1336. */
void p3o_blend(const float *palette, const int32_t *xf, const float *w, int n, float *out16) {
  int i, k;
  if (n == 0) {
    /* An empty TransformBlend is never "modified" (seq stays UpdateSeq::initial,
     * transformBlend.cxx:217-222), so its result keeps the CData constructor's
     * value, the identity (transformBlend.I:362-366). */
    for (k = 0; k < 16; ++k) out16[k] = (k % 5 == 0) ? 1.0f : 0.0f;
    return;
  }
  for (k = 0; k < 16; ++k) out16[k] = 0.0f;
  for (i = 0; i < n; ++i) {
    const float *m = palette + (size_t)xf[i] * 16;
    float weight = w[i];
    for (k = 0; k < 16; ++k) {
      out16[k] += m[k] * weight;
    }
  }
}

/* LMatrix3f::invert_from, panda/src/linmath/lmatrix3_src.I:917-965
 * (DET2 :885, MATRIX3_DETERMINANT :886-889). */
#define DET2(e00, e01, e10, e11) ((e00) * (e11) - (e10) * (e01))
int p3o_invert3(const float *o, float *r) {
  float det = (M3(o, 0, 0) * DET2(M3(o, 1, 1), M3(o, 1, 2), M3(o, 2, 1), M3(o, 2, 2))
             - M3(o, 0, 1) * DET2(M3(o, 1, 0), M3(o, 1, 2), M3(o, 2, 0), M3(o, 2, 2))
             + M3(o, 0, 2) * DET2(M3(o, 1, 0), M3(o, 1, 1), M3(o, 2, 0), M3(o, 2, 1)));
  if (IS_THRESHOLD_ZERO(det, (NEARLY_ZERO_F * NEARLY_ZERO_F))) {
    static const float ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(r, ident, sizeof(ident));
    return 0;
  }
  det = 1.0f / det;
  M3(r, 0, 0) = det * DET2(M3(o, 1, 1), M3(o, 1, 2), M3(o, 2, 1), M3(o, 2, 2));
  M3(r, 1, 0) = -det * DET2(M3(o, 1, 0), M3(o, 1, 2), M3(o, 2, 0), M3(o, 2, 2));
  M3(r, 2, 0) = det * DET2(M3(o, 1, 0), M3(o, 1, 1), M3(o, 2, 0), M3(o, 2, 1));

  M3(r, 0, 1) = -det * DET2(M3(o, 0, 1), M3(o, 0, 2), M3(o, 2, 1), M3(o, 2, 2));
  M3(r, 1, 1) = det * DET2(M3(o, 0, 0), M3(o, 0, 2), M3(o, 2, 0), M3(o, 2, 2));
  M3(r, 2, 1) = -det * DET2(M3(o, 0, 0), M3(o, 0, 1), M3(o, 2, 0), M3(o, 2, 1));

  M3(r, 0, 2) = det * DET2(M3(o, 0, 1), M3(o, 0, 2), M3(o, 1, 1), M3(o, 1, 2));
  M3(r, 1, 2) = -det * DET2(M3(o, 0, 0), M3(o, 0, 2), M3(o, 1, 0), M3(o, 1, 2));
  M3(r, 2, 2) = det * DET2(M3(o, 0, 0), M3(o, 0, 1), M3(o, 1, 0), M3(o, 1, 1));
  return 1;
}

/* LMatrix4f::decompose_mat, panda/src/linmath/lmatrix4_src.cxx:380-450 */
static int lu_decompose4(float *a, int index[4]) {
  int i, j, k;
  float vv[4];
  for (i = 0; i < 4; i++) {
    float big = 0.0f;
    for (j = 0; j < 4; j++) {
      float temp = fabsf(M4(a, i, j));
      if (temp > big) big = temp;
    }
    if (IS_THRESHOLD_ZERO(big, (NEARLY_ZERO_F * NEARLY_ZERO_F))) return 0;
    vv[i] = 1.0f / big;
  }
  for (j = 0; j < 4; j++) {
    float big;
    int imax = -1;
    for (i = 0; i < j; i++) {
      float sum = M4(a, i, j);
      for (k = 0; k < i; k++) sum -= M4(a, i, k) * M4(a, k, j);
      M4(a, i, j) = sum;
    }
    big = 0.0f;
    for (i = j; i < 4; i++) {
      float sum = M4(a, i, j), dum;
      for (k = 0; k < j; k++) sum -= M4(a, i, k) * M4(a, k, j);
      M4(a, i, j) = sum;
      dum = vv[i] * fabsf(sum);
      if (dum >= big) {
        big = dum;
        imax = i;
      }
    }
    if (imax < 0) return 0;
    if (j != imax) {
      for (k = 0; k < 4; k++) {
        float dum = M4(a, imax, k);
        M4(a, imax, k) = M4(a, j, k);
        M4(a, j, k) = dum;
      }
      vv[imax] = vv[j];
    }
    index[j] = imax;
    if (M4(a, j, j) == 0.0f) M4(a, j, j) = NEARLY_ZERO_F;
    if (j != 4 - 1) {
      float dum = 1.0f / M4(a, j, j);
      for (i = j + 1; i < 4; i++) M4(a, i, j) *= dum;
    }
  }
  return 1;
}

/* LMatrix4f::back_sub_mat, panda/src/linmath/lmatrix4_src.cxx:455-483 */
static void lu_back_sub4(const float *a, const int index[4], float *inv, int row) {
  int ii = -1, i, j;
  for (i = 0; i < 4; i++) {
    int ip = index[i];
    float sum = M4(inv, row, ip);
    M4(inv, row, ip) = M4(inv, row, i);
    if (ii >= 0) {
      for (j = ii; j <= i - 1; j++) sum -= M4(a, i, j) * M4(inv, row, j);
    } else if (sum) {
      ii = i;
    }
    M4(inv, row, i) = sum;
  }
  for (i = 4 - 1; i >= 0; i--) {
    float sum = M4(inv, row, i);
    for (j = i + 1; j < 4; j++) sum -= M4(a, i, j) * M4(inv, row, j);
    M4(inv, row, i) = sum / M4(a, i, i);
  }
}

static void ident4(float *m) {
  memset(m, 0, 16 * sizeof(float));
  M4(m, 0, 0) = M4(m, 1, 1) = M4(m, 2, 2) = M4(m, 3, 3) = 1.0f;
}

/* LMatrix4f::invert_from (non-Eigen), panda/src/linmath/lmatrix4_src.I:1194-1247;
 * invert_affine_from :1263-1295.  Note: on a singular affine matrix the
 * reference returns false and leaves `this` untouched (uninitialised `xform`
 * at geomVertexData.cxx:1811); the oracle writes identity there and flags it. */
int p3o_invert4(const float *o, float *r) {
  if (IS_NEARLY_EQUAL(M4(o, 0, 3), 0.0f) && IS_NEARLY_EQUAL(M4(o, 1, 3), 0.0f) &&
      IS_NEARLY_EQUAL(M4(o, 2, 3), 0.0f) && IS_NEARLY_EQUAL(M4(o, 3, 3), 1.0f)) {
    float up[9], rot[9];
    int i, j;
    for (i = 0; i < 3; ++i)
      for (j = 0; j < 3; ++j) M3(up, i, j) = M4(o, i, j);
    if (!p3o_invert3(up, rot)) {
      ident4(r);
      return 0;
    }
    for (i = 0; i < 3; ++i)
      for (j = 0; j < 3; ++j) M4(r, i, j) = M3(rot, i, j);
    M4(r, 0, 3) = 0.0f;
    M4(r, 1, 3) = 0.0f;
    M4(r, 2, 3) = 0.0f;
    M4(r, 3, 3) = 1.0f;
    M4(r, 3, 0) = -(M4(o, 3, 0) * M4(r, 0, 0) + M4(o, 3, 1) * M4(r, 1, 0) + M4(o, 3, 2) * M4(r, 2, 0));
    M4(r, 3, 1) = -(M4(o, 3, 0) * M4(r, 0, 1) + M4(o, 3, 1) * M4(r, 1, 1) + M4(o, 3, 2) * M4(r, 2, 1));
    M4(r, 3, 2) = -(M4(o, 3, 0) * M4(r, 0, 2) + M4(o, 3, 1) * M4(r, 1, 2) + M4(o, 3, 2) * M4(r, 2, 2));
    return 1;
  } else {
    float a[16], inv[16];
    int index[4], row, i, j;
    memcpy(a, o, sizeof(a));
    if (!lu_decompose4(a, index)) {
      ident4(r);
      return 0;
    }
    ident4(inv);
    for (row = 0; row < 4; row++) lu_back_sub4(a, index, inv, row);
    /* transpose_from(inv) */
    for (i = 0; i < 4; ++i)
      for (j = 0; j < 4; ++j) M4(r, i, j) = M4(inv, j, i);
    return 1;
  }
}

/* LVecBase2f::normalize, panda/src/linmath/lvecBase2_src.I:274-286;
 * operator /= multiplies by the reciprocal (:557-566). */
static int normalize2(float *v) {
  float l2 = v[0] * v[0] + v[1] * v[1];
  if (l2 == 0.0f) {
    v[0] = v[1] = 0.0f;
    return 0;
  } else if (!IS_THRESHOLD_EQUAL(l2, 1.0f, NEARLY_ZERO_F * NEARLY_ZERO_F)) {
    float recip = 1.0f / sqrtf(l2);
    v[0] *= recip;
    v[1] *= recip;
  }
  return 1;
}

/* LVecBase3f::normalize, panda/src/linmath/lvecBase3_src.I:347-361 (+ /= :708-718) */
static int normalize3(float *v) {
  float l2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  if (l2 == 0.0f) {
    v[0] = v[1] = v[2] = 0.0f;
    return 0;
  } else if (!IS_THRESHOLD_EQUAL(l2, 1.0f, (NEARLY_ZERO_F * NEARLY_ZERO_F))) {
    float recip = 1.0f / sqrtf(l2);
    v[0] *= recip;
    v[1] *= recip;
    v[2] *= recip;
  }
  return 1;
}

/* LMatrix3f::xform, VECTOR3_MATRIX3_PRODUCT, panda/src/linmath/lmatrix3_src.I:496-515 */
static void xform3(const float *m, float *v) {
  float r0 = v[0] * M3(m, 0, 0) + v[1] * M3(m, 1, 0) + v[2] * M3(m, 2, 0);
  float r1 = v[0] * M3(m, 0, 1) + v[1] * M3(m, 1, 1) + v[2] * M3(m, 2, 1);
  float r2 = v[0] * M3(m, 0, 2) + v[1] * M3(m, 1, 2) + v[2] * M3(m, 2, 2);
  v[0] = r0;
  v[1] = r1;
  v[2] = r2;
}

/* MATRIX3_PRODUCT, panda/src/linmath/lmatrix3_src.I:662-671; res = a * b */
static void mul3(float *res, const float *a, const float *b) {
  int i, j;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j)
      M3(res, i, j) = M3(a, i, 0) * M3(b, 0, j) + M3(a, i, 1) * M3(b, 1, j) + M3(a, i, 2) * M3(b, 2, j);
}

/* unwind_zup_rotation, panda/src/linmath/compose_matrix_src.cxx:422-504 */
static void unwind_zup_rotation(float *mat, float *hpr) {
  float x[3], y[3], z[3], heading = 0, pitch = 0, roll = 0, p[2], rot[9];
  memcpy(x, mat + 0, sizeof(x));
  memcpy(y, mat + 3, sizeof(y));
  memcpy(z, mat + 6, sizeof(z));

  p[0] = y[0];
  p[1] = y[1];
  if (normalize2(p)) {
    heading = -atan2f(p[0], p[1]);
    M3(rot, 0, 0) = p[1];  M3(rot, 0, 1) = p[0];  M3(rot, 0, 2) = 0;
    M3(rot, 1, 0) = -p[0]; M3(rot, 1, 1) = p[1];  M3(rot, 1, 2) = 0;
    M3(rot, 2, 0) = 0;     M3(rot, 2, 1) = 0;     M3(rot, 2, 2) = 1;
    xform3(rot, x);
    xform3(rot, y);
    xform3(rot, z);
  }
  p[0] = y[1];
  p[1] = y[2];
  if (normalize2(p)) {
    pitch = atan2f(p[1], p[0]);
    M3(rot, 0, 0) = 1; M3(rot, 0, 1) = 0;     M3(rot, 0, 2) = 0;
    M3(rot, 1, 0) = 0; M3(rot, 1, 1) = p[0];  M3(rot, 1, 2) = -p[1];
    M3(rot, 2, 0) = 0; M3(rot, 2, 1) = p[1];  M3(rot, 2, 2) = p[0];
    xform3(rot, x);
    xform3(rot, y);
    xform3(rot, z);
  }
  p[0] = x[0];
  p[1] = x[2];
  if (normalize2(p)) {
    roll = -atan2f(p[1], p[0]);
    M3(rot, 0, 0) = p[0];  M3(rot, 0, 1) = 0; M3(rot, 0, 2) = -p[1];
    M3(rot, 1, 0) = 0;     M3(rot, 1, 1) = 1; M3(rot, 1, 2) = 0;
    M3(rot, 2, 0) = p[1];  M3(rot, 2, 1) = 0; M3(rot, 2, 2) = p[0];
    xform3(rot, x);
    xform3(rot, y);
    xform3(rot, z);
  }
  memcpy(mat + 0, x, sizeof(x));
  memcpy(mat + 3, y, sizeof(y));
  memcpy(mat + 6, z, sizeof(z));
  hpr[0] = heading * RAD_2_DEG_F;
  hpr[1] = pitch * RAD_2_DEG_F;
  hpr[2] = roll * RAD_2_DEG_F;
}

/* unwind_yup_rotation, panda/src/linmath/compose_matrix_src.cxx:315-412 */
static void unwind_yup_rotation(float *mat, float *hpr) {
  float x[3], y[3], z[3], heading = 0, pitch = 0, roll = 0, p[2], rot[9];
  memcpy(x, mat + 0, sizeof(x));
  memcpy(y, mat + 3, sizeof(y));
  memcpy(z, mat + 6, sizeof(z));

  p[0] = z[0];
  p[1] = z[2];
  if (normalize2(p)) {
    heading = atan2f(p[0], p[1]);
    M3(rot, 0, 0) = p[1];  M3(rot, 0, 1) = 0; M3(rot, 0, 2) = p[0];
    M3(rot, 1, 0) = 0;     M3(rot, 1, 1) = 1; M3(rot, 1, 2) = 0;
    M3(rot, 2, 0) = -p[0]; M3(rot, 2, 1) = 0; M3(rot, 2, 2) = p[1];
    xform3(rot, x);
    xform3(rot, y);
    xform3(rot, z);
  }
  p[0] = z[1];
  p[1] = z[2];
  if (normalize2(p)) {
    pitch = -atan2f(p[0], p[1]);
    M3(rot, 0, 0) = 1; M3(rot, 0, 1) = 0;     M3(rot, 0, 2) = 0;
    M3(rot, 1, 0) = 0; M3(rot, 1, 1) = p[1];  M3(rot, 1, 2) = p[0];
    M3(rot, 2, 0) = 0; M3(rot, 2, 1) = -p[0]; M3(rot, 2, 2) = p[1];
    xform3(rot, x);
    xform3(rot, y);
    xform3(rot, z);
  }
  p[0] = x[0];
  p[1] = x[1];
  if (normalize2(p)) {
    roll = -atan2f(p[1], p[0]);
    M3(rot, 0, 0) = p[0];  M3(rot, 0, 1) = -p[1]; M3(rot, 0, 2) = 0;
    M3(rot, 1, 0) = p[1];  M3(rot, 1, 1) = p[0];  M3(rot, 1, 2) = 0;
    M3(rot, 2, 0) = 0;     M3(rot, 2, 1) = 0;     M3(rot, 2, 2) = 1;
    xform3(rot, x);
    xform3(rot, y);
    xform3(rot, z);
  }
  memcpy(mat + 0, x, sizeof(x));
  memcpy(mat + 3, y, sizeof(y));
  memcpy(mat + 6, z, sizeof(z));
  hpr[0] = heading * RAD_2_DEG_F;
  hpr[1] = pitch * RAD_2_DEG_F;
  hpr[2] = roll * RAD_2_DEG_F;
}

/* decompose_matrix(LMatrix3f, scale, shear, hpr, cs),
 * panda/src/linmath/compose_matrix_src.cxx:525-627 */
int p3o_decompose3(const float *in9, int cs, float *scale, float *shear, float *hpr) {
  float m[9];
  memcpy(m, in9, sizeof(m));
  switch (cs) {
  case P3O_CS_ZUP_RIGHT:
    unwind_zup_rotation(m, hpr);
    break;
  case P3O_CS_YUP_RIGHT:
    unwind_yup_rotation(m, hpr);
    break;
  case P3O_CS_ZUP_LEFT:
    M3(m, 0, 2) = -M3(m, 0, 2);
    M3(m, 1, 2) = -M3(m, 1, 2);
    M3(m, 2, 0) = -M3(m, 2, 0);
    M3(m, 2, 1) = -M3(m, 2, 1);
    unwind_zup_rotation(m, hpr);
    hpr[0] = -hpr[0];
    hpr[2] = -hpr[2];
    break;
  case P3O_CS_YUP_LEFT:
    M3(m, 0, 2) = -M3(m, 0, 2);
    M3(m, 1, 2) = -M3(m, 1, 2);
    M3(m, 2, 0) = -M3(m, 2, 0);
    M3(m, 2, 1) = -M3(m, 2, 1);
    unwind_yup_rotation(m, hpr);
    break;
  default:
    return 0;
  }
  scale[0] = M3(m, 0, 0);
  scale[1] = M3(m, 1, 1);
  scale[2] = M3(m, 2, 2);
  if (scale[0] != 0.0f) {
    M3(m, 0, 1) /= scale[0];
    M3(m, 0, 2) /= scale[0];
  }
  if (scale[1] != 0.0f) {
    M3(m, 1, 0) /= scale[1];
    M3(m, 1, 2) /= scale[1];
  }
  if (scale[2] != 0.0f) {
    M3(m, 2, 0) /= scale[2];
    M3(m, 2, 1) /= scale[2];
  }
  shear[0] = M3(m, 0, 1) + M3(m, 1, 0);
  shear[1] = M3(m, 2, 0) + M3(m, 0, 2);
  shear[2] = M3(m, 2, 1) + M3(m, 1, 2);
  return 1;
}

/* LMatrix3f::set_rotate_mat_normaxis, panda/src/linmath/lmatrix3_src.cxx:258-298 */
static void rotate_mat_normaxis3(float *m, float angle, const float *axis, int cs) {
  float a0 = axis[0], a1 = axis[1], a2 = axis[2];
  float angle_rad, s, c, t, t0, t1, t2, s0, s1, s2;
  if (cs == P3O_CS_ZUP_LEFT || cs == P3O_CS_YUP_LEFT) angle = -angle;
  angle_rad = angle * DEG_2_RAD_F;
  s = sinf(angle_rad);   /* csincos -> sincosf, dtool/src/dtoolbase/cmath.I:55-75 */
  c = cosf(angle_rad);
  t = 1.0f - c;
  t0 = t * a0;
  t1 = t * a1;
  t2 = t * a2;
  s0 = s * a0;
  s1 = s * a1;
  s2 = s * a2;
  M3(m, 0, 0) = t0 * a0 + c;
  M3(m, 0, 1) = t0 * a1 + s2;
  M3(m, 0, 2) = t0 * a2 - s1;
  M3(m, 1, 0) = t1 * a0 - s2;
  M3(m, 1, 1) = t1 * a1 + c;
  M3(m, 1, 2) = t1 * a2 + s0;
  M3(m, 2, 0) = t2 * a0 + s1;
  M3(m, 2, 1) = t2 * a1 - s0;
  M3(m, 2, 2) = t2 * a2 + c;
}

/* compose_matrix(LMatrix3f &, scale, shear, hpr, cs),
 * panda/src/linmath/compose_matrix_src.cxx:283-306;
 * LMatrix3f::set_scale_shear_mat, lmatrix3_src.cxx:50-96;
 * LVector3f::forward/right/up, lvector3_src.I:267-322. */
void p3o_compose3(float *mat, const float *scale, const float *shear, const float *hpr, int cs) {
  float fwd[3] = {0, 0, 0}, right[3] = {1, 0, 0}, up[3] = {0, 0, 0}, r[9], tmp[9];
  switch (cs) {
  case P3O_CS_ZUP_RIGHT:
    M3(mat, 0, 0) = scale[0]; M3(mat, 0, 1) = shear[0] * scale[0]; M3(mat, 0, 2) = 0.0f;
    M3(mat, 1, 0) = 0.0f; M3(mat, 1, 1) = scale[1]; M3(mat, 1, 2) = 0.0f;
    M3(mat, 2, 0) = shear[1] * scale[2]; M3(mat, 2, 1) = shear[2] * scale[2]; M3(mat, 2, 2) = scale[2];
    fwd[1] = 1; up[2] = 1;
    break;
  case P3O_CS_ZUP_LEFT:
    M3(mat, 0, 0) = scale[0]; M3(mat, 0, 1) = shear[0] * scale[0]; M3(mat, 0, 2) = 0.0f;
    M3(mat, 1, 0) = 0.0f; M3(mat, 1, 1) = scale[1]; M3(mat, 1, 2) = 0.0f;
    M3(mat, 2, 0) = -shear[1] * scale[2]; M3(mat, 2, 1) = -shear[2] * scale[2]; M3(mat, 2, 2) = scale[2];
    fwd[1] = -1; up[2] = 1;
    break;
  case P3O_CS_YUP_RIGHT:
    M3(mat, 0, 0) = scale[0]; M3(mat, 0, 1) = 0.0f; M3(mat, 0, 2) = shear[1] * scale[0];
    M3(mat, 1, 0) = shear[0] * scale[1]; M3(mat, 1, 1) = scale[1]; M3(mat, 1, 2) = shear[2] * scale[1];
    M3(mat, 2, 0) = 0.0f; M3(mat, 2, 1) = 0.0f; M3(mat, 2, 2) = scale[2];
    fwd[2] = -1; up[1] = 1;
    break;
  case P3O_CS_YUP_LEFT:
  default:
    M3(mat, 0, 0) = scale[0]; M3(mat, 0, 1) = 0.0f; M3(mat, 0, 2) = -shear[1] * scale[0];
    M3(mat, 1, 0) = shear[0] * scale[1]; M3(mat, 1, 1) = scale[1]; M3(mat, 1, 2) = -shear[2] * scale[1];
    M3(mat, 2, 0) = 0.0f; M3(mat, 2, 1) = 0.0f; M3(mat, 2, 2) = scale[2];
    fwd[2] = 1; up[1] = 1;
    break;
  }
  if (!IS_NEARLY_ZERO(hpr[2])) {
    rotate_mat_normaxis3(r, hpr[2], fwd, cs);
    memcpy(tmp, mat, sizeof(tmp));
    mul3(mat, tmp, r);   /* mat *= r, lmatrix3_src.I:775-783 */
  }
  if (!IS_NEARLY_ZERO(hpr[1])) {
    rotate_mat_normaxis3(r, hpr[1], right, cs);
    memcpy(tmp, mat, sizeof(tmp));
    mul3(mat, tmp, r);
  }
  if (!IS_NEARLY_ZERO(hpr[0])) {
    rotate_mat_normaxis3(r, hpr[0], up, cs);
    memcpy(tmp, mat, sizeof(tmp));
    mul3(mat, tmp, r);
  }
}

/* The C_normal prelude of GeomVertexData::do_transform_vector_column,
 * panda/src/gobj/geomVertexData.cxx:1811-1840. */
int p3o_normal_xform(const float *mat, int cs, float *xform, int *normalize) {
  float sq0 = M4(mat, 0, 0) * M4(mat, 0, 0) + M4(mat, 0, 1) * M4(mat, 0, 1) + M4(mat, 0, 2) * M4(mat, 0, 2);
  float sq1 = M4(mat, 1, 0) * M4(mat, 1, 0) + M4(mat, 1, 1) * M4(mat, 1, 1) + M4(mat, 1, 2) * M4(mat, 1, 2);
  float sq2 = M4(mat, 2, 0) * M4(mat, 2, 0) + M4(mat, 2, 1) * M4(mat, 2, 1) + M4(mat, 2, 2) * M4(mat, 2, 2);
  *normalize = 0;
  if (IS_THRESHOLD_EQUAL(sq0, sq1, 2.0e-3f) && IS_THRESHOLD_EQUAL(sq0, sq2, 2.0e-3f)) {
    float up3[9], scale[3], shear[3], hpr[3], ones[3] = {1, 1, 1}, c3[9];
    int i, j;
    if (IS_THRESHOLD_EQUAL(sq0, 1, 2.0e-3f)) {
      memcpy(xform, mat, 16 * sizeof(float));
      return 0;
    }
    for (i = 0; i < 3; ++i)
      for (j = 0; j < 3; ++j) M3(up3, i, j) = M4(mat, i, j);
    if (p3o_decompose3(up3, cs, scale, shear, hpr)) {
      /* compose_matrix(LMatrix4f&, ...) = LMatrix4f(upper3, translate=0),
       * panda/src/linmath/compose_matrix_src.I:18-28 */
      p3o_compose3(c3, ones, shear, hpr, cs);
      ident4(xform);
      for (i = 0; i < 3; ++i)
        for (j = 0; j < 3; ++j) M4(xform, i, j) = M3(c3, i, j);
      return 1;
    }
    memcpy(xform, mat, 16 * sizeof(float)); /* unreachable for a valid cs */
    *normalize = 1;
    return 2;
  } else {
    float inv[16];
    int i, j;
    p3o_invert4(mat, inv);
    for (i = 0; i < 4; ++i)
      for (j = 0; j < 4; ++j) M4(xform, i, j) = M4(inv, j, i); /* transpose_in_place */
    *normalize = 1;
    return 3;
  }
}

/* ---- column transforms: geomVertexData.cxx:1886-1956 + lmatrix4_src.I:704-784 ---- */

static void xform_point3(float *v, const float *m) { /* table_xform_point3f / xform_point */
  float r0 = v[0] * M4(m, 0, 0) + v[1] * M4(m, 1, 0) + v[2] * M4(m, 2, 0) + M4(m, 3, 0);
  float r1 = v[0] * M4(m, 0, 1) + v[1] * M4(m, 1, 1) + v[2] * M4(m, 2, 1) + M4(m, 3, 1);
  float r2 = v[0] * M4(m, 0, 2) + v[1] * M4(m, 1, 2) + v[2] * M4(m, 2, 2) + M4(m, 3, 2);
  v[0] = r0;
  v[1] = r1;
  v[2] = r2;
}

static void xform_vec3(float *v, const float *m) { /* table_xform_vector3f / xform_vec */
  float r0 = v[0] * M4(m, 0, 0) + v[1] * M4(m, 1, 0) + v[2] * M4(m, 2, 0);
  float r1 = v[0] * M4(m, 0, 1) + v[1] * M4(m, 1, 1) + v[2] * M4(m, 2, 1);
  float r2 = v[0] * M4(m, 0, 2) + v[1] * M4(m, 1, 2) + v[2] * M4(m, 2, 2);
  v[0] = r0;
  v[1] = r1;
  v[2] = r2;
}

static void xform_vec4(float *v, const float *m) { /* table_xform_vecbase4f / xform */
  float r[4];
  int j;
  for (j = 0; j < 4; ++j)
    r[j] = v[0] * M4(m, 0, j) + v[1] * M4(m, 1, j) + v[2] * M4(m, 2, j) + v[3] * M4(m, 3, j);
  memcpy(v, r, sizeof(r));
}

/* GeomVertexReader::get_data1i on an integer column */
static int read_index(const uint8_t *p, int numeric_type) {
  switch (numeric_type) {
  case P3O_NT_UINT8: return *p;
  case P3O_NT_UINT16: { uint16_t v; memcpy(&v, p, 2); return v; }
  case P3O_NT_UINT32: { uint32_t v; memcpy(&v, p, 4); return (int)v; }
  default: return -1;
  }
}

/* ------------------------------------------------------------------------------------------------
 * GeomVertexReader::get_data3/4 and GeomVertexWriter::set_data3/4 (PN_stdfloat = float) on a column of
 * any numeric type: the GeomVertexColumn::Packer family, panda/src/gobj/geomVertexColumn.cxx.
 *
 * make_packer (:240-372) picks Packer_point for C_point / C_clip_point / C_texcoord, a colour packer for
 * C_color and the plain Packer otherwise.  Packer_point differs from Packer in get_data4f on fewer than
 * four values (w = 1, :2814-2828), in get_data3f on four values (divide by w, :2800-2808) and in the
 * mirrored setters (:3109-3140) -- calls this path never makes: columns of four values are read and
 * written with get_data4 / set_data4 and narrower ones with get_data3 / set_data3
 * (geomVertexData.cxx:1489-1537, 1782-1799), so both reduce to the plain Packer here.
 * ------------------------------------------------------------------------------------------------ */
enum { PK_PLAIN, PK_COLOR, PK_COLOR_CLAMPED };

static int packer_kind(const p3o_column *c) {
  if (c->contents != P3O_C_COLOR) return PK_PLAIN;
  if (c->num_values == 4 && (c->numeric_type == P3O_NT_UINT8 || c->numeric_type == P3O_NT_PACKED_DABC))
    return PK_COLOR_CLAMPED;  /* Packer_rgba_uint8_4 / Packer_argb_packed, :304-309 */
  return PK_COLOR;            /* Packer_color (and its float32 variants, which store the values as they are) */
}

/* Bytes a column occupies in its row (GeomVertexColumn::setup, :154-234: the packed types hold all
 * their values in one 32-bit component). */
static int column_bytes(const p3o_column *c) {
  switch (c->numeric_type) {
  case P3O_NT_UINT8: case P3O_NT_INT8: return c->num_values;
  case P3O_NT_UINT16: case P3O_NT_INT16: return 2 * c->num_values;
  case P3O_NT_FLOAT64: return 8 * c->num_values;
  case P3O_NT_PACKED_DCBA: case P3O_NT_PACKED_DABC: case P3O_NT_PACKED_UFLOAT: return 4;
  default: return 4 * c->num_values;
  }
}

/* Can the reference read and write this column at all?  (The combinations its packers answer with
 * nassert(false) are refused.) */
static int column_supported(const p3o_column *c) {
  int nt = c->numeric_type, nv = c->num_values;
  if (nv < 1 || nv > 4 || nt < 0 || nt > P3O_NT_PACKED_UFLOAT || nt == P3O_NT_STDFLOAT) return 0;
  if ((nt == P3O_NT_PACKED_DCBA || nt == P3O_NT_PACKED_DABC) && nv != 4) return 0;
  if (nt == P3O_NT_PACKED_UFLOAT && nv != 3) return 0;
  if (packer_kind(c) != PK_PLAIN) {
    if (nv < 3) return 0;                                                                  /* :325-328 */
    if (nt == P3O_NT_INT8 || nt == P3O_NT_INT16 || nt == P3O_NT_INT32) return 0;           /* :3500-3505 */
  }
  return 1;
}

/* `(unsigned int)x` and `(int)x` of a float as the reference's x86-64 build executes them: cvttss2si
 * into a 64-bit register of which the low half is kept, and cvttss2si into a 32-bit register
 * ("integer indefinite" 0x80..0 when the truncated value does not fit).  Spelled out because the plain
 * C casts are undefined outside the target type's range. */
static uint32_t cast_uint(float x) {
  if (!(fabsf(x) < 9223372036854775808.0f)) return 0u;
  return (uint32_t)(uint64_t)(int64_t)x;
}
static int32_t cast_int(float x) {
  if (!(x >= -2147483648.0f && x < 2147483648.0f)) return INT32_MIN;
  return (int32_t)x;
}

/* GeomVertexData::unpack_ufloat_a/b/c (geomVertexData.I:431-505): an unsigned float of `mbits`
 * mantissa bits and 5 exponent bits (bias 15), given right-aligned in `field`. */
static float ufloat_field_to_float(uint32_t field, int mbits) {
  uint32_t exponent = field >> mbits, mantissa = field & ((1u << mbits) - 1u), bits;
  float f;
  if (exponent == 0) return ldexpf((float)mantissa / (float)(1u << mbits), -14);  /* denormal, zero */
  bits = field << (23 - mbits);
  if (exponent == 31) bits |= 0x7f800000u; else bits += 0x38000000u;             /* inf / nan; rebias 15 -> 127 */
  memcpy(&f, &bits, 4);
  return f;
}

/* GeomVertexData::pack_ufloat (geomVertexData.I:375-426): r11 g11 b10, per component: nan / +inf keep
 * their top mantissa bits, values from 65536 up clamp to the largest finite one, normals are re-biased
 * and truncated, 2^-21 .. 2^-14 become denormals, everything smaller or negative becomes 0.  The
 * constants are the reference's; note the denormal branch keeps one mantissa bit fewer than the field
 * holds, and a nan in the third component spills one bit into the second field. */
static uint32_t pack_ufloat(float a, float b, float c) {
  int32_t ia, ib, ic;
  uint32_t word = 0;
  memcpy(&ia, &a, 4); memcpy(&ib, &b, 4); memcpy(&ic, &c, 4);
#define P3O_NAN_OR_INF(i) ((((i) & 0x7f800000) == 0x7f800000) && (uint32_t)(i) != 0xff800000u)
  if (P3O_NAN_OR_INF(ia)) word |= (uint32_t)(ia >> 17) & 0x7ffu;
  else if (ia >= 0x47800000) word |= 0x7bfu;
  else if (ia >= 0x38800000) word |= (uint32_t)((ia >> 17) - 0x1c00);
  else if (ia >= 0x35000000) word |= (((uint32_t)ia & 0x7c0000u) | 0x800000u) >> (130 - (ia >> 23));

  if (P3O_NAN_OR_INF(ib)) word |= (uint32_t)(ib >> 6) & 0x3ff800u;
  else if (ib >= 0x47800000) word |= 0x3df800u;
  else if (ib >= 0x38800000) word |= (uint32_t)((ib >> 6) - 0xe00000) & 0x3ff800u;
  else if (ib >= 0x35000000) word |= ((((uint32_t)ib & 0x7c0000u) | 0x800000u) >> (119 - (ib >> 23))) & 0x1f800u;

  if (P3O_NAN_OR_INF(ic)) word |= ((uint32_t)ic & 0x0ffe0000u) << 4;
  else if (ic >= 0x47800000) word |= 0xf7c00000u;
  else if (ic >= 0x38800000) word |= ((uint32_t)(ic - 0x38000000) << 4) & 0xffc00000u;
  else if (ic >= 0x35000000) word |= (((((uint32_t)ic << 3) & 0x03c00000u) | 0x04000000u) >> (112 - (ic >> 23))) & 0x07c00000u;
#undef P3O_NAN_OR_INF
  return word;
}

/* The column's values in the order the packers hand them out, as floats, no scaling
 * (the switch bodies of Packer::get_data1f..4f, :457-830).  Returns how many. */
static int raw_get(const p3o_column *c, const uint8_t *p, float *v) {
  int k, n = c->num_values;
  uint32_t w;
  switch (c->numeric_type) {
  case P3O_NT_UINT8: for (k = 0; k < n; ++k) v[k] = (float)p[k]; break;
  case P3O_NT_INT8: for (k = 0; k < n; ++k) v[k] = (float)(int8_t)p[k]; break;
  case P3O_NT_UINT16: for (k = 0; k < n; ++k) { uint16_t x; memcpy(&x, p + 2 * k, 2); v[k] = (float)x; } break;
  case P3O_NT_INT16: for (k = 0; k < n; ++k) { int16_t x; memcpy(&x, p + 2 * k, 2); v[k] = (float)x; } break;
  case P3O_NT_UINT32: for (k = 0; k < n; ++k) { uint32_t x; memcpy(&x, p + 4 * k, 4); v[k] = (float)x; } break;
  case P3O_NT_INT32: for (k = 0; k < n; ++k) { int32_t x; memcpy(&x, p + 4 * k, 4); v[k] = (float)x; } break;
  case P3O_NT_FLOAT32: memcpy(v, p, (size_t)n * 4); break;
  case P3O_NT_FLOAT64: for (k = 0; k < n; ++k) { double x; memcpy(&x, p + 8 * k, 8); v[k] = (float)x; } break;
  case P3O_NT_PACKED_DCBA:   /* d c b a = the four bytes in memory order (:787-796, unpack_abcd_* geomVertexData.I:343-370) */
    memcpy(&w, p, 4);
    v[0] = (float)(w & 0xffu); v[1] = (float)((w >> 8) & 0xffu); v[2] = (float)((w >> 16) & 0xffu); v[3] = (float)(w >> 24);
    break;
  case P3O_NT_PACKED_DABC:   /* b c d a (:798-807) */
    memcpy(&w, p, 4);
    v[0] = (float)((w >> 16) & 0xffu); v[1] = (float)((w >> 8) & 0xffu); v[2] = (float)(w & 0xffu); v[3] = (float)(w >> 24);
    break;
  case P3O_NT_PACKED_UFLOAT:  /* :700-708 */
    memcpy(&w, p, 4);
    v[0] = ufloat_field_to_float(w & 0x7ffu, 6);
    v[1] = ufloat_field_to_float((w >> 11) & 0x7ffu, 6);
    v[2] = ufloat_field_to_float(w >> 22, 5);
    break;
  default: break;
  }
  return n;
}

/* ... and back: the first `n` values of v, n <= num_values (the switch bodies of Packer::set_data1f..4f,
 * :1582-1975; packed words take all their values at once). */
static void raw_set(const p3o_column *c, uint8_t *p, int n, const float *v) {
  int k;
  uint32_t w;
  switch (c->numeric_type) {
  case P3O_NT_UINT8: for (k = 0; k < n; ++k) p[k] = (uint8_t)cast_uint(v[k]); break;
  case P3O_NT_INT8: for (k = 0; k < n; ++k) p[k] = (uint8_t)cast_int(v[k]); break;
  case P3O_NT_UINT16: for (k = 0; k < n; ++k) { uint16_t x = (uint16_t)cast_uint(v[k]); memcpy(p + 2 * k, &x, 2); } break;
  case P3O_NT_INT16: for (k = 0; k < n; ++k) { uint16_t x = (uint16_t)cast_int(v[k]); memcpy(p + 2 * k, &x, 2); } break;
  case P3O_NT_UINT32: for (k = 0; k < n; ++k) { uint32_t x = cast_uint(v[k]); memcpy(p + 4 * k, &x, 4); } break;
  case P3O_NT_INT32: for (k = 0; k < n; ++k) { int32_t x = cast_int(v[k]); memcpy(p + 4 * k, &x, 4); } break;
  case P3O_NT_FLOAT32: memcpy(p, v, (size_t)n * 4); break;
  case P3O_NT_FLOAT64: for (k = 0; k < n; ++k) { double x = (double)v[k]; memcpy(p + 8 * k, &x, 8); } break;
  case P3O_NT_PACKED_DCBA:   /* pack_abcd(v3, v2, v1, v0), :1925-1927, geomVertexData.I:331-338 */
    w = ((cast_uint(v[3]) & 0xffu) << 24) | ((cast_uint(v[2]) & 0xffu) << 16) | ((cast_uint(v[1]) & 0xffu) << 8) | (cast_uint(v[0]) & 0xffu);
    memcpy(p, &w, 4);
    break;
  case P3O_NT_PACKED_DABC:   /* pack_abcd(v3, v0, v1, v2), :1929-1931 */
    w = ((cast_uint(v[3]) & 0xffu) << 24) | ((cast_uint(v[0]) & 0xffu) << 16) | ((cast_uint(v[1]) & 0xffu) << 8) | (cast_uint(v[2]) & 0xffu);
    memcpy(p, &w, 4);
    break;
  case P3O_NT_PACKED_UFLOAT:  /* :1844-1846 */
    w = pack_ufloat(v[0], v[1], v[2]);
    memcpy(p, &w, 4);
    break;
  default: break;
  }
}

/* What a colour packer divides by on reading and multiplies by on writing (Packer_color, :3397-3590,
 * :3851-4025; `_v4 /= 255.0f` multiplies by the rounded reciprocal, lvecBase4_src.I:735-745). */
static float color_scale(const p3o_column *c) {
  switch (c->numeric_type) {
  case P3O_NT_UINT8: case P3O_NT_PACKED_DCBA: case P3O_NT_PACKED_DABC: return 255.0f;
  case P3O_NT_UINT16: return 65535.0f;
  case P3O_NT_UINT32: return 4294967295.0f;
  default: return 0.0f;  /* float32, float64, packed_ufloat: stored as they are */
  }
}

static void col_get4(const p3o_column *c, const uint8_t *p, float *v) {
  int k, n, kind = packer_kind(c);
  v[0] = v[1] = v[2] = v[3] = 0.0f;               /* Packer::get_data4f pads with 0, :719-737 */
  n = raw_get(c, p, v);
  if (kind != PK_PLAIN) {
    float scale = color_scale(c);
    if (scale != 0.0f) {
      float recip = 1.0f / scale;
      for (k = 0; k < n; ++k) v[k] = v[k] * recip;
    }
    if (n == 3) v[3] = 1.0f;                      /* Packer_color::get_data4f on three values, :3530-3534 */
  }
}
static void col_get3(const p3o_column *c, const uint8_t *p, float *v) {
  float t[4];
  col_get4(c, p, t);                              /* a colour of four values is read whole and cut, :3521-3525 */
  v[0] = t[0]; v[1] = t[1]; v[2] = t[2];
}

static void col_set4(const p3o_column *c, uint8_t *p, const float *v) {
  int k, n = c->num_values, kind = packer_kind(c);
  float t[4];
  for (k = 0; k < 4; ++k) t[k] = v[k];
  if (kind == PK_COLOR_CLAMPED) {                 /* :4493-4505, :4543-4552: std::clamp(x, 0, 1) * 255 */
    for (k = 0; k < 4; ++k) {
      float x = t[k];
      x = x < 0.0f ? 0.0f : (1.0f < x ? 1.0f : x);
      t[k] = x * 255.0f;
    }
  } else if (kind == PK_COLOR) {
    float scale = color_scale(c);
    if (scale != 0.0f)
      for (k = 0; k < n; ++k) t[k] = t[k] * scale;  /* `scaled = data * 255.0f`, :3940 ff. */
  }
  raw_set(c, p, n, t);                            /* fewer than four values: the leading ones, :1863-1875 */
}
static void col_set3(const p3o_column *c, uint8_t *p, const float *v) {
  float t[4];
  t[0] = v[0]; t[1] = v[1]; t[2] = v[2];
  /* on four values: Packer::set_data3f stores w = 0 (:1848-1850), the colour packers alpha = 1 (:3933-3935) */
  t[3] = packer_kind(c) == PK_PLAIN ? 0.0f : 1.0f;
  col_set4(c, p, t);
}

/* do_transform_point_column / do_transform_vector_column for one run
 * [begin,end) sharing one blend matrix (geomVertexData.cxx:1758-1880). */
static void transform_run(const p3o_mesh *mesh, const p3o_column *col, uint8_t *out, int stride,
                          const float *mat, int begin, int end) {
  int nv = col->num_values, i;
  int fast = col->numeric_type == P3O_NT_FLOAT32 && (nv == 3 || nv == 4);   /* the table_xform_* branches */
  uint8_t *datat = out + col->start + (size_t)begin * stride;
  if (col->contents == P3O_C_POINT) {
    for (i = 0; i < end - begin; ++i) {
      if (fast) {
        float *v = (float *)(datat + (size_t)i * stride);
        if (nv == 3) xform_point3(v, mat); else xform_vec4(v, mat);
      } else {  /* GeomVertexRewriter branches :1782-1799 */
        float v[4];
        uint8_t *p = datat + (size_t)i * stride;
        if (nv == 4) {
          col_get4(col, p, v);
          xform_vec4(v, mat);
          col_set4(col, p, v);
        } else {
          col_get3(col, p, v);
          xform_point3(v, mat);
          col_set3(col, p, v);
        }
      }
    }
  } else {
    float xform[16];
    int normalize = 0;
    if (col->contents == P3O_C_NORMAL) {
      p3o_normal_xform(mat, mesh->coordinate_system, xform, &normalize);
    } else {
      memcpy(xform, mat, sizeof(xform));
    }
    for (i = 0; i < end - begin; ++i) {
      if (fast) {
        float *v = (float *)(datat + (size_t)i * stride);
        if (normalize) {          /* table_xform_normal3f, also for 4-component columns */
          xform_vec3(v, xform);
          normalize3(v);
        } else if (nv == 3) {
          xform_vec3(v, xform);
        } else {
          xform_vec4(v, xform);
        }
      } else {  /* :1862-1879: get_data3 / set_data3 whatever the number of values */
        float v[4];
        uint8_t *p = datat + (size_t)i * stride;
        col_get3(col, p, v);
        xform_vec3(v, xform);
        if (normalize) normalize3(v);
        col_set3(col, p, v);
      }
    }
  }
}

static int output_index(const p3o_mesh *mesh, int array) {
  int i, o = 0;
  for (i = 0; i < array; ++i) o += mesh->array_is_output[i] ? 1 : 0;
  return mesh->array_is_output[array] ? o : -1;
}

int p3o_animate(const p3o_mesh *mesh, const float *palette, const float *sliders,
                uint8_t *const *out_arrays, float *blends_out, int32_t *selection_out) {
  int n = mesh->num_rows, a, o, ci, mi, i, j;
  float *blends = NULL;

  /* validate */
  for (ci = 0; ci < mesh->num_skin_columns; ++ci) {
    const p3o_column *c = &mesh->skin_columns[ci];
    if (!column_supported(c) || packer_kind(c) != PK_PLAIN) return -1;
    if (c->start < 0 || c->start + column_bytes(c) > mesh->array_stride[c->array]) return -1;
    if (output_index(mesh, c->array) < 0) return -1;
  }

  /* (i) new_data->copy_from(this, true), geomVertexData.cxx:1460: whole-array
   * memcpy because the post-animated array formats are data subsets with the
   * same stride (geomVertexArrayFormat.cxx:414-419, remove_column :278-304). */
  for (a = 0, o = 0; a < mesh->num_arrays; ++a) {
    if (mesh->array_is_output[a]) {
      memcpy(out_arrays[o], mesh->array_data[a], (size_t)n * mesh->array_stride[a]);
      ++o;
    }
  }

  /* (ii) morphs, geomVertexData.cxx:1462-1543 */
  for (mi = 0; mi < mesh->num_morphs; ++mi) {
    const p3o_morph *m = &mesh->morphs[mi];
    float slider_value = sliders[m->slider];
    int bo = output_index(mesh, m->base.array);
    int bstride, dstride;
    uint8_t *bdata;
    const uint8_t *ddata;
    if (bo < 0 || !column_supported(&m->base) || !column_supported(&m->delta)) return -1;
    if (slider_value == 0.0f) continue;
    bstride = mesh->array_stride[m->base.array];
    dstride = mesh->array_stride[m->delta.array];
    bdata = out_arrays[bo] + m->base.start;
    ddata = mesh->array_data[m->delta.array] + m->delta.start;
    for (i = 0; i < m->num_row_ranges; ++i) {
      int begin = m->row_ranges[2 * i], end = m->row_ranges[2 * i + 1];
      for (j = begin; j < end; ++j) {
        uint8_t *vp = bdata + (size_t)j * bstride;
        const uint8_t *dp = ddata + (size_t)j * dstride;
        float vertex[4], d[4];
        int k;
        if (m->base.num_values == 4) {
          /* GeomVertexColumn::has_homogeneous_coord, geomVertexColumn.I:182-192 */
          int homogeneous = (m->base.contents == P3O_C_POINT || m->base.contents == P3O_C_TEXCOORD);
          col_get4(&m->base, vp, vertex);
          if (homogeneous) {
            /* :1499-1507  d(3) *= slider_value * vertex[3] */
            float s = slider_value * vertex[3];
            col_get3(&m->delta, dp, d);
            d[0] *= s; d[1] *= s; d[2] *= s;
            vertex[0] = vertex[0] + d[0];
            vertex[1] = vertex[1] + d[1];
            vertex[2] = vertex[2] + d[2];
          } else {
            /* :1516-1520  vertex + d * slider_value (4 components) */
            col_get4(&m->delta, dp, d);
            for (k = 0; k < 4; ++k) vertex[k] = vertex[k] + d[k] * slider_value;
          }
          col_set4(&m->base, vp, vertex);
        } else {
          /* :1531-1535  vertex + d * slider_value (<= 3 components) */
          col_get3(&m->base, vp, vertex);
          col_get3(&m->delta, dp, d);
          for (k = 0; k < 3; ++k) vertex[k] = vertex[k] + d[k] * slider_value;
          col_set3(&m->base, vp, vertex);
        }
      }
    }
  }

  if (selection_out) {
    for (i = 0; i < n; ++i) selection_out[i] = -1;
  }
  if (mesh->blend_column.array < 0) return 0;

  /* (iii) update_blend for every blend, geomVertexData.cxx:1550-1556 */
  blends = (float *)malloc((size_t)(mesh->num_blends > 0 ? mesh->num_blends : 1) * 16 * sizeof(float));
  if (!blends) return -2;
  for (i = 0; i < mesh->num_blends; ++i) {
    int b = mesh->blend_offsets[i], e = mesh->blend_offsets[i + 1];
    p3o_blend(palette, mesh->blend_transform + b, mesh->blend_weight + b, e - b, blends + (size_t)i * 16);
  }
  if (blends_out) memcpy(blends_out, blends, (size_t)mesh->num_blends * 16 * sizeof(float));

  /* (iv) run-length scan over the rows subranges, per column: points first,
   * then vectors (geomVertexData.cxx:1574-1750; both the ushort fast path and
   * the generic reader path visit the same runs). */
  {
    const p3o_column *bc = &mesh->blend_column;
    const uint8_t *bt = mesh->array_data[bc->array] + bc->start;
    int bstride = mesh->array_stride[bc->array];
    for (ci = 0; ci < mesh->num_skin_columns; ++ci) {
      const p3o_column *col = &mesh->skin_columns[ci];
      int oa = output_index(mesh, col->array);
      int stride = mesh->array_stride[col->array];
      for (i = 0; i < mesh->num_row_ranges; ++i) {
        int begin = mesh->row_ranges[2 * i], end = mesh->row_ranges[2 * i + 1];
        int first_vertex = begin;
        int first_bi;
        if (begin >= end) continue;
        first_bi = read_index(bt + (size_t)first_vertex * bstride, bc->numeric_type);
        while (first_vertex < end) {
          int next_vertex = first_vertex, next_bi = first_bi;
          ++next_vertex;
          while (next_vertex < end) {
            next_bi = read_index(bt + (size_t)next_vertex * bstride, bc->numeric_type);
            if (next_bi != first_bi) break;
            ++next_vertex;
          }
          if (first_bi < 0 || first_bi >= mesh->num_blends) {
            free(blends);
            return -3;
          }
          transform_run(mesh, col, out_arrays[oa], stride, blends + (size_t)first_bi * 16,
                        first_vertex, next_vertex);
          if (selection_out) {
            for (j = first_vertex; j < next_vertex; ++j) selection_out[j] = first_bi;
          }
          first_vertex = next_vertex;
          first_bi = next_bi;
        }
      }
    }
    if (selection_out && mesh->num_skin_columns == 0) {
      for (i = 0; i < mesh->num_row_ranges; ++i)
        for (j = mesh->row_ranges[2 * i]; j < mesh->row_ranges[2 * i + 1]; ++j)
          selection_out[j] = read_index(bt + (size_t)j * bstride, bc->numeric_type);
    }
  }
  free(blends);
  return 0;
}

/* ---- joint hierarchy ("next" row f1) ------------------------------------------------------- */

/* MATRIX4_PRODUCT, panda/src/linmath/lmatrix4_src.I:874-893: res = a * b */
static void mul4(float *res, const float *a, const float *b) {
  int i, j;
  for (i = 0; i < 4; ++i)
    for (j = 0; j < 4; ++j)
      M4(res, i, j) = M4(a, i, 0) * M4(b, 0, j) + M4(a, i, 1) * M4(b, 1, j) + M4(a, i, 2) * M4(b, 2, j) + M4(a, i, 3) * M4(b, 3, j);
}

/* AnimChannelMatrixXfmTable::get_value (chan/animChannelMatrixXfmTable.cxx:112-124): components[i] =
 * table[frame % size] or the default (1 for the scales, 0 otherwise; animChannelMatrixXfmTable.I:75-85),
 * then compose_matrix -> LMatrix4f(upper3, translate) (compose_matrix_src.I:18-28). */
static void channel_value(const p3o_skeleton *skel, int j, int frame, float *value) {
  float comp[12], up3[9];
  int i, r, c;
  for (i = 0; i < 12; ++i) {
    int len = skel->table_length[j * 12 + i];
    comp[i] = len == 0 ? (i < 3 ? 1.0f : 0.0f) : skel->tables[skel->table_offset[j * 12 + i] + frame % len];
  }
  p3o_compose3(up3, comp + 0, comp + 3, comp + 6, skel->coordinate_system);
  for (r = 0; r < 3; ++r) {
    for (c = 0; c < 3; ++c) M4(value, r, c) = M3(up3, r, c);
    M4(value, r, 3) = 0.0f;
  }
  M4(value, 3, 0) = comp[9];
  M4(value, 3, 1) = comp[10];
  M4(value, 3, 2) = comp[11];
  M4(value, 3, 3) = 1.0f;
}

/* One table component at a frame (AnimChannelMatrixXfmTable::get_scale / get_shear / get_value). */
static float channel_component(const p3o_skeleton *skel, int j, int i, int frame) {
  int len = skel->table_length[j * 12 + i];
  return len == 0 ? (i < 3 ? 1.0f : 0.0f) : skel->tables[skel->table_offset[j * 12 + i] + frame % len];
}

/* AnimChannelMatrixXfmTable::get_value_no_scale_shear (animChannelMatrixXfmTable.cxx:131-149) */
static void channel_value_no_scale_shear(const p3o_skeleton *skel, int j, int frame, float *value) {
  float comp[12], up3[9];
  int i, r, c;
  comp[0] = comp[1] = comp[2] = 1.0f;
  comp[3] = comp[4] = comp[5] = 0.0f;
  for (i = 6; i < 12; ++i) comp[i] = channel_component(skel, j, i, frame);
  p3o_compose3(up3, comp + 0, comp + 3, comp + 6, skel->coordinate_system);
  for (r = 0; r < 3; ++r) {
    for (c = 0; c < 3; ++c) M4(value, r, c) = M3(up3, r, c);
    M4(value, r, 3) = 0.0f;
  }
  M4(value, 3, 0) = comp[9];
  M4(value, 3, 1) = comp[10];
  M4(value, 3, 2) = comp[11];
  M4(value, 3, 3) = 1.0f;
}

/* LQuaternionf::multiply (linmath/lquaternion_src.I:78-85): a = *this, b = rhs; components (r, i, j, k) */
static void quat_mul(float *res, const float *a, const float *b) {
  float r = (b[0] * a[0]) - (b[1] * a[1]) - (b[2] * a[2]) - (b[3] * a[3]);
  float i = (b[1] * a[0]) + (b[0] * a[1]) - (b[3] * a[2]) + (b[2] * a[3]);
  float j = (b[2] * a[0]) + (b[3] * a[1]) + (b[0] * a[2]) - (b[1] * a[3]);
  float k = (b[3] * a[0]) - (b[2] * a[1]) + (b[1] * a[2]) + (b[0] * a[3]);
  res[0] = r; res[1] = i; res[2] = j; res[3] = k;
}

/* LQuaternionf::set_hpr (linmath/lquaternion_src.cxx:95-120): half-angle quaternions about up / right / forward
 * (lvector3_src.I:267-322), quat_r * quat_p * quat_h on right-handed systems, invert(quat_h * quat_p * quat_r)
 * (invert_from: the real part negated, lquaternion_src.I:474-478) on left-handed ones. */
static void quat_set_hpr(float *q, const float *hpr, int cs) {
  float fwd[3] = {0, 0, 0}, right[3] = {1, 0, 0}, up[3] = {0, 0, 0}, qh[4], qp[4], qr[4], t[4], a, s, c;
  switch (cs) {
  case P3O_CS_ZUP_RIGHT: fwd[1] = 1; up[2] = 1; break;
  case P3O_CS_ZUP_LEFT: fwd[1] = -1; up[2] = 1; break;
  case P3O_CS_YUP_RIGHT: fwd[2] = -1; up[1] = 1; break;
  default: fwd[2] = 1; up[1] = 1; break;
  }
  a = (hpr[0] * 0.5f) * DEG_2_RAD_F; s = sinf(a); c = cosf(a);
  qh[0] = c; qh[1] = up[0] * s; qh[2] = up[1] * s; qh[3] = up[2] * s;
  a = (hpr[1] * 0.5f) * DEG_2_RAD_F; s = sinf(a); c = cosf(a);
  qp[0] = c; qp[1] = right[0] * s; qp[2] = right[1] * s; qp[3] = right[2] * s;
  a = (hpr[2] * 0.5f) * DEG_2_RAD_F; s = sinf(a); c = cosf(a);
  qr[0] = c; qr[1] = fwd[0] * s; qr[2] = fwd[1] * s; qr[3] = fwd[2] * s;
  if (cs == P3O_CS_ZUP_RIGHT || cs == P3O_CS_YUP_RIGHT) {
    quat_mul(t, qr, qp);
    quat_mul(q, t, qh);
  } else {
    quat_mul(t, qh, qp);
    quat_mul(q, t, qr);
    q[0] = -q[0];
  }
}

/* LQuaternionf::extract_to_matrix(LMatrix4f &) (linmath/lquaternion_src.cxx:74-89) */
static void quat_to_matrix4(float *m, const float *q) {
  float N = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];   /* LVecBase4f::dot, left to right */
  float s = (N == 0.0f) ? 0.0f : (2.0f / N);
  float xs = q[1] * s, ys = q[2] * s, zs = q[3] * s;
  float wx = q[0] * xs, wy = q[0] * ys, wz = q[0] * zs;
  float xx = q[1] * xs, xy = q[1] * ys, xz = q[1] * zs;
  float yy = q[2] * ys, yz = q[2] * zs, zz = q[3] * zs;
  M4(m, 0, 0) = 1.0f - (yy + zz); M4(m, 0, 1) = xy + wz; M4(m, 0, 2) = xz - wy; M4(m, 0, 3) = 0.0f;
  M4(m, 1, 0) = xy - wz; M4(m, 1, 1) = 1.0f - (xx + zz); M4(m, 1, 2) = yz + wx; M4(m, 1, 3) = 0.0f;
  M4(m, 2, 0) = xz + wy; M4(m, 2, 1) = yz - wx; M4(m, 2, 2) = 1.0f - (xx + yy); M4(m, 2, 3) = 0.0f;
  M4(m, 3, 0) = 0.0f; M4(m, 3, 1) = 0.0f; M4(m, 3, 2) = 0.0f; M4(m, 3, 3) = 1.0f;
}

/* `LMatrix4f::scale_shear_mat(scale, shear) * quat` followed by set_row(3, pos): the tail of the
 * BT_componentwise_quat branch (chan/movingPartMatrix.cxx:345-355; operator * (LMatrix4f, LQuaternionf),
 * lquaternion_src.I:547-561: the full 4x4 product, then row 3 and column 3 of the left matrix put back). */
static void quat_compose_value(float *value, const float *scale, const float *shear, const float *quat, const float *pos, int cs) {
  static const float zero_hpr[3] = {0.0f, 0.0f, 0.0f};
  float s3[9], m[16], qm[16];
  int r, c;
  p3o_compose3(s3, scale, shear, zero_hpr, cs);   /* hpr == 0 skips every rotation: set_scale_shear_mat alone */
  for (r = 0; r < 3; ++r) {
    for (c = 0; c < 3; ++c) M4(m, r, c) = M3(s3, r, c);
    M4(m, r, 3) = 0.0f;
  }
  M4(m, 3, 0) = 0.0f; M4(m, 3, 1) = 0.0f; M4(m, 3, 2) = 0.0f; M4(m, 3, 3) = 1.0f;
  quat_to_matrix4(qm, quat);
  mul4(value, m, qm);
  for (c = 0; c < 4; ++c) M4(value, 3, c) = M4(m, 3, c);
  for (r = 0; r < 4; ++r) M4(value, r, 3) = M4(m, r, 3);
  M4(value, 3, 0) = pos[0];
  M4(value, 3, 1) = pos[1];
  M4(value, 3, 2) = pos[2];
}

/* blend_type: 0 BT_linear, 1 BT_normalized_linear, 2 BT_componentwise, 3 BT_componentwise_quat (PartBundle::BlendType, chan/partBundle.h; the default
 * comes from the PRC variable anim-blend-type = normalized-linear, chan/partBundle.cxx:41-42).
 *
 * K AnimControls in the order PartBundle iterates its _blend map; ctrl_skel[k] carries the channel tables of
 * control k's animation (hierarchy, default values and inverses are taken from ctrl_skel[0]); a control of
 * effect 0 is not in the map.  Per joint (MovingPartBase::determine_effective_channels, movingPartBase.cxx:303-333,
 * and MovingPartMatrix::get_blend_value, movingPartMatrix.cxx:52-198): no control with a channel -> the default
 * value; exactly one and no frame blending -> its value; else the blend. */
int p3o_pose_controls(const p3o_skeleton *const *ctrl_skel, int K, const int *frames, const int *next_frames, const float *fracs,
                      const float *effects, int frame_blend, int blend_type, float *skinning_out) {
  return p3o_pose_controls_forced(ctrl_skel, K, frames, next_frames, fracs, effects, frame_blend, blend_type, 0, NULL, NULL, skinning_out);
}

/* The same with forced channels (PartBundle::freeze_joint / control_joint): a joint with a _forced_channel takes that
 * channel's value whatever the controls say (MovingPartMatrix::get_blend_value, chan/movingPartMatrix.cxx:53-60);
 * forced_values[f] is what the channel returns (AnimChannelMatrixFixed / AnimChannelMatrixDynamic::get_value). */
int p3o_pose_controls_forced(const p3o_skeleton *const *ctrl_skel, int K, const int *frames, const int *next_frames, const float *fracs,
                             const float *effects, int frame_blend, int blend_type, int num_forced, const int *forced_joints,
                             const float *forced_values, float *skinning_out) {
  const p3o_skeleton *skel = ctrl_skel[0];
  int J = skel->num_joints, j, k, c;
  float *net = (float *)malloc((size_t)(J > 0 ? J : 1) * 16 * sizeof(float));
  if (!net) return -2;
  for (j = 0; j < J; ++j) {
    float value[16];
    int n_eff = 0, only = -1, forced = -1;
    for (c = 0; c < num_forced; ++c)
      if (forced_joints[c] == j) forced = c;
    for (c = 0; c < K; ++c)
      if (effects[c] != 0.0f && ctrl_skel[c]->has_channel[j]) { ++n_eff; only = c; }
    if (forced >= 0) {
      memcpy(value, forced_values + (size_t)forced * 16, sizeof(value));
    } else if (n_eff == 0) {
      memcpy(value, skel->default_value + (size_t)j * 16, sizeof(value));
    } else if (n_eff == 1 && !frame_blend) {
      channel_value(ctrl_skel[only], j, frames[only], value);  /* the single-value case, movingPartMatrix.cxx:71-76 */
    } else {
      /* sums start from zero, in map order; the division by net_effect multiplies by its reciprocal
       * (LMatrix4 / LVecBase3 operator /=, lmatrix4_src.I:1094-1098, lvecBase3_src.I:709-718) */
      float nv[16], scale[3] = {0.0f, 0.0f, 0.0f}, shear[3] = {0.0f, 0.0f, 0.0f}, net_effect = 0.0f, recip;
      float comps[12] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
      float quat[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (k = 0; k < 16; ++k) nv[k] = 0.0f;
      for (c = 0; c < K; ++c) {
        const p3o_skeleton *cs = ctrl_skel[c];
        float v[16], effect = effects[c];
        int pass, passes = frame_blend ? 2 : 1;
        if (effect == 0.0f || !cs->has_channel[j]) continue;
        for (pass = 0; pass < passes; ++pass) {
          int fr = pass == 0 ? frames[c] : next_frames[c];
          float e = !frame_blend ? effect : pass == 0 ? effect * (1.0f - fracs[c]) : effect * fracs[c];
          if (blend_type == 2) {  /* BT_componentwise, :200-271: scale, shear, hpr and pos blend on their own */
            for (k = 0; k < 12; ++k) comps[k] = comps[k] + channel_component(cs, j, k, fr) * e;
            continue;
          }
          if (blend_type == 3) {  /* BT_componentwise_quat, :273-357: the rotation blends as a quaternion (get_quat, animChannelMatrixXfmTable.cxx:186-197) */
            float ihpr[3], iq[4];
            for (k = 0; k < 3; ++k) ihpr[k] = channel_component(cs, j, 6 + k, fr);
            quat_set_hpr(iq, ihpr, skel->coordinate_system);
            for (k = 0; k < 6; ++k) comps[k] = comps[k] + channel_component(cs, j, k, fr) * e;
            for (k = 9; k < 12; ++k) comps[k] = comps[k] + channel_component(cs, j, k, fr) * e;
            for (k = 0; k < 4; ++k) quat[k] = quat[k] + iq[k] * e;
            continue;
          }
          if (blend_type == 0) channel_value(cs, j, fr, v);   /* BT_linear, :82-122 */
          else channel_value_no_scale_shear(cs, j, fr, v);     /* BT_normalized_linear, :125-198 */
          for (k = 0; k < 16; ++k) nv[k] = nv[k] + v[k] * e;
          if (blend_type != 0)
            for (k = 0; k < 3; ++k) {
              scale[k] = scale[k] + channel_component(cs, j, k, fr) * e;
              shear[k] = shear[k] + channel_component(cs, j, 3 + k, fr) * e;
            }
        }
        net_effect = net_effect + effect;
      }
      recip = 1.0f / net_effect;
      for (k = 0; k < 16; ++k) nv[k] = nv[k] * recip;
      if (blend_type == 3) {
        for (k = 0; k < 12; ++k) comps[k] = comps[k] * recip;
        for (k = 0; k < 4; ++k) quat[k] = quat[k] * recip;
        quat_compose_value(value, comps + 0, comps + 3, quat, comps + 9, skel->coordinate_system);
      } else if (blend_type == 2) {
        float out3[9];
        int r, cc;
        for (k = 0; k < 12; ++k) comps[k] = comps[k] * recip;
        p3o_compose3(out3, comps + 0, comps + 3, comps + 6, skel->coordinate_system);
        for (r = 0; r < 3; ++r) {
          for (cc = 0; cc < 3; ++cc) M4(value, r, cc) = M3(out3, r, cc);
          M4(value, r, 3) = 0.0f;
        }
        M4(value, 3, 0) = comps[9];
        M4(value, 3, 1) = comps[10];
        M4(value, 3, 2) = comps[11];
        M4(value, 3, 3) = 1.0f;
      } else if (blend_type == 0) {
        memcpy(value, nv, sizeof(value));
      } else {
        /* decompose_matrix(net_value, false_scale, false_shear, hpr, translate); compose_matrix(_value, scale,
         * shear, hpr, translate) (compose_matrix_src.I:18-60) */
        float false_scale[3], false_shear[3], hpr[3], up3[9], out3[9];
        int r, cc;
        for (k = 0; k < 3; ++k) {
          scale[k] = scale[k] * recip;
          shear[k] = shear[k] * recip;
        }
        for (r = 0; r < 3; ++r)
          for (cc = 0; cc < 3; ++cc) M3(up3, r, cc) = M4(nv, r, cc);
        p3o_decompose3(up3, skel->coordinate_system, false_scale, false_shear, hpr);
        p3o_compose3(out3, scale, shear, hpr, skel->coordinate_system);
        for (r = 0; r < 3; ++r) {
          for (cc = 0; cc < 3; ++cc) M4(value, r, cc) = M3(out3, r, cc);
          M4(value, r, 3) = 0.0f;
        }
        M4(value, 3, 0) = M4(nv, 3, 0);
        M4(value, 3, 1) = M4(nv, 3, 1);
        M4(value, 3, 2) = M4(nv, 3, 2);
        M4(value, 3, 3) = 1.0f;
      }
    }
    /* CharacterJoint::update_internals: _net_transform = _value * parent net (or root xform) */
    mul4(net + (size_t)j * 16, value, skel->parent[j] >= 0 ? net + (size_t)skel->parent[j] * 16 : skel->root_xform);
    /* _skinning_matrix = _initial_net_transform_inverse * _net_transform */
    mul4(skinning_out + (size_t)j * 16, skel->initial_inverse + (size_t)j * 16, net + (size_t)j * 16);
  }
  free(net);
  return 0;
}

/* AnimChannelScalarTable::get_value (chan/animChannelScalarTable.cxx:88-94) */
static float slider_channel_value(const p3o_skeleton *sk, int s, int frame) {
  int len = sk->slider_table_length[s];
  return len == 0 ? 0.0f : sk->tables[sk->slider_table_offset[s] + frame % len];
}

int p3o_pose_sliders(const p3o_skeleton *const *ctrl_skel, int K, const int *frames, const int *next_frames, const float *fracs,
                     const float *effects, int frame_blend, float *sliders_out) {
  const p3o_skeleton *skel = ctrl_skel[0];
  int S = skel->num_sliders, s, c;
  for (s = 0; s < S; ++s) {
    int n_eff = 0, only = -1;
    for (c = 0; c < K; ++c)
      if (effects[c] != 0.0f && ctrl_skel[c]->slider_has_channel[s]) { ++n_eff; only = c; }
    if (n_eff == 0) {
      sliders_out[s] = skel->slider_default[s];
    } else if (n_eff == 1 && !frame_blend) {
      sliders_out[s] = slider_channel_value(ctrl_skel[only], s, frames[only]);
    } else {  /* movingPartScalar.cxx:63-104: the sum starts from zero, `_value /= net` is a true division */
      float value = 0.0f, net = 0.0f;
      for (c = 0; c < K; ++c) {
        float effect = effects[c];
        if (effect == 0.0f || !ctrl_skel[c]->slider_has_channel[s]) continue;
        if (!frame_blend) {
          value = value + slider_channel_value(ctrl_skel[c], s, frames[c]) * effect;
        } else {
          value = value + slider_channel_value(ctrl_skel[c], s, frames[c]) * (effect * (1.0f - fracs[c]));
          value = value + slider_channel_value(ctrl_skel[c], s, next_frames[c]) * (effect * fracs[c]);
        }
        net = net + effect;
      }
      sliders_out[s] = value / net;
    }
  }
  return 0;
}

int p3o_pose_blend(const p3o_skeleton *skel, int frame, int next_frame, float frac, int frame_blend, int blend_type,
                   float *skinning_out) {
  const float effect = 1.0f;
  return p3o_pose_controls(&skel, 1, &frame, &next_frame, &frac, &effect, frame_blend, blend_type, skinning_out);
}

int p3o_pose(const p3o_skeleton *skel, int frame, float *skinning_out) {
  return p3o_pose_blend(skel, frame, frame, 0.0f, 0, 1, skinning_out);
}
