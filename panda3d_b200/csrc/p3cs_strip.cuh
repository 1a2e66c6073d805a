// p3cs_strip.cuh -- the fused hot-path kernel (sm_100a): one launch animates a whole crowd.
//
// Work unit = a strip: consecutive record groups (runs of 256- or 512-row tiles sharing one record list) of ONE
// instance.  Per strip the CTA stages the instance's joint-matrix palette in shared memory once;
// then, per record group:
//   * phase A: one thread per DISTINCT blend of the group accumulates sum_i w_i * M_i from the
//     shared palette (TransformBlend::update_blend, panda/src/gobj/transformBlend.cxx:211-233),
//     derives the normal matrix (geomVertexData.cxx:1811-1840) and leaves a record in smem;
// and per tile of the group:
//   * the tile's rows arrive by a TMA bulk copy (cp.async.bulk global->shared, completion on an
//     mbarrier), two or three buffers so the next tile is in flight while this one is computed;
//   * phase B: one thread per row (one or two rows per thread and tile) applies the sparse morph deltas (geomVertexData.cxx:1462-1543)
//     and the blend record of the row (:1558-1751) to the staged row, in place;
//   * the finished tile leaves by a TMA bulk store (shared->global, bulk_group).
// Blend records never touch HBM, the static mesh is read through L2 (shared by every instance),
// so the compulsory HBM traffic of a crowd is the output rows plus one palette per instance.
// Global accesses of rows are whole-tile bulk copies; every layout-dependent access (arbitrary
// stride / column offsets of GeomVertexArrayFormat) happens on shared memory.
//
// Kernel variants (template MODE, see StripMode): the loader's common formats -- one output array, float32 x 3 or
// aligned float32 x 4 columns -- get phase B code specialised on the column kinds (point / normal / vector) and,
// for the usual strides, on the row layout; everything else runs the generic column loop.  AUX adds the
// selection read-back (parity hook) or the fused tight-bounds reduction, MORPH the sparse morph-delta loop.
#pragma once
#include "p3cs_device.cuh"

namespace p3cs {

// kRow24 / kRow32 / kRow48 / kRow56: the loader's usual row layouts ([vertex normal], + texcoord, + tangent
// binormal, + both).  The whole row lives in registers and moves between them and the staged tile with the
// widest shared-memory accesses that are bank-conflict free for that stride: 64-bit when the stride is an odd
// number of 8-byte units (24, 56), 128-bit when it is an odd number of 16-byte units (48), and for the
// power-of-two stride 32 two 128-bit accesses whose order alternates every four rows.
// kAlignedPN / kAlignedPNVV: the vertex-animation-align-16 layouts of Eigen builds (every animated column float32 x 4
// on a 16-byte boundary): one 128-bit access per column, compact records extended by column 3 of the blend.
enum StripMode {
  kGenericCompact = 0, kGenericFull = 1, kFastP = 2, kFastPN = 3, kFastPNVV = 4, kRow24 = 5, kRow32 = 6, kRow48 = 7, kRow56 = 8,
  kAlignedPN = 9, kAlignedPNVV = 10
};

// Row layouts of the kRow modes, in 32-bit words: row length W, access width VEC, column offsets (-1: absent).
template <int MODE> struct RowLayout { static constexpr int W = 2, VEC = 2, P = 0, N = -1, T = -1, B = -1; };
template <> struct RowLayout<kRow24> { static constexpr int W = 6, VEC = 2, P = 0, N = 3, T = -1, B = -1; };
template <> struct RowLayout<kRow32> { static constexpr int W = 8, VEC = 4, P = 0, N = 3, T = -1, B = -1; };
template <> struct RowLayout<kRow48> { static constexpr int W = 12, VEC = 4, P = 0, N = 3, T = 6, B = 9; };
template <> struct RowLayout<kRow56> { static constexpr int W = 14, VEC = 2, P = 0, N = 3, T = 8, B = 11; };

template <class L> __device__ __forceinline__ constexpr int kRowN() { return L::N >= 0 ? L::N : 0; }
template <class L> __device__ __forceinline__ constexpr int kRowT() { return L::T >= 0 ? L::T : 0; }
template <class L> __device__ __forceinline__ constexpr int kRowB() { return L::B >= 0 ? L::B : 0; }

// does the access unit [u*VEC, (u+1)*VEC) of the row hold a word some animated column writes?
template <class L>
__device__ __forceinline__ constexpr bool row_unit_dirty(int u) {
  const int lo = u * L::VEC, hi = lo + L::VEC;
  return (L::P >= 0 && L::P < hi && L::P + 3 > lo) || (L::N >= 0 && L::N < hi && L::N + 3 > lo) ||
         (L::T >= 0 && L::T < hi && L::T + 3 > lo) || (L::B >= 0 && L::B < hi && L::B + 3 > lo);
}

template <class L>
__device__ __forceinline__ void row_load(const uint8_t *rowp, int trow, float (&r)[L::W]) {
  if (L::W == 8) {  // stride 32: rows r and r+4 share their banks, so rows 4..7 of every eight start with the upper half
    const int sel = (trow >> 2) & 1;
    const float4 a = *reinterpret_cast<const float4 *>(rowp + sel * 16);
    const float4 b = *reinterpret_cast<const float4 *>(rowp + (sel ^ 1) * 16);
    const float4 lo = sel ? b : a, hi = sel ? a : b;
    r[0] = lo.x; r[1] = lo.y; r[2] = lo.z; r[3] = lo.w;
    r[4] = hi.x; r[5] = hi.y; r[6] = hi.z; r[7] = hi.w;
  } else if (L::VEC == 4) {
#pragma unroll
    for (int u = 0; u < L::W / 4; ++u) {
      const float4 v = reinterpret_cast<const float4 *>(rowp)[u];
      r[4 * u] = v.x; r[4 * u + 1] = v.y; r[4 * u + 2] = v.z; r[4 * u + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int u = 0; u < L::W / 2; ++u) {
      const float2 v = reinterpret_cast<const float2 *>(rowp)[u];
      r[2 * u] = v.x; r[2 * u + 1] = v.y;
    }
  }
}

template <class L>
__device__ __forceinline__ void row_store(uint8_t *rowp, int trow, const float (&r)[L::W]) {
  if (L::W == 8) {
    const int sel = (trow >> 2) & 1;
    const float4 lo = make_float4(r[0], r[1], r[2], r[3]), hi = make_float4(r[4], r[5], r[6], r[7]);
    *reinterpret_cast<float4 *>(rowp + sel * 16) = sel ? hi : lo;
    *reinterpret_cast<float4 *>(rowp + (sel ^ 1) * 16) = sel ? lo : hi;
  } else if (L::VEC == 4) {
#pragma unroll
    for (int u = 0; u < L::W / 4; ++u)
      if (row_unit_dirty<L>(u)) reinterpret_cast<float4 *>(rowp)[u] = make_float4(r[4 * u], r[4 * u + 1], r[4 * u + 2], r[4 * u + 3]);
  } else {
#pragma unroll
    for (int u = 0; u < L::W / 2; ++u)
      if (row_unit_dirty<L>(u)) reinterpret_cast<float2 *>(rowp)[u] = make_float2(r[2 * u], r[2 * u + 1]);
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copies (SASS: UBLKCP); sizes and addresses are multiples of 16 bytes.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Shared-memory plan: [2 mbarriers][palette rows x cols 0..2: T x 48 B][palette column 3: T x 16 B]
// [records][tile buffer 0][tile buffer 1].  The 48-byte palette stride keeps the LDS.128 of
// phase A spread over all bank groups; column 3 is only read when a palette is not affine.
struct StripSmem {
  int pal3_off, palc_off, records_off, tiles_off, total;
};

__host__ __device__ inline StripSmem strip_layout(const MeshDev &m) {
  StripSmem s;
  s.pal3_off = 32;
  s.palc_off = s.pal3_off + m.num_transforms * 48;
  s.records_off = s.palc_off + m.num_transforms * 16;
  const int rec_bytes = m.strip_rec_floats * 4;
  const int recs_bytes = m.max_group_records * rec_bytes > 256 ? m.max_group_records * rec_bytes : 256;  // >= the bounds scratch
  s.tiles_off = (s.records_off + recs_bytes + 127) & ~127;
  // staged tiles (contiguous 32-row slices), tile_bufs of them in flight
  s.total = s.tiles_off + m.tile_bufs * (kStripThreads * m.rows_per_thread / 32) * m.slice_bytes;
  return s;
}

// ------------------------------------------------------------------ phase A, compact records
// Rare normal-matrix branches, out of line so the common path stays in registers.  They write
// the 3x3 normal matrix + normalize flag (rec[12..21]) straight into the shared-memory record.

// uniform scale != 1: decompose_matrix / compose_matrix (geomVertexData.cxx:1825-1827)
static __device__ __noinline__ void normal_uniform_scale(float m00, float m01, float m02, float m10, float m11, float m12, float m20,
                                                         float m21, float m22, int cs, float *rec) {
  Mat3 up3, c3;
  set3(up3, m00, m01, m02, m10, m11, m12, m20, m21, m22);
  float scale[3], shear[3], hpr[3];
  decompose3(up3, cs, scale, shear, hpr);
  compose3_unit_scale(c3, shear, hpr, cs);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) rec[12 + i * 3 + j] = c3.m[i][j];
  rec[21] = 0.0f;
}

// non-uniform scale on a blended matrix that is not affine within 1e-6: the LU inverse
// (LMatrix4f::invert_from general branch, lmatrix4_src.I:1225-1246)
static __device__ __noinline__ void normal_general_inverse(float m00, float m01, float m02, float m03, float m10, float m11, float m12,
                                                           float m13, float m20, float m21, float m22, float m23, float m30, float m31,
                                                           float m32, float m33, float *rec) {
  Mat4 o, inv;
  o.m[0][0] = m00; o.m[0][1] = m01; o.m[0][2] = m02; o.m[0][3] = m03;
  o.m[1][0] = m10; o.m[1][1] = m11; o.m[1][2] = m12; o.m[1][3] = m13;
  o.m[2][0] = m20; o.m[2][1] = m21; o.m[2][2] = m22; o.m[2][3] = m23;
  o.m[3][0] = m30; o.m[3][1] = m31; o.m[3][2] = m32; o.m[3][3] = m33;
  invert4_general(o, inv);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) rec[12 + i * 3 + j] = inv.m[j][i];  // transpose_in_place
  rec[21] = 1.0f;
}

// acc += M_j * weight for one entry (LMatrix4f::accumulate, lmatrix4_src.I:1311-1336).
// AFFINE: every palette matrix has column 3 == (0,0,0,1) exactly, so the blended column 3 is
// (0,0,0,sum w) and needs no multiplies (sum w is accumulated in the reference's order).
template <bool AFFINE>
__device__ __forceinline__ void accumulate_entry(const float4 *pal3, const float4 *palc, uint32_t j, float weight, float (&m)[12],
                                                 float (&c3)[4]) {
  const float4 a = pal3[j * 3], b = pal3[j * 3 + 1], c = pal3[j * 3 + 2];
  m[0] += a.x * weight; m[1] += a.y * weight; m[2] += a.z * weight; m[3] += a.w * weight;
  m[4] += b.x * weight; m[5] += b.y * weight; m[6] += b.z * weight; m[7] += b.w * weight;
  m[8] += c.x * weight; m[9] += c.y * weight; m[10] += c.z * weight; m[11] += c.w * weight;
  if (AFFINE) {
    c3[3] += weight;  // 1.0f * weight == weight
  } else {
    const float4 d = palc[j];
    c3[0] += d.x * weight; c3[1] += d.y * weight; c3[2] += d.z * weight; c3[3] += d.w * weight;
  }
}

// One compact record (22 floats) of one blend, written to shared memory.  `q0`,`q1` hold the
// record's fixed block of up to four (palette index, weight) entries, count in q0.x >> 16;
// longer blends read the CSR arrays.
template <bool AFFINE>
__device__ __forceinline__ void compact_record(const float4 *pal3, const float4 *palc, uint4 q0, uint4 q1, const MeshDev &mesh,
                                               int grec, int want_normal, bool store_c3, float *rec) {
  float m[12];  // m[r*3+c]: rows 0..3 x columns 0..2
  float c3[4];
  const uint32_t count = q0.x >> 16;
  if (count == 0) {  // empty TransformBlend: identity (transformBlend.I:362-366)
#pragma unroll
    for (int i = 0; i < 12; ++i) m[i] = (i == 0 || i == 4 || i == 8) ? 1.0f : 0.0f;
    c3[0] = c3[1] = c3[2] = 0.0f;
    c3[3] = 1.0f;
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) m[i] = 0.0f;
    c3[0] = c3[1] = c3[2] = c3[3] = 0.0f;
    if (count <= 4) {
      accumulate_entry<AFFINE>(pal3, palc, q0.x & 0xffffu, __uint_as_float(q0.y), m, c3);
      if (count > 1) accumulate_entry<AFFINE>(pal3, palc, q0.z, __uint_as_float(q0.w), m, c3);
      if (count > 2) accumulate_entry<AFFINE>(pal3, palc, q1.x, __uint_as_float(q1.y), m, c3);
      if (count > 3) accumulate_entry<AFFINE>(pal3, palc, q1.z, __uint_as_float(q1.w), m, c3);
    } else {
      const int beg = __ldg(mesh.grec_entry_offsets + grec), end = __ldg(mesh.grec_entry_offsets + grec + 1);
      for (int e = beg; e < end; ++e) {
        const uint2 en = __ldg(mesh.grec_entries + e);
        accumulate_entry<AFFINE>(pal3, palc, en.x, __uint_as_float(en.y), m, c3);
      }
    }
  }
  float2 *r2 = reinterpret_cast<float2 *>(rec);
#pragma unroll
  for (int i = 0; i < 6; ++i) r2[i] = make_float2(m[2 * i], m[2 * i + 1]);
  if (store_c3) {  // aligned modes: column 3 of the blend, for the 4-component products
    r2[11] = make_float2(c3[0], c3[1]);
    r2[12] = make_float2(c3[2], c3[3]);
  }
  if (!want_normal) return;

  // the C_normal prelude of do_transform_vector_column, geomVertexData.cxx:1811-1840
  const float sq0 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
  const float sq1 = m[3] * m[3] + m[4] * m[4] + m[5] * m[5];
  const float sq2 = m[6] * m[6] + m[7] * m[7] + m[8] * m[8];
  if (thr_equal(sq0, sq1, 2.0e-3f) && thr_equal(sq0, sq2, 2.0e-3f)) {
    if (thr_equal(sq0, 1.0f, 2.0e-3f)) {  // xform = mat
      r2[6] = make_float2(m[0], m[1]);
      r2[7] = make_float2(m[2], m[3]);
      r2[8] = make_float2(m[4], m[5]);
      r2[9] = make_float2(m[6], m[7]);
      r2[10] = make_float2(m[8], 2.0f);  // flag 2: the normal matrix IS the blend matrix (4-component normals need all of it)
    } else {
      normal_uniform_scale(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], mesh.coordinate_system, rec);
    }
    return;
  }
  // non-uniform scale: xform = transpose(invert(mat)), normalize = true
  const bool affine = AFFINE ? thr_equal(c3[3], 1.0f, P3CS_NEARLY_ZERO)
                             : (thr_equal(c3[0], 0.0f, P3CS_NEARLY_ZERO) && thr_equal(c3[1], 0.0f, P3CS_NEARLY_ZERO) &&
                                thr_equal(c3[2], 0.0f, P3CS_NEARLY_ZERO) && thr_equal(c3[3], 1.0f, P3CS_NEARLY_ZERO));
  if (!affine) {
    normal_general_inverse(m[0], m[1], m[2], c3[0], m[3], m[4], m[5], c3[1], m[6], m[7], m[8], c3[2], m[9], m[10], m[11], c3[3], rec);
    return;
  }
  // LMatrix3f::invert_from (lmatrix3_src.I:917-965), transposed on the fly: N(i,j) = inv(j,i)
  float det = (m[0] * det2(m[4], m[5], m[7], m[8]) - m[1] * det2(m[3], m[5], m[6], m[8]) + m[2] * det2(m[3], m[4], m[6], m[7]));
  if (thr_zero(det, P3CS_NEARLY_ZERO * P3CS_NEARLY_ZERO)) {
    // singular: the reference leaves `xform` uninitialised here; we define it as the identity
    r2[6] = make_float2(1.0f, 0.0f);
    r2[7] = make_float2(0.0f, 0.0f);
    r2[8] = make_float2(1.0f, 0.0f);
    r2[9] = make_float2(0.0f, 0.0f);
    r2[10] = make_float2(1.0f, 1.0f);
    return;
  }
  det = 1.0f / det;
  const float i00 = det * det2(m[4], m[5], m[7], m[8]);
  const float i10 = -det * det2(m[3], m[5], m[6], m[8]);
  const float i20 = det * det2(m[3], m[4], m[6], m[7]);
  const float i01 = -det * det2(m[1], m[2], m[7], m[8]);
  const float i11 = det * det2(m[0], m[2], m[6], m[8]);
  const float i21 = -det * det2(m[0], m[1], m[6], m[7]);
  const float i02 = det * det2(m[1], m[2], m[4], m[5]);
  const float i12 = -det * det2(m[0], m[2], m[3], m[5]);
  const float i22 = det * det2(m[0], m[1], m[3], m[4]);
  r2[6] = make_float2(i00, i10);    // N row 0 = inv column 0
  r2[7] = make_float2(i20, i01);    // N(0,2), N(1,0)
  r2[8] = make_float2(i11, i21);    // N(1,1), N(1,2)
  r2[9] = make_float2(i02, i12);    // N(2,0), N(2,1)
  r2[10] = make_float2(i22, 1.0f);  // N(2,2), normalize
}

// ------------------------------------------------------------------ phase B, align-16 columns
// LMatrix4f::xform (VECTOR4_MATRIX4_PRODUCT, lmatrix4_src.I:703-725) with the blend matrix of a 26-float record:
// f[r*3+c] = M(r,c) for c < 3, f[22+r] = M(r,3).
__device__ __forceinline__ float4 aligned_xform_blend(const float4 v, const float (&f)[26]) {
  float4 r;
  r.x = v.x * f[0] + v.y * f[3] + v.z * f[6] + v.w * f[9];
  r.y = v.x * f[1] + v.y * f[4] + v.z * f[7] + v.w * f[10];
  r.z = v.x * f[2] + v.y * f[5] + v.z * f[8] + v.w * f[11];
  r.w = v.x * f[22] + v.y * f[23] + v.z * f[24] + v.w * f[25];
  return r;
}
// A C_normal column of four values (do_transform_vector_column, geomVertexData.cxx:1811-1858): normalize ->
// table_xform_normal3f on the first three values; else table_xform_vecbase4f with the 4x4 `xform`, which is the
// blend matrix itself (flag 2) or compose_matrix(unit scale, shear, hpr, zero translation) (flag 0: the 3x3 in
// f[12..20], row 3 = (0,0,0,1), column 3 = (0,0,0,1) -- the products with those constants are evaluated, not
// folded, so signed zeros and non-finite inputs come out as on the CPU).
__device__ __forceinline__ float4 aligned_normal(const float4 v, const float (&f)[26]) {
  if (f[21] == 1.0f) {
    float r0 = v.x * f[12] + v.y * f[15] + v.z * f[18];
    float r1 = v.x * f[13] + v.y * f[16] + v.z * f[19];
    float r2 = v.x * f[14] + v.y * f[17] + v.z * f[20];
    normalize3_hot(r0, r1, r2);
    return make_float4(r0, r1, r2, v.w);
  }
  if (f[21] == 2.0f) return aligned_xform_blend(v, f);
  float4 r;
  r.x = v.x * f[12] + v.y * f[15] + v.z * f[18] + v.w * 0.0f;
  r.y = v.x * f[13] + v.y * f[16] + v.z * f[19] + v.w * 0.0f;
  r.z = v.x * f[14] + v.y * f[17] + v.z * f[20] + v.w * 0.0f;
  r.w = v.x * 0.0f + v.y * 0.0f + v.z * 0.0f + v.w * 1.0f;
  return r;
}

// ------------------------------------------------------------------ phase B, specialised columns
// KIND: 1 point, 2 vector, 3 normal (float32 x 3); f = the 22-float record in registers.
template <int KIND>
__device__ __forceinline__ void fast_column(float *cell, const float (&f)[22]) {
  const float v0 = cell[0], v1 = cell[1], v2 = cell[2];
  float r0, r1, r2;
  if (KIND == 1) {  // LMatrix4f::xform_point
    r0 = v0 * f[0] + v1 * f[3] + v2 * f[6] + f[9];
    r1 = v0 * f[1] + v1 * f[4] + v2 * f[7] + f[10];
    r2 = v0 * f[2] + v1 * f[5] + v2 * f[8] + f[11];
  } else if (KIND == 2) {  // xform_vec with the blend matrix
    r0 = v0 * f[0] + v1 * f[3] + v2 * f[6];
    r1 = v0 * f[1] + v1 * f[4] + v2 * f[7];
    r2 = v0 * f[2] + v1 * f[5] + v2 * f[8];
  } else {  // xform_vec with the normal matrix (+ normalize)
    r0 = v0 * f[12] + v1 * f[15] + v2 * f[18];
    r1 = v0 * f[13] + v1 * f[16] + v2 * f[19];
    r2 = v0 * f[14] + v1 * f[17] + v2 * f[20];
    if (f[21] == 1.0f) normalize3_hot(r0, r1, r2);
  }
  cell[0] = r0;
  cell[1] = r1;
  cell[2] = r2;
}

// AUX: 0 plain, 1 also writes the per-row selection (parity hook), 2 also reduces the animated tight bounds
template <int MODE, int kRowsPerThread, int AUX, bool MORPH>
__global__ void __launch_bounds__(kStripThreads, MODE == kGenericFull ? 2 : 4)
    strip_kernel(MeshDev mesh, GroupDev grp, int first_instance, int groups_per_cta, int want_normal) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr bool SELECT = AUX == 1;
  constexpr bool BOUNDS = AUX == 2;
  constexpr int kTileRows = kStripThreads * kRowsPerThread;
  constexpr int kSlicesPerTile = kTileRows / 32;
  const int kTileBufs = kRowsPerThread == 2 ? 2 : mesh.tile_bufs;  // 512-row tiles: always 2; else the host's plan (2 or 3)
  constexpr bool FULL = MODE == kGenericFull;
  constexpr bool FAST = MODE >= kFastP;
  constexpr bool ALIGNED = MODE >= kAlignedPN;                   // float32 x 4 columns on 16-byte boundaries
  constexpr bool ROWREG = MODE >= kRow24 && MODE <= kRow56;  // the row is held in registers (RowLayout<MODE>)
  constexpr int REC = FULL ? kRecFull : ALIGNED ? kRecAligned : kRecCompact;
  const StripSmem lay = strip_layout(mesh);
  float *pal3f = reinterpret_cast<float *>(smem + lay.pal3_off);
  float *palcf = reinterpret_cast<float *>(smem + lay.palc_off);
  const float4 *pal3 = reinterpret_cast<const float4 *>(pal3f);
  const float4 *palc = reinterpret_cast<const float4 *>(palcf);
  float *recs = reinterpret_cast<float *>(smem + lay.records_off);
  uint8_t *tiles = smem + lay.tiles_off;
  const uint32_t mbar0 = smem_u32(smem);
  const uint32_t tiles_s = smem_u32(tiles);
  const int tile_bytes = kSlicesPerTile * mesh.slice_bytes;
  const int stride0 = mesh.stride[0];

  int tid;  // read once and kept in a register: the compiler otherwise re-reads SR_TID.X (long-latency S2R) per tile
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int inst = grp.instance_list != nullptr ? __ldg(grp.instance_list + first_instance + blockIdx.y) : first_instance + (int)blockIdx.y;
  // groups_per_cta > 0: the CTA owns that many consecutive record groups.  groups_per_cta < 0 (few CTAs in the
  // launch: a single small mesh): one group per strip, its tiles split over -groups_per_cta CTAs, each of which
  // computes the group's records itself -- more CTAs in flight instead of one CTA walking every tile.
  const int tile_split = groups_per_cta < 0 ? -groups_per_cta : 1;
  const int gpc = groups_per_cta < 0 ? 1 : groups_per_cta;
  const int g0 = ((int)blockIdx.x / tile_split) * gpc;
  const int g1 = min(mesh.num_groups, g0 + gpc);
  if (g0 >= g1) return;
  int t0 = __ldg(mesh.group_tile_offsets + g0);
  int t1 = __ldg(mesh.group_tile_offsets + g1);
  if (tile_split > 1) {
    const int per = (t1 - t0 + tile_split - 1) / tile_split;
    t0 += ((int)blockIdx.x % tile_split) * per;
    t1 = min(t1, t0 + per);
    if (t0 >= t1) return;
  }

  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < kTileBufs; ++k) mbar_init(mbar0 + 8 * k, 1);
    fence_proxy_async();  // make the initialised barriers visible to the async (TMA) proxy
  }
  // one elected thread drives the TMA traffic; the first tile and the table prefetches below are issued before the
  // palette is staged, so their latencies overlap it
  auto issue_tile_load = [&](int t, int buf) {
    const int row0 = t * kTileRows;
    const int rows = min(kTileRows, mesh.num_rows - row0);
    const uint32_t bar = mbar0 + 8 * buf;
    if (FAST) {  // single output array
      const uint32_t bytes = (uint32_t)((rows * stride0 + 15) & ~15);
      mbar_expect_tx(bar, bytes);
      bulk_g2s(tiles_s + buf * tile_bytes, mesh.src[0] + (size_t)row0 * stride0, bytes, bar);
      return;
    }
    uint32_t total = 0;
    for (int o = 0; o < mesh.num_outputs; ++o) total += (uint32_t)((rows * mesh.stride[o] + 15) & ~15);
    mbar_expect_tx(bar, total);
    for (int o = 0; o < mesh.num_outputs; ++o) {
      const uint32_t bytes = (uint32_t)((rows * mesh.stride[o] + 15) & ~15);
      bulk_g2s(tiles_s + buf * tile_bytes + kSlicesPerTile * mesh.slice_out_offset[o], mesh.src[o] + (size_t)row0 * mesh.stride[o], bytes, bar);
    }
  };
  if (tid == 0) issue_tile_load(t0, 0);

  const float *sliders = grp.sliders + (size_t)inst * mesh.num_sliders;
  const bool has_index = mesh.index_bytes != 0;
  int buf = 0;      // (tiles processed so far) % kTileBufs
  uint32_t ph = 0;  // mbarrier phase parity of buffer `buf`

  // software prefetch (hides the L2 latency of the static tables): the row's record slot for the
  // next tile, and this thread's entry block of the next record group
  int next_local[kRowsPerThread];
  auto fetch_local = [&](int tile_index, bool live) {  // the record slots of this thread's rows of a tile
    if (kRowsPerThread == 2) {                         // rows t and t + 256: one 32-bit load (MeshDev::row_local2)
      const uint32_t two = (live && has_index) ? __ldg(mesh.row_local2 + (size_t)tile_index * kStripThreads + tid) : 0xFFFFFFFFu;
      next_local[0] = (int)(two & 0xFFFFu);
      next_local[kRowsPerThread - 1] = (int)(two >> 16);
    } else {
      const int row = tile_index * kTileRows + tid;
      next_local[0] = (live && has_index && row < mesh.num_rows) ? (int)__ldg(mesh.row_local + row) : 0xFFFF;
    }
  };
  fetch_local(t0, true);
  int next_mb[kRowsPerThread], next_me[kRowsPerThread];  // the rows' morph entry ranges, prefetched the same way
#pragma unroll
  for (int rr = 0; rr < kRowsPerThread; ++rr) {
    const int row = t0 * kTileRows + rr * kStripThreads + tid;
    next_mb[rr] = next_me[rr] = 0;
    if (MORPH && FAST && row < mesh.num_rows) {
      next_mb[rr] = __ldg(mesh.morph_row_offsets + row);
      next_me[rr] = __ldg(mesh.morph_row_offsets + row + 1);
    }
  }
  int rbeg = __ldg(mesh.group_rec_offsets + g0);
  int rend = __ldg(mesh.group_rec_offsets + g0 + 1);
  int gt1 = __ldg(mesh.group_tile_offsets + g0 + 1);
  uint4 pq0 = make_uint4(0, 0, 0, 0), pq1 = pq0;
  if (!FULL) {  // the table is padded by kStripThreads records, so this never reads out of bounds
    pq0 = __ldg(mesh.grec_blocks + (rbeg + tid));
    pq1 = __ldg(mesh.grec_blocks_hi + (rbeg + tid));
  }

  BoundsAcc bounds;
  bounds.reset();
  const int bounds_off = BOUNDS ? kSlicesPerTile * mesh.slice_out_offset[grp.bounds_output] + grp.bounds_start : 0;
  const int bounds_stride = BOUNDS ? mesh.stride[grp.bounds_output] : 0;


  // palette: T x 64 B by coalesced 128-bit loads; rows x columns 0..2 packed to 48 B, column 3
  // aside; on the way check whether every matrix is affine (column 3 == (0,0,0,1) exactly)
  int non_affine = 0;
  {
    const float4 *gp = reinterpret_cast<const float4 *>(grp.palettes + (size_t)inst * mesh.num_transforms * 16);
    const int n4 = mesh.num_transforms * 4;
    for (int i = tid; i < n4; i += kStripThreads) {
      const float4 row = __ldg(gp + i);
      const int j = i >> 2, r = i & 3;
      float *d = pal3f + j * 12 + r * 3;
      d[0] = row.x;
      d[1] = row.y;
      d[2] = row.z;
      palcf[i] = row.w;
      non_affine |= (row.w != (r == 3 ? 1.0f : 0.0f));
    }
  }
  non_affine = __syncthreads_or(non_affine);


  int t = t0;
  for (int g = g0; g < g1; ++g) {
    // ---------------- phase A: one thread per distinct blend of this record group
    // (the tile loop's trailing barrier guarantees nobody still reads the previous records)
    const int nrec = rend - rbeg;
    for (int b = tid; b < nrec; b += kStripThreads) {
      if (FULL) {
        Mat4 acc, x;
        const int beg = __ldg(mesh.grec_entry_offsets + rbeg + b), end = __ldg(mesh.grec_entry_offsets + rbeg + b + 1);
        blend_accumulate_packed(
            [&](int j, int r) { return make_float4(pal3f[j * 12 + r * 3], pal3f[j * 12 + r * 3 + 1], pal3f[j * 12 + r * 3 + 2], palcf[j * 4 + r]); },
            mesh.grec_entries, beg, end, acc);
        bool normalize = false;
        if (want_normal) normalize = normal_xform(acc, mesh.coordinate_system, x);
        else x = acc;
        store_record<true>(recs + (size_t)b * REC, acc, x, normalize);
      } else {
        uint4 q0 = pq0, q1 = pq1;
        if (b != tid) {
          q0 = __ldg(mesh.grec_blocks + (rbeg + b));
          q1 = __ldg(mesh.grec_blocks_hi + (rbeg + b));
        }
        if (non_affine) compact_record<false>(pal3, palc, q0, q1, mesh, rbeg + b, want_normal, ALIGNED, recs + (size_t)b * REC);
        else compact_record<true>(pal3, palc, q0, q1, mesh, rbeg + b, want_normal, ALIGNED, recs + (size_t)b * REC);
      }
    }
    __syncthreads();  // every record of the group is in shared memory
    const int cur_rbeg = rbeg;
    const int cur_gt1 = min(gt1, t1);
    if (g + 1 < g1) {  // prefetch for the next group; consumed after this group's tiles
      rbeg = rend;
      rend = __ldg(mesh.group_rec_offsets + g + 2);
      gt1 = __ldg(mesh.group_tile_offsets + g + 2);
      if (!FULL) {
        pq0 = __ldg(mesh.grec_blocks + (rbeg + tid));
        pq1 = __ldg(mesh.grec_blocks_hi + (rbeg + tid));
      }
    }

    for (; t < cur_gt1; ++t) {
      if (tid == 0 && t + 1 < t1) {
        // tile t+1 lands in the buffer last read by the store of tile t+1-kTileBufs
        if (kTileBufs >= 3) bulk_wait_read_1();
        else bulk_wait_read_all();
        issue_tile_load(t + 1, buf + 1 == kTileBufs ? 0 : buf + 1);
      }

      // -------------- per-row selection (integer work, bit-exact); the next tile's slots are prefetched
      int local[kRowsPerThread];
#pragma unroll
      for (int rr = 0; rr < kRowsPerThread; ++rr) local[rr] = next_local[rr];
      fetch_local(t + 1, t + 1 < t1);
      if (SELECT) {
#pragma unroll
        for (int rr = 0; rr < kRowsPerThread; ++rr) {
          const int row = t * kTileRows + rr * kStripThreads + tid;
          if (row < mesh.num_rows && blockIdx.y == 0)
            grp.selection[row] = local[rr] == 0xFFFF ? -1 : __ldg(mesh.group_rec_blend + cur_rbeg + local[rr]);
        }
      }
      int mb[kRowsPerThread], me[kRowsPerThread];
#pragma unroll
      for (int rr = 0; rr < kRowsPerThread; ++rr) {
        mb[rr] = next_mb[rr];
        me[rr] = next_me[rr];
        if (MORPH && FAST) {
          const int nrow = (t + 1) * kTileRows + rr * kStripThreads + tid;
          next_mb[rr] = next_me[rr] = 0;
          if (t + 1 < t1 && nrow < mesh.num_rows) {
            next_mb[rr] = __ldg(mesh.morph_row_offsets + nrow);
            next_me[rr] = __ldg(mesh.morph_row_offsets + nrow + 1);
          }
        }
      }

      mbar_wait(mbar0 + 8 * buf, ph);  // the tile's rows have landed

      // -------------- phase B: kRowsPerThread rows per thread, in place on the staged tile
      uint8_t *tile = tiles + buf * tile_bytes;
#pragma unroll
      for (int rr = 0; rr < kRowsPerThread; ++rr) {
        const int trow = rr * kStripThreads + tid;  // row within the tile
        const int row = t * kTileRows + trow;
        const bool skinned = local[rr] != 0xFFFF;
        if (FAST) {
          if (MORPH && ROWREG) {
            // Row-in-registers modes: the row is loaded first, the sparse morph deltas are added to the registers
            // (reference order; 3-component branch, geomVertexData.cxx:1531-1535), then the blend record applies --
            // no shared-memory round trip per morph entry.  Rows without entries that are not skinned stay untouched.
            using L = RowLayout<MODE>;
            if (row < mesh.num_rows && (skinned || me[rr] > mb[rr])) {
              uint8_t *rowp = tile + trow * (L::W * 4);
              float v[L::W];
              row_load<L>(rowp, trow, v);
              for (int e = mb[rr]; e < me[rr]; ++e) {
                const float4 d = __ldg(mesh.morph_deltas + e);  // .w carries the key of a 3-component delta
                const uint32_t key = __float_as_uint(d.w);
                const float value = __ldg(sliders + (key & 0xffffu));
                if (value == 0.0f) continue;  // `if (slider_value != 0.0f)`, :1483
                const int c = (int)(key >> 16);  // animated columns in canonical order: point, normal, vector, vector
                constexpr int kN = L::N >= 0 ? L::N : 0, kT = L::T >= 0 ? L::T : 0, kB = L::B >= 0 ? L::B : 0;
                if (c == 0) {
                  v[L::P] = v[L::P] + d.x * value;
                  v[L::P + 1] = v[L::P + 1] + d.y * value;
                  v[L::P + 2] = v[L::P + 2] + d.z * value;
                } else if (L::N >= 0 && c == 1) {
                  v[kN] = v[kN] + d.x * value;
                  v[kN + 1] = v[kN + 1] + d.y * value;
                  v[kN + 2] = v[kN + 2] + d.z * value;
                } else if (L::T >= 0 && c == 2) {
                  v[kT] = v[kT] + d.x * value;
                  v[kT + 1] = v[kT + 1] + d.y * value;
                  v[kT + 2] = v[kT + 2] + d.z * value;
                } else if (L::B >= 0 && c == 3) {
                  v[kB] = v[kB] + d.x * value;
                  v[kB + 1] = v[kB + 1] + d.y * value;
                  v[kB + 2] = v[kB + 2] + d.z * value;
                }
              }
              if (skinned) {
                const float2 *p = reinterpret_cast<const float2 *>(recs + (size_t)local[rr] * kRecCompact);
                float f[22];
#pragma unroll
                for (int i = 0; i < 11; ++i) {
                  const float2 q = p[i];
                  f[2 * i] = q.x;
                  f[2 * i + 1] = q.y;
                }
                fast_column<1>(v + L::P, f);
                if (L::N >= 0) fast_column<3>(v + kRowN<L>(), f);
                if (L::T >= 0) fast_column<2>(v + kRowT<L>(), f);
                if (L::B >= 0) fast_column<2>(v + kRowB<L>(), f);
              }
              if (BOUNDS && bounds_off == 0 && bounds_wants(grp, row)) bounds.add(v[0], v[1], v[2]);
              row_store<L>(rowp, trow, v);
            }
          } else if (MORPH && row < mesh.num_rows) {
            // sparse morph deltas, added in place in the reference's order (3-component branch,
            // geomVertexData.cxx:1531-1535); columns are independent, so walking the row's
            // entries once is the same arithmetic as the reference's per-column passes
            uint8_t *rowp = tile + trow * stride0;
            for (int e = mb[rr]; e < me[rr]; ++e) {
              const float4 d = __ldg(mesh.morph_deltas + e);  // .w carries the key of a 3-component delta
              // four-component columns may have a real fourth delta component: their key comes from its own table
              const uint32_t key = ALIGNED ? __ldg(mesh.morph_keys + e) : __float_as_uint(d.w);
              const float value = __ldg(sliders + (key & 0xffffu));
              if (value == 0.0f) continue;  // `if (slider_value != 0.0f)`, :1483
              const DynColumn col = mesh.dyn[key >> 16];
              float *cell = reinterpret_cast<float *>(rowp + col.start);
              if (!ALIGNED) {
                cell[0] = cell[0] + d.x * value;
                cell[1] = cell[1] + d.y * value;
                cell[2] = cell[2] + d.z * value;
              } else if (col.homogeneous) {  // :1499-1507: the delta is scaled by the homogeneous coordinate
                const float sc = value * cell[3];
                cell[0] = cell[0] + d.x * sc;
                cell[1] = cell[1] + d.y * sc;
                cell[2] = cell[2] + d.z * sc;
              } else {  // :1516-1520
                cell[0] = cell[0] + d.x * value;
                cell[1] = cell[1] + d.y * value;
                cell[2] = cell[2] + d.z * value;
                cell[3] = cell[3] + d.w * value;
              }
            }
          }
          if (MORPH && ROWREG) {
            // handled above
          } else if (ALIGNED) {
            if (skinned) {
              const float2 *p = reinterpret_cast<const float2 *>(recs + (size_t)local[rr] * kRecAligned);
              float f[26];
#pragma unroll
              for (int i = 0; i < 13; ++i) {
                const float2 v = p[i];
                f[2 * i] = v.x;
                f[2 * i + 1] = v.y;
              }
              uint8_t *rowp = tile + trow * stride0;
              if (MODE == kAlignedPN && stride0 == 32) {
                // [vertex normal] rows of 32 bytes: rows r and r+4 share their banks, so rows 4..7 of every
                // eight touch the normal first (as in kRow32)
                const int sel = (trow >> 2) & 1;
                const float4 a = *reinterpret_cast<const float4 *>(rowp + sel * 16);
                const float4 b = *reinterpret_cast<const float4 *>(rowp + (sel ^ 1) * 16);
                const float4 pv = aligned_xform_blend(sel ? b : a, f);
                const float4 nv = aligned_normal(sel ? a : b, f);
                *reinterpret_cast<float4 *>(rowp + sel * 16) = sel ? nv : pv;
                *reinterpret_cast<float4 *>(rowp + (sel ^ 1) * 16) = sel ? pv : nv;
              } else {
                float4 *c0 = reinterpret_cast<float4 *>(rowp + mesh.dyn[0].start);
                float4 *c1 = reinterpret_cast<float4 *>(rowp + mesh.dyn[1].start);
                *c0 = aligned_xform_blend(*c0, f);
                *c1 = aligned_normal(*c1, f);
                if (MODE == kAlignedPNVV) {
                  float4 *c2 = reinterpret_cast<float4 *>(rowp + mesh.dyn[2].start);
                  float4 *c3p = reinterpret_cast<float4 *>(rowp + mesh.dyn[3].start);
                  *c2 = aligned_xform_blend(*c2, f);
                  *c3p = aligned_xform_blend(*c3p, f);
                }
              }
            }
          } else if (skinned) {
            const float2 *p = reinterpret_cast<const float2 *>(recs + (size_t)local[rr] * kRecCompact);
            float f[22];
#pragma unroll
            for (int i = 0; i < 11; ++i) {
              const float2 v = p[i];
              f[2 * i] = v.x;
              f[2 * i + 1] = v.y;
            }
            if (ROWREG) {
              using L = RowLayout<MODE>;
              uint8_t *rowp = tile + trow * (L::W * 4);
              float v[L::W];
              row_load<L>(rowp, trow, v);
              fast_column<1>(v + L::P, f);
              if (L::N >= 0) fast_column<3>(v + (L::N >= 0 ? L::N : 0), f);
              if (L::T >= 0) fast_column<2>(v + (L::T >= 0 ? L::T : 0), f);
              if (L::B >= 0) fast_column<2>(v + (L::B >= 0 ? L::B : 0), f);
              if (BOUNDS && bounds_off == 0 && bounds_wants(grp, row)) bounds.add(v[0], v[1], v[2]);  // the vertex column, straight from the registers
              row_store<L>(rowp, trow, v);
            } else {
              uint8_t *rowp = tile + trow * stride0;
              fast_column<1>(reinterpret_cast<float *>(rowp + mesh.dyn[0].start), f);
              if (MODE >= kFastPN) fast_column<3>(reinterpret_cast<float *>(rowp + mesh.dyn[1].start), f);
              if (MODE >= kFastPNVV) {
                fast_column<2>(reinterpret_cast<float *>(rowp + mesh.dyn[2].start), f);
                fast_column<2>(reinterpret_cast<float *>(rowp + mesh.dyn[3].start), f);
              }
            }
          }
        } else if (row < mesh.num_rows) {
          int morph_beg = 0, morph_end = 0;
          if (mesh.has_morphs) {
            morph_beg = __ldg(mesh.morph_row_offsets + row);
            morph_end = __ldg(mesh.morph_row_offsets + row + 1);
          }
          Record<FULL> rec;
          if (skinned) load_record<FULL, false>(recs, (size_t)local[rr], rec);
          for (int c = 0; c < mesh.num_dyn; ++c) {
            const DynColumn col = mesh.dyn[c];
            float *cell = reinterpret_cast<float *>(tile + kSlicesPerTile * mesh.slice_out_offset[col.output] +
                                                    (size_t)trow * mesh.stride[col.output] + col.start);
            animate_column<FULL>(mesh, col, c, cell, morph_beg, morph_end, sliders, skinned, rec);
          }
        }
      }
      if (BOUNDS) {  // the animated vertex column of this thread's rows, as just written to the tile
#pragma unroll
        for (int rr = 0; rr < kRowsPerThread; ++rr) {
          const int trow = rr * kStripThreads + tid;
          const bool in_registers = ROWREG && bounds_off == 0 && (local[rr] != 0xFFFF || (MORPH && me[rr] > mb[rr]));
          if (t * kTileRows + trow < mesh.num_rows && !in_registers && bounds_wants(grp, t * kTileRows + trow)) {
            const float *c = reinterpret_cast<const float *>(tile + bounds_off + (size_t)trow * bounds_stride);
            bounds.add(c[0], c[1], c[2]);
          }
        }
      }
      fence_proxy_async();  // generic-proxy writes to the tile -> visible to the TMA store
      __syncthreads();

      if (tid == 0) {
        const int row0 = t * kTileRows;
        const int rows = min(kTileRows, mesh.num_rows - row0);
        const uint32_t tile_s = tiles_s + buf * tile_bytes;
        if (FAST) {
          const uint32_t bytes = (uint32_t)((rows * stride0 + 15) & ~15);
          bulk_s2g(output_base(grp, 0, inst) + (size_t)row0 * stride0, tile_s, bytes);
        } else {
          for (int o = 0; o < mesh.num_outputs; ++o) {
            const uint32_t bytes = (uint32_t)((rows * mesh.stride[o] + 15) & ~15);
            bulk_s2g(output_base(grp, o, inst) + (size_t)row0 * mesh.stride[o], tile_s + kSlicesPerTile * mesh.slice_out_offset[o], bytes);
          }
        }
        bulk_commit();
      }
      if (++buf == kTileBufs) {
        buf = 0;
        ph ^= 1;
      }
    }
  }
  if (BOUNDS)  // the records area is free after the last tile's barrier
    bounds_commit(bounds, recs, grp.bounds + (size_t)inst * 8, gridDim.x == 1, tid, kStripThreads);
  if (tid == 0) bulk_wait_read_all();  // shared memory must outlive the last bulk store's reads
}

template <int MODE, int R, bool MORPH>
static void launch_mode_rm(const MeshDev &mesh, const GroupDev &grp, int first, dim3 grid, size_t smem, int gpc, int want_normal,
                           cudaStream_t stream) {
  if (grp.selection != nullptr) {
    cudaFuncSetAttribute(strip_kernel<MODE, R, 1, MORPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    strip_kernel<MODE, R, 1, MORPH><<<grid, kStripThreads, smem, stream>>>(mesh, grp, first, gpc, want_normal);
  } else if (grp.bounds != nullptr) {
    cudaFuncSetAttribute(strip_kernel<MODE, R, 2, MORPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    strip_kernel<MODE, R, 2, MORPH><<<grid, kStripThreads, smem, stream>>>(mesh, grp, first, gpc, want_normal);
  } else {
    static thread_local int configured_device = -1;  // the attribute is sticky per function and device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != configured_device) {
      cudaFuncSetAttribute(strip_kernel<MODE, R, 0, MORPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      configured_device = dev;
    }
    strip_kernel<MODE, R, 0, MORPH><<<grid, kStripThreads, smem, stream>>>(mesh, grp, first, gpc, want_normal);
  }
}

// The specialised modes carry the morph loop only when the mesh has morphs (the generic modes
// test MeshDev::has_morphs at run time).
template <int MODE, int R>
static void launch_mode_r(const MeshDev &mesh, const GroupDev &grp, int first, dim3 grid, size_t smem, int gpc, int want_normal,
                          cudaStream_t stream) {
  if (MODE >= kFastP && mesh.has_morphs) launch_mode_rm<MODE, R, (MODE >= kFastP)>(mesh, grp, first, grid, smem, gpc, want_normal, stream);
  else launch_mode_rm<MODE, R, false>(mesh, grp, first, grid, smem, gpc, want_normal, stream);
}

// Defined here, instantiated once per mode in p3cs_strip_inst_*.cu (so the modes compile in parallel).
template <int MODE>
void launch_mode(const MeshDev &mesh, const GroupDev &grp, int first, dim3 grid, size_t smem, int gpc, int want_normal,
                        cudaStream_t stream) {
  if (MODE != kGenericFull && mesh.rows_per_thread == 2) launch_mode_r<MODE, 2>(mesh, grp, first, grid, smem, gpc, want_normal, stream);
  else launch_mode_r<MODE, 1>(mesh, grp, first, grid, smem, gpc, want_normal, stream);
}

}  // namespace p3cs
