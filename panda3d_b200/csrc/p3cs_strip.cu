// p3cs_strip.cu -- the fused hot-path kernel (sm_100a): one launch animates a whole crowd.
//
// Work unit = a strip: `tiles_per_cta` consecutive 256-row tiles of ONE instance.  Per strip the
// CTA stages the instance's joint-matrix palette in shared memory once; then, per tile:
//   * the tile's rows arrive by a TMA bulk copy (cp.async.bulk global->shared, completion on an
//     mbarrier), double buffered so the next tile is in flight while this one is computed;
//   * phase A: one thread per DISTINCT blend of the tile accumulates sum_i w_i * M_i from the
//     shared palette (TransformBlend::update_blend, panda/src/gobj/transformBlend.cxx:211-233),
//     derives the normal matrix (geomVertexData.cxx:1811-1840) and leaves a record in smem;
//   * phase B: one thread per row applies the sparse morph deltas (geomVertexData.cxx:1462-1543)
//     and the blend record of the row (:1558-1751) to the staged row, in place;
//   * the finished tile leaves by a TMA bulk store (shared->global, bulk_group).
// Blend records never touch HBM, the static mesh is read through L2 (shared by every instance),
// so the compulsory HBM traffic of a crowd is the output rows plus one palette per instance.
// Global accesses of rows are whole-tile bulk copies; every layout-dependent access (arbitrary
// stride / column offsets of GeomVertexArrayFormat) happens on shared memory.
//
// Kernel variants (template MODE): the loader's common formats -- one output array, float32 x 3
// columns, no morphs -- get phase B code specialised on the column kinds (point / normal /
// vector); everything else runs the generic column loop.
#include "p3cs_device.cuh"

namespace p3cs {

constexpr int kStripThreads = kTileRows;  // one thread per row of a tile
constexpr int kPalStride4 = 5;            // palette matrix stride in smem: 5 x float4 (80 B): LDS.128 without 4-way conflicts

// kFastPN24: the egg loader's packed [vertex(3f) normal(3f)] rows (stride 24): the whole row is
// moved with three conflict-free 64-bit shared-memory accesses each way.
enum StripMode { kGenericCompact = 0, kGenericFull = 1, kFastP = 2, kFastPN = 3, kFastPNVV = 4, kFastPN24 = 5 };

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copies (SASS: UBLKCP); sizes and addresses are multiples of 16 bytes.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct StripSmem {
  int palette_off, records_off, tiles_off, total;
};

__host__ __device__ inline StripSmem strip_layout(const MeshDev &m) {
  StripSmem s;
  s.palette_off = 16;  // two mbarriers
  s.records_off = s.palette_off + m.num_transforms * kPalStride4 * 16;
  const int rec_bytes = (m.full_records ? kRecFull : kRecCompact) * 4;
  s.tiles_off = (s.records_off + (m.max_group_records > 0 ? m.max_group_records : 1) * rec_bytes + 127) & ~127;
  s.total = s.tiles_off + 2 * m.tile_bytes;
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

// One compact record (22 floats) of one blend, written to shared memory.
// AFFINE: every palette matrix has column 3 == (0,0,0,1) exactly, so the blended column 3 is
// (0,0,0,sum w) and needs no multiplies (sum w is accumulated in the reference's order).
template <bool AFFINE>
__device__ __forceinline__ void accumulate_entry(const float4 *pal, uint2 e, float (&m)[4][3], float (&c3)[4]) {
  const float weight = __uint_as_float(e.y);
  const float4 *mat = pal + e.x * kPalStride4;
#pragma unroll
  for (int r = 0; r < 4; ++r) {  // LMatrix4f::accumulate, lmatrix4_src.I:1311-1336
    const float4 row = mat[r];
    m[r][0] += row.x * weight;
    m[r][1] += row.y * weight;
    m[r][2] += row.z * weight;
    if (!AFFINE) c3[r] += row.w * weight;
  }
  if (AFFINE) c3[3] += weight;  // 1.0f * weight == weight
}

template <bool AFFINE>
__device__ __forceinline__ void compact_record(const float4 *pal, const uint2 *entries, int beg, int end, int cs, int want_normal,
                                               float *rec) {
  float m[4][3];
  float c3[4];
  if (beg == end) {  // empty TransformBlend: identity (transformBlend.I:362-366)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) m[i][j] = (i == j) ? 1.0f : 0.0f;
      c3[i] = (i == 3) ? 1.0f : 0.0f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) m[i][j] = 0.0f;
      c3[i] = 0.0f;
    }
    if (end - beg == 4) {  // the common 4-weight blend: straight-line, all loads issued up front
      const uint2 e0 = __ldg(entries + beg), e1 = __ldg(entries + beg + 1), e2 = __ldg(entries + beg + 2), e3 = __ldg(entries + beg + 3);
      accumulate_entry<AFFINE>(pal, e0, m, c3);
      accumulate_entry<AFFINE>(pal, e1, m, c3);
      accumulate_entry<AFFINE>(pal, e2, m, c3);
      accumulate_entry<AFFINE>(pal, e3, m, c3);
    } else {
      for (int e = beg; e < end; ++e) accumulate_entry<AFFINE>(pal, __ldg(entries + e), m, c3);
    }
  }
  float2 *r2 = reinterpret_cast<float2 *>(rec);
  r2[0] = make_float2(m[0][0], m[0][1]);
  r2[1] = make_float2(m[0][2], m[1][0]);
  r2[2] = make_float2(m[1][1], m[1][2]);
  r2[3] = make_float2(m[2][0], m[2][1]);
  r2[4] = make_float2(m[2][2], m[3][0]);
  r2[5] = make_float2(m[3][1], m[3][2]);
  if (!want_normal) return;

  // the C_normal prelude of do_transform_vector_column, geomVertexData.cxx:1811-1840
  const float sq0 = m[0][0] * m[0][0] + m[0][1] * m[0][1] + m[0][2] * m[0][2];
  const float sq1 = m[1][0] * m[1][0] + m[1][1] * m[1][1] + m[1][2] * m[1][2];
  const float sq2 = m[2][0] * m[2][0] + m[2][1] * m[2][1] + m[2][2] * m[2][2];
  if (thr_equal(sq0, sq1, 2.0e-3f) && thr_equal(sq0, sq2, 2.0e-3f)) {
    if (thr_equal(sq0, 1.0f, 2.0e-3f)) {  // xform = mat
      r2[6] = make_float2(m[0][0], m[0][1]);
      r2[7] = make_float2(m[0][2], m[1][0]);
      r2[8] = make_float2(m[1][1], m[1][2]);
      r2[9] = make_float2(m[2][0], m[2][1]);
      r2[10] = make_float2(m[2][2], 0.0f);
    } else {
      normal_uniform_scale(m[0][0], m[0][1], m[0][2], m[1][0], m[1][1], m[1][2], m[2][0], m[2][1], m[2][2], cs, rec);
    }
    return;
  }
  // non-uniform scale: xform = transpose(invert(mat)), normalize = true
  const bool affine = AFFINE ? thr_equal(c3[3], 1.0f, P3CS_NEARLY_ZERO)
                             : (thr_equal(c3[0], 0.0f, P3CS_NEARLY_ZERO) && thr_equal(c3[1], 0.0f, P3CS_NEARLY_ZERO) &&
                                thr_equal(c3[2], 0.0f, P3CS_NEARLY_ZERO) && thr_equal(c3[3], 1.0f, P3CS_NEARLY_ZERO));
  if (!affine) {
    normal_general_inverse(m[0][0], m[0][1], m[0][2], c3[0], m[1][0], m[1][1], m[1][2], c3[1], m[2][0], m[2][1], m[2][2], c3[2],
                           m[3][0], m[3][1], m[3][2], c3[3], rec);
    return;
  }
  // LMatrix3f::invert_from (lmatrix3_src.I:917-965), transposed on the fly: N(i,j) = inv(j,i)
  float det = (m[0][0] * det2(m[1][1], m[1][2], m[2][1], m[2][2]) - m[0][1] * det2(m[1][0], m[1][2], m[2][0], m[2][2]) +
               m[0][2] * det2(m[1][0], m[1][1], m[2][0], m[2][1]));
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
  const float i00 = det * det2(m[1][1], m[1][2], m[2][1], m[2][2]);
  const float i10 = -det * det2(m[1][0], m[1][2], m[2][0], m[2][2]);
  const float i20 = det * det2(m[1][0], m[1][1], m[2][0], m[2][1]);
  const float i01 = -det * det2(m[0][1], m[0][2], m[2][1], m[2][2]);
  const float i11 = det * det2(m[0][0], m[0][2], m[2][0], m[2][2]);
  const float i21 = -det * det2(m[0][0], m[0][1], m[2][0], m[2][1]);
  const float i02 = det * det2(m[0][1], m[0][2], m[1][1], m[1][2]);
  const float i12 = -det * det2(m[0][0], m[0][2], m[1][0], m[1][2]);
  const float i22 = det * det2(m[0][0], m[0][1], m[1][0], m[1][1]);
  r2[6] = make_float2(i00, i10);   // N row 0 = inv column 0
  r2[7] = make_float2(i20, i01);   // N(0,2), N(1,0)
  r2[8] = make_float2(i11, i21);   // N(1,1), N(1,2)
  r2[9] = make_float2(i02, i12);   // N(2,0), N(2,1)
  r2[10] = make_float2(i22, 1.0f); // N(2,2), normalize
}

// ------------------------------------------------------------------ phase B, specialised columns
// KIND: 1 point, 2 vector, 3 normal (float32 x 3), record already in registers.
template <int KIND>
__device__ __forceinline__ void fast_column(float *cell, const float (&m)[4][3], const float (&n)[3][3], bool normalize) {
  const float v0 = cell[0], v1 = cell[1], v2 = cell[2];
  float r0, r1, r2;
  if (KIND == 1) {  // LMatrix4f::xform_point
    r0 = v0 * m[0][0] + v1 * m[1][0] + v2 * m[2][0] + m[3][0];
    r1 = v0 * m[0][1] + v1 * m[1][1] + v2 * m[2][1] + m[3][1];
    r2 = v0 * m[0][2] + v1 * m[1][2] + v2 * m[2][2] + m[3][2];
  } else if (KIND == 2) {  // xform_vec with the blend matrix
    r0 = v0 * m[0][0] + v1 * m[1][0] + v2 * m[2][0];
    r1 = v0 * m[0][1] + v1 * m[1][1] + v2 * m[2][1];
    r2 = v0 * m[0][2] + v1 * m[1][2] + v2 * m[2][2];
  } else {  // xform_vec with the normal matrix (+ normalize)
    r0 = v0 * n[0][0] + v1 * n[1][0] + v2 * n[2][0];
    r1 = v0 * n[0][1] + v1 * n[1][1] + v2 * n[2][1];
    r2 = v0 * n[0][2] + v1 * n[1][2] + v2 * n[2][2];
    if (normalize) normalize3(r0, r1, r2);
  }
  cell[0] = r0;
  cell[1] = r1;
  cell[2] = r2;
}

template <int MODE>
__global__ void __launch_bounds__(kStripThreads, MODE == kGenericFull ? 2 : 4)
    strip_kernel(MeshDev mesh, GroupDev grp, int first_instance, int groups_per_cta, int want_normal) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr bool FULL = MODE == kGenericFull;
  constexpr int REC = FULL ? kRecFull : kRecCompact;
  const StripSmem lay = strip_layout(mesh);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
  float4 *pal = reinterpret_cast<float4 *>(smem + lay.palette_off);
  float *recs = reinterpret_cast<float *>(smem + lay.records_off);
  uint8_t *tiles = smem + lay.tiles_off;

  const int tid = threadIdx.x;
  const int inst = first_instance + blockIdx.y;
  const int g0 = blockIdx.x * groups_per_cta;
  const int g1 = min(mesh.num_groups, g0 + groups_per_cta);
  if (g0 >= g1) return;
  const int t0 = g0 * mesh.group_tiles;
  const int t1 = min(mesh.num_tiles, g1 * mesh.group_tiles);

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_proxy_async();  // make the initialised barriers visible to the async (TMA) proxy
  }
  // palette: T x 64 B, coalesced 128-bit loads, re-strided to 80 B in shared memory; on the way
  // check whether every matrix is affine (column 3 == (0,0,0,1) exactly)
  int non_affine = 0;
  {
    const float4 *gp = reinterpret_cast<const float4 *>(grp.palettes + (size_t)inst * mesh.num_transforms * 16);
    const int n4 = mesh.num_transforms * 4;
    for (int i = tid; i < n4; i += kStripThreads) {
      const float4 row = __ldg(gp + i);
      pal[(i >> 2) * kPalStride4 + (i & 3)] = row;
      non_affine |= (row.w != ((i & 3) == 3 ? 1.0f : 0.0f));
    }
  }
  non_affine = __syncthreads_or(non_affine);

  // one elected thread drives the TMA traffic
  auto issue_tile_load = [&](int t, int buf) {
    const int row0 = t * kTileRows;
    const int rows = min(kTileRows, mesh.num_rows - row0);
    if (MODE >= kFastP) {  // single output array
      const uint32_t bytes = (uint32_t)((rows * mesh.stride[0] + 15) & ~15);
      mbar_expect_tx(&mbar[buf], bytes);
      bulk_g2s(tiles + buf * mesh.tile_bytes, mesh.src[0] + (size_t)row0 * mesh.stride[0], bytes, &mbar[buf]);
      return;
    }
    uint32_t total = 0;
    for (int o = 0; o < mesh.num_outputs; ++o) total += (uint32_t)((rows * mesh.stride[o] + 15) & ~15);
    mbar_expect_tx(&mbar[buf], total);
    for (int o = 0; o < mesh.num_outputs; ++o) {
      const uint32_t bytes = (uint32_t)((rows * mesh.stride[o] + 15) & ~15);
      bulk_g2s(tiles + buf * mesh.tile_bytes + mesh.tile_out_offset[o], mesh.src[o] + (size_t)row0 * mesh.stride[o], bytes, &mbar[buf]);
    }
  };
  if (tid == 0) issue_tile_load(t0, 0);

  const float *sliders = grp.sliders + (size_t)inst * mesh.num_sliders;
  int it = 0;  // tiles processed by this CTA so far (buffer / mbarrier phase bookkeeping)

  for (int g = g0; g < g1; ++g) {
    // ---------------- phase A: one thread per distinct blend of this record group
    const int rbeg = __ldg(mesh.group_rec_offsets + g);
    const int nrec = __ldg(mesh.group_rec_offsets + g + 1) - rbeg;
    for (int b = tid; b < nrec; b += kStripThreads) {
      const int beg = __ldg(mesh.grec_entry_offsets + rbeg + b);
      const int end = __ldg(mesh.grec_entry_offsets + rbeg + b + 1);
      if (FULL) {
        Mat4 acc, x;
        blend_accumulate_packed([&](int j, int r) { return pal[j * kPalStride4 + r]; }, mesh.grec_entries, beg, end, acc);
        bool normalize = false;
        if (want_normal) normalize = normal_xform(acc, mesh.coordinate_system, x);
        else x = acc;
        store_record<true>(recs + (size_t)b * REC, acc, x, normalize);
      } else if (non_affine) {
        compact_record<false>(pal, mesh.grec_entries, beg, end, mesh.coordinate_system, want_normal, recs + (size_t)b * REC);
      } else {
        compact_record<true>(pal, mesh.grec_entries, beg, end, mesh.coordinate_system, want_normal, recs + (size_t)b * REC);
      }
    }
    __syncthreads();  // every record of the group is in shared memory

    const int gt0 = g * mesh.group_tiles;
    const int gt1 = min(mesh.num_tiles, gt0 + mesh.group_tiles);
    for (int t = gt0; t < gt1; ++t, ++it) {
      const int buf = it & 1;
      if (tid == 0 && t + 1 < t1) {
        bulk_wait_read_all();  // the store that last read buffer buf^1 has drained it
        issue_tile_load(t + 1, buf ^ 1);
      }

      // -------------- per-row selection (integer work, bit-exact)
      const int row = t * kTileRows + tid;
      const bool live = row < mesh.num_rows;
      int local = 0xFFFF;
      if (live && mesh.index_bytes != 0) local = __ldg(mesh.row_local + row);
      if (live && grp.selection != nullptr && blockIdx.y == 0)
        grp.selection[row] = local == 0xFFFF ? -1 : __ldg(mesh.group_rec_blend + rbeg + local);
      const bool skinned = local != 0xFFFF;

      mbar_wait(&mbar[buf], (it >> 1) & 1);  // the tile's rows have landed

      // -------------- phase B: one thread per row, in place on the staged tile
      uint8_t *tile = tiles + buf * mesh.tile_bytes;
      if (MODE >= kFastP) {
        if (live && skinned) {
          const float2 *p = reinterpret_cast<const float2 *>(recs + (size_t)local * kRecCompact);
          float f[22];
#pragma unroll
          for (int i = 0; i < 11; ++i) {
            const float2 v = p[i];
            f[2 * i] = v.x;
            f[2 * i + 1] = v.y;
          }
          float m[4][3], n[3][3];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) m[i][j] = f[i * 3 + j];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) n[i][j] = f[12 + i * 3 + j];
          const bool normalize = f[21] != 0.0f;
          if (MODE == kFastPN24) {
            float2 *rowp = reinterpret_cast<float2 *>(tile + (size_t)tid * 24);
            const float2 a = rowp[0], b2 = rowp[1], c = rowp[2];
            float v[6] = {a.x, a.y, b2.x, b2.y, c.x, c.y};
            fast_column<1>(v, m, n, normalize);
            fast_column<3>(v + 3, m, n, normalize);
            rowp[0] = make_float2(v[0], v[1]);
            rowp[1] = make_float2(v[2], v[3]);
            rowp[2] = make_float2(v[4], v[5]);
          } else {
            uint8_t *rowp = tile + (size_t)tid * mesh.stride[0];
            fast_column<1>(reinterpret_cast<float *>(rowp + mesh.dyn[0].start), m, n, normalize);
            if (MODE >= kFastPN) fast_column<3>(reinterpret_cast<float *>(rowp + mesh.dyn[1].start), m, n, normalize);
            if (MODE >= kFastPNVV) {
              fast_column<2>(reinterpret_cast<float *>(rowp + mesh.dyn[2].start), m, n, normalize);
              fast_column<2>(reinterpret_cast<float *>(rowp + mesh.dyn[3].start), m, n, normalize);
            }
          }
        }
      } else if (live) {
        int morph_beg = 0, morph_end = 0;
        if (mesh.has_morphs) {
          morph_beg = __ldg(mesh.morph_row_offsets + row);
          morph_end = __ldg(mesh.morph_row_offsets + row + 1);
        }
        Record<FULL> rec;
        if (skinned) load_record<FULL, false>(recs, (size_t)local, rec);
        for (int c = 0; c < mesh.num_dyn; ++c) {
          const DynColumn col = mesh.dyn[c];
          float *cell = reinterpret_cast<float *>(tile + mesh.tile_out_offset[col.output] + (size_t)tid * mesh.stride[col.output] + col.start);
          animate_column<FULL>(mesh, col, c, cell, morph_beg, morph_end, sliders, skinned, rec);
        }
      }
      fence_proxy_async();  // generic-proxy writes to the tile -> visible to the TMA store
      __syncthreads();

      if (tid == 0) {
        const int row0 = t * kTileRows;
        const int rows = min(kTileRows, mesh.num_rows - row0);
        if (MODE >= kFastP) {
          const uint32_t bytes = (uint32_t)((rows * mesh.stride[0] + 15) & ~15);
          bulk_s2g(grp.out[0] + (size_t)inst * grp.pitch[0] + (size_t)row0 * mesh.stride[0], tile, bytes);
        } else {
          for (int o = 0; o < mesh.num_outputs; ++o) {
            const uint32_t bytes = (uint32_t)((rows * mesh.stride[o] + 15) & ~15);
            bulk_s2g(grp.out[o] + (size_t)inst * grp.pitch[o] + (size_t)row0 * mesh.stride[o], tile + mesh.tile_out_offset[o], bytes);
          }
        }
        bulk_commit();
      }
    }
  }
  if (tid == 0) bulk_wait_read_all();  // shared memory must outlive the last bulk store's reads
}

size_t strip_smem_bytes(const MeshDev &mesh) { return (size_t)strip_layout(mesh).total; }

// Which specialisation fits this mesh (see the header comment).
static int pick_mode(const MeshDev &mesh) {
  if (mesh.full_records) return kGenericFull;
  if (mesh.num_outputs != 1 || mesh.has_morphs || mesh.index_bytes == 0) return kGenericCompact;
  for (int c = 0; c < mesh.num_dyn; ++c)
    if (mesh.dyn[c].num_values != 3) return kGenericCompact;
  const DynColumn *d = mesh.dyn;
  if (mesh.num_dyn == 1 && d[0].skin == 1) return kFastP;
  if (mesh.num_dyn == 2 && d[0].skin == 1 && d[1].skin == 3)
    return (mesh.stride[0] == 24 && d[0].start == 0 && d[1].start == 12) ? kFastPN24 : kFastPN;
  if (mesh.num_dyn == 4 && d[0].skin == 1 && d[1].skin == 3 && d[2].skin == 2 && d[3].skin == 2) return kFastPNVV;
  return kGenericCompact;
}

template <int MODE>
static void launch_mode(const MeshDev &mesh, const GroupDev &grp, int first, dim3 grid, size_t smem, int tpc, int want_normal,
                        cudaStream_t stream) {
  cudaFuncSetAttribute(strip_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  strip_kernel<MODE><<<grid, kStripThreads, smem, stream>>>(mesh, grp, first, tpc, want_normal);
}

int launch_strip(const MeshDev &mesh, const GroupDev &grp, int first_instance, int count, int sm_count, cudaStream_t stream) {
  if (mesh.num_rows == 0 || count == 0 || mesh.num_tiles == 0) return 0;
  int want_normal = 0;
  for (int c = 0; c < mesh.num_dyn; ++c) want_normal |= (mesh.dyn[c].skin == 3);
  const size_t smem = strip_smem_bytes(mesh);
  const int mode = pick_mode(mesh);
  // strips of whole record groups: enough of them to fill every SM several times over, yet as
  // long as possible so the palette is staged rarely
  const long long target_ctas = (long long)sm_count * 24;
  long long tpc = ((long long)mesh.num_groups * count + target_ctas - 1) / target_ctas;
  if (tpc < 1) tpc = 1;
  if (tpc > mesh.num_groups) tpc = mesh.num_groups;
  const int strips = (mesh.num_groups + (int)tpc - 1) / (int)tpc;
  tpc = (mesh.num_groups + strips - 1) / strips;  // equal-length strips
  int launches = 0;
  for (int c0 = 0; c0 < count; c0 += 65535) {
    const int cn = count - c0 < 65535 ? count - c0 : 65535;
    const dim3 grid(strips, cn);
    const int first = first_instance + c0;
    switch (mode) {
    case kGenericFull: launch_mode<kGenericFull>(mesh, grp, first, grid, smem, (int)tpc, want_normal, stream); break;
    case kFastP: launch_mode<kFastP>(mesh, grp, first, grid, smem, (int)tpc, want_normal, stream); break;
    case kFastPN: launch_mode<kFastPN>(mesh, grp, first, grid, smem, (int)tpc, want_normal, stream); break;
    case kFastPNVV: launch_mode<kFastPNVV>(mesh, grp, first, grid, smem, (int)tpc, want_normal, stream); break;
    case kFastPN24: launch_mode<kFastPN24>(mesh, grp, first, grid, smem, (int)tpc, want_normal, stream); break;
    default: launch_mode<kGenericCompact>(mesh, grp, first, grid, smem, (int)tpc, want_normal, stream); break;
    }
    ++launches;
  }
  return launches;
}

}  // namespace p3cs
