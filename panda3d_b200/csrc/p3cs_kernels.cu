// p3cs_kernels.cu -- hand-written sm_100a kernels of the vertex-animation hot path.
//
//   blend_kernel  (K1)  one thread per (instance, blend): TransformBlend::update_blend
//                       (panda/src/gobj/transformBlend.cxx:211-233) + the per-matrix normal
//                       prelude of do_transform_vector_column (geomVertexData.cxx:1811-1840),
//                       written as one record per blend.
//   skin_kernel   (K2)  one thread per (instance, row): the copy_from + morph + skinning
//                       loops of GeomVertexData::update_animated_vertices
//                       (geomVertexData.cxx:1460-1751).  Rows travel HBM -> smem -> HBM as
//                       whole tiles with 128-bit coalesced accesses; all layout-dependent
//                       work (arbitrary strides / column offsets) happens on shared memory.
//
// Compiled with -fmad=false: see p3cs_math.cuh.
#include "p3cs_kernels.cuh"
#include "p3cs_math.cuh"

namespace p3cs {

// ---------------------------------------------------------------------------- K1
template <bool FULL>
__global__ void __launch_bounds__(128) blend_kernel(MeshDev mesh, GroupDev grp, int first_instance, int want_normal) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= mesh.num_blends) return;
  const int inst = first_instance + blockIdx.y;
  const float4 *pal = reinterpret_cast<const float4 *>(grp.palettes + (size_t)inst * mesh.num_transforms * 16);
  const int beg = mesh.blend_offsets[b], end = mesh.blend_offsets[b + 1];

  Mat4 acc;
  if (beg == end) {
    set_ident4(acc);  // empty TransformBlend keeps CData's identity (transformBlend.I:362-366)
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc.m[i][j] = 0.0f;
    for (int e = beg; e < end; ++e) {
      const int xf = mesh.blend_transform[e];
      const float w = mesh.blend_weight[e];
      const float4 *m = pal + (size_t)xf * 4;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 row = __ldg(m + r);
        // LMatrix4f::accumulate (lmatrix4_src.I:1311-1336): acc += m * w, element by element
        acc.m[r][0] += row.x * w;
        acc.m[r][1] += row.y * w;
        acc.m[r][2] += row.z * w;
        acc.m[r][3] += row.w * w;
      }
    }
  }

  Mat4 x;
  bool normalize = false;
  if (want_normal) {
    normalize = normal_xform(acc, mesh.coordinate_system, x);
  } else {
    x = acc;
  }

  if (FULL) {
    float4 *rec = reinterpret_cast<float4 *>(grp.records + ((size_t)blockIdx.y * mesh.num_blends + b) * kRecFull);
#pragma unroll
    for (int r = 0; r < 4; ++r) rec[r] = make_float4(acc.m[r][0], acc.m[r][1], acc.m[r][2], acc.m[r][3]);
#pragma unroll
    for (int r = 0; r < 4; ++r) rec[4 + r] = make_float4(x.m[r][0], x.m[r][1], x.m[r][2], x.m[r][3]);
    rec[8] = make_float4(normalize ? 1.0f : 0.0f, 0.0f, 0.0f, 0.0f);
  } else {
    float2 *rec = reinterpret_cast<float2 *>(grp.records + ((size_t)blockIdx.y * mesh.num_blends + b) * kRecCompact);
    rec[0] = make_float2(acc.m[0][0], acc.m[0][1]);
    rec[1] = make_float2(acc.m[0][2], acc.m[1][0]);
    rec[2] = make_float2(acc.m[1][1], acc.m[1][2]);
    rec[3] = make_float2(acc.m[2][0], acc.m[2][1]);
    rec[4] = make_float2(acc.m[2][2], acc.m[3][0]);
    rec[5] = make_float2(acc.m[3][1], acc.m[3][2]);
    rec[6] = make_float2(x.m[0][0], x.m[0][1]);
    rec[7] = make_float2(x.m[0][2], x.m[1][0]);
    rec[8] = make_float2(x.m[1][1], x.m[1][2]);
    rec[9] = make_float2(x.m[2][0], x.m[2][1]);
    rec[10] = make_float2(x.m[2][2], normalize ? 1.0f : 0.0f);
  }
}

// ---------------------------------------------------------------------------- K2
constexpr int kSkinThreads = 256;

// Blend record in registers.
template <bool FULL>
struct Record {
  float m[4][4];  // blended matrix (column 3 only meaningful when FULL)
  float x[4][4];  // normal matrix (upper 3x3 when !FULL)
  bool normalize;
};

template <bool FULL>
__device__ __forceinline__ void load_record(const float *records, size_t index, Record<FULL> &r) {
  if (FULL) {
    const float4 *p = reinterpret_cast<const float4 *>(records + index * kRecFull);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v = __ldg(p + i);
      r.m[i][0] = v.x; r.m[i][1] = v.y; r.m[i][2] = v.z; r.m[i][3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v = __ldg(p + 4 + i);
      r.x[i][0] = v.x; r.x[i][1] = v.y; r.x[i][2] = v.z; r.x[i][3] = v.w;
    }
    r.normalize = __ldg(p + 8).x != 0.0f;
  } else {
    const float2 *p = reinterpret_cast<const float2 *>(records + index * kRecCompact);
    float f[22];
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      float2 v = __ldg(p + i);
      f[2 * i] = v.x;
      f[2 * i + 1] = v.y;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) r.m[i][j] = f[i * 3 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) r.x[i][j] = f[12 + i * 3 + j];
    r.normalize = f[21] != 0.0f;
  }
}

// LMatrix4f::xform_point (lmatrix4_src.I:735-753): ((v0*m00 + v1*m10) + v2*m20) + m30
__device__ __forceinline__ void xform_point3(float *v, const float m[4][4]) {
  float r0 = v[0] * m[0][0] + v[1] * m[1][0] + v[2] * m[2][0] + m[3][0];
  float r1 = v[0] * m[0][1] + v[1] * m[1][1] + v[2] * m[2][1] + m[3][1];
  float r2 = v[0] * m[0][2] + v[1] * m[1][2] + v[2] * m[2][2] + m[3][2];
  v[0] = r0; v[1] = r1; v[2] = r2;
}
// LMatrix4f::xform_vec (lmatrix4_src.I:767-784)
__device__ __forceinline__ void xform_vec3(float *v, const float m[4][4]) {
  float r0 = v[0] * m[0][0] + v[1] * m[1][0] + v[2] * m[2][0];
  float r1 = v[0] * m[0][1] + v[1] * m[1][1] + v[2] * m[2][1];
  float r2 = v[0] * m[0][2] + v[1] * m[1][2] + v[2] * m[2][2];
  v[0] = r0; v[1] = r1; v[2] = r2;
}
// LMatrix4f::xform (VECTOR4_MATRIX4_PRODUCT, lmatrix4_src.I:703-725)
__device__ __forceinline__ void xform_vec4(float *v, const float m[4][4]) {
  float r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) r[j] = v[0] * m[0][j] + v[1] * m[1][j] + v[2] * m[2][j] + v[3] * m[3][j];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = r[j];
}

// One dynamic column of one row, in shared memory: morph deltas first
// (geomVertexData.cxx:1462-1543), then the blend transform (:1758-1880).
template <bool FULL>
__device__ __forceinline__ void animate_column(const MeshDev &mesh, const DynColumn &col, int dyn_index, float *cell,
                                               int morph_beg, int morph_end, const float *sliders, bool skinned,
                                               const Record<FULL> &rec) {
  const int nv = col.num_values;
  float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nv) v[k] = cell[k];

  for (int e = morph_beg; e < morph_end; ++e) {
    const uint32_t key = __ldg(mesh.morph_keys + e);
    if ((int)(key >> 16) != dyn_index) continue;
    const float value = __ldg(sliders + (key & 0xffffu));
    if (value == 0.0f) continue;  // `if (slider_value != 0.0f)`, :1483
    const float4 d = __ldg(mesh.morph_deltas + e);
    if (nv == 4) {
      if (col.homogeneous) {  // :1499-1507
        const float s = value * v[3];
        v[0] = v[0] + d.x * s;
        v[1] = v[1] + d.y * s;
        v[2] = v[2] + d.z * s;
      } else {  // :1516-1520
        v[0] = v[0] + d.x * value;
        v[1] = v[1] + d.y * value;
        v[2] = v[2] + d.z * value;
        v[3] = v[3] + d.w * value;
      }
    } else {  // :1531-1535
      v[0] = v[0] + d.x * value;
      v[1] = v[1] + d.y * value;
      v[2] = v[2] + d.z * value;
    }
  }

  if (skinned && col.skin != 0) {
    if (col.skin == 1) {  // C_point
      if (nv == 3) xform_point3(v, rec.m);
      else if (FULL) xform_vec4(v, rec.m);
    } else if (col.skin == 2) {  // C_vector: xform = mat
      if (nv == 3) xform_vec3(v, rec.m);
      else if (FULL) xform_vec4(v, rec.m);
    } else {  // C_normal
      if (rec.normalize) {  // table_xform_normal3f, also on 4-component columns (:1854-1855)
        xform_vec3(v, rec.x);
        normalize3(v[0], v[1], v[2]);
      } else if (nv == 3) {
        xform_vec3(v, rec.x);
      } else if (FULL) {
        xform_vec4(v, rec.x);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nv) cell[k] = v[k];
}

template <bool FULL>
__global__ void __launch_bounds__(kSkinThreads) skin_kernel(MeshDev mesh, GroupDev grp, int first_instance, int rows_per_tile) {
  extern __shared__ uint4 smem_raw[];
  uint8_t *tile = reinterpret_cast<uint8_t *>(smem_raw);

  const int inst_local = blockIdx.y;
  const int inst = first_instance + inst_local;
  const int row0 = blockIdx.x * rows_per_tile;
  const int rows = min(rows_per_tile, mesh.num_rows - row0);
  const int tid = threadIdx.x;

  for (int rbase = 0; rbase < rows_per_tile; rbase += kSkinThreads) {
    // -------- per-row selection: which blend, which morph entries (bit-exact integer work)
    const int r = rbase + tid;
    const int row = row0 + r;
    const bool live = r < rows;
    bool skinned = false;
    int bi = -1;
    int morph_beg = 0, morph_end = 0;
    Record<FULL> rec;
    if (live) {
      if (mesh.index_bytes != 0) {
        skinned = mesh.rows_mask == nullptr || ((__ldg(mesh.rows_mask + (row >> 5)) >> (row & 31)) & 1u);
        if (skinned) {
          bi = mesh.index_bytes == 2 ? (int)__ldg(reinterpret_cast<const uint16_t *>(mesh.index) + row)
                                     : (int)__ldg(reinterpret_cast<const uint32_t *>(mesh.index) + row);
          load_record<FULL>(grp.records, (size_t)inst_local * mesh.num_blends + bi, rec);
        }
      }
      if (mesh.has_morphs) {
        morph_beg = __ldg(mesh.morph_row_offsets + row);
        morph_end = __ldg(mesh.morph_row_offsets + row + 1);
      }
      if (grp.selection != nullptr && inst_local == 0) grp.selection[row] = bi;
    }
    const float *sliders = grp.sliders + (size_t)inst * mesh.num_sliders;

    for (int o = 0; o < mesh.num_outputs; ++o) {
      const int stride = mesh.stride[o];
      const int chunk_rows = min(kSkinThreads, rows - rbase);
      if (chunk_rows <= 0) break;
      const size_t offset = (size_t)(row0 + rbase) * stride;  // multiple of 16: row0, rbase are multiples of 64... see launch
      const int bytes = chunk_rows * stride;
      const uint8_t *src = mesh.src[o] + offset;
      uint8_t *dst = grp.out[o] + (size_t)inst * grp.pitch[o] + offset;

      // -------- HBM -> smem, 128-bit coalesced (tail in 32-bit words)
      const int n16 = bytes >> 4;
      const uint4 *src16 = reinterpret_cast<const uint4 *>(src);
      uint4 *tile16 = reinterpret_cast<uint4 *>(tile);
      for (int i = tid; i < n16; i += kSkinThreads) tile16[i] = __ldg(src16 + i);
      for (int i = (n16 << 2) + tid; i < (bytes >> 2); i += kSkinThreads)
        reinterpret_cast<uint32_t *>(tile)[i] = __ldg(reinterpret_cast<const uint32_t *>(src) + i);
      __syncthreads();

      // -------- one thread per row: morph + skin every dynamic column of this array
      if (live) {
        uint8_t *myrow = tile + (size_t)tid * stride;
        for (int c = 0; c < mesh.num_dyn; ++c) {
          const DynColumn col = mesh.dyn[c];
          if (col.output != o) continue;
          animate_column<FULL>(mesh, col, c, reinterpret_cast<float *>(myrow + col.start), morph_beg, morph_end,
                               sliders, skinned, rec);
        }
      }
      __syncthreads();

      // -------- smem -> HBM, whole rows (static bytes included), 128-bit coalesced
      uint4 *dst16 = reinterpret_cast<uint4 *>(dst);
      for (int i = tid; i < n16; i += kSkinThreads) dst16[i] = tile16[i];
      for (int i = (n16 << 2) + tid; i < (bytes >> 2); i += kSkinThreads)
        reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(tile)[i];
      __syncthreads();
    }
  }
}

int skin_rows_per_tile(const MeshDev &) { return kSkinThreads; }

void launch_blend(const MeshDev &mesh, const GroupDev &grp, int first_instance, int count, cudaStream_t stream) {
  if (mesh.num_blends == 0 || count == 0) return;
  int want_normal = 0;
  for (int c = 0; c < mesh.num_dyn; ++c) want_normal |= (mesh.dyn[c].skin == 3);
  dim3 grid((mesh.num_blends + 127) / 128, count);
  if (mesh.full_records) blend_kernel<true><<<grid, 128, 0, stream>>>(mesh, grp, first_instance, want_normal);
  else blend_kernel<false><<<grid, 128, 0, stream>>>(mesh, grp, first_instance, want_normal);
}

void launch_skin(const MeshDev &mesh, const GroupDev &grp, int first_instance, int count, cudaStream_t stream) {
  if (mesh.num_rows == 0 || count == 0) return;
  const int rows_per_tile = skin_rows_per_tile(mesh);
  int max_stride = 4;
  for (int o = 0; o < mesh.num_outputs; ++o) max_stride = max_stride > mesh.stride[o] ? max_stride : mesh.stride[o];
  const size_t smem = (size_t)kSkinThreads * max_stride;
  dim3 grid((mesh.num_rows + rows_per_tile - 1) / rows_per_tile, count);
  if (mesh.full_records) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(skin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    skin_kernel<true><<<grid, kSkinThreads, smem, stream>>>(mesh, grp, first_instance, rows_per_tile);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(skin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    skin_kernel<false><<<grid, kSkinThreads, smem, stream>>>(mesh, grp, first_instance, rows_per_tile);
  }
}

}  // namespace p3cs
