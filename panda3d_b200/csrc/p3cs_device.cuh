// p3cs_device.cuh -- device code shared by the kernels: the per-blend record, the column
// transforms (LMatrix4f::xform_point / xform_vec / xform) and the per-row morph + skin step.
#pragma once
#include "p3cs_kernels.cuh"
#include "p3cs_packer.cuh"
#include "p3cs_math.cuh"

namespace p3cs {

// TransformBlend::recompute_result (panda/src/gobj/transformBlend.cxx:226-231): result = zeros;
// for each entry in stored order: LMatrix4f::accumulate(M_i, w_i) (lmatrix4_src.I:1311-1336).
// An empty blend keeps the CData constructor's identity (transformBlend.I:362-366).
// `pal` points at float4 rows; matrix j starts at pal + j * pal_stride4.
template <class LoadRow>
__device__ __forceinline__ void blend_accumulate(LoadRow load_row, const int *xf, const float *w, int beg, int end, Mat4 &acc) {
  if (beg == end) {
    set_ident4(acc);
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc.m[i][j] = 0.0f;
  for (int e = beg; e < end; ++e) {
    const int t = __ldg(xf + e);
    const float weight = __ldg(w + e);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 row = load_row(t, r);
      acc.m[r][0] += row.x * weight;
      acc.m[r][1] += row.y * weight;
      acc.m[r][2] += row.z * weight;
      acc.m[r][3] += row.w * weight;
    }
  }
}

// The same over entries packed as (palette index, weight bits) pairs (strip kernel tables).
template <class LoadRow>
__device__ __forceinline__ void blend_accumulate_packed(LoadRow load_row, const uint2 *entries, int beg, int end, Mat4 &acc) {
  if (beg == end) {
    set_ident4(acc);
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc.m[i][j] = 0.0f;
  for (int e = beg; e < end; ++e) {
    const uint2 en = __ldg(entries + e);
    const float weight = __uint_as_float(en.y);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 row = load_row((int)en.x, r);
      acc.m[r][0] += row.x * weight;
      acc.m[r][1] += row.y * weight;
      acc.m[r][2] += row.z * weight;
      acc.m[r][3] += row.w * weight;
    }
  }
}

// Writes one record (layouts: p3cs_kernels.cuh) to global or shared memory.
template <bool FULL>
__device__ __forceinline__ void store_record(float *dst, const Mat4 &acc, const Mat4 &x, bool normalize) {
  if (FULL) {
    float4 *rec = reinterpret_cast<float4 *>(dst);
#pragma unroll
    for (int r = 0; r < 4; ++r) rec[r] = make_float4(acc.m[r][0], acc.m[r][1], acc.m[r][2], acc.m[r][3]);
#pragma unroll
    for (int r = 0; r < 4; ++r) rec[4 + r] = make_float4(x.m[r][0], x.m[r][1], x.m[r][2], x.m[r][3]);
    rec[8] = make_float4(normalize ? 1.0f : 0.0f, 0.0f, 0.0f, 0.0f);
  } else {
    float2 *rec = reinterpret_cast<float2 *>(dst);
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

// Blend record in registers.
template <bool FULL>
struct Record {
  float m[4][4];  // blended matrix (column 3 only meaningful when FULL)
  float x[4][4];  // normal matrix (upper 3x3 when !FULL)
  bool normalize;
};

template <bool FULL, bool GLOBAL>
__device__ __forceinline__ void load_record(const float *records, size_t index, Record<FULL> &r) {
  if (FULL) {
    const float4 *p = reinterpret_cast<const float4 *>(records + index * kRecFull);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v = GLOBAL ? __ldg(p + i) : p[i];
      r.m[i][0] = v.x; r.m[i][1] = v.y; r.m[i][2] = v.z; r.m[i][3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v = GLOBAL ? __ldg(p + 4 + i) : p[4 + i];
      r.x[i][0] = v.x; r.x[i][1] = v.y; r.x[i][2] = v.z; r.x[i][3] = v.w;
    }
    r.normalize = (GLOBAL ? __ldg(p + 8) : p[8]).x != 0.0f;
  } else {
    const float2 *p = reinterpret_cast<const float2 *>(records + index * kRecCompact);
    float f[22];
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      float2 v = GLOBAL ? __ldg(p + i) : p[i];
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
    r.normalize = f[21] == 1.0f;  // 2 = identity scale: x equals the blend matrix (strip kernel's aligned modes)
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

// An integer or packed dynamic column of one row, in shared memory.  The reference works on the stored value:
// every morph reads it (get_data3/4), adds its delta and stores it (set_data3/4) -- truncating to the column's
// type each time -- and the blend transform does the same once more (geomVertexData.cxx:1489-1537, 1782-1799,
// 1862-1879).  Four values: get_data4 / set_data4; fewer: get_data3 / set_data3; vectors always get_data3 / set_data3.
template <bool FULL>
__device__ __noinline__ void animate_typed_column(const MeshDev &mesh, const DynColumn &col, int dyn_index, uint8_t *cell,
                                                  int morph_beg, int morph_end, const float *sliders, bool skinned,
                                                  const Record<FULL> &rec) {
  const int nv = col.num_values, typed = col.typed;
  float v[4];
  for (int e = morph_beg; e < morph_end; ++e) {
    const uint32_t key = __ldg(mesh.morph_keys + e);
    if ((int)(key >> 16) != dyn_index) continue;
    const float value = __ldg(sliders + (key & 0xffffu));
    if (value == 0.0f) continue;  // `if (slider_value != 0.0f)`, :1483
    const float4 d = __ldg(mesh.morph_deltas + e);
    typed_get4(typed, nv, cell, v);
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
      typed_set4(typed, nv, cell, v);
    } else {  // :1531-1535
      v[0] = v[0] + d.x * value;
      v[1] = v[1] + d.y * value;
      v[2] = v[2] + d.z * value;
      typed_set3(typed, nv, cell, v);
    }
  }
  if (!skinned || col.skin == 0) return;
  typed_get4(typed, nv, cell, v);
  if (col.skin == 1) {  // C_point (:1782-1799)
    if (nv == 4) {
      if (FULL) xform_vec4(v, rec.m);
      typed_set4(typed, nv, cell, v);
    } else {
      xform_point3(v, rec.m);
      typed_set3(typed, nv, cell, v);
    }
  } else {  // C_vector / C_normal (:1862-1879)
    if (col.skin == 2) {
      xform_vec3(v, rec.m);
    } else {
      xform_vec3(v, rec.x);
      if (rec.normalize) normalize3(v[0], v[1], v[2]);
    }
    typed_set3(typed, nv, cell, v);
  }
}

// One dynamic column of one row, in shared memory: morph deltas first
// (geomVertexData.cxx:1462-1543), then the blend transform (:1758-1880).
// float64 columns take the reference's GeomVertexRewriter branches (:1782-1799, :1862-1879): components
// are read as float (get_data3/4 with PN_stdfloat = float), 3-component vector math even on 4-component
// columns, and only the components set_data3/4 touches are written back.
template <bool FULL>
__device__ __forceinline__ void animate_column(const MeshDev &mesh, const DynColumn &col, int dyn_index, float *cell,
                                               int morph_beg, int morph_end, const float *sliders, bool skinned,
                                               const Record<FULL> &rec) {
  const int nv = col.num_values;
  if (col.typed != 0) {  // integer / packed column: every step reads the stored value and stores its result (p3cs_packer.cuh)
    animate_typed_column<FULL>(mesh, col, dyn_index, reinterpret_cast<uint8_t *>(cell), morph_beg, morph_end, sliders, skinned, rec);
    return;
  }
  const bool f64 = col.f64 != 0;
  float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (!f64) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nv) v[k] = cell[k];
  } else {  // the row stride is only 4-byte aligned: assemble the doubles from 32-bit halves
    const int *w = reinterpret_cast<const int *>(cell);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nv) v[k] = (float)__hiloint2double(w[2 * k + 1], w[2 * k]);
  }
  bool dirty3 = false, dirty4 = false;  // what set_data3 / set_data4 rewrote (matters for float64 only)

  for (int e = morph_beg; e < morph_end; ++e) {
    const uint32_t key = __ldg(mesh.morph_keys + e);
    if ((int)(key >> 16) != dyn_index) continue;
    const float value = __ldg(sliders + (key & 0xffffu));
    if (value == 0.0f) continue;  // `if (slider_value != 0.0f)`, :1483
    const float4 d = __ldg(mesh.morph_deltas + e);
    dirty3 = true;
    if (nv == 4) {
      dirty4 = true;
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
    // a column of fewer than three values: set_data3 stored only those, the transform reads the rest back as 0
    if (nv < 3) v[2] = 0.0f;
    if (nv < 2) v[1] = 0.0f;
    dirty3 = true;
    if (col.skin == 1) {  // C_point; fewer than three values: get_data3 pads with 0, set_data3 keeps the leading ones (:1793-1798)
      if (nv < 4) xform_point3(v, rec.m);
      else if (FULL) { xform_vec4(v, rec.m); dirty4 = true; }
    } else if (col.skin == 2) {  // C_vector: xform = mat
      if (nv < 4 || f64) xform_vec3(v, rec.m);
      else if (FULL) xform_vec4(v, rec.m);
    } else {  // C_normal
      if (rec.normalize) {  // table_xform_normal3f, also on 4-component columns (:1854-1855)
        xform_vec3(v, rec.x);
        normalize3(v[0], v[1], v[2]);
      } else if (nv < 4 || f64) {
        xform_vec3(v, rec.x);
      } else if (FULL) {
        xform_vec4(v, rec.x);
      }
    }
    if (f64 && nv == 4 && col.skin != 1) {  // Packer::set_data3f on 4 values stores w = 0 (geomVertexColumn.cxx:1848-1850)
      v[3] = 0.0f;
      dirty4 = true;
    }
  }
  if (!f64) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nv) cell[k] = v[k];
  } else {
    int *w = reinterpret_cast<int *>(cell);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nv && (k < 3 ? dirty3 : dirty4)) {
        const double dv = (double)v[k];
        w[2 * k] = __double2loint(dv);
        w[2 * k + 1] = __double2hiint(dv);
      }
  }
}

// ---- animated tight bounds (GeomPrimitive::calc_tight_bounds, gobj/geomPrimitive.cxx:1646-1830, non-indexed
// no-matrix branch): running min / max of the vertex column and max of length_squared(); NaN components
// never win a comparison there (std::min/max) nor here (fminf/fmaxf).
struct BoundsAcc {
  float lo[3], hi[3], sq;
  __device__ __forceinline__ void reset() {
    lo[0] = lo[1] = lo[2] = __int_as_float(0x7f800000);
    hi[0] = hi[1] = hi[2] = __int_as_float(0xff800000);
    sq = 0.0f;
  }
  __device__ __forceinline__ void add(float x, float y, float z) {
    lo[0] = fminf(lo[0], x); lo[1] = fminf(lo[1], y); lo[2] = fminf(lo[2], z);
    hi[0] = fmaxf(hi[0], x); hi[1] = fmaxf(hi[1], y); hi[2] = fmaxf(hi[2], z);
    sq = fmaxf(sq, x * x + y * y + z * z);  // LVecBase3f::length_squared = dot(*this), left to right
  }
  __device__ __forceinline__ void merge(const BoundsAcc &o) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      lo[k] = fminf(lo[k], o.lo[k]);
      hi[k] = fmaxf(hi[k], o.hi[k]);
    }
    sq = fmaxf(sq, o.sq);
  }
  __device__ __forceinline__ void warp_reduce() {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      BoundsAcc o;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        o.lo[k] = __shfl_xor_sync(0xffffffffu, lo[k], off);
        o.hi[k] = __shfl_xor_sync(0xffffffffu, hi[k], off);
      }
      o.sq = __shfl_xor_sync(0xffffffffu, sq, off);
      merge(o);
    }
  }
};

// float atomic min / max through the integer ordering of IEEE bit patterns (-0 is folded into +0 first)
__device__ __forceinline__ void atomic_min_float(float *addr, float v) {
  v = v + 0.0f;
  if (v != v) return;
  if (v >= 0.0f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
  v = v + 0.0f;
  if (v != v) return;
  if (v >= 0.0f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

// One CTA's contribution to the bounds of one instance: warp shuffles, then `scratch` (>= 8 warps x 7 floats of
// shared memory nobody else uses any more), then thread 0 stores (a CTA owning the whole instance) or merges.
__device__ __forceinline__ void bounds_commit(BoundsAcc acc, float *scratch, float *dst, bool whole_instance, int tid, int nthreads) {
  acc.warp_reduce();
  const int warp = tid >> 5, nwarps = nthreads >> 5;
  if ((tid & 31) == 0) {
    float *w = scratch + warp * 7;
    w[0] = acc.lo[0]; w[1] = acc.lo[1]; w[2] = acc.lo[2];
    w[3] = acc.hi[0]; w[4] = acc.hi[1]; w[5] = acc.hi[2];
    w[6] = acc.sq;
  }
  __syncthreads();
  if (tid != 0) return;
  for (int i = 1; i < nwarps; ++i) {
    const float *w = scratch + i * 7;
    BoundsAcc o;
    o.lo[0] = w[0]; o.lo[1] = w[1]; o.lo[2] = w[2];
    o.hi[0] = w[3]; o.hi[1] = w[4]; o.hi[2] = w[5];
    o.sq = w[6];
    acc.merge(o);
  }
  const bool found = acc.lo[0] <= acc.hi[0] || acc.lo[1] <= acc.hi[1] || acc.lo[2] <= acc.hi[2];
  if (whole_instance) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      dst[k] = acc.lo[k] + 0.0f;
      dst[3 + k] = acc.hi[k] + 0.0f;
    }
    dst[6] = acc.sq;
    dst[7] = found ? 1.0f : 0.0f;
  } else if (found) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      atomic_min_float(dst + k, acc.lo[k]);
      atomic_max_float(dst + 3 + k, acc.hi[k]);
    }
    atomic_max_float(dst + 6, acc.sq);
    dst[7] = 1.0f;
  }
}

}  // namespace p3cs
