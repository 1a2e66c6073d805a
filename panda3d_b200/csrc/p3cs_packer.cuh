// p3cs_packer.cuh -- columns that are not float32 / float64: what GeomVertexReader::get_data3/4 and
// GeomVertexWriter::set_data3/4 do on them (PN_stdfloat = float), i.e. the GeomVertexColumn::Packer family of
// panda/src/gobj/geomVertexColumn.cxx.  The reference animates every column that is not float32 x 3 / x 4 through
// GeomVertexRewriter (geomVertexData.cxx:1782-1799, 1862-1879) and morphs every base column through it (:1489-1537).
//
// Which packer (make_packer, :240-372): colour columns get Packer_color (values scaled by the type's maximum) or,
// for 4 x uint8 and packed_dabc, Packer_rgba_uint8_4 / Packer_argb_packed (clamped to [0,1] on writing); everything
// else the plain Packer -- Packer_point included, because it differs from Packer only in get_data4 on fewer than
// four values and get_data3 on four, which this path never calls (four values: get_data4 / set_data4, fewer:
// get_data3 / set_data3).
//
// Integer conversions are those of the reference's x86-64 build, spelled out because CUDA's saturate where x86
// answers "integer indefinite": `(unsigned int)x` is cvttss2si into a 64-bit register, low half kept;
// `(int)x` is cvttss2si into a 32-bit register.  Bytes are addressed one by one: a column may start anywhere.
#pragma once

#include <cstdint>

#include "../../include/p3cudaskin.h"

namespace p3cs {

enum { kPackPlain = 0, kPackColor = 1, kPackColorClamped = 2 };

// DynColumn::typed: 0 = the float32 / float64 path of animate_column; else (1 + numeric type) | packer kind << 4
__host__ __device__ inline int typed_code(int numeric_type, int kind) { return (1 + numeric_type) | (kind << 4); }
__host__ __device__ inline int typed_numeric(int typed) { return (typed & 15) - 1; }
__host__ __device__ inline int typed_kind(int typed) { return typed >> 4; }

// bytes of a column in its row (GeomVertexColumn::setup, :154-234)
__host__ __device__ inline int column_bytes(int numeric_type, int num_values) {
  switch (numeric_type) {
  case P3CS_NT_UINT8: case P3CS_NT_INT8: return num_values;
  case P3CS_NT_UINT16: case P3CS_NT_INT16: return 2 * num_values;
  case P3CS_NT_FLOAT64: return 8 * num_values;
  case P3CS_NT_PACKED_DCBA: case P3CS_NT_PACKED_DABC: case P3CS_NT_PACKED_UFLOAT: return 4;
  default: return 4 * num_values;
  }
}

// the packer make_packer gives a column of these contents / type / width
__host__ __device__ inline int packer_kind(int contents, int numeric_type, int num_values) {
  if (contents != P3CS_C_COLOR) return kPackPlain;
  if (num_values == 4 && (numeric_type == P3CS_NT_UINT8 || numeric_type == P3CS_NT_PACKED_DABC)) return kPackColorClamped;
  return kPackColor;
}

// Can the reference read and write such a column at all?  (Its packers answer the other combinations with
// nassert(false) and stale values.)
__host__ __device__ inline bool column_supported(int contents, int numeric_type, int num_values) {
  if (num_values < 1 || num_values > 4 || numeric_type < 0 || numeric_type > P3CS_NT_PACKED_UFLOAT || numeric_type == P3CS_NT_STDFLOAT)
    return false;
  if ((numeric_type == P3CS_NT_PACKED_DCBA || numeric_type == P3CS_NT_PACKED_DABC) && num_values != 4) return false;
  if (numeric_type == P3CS_NT_PACKED_UFLOAT && num_values != 3) return false;
  if (packer_kind(contents, numeric_type, num_values) != kPackPlain) {
    if (num_values < 3) return false;                                                                       // :325-328
    if (numeric_type == P3CS_NT_INT8 || numeric_type == P3CS_NT_INT16 || numeric_type == P3CS_NT_INT32) return false;  // :3500-3505
  }
  return true;
}

// what a colour packer divides by on reading and multiplies by on writing; 0: stored as they are
__host__ __device__ inline float color_scale(int numeric_type) {
  switch (numeric_type) {
  case P3CS_NT_UINT8: case P3CS_NT_PACKED_DCBA: case P3CS_NT_PACKED_DABC: return 255.0f;
  case P3CS_NT_UINT16: return 65535.0f;
  case P3CS_NT_UINT32: return 4294967295.0f;
  default: return 0.0f;
  }
}

#ifdef __CUDACC__

// for k in [0, n), n <= 4, unrolled so the values stay in registers
#define P3CS_EACH_VALUE(k, n) _Pragma("unroll") for (int k = 0; k < 4; ++k) if (k < (n))

__device__ __forceinline__ uint32_t x86_cast_uint(float x) {
  return fabsf(x) < 9223372036854775808.0f ? (uint32_t)(unsigned long long)__float2ll_rz(x) : 0u;
}
__device__ __forceinline__ int32_t x86_cast_int(float x) {
  return (x >= -2147483648.0f && x < 2147483648.0f) ? __float2int_rz(x) : (int32_t)0x80000000;
}

__device__ __forceinline__ uint32_t load_u16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
__device__ __forceinline__ uint32_t load_u32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ void store_u16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
__device__ __forceinline__ void store_u32(uint8_t *p, uint32_t v) {
  p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

// GeomVertexData::unpack_ufloat_a/b/c (geomVertexData.I:431-505): an unsigned float of `MBITS` mantissa bits and 5
// exponent bits (bias 15), right-aligned in `field`.  Denormals: ldexpf(mantissa / 2^MBITS, -14), exact in float.
template <int MBITS>
__device__ __forceinline__ float ufloat_field_to_float(uint32_t field) {
  const uint32_t exponent = field >> MBITS;
  if (exponent == 0) return (float)(field & ((1u << MBITS) - 1u)) * (MBITS == 6 ? 9.5367431640625e-07f : 1.9073486328125e-06f);
  uint32_t bits = field << (23 - MBITS);
  bits = exponent == 31 ? (bits | 0x7f800000u) : bits + 0x38000000u;
  return __uint_as_float(bits);
}

// GeomVertexData::pack_ufloat (geomVertexData.I:375-426), r11 g11 b10; the constants (and their quirks: one
// mantissa bit fewer in the denormal branch than the field holds) are the reference's.
__device__ __forceinline__ uint32_t pack_ufloat(float a, float b, float c) {
  const int32_t ia = __float_as_int(a), ib = __float_as_int(b), ic = __float_as_int(c);
  auto nan_or_inf = [](int32_t i) { return (i & 0x7f800000) == 0x7f800000 && (uint32_t)i != 0xff800000u; };
  uint32_t word = 0;
  if (nan_or_inf(ia)) word |= (uint32_t)(ia >> 17) & 0x7ffu;
  else if (ia >= 0x47800000) word |= 0x7bfu;
  else if (ia >= 0x38800000) word |= (uint32_t)((ia >> 17) - 0x1c00);
  else if (ia >= 0x35000000) word |= (((uint32_t)ia & 0x7c0000u) | 0x800000u) >> (130 - (ia >> 23));

  if (nan_or_inf(ib)) word |= (uint32_t)(ib >> 6) & 0x3ff800u;
  else if (ib >= 0x47800000) word |= 0x3df800u;
  else if (ib >= 0x38800000) word |= (uint32_t)((ib >> 6) - 0xe00000) & 0x3ff800u;
  else if (ib >= 0x35000000) word |= ((((uint32_t)ib & 0x7c0000u) | 0x800000u) >> (119 - (ib >> 23))) & 0x1f800u;

  if (nan_or_inf(ic)) word |= ((uint32_t)ic & 0x0ffe0000u) << 4;
  else if (ic >= 0x47800000) word |= 0xf7c00000u;
  else if (ic >= 0x38800000) word |= ((uint32_t)(ic - 0x38000000) << 4) & 0xffc00000u;
  else if (ic >= 0x35000000) word |= (((((uint32_t)ic << 3) & 0x03c00000u) | 0x04000000u) >> (112 - (ic >> 23))) & 0x07c00000u;
  return word;
}

// get_data4 of a typed column (the float32 / float64 ones never come here): its values in the packer's order,
// padded with 0 (a colour of three values: alpha 1), colour types scaled.
__device__ __forceinline__ void typed_get4(int typed, int nv, const uint8_t *p, float v[4]) {
  const int nt = typed_numeric(typed);
  v[0] = v[1] = v[2] = v[3] = 0.0f;
  switch (nt) {
  case P3CS_NT_UINT8:
    P3CS_EACH_VALUE(k, nv) v[k] = (float)p[k];
    break;
  case P3CS_NT_INT8:
    P3CS_EACH_VALUE(k, nv) v[k] = (float)(int8_t)p[k];
    break;
  case P3CS_NT_UINT16:
    P3CS_EACH_VALUE(k, nv) v[k] = (float)load_u16(p + 2 * k);
    break;
  case P3CS_NT_INT16:
    P3CS_EACH_VALUE(k, nv) v[k] = (float)(int16_t)load_u16(p + 2 * k);
    break;
  case P3CS_NT_UINT32:
    P3CS_EACH_VALUE(k, nv) v[k] = (float)load_u32(p + 4 * k);
    break;
  case P3CS_NT_INT32:
    P3CS_EACH_VALUE(k, nv) v[k] = (float)(int32_t)load_u32(p + 4 * k);
    break;
  case P3CS_NT_PACKED_DCBA:  // d c b a = the bytes in memory order (:787-796)
    v[0] = (float)p[0]; v[1] = (float)p[1]; v[2] = (float)p[2]; v[3] = (float)p[3];
    break;
  case P3CS_NT_PACKED_DABC:  // b c d a (:798-807)
    v[0] = (float)p[2]; v[1] = (float)p[1]; v[2] = (float)p[0]; v[3] = (float)p[3];
    break;
  case P3CS_NT_PACKED_UFLOAT: {  // :700-708
    const uint32_t w = load_u32(p);
    v[0] = ufloat_field_to_float<6>(w & 0x7ffu);
    v[1] = ufloat_field_to_float<6>((w >> 11) & 0x7ffu);
    v[2] = ufloat_field_to_float<5>(w >> 22);
    break;
  }
  default:
    break;
  }
  if (typed_kind(typed) != kPackPlain) {
    const float scale = color_scale(nt);
    if (scale != 0.0f) {
      const float recip = __fdiv_rn(1.0f, scale);  // `_v4 /= 255.0f` multiplies by the rounded reciprocal (lvecBase4_src.I:735-745)
      P3CS_EACH_VALUE(k, nv) v[k] = v[k] * recip;
    }
    if (nv == 3) v[3] = 1.0f;  // Packer_color::get_data4f on three values (:3530-3534)
  }
}

// set_data4: fewer than four values take the leading ones (:1863-1875)
__device__ __forceinline__ void typed_set4(int typed, int nv, uint8_t *p, const float in[4]) {
  const int nt = typed_numeric(typed), kind = typed_kind(typed);
  float v[4] = {in[0], in[1], in[2], in[3]};
  if (kind == kPackColorClamped) {  // :4493-4505, :4543-4552: std::clamp(x, 0, 1) * 255
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x = v[k];
      x = x < 0.0f ? 0.0f : (1.0f < x ? 1.0f : x);
      v[k] = x * 255.0f;
    }
  } else if (kind == kPackColor) {
    const float scale = color_scale(nt);
    if (scale != 0.0f)
      P3CS_EACH_VALUE(k, nv) v[k] = v[k] * scale;  // `scaled = data * 255.0f` (:3940 ff.)
  }
  switch (nt) {
  case P3CS_NT_UINT8:
    P3CS_EACH_VALUE(k, nv) p[k] = (uint8_t)x86_cast_uint(v[k]);
    break;
  case P3CS_NT_INT8:
    P3CS_EACH_VALUE(k, nv) p[k] = (uint8_t)x86_cast_int(v[k]);
    break;
  case P3CS_NT_UINT16:
    P3CS_EACH_VALUE(k, nv) store_u16(p + 2 * k, x86_cast_uint(v[k]));
    break;
  case P3CS_NT_INT16:
    P3CS_EACH_VALUE(k, nv) store_u16(p + 2 * k, (uint32_t)x86_cast_int(v[k]));
    break;
  case P3CS_NT_UINT32:
    P3CS_EACH_VALUE(k, nv) store_u32(p + 4 * k, x86_cast_uint(v[k]));
    break;
  case P3CS_NT_INT32:
    P3CS_EACH_VALUE(k, nv) store_u32(p + 4 * k, (uint32_t)x86_cast_int(v[k]));
    break;
  case P3CS_NT_PACKED_DCBA:  // pack_abcd(v3, v2, v1, v0) (:1925-1927): memory bytes v0 v1 v2 v3
    p[0] = (uint8_t)x86_cast_uint(v[0]); p[1] = (uint8_t)x86_cast_uint(v[1]);
    p[2] = (uint8_t)x86_cast_uint(v[2]); p[3] = (uint8_t)x86_cast_uint(v[3]);
    break;
  case P3CS_NT_PACKED_DABC:  // pack_abcd(v3, v0, v1, v2) (:1929-1931): memory bytes v2 v1 v0 v3
    p[0] = (uint8_t)x86_cast_uint(v[2]); p[1] = (uint8_t)x86_cast_uint(v[1]);
    p[2] = (uint8_t)x86_cast_uint(v[0]); p[3] = (uint8_t)x86_cast_uint(v[3]);
    break;
  case P3CS_NT_PACKED_UFLOAT:  // :1844-1846
    store_u32(p, pack_ufloat(v[0], v[1], v[2]));
    break;
  default:
    break;
  }
}

// set_data3: on four values Packer::set_data3f stores w = 0 (:1848-1850), the colour packers alpha = 1 (:3933-3935)
__device__ __forceinline__ void typed_set3(int typed, int nv, uint8_t *p, const float in[3]) {
  const float v[4] = {in[0], in[1], in[2], typed_kind(typed) == kPackPlain ? 0.0f : 1.0f};
  typed_set4(typed, nv, p, v);
}

#undef P3CS_EACH_VALUE
#endif  // __CUDACC__

}  // namespace p3cs
