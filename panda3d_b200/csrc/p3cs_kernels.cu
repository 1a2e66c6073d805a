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
#include "p3cs_device.cuh"

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
  blend_accumulate([&](int t, int r) { return __ldg(pal + (size_t)t * 4 + r); }, mesh.blend_transform, mesh.blend_weight,
                   beg, end, acc);
  Mat4 x;
  bool normalize = false;
  if (want_normal) {
    normalize = normal_xform(acc, mesh.coordinate_system, x);
  } else {
    x = acc;
  }
  const size_t rf = FULL ? kRecFull : kRecCompact;
  store_record<FULL>(grp.records + ((size_t)blockIdx.y * mesh.num_blends + b) * rf, acc, x, normalize);
}

// ---------------------------------------------------------------------------- K2
constexpr int kSkinThreads = 256;

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
          load_record<FULL, true>(grp.records, (size_t)inst_local * mesh.num_blends + bi, rec);
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
      uint8_t *dst = output_base(grp, o, inst) + offset;

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

// ---- animated tight bounds for the two-kernel path; the strip kernel fuses the reduction
__global__ void bounds_init_kernel(float *bounds, const int *instance_list, int first, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count * 8) return;
  const int k = i & 7;
  const int inst = instance_list != nullptr ? instance_list[first + (i >> 3)] : first + (i >> 3);
  bounds[(size_t)inst * 8 + k] = k < 3 ? __int_as_float(0x7f800000) : k < 6 ? __int_as_float(0xff800000) : 0.0f;
}

__global__ void __launch_bounds__(256) bounds_scan_kernel(MeshDev mesh, GroupDev grp, int first_instance, int rows_per_cta) {
  __shared__ float scratch[8 * 7];
  const int inst = first_instance + blockIdx.y;
  const int o = grp.bounds_output;
  const uint8_t *base = output_base(grp, o, inst) + grp.bounds_start;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(mesh.num_rows, r0 + rows_per_cta);
  BoundsAcc acc;
  acc.reset();
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    const float *c = reinterpret_cast<const float *>(base + (size_t)r * mesh.stride[o]);
    if (bounds_wants(grp, r)) acc.add(c[0], c[1], c[2]);
  }
  bounds_commit(acc, scratch, grp.bounds + (size_t)inst * 8, false, threadIdx.x, blockDim.x);
}

void launch_bounds_init(const GroupDev &grp, int first_instance, int count, cudaStream_t stream) {
  if (grp.bounds == nullptr || count == 0) return;
  bounds_init_kernel<<<(count * 8 + 255) / 256, 256, 0, stream>>>(grp.bounds, grp.instance_list, first_instance, count);
}

void launch_bounds_scan(const MeshDev &mesh, const GroupDev &grp, int first_instance, int count, cudaStream_t stream) {
  if (grp.bounds == nullptr || count == 0 || mesh.num_rows == 0) return;
  const int rows_per_cta = 8192;
  for (int c0 = 0; c0 < count; c0 += 65535) {
    const int cn = count - c0 < 65535 ? count - c0 : 65535;
    dim3 grid((mesh.num_rows + rows_per_cta - 1) / rows_per_cta, cn);
    bounds_scan_kernel<<<grid, 256, 0, stream>>>(mesh, grp, first_instance + c0, rows_per_cta);
  }
}

}  // namespace p3cs
