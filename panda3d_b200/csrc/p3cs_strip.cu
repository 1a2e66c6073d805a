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
#include "p3cs_device.cuh"

namespace p3cs {

constexpr int kStripThreads = kTileRows;  // one thread per row of a tile
constexpr int kPalStride4 = 5;            // palette matrix stride in smem: 5 x float4 (80 B) -> LDS.128 without 4-way conflicts

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
  s.tiles_off = (s.records_off + (m.max_tile_records > 0 ? m.max_tile_records : 1) * rec_bytes + 127) & ~127;
  s.total = s.tiles_off + 2 * m.tile_bytes;
  return s;
}

template <bool FULL>
__global__ void __launch_bounds__(kStripThreads) strip_kernel(MeshDev mesh, GroupDev grp, int first_instance, int tiles_per_cta,
                                                            int want_normal) {
  extern __shared__ __align__(128) uint8_t smem[];
  const StripSmem lay = strip_layout(mesh);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
  float4 *pal = reinterpret_cast<float4 *>(smem + lay.palette_off);
  float *recs = reinterpret_cast<float *>(smem + lay.records_off);
  uint8_t *tiles = smem + lay.tiles_off;
  constexpr int REC = FULL ? kRecFull : kRecCompact;

  const int tid = threadIdx.x;
  const int inst = first_instance + blockIdx.y;
  const int t0 = blockIdx.x * tiles_per_cta;
  const int t1 = min(mesh.num_tiles, t0 + tiles_per_cta);
  if (t0 >= t1) return;

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_proxy_async();  // make the initialised barriers visible to the async (TMA) proxy
  }
  // palette: T x 64 B, coalesced 128-bit loads, re-strided to 80 B in shared memory
  {
    const float4 *gp = reinterpret_cast<const float4 *>(grp.palettes + (size_t)inst * mesh.num_transforms * 16);
    const int n4 = mesh.num_transforms * 4;
    for (int i = tid; i < n4; i += kStripThreads) pal[(i >> 2) * kPalStride4 + (i & 3)] = __ldg(gp + i);
  }
  __syncthreads();

  // one elected thread drives the TMA traffic
  auto issue_tile_load = [&](int t, int buf) {
    const int row0 = t * kTileRows;
    const int rows = min(kTileRows, mesh.num_rows - row0);
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

  for (int t = t0; t < t1; ++t) {
    const int it = t - t0;
    const int buf = it & 1;
    if (tid == 0 && t + 1 < t1) {
      bulk_wait_read_all();  // the store that last read buffer buf^1 has drained it
      issue_tile_load(t + 1, buf ^ 1);
    }

    // ---------------- phase A: one thread per distinct blend of this tile
    const int rbeg = __ldg(mesh.tile_rec_offsets + t);
    const int nrec = __ldg(mesh.tile_rec_offsets + t + 1) - rbeg;
    for (int b = tid; b < nrec; b += kStripThreads) {
      const int beg = __ldg(mesh.trec_entry_offsets + rbeg + b);
      const int end = __ldg(mesh.trec_entry_offsets + rbeg + b + 1);
      Mat4 acc, x;
      blend_accumulate([&](int j, int r) { return pal[j * kPalStride4 + r]; }, mesh.trec_transform, mesh.trec_weight, beg, end, acc);
      bool normalize = false;
      if (want_normal) normalize = normal_xform(acc, mesh.coordinate_system, x);
      else x = acc;
      store_record<FULL>(recs + (size_t)b * REC, acc, x, normalize);
    }

    // ---------------- per-row selection (integer work, bit-exact)
    const int row = t * kTileRows + tid;
    const bool live = row < mesh.num_rows;
    int local = 0xFFFF;
    int morph_beg = 0, morph_end = 0;
    if (live) {
      if (mesh.index_bytes != 0) local = __ldg(mesh.tile_local + row);
      if (mesh.has_morphs) {
        morph_beg = __ldg(mesh.morph_row_offsets + row);
        morph_end = __ldg(mesh.morph_row_offsets + row + 1);
      }
      if (grp.selection != nullptr && blockIdx.y == 0)
        grp.selection[row] = local == 0xFFFF ? -1 : __ldg(mesh.tile_rec_blend + rbeg + local);
    }
    const bool skinned = local != 0xFFFF;

    mbar_wait(&mbar[buf], (it >> 1) & 1);  // the tile's rows have landed
    __syncthreads();                       // ... and every record of phase A is written

    // ---------------- phase B: one thread per row, in place on the staged tile
    if (live) {
      Record<FULL> rec;
      if (skinned) load_record<FULL, false>(recs, (size_t)local, rec);
      uint8_t *tile = tiles + buf * mesh.tile_bytes;
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
      for (int o = 0; o < mesh.num_outputs; ++o) {
        const uint32_t bytes = (uint32_t)((rows * mesh.stride[o] + 15) & ~15);
        bulk_s2g(grp.out[o] + (size_t)inst * grp.pitch[o] + (size_t)row0 * mesh.stride[o],
                 tiles + buf * mesh.tile_bytes + mesh.tile_out_offset[o], bytes);
      }
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_read_all();  // shared memory must outlive the last bulk store's reads
}

size_t strip_smem_bytes(const MeshDev &mesh) { return (size_t)strip_layout(mesh).total; }

int launch_strip(const MeshDev &mesh, const GroupDev &grp, int first_instance, int count, int sm_count, cudaStream_t stream) {
  if (mesh.num_rows == 0 || count == 0 || mesh.num_tiles == 0) return 0;
  int want_normal = 0;
  for (int c = 0; c < mesh.num_dyn; ++c) want_normal |= (mesh.dyn[c].skin == 3);
  const size_t smem = strip_smem_bytes(mesh);
  if (mesh.full_records) cudaFuncSetAttribute(strip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  else cudaFuncSetAttribute(strip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // enough strips to fill every SM several times over, as long as possible so the palette is staged rarely
  const long long target_ctas = (long long)sm_count * 24;
  long long tpc = ((long long)mesh.num_tiles * count + target_ctas - 1) / target_ctas;
  if (tpc < 1) tpc = 1;
  if (tpc > mesh.num_tiles) tpc = mesh.num_tiles;
  // balance: equal-length strips
  const int strips = (mesh.num_tiles + (int)tpc - 1) / (int)tpc;
  tpc = (mesh.num_tiles + strips - 1) / strips;
  int launches = 0;
  for (int c0 = 0; c0 < count; c0 += 65535) {
    const int cn = count - c0 < 65535 ? count - c0 : 65535;
    dim3 grid(strips, cn);
    if (mesh.full_records) strip_kernel<true><<<grid, kStripThreads, smem, stream>>>(mesh, grp, first_instance + c0, (int)tpc, want_normal);
    else strip_kernel<false><<<grid, kStripThreads, smem, stream>>>(mesh, grp, first_instance + c0, (int)tpc, want_normal);
    ++launches;
  }
  return launches;
}

}  // namespace p3cs
