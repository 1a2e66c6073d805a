// p3cs_strip.cu -- host side of the fused strip kernel (p3cs_strip.cuh): mode choice, strip partition, launch.
// The kernel template is instantiated per mode in p3cs_strip_inst_*.cu.
#include "p3cs_run.cuh"

#include <cstdlib>
#include <cstring>

namespace p3cs {

#define P3CS_EXTERN_MODE(M)                                                                                             \
  extern template void launch_mode<M>(const MeshDev &, const GroupDev &, int, dim3, size_t, int, int, cudaStream_t);
P3CS_EXTERN_MODE(kGenericCompact)
P3CS_EXTERN_MODE(kGenericFull)
P3CS_EXTERN_MODE(kFastP)
P3CS_EXTERN_MODE(kFastPN)
P3CS_EXTERN_MODE(kFastPNVV)
P3CS_EXTERN_MODE(kRow24)
P3CS_EXTERN_MODE(kRow32)
P3CS_EXTERN_MODE(kRow48)
P3CS_EXTERN_MODE(kRow56)
P3CS_EXTERN_MODE(kAlignedPN)
P3CS_EXTERN_MODE(kAlignedPNVV)
#undef P3CS_EXTERN_MODE

// run kernel (p3cs_run.cuh), instantiated in p3cs_run_inst_*.cu
#define P3CS_EXTERN_RUN(M)                                                                                                          \
  extern template void launch_run_mode<M>(const MeshDev &, const RunPlanDev &, const GroupDev &, int, dim3, size_t, int, RunTail, int, cudaStream_t);
P3CS_EXTERN_RUN(kRow24)
P3CS_EXTERN_RUN(kRow32)
P3CS_EXTERN_RUN(kRow48)
P3CS_EXTERN_RUN(kRow56)
#undef P3CS_EXTERN_RUN

size_t strip_smem_bytes(const MeshDev &mesh) { return (size_t)strip_layout(mesh).total; }

// Which specialisation fits this mesh (see the header comment).
int strip_pick_mode(const MeshDev &mesh) {
  if (mesh.aligned) return mesh.num_dyn == 2 ? kAlignedPN : kAlignedPNVV;
  if (mesh.full_records) return kGenericFull;
  if (mesh.num_outputs != 1 || mesh.index_bytes == 0) return kGenericCompact;
  for (int c = 0; c < mesh.num_dyn; ++c)
    if (mesh.dyn[c].num_values != 3 || mesh.dyn[c].f64 || mesh.dyn[c].typed) return kGenericCompact;
  const DynColumn *d = mesh.dyn;
  if (mesh.num_dyn == 1 && d[0].skin == 1) return kFastP;
  const int stride = mesh.stride[0];
  if (mesh.num_dyn == 2 && d[0].skin == 1 && d[1].skin == 3) {
    if (d[0].start == 0 && d[1].start == 12 && stride == 24) return kRow24;
    if (d[0].start == 0 && d[1].start == 12 && stride == 32) return kRow32;
    return kFastPN;
  }
  if (mesh.num_dyn == 4 && d[0].skin == 1 && d[1].skin == 3 && d[2].skin == 2 && d[3].skin == 2) {
    const int t = d[2].start, b = d[3].start;  // p3cs_mesh_create sorts same-kind columns by offset: dyn[2] is the lower one
    if (d[0].start == 0 && d[1].start == 12 && t == 24 && b == 36 && stride == 48) return kRow48;
    if (d[0].start == 0 && d[1].start == 12 && t == 32 && b == 44 && stride == 56) return kRow56;
    return kFastPNVV;
  }
  return kGenericCompact;
}

// strips of whole record groups: enough of them to fill every SM several times over, yet as
// long as possible so the palette is staged rarely
static int strip_partition(const MeshDev &mesh, int count, int sm_count, int *groups_per_cta) {
  static const int per_sm = [] {  // CTAs per SM the partition aims for (P3CS_STRIP_CTAS_PER_SM: tuning aid)
    const char *e = getenv("P3CS_STRIP_CTAS_PER_SM");
    const int v = e != nullptr ? atoi(e) : 0;
    return v > 0 ? v : 24;
  }();
  const long long target_ctas = (long long)sm_count * per_sm;
  long long gpc = ((long long)mesh.num_groups * count + target_ctas - 1) / target_ctas;
  if (gpc < 1) gpc = 1;
  if (gpc > mesh.num_groups) gpc = mesh.num_groups;
  const int strips = (mesh.num_groups + (int)gpc - 1) / (int)gpc;
  *groups_per_cta = (mesh.num_groups + strips - 1) / strips;  // equal-length strips
  // Few CTAs in the whole launch (a single mesh, fewer than two CTAs per SM): split every group's tiles over several
  // CTAs, up to about four CTAs per SM, which is encoded as a negative groups_per_cta (see strip_kernel).
  static const int split_target = [] {  // CTAs per SM below which groups are split (P3CS_STRIP_SPLIT_PER_SM: tuning aid)
    const char *e = getenv("P3CS_STRIP_SPLIT_PER_SM");
    const int v = e != nullptr ? atoi(e) : 0;
    return v > 0 ? v : 4;  // measured: the 250 k-vertex morph mesh 28.2 -> 24.4 us; nothing gained beyond, small crowds lose at 6
  }();
  if (*groups_per_cta == 1 && (long long)strips * count * 2 <= (long long)sm_count * split_target && mesh.num_groups > 0) {
    const int tiles_per_group = (mesh.num_tiles + mesh.num_groups - 1) / mesh.num_groups;
    int split = (int)(((long long)sm_count * split_target) / ((long long)strips * count));
    if (split > tiles_per_group) split = tiles_per_group;
    if (split > 1) {
      *groups_per_cta = -split;
      return strips * split;
    }
  }
  return strips;
}

int strip_mode_row_bytes(int mode) {
  switch (mode) {
  case kRow24: return 24;
  case kRow32: return 32;
  case kRow48: return 48;
  case kRow56: return 56;
  default: return 0;
  }
}

// ---- run kernel: which plan, how many CTAs per instance
// Which kernel takes a launch of `count` instances.  The run kernel (p3cs_run.cuh) wins where the strip kernel is
// bound by its per-row record reads: 24-, 32- and 48-byte rows without morphs (measured on B200, crowds of the C3
// character: 1.25 / 1.81 / 2.14 ms against 1.46 / 1.86 / 2.42 ms).  The strip kernel keeps 56-byte rows (2.24 ms against
// 2.39), meshes with morph sliders (its per-row morph ranges are prefetched a tile ahead), plans that walk badly in
// lockstep, and launches of a few thousand rows.  P3CS_PATH=run / strip forces one of them (A/B runs, parity tests).
static bool use_run_kernel(const MeshDev &mesh, int count) {
  const char *e = getenv("P3CS_PATH");  // read per launch: the parity tests switch it between calls
  const int forced = e == nullptr ? 0 : strcmp(e, "run") == 0 ? 1 : strcmp(e, "strip") == 0 ? -1 : 0;
  if (mesh.run_mode == 0 || forced < 0) return false;
  if (forced > 0) return true;
  return mesh.run_mode != kRow56 && !mesh.has_morphs && mesh.run_lockstep_pct >= 62 && (long long)mesh.num_rows * count >= 8192;
}

struct RunLaunch {
  int plan, ctas_per_instance, tiles_per_cta;
  RunTail tail;  // {-1, 0, 0}: none
};

static RunLaunch run_partition(const MeshDev &mesh, int count, int sm_count, bool grp_plain = false) {
  RunLaunch r{0, 1, 1, {-1, 0, 0}};
  // the crowd plan (large tiles) needs enough instances to fill every SM four times over; else the small tiles
  const RunPlanDev &big = mesh.run_plan[0];
  if (mesh.run_plan[1].tile_rows != 0 && (long long)big.num_tiles * count < 4LL * sm_count) r.plan = 1;
  const RunPlanDev &p = mesh.run_plan[r.plan];
  // One CTA per instance as soon as the instances alone fill the machine about twice over (a CTA stages the instance's
  // palette and sits out its fetch: 10 000 x C3 1.235 ms with one CTA per instance, 1.276 with two; 1250 x C3 0.179 / 0.181
  // / 0.183 ms with 1 / 2 / 3); below that the tiles are spread over more CTAs.
  int threads = 256, ctas = 3;
  run_shape(mesh.run_mode, &threads, &ctas);
  const long long target = 2LL * ctas * sm_count;
  long long cpi = (target + count - 1) / count;
  if (cpi < 1) cpi = 1;
  if (const char *e = getenv("P3CS_RUN_CPI")) {  // tuning aid: CTAs per instance
    const int v = atoi(e);
    if (v > 0) cpi = v;
  }
  if (cpi > p.num_tiles) cpi = p.num_tiles;
  r.tiles_per_cta = (p.num_tiles + (int)cpi - 1) / (int)cpi;
  r.ctas_per_instance = (p.num_tiles + r.tiles_per_cta - 1) / r.tiles_per_cta;
  // Tail split.  With one CTA per instance a launch drains for a good part of an instance's time (the CTAs of the last wave
  // end one by one; an instance of the C3 character takes 55 us, a third of the whole launch at 1250 instances).  So the
  // last `resident` instances go in two pieces each, dealt last (RunGeom): the drain shrinks with the pieces, at the price
  // of staging those instances' palettes twice.  Measured on B200 (C3 character): 1250 instances 0.172 -> 0.167 ms, 10 000
  // instances 1.186 -> 1.185 ms; three pieces 0.169 / 1.183, four 0.172 / 1.187, twice as many instances in three 0.174 /
  // 1.188.  Plain launches only (no bounds, whose per-instance merge depends on the grid shape; an instance list is
  // indexed the same way).  P3CS_RUN_TAIL="instances,parts" tunes, "0" turns off.
  if (r.ctas_per_instance == 1 && grp_plain && p.num_tiles >= 4) {
    int tail_instances = ctas * sm_count, parts = 2;
    if (const char *e = getenv("P3CS_RUN_TAIL")) {
      int a = 0, b = 0;
      if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 0 && b >= 2) tail_instances = a, parts = b;
      else tail_instances = 0;
    }
    if (tail_instances > count) tail_instances = count;
    if (parts > p.num_tiles) parts = p.num_tiles;
    if (tail_instances > 0) {
      const int tpc = (p.num_tiles + parts - 1) / parts;
      r.tail = RunTail{count - tail_instances, (p.num_tiles + tpc - 1) / tpc, tpc};
    }
  }
  return r;
}

int strip_ctas_per_instance(const MeshDev &mesh, int count, int sm_count) {
  if (mesh.num_rows == 0 || count == 0 || mesh.num_tiles == 0) return 0;
  if (use_run_kernel(mesh, count)) return run_partition(mesh, count, sm_count).ctas_per_instance;
  int gpc;
  return strip_partition(mesh, count, sm_count, &gpc);
}

static int launch_run(const MeshDev &mesh, const GroupDev &grp, int first_instance, int count, int sm_count, int want_normal, cudaStream_t stream) {
  const bool plain = grp.selection == nullptr && grp.bounds == nullptr;
  const RunLaunch rl = run_partition(mesh, count, sm_count, plain);
  const RunPlanDev &plan = mesh.run_plan[rl.plan];
  int threads, ctas;
  run_shape(mesh.run_mode, &threads, &ctas);
  const size_t smem = (size_t)run_layout(plan.pal_entries, mesh.num_transforms, plan.tile_rows, strip_mode_row_bytes(mesh.run_mode), plan.tile_bufs, threads).total;
  int launches = 0;
  const int chunk = rl.tail.whole_instances >= 0 ? count : 65535;  // the one-dimensional grid of the tail split has no 65535 limit
  for (int c0 = 0; c0 < count; c0 += chunk) {
    const int cn = count - c0 < chunk ? count - c0 : chunk;
    const dim3 grid = rl.tail.whole_instances >= 0 ? dim3((unsigned)(rl.tail.whole_instances + (count - rl.tail.whole_instances) * rl.tail.parts), 1)
                                                   : dim3(rl.ctas_per_instance, cn);
    const int first = first_instance + c0;
    switch (mesh.run_mode) {
    case kRow24: launch_run_mode<kRow24>(mesh, plan, grp, first, grid, smem, rl.tiles_per_cta, rl.tail, want_normal, stream); break;
    case kRow32: launch_run_mode<kRow32>(mesh, plan, grp, first, grid, smem, rl.tiles_per_cta, rl.tail, want_normal, stream); break;
    case kRow48: launch_run_mode<kRow48>(mesh, plan, grp, first, grid, smem, rl.tiles_per_cta, rl.tail, want_normal, stream); break;
    default: launch_run_mode<kRow56>(mesh, plan, grp, first, grid, smem, rl.tiles_per_cta, rl.tail, want_normal, stream); break;
    }
    ++launches;
  }
  return launches;
}

int launch_strip(const MeshDev &mesh, const GroupDev &grp, int first_instance, int count, int sm_count, cudaStream_t stream) {
  if (mesh.num_rows == 0 || count == 0 || mesh.num_tiles == 0) return 0;
  int want_normal = 0;
  for (int c = 0; c < mesh.num_dyn; ++c) want_normal |= (mesh.dyn[c].skin == 3);
  if (use_run_kernel(mesh, count)) return launch_run(mesh, grp, first_instance, count, sm_count, want_normal, stream);
  const size_t smem = strip_smem_bytes(mesh);
  const int mode = strip_pick_mode(mesh);
  int gpc;
  const int strips = strip_partition(mesh, count, sm_count, &gpc);
  int launches = 0;
  for (int c0 = 0; c0 < count; c0 += 65535) {
    const int cn = count - c0 < 65535 ? count - c0 : 65535;
    const dim3 grid(strips, cn);
    const int first = first_instance + c0;
    switch (mode) {
    case kGenericFull: launch_mode<kGenericFull>(mesh, grp, first, grid, smem, (int)gpc, want_normal, stream); break;
    case kFastP: launch_mode<kFastP>(mesh, grp, first, grid, smem, (int)gpc, want_normal, stream); break;
    case kFastPN: launch_mode<kFastPN>(mesh, grp, first, grid, smem, (int)gpc, want_normal, stream); break;
    case kFastPNVV: launch_mode<kFastPNVV>(mesh, grp, first, grid, smem, (int)gpc, want_normal, stream); break;
    case kRow24: launch_mode<kRow24>(mesh, grp, first, grid, smem, (int)gpc, want_normal, stream); break;
    case kRow32: launch_mode<kRow32>(mesh, grp, first, grid, smem, (int)gpc, want_normal, stream); break;
    case kRow48: launch_mode<kRow48>(mesh, grp, first, grid, smem, (int)gpc, want_normal, stream); break;
    case kRow56: launch_mode<kRow56>(mesh, grp, first, grid, smem, (int)gpc, want_normal, stream); break;
    case kAlignedPN: launch_mode<kAlignedPN>(mesh, grp, first, grid, smem, (int)gpc, want_normal, stream); break;
    case kAlignedPNVV: launch_mode<kAlignedPNVV>(mesh, grp, first, grid, smem, (int)gpc, want_normal, stream); break;
    default: launch_mode<kGenericCompact>(mesh, grp, first, grid, smem, (int)gpc, want_normal, stream); break;
    }
    ++launches;
  }
  return launches;
}

}  // namespace p3cs
