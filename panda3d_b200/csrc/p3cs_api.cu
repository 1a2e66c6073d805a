// p3cs_api.cu -- the C-ABI of libp3cudaskin.so (include/p3cudaskin.h): library and context entry points,
// instance groups and the per-frame animate calls, the joint hierarchy (f1) and the plugin table.  Host C++.
// Mesh upload is in p3cs_mesh.cu, memory sharing (peer / IPC / exportable / page-locked host) in
// p3cs_interop.cu, the AT_hardware format in p3cs_format.cu, the kernels in p3cs_strip.cuh, p3cs_kernels.cu
// and p3cs_pose.cu.
#include "p3cs_internal.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace p3cs;

// ------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

int p3cs_fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

// ------------------------------------------------------------------ library
extern "C" int p3cs_api_version(void) { return P3CS_API_VERSION; }

extern "C" int p3cs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" const char *p3cs_last_error(void) { return g_last_error.c_str(); }

// ------------------------------------------------------------------ context
extern "C" int p3cs_context_create(int device, void *stream, p3cs_context *out) {
  if (out == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_context_create: out is NULL");
  *out = nullptr;
  int n = p3cs_device_count();
  if (n <= 0) return p3cs_fail(P3CS_ERR_NO_DEVICE, "no CUDA device available (this backend has no CPU fallback)");
  if (device < 0 || device >= n) return p3cs_fail(P3CS_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
  DeviceGuard g(device);  // the caller's current device is restored on return
  p3cs_context ctx = new (std::nothrow) p3cs_context_s;
  if (ctx == nullptr) return p3cs_fail(P3CS_ERR_NOMEM, "out of host memory");
  ctx->device = device;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (stream != nullptr) {
    ctx->stream = static_cast<cudaStream_t>(stream);
  } else {
    cudaError_t err = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) {
      delete ctx;
      return p3cs_fail(P3CS_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(err));
    }
    ctx->owns_stream = true;
  }
  *out = ctx;
  return P3CS_OK;
}

extern "C" int p3cs_context_destroy(p3cs_context ctx) {
  if (ctx == nullptr) return P3CS_OK;
  DeviceGuard g(ctx->device);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  for (auto st : ctx->side)
    if (st) cudaStreamDestroy(st);
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  for (auto e : ctx->join_ev)
    if (e) cudaEventDestroy(e);
  for (auto e : ctx->pipe_ev) cudaEventDestroy(e);
  for (auto &gg : ctx->graphs) cudaGraphExecDestroy(gg.exec);
  delete ctx;
  return P3CS_OK;
}

extern "C" int p3cs_context_sync(p3cs_context ctx) {
  if (ctx == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "context is NULL");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return P3CS_OK;
}

extern "C" void *p3cs_context_stream(p3cs_context ctx) { return ctx ? ctx->stream : nullptr; }


// ------------------------------------------------------------------ group
static size_t record_floats(const MeshDev &m) { return m.full_records ? kRecFull : kRecCompact; }

extern "C" int p3cs_group_create(p3cs_mesh mesh, int n, const p3cs_group_buffers *buffers, p3cs_group *out) {
  if (out == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_create: out is NULL");
  *out = nullptr;
  if (mesh == nullptr || n <= 0) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_create: bad mesh / num_instances");
  DeviceGuard g(mesh->ctx->device);
  const MeshDev &m = mesh->dev;
  p3cs_group grp = new (std::nothrow) p3cs_group_s;
  if (grp == nullptr) return p3cs_fail(P3CS_ERR_NOMEM, "out of host memory");
  grp->mesh = mesh;
  grp->n = n;
  auto bail = [&](int code) {
    p3cs_group_destroy(grp);
    return code;
  };
  auto alloc = [&](size_t bytes, void **p) -> int {
    CUDA_TRY(cudaMalloc(p, std::max<size_t>(bytes, 256)));
    grp->allocations.push_back(*p);
    return P3CS_OK;
  };
  int rc;
  void *p = nullptr;
  if (buffers != nullptr && buffers->palettes != nullptr) p = buffers->palettes;
  else if ((rc = alloc((size_t)n * m.num_transforms * 64, &p)) != P3CS_OK) return bail(rc);
  grp->dev.palettes = static_cast<const float *>(p);
  if (buffers != nullptr && buffers->sliders != nullptr) p = buffers->sliders;
  else {
    if ((rc = alloc((size_t)n * m.num_sliders * 4, &p)) != P3CS_OK) return bail(rc);
    cudaMemsetAsync(p, 0, std::max<size_t>((size_t)n * m.num_sliders * 4, 256), mesh->ctx->stream);
  }
  grp->dev.sliders = static_cast<const float *>(p);
  for (int o = 0; o < m.num_outputs; ++o) {
    grp->dev.pitch[o] = round_up((size_t)m.num_rows * m.stride[o], 256);
    if (buffers != nullptr && buffers->outputs != nullptr && buffers->outputs[o] != nullptr) p = buffers->outputs[o];
    else if ((rc = alloc(grp->dev.pitch[o] * n, &p)) != P3CS_OK) return bail(rc);
    grp->dev.out[o] = static_cast<uint8_t *>(p);
  }
  // records scratch: sized so one chunk of instances stays L2-resident between K1 and K2
  const size_t rec_bytes = std::max<size_t>((size_t)m.num_blends * record_floats(m) * 4, 4);
  const size_t budget = 32u << 20;
  grp->chunk = (int)std::max<size_t>(1, std::min<size_t>({(size_t)n, budget / rec_bytes, (size_t)65535}));
  if ((rc = alloc(rec_bytes * grp->chunk, &p)) != P3CS_OK) return bail(rc);
  grp->dev.records = static_cast<float *>(p);
  for (auto &e : grp->ev) {
    cudaError_t err = cudaEventCreate(&e);
    if (err != cudaSuccess) return bail(p3cs_fail(P3CS_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(err)));
  }
  *out = grp;
  return P3CS_OK;
}

extern "C" int p3cs_group_destroy(p3cs_group grp) {
  if (grp == nullptr) return P3CS_OK;
  DeviceGuard g(grp->mesh->ctx->device);
  cudaStreamSynchronize(grp->mesh->ctx->stream);
  auto &graphs = grp->mesh->ctx->graphs;  // captured launches of this group die with it
  for (size_t i = 0; i < graphs.size();) {
    if (std::find(graphs[i].groups.begin(), graphs[i].groups.end(), grp) != graphs[i].groups.end()) {
      cudaGraphExecDestroy(graphs[i].exec);
      graphs.erase(graphs.begin() + (long)i);
    } else {
      ++i;
    }
  }
  for (void *p : grp->allocations) cudaFree(p);
  for (auto e : grp->ev)
    if (e) cudaEventDestroy(e);
  for (auto e : grp->prof) cudaEventDestroy(e);
  if (grp->frames_dev != nullptr) cudaFree(grp->frames_dev);
  if (grp->forced_slot_dev != nullptr) cudaFree(grp->forced_slot_dev);
  if (grp->bounds_rows_dev != nullptr) cudaFree(grp->bounds_rows_dev);
  if (grp->forced_values_dev != nullptr) cudaFree(grp->forced_values_dev);
  if (grp->list_dev != nullptr) cudaFree(grp->list_dev);
  delete grp;
  return P3CS_OK;
}

extern "C" size_t p3cs_group_output_pitch(p3cs_group grp, int o) {
  return (grp && o >= 0 && o < grp->mesh->dev.num_outputs) ? grp->dev.pitch[o] : 0;
}
extern "C" void *p3cs_group_palettes_device(p3cs_group grp) { return grp ? const_cast<float *>(grp->dev.palettes) : nullptr; }
extern "C" void *p3cs_group_sliders_device(p3cs_group grp) { return grp ? const_cast<float *>(grp->dev.sliders) : nullptr; }
extern "C" void *p3cs_group_output_device(p3cs_group grp, int o) {
  return (grp && o >= 0 && o < grp->mesh->dev.num_outputs) ? grp->dev.out[o] : nullptr;
}

static int check_span(p3cs_group grp, int first, int count, const char *what) {
  if (grp == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "%s: group is NULL", what);
  if (first < 0 || count < 0 || first + count > grp->n)
    return p3cs_fail(P3CS_ERR_INVALID, "%s: instances [%d,%d) outside the group of %d", what, first, first + count, grp->n);
  return P3CS_OK;
}

extern "C" int p3cs_group_set_palettes(p3cs_group grp, int first, int count, const float *host) {
  int rc = check_span(grp, first, count, "p3cs_group_set_palettes");
  if (rc != P3CS_OK) return rc;
  if (host == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_palettes: host pointer is NULL");
  DeviceGuard g(grp->mesh->ctx->device);
  const size_t per = (size_t)grp->mesh->dev.num_transforms * 16;
  if (per * count == 0) return P3CS_OK;
  CUDA_TRY(cudaMemcpyAsync(const_cast<float *>(grp->dev.palettes) + per * first, host, per * count * 4, cudaMemcpyHostToDevice,
                           grp->mesh->ctx->stream));
  return P3CS_OK;
}

extern "C" int p3cs_group_set_sliders(p3cs_group grp, int first, int count, const float *host) {
  int rc = check_span(grp, first, count, "p3cs_group_set_sliders");
  if (rc != P3CS_OK) return rc;
  const size_t per = (size_t)grp->mesh->dev.num_sliders;
  if (per * count == 0) return P3CS_OK;
  if (host == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_sliders: host pointer is NULL");
  DeviceGuard g(grp->mesh->ctx->device);
  CUDA_TRY(cudaMemcpyAsync(const_cast<float *>(grp->dev.sliders) + per * first, host, per * count * 4, cudaMemcpyHostToDevice,
                           grp->mesh->ctx->stream));
  return P3CS_OK;
}

static int prof_event(p3cs_group grp, size_t i, cudaEvent_t *out) {
  while (grp->prof.size() <= i) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    grp->prof.push_back(e);
  }
  *out = grp->prof[i];
  return P3CS_OK;
}

static int animate_range(p3cs_group grp, int first, int count, bool timed, cudaStream_t stream_override = nullptr,
                         const GroupDev *dev_override = nullptr) {
  const MeshDev &m = grp->mesh->dev;
  const GroupDev &gd = dev_override != nullptr ? *dev_override : grp->dev;
  cudaStream_t s = stream_override != nullptr ? stream_override : grp->mesh->ctx->stream;
  const bool prof = timed && grp->profiling;
  int launches = 0, nblend = 0, nskin = 0, rc;
  size_t pe = 0;
  cudaEvent_t e0, e1;
  grp->prof_kind.clear();
  if (timed) CUDA_TRY(cudaEventRecord(grp->ev[0], s));
  if (grp->mesh->use_strip) {
    if (prof) {
      if ((rc = prof_event(grp, pe++, &e0)) != P3CS_OK || (rc = prof_event(grp, pe++, &e1)) != P3CS_OK) return rc;
      CUDA_TRY(cudaEventRecord(e0, s));
    }
    if (gd.bounds != nullptr && strip_ctas_per_instance(m, count, grp->mesh->ctx->sm_count) != 1) {
      launch_bounds_init(gd, first, count, s);  // several CTAs per instance merge with atomics
      ++launches;
    }
    nskin = launch_strip(m, gd, first, count, grp->mesh->ctx->sm_count, s);
    launches += nskin;
    if (prof) {
      CUDA_TRY(cudaEventRecord(e1, s));
      grp->prof_kind.push_back(1);
    }
  } else
  for (int c0 = first; c0 < first + count; c0 += grp->chunk) {
    const int cn = std::min(grp->chunk, first + count - c0);
    if (m.num_blends > 0) {
      if (prof) {
        if ((rc = prof_event(grp, pe++, &e0)) != P3CS_OK || (rc = prof_event(grp, pe++, &e1)) != P3CS_OK) return rc;
        CUDA_TRY(cudaEventRecord(e0, s));
      }
      launch_blend(m, gd, c0, cn, s);
      if (prof) {
        CUDA_TRY(cudaEventRecord(e1, s));
        grp->prof_kind.push_back(0);
      }
      ++launches;
      ++nblend;
    }
    if (prof) {
      if ((rc = prof_event(grp, pe++, &e0)) != P3CS_OK || (rc = prof_event(grp, pe++, &e1)) != P3CS_OK) return rc;
      CUDA_TRY(cudaEventRecord(e0, s));
    }
    launch_skin(m, gd, c0, cn, s);
    if (prof) {
      CUDA_TRY(cudaEventRecord(e1, s));
      grp->prof_kind.push_back(1);
    }
    ++launches;
    ++nskin;
  }
  if (!grp->mesh->use_strip && gd.bounds != nullptr) {
    launch_bounds_init(gd, first, count, s);
    launch_bounds_scan(m, gd, first, count, s);
    launches += 2;
  }
  if (timed) CUDA_TRY(cudaEventRecord(grp->ev[1], s));
  CUDA_TRY(cudaGetLastError());
  grp->last.launches = launches;
  grp->last.blend_launches = nblend;
  grp->last.skin_launches = nskin;
  grp->last.blend_ms = grp->last.skin_ms = -1.0f;
  grp->timed = timed;
  return P3CS_OK;
}

extern "C" int p3cs_group_animate(p3cs_group grp) {
  if (grp == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_animate: group is NULL");
  DeviceGuard g(grp->mesh->ctx->device);
  return animate_range(grp, 0, grp->n, true);
}

static int copy_output_rows(p3cs_group grp, int o, int first, int count, void *host_dst, cudaStream_t stream);

static int ensure_side_streams(p3cs_context ctx) {
  if (ctx->side_ready) return P3CS_OK;
  // created one by one, marked ready only when all exist: a partial failure is retried (and reported) by the next call
  if (ctx->fork_ev == nullptr) CUDA_TRY(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
  for (int k = 0; k < p3cs_context_s::kSideStreams; ++k) {
    if (ctx->side[k] == nullptr) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->side[k], cudaStreamNonBlocking));
    if (ctx->join_ev[k] == nullptr) CUDA_TRY(cudaEventCreateWithFlags(&ctx->join_ev[k], cudaEventDisableTiming));
  }
  ctx->side_ready = true;
  return P3CS_OK;
}

// The crowd form of p3cs_animate_vertices: host palettes / sliders in, host arrays out, for ALL instances of the
// group, software-pipelined in chunks of instances over three streams (upload, kernels, download) so that the
// PCIe download -- the bound of this call -- never waits for an upload or a kernel after the first chunk.
extern "C" int p3cs_group_animate_host(p3cs_group grp, const float *host_palettes, const float *host_sliders,
                                       void *const *host_outputs, int chunk_instances) {
  if (grp == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_animate_host: group is NULL");
  const MeshDev &m = grp->mesh->dev;
  p3cs_context ctx = grp->mesh->ctx;
  if (m.num_transforms > 0 && host_palettes == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_animate_host: palettes are NULL");
  if (grp->n == 0) return P3CS_OK;
  DeviceGuard g(ctx->device);
  int rc = ensure_side_streams(ctx);
  if (rc != P3CS_OK) return rc;
  int chunk = chunk_instances;
  if (chunk <= 0) {  // ~256 MB of output per chunk: long enough to run the link at full rate, short enough to overlap
    size_t per = 0;
    for (int o = 0; o < m.num_outputs; ++o) per += (size_t)m.num_rows * m.stride[o];
    chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)grp->n, (256u << 20) / std::max<size_t>(per, 1)));
  }
  const int nchunks = (grp->n + chunk - 1) / chunk;
  while ((int)ctx->pipe_ev.size() < 2 * nchunks) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->pipe_ev.push_back(e);
  }
  cudaStream_t up = ctx->side[0], down = ctx->side[1], run = ctx->stream;
  CUDA_TRY(cudaEventRecord(ctx->fork_ev, run));  // both copy streams start after whatever the context did before
  CUDA_TRY(cudaStreamWaitEvent(up, ctx->fork_ev, 0));
  CUDA_TRY(cudaStreamWaitEvent(down, ctx->fork_ev, 0));
  const size_t pal_per = (size_t)m.num_transforms * 16, sl_per = (size_t)m.num_sliders;
  int total_launches = 0;
  // From here on the three streams are forked: every failure falls through to the join below (never an early return),
  // so the copy streams cannot run ahead of the context's stream.
  auto cuda_step = [&](cudaError_t err, const char *what) {
    if (err != cudaSuccess && rc == P3CS_OK)
      rc = p3cs_fail(err == cudaErrorMemoryAllocation ? P3CS_ERR_NOMEM : P3CS_ERR_CUDA, "p3cs_group_animate_host: %s failed: %s", what,
                     cudaGetErrorString(err));
    return rc == P3CS_OK;
  };
  for (int c = 0; c < nchunks && rc == P3CS_OK; ++c) {
    const int first = c * chunk, count = std::min(chunk, grp->n - first);
    if (pal_per > 0 &&
        !cuda_step(cudaMemcpyAsync(const_cast<float *>(grp->dev.palettes) + pal_per * first, host_palettes + pal_per * first,
                                   pal_per * count * 4, cudaMemcpyHostToDevice, up), "palette upload"))
      break;
    if (sl_per > 0 && host_sliders != nullptr &&
        !cuda_step(cudaMemcpyAsync(const_cast<float *>(grp->dev.sliders) + sl_per * first, host_sliders + sl_per * first,
                                   sl_per * count * 4, cudaMemcpyHostToDevice, up), "slider upload"))
      break;
    if (!cuda_step(cudaEventRecord(ctx->pipe_ev[2 * c], up), "cudaEventRecord")) break;
    if (!cuda_step(cudaStreamWaitEvent(run, ctx->pipe_ev[2 * c], 0), "cudaStreamWaitEvent")) break;
    if ((rc = animate_range(grp, first, count, false)) != P3CS_OK) break;
    total_launches += grp->last.launches;
    if (!cuda_step(cudaEventRecord(ctx->pipe_ev[2 * c + 1], run), "cudaEventRecord")) break;
    if (host_outputs != nullptr) {
      if (!cuda_step(cudaStreamWaitEvent(down, ctx->pipe_ev[2 * c + 1], 0), "cudaStreamWaitEvent")) break;
      for (int o = 0; o < m.num_outputs && rc == P3CS_OK; ++o)
        if (host_outputs[o] != nullptr) {
          uint8_t *dst = static_cast<uint8_t *>(host_outputs[o]) + (size_t)m.num_rows * m.stride[o] * first;
          rc = copy_output_rows(grp, o, first, count, dst, down);
        }
    }
  }
  // join: the context's stream (what p3cs_context_sync waits for) ends after the last download
  cudaEventRecord(ctx->join_ev[0], up);
  cudaEventRecord(ctx->join_ev[1], down);
  cudaStreamWaitEvent(run, ctx->join_ev[0], 0);
  cudaStreamWaitEvent(run, ctx->join_ev[1], 0);
  grp->last.launches = total_launches;
  return rc;
}

extern "C" int p3cs_animate_groups(p3cs_group const *groups, int num_groups) {
  if (num_groups < 0 || (num_groups > 0 && groups == nullptr)) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_animate_groups: bad arguments");
  if (num_groups == 0) return P3CS_OK;
  p3cs_context ctx = nullptr;
  for (int i = 0; i < num_groups; ++i) {
    if (groups[i] == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_animate_groups: group %d is NULL", i);
    if (ctx == nullptr) ctx = groups[i]->mesh->ctx;
    else if (groups[i]->mesh->ctx != ctx) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_animate_groups: group %d belongs to another context", i);
  }
  DeviceGuard g(ctx->device);
  if (num_groups == 1) return animate_range(groups[0], 0, groups[0]->n, true);
  {
    const int rc0 = ensure_side_streams(ctx);
    if (rc0 != P3CS_OK) return rc0;
  }
  // fork: group i runs on lane i % (1 + kSideStreams); lane 0 is the context's own stream
  const int lanes = std::min(num_groups, 1 + p3cs_context_s::kSideStreams);
  auto fork_join = [&](bool timed) {
    int rc = P3CS_OK;
    if (cudaEventRecord(ctx->fork_ev, ctx->stream) != cudaSuccess) rc = p3cs_fail(P3CS_ERR_CUDA, "cudaEventRecord failed");
    for (int k = 1; k < lanes && rc == P3CS_OK; ++k)
      if (cudaStreamWaitEvent(ctx->side[k - 1], ctx->fork_ev, 0) != cudaSuccess) rc = p3cs_fail(P3CS_ERR_CUDA, "cudaStreamWaitEvent failed");
    for (int i = 0; i < num_groups && rc == P3CS_OK; ++i) {
      const int lane = i % lanes;
      rc = animate_range(groups[i], 0, groups[i]->n, timed, lane == 0 ? ctx->stream : ctx->side[lane - 1]);
    }
    // join (also on failure, so the side streams never run ahead of the context's stream)
    for (int k = 1; k < lanes; ++k) {
      cudaEventRecord(ctx->join_ev[k - 1], ctx->side[k - 1]);
      cudaStreamWaitEvent(ctx->stream, ctx->join_ev[k - 1], 0);
    }
    return rc;
  };

  // The same set of groups comes back every frame: its fork / join is captured once as a CUDA graph and replayed
  // with one launch call (the ~14 stream calls otherwise delay the later kernels).  Profiling mode, which wants
  // per-launch events, keeps the direct path.
  bool profiling = false;
  for (int i = 0; i < num_groups; ++i) profiling = profiling || groups[i]->profiling;
  static const bool use_graphs = getenv("P3CS_NO_GRAPHS") == nullptr;
  if (profiling || !use_graphs) return fork_join(true);

  p3cs_context_s::GroupsGraph *hit = nullptr;
  for (auto &gg : ctx->graphs) {
    if ((int)gg.groups.size() != num_groups) continue;
    bool same = true;
    for (int i = 0; i < num_groups && same; ++i) same = gg.groups[i] == groups[i] && gg.epochs[i] == groups[i]->epoch;
    if (same) hit = &gg;
  }
  if (hit == nullptr) {
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    const int rc = fork_join(false);
    const cudaError_t end = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc != P3CS_OK || end != cudaSuccess || graph == nullptr) {
      if (graph != nullptr) cudaGraphDestroy(graph);
      cudaGetLastError();
      return rc != P3CS_OK ? rc : fork_join(true);  // capture refused (e.g. a foreign stream state): run directly
    }
    p3cs_context_s::GroupsGraph gg;
    const cudaError_t inst = cudaGraphInstantiate(&gg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (inst != cudaSuccess) {
      cudaGetLastError();
      return fork_join(true);
    }
    for (int i = 0; i < num_groups; ++i) {
      gg.groups.push_back(groups[i]);
      gg.epochs.push_back(groups[i]->epoch);
      gg.launches.push_back(groups[i]->last.launches);
    }
    if (ctx->graphs.size() >= 8) {  // a handful of crowd compositions at most; drop the oldest
      cudaGraphExecDestroy(ctx->graphs.front().exec);
      ctx->graphs.erase(ctx->graphs.begin());
    }
    ctx->graphs.push_back(gg);
    hit = &ctx->graphs.back();
  }
  CUDA_TRY(cudaGraphLaunch(hit->exec, ctx->stream));
  for (int i = 0; i < num_groups; ++i) {
    groups[i]->last.launches = hit->launches[i];
    groups[i]->last.blend_ms = groups[i]->last.skin_ms = groups[i]->last.total_ms = -1.0f;
    groups[i]->timed = false;
  }
  return P3CS_OK;
}

// Only the instances that changed: what the UpdateSeq test of GeomVertexData::animate_vertices
// (geomVertexData.cxx:937-946) does per character, for a crowd.
extern "C" int p3cs_group_animate_instances(p3cs_group grp, const int32_t *host_instances, int count) {
  if (grp == nullptr || count < 0 || (count > 0 && host_instances == nullptr)) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_animate_instances: bad arguments");
  for (int i = 0; i < count; ++i)
    if (host_instances[i] < 0 || host_instances[i] >= grp->n)
      return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_animate_instances: entry %d = %d outside the group of %d", i, host_instances[i], grp->n);
  if (count == 0) return P3CS_OK;
  DeviceGuard g(grp->mesh->ctx->device);
  if (!grp->mesh->use_strip) {  // two-kernel path: its record scratch is indexed by launch position, go one by one
    int rc = P3CS_OK;
    for (int i = 0; i < count && rc == P3CS_OK; ++i) rc = animate_range(grp, host_instances[i], 1, false);
    return rc;
  }
  if (grp->list_capacity < count) {
    if (grp->list_dev != nullptr) {
      CUDA_TRY(cudaStreamSynchronize(grp->mesh->ctx->stream));
      CUDA_TRY(cudaFree(grp->list_dev));
      grp->list_dev = nullptr;
    }
    CUDA_TRY(cudaMalloc(&grp->list_dev, (size_t)count * 4));
    grp->list_capacity = count;
  }
  CUDA_TRY(cudaMemcpyAsync(grp->list_dev, host_instances, (size_t)count * 4, cudaMemcpyHostToDevice, grp->mesh->ctx->stream));
  GroupDev listed = grp->dev;
  listed.instance_list = grp->list_dev;
  return animate_range(grp, 0, count, true, nullptr, &listed);
}

extern "C" int p3cs_group_set_output_pointers(p3cs_group grp, int output, void *const *ptrs) {
  if (grp == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_output_pointers: group is NULL");
  const MeshDev &m = grp->mesh->dev;
  if (output < 0 || output >= m.num_outputs) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_output_pointers: output %d out of range", output);
  DeviceGuard g(grp->mesh->ctx->device);
  cudaStream_t s = grp->mesh->ctx->stream;
  ++grp->epoch;
  if (ptrs == nullptr) {
    grp->dev.out_ptrs[output] = nullptr;
    return P3CS_OK;
  }
  std::vector<void *> dev((size_t)grp->n);
  for (int i = 0; i < grp->n; ++i) {
    if (ptrs[i] == nullptr || reinterpret_cast<uintptr_t>(ptrs[i]) % 16 != 0)
      return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_output_pointers: destination %d is NULL or not 16-byte aligned", i);
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, ptrs[i]) != cudaSuccess || attr.devicePointer == nullptr ||
        (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeHost && attr.type != cudaMemoryTypeManaged)) {
      cudaGetLastError();
      return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_output_pointers: destination %d is neither device memory nor page-locked host memory "
                       "(p3cs_host_register)", i);
    }
    dev[(size_t)i] = attr.devicePointer;
  }
  if (grp->out_ptrs_dev[output] == nullptr) {
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, (size_t)grp->n * sizeof(void *)));
    grp->allocations.push_back(p);
    grp->out_ptrs_dev[output] = static_cast<void **>(p);
  } else {
    CUDA_TRY(cudaStreamSynchronize(s));  // an animate in flight may still read the old table
  }
  CUDA_TRY(cudaMemcpyAsync(grp->out_ptrs_dev[output], dev.data(), dev.size() * sizeof(void *), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));  // `dev` dies with this frame
  grp->dev.out_ptrs[output] = reinterpret_cast<uint8_t *const *>(grp->out_ptrs_dev[output]);
  return P3CS_OK;
}

extern "C" int p3cs_group_set_profiling(p3cs_group grp, int enable) {
  if (grp == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_profiling: group is NULL");
  grp->profiling = enable != 0;
  ++grp->epoch;
  return P3CS_OK;
}

extern "C" int p3cs_group_last_timings(p3cs_group grp, p3cs_timings *out) {
  if (grp == nullptr || out == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_last_timings: NULL argument");
  DeviceGuard g(grp->mesh->ctx->device);
  if (grp->timed) {
    CUDA_TRY(cudaEventSynchronize(grp->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&grp->last.total_ms, grp->ev[0], grp->ev[1]));
    if (!grp->prof_kind.empty()) {
      float sum[2] = {0.0f, 0.0f};
      for (size_t i = 0; i < grp->prof_kind.size(); ++i) {
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, grp->prof[2 * i], grp->prof[2 * i + 1]));
        sum[grp->prof_kind[i]] += ms;
      }
      grp->last.blend_ms = sum[0];
      grp->last.skin_ms = sum[1];
    }
  }
  *out = grp->last;
  return P3CS_OK;
}

extern "C" int p3cs_group_enable_bounds(p3cs_group grp, const p3cs_column_desc *vertex_column) {
  if (grp == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_enable_bounds: group is NULL");
  ++grp->epoch;
  if (vertex_column == nullptr) {  // off (the buffer stays allocated until the group is destroyed)
    grp->dev.bounds = nullptr;
    return P3CS_OK;
  }
  const p3cs_mesh_s *mesh = grp->mesh;
  const MeshDev &m = mesh->dev;
  const p3cs_column_desc &c = *vertex_column;
  if (c.array < 0 || c.array >= (int)mesh->out_of_array.size() || mesh->out_of_array[c.array] < 0)
    return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_enable_bounds: column is not in an output array");
  const int o = mesh->out_of_array[c.array];
  // GeomVertexReader::get_data3 on anything else divides by w or converts (geomVertexColumn.cxx:2800-2808)
  if (c.numeric_type != P3CS_NT_FLOAT32 || c.num_values != 3 || c.start < 0 || c.start % 4 != 0 || c.start + 12 > m.stride[o])
    return p3cs_fail(P3CS_ERR_UNSUPPORTED, "p3cs_group_enable_bounds: only a 4-byte aligned float32 x 3 column is supported");
  DeviceGuard g(mesh->ctx->device);
  if (grp->bounds_buf == nullptr) {
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, std::max<size_t>((size_t)grp->n * 32, 256)));
    grp->allocations.push_back(p);
    grp->bounds_buf = static_cast<float *>(p);
    GroupDev only_bounds{};
    only_bounds.bounds = grp->bounds_buf;
    launch_bounds_init(only_bounds, 0, grp->n, mesh->ctx->stream);
  }
  grp->dev.bounds = grp->bounds_buf;
  grp->dev.bounds_output = o;
  grp->dev.bounds_start = c.start;
  return P3CS_OK;
}

extern "C" int p3cs_group_set_bounds_rows(p3cs_group grp, const uint32_t *host_row_bits) {
  if (grp == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_bounds_rows: group is NULL");
  ++grp->epoch;
  DeviceGuard g(grp->mesh->ctx->device);
  cudaStream_t s = grp->mesh->ctx->stream;
  if (host_row_bits == nullptr) {  // back to "every row"; the buffer stays until the group goes
    grp->dev.bounds_rows = nullptr;
    return P3CS_OK;
  }
  const size_t words = ((size_t)grp->mesh->dev.num_rows + 31) / 32;
  if (grp->bounds_rows_dev == nullptr) CUDA_TRY(cudaMalloc(&grp->bounds_rows_dev, std::max<size_t>(words, 1) * 4));
  else CUDA_TRY(cudaStreamSynchronize(s));  // an animate in flight may still read the old bits
  CUDA_TRY(cudaMemcpyAsync(grp->bounds_rows_dev, host_row_bits, words * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  grp->dev.bounds_rows = grp->bounds_rows_dev;
  return P3CS_OK;
}

extern "C" int p3cs_group_read_bounds(p3cs_group grp, int first, int count, float *host_dst) {
  int rc = check_span(grp, first, count, "p3cs_group_read_bounds");
  if (rc != P3CS_OK) return rc;
  if (host_dst == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_read_bounds: destination is NULL");
  if (grp->dev.bounds == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_read_bounds: bounds are not enabled (p3cs_group_enable_bounds)");
  if (count == 0) return P3CS_OK;
  DeviceGuard g(grp->mesh->ctx->device);
  CUDA_TRY(cudaMemcpyAsync(host_dst, grp->dev.bounds + (size_t)first * 8, (size_t)count * 32, cudaMemcpyDeviceToHost, grp->mesh->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(grp->mesh->ctx->stream));
  return P3CS_OK;
}

// Device -> host copy of `count` instances of output `o`, tightly packed at the destination.  When the device
// pitch equals the packed size (rows * stride a multiple of 256) the span is one linear copy, which the copy
// engine moves faster than a pitched one.
static int copy_output_rows(p3cs_group grp, int o, int first, int count, void *host_dst, cudaStream_t stream) {
  const MeshDev &m = grp->mesh->dev;
  const size_t row_bytes = (size_t)m.num_rows * m.stride[o];
  const uint8_t *src = grp->dev.out[o] + grp->dev.pitch[o] * first;
  if (grp->dev.pitch[o] == row_bytes) CUDA_TRY(cudaMemcpyAsync(host_dst, src, row_bytes * count, cudaMemcpyDeviceToHost, stream));
  else CUDA_TRY(cudaMemcpy2DAsync(host_dst, row_bytes, src, grp->dev.pitch[o], row_bytes, count, cudaMemcpyDeviceToHost, stream));
  return P3CS_OK;
}

extern "C" int p3cs_group_read_output(p3cs_group grp, int o, int first, int count, void *host_dst) {
  int rc = check_span(grp, first, count, "p3cs_group_read_output");
  if (rc != P3CS_OK) return rc;
  const MeshDev &m = grp->mesh->dev;
  if (o < 0 || o >= m.num_outputs || host_dst == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_read_output: bad output / destination");
  if (grp->dev.out_ptrs[o] != nullptr)
    return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_read_output: output %d goes to per-instance destinations (p3cs_group_set_output_pointers)", o);
  DeviceGuard g(grp->mesh->ctx->device);
  const size_t row_bytes = (size_t)m.num_rows * m.stride[o];
  if (row_bytes == 0 || count == 0) return P3CS_OK;
  return copy_output_rows(grp, o, first, count, host_dst, grp->mesh->ctx->stream);
}

extern "C" int p3cs_group_read_selection(p3cs_group grp, int32_t *host_dst) {
  if (grp == nullptr || host_dst == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_read_selection: NULL argument");
  const MeshDev &m = grp->mesh->dev;
  DeviceGuard g(grp->mesh->ctx->device);
  cudaStream_t s = grp->mesh->ctx->stream;
  int32_t *sel = nullptr;
  CUDA_TRY(cudaMalloc(&sel, std::max<size_t>((size_t)m.num_rows * 4, 4)));
  GroupDev tmp = grp->dev;
  tmp.selection = sel;
  if (grp->mesh->use_strip) {
    launch_strip(m, tmp, 0, 1, grp->mesh->ctx->sm_count, s);
  } else {
    if (m.num_blends > 0) launch_blend(m, tmp, 0, 1, s);
    launch_skin(m, tmp, 0, 1, s);
  }
  cudaError_t err = cudaMemcpyAsync(host_dst, sel, (size_t)m.num_rows * 4, cudaMemcpyDeviceToHost, s);
  if (err == cudaSuccess) err = cudaStreamSynchronize(s);
  cudaFree(sel);
  if (err != cudaSuccess) return p3cs_fail(P3CS_ERR_CUDA, "p3cs_group_read_selection: %s", cudaGetErrorString(err));
  return P3CS_OK;
}

extern "C" int p3cs_group_read_blends(p3cs_group grp, int instance, float *host_dst) {
  int rc = check_span(grp, instance, 1, "p3cs_group_read_blends");
  if (rc != P3CS_OK) return rc;
  if (host_dst == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_read_blends: destination is NULL");
  const MeshDev &m = grp->mesh->dev;
  if (m.num_blends == 0) return P3CS_OK;
  DeviceGuard g(grp->mesh->ctx->device);
  cudaStream_t s = grp->mesh->ctx->stream;
  launch_blend(m, grp->dev, instance, 1, s);
  const size_t rf = record_floats(m);
  std::vector<float> rec((size_t)m.num_blends * rf);
  CUDA_TRY(cudaMemcpyAsync(rec.data(), grp->dev.records, rec.size() * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  for (int b = 0; b < m.num_blends; ++b) {
    const float *r = rec.data() + (size_t)b * rf;
    float *o = host_dst + (size_t)b * 16;
    if (m.full_records) {
      memcpy(o, r, 64);
    } else {
      // compact records keep rows 0..3 x columns 0..2; column 3 is not stored
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 3; ++j) o[i * 4 + j] = r[i * 3 + j];
        o[i * 4 + 3] = NAN;
      }
    }
  }
  return P3CS_OK;
}

// Waits for the context's stream the way a latency-bound caller wants it: polling (a blocking
// cudaStreamSynchronize costs several microseconds of wake-up), falling back to the blocking wait for long work.
static int sync_polling(p3cs_context ctx) {
  for (int spin = 0; spin < 4000; ++spin) {
    const cudaError_t q = cudaStreamQuery(ctx->stream);
    if (q == cudaSuccess) return P3CS_OK;
    if (q != cudaErrorNotReady) return p3cs_fail(P3CS_ERR_CUDA, "stream failed: %s", cudaGetErrorString(q));
  }
  return p3cs_context_sync(ctx);
}

extern "C" int p3cs_animate_vertices(p3cs_group grp, int instance, const float *host_palette, const float *host_sliders,
                                     void *const *host_outputs) {
  int rc = check_span(grp, instance, 1, "p3cs_animate_vertices");
  if (rc != P3CS_OK) return rc;
  const MeshDev &m = grp->mesh->dev;
  p3cs_context ctx = grp->mesh->ctx;
  if (m.num_transforms > 0 && host_palette == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_animate_vertices: palette is NULL");
  DeviceGuard g(ctx->device);
  GroupDev direct = grp->dev;
  if (m.num_transforms > 0 && (rc = p3cs_group_set_palettes(grp, instance, 1, host_palette)) != P3CS_OK) return rc;
  if (m.num_sliders > 0 && (rc = p3cs_group_set_sliders(grp, instance, 1, host_sliders)) != P3CS_OK) return rc;
  // ---- outputs: page-locked, device-mapped destinations (p3cs_host_register / cudaHostRegister): the kernel's row stores
  // go straight to host memory, no device->host copy afterwards.  The bulk stores move 16-byte units, hence the alignment
  // and size conditions; anything else takes the copy below.  The destinations of the previous call are remembered.
  if (host_outputs != nullptr && m.num_outputs > 0) {
    bool ok = true;
    for (int o = 0; o < m.num_outputs && ok; ++o) {
      const size_t bytes = (size_t)m.num_rows * m.stride[o];
      ok = host_outputs[o] != nullptr && bytes % 16 == 0 && reinterpret_cast<uintptr_t>(host_outputs[o]) % 16 == 0;
      if (!ok) break;
      void *dev = nullptr;
      if (ctx->seen_host[o] == host_outputs[o]) {
        dev = ctx->seen_dev[o];
      } else {
        cudaPointerAttributes attr{};
        ok = cudaPointerGetAttributes(&attr, host_outputs[o]) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
        cudaGetLastError();  // cudaPointerGetAttributes on plain malloc memory may leave a sticky-free error behind
        if (ok) {
          dev = attr.devicePointer;
          ctx->seen_host[o] = host_outputs[o];
          ctx->seen_dev[o] = dev;
        }
      }
      if (ok) direct.out[o] = reinterpret_cast<uint8_t *>(reinterpret_cast<uintptr_t>(dev) - (uintptr_t)instance * grp->dev.pitch[o]);
      direct.out_ptrs[o] = nullptr;
    }
    if (ok) {
      if ((rc = animate_range(grp, instance, 1, false, nullptr, &direct)) != P3CS_OK) return rc;
      return sync_polling(ctx);
    }
    for (int o = 0; o < m.num_outputs; ++o) {
      direct.out[o] = grp->dev.out[o];
      direct.out_ptrs[o] = grp->dev.out_ptrs[o];
    }
  }
  if ((rc = animate_range(grp, instance, 1, false, nullptr, &direct)) != P3CS_OK) return rc;
  if (host_outputs != nullptr)
    for (int o = 0; o < m.num_outputs; ++o)
      if (host_outputs[o] != nullptr && (rc = p3cs_group_read_output(grp, o, instance, 1, host_outputs[o])) != P3CS_OK) return rc;
  return sync_polling(ctx);
}

// ------------------------------------------------------------------ joint hierarchy (f1)
struct p3cs_skeleton_s {
  p3cs_context ctx = nullptr;
  SkeletonDev dev{};
  int num_palette = 0;
  std::vector<void *> allocations;
};

template <class T>
static int upload_skel(p3cs_skeleton sk, const T *host, size_t count, const T **out) {
  void *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, std::max<size_t>(count, 1) * sizeof(T)));
  sk->allocations.push_back(d);
  if (count > 0) CUDA_TRY(cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, sk->ctx->stream));
  *out = static_cast<const T *>(d);
  return P3CS_OK;
}

extern "C" int p3cs_skeleton_create(p3cs_context ctx, const p3cs_skeleton_desc *d, p3cs_skeleton *out) {
  if (out == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_skeleton_create: out is NULL");
  *out = nullptr;
  if (ctx == nullptr || d == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_skeleton_create: NULL argument");
  if (d->struct_size != sizeof(p3cs_skeleton_desc)) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_skeleton_desc::struct_size mismatch");
  const int J = d->num_joints;
  // pose_kernel keeps the local and the net transform of every joint of an instance in shared memory (128 B per
  // joint, 227 KB per CTA on sm_100a)
  constexpr int kMaxPoseJoints = (227 * 1024) / 128;
  if (J < 0 || J > kMaxPoseJoints) return p3cs_fail(P3CS_ERR_UNSUPPORTED, "num_joints %d outside 0..%d", J, kMaxPoseJoints);
  int cs = d->coordinate_system == 0 ? P3CS_CS_ZUP_RIGHT : d->coordinate_system;
  if (cs < P3CS_CS_ZUP_RIGHT || cs > P3CS_CS_YUP_LEFT) return p3cs_fail(P3CS_ERR_INVALID, "bad coordinate_system %d", cs);
  std::vector<int> level(std::max(J, 1), 0);
  int max_level = 0;
  for (int j = 0; j < J; ++j) {
    const int p = d->parent[j];
    if (p >= j) return p3cs_fail(P3CS_ERR_INVALID, "joint %d: parent %d does not precede it (joints must be listed parents-first)", j, p);
    level[j] = p < 0 ? 0 : level[p] + 1;
    max_level = std::max(max_level, level[j]);
    for (int i = 0; i < 12; ++i) {
      const int len = d->table_length[j * 12 + i], off = d->table_offset[j * 12 + i];
      if (len < 0 || off < 0 || off + len > d->num_table_values) return p3cs_fail(P3CS_ERR_INVALID, "joint %d: table %d out of range", j, i);
    }
  }
  for (int t = 0; t < d->num_palette; ++t)
    if (d->palette_joint[t] < 0 || d->palette_joint[t] >= J) return p3cs_fail(P3CS_ERR_INVALID, "palette slot %d has no joint", t);
  if (d->num_sliders < 0 || (d->num_sliders > 0 && (d->slider_has_channel == nullptr || d->slider_table_length == nullptr ||
                                                    d->slider_table_offset == nullptr || d->slider_default == nullptr)))
    return p3cs_fail(P3CS_ERR_INVALID, "p3cs_skeleton_create: slider arrays missing");
  for (int sl = 0; sl < d->num_sliders; ++sl) {
    const int len = d->slider_table_length[sl], off = d->slider_table_offset[sl];
    if (len < 0 || off < 0 || off + len > d->num_table_values) return p3cs_fail(P3CS_ERR_INVALID, "slider %d: table out of range", sl);
  }
  DeviceGuard g(ctx->device);
  p3cs_skeleton sk = new (std::nothrow) p3cs_skeleton_s;
  if (sk == nullptr) return p3cs_fail(P3CS_ERR_NOMEM, "out of host memory");
  sk->ctx = ctx;
  sk->num_palette = d->num_palette;
  SkeletonDev &s = sk->dev;
  s.num_joints = J;
  s.coordinate_system = cs;
  s.max_level = max_level;
  int rc;
  auto bail = [&](int code) {
    p3cs_skeleton_destroy(sk);
    return code;
  };
  if ((rc = upload_skel<int>(sk, d->parent, J, &s.parent)) != P3CS_OK) return bail(rc);
  std::vector<int> order(std::max(J, 1), 0), begin(max_level + 2, 0);
  for (int j = 0; j < J; ++j) begin[level[j] + 1]++;
  for (int l = 0; l <= max_level; ++l) begin[l + 1] += begin[l];
  {
    std::vector<int> cursor(begin.begin(), begin.end() - 1);
    for (int j = 0; j < J; ++j) order[cursor[level[j]]++] = j;
  }
  if ((rc = upload_skel<int>(sk, order.data(), J, &s.level_order)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<int>(sk, begin.data(), begin.size(), &s.level_begin)) != P3CS_OK) return bail(rc);
  {
    // every joint's chain of ancestors, outermost first: the kernel multiplies down the chain per joint instead of
    // synchronising the CTA once per tree level
    std::vector<int> poff((size_t)J + 1, 0), pj;
    std::vector<int> chain;
    for (int j = 0; j < J; ++j) {
      chain.clear();
      for (int a = j; a >= 0; a = d->parent[a]) chain.push_back(a);
      pj.insert(pj.end(), chain.rbegin(), chain.rend());
      poff[j + 1] = (int)pj.size();
    }
    if (pj.empty()) pj.push_back(0);
    if ((rc = upload_skel<int>(sk, poff.data(), poff.size(), &s.path_offsets)) != P3CS_OK) return bail(rc);
    if ((rc = upload_skel<int>(sk, pj.data(), pj.size(), &s.path_joints)) != P3CS_OK) return bail(rc);
  }
  s.num_animations = 1;  // animation 0: the channel tables of the descriptor
  if ((rc = upload_skel<int>(sk, d->has_channel, J, &s.anim[0].has_channel)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<int>(sk, d->table_length, (size_t)J * 12, &s.anim[0].table_length)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<int>(sk, d->table_offset, (size_t)J * 12, &s.anim[0].table_offset)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<float>(sk, d->tables, d->num_table_values, &s.anim[0].tables)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<float>(sk, d->default_value, (size_t)J * 16, &s.default_value)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<float>(sk, d->initial_inverse, (size_t)J * 16, &s.initial_inverse)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<float>(sk, d->root_xform, 16, &s.root_xform)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<int>(sk, d->palette_joint, d->num_palette, &s.palette_joint)) != P3CS_OK) return bail(rc);
  s.num_sliders = d->num_sliders;
  if (d->num_sliders > 0) {
    if ((rc = upload_skel<float>(sk, d->slider_default, d->num_sliders, &s.slider_default)) != P3CS_OK) return bail(rc);
    if ((rc = upload_skel<int>(sk, d->slider_has_channel, d->num_sliders, &s.anim[0].slider_has_channel)) != P3CS_OK) return bail(rc);
    if ((rc = upload_skel<int>(sk, d->slider_table_length, d->num_sliders, &s.anim[0].slider_table_length)) != P3CS_OK) return bail(rc);
    if ((rc = upload_skel<int>(sk, d->slider_table_offset, d->num_sliders, &s.anim[0].slider_table_offset)) != P3CS_OK) return bail(rc);
  }
  cudaError_t err = cudaStreamSynchronize(ctx->stream);
  if (err != cudaSuccess) return bail(p3cs_fail(P3CS_ERR_CUDA, "skeleton upload failed: %s", cudaGetErrorString(err)));
  *out = sk;
  return P3CS_OK;
}

extern "C" int p3cs_skeleton_destroy(p3cs_skeleton sk) {
  if (sk == nullptr) return P3CS_OK;
  DeviceGuard g(sk->ctx->device);
  for (void *p : sk->allocations) cudaFree(p);
  delete sk;
  return P3CS_OK;
}

extern "C" int p3cs_skeleton_add_animation(p3cs_skeleton sk, const p3cs_animation_desc *d, int32_t *out_index) {
  if (sk == nullptr || d == nullptr || out_index == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_skeleton_add_animation: NULL argument");
  if (d->struct_size != sizeof(p3cs_animation_desc)) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_animation_desc::struct_size mismatch");
  SkeletonDev &s = sk->dev;
  if (s.num_animations >= kMaxAnimations) return p3cs_fail(P3CS_ERR_UNSUPPORTED, "a skeleton holds at most %d animations", kMaxAnimations);
  const int J = s.num_joints;
  for (int j = 0; j < J; ++j)
    for (int i = 0; i < 12; ++i) {
      const int len = d->table_length[j * 12 + i], off = d->table_offset[j * 12 + i];
      if (len < 0 || off < 0 || off + len > d->num_table_values) return p3cs_fail(P3CS_ERR_INVALID, "joint %d: table %d out of range", j, i);
    }
  DeviceGuard g(sk->ctx->device);
  AnimationDev an{};
  int rc;
  if ((rc = upload_skel<int>(sk, d->has_channel, J, &an.has_channel)) != P3CS_OK) return rc;
  if ((rc = upload_skel<int>(sk, d->table_length, (size_t)J * 12, &an.table_length)) != P3CS_OK) return rc;
  if ((rc = upload_skel<int>(sk, d->table_offset, (size_t)J * 12, &an.table_offset)) != P3CS_OK) return rc;
  if ((rc = upload_skel<float>(sk, d->tables, d->num_table_values, &an.tables)) != P3CS_OK) return rc;
  if (s.num_sliders > 0) {
    if (d->slider_has_channel == nullptr || d->slider_table_length == nullptr || d->slider_table_offset == nullptr)
      return p3cs_fail(P3CS_ERR_INVALID, "p3cs_skeleton_add_animation: the skeleton has sliders, the animation's slider arrays are missing");
    for (int sl = 0; sl < s.num_sliders; ++sl) {
      const int len = d->slider_table_length[sl], off = d->slider_table_offset[sl];
      if (len < 0 || off < 0 || off + len > d->num_table_values) return p3cs_fail(P3CS_ERR_INVALID, "slider %d: table out of range", sl);
    }
    if ((rc = upload_skel<int>(sk, d->slider_has_channel, s.num_sliders, &an.slider_has_channel)) != P3CS_OK) return rc;
    if ((rc = upload_skel<int>(sk, d->slider_table_length, s.num_sliders, &an.slider_table_length)) != P3CS_OK) return rc;
    if ((rc = upload_skel<int>(sk, d->slider_table_offset, s.num_sliders, &an.slider_table_offset)) != P3CS_OK) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(sk->ctx->stream));
  *out_index = s.num_animations;
  s.anim[s.num_animations++] = an;
  return P3CS_OK;
}

// states: [count][num_controls]; NULL = the unblended single-control form with host_frames.
static int pose_common(p3cs_group grp, p3cs_skeleton sk, int first, int count, const int32_t *host_frames, const ControlState *states,
                       int num_controls, int frame_blend, int blend_type, const char *what) {
  int rc = check_span(grp, first, count, what);
  if (rc != P3CS_OK) return rc;
  if (sk == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "%s: skeleton is NULL", what);
  if (sk->num_palette != grp->mesh->dev.num_transforms)
    return p3cs_fail(P3CS_ERR_INVALID, "skeleton feeds %d palette slots, the mesh has %d transforms", sk->num_palette, grp->mesh->dev.num_transforms);
  if (sk->dev.num_sliders > 0 && sk->dev.num_sliders != grp->mesh->dev.num_sliders)
    return p3cs_fail(P3CS_ERR_INVALID, "skeleton drives %d sliders, the mesh has %d", sk->dev.num_sliders, grp->mesh->dev.num_sliders);
  if (grp->num_forced > 0 && grp->forced_num_joints != sk->dev.num_joints)
    return p3cs_fail(P3CS_ERR_INVALID, "%s: the group's forced joints were set for a skeleton of %d joints, this one has %d", what,
                     grp->forced_num_joints, sk->dev.num_joints);
  if (count == 0) return P3CS_OK;
  DeviceGuard g(grp->mesh->ctx->device);
  cudaStream_t s = grp->mesh->ctx->stream;
  const ForcedDev forced{grp->forced_slot_dev, grp->forced_values_dev, grp->num_forced};
  const size_t words = states != nullptr ? (size_t)count * num_controls * (sizeof(ControlState) / 4) : (size_t)count;
  if ((size_t)grp->frames_capacity < words) {
    if (grp->frames_dev != nullptr) {
      CUDA_TRY(cudaStreamSynchronize(s));
      CUDA_TRY(cudaFree(grp->frames_dev));
    }
    grp->frames_dev = nullptr;
    CUDA_TRY(cudaMalloc(&grp->frames_dev, words * 4));
    grp->frames_capacity = (int)words;
  }
  CUDA_TRY(cudaMemcpyAsync(grp->frames_dev, states != nullptr ? static_cast<const void *>(states) : static_cast<const void *>(host_frames),
                           words * 4, cudaMemcpyHostToDevice, s));
  launch_pose(sk->dev, const_cast<float *>(grp->dev.palettes), sk->num_palette, const_cast<float *>(grp->dev.sliders),
              states != nullptr ? nullptr : grp->frames_dev,
              states != nullptr ? reinterpret_cast<const ControlState *>(grp->frames_dev) : nullptr, num_controls, frame_blend, blend_type,
              forced, first, count, s);
  CUDA_TRY(cudaGetLastError());
  return P3CS_OK;
}

static int check_blend_type(int blend_type, const char *what) {
  if (blend_type != P3CS_BT_LINEAR && blend_type != P3CS_BT_NORMALIZED_LINEAR && blend_type != P3CS_BT_COMPONENTWISE &&
      blend_type != P3CS_BT_COMPONENTWISE_QUAT)
    return p3cs_fail(P3CS_ERR_INVALID, "%s: blend type %d is not a PartBundle::BlendType", what, blend_type);
  return P3CS_OK;
}

extern "C" int p3cs_group_pose(p3cs_group grp, p3cs_skeleton sk, int first, int count, const int32_t *host_frames) {
  if (host_frames == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_pose: NULL argument");
  return pose_common(grp, sk, first, count, host_frames, nullptr, 0, 0, P3CS_BT_NORMALIZED_LINEAR, "p3cs_group_pose");
}

extern "C" int p3cs_group_pose_blend(p3cs_group grp, p3cs_skeleton sk, int first, int count, const int32_t *host_frames,
                                     const int32_t *host_next_frames, const float *host_fracs, int blend_type) {
  if (host_frames == nullptr || host_next_frames == nullptr || host_fracs == nullptr)
    return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_pose_blend: NULL argument");
  int rc = check_blend_type(blend_type, "p3cs_group_pose_blend");
  if (rc != P3CS_OK) return rc;
  std::vector<ControlState> states((size_t)std::max(count, 0));
  for (int i = 0; i < count; ++i) states[i] = ControlState{0, host_frames[i], host_next_frames[i], host_fracs[i], 1.0f};
  return pose_common(grp, sk, first, count, nullptr, states.data(), 1, 1, blend_type, "p3cs_group_pose_blend");
}

extern "C" int p3cs_group_pose_controls(p3cs_group grp, p3cs_skeleton sk, int first, int count, int num_controls,
                                        const p3cs_control_state *host_states, int frame_blend, int blend_type) {
  static_assert(sizeof(p3cs_control_state) == sizeof(ControlState), "p3cs_control_state layout");
  if (host_states == nullptr || num_controls < 1 || num_controls > 16)
    return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_pose_controls: bad arguments (1..16 controls)");
  int rc = check_blend_type(blend_type, "p3cs_group_pose_controls");
  if (rc != P3CS_OK) return rc;
  if (sk != nullptr)
    for (size_t i = 0; i < (size_t)std::max(count, 0) * num_controls; ++i)
      if (host_states[i].anim < 0 || host_states[i].anim >= sk->dev.num_animations)
        return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_pose_controls: state %zu plays animation %d, the skeleton holds %d", i,
                         host_states[i].anim, sk->dev.num_animations);
  return pose_common(grp, sk, first, count, nullptr, reinterpret_cast<const ControlState *>(host_states), num_controls,
                     frame_blend != 0, blend_type, "p3cs_group_pose_controls");
}

extern "C" int p3cs_group_set_forced_joints(p3cs_group grp, p3cs_skeleton sk, int num_forced, const int32_t *joints) {
  if (grp == nullptr || sk == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_forced_joints: NULL argument");
  if (num_forced < 0 || (num_forced > 0 && joints == nullptr)) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_forced_joints: bad joint list");
  const int J = sk->dev.num_joints;
  std::vector<int> slot((size_t)std::max(J, 1), -1);
  for (int f = 0; f < num_forced; ++f) {
    if (joints[f] < 0 || joints[f] >= J) return p3cs_fail(P3CS_ERR_INVALID, "forced joint %d is not a joint of the skeleton (%d joints)", joints[f], J);
    if (slot[joints[f]] >= 0) return p3cs_fail(P3CS_ERR_INVALID, "joint %d is listed twice", joints[f]);
    slot[joints[f]] = f;
  }
  DeviceGuard g(grp->mesh->ctx->device);
  cudaStream_t s = grp->mesh->ctx->stream;
  CUDA_TRY(cudaStreamSynchronize(s));  // a pose kernel in flight may still read the old tables
  if (grp->forced_slot_dev != nullptr) CUDA_TRY(cudaFree(grp->forced_slot_dev));
  if (grp->forced_values_dev != nullptr) CUDA_TRY(cudaFree(grp->forced_values_dev));
  grp->forced_slot_dev = nullptr;
  grp->forced_values_dev = nullptr;
  grp->num_forced = 0;
  grp->forced_num_joints = J;
  if (num_forced == 0) return P3CS_OK;
  CUDA_TRY(cudaMalloc(&grp->forced_slot_dev, slot.size() * sizeof(int)));
  CUDA_TRY(cudaMemcpyAsync(grp->forced_slot_dev, slot.data(), slot.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  // every instance starts with identity values
  std::vector<float> ident((size_t)grp->n * num_forced * 16, 0.0f);
  for (size_t i = 0; i < (size_t)grp->n * num_forced; ++i) ident[i * 16] = ident[i * 16 + 5] = ident[i * 16 + 10] = ident[i * 16 + 15] = 1.0f;
  CUDA_TRY(cudaMalloc(&grp->forced_values_dev, std::max<size_t>(ident.size(), 1) * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(grp->forced_values_dev, ident.data(), ident.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  grp->num_forced = num_forced;
  return P3CS_OK;
}

extern "C" int p3cs_group_set_forced_values(p3cs_group grp, int first, int count, const float *host_values) {
  int rc = check_span(grp, first, count, "p3cs_group_set_forced_values");
  if (rc != P3CS_OK) return rc;
  if (grp->num_forced == 0) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_forced_values: the group has no forced joints");
  if (count == 0) return P3CS_OK;
  if (host_values == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_set_forced_values: values are NULL");
  DeviceGuard g(grp->mesh->ctx->device);
  cudaStream_t s = grp->mesh->ctx->stream;
  const size_t per = (size_t)grp->num_forced * 16;
  CUDA_TRY(cudaMemcpyAsync(grp->forced_values_dev + (size_t)first * per, host_values, (size_t)count * per * sizeof(float), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));  // the host buffer is the caller's again on return
  return P3CS_OK;
}

extern "C" int p3cs_group_read_palettes(p3cs_group grp, int first, int count, float *host_dst) {
  int rc = check_span(grp, first, count, "p3cs_group_read_palettes");
  if (rc != P3CS_OK) return rc;
  if (host_dst == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_group_read_palettes: destination is NULL");
  DeviceGuard g(grp->mesh->ctx->device);
  const size_t per = (size_t)grp->mesh->dev.num_transforms * 16;
  if (per * count == 0) return P3CS_OK;
  CUDA_TRY(cudaMemcpyAsync(host_dst, grp->dev.palettes + per * first, per * count * 4, cudaMemcpyDeviceToHost, grp->mesh->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(grp->mesh->ctx->stream));
  return P3CS_OK;
}

// ------------------------------------------------------------------ plugin table
extern "C" const p3cs_api *p3cs_get_api(void) {
  static const p3cs_api api = {
      P3CS_API_VERSION,        p3cs_device_count,  p3cs_last_error,        p3cs_context_create, p3cs_context_destroy,
      p3cs_context_sync,       p3cs_mesh_create,   p3cs_mesh_destroy,      p3cs_group_create,   p3cs_group_destroy,
      p3cs_group_set_palettes, p3cs_group_set_sliders, p3cs_group_animate, p3cs_group_read_output, p3cs_animate_vertices,
      p3cs_host_register,      p3cs_host_unregister, p3cs_animate_groups,
      p3cs_group_set_output_pointers, p3cs_group_animate_instances};
  return &api;
}
