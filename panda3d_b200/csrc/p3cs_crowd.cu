// p3cs_crowd.cu -- one process, several GPUs: a crowd of instances of one mesh spread over the devices of a box, every
// device's kernel storing its finished rows straight into ONE buffer on the rendering GPU (peer-mapped over NVLink,
// the same TMA bulk stores as on one device).  The C face of what panda3d_b200/crowd.py does across processes; host
// C++ over the single-device entry points of this library, no collective library involved.
//
// The split is SINK-AWARE: all rows end on devices[0], whose inbound NVLink carries at most ~720 GB/s (measured on
// 8 x B200, 0.93 of the 770 GB/s peer-copy reference), while its own rows cost it only compute.  devices[0] keeps the
// share f that makes f * C equal to (1 - f) * max(G, C / (n - 1)) -- C: one GPU animating the whole crowd, G: the whole
// crowd crossing the link.  An even deal is link-bound and slower than one GPU (5.84 ms against 1.46 ms).
#include "p3cs_internal.h"

#include <algorithm>
#include <new>
#include <vector>

using namespace p3cs;

struct p3cs_crowd_s {
  int num_instances = 0;
  std::vector<int> devices;
  std::vector<p3cs_context> ctx;
  std::vector<p3cs_mesh> mesh;
  std::vector<p3cs_group> group;   // NULL for a device without instances
  std::vector<int> first, count;
  std::vector<void *> out;         // per output array: [num_instances][pitch] on devices[0]
  std::vector<size_t> pitch;
  std::vector<cudaEvent_t> ev0, ev1;
  int num_outputs = 0, num_transforms = 0, num_sliders = 0;
  double bytes_per_instance = 0.0;
};

static void split_counts(int n, int ndev, double share0, std::vector<int> &first, std::vector<int> &count) {
  first.assign(ndev, 0);
  count.assign(ndev, 0);
  if (ndev == 1) {
    count[0] = n;
    return;
  }
  share0 = std::min(1.0, std::max(1.0 / ndev, share0));
  count[0] = std::min(n, (int)(share0 * n + 0.5));
  const int rest = n - count[0];
  for (int d = 1; d < ndev; ++d) count[d] = rest / (ndev - 1) + (d - 1 < rest % (ndev - 1) ? 1 : 0);
  for (int d = 1; d < ndev; ++d) first[d] = first[d - 1] + count[d - 1];
}

// f = R / (C + R), R = max(G, C / (n - 1))
static double sink_share(double compute_ms_all, double gather_ms_all, int ndev) {
  if (ndev <= 1) return 1.0;
  const double remote = std::max(gather_ms_all, compute_ms_all / (ndev - 1));
  return remote / (compute_ms_all + remote);
}

static int build_groups(p3cs_crowd c) {
  for (size_t d = 0; d < c->devices.size(); ++d) {
    if (c->group[d] != nullptr) {
      p3cs_group_destroy(c->group[d]);
      c->group[d] = nullptr;
    }
    if (c->count[d] == 0) continue;
    std::vector<void *> outs(c->num_outputs);
    for (int o = 0; o < c->num_outputs; ++o) outs[o] = static_cast<uint8_t *>(c->out[o]) + (size_t)c->first[d] * c->pitch[o];
    p3cs_group_buffers bufs{nullptr, nullptr, outs.data()};
    const int rc = p3cs_group_create(c->mesh[d], c->count[d], &bufs, &c->group[d]);
    if (rc != P3CS_OK) return rc;
  }
  return P3CS_OK;
}

extern "C" int p3cs_crowd_create(const p3cs_crowd_desc *d, const p3cs_mesh_desc *mesh, p3cs_crowd *out) {
  if (out == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_create: out is NULL");
  *out = nullptr;
  if (d == nullptr || mesh == nullptr || d->struct_size != sizeof(p3cs_crowd_desc))
    return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_create: bad descriptor");
  if (d->num_devices < 1 || d->devices == nullptr || d->num_instances < 1) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_create: devices / instances");
  for (int i = 0; i < d->num_devices; ++i)
    for (int j = 0; j < i; ++j)
      if (d->devices[i] == d->devices[j]) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_create: device %d listed twice", d->devices[i]);
  p3cs_crowd c = new (std::nothrow) p3cs_crowd_s;
  if (c == nullptr) return p3cs_fail(P3CS_ERR_NOMEM, "out of host memory");
  const int ndev = d->num_devices;
  c->num_instances = d->num_instances;
  c->devices.assign(d->devices, d->devices + ndev);
  c->ctx.assign(ndev, nullptr);
  c->mesh.assign(ndev, nullptr);
  c->group.assign(ndev, nullptr);
  c->ev0.assign(ndev, nullptr);
  c->ev1.assign(ndev, nullptr);
  int rc = P3CS_OK;
  auto bail = [&](int code) {
    p3cs_crowd_destroy(c);
    return code;
  };
  for (int i = 0; i < ndev; ++i) {
    if ((rc = p3cs_context_create(c->devices[i], nullptr, &c->ctx[i])) != P3CS_OK) return bail(rc);
    if (i > 0 && (rc = p3cs_context_enable_peer_access(c->ctx[i], c->devices[0])) != P3CS_OK) return bail(rc);
    if ((rc = p3cs_mesh_create(c->ctx[i], mesh, &c->mesh[i])) != P3CS_OK) return bail(rc);
    DeviceGuard g(c->devices[i]);
    if (cudaEventCreate(&c->ev0[i]) != cudaSuccess || cudaEventCreate(&c->ev1[i]) != cudaSuccess)
      return bail(p3cs_fail(P3CS_ERR_CUDA, "cudaEventCreate failed"));
  }
  const MeshDev &m = c->mesh[0]->dev;
  c->num_outputs = m.num_outputs;
  c->num_transforms = m.num_transforms;
  c->num_sliders = m.num_sliders;
  c->out.assign(c->num_outputs, nullptr);
  c->pitch.assign(c->num_outputs, 0);
  for (int o = 0; o < c->num_outputs; ++o) {
    c->pitch[o] = round_up((size_t)m.num_rows * m.stride[o], 256);
    c->bytes_per_instance += (double)c->pitch[o];
    if ((rc = p3cs_device_alloc(c->ctx[0], c->pitch[o] * (size_t)c->num_instances, &c->out[o])) != P3CS_OK) return bail(rc);
  }
  // the first split: planned from rates (a kernel that writes ~3.9 TB/s of rows, a link of `link_gbs`); p3cs_crowd_tune
  // replaces it by one from measured times
  double share = d->root_share;
  if (share <= 0.0) {
    const double link = d->link_gbs > 0.0f ? d->link_gbs : 720.0;
    const double total = c->bytes_per_instance * c->num_instances;
    share = sink_share(total / 3.9e12 * 1e3, total / (link * 1e9) * 1e3, ndev);
  }
  split_counts(c->num_instances, ndev, share, c->first, c->count);
  if ((rc = build_groups(c)) != P3CS_OK) return bail(rc);
  *out = c;
  return P3CS_OK;
}

extern "C" int p3cs_crowd_destroy(p3cs_crowd c) {
  if (c == nullptr) return P3CS_OK;
  for (size_t d = 0; d < c->devices.size(); ++d) {
    if (c->group[d] != nullptr) p3cs_group_destroy(c->group[d]);
    if (c->mesh[d] != nullptr) p3cs_mesh_destroy(c->mesh[d]);
  }
  if (!c->ctx.empty() && c->ctx[0] != nullptr)
    for (void *p : c->out)
      if (p != nullptr) p3cs_device_free(c->ctx[0], p);
  for (size_t d = 0; d < c->devices.size(); ++d) {
    if (c->ctx[d] == nullptr) continue;
    {
      DeviceGuard g(c->devices[d]);
      if (c->ev0[d] != nullptr) cudaEventDestroy(c->ev0[d]);
      if (c->ev1[d] != nullptr) cudaEventDestroy(c->ev1[d]);
    }
    p3cs_context_destroy(c->ctx[d]);
  }
  delete c;
  return P3CS_OK;
}

extern "C" int p3cs_crowd_num_devices(p3cs_crowd c) { return c != nullptr ? (int)c->devices.size() : 0; }

extern "C" int p3cs_crowd_range(p3cs_crowd c, int device_index, int32_t *first, int32_t *count) {
  if (c == nullptr || device_index < 0 || device_index >= (int)c->devices.size() || first == nullptr || count == nullptr)
    return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_range: bad arguments");
  *first = c->first[device_index];
  *count = c->count[device_index];
  return P3CS_OK;
}

// [first, first + count) in crowd numbering -> the devices that hold the pieces
template <class F>
static int for_each_piece(p3cs_crowd c, int first, int count, const char *what, F f) {
  if (c == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "%s: crowd is NULL", what);
  if (first < 0 || count < 0 || first + count > c->num_instances)
    return p3cs_fail(P3CS_ERR_INVALID, "%s: instances [%d,%d) outside the crowd of %d", what, first, first + count, c->num_instances);
  for (size_t d = 0; d < c->devices.size(); ++d) {
    const int lo = std::max(first, c->first[d]), hi = std::min(first + count, c->first[d] + c->count[d]);
    if (lo >= hi) continue;
    const int rc = f((int)d, lo - c->first[d], hi - lo, lo - first);
    if (rc != P3CS_OK) return rc;
  }
  return P3CS_OK;
}

extern "C" int p3cs_crowd_set_palettes(p3cs_crowd c, int first, int count, const float *host) {
  if (host == nullptr && count > 0) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_set_palettes: host pointer is NULL");
  return for_each_piece(c, first, count, "p3cs_crowd_set_palettes", [&](int d, int local_first, int n, int offset) {
    return p3cs_group_set_palettes(c->group[d], local_first, n, host + (size_t)offset * c->num_transforms * 16);
  });
}

extern "C" int p3cs_crowd_set_sliders(p3cs_crowd c, int first, int count, const float *host) {
  if (c != nullptr && c->num_sliders == 0) return P3CS_OK;
  if (host == nullptr && count > 0) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_set_sliders: host pointer is NULL");
  return for_each_piece(c, first, count, "p3cs_crowd_set_sliders", [&](int d, int local_first, int n, int offset) {
    return p3cs_group_set_sliders(c->group[d], local_first, n, host + (size_t)offset * c->num_sliders);
  });
}

extern "C" int p3cs_crowd_animate(p3cs_crowd c) {
  if (c == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_animate: crowd is NULL");
  // remote devices first: their rows have the longer way
  for (int pass = 0; pass < 2; ++pass)
    for (size_t d = 0; d < c->devices.size(); ++d) {
      if ((pass == 0) == (d == 0) || c->group[d] == nullptr) continue;
      DeviceGuard g(c->devices[d]);
      cudaEventRecord(c->ev0[d], c->ctx[d]->stream);
      const int rc = p3cs_group_animate(c->group[d]);
      cudaEventRecord(c->ev1[d], c->ctx[d]->stream);
      if (rc != P3CS_OK) return rc;
    }
  return P3CS_OK;
}

extern "C" int p3cs_crowd_sync(p3cs_crowd c) {
  if (c == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_sync: crowd is NULL");
  for (size_t d = 0; d < c->devices.size(); ++d) {
    const int rc = p3cs_context_sync(c->ctx[d]);
    if (rc != P3CS_OK) return rc;
  }
  return P3CS_OK;
}

extern "C" int p3cs_crowd_last_ms(p3cs_crowd c, float *per_device_ms) {
  if (c == nullptr || per_device_ms == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_last_ms: NULL argument");
  for (size_t d = 0; d < c->devices.size(); ++d) {
    per_device_ms[d] = 0.0f;
    if (c->group[d] == nullptr) continue;
    DeviceGuard g(c->devices[d]);
    CUDA_TRY(cudaEventSynchronize(c->ev1[d]));
    CUDA_TRY(cudaEventElapsedTime(&per_device_ms[d], c->ev0[d], c->ev1[d]));
  }
  return P3CS_OK;
}

extern "C" void *p3cs_crowd_output_device(p3cs_crowd c, int o) { return (c != nullptr && o >= 0 && o < c->num_outputs) ? c->out[o] : nullptr; }
extern "C" size_t p3cs_crowd_output_pitch(p3cs_crowd c, int o) { return (c != nullptr && o >= 0 && o < c->num_outputs) ? c->pitch[o] : 0; }

extern "C" int p3cs_crowd_read_output(p3cs_crowd c, int o, int first, int count, void *host_dst) {
  if (c == nullptr || o < 0 || o >= c->num_outputs || host_dst == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_read_output: bad arguments");
  if (first < 0 || count < 0 || first + count > c->num_instances) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_read_output: instances outside the crowd");
  int rc = p3cs_crowd_sync(c);
  if (rc != P3CS_OK) return rc;
  const MeshDev &m = c->mesh[0]->dev;
  const size_t row_bytes = (size_t)m.num_rows * m.stride[o];
  DeviceGuard g(c->devices[0]);
  CUDA_TRY(cudaMemcpy2D(host_dst, row_bytes, static_cast<uint8_t *>(c->out[o]) + (size_t)first * c->pitch[o], c->pitch[o], row_bytes, count,
                        cudaMemcpyDeviceToHost));
  return P3CS_OK;
}

// Measures one animate pass per step with the palettes the groups hold and re-splits the crowd from the times seen:
// C = root time scaled to the whole crowd, G = slowest remote time scaled likewise.  Palettes and slider values must
// be set again afterwards (the instances change device).
extern "C" int p3cs_crowd_tune(p3cs_crowd c, int steps) {
  if (c == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_crowd_tune: crowd is NULL");
  const int ndev = (int)c->devices.size();
  if (ndev == 1 || c->count[0] == 0 || c->count[0] == c->num_instances) return P3CS_OK;
  steps = std::max(1, steps);
  std::vector<float> ms(ndev), acc(ndev, 0.0f);
  int rc;
  for (int s = 0; s < steps + 1; ++s) {
    if ((rc = p3cs_crowd_animate(c)) != P3CS_OK || (rc = p3cs_crowd_sync(c)) != P3CS_OK || (rc = p3cs_crowd_last_ms(c, ms.data())) != P3CS_OK) return rc;
    if (s > 0)  // the first pass warms up
      for (int d = 0; d < ndev; ++d) acc[d] += ms[d];
  }
  const double root_ms = acc[0] / steps;
  double remote_ms = 0.0;
  for (int d = 1; d < ndev; ++d) remote_ms = std::max(remote_ms, (double)acc[d] / steps);
  if (root_ms <= 0.0 || remote_ms <= 0.0) return P3CS_OK;
  const double compute_all = root_ms * c->num_instances / c->count[0];
  const double gather_all = remote_ms * c->num_instances / (c->num_instances - c->count[0]);
  std::vector<int> first, count;
  split_counts(c->num_instances, ndev, sink_share(compute_all, gather_all, ndev), first, count);
  if (count == c->count) return P3CS_OK;
  c->first = first;
  c->count = count;
  return build_groups(c);
}
