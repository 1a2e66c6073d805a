// p3cs_internal.h -- what the host-side translation units of libp3cudaskin.so share: the objects behind the
// opaque handles of include/p3cudaskin.h, error reporting and small helpers.  Not installed.
#pragma once
#include "../../include/p3cudaskin.h"
#include "p3cs_kernels.cuh"

#include <cuda_runtime.h>

#include <cstddef>
#include <vector>

// Records the message p3cs_last_error() returns (per thread) and hands `code` back.
int p3cs_fail(int code, const char *fmt, ...);

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t err__ = (expr);                                                                 \
    if (err__ != cudaSuccess)                                                                   \
      return p3cs_fail(err__ == cudaErrorMemoryAllocation ? P3CS_ERR_NOMEM : P3CS_ERR_CUDA,     \
                       "%s failed: %s", #expr, cudaGetErrorString(err__));                      \
  } while (0)

// ------------------------------------------------------------------ objects
struct p3cs_context_s {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  // fork / join plumbing of p3cs_animate_groups: independent groups run on side streams
  static constexpr int kSideStreams = 3;
  cudaStream_t side[kSideStreams] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork_ev = nullptr;
  bool side_ready = false;  // every side stream / event above exists (ensure_side_streams)
  cudaEvent_t join_ev[kSideStreams] = {nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> pipe_ev;  // chunk hand-off events of p3cs_group_animate_host
  // p3cs_animate_vertices: the destinations last seen are remembered with their device address (cudaPointerGetAttributes
  // costs a microsecond per array and call)
  const void *seen_host[4] = {nullptr, nullptr, nullptr, nullptr};
  void *seen_dev[4] = {nullptr, nullptr, nullptr, nullptr};
  // p3cs_animate_groups: the fork / join of a set of groups captured once as a CUDA graph and replayed
  struct GroupsGraph {
    std::vector<p3cs_group> groups;
    std::vector<unsigned> epochs;  // p3cs_group_s::epoch of every group at capture time
    std::vector<int> launches;
    cudaGraphExec_t exec = nullptr;
  };
  std::vector<GroupsGraph> graphs;
};

struct p3cs_mesh_s {
  p3cs_context ctx = nullptr;
  p3cs::MeshDev dev{};
  std::vector<void *> allocations;
  bool use_strip = true;  // fused strip kernel (p3cs_strip.cu); false: blend + skin kernels
  std::vector<int> out_of_array;  // source array -> output index (-1: not an output)
};

struct p3cs_group_s {
  p3cs_mesh mesh = nullptr;
  int n = 0;
  int chunk = 1;  // instances per blend/skin launch pair (records scratch stays L2-sized)
  p3cs::GroupDev dev{};
  std::vector<void *> allocations;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> prof;  // event pairs per launch when profiling
  std::vector<int> prof_kind;     // 0 blend, 1 skin, per pair
  bool profiling = false;
  float *bounds_buf = nullptr;  // [n][8] animated tight bounds (p3cs_group_enable_bounds)
  void **out_ptrs_dev[p3cs::kMaxOutputs] = {nullptr, nullptr, nullptr, nullptr};  // p3cs_group_set_output_pointers
  int *list_dev = nullptr;      // instance list of the last p3cs_group_animate_instances
  int list_capacity = 0;
  int *frames_dev = nullptr;  // per-instance animation frames of the last p3cs_group_pose
  int frames_capacity = 0;
  uint32_t *bounds_rows_dev = nullptr;  // p3cs_group_set_bounds_rows
  int *forced_slot_dev = nullptr;      // p3cs_group_set_forced_joints: [joints of the skeleton] slot or -1
  float *forced_values_dev = nullptr;  // [n][num_forced][16]
  int num_forced = 0, forced_num_joints = 0;
  p3cs_timings last{};
  bool timed = false;
  unsigned epoch = 0;  // bumped whenever what a launch of this group does changes (bounds on / off, profiling)
};

namespace p3cs {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <class T>
int upload(p3cs_mesh mesh, const T *host, size_t count, const T **out) {
  *out = nullptr;
  if (count == 0) count = 1;  // keep pointers non-null
  void *d = nullptr;
  // +16: the strip kernel's bulk copies round the last tile of an array up to 16 bytes
  CUDA_TRY(cudaMalloc(&d, count * sizeof(T) + 16));
  mesh->allocations.push_back(d);
  if (host != nullptr) CUDA_TRY(cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, mesh->ctx->stream));
  *out = static_cast<const T *>(d);
  return P3CS_OK;
}

inline int numeric_bytes(int nt) {
  switch (nt) {
  case P3CS_NT_UINT8: case P3CS_NT_INT8: return 1;
  case P3CS_NT_UINT16: case P3CS_NT_INT16: return 2;
  case P3CS_NT_FLOAT64: return 8;
  default: return 4;
  }
}

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace p3cs

