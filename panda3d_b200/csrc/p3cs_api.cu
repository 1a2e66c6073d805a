// p3cs_api.cu -- the C-ABI of libp3cudaskin.so (include/p3cudaskin.h): contexts, the
// one-time flattening/upload of a GeomVertexData's static half, instance groups and the
// per-frame animate call.  Host C++; the kernels are in p3cs_kernels.cu.
#include "../../include/p3cudaskin.h"
#include "p3cs_kernels.cuh"

#include <cuda.h>  // driver API types only; entry points come from cudaGetDriverEntryPoint
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace p3cs;

// ------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t err__ = (expr);                                                                 \
    if (err__ != cudaSuccess)                                                                   \
      return fail(err__ == cudaErrorMemoryAllocation ? P3CS_ERR_NOMEM : P3CS_ERR_CUDA,          \
                  "%s failed: %s", #expr, cudaGetErrorString(err__));                           \
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
  cudaEvent_t join_ev[kSideStreams] = {nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> pipe_ev;  // chunk hand-off events of p3cs_group_animate_host
};

struct p3cs_mesh_s {
  p3cs_context ctx = nullptr;
  MeshDev dev{};
  std::vector<void *> allocations;
  bool use_strip = true;  // fused strip kernel (p3cs_strip.cu); false: blend + skin kernels
  std::vector<int> out_of_array;  // source array -> output index (-1: not an output)
};

struct p3cs_group_s {
  p3cs_mesh mesh = nullptr;
  int n = 0;
  int chunk = 1;  // instances per blend/skin launch pair (records scratch stays L2-sized)
  GroupDev dev{};
  std::vector<void *> allocations;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> prof;  // event pairs per launch when profiling
  std::vector<int> prof_kind;     // 0 blend, 1 skin, per pair
  bool profiling = false;
  float *bounds_buf = nullptr;  // [n][8] animated tight bounds (p3cs_group_enable_bounds)
  int *list_dev = nullptr;      // instance list of the last p3cs_group_animate_instances
  int list_capacity = 0;
  int *frames_dev = nullptr;  // per-instance animation frames of the last p3cs_group_pose
  int frames_capacity = 0;
  p3cs_timings last{};
  bool timed = false;
};

namespace {

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

int numeric_bytes(int nt) {
  switch (nt) {
  case P3CS_NT_UINT8: case P3CS_NT_INT8: return 1;
  case P3CS_NT_UINT16: case P3CS_NT_INT16: return 2;
  case P3CS_NT_FLOAT64: return 8;
  default: return 4;
  }
}

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

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
  if (out == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_context_create: out is NULL");
  *out = nullptr;
  int n = p3cs_device_count();
  if (n <= 0) return fail(P3CS_ERR_NO_DEVICE, "no CUDA device available (this backend has no CPU fallback)");
  if (device < 0 || device >= n) return fail(P3CS_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
  CUDA_TRY(cudaSetDevice(device));
  p3cs_context ctx = new (std::nothrow) p3cs_context_s;
  if (ctx == nullptr) return fail(P3CS_ERR_NOMEM, "out of host memory");
  ctx->device = device;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (stream != nullptr) {
    ctx->stream = static_cast<cudaStream_t>(stream);
  } else {
    cudaError_t err = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) {
      delete ctx;
      return fail(P3CS_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(err));
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
  delete ctx;
  return P3CS_OK;
}

extern "C" int p3cs_context_sync(p3cs_context ctx) {
  if (ctx == nullptr) return fail(P3CS_ERR_INVALID, "context is NULL");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return P3CS_OK;
}

extern "C" int p3cs_context_enable_peer_access(p3cs_context ctx, int peer_device) {
  if (ctx == nullptr) return fail(P3CS_ERR_INVALID, "context is NULL");
  if (peer_device == ctx->device) return P3CS_OK;
  DeviceGuard g(ctx->device);
  int can = 0;
  CUDA_TRY(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
  if (!can) return fail(P3CS_ERR_UNSUPPORTED, "device %d cannot access device %d", ctx->device, peer_device);
  cudaError_t err = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (err == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return P3CS_OK;
  }
  if (err != cudaSuccess) return fail(P3CS_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(err));
  return P3CS_OK;
}

extern "C" int p3cs_device_alloc(p3cs_context ctx, size_t bytes, void **out) {
  if (ctx == nullptr || out == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_device_alloc: NULL argument");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaMalloc(out, bytes > 0 ? bytes : 256));
  return P3CS_OK;
}

extern "C" int p3cs_device_free(p3cs_context ctx, void *ptr) {
  if (ctx == nullptr) return fail(P3CS_ERR_INVALID, "context is NULL");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaFree(ptr));
  return P3CS_OK;
}

extern "C" int p3cs_device_read(p3cs_context ctx, const void *device_src, void *host_dst, size_t bytes) {
  if (ctx == nullptr || device_src == nullptr || host_dst == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_device_read: NULL argument");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return P3CS_OK;
}

extern "C" int p3cs_ipc_export(p3cs_context ctx, void *ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (ctx == nullptr || ptr == nullptr || handle == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_ipc_export: NULL argument");
  DeviceGuard g(ctx->device);
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, 64);
  return P3CS_OK;
}

extern "C" int p3cs_ipc_open(p3cs_context ctx, const unsigned char handle[64], void **out) {
  if (ctx == nullptr || handle == nullptr || out == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_ipc_open: NULL argument");
  DeviceGuard g(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return P3CS_OK;
}

extern "C" int p3cs_ipc_close(p3cs_context ctx, void *ptr) {
  if (ctx == nullptr) return fail(P3CS_ERR_INVALID, "context is NULL");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return P3CS_OK;
}

// ---- "next" row f2: memory the renderer can import (Vulkan VK_KHR_external_memory_fd / OpenGL
// GL_EXT_memory_object_fd).  cudaMalloc memory cannot be exported, so this goes through the driver's virtual
// memory management API; its entry points are fetched at run time (cudaGetDriverEntryPoint), the library keeps
// no link-time dependency on libcuda.
namespace {

struct VmmApi {
  CUresult (*get_granularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*reserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*address_free)(CUdeviceptr, size_t) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*export_handle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*import_handle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
  bool ok = false;
};

template <class F>
bool driver_entry(const char *name, F *out) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || fn == nullptr) {
    cudaGetLastError();
    return false;
  }
  *out = reinterpret_cast<F>(fn);
  return true;
}

const VmmApi &vmm() {
  static VmmApi api = [] {
    VmmApi a;
    a.ok = driver_entry("cuMemGetAllocationGranularity", &a.get_granularity) && driver_entry("cuMemCreate", &a.create) &&
           driver_entry("cuMemRelease", &a.release) && driver_entry("cuMemAddressReserve", &a.reserve) &&
           driver_entry("cuMemAddressFree", &a.address_free) && driver_entry("cuMemMap", &a.map) &&
           driver_entry("cuMemUnmap", &a.unmap) && driver_entry("cuMemSetAccess", &a.set_access) &&
           driver_entry("cuMemExportToShareableHandle", &a.export_handle) &&
           driver_entry("cuMemImportFromShareableHandle", &a.import_handle);
    return a;
  }();
  return api;
}

struct Shareable {
  CUmemGenericAllocationHandle handle;
  size_t bytes;
};
std::mutex g_shareable_lock;
std::map<void *, Shareable> g_shareable;

CUmemAllocationProp shareable_prop(int device) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

int map_shareable(p3cs_context ctx, CUmemGenericAllocationHandle handle, size_t bytes, void **out) {
  const VmmApi &a = vmm();
  CUdeviceptr ptr = 0;
  CUresult r = a.reserve(&ptr, bytes, 0, 0, 0);
  if (r != CUDA_SUCCESS) return fail(P3CS_ERR_CUDA, "cuMemAddressReserve failed (%d)", (int)r);
  r = a.map(ptr, bytes, 0, handle, 0);
  if (r != CUDA_SUCCESS) {
    a.address_free(ptr, bytes);
    return fail(P3CS_ERR_CUDA, "cuMemMap failed (%d)", (int)r);
  }
  CUmemAccessDesc access;
  memset(&access, 0, sizeof(access));
  access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  access.location.id = ctx->device;
  access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = a.set_access(ptr, bytes, &access, 1);
  if (r != CUDA_SUCCESS) {
    a.unmap(ptr, bytes);
    a.address_free(ptr, bytes);
    return fail(P3CS_ERR_CUDA, "cuMemSetAccess failed (%d)", (int)r);
  }
  *out = reinterpret_cast<void *>(ptr);
  std::lock_guard<std::mutex> lock(g_shareable_lock);
  g_shareable[*out] = Shareable{handle, bytes};
  return P3CS_OK;
}

}  // namespace

extern "C" int p3cs_shareable_alloc(p3cs_context ctx, size_t bytes, void **out_ptr, int *out_fd, size_t *out_bytes) {
  if (ctx == nullptr || out_ptr == nullptr || out_fd == nullptr || bytes == 0) return fail(P3CS_ERR_INVALID, "p3cs_shareable_alloc: bad arguments");
  *out_ptr = nullptr;
  *out_fd = -1;
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaFree(nullptr));  // make sure the primary context exists
  const VmmApi &a = vmm();
  if (!a.ok) return fail(P3CS_ERR_UNSUPPORTED, "the CUDA driver does not expose the virtual memory management API");
  const CUmemAllocationProp prop = shareable_prop(ctx->device);
  size_t gran = 0;
  CUresult r = a.get_granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
  if (r != CUDA_SUCCESS || gran == 0) return fail(P3CS_ERR_CUDA, "cuMemGetAllocationGranularity failed (%d)", (int)r);
  const size_t size = round_up(bytes, gran);
  CUmemGenericAllocationHandle handle;
  r = a.create(&handle, size, &prop, 0);
  if (r != CUDA_SUCCESS) return fail(r == CUDA_ERROR_OUT_OF_MEMORY ? P3CS_ERR_NOMEM : P3CS_ERR_CUDA, "cuMemCreate(%zu bytes) failed (%d)", size, (int)r);
  int fd = -1;
  r = a.export_handle(&fd, handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    a.release(handle);
    return fail(P3CS_ERR_CUDA, "cuMemExportToShareableHandle failed (%d)", (int)r);
  }
  int rc = map_shareable(ctx, handle, size, out_ptr);
  if (rc != P3CS_OK) {
    close(fd);
    a.release(handle);
    return rc;
  }
  *out_fd = fd;
  if (out_bytes != nullptr) *out_bytes = size;
  return P3CS_OK;
}

extern "C" int p3cs_shareable_import(p3cs_context ctx, int fd, size_t bytes, void **out_ptr) {
  if (ctx == nullptr || out_ptr == nullptr || fd < 0 || bytes == 0) return fail(P3CS_ERR_INVALID, "p3cs_shareable_import: bad arguments");
  *out_ptr = nullptr;
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaFree(nullptr));
  const VmmApi &a = vmm();
  if (!a.ok) return fail(P3CS_ERR_UNSUPPORTED, "the CUDA driver does not expose the virtual memory management API");
  CUmemGenericAllocationHandle handle;
  CUresult r = a.import_handle(&handle, reinterpret_cast<void *>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (r != CUDA_SUCCESS) return fail(P3CS_ERR_CUDA, "cuMemImportFromShareableHandle failed (%d)", (int)r);
  int rc = map_shareable(ctx, handle, bytes, out_ptr);
  if (rc != P3CS_OK) a.release(handle);
  return rc;
}

extern "C" int p3cs_shareable_free(p3cs_context ctx, void *ptr) {
  if (ctx == nullptr || ptr == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_shareable_free: bad arguments");
  DeviceGuard g(ctx->device);
  Shareable sh;
  {
    std::lock_guard<std::mutex> lock(g_shareable_lock);
    auto it = g_shareable.find(ptr);
    if (it == g_shareable.end()) return fail(P3CS_ERR_INVALID, "p3cs_shareable_free: not a shareable allocation");
    sh = it->second;
    g_shareable.erase(it);
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  const VmmApi &a = vmm();
  const CUdeviceptr p = reinterpret_cast<CUdeviceptr>(ptr);
  a.unmap(p, sh.bytes);
  a.address_free(p, sh.bytes);
  a.release(sh.handle);
  return P3CS_OK;
}

extern "C" void *p3cs_context_stream(p3cs_context ctx) { return ctx ? ctx->stream : nullptr; }

// ------------------------------------------------------------------ mesh
static int check_column(const p3cs_mesh_desc *d, const p3cs_column_desc &c, const char *what) {
  if (c.array < 0 || c.array >= d->num_arrays) return fail(P3CS_ERR_INVALID, "%s: array %d out of range", what, c.array);
  const int stride = d->arrays[c.array].stride;
  if (c.start < 0 || c.num_values < 1 || c.start + c.num_values * numeric_bytes(c.numeric_type) > stride)
    return fail(P3CS_ERR_INVALID, "%s: column [%d,+%d x %d) exceeds stride %d", what, c.start, c.num_values,
                numeric_bytes(c.numeric_type), stride);
  return P3CS_OK;
}

static int check_ranges(const int32_t *ranges, int count, int num_rows, const char *what) {
  int prev_end = 0;
  for (int i = 0; i < count; ++i) {
    const int b = ranges[2 * i], e = ranges[2 * i + 1];
    if (b < prev_end || e <= b || e > num_rows)
      return fail(P3CS_ERR_INVALID, "%s: subrange %d = [%d,%d) is not a sorted, disjoint, non-empty range within %d rows", what, i, b, e, num_rows);
    prev_end = e;
  }
  return P3CS_OK;
}

extern "C" int p3cs_mesh_create(p3cs_context ctx, const p3cs_mesh_desc *d, p3cs_mesh *out) {
  if (out == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_mesh_create: out is NULL");
  *out = nullptr;
  if (ctx == nullptr || d == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_mesh_create: NULL argument");
  if (d->struct_size != sizeof(p3cs_mesh_desc))
    return fail(P3CS_ERR_INVALID, "p3cs_mesh_desc::struct_size %u != %zu (header mismatch)", d->struct_size, sizeof(p3cs_mesh_desc));
  if (d->num_rows < 0 || d->num_arrays < 0 || (d->num_arrays > 0 && d->arrays == nullptr))
    return fail(P3CS_ERR_INVALID, "bad num_rows / arrays");
  int cs = d->coordinate_system == 0 ? P3CS_CS_ZUP_RIGHT : d->coordinate_system;
  if (cs < P3CS_CS_ZUP_RIGHT || cs > P3CS_CS_YUP_LEFT) return fail(P3CS_ERR_INVALID, "bad coordinate_system %d", cs);
  const int N = d->num_rows;
  DeviceGuard g(ctx->device);

  p3cs_mesh mesh = new (std::nothrow) p3cs_mesh_s;
  if (mesh == nullptr) return fail(P3CS_ERR_NOMEM, "out of host memory");
  mesh->ctx = ctx;
  MeshDev &m = mesh->dev;
  m.num_rows = N;
  m.coordinate_system = cs;
  m.num_transforms = d->num_transforms;
  m.num_blends = d->num_blends;
  m.num_sliders = d->num_sliders;
  int rc = P3CS_OK;
  auto bail = [&](int code) {
    p3cs_mesh_destroy(mesh);
    return code;
  };

  // ---- output arrays: static image uploaded once (replaces the per-frame copy_from)
  std::vector<int> out_of_array(d->num_arrays, -1);
  for (int a = 0; a < d->num_arrays; ++a) {
    const p3cs_array_desc &ad = d->arrays[a];
    if (ad.stride <= 0 || (N > 0 && ad.data == nullptr)) return bail(fail(P3CS_ERR_INVALID, "array %d: bad stride/data", a));
    if (!ad.is_output) continue;
    if (m.num_outputs >= kMaxOutputs) return bail(fail(P3CS_ERR_UNSUPPORTED, "more than %d output arrays", kMaxOutputs));
    if (ad.stride % 4 != 0) return bail(fail(P3CS_ERR_UNSUPPORTED, "output array %d: stride %d is not a multiple of 4", a, ad.stride));
    out_of_array[a] = m.num_outputs;
    m.stride[m.num_outputs] = ad.stride;
    if ((rc = upload<uint8_t>(mesh, static_cast<const uint8_t *>(ad.data), (size_t)N * ad.stride, &m.src[m.num_outputs])) != P3CS_OK) return bail(rc);
    ++m.num_outputs;
  }

  // ---- dynamic columns: skinned columns, then morph bases
  auto find_dyn = [&](int output, int start) {
    for (int c = 0; c < m.num_dyn; ++c)
      if (m.dyn[c].output == output && m.dyn[c].start == start) return c;
    return -1;
  };
  const bool has_blend = d->blend_column.array >= 0 && d->num_skin_columns > 0;
  for (int i = 0; i < d->num_skin_columns; ++i) {
    const p3cs_column_desc &c = d->skin_columns[i];
    if ((rc = check_column(d, c, "skin column")) != P3CS_OK) return bail(rc);
    if (out_of_array[c.array] < 0) return bail(fail(P3CS_ERR_INVALID, "skin column %d lives in a non-output array", i));
    if ((c.numeric_type != P3CS_NT_FLOAT32 && c.numeric_type != P3CS_NT_FLOAT64) || (c.num_values != 3 && c.num_values != 4))
      return bail(fail(P3CS_ERR_UNSUPPORTED, "skin column %d: only float32 / float64 x 3 / x 4 columns are supported (got type %d x %d)", i, c.numeric_type, c.num_values));
    if (c.start % 4 != 0) return bail(fail(P3CS_ERR_UNSUPPORTED, "skin column %d: start %d not 4-byte aligned", i, c.start));
    int skin = c.contents == P3CS_C_POINT ? 1 : c.contents == P3CS_C_VECTOR ? 2 : c.contents == P3CS_C_NORMAL ? 3 : 0;
    if (skin == 0) return bail(fail(P3CS_ERR_INVALID, "skin column %d: contents %d is not point/vector/normal", i, c.contents));
    if (m.num_dyn >= kMaxDynColumns) return bail(fail(P3CS_ERR_UNSUPPORTED, "more than %d animated columns", kMaxDynColumns));
    DynColumn &dc = m.dyn[m.num_dyn++];
    dc.output = (int16_t)out_of_array[c.array];
    dc.start = (int16_t)c.start;
    dc.num_values = (int8_t)c.num_values;
    dc.skin = (int8_t)skin;
    dc.homogeneous = 0;
    dc.f64 = c.numeric_type == P3CS_NT_FLOAT64 ? 1 : 0;
    if (c.num_values == 4) m.full_records = 1;
  }

  // column order is semantically irrelevant (columns are independent); a canonical order
  // point(s), normal(s), vector(s) lets the strip kernel pick a specialised variant
  std::stable_sort(m.dyn, m.dyn + m.num_dyn, [](const DynColumn &a, const DynColumn &b) {
    auto key = [](int skin) { return skin == 1 ? 0 : skin == 3 ? 1 : 2; };
    return key(a.skin) < key(b.skin);
  });

  // ---- transform_blend index column -> dense u16/u32
  if (d->blend_column.array >= 0) {
    const p3cs_column_desc &bc = d->blend_column;
    if ((rc = check_column(d, bc, "transform_blend column")) != P3CS_OK) return bail(rc);
    if (bc.numeric_type != P3CS_NT_UINT8 && bc.numeric_type != P3CS_NT_UINT16 && bc.numeric_type != P3CS_NT_UINT32)
      return bail(fail(P3CS_ERR_UNSUPPORTED, "transform_blend column of numeric type %d", bc.numeric_type));
    if (d->num_blends < 0 || d->num_transforms < 0 || d->blend_offsets == nullptr)
      return bail(fail(P3CS_ERR_INVALID, "blend table missing"));
    if ((rc = check_ranges(d->row_ranges, d->num_row_ranges, N, "TransformBlendTable rows")) != P3CS_OK) return bail(rc);
    const int nnz = d->blend_offsets[d->num_blends];
    if (d->blend_offsets[0] != 0) return bail(fail(P3CS_ERR_INVALID, "blend_offsets[0] != 0"));
    for (int b = 0; b < d->num_blends; ++b)
      if (d->blend_offsets[b + 1] < d->blend_offsets[b]) return bail(fail(P3CS_ERR_INVALID, "blend_offsets not monotonic at %d", b));
    for (int e = 0; e < nnz; ++e)
      if (d->blend_transform[e] < 0 || d->blend_transform[e] >= d->num_transforms)
        return bail(fail(P3CS_ERR_INVALID, "blend entry %d refers to transform %d outside the palette (%d)", e, d->blend_transform[e], d->num_transforms));

    const uint8_t *base = static_cast<const uint8_t *>(d->arrays[bc.array].data) + bc.start;
    const int istride = d->arrays[bc.array].stride;
    const bool wide = bc.numeric_type == P3CS_NT_UINT32;
    std::vector<uint16_t> idx16;
    std::vector<uint32_t> idx32;
    if (wide) idx32.resize(std::max(N, 1)); else idx16.resize(std::max(N, 1));
    for (int r = 0; r < N; ++r) {
      const uint8_t *p = base + (size_t)r * istride;
      uint32_t v;
      if (bc.numeric_type == P3CS_NT_UINT8) v = *p;
      else if (bc.numeric_type == P3CS_NT_UINT16) { uint16_t t; memcpy(&t, p, 2); v = t; }
      else memcpy(&v, p, 4);
      if (wide) idx32[r] = v; else idx16[r] = (uint16_t)v;
    }
    // every row the reference would transform must select an existing blend
    bool full = d->num_row_ranges == 1 && d->row_ranges[0] == 0 && d->row_ranges[1] == N;
    std::vector<uint32_t> mask;
    if (!full) mask.assign(((size_t)N + 31) / 32 + 1, 0u);
    for (int i = 0; i < d->num_row_ranges; ++i)
      for (int r = d->row_ranges[2 * i]; r < d->row_ranges[2 * i + 1]; ++r) {
        uint32_t v = wide ? idx32[r] : idx16[r];
        if (v >= (uint32_t)d->num_blends)
          return bail(fail(P3CS_ERR_INVALID, "row %d selects blend %u but the table has %d blends", r, v, d->num_blends));
        if (!full) mask[r >> 5] |= 1u << (r & 31);
      }
    // ---- vertex-animation-align-16 layout (gobj/config_gobj.cxx:286-300, geomVertexArrayFormat.cxx:349-373): the
    // strip kernel's kAligned modes
    {
      const DynColumn *c = m.dyn;
      bool ok = m.full_records && m.num_outputs == 1 && m.stride[0] % 16 == 0 &&
                ((m.num_dyn == 2 && c[0].skin == 1 && c[1].skin == 3) ||
                 (m.num_dyn == 4 && c[0].skin == 1 && c[1].skin == 3 && c[2].skin == 2 && c[3].skin == 2));
      for (int i = 0; ok && i < m.num_dyn; ++i) ok = c[i].num_values == 4 && !c[i].f64 && c[i].start % 16 == 0;
      for (int mi = 0; ok && mi < d->num_morphs; ++mi) {  // morphs only on those columns (no morph-only column)
        const p3cs_column_desc &b = d->morphs[mi].base;
        ok = b.array >= 0 && b.array < d->num_arrays && out_of_array[b.array] == 0 && b.num_values == 4 &&
             b.numeric_type == P3CS_NT_FLOAT32 && find_dyn(0, b.start) >= 0;
      }
      m.aligned = ok ? 1 : 0;
    }
    m.strip_rec_floats = m.aligned ? kRecAligned : m.full_records ? kRecFull : kRecCompact;
    // ---- shared-memory plan of the strip kernel.  Four CTAs per SM (the register-file limit) need <= ~55 KB each:
    // palette + records + staged tiles.  Preference: two rows per thread (512-row tiles halve the per-tile
    // bookkeeping), else 256-row tiles with three, then two buffers, then a smaller record group.
    {
      size_t slice = 0;
      for (int o = 0; o < m.num_outputs; ++o) slice += round_up((size_t)32 * m.stride[o], 128);
      const size_t rec_bytes = (size_t)m.strip_rec_floats * 4;
      const size_t fixed = 32 + (size_t)d->num_transforms * 64 + 128;
      const size_t budget = 55 * 1024;
      const int full_cap = std::min(kStripThreads, (int)((32 * 1024) / rec_bytes));  // one phase-A pass: a thread per record
      auto fits = [&](int r, int bufs, int cap) { return fixed + (size_t)cap * rec_bytes + (size_t)bufs * 8 * r * slice <= budget; };
      m.rows_per_thread = 1;
      m.tile_bufs = 3;
      m.group_record_cap = full_cap;
      if (!m.full_records || m.aligned) {
        if (fits(2, 2, full_cap)) {
          m.rows_per_thread = 2;
          m.tile_bufs = 2;
        } else if (fits(1, 3, full_cap)) {
        } else if (fits(1, 2, full_cap)) {
          m.tile_bufs = 2;
        } else {
          int cap = full_cap;
          while (cap > 128 && !fits(1, 2, cap)) cap -= 8;
          if (fits(1, 2, cap)) {
            m.tile_bufs = 2;
            m.group_record_cap = cap;
          }
        }
      }
    }
    const int kTileRows = kStripThreads * m.rows_per_thread;
    // ---- strip-kernel tables: per record group the distinct blends, per row its slot in that list.
    // Groups are grown greedily, tile by tile, while their distinct blends fit one phase-A pass
    // (one thread per record) and the shared-memory record budget.
    {
      const int nt = (N + kTileRows - 1) / kTileRows;
      const int cap = m.group_record_cap;
      std::vector<uint16_t> local(std::max(N, 1), 0xFFFFu);
      std::vector<int> rec_off(1, 0), rec_blend, ent_off(1, 0), tile_off(1, 0);
      std::vector<uint2> ent;
      std::vector<int> in_group(std::max(d->num_blends, 1), -1), seen_in_tile(std::max(d->num_blends, 1), -1),
          slot(std::max(d->num_blends, 1), 0);
      int ng = 0, max_rec = 0;
      std::vector<int> open;   // distinct blends of the open group, first-seen order
      std::vector<int> fresh;  // blends first seen in the tile being added
      auto row_live = [&](int r) { return full || ((mask[r >> 5] >> (r & 31)) & 1u); };
      auto row_blend = [&](int r) { return (int)(wide ? idx32[r] : idx16[r]); };

      // Order of the records inside a group.  Phase A runs one thread per record and gathers the record's
      // first four palette matrices with 128-bit shared-memory loads, eight lanes per wavefront; two lanes
      // collide when their joints differ but agree modulo 8 (48-byte palette rows).  The order of a group's
      // records is free (rows address them through their slot), so they are packed first-fit into
      // eight-lane cells whose k-th joints are pairwise distinct modulo 8 (or equal: a broadcast).
      auto pack_group = [&](std::vector<int> &blends) {
        const int n = (int)blends.size();
        const int cells = (n + 7) / 8;
        std::vector<int> occ((size_t)cells * 32, -1), fill(cells, 0), cell_of(n, -1);
        std::vector<int> leftover;
        int first_open = 0;
        for (int i = 0; i < n; ++i) {
          const int b = blends[i];
          const int e0 = d->blend_offsets[b], cnt = std::min(4, d->blend_offsets[b + 1] - e0);
          int placed = -1;
          for (int c = first_open; c < cells && placed < 0; ++c) {
            if (fill[c] >= 8) continue;
            bool ok = true;
            for (int k = 0; k < cnt && ok; ++k) {
              const int j = d->blend_transform[e0 + k];
              const int o = occ[(size_t)c * 32 + k * 8 + (j & 7)];
              ok = o < 0 || o == j;
            }
            if (ok) placed = c;
          }
          if (placed < 0) {
            leftover.push_back(i);
            continue;
          }
          for (int k = 0; k < cnt; ++k) {
            const int j = d->blend_transform[e0 + k];
            occ[(size_t)placed * 32 + k * 8 + (j & 7)] = j;
          }
          cell_of[i] = placed;
          ++fill[placed];
          while (first_open < cells && fill[first_open] >= 8) ++first_open;
        }
        int c = 0;
        for (int i : leftover) {  // whatever did not fit conflict-free fills the remaining lanes
          while (fill[c] >= 8) ++c;
          cell_of[i] = c;
          ++fill[c];
        }
        // dense positions: cells in order; only the last cell may be short, so move its surplus forward
        std::vector<int> count_in(cells, 0);
        for (int i = 0; i < n; ++i) ++count_in[cell_of[i]];
        for (int cc = 0, donor = cells - 1; cc < donor; ++cc)
          while (count_in[cc] < 8 && cc < donor) {
            if (count_in[donor] == 0) { --donor; continue; }
            for (int i = n - 1; i >= 0; --i)
              if (cell_of[i] == donor) { cell_of[i] = cc; break; }
            --count_in[donor];
            ++count_in[cc];
          }
        // Lane inside the cell.  Phase B reads a row's record with 64-bit loads, sixteen rows per wavefront; the
        // records of neighbouring runs (adjacent in first-seen order) collide there when their slots agree
        // modulo 16, so each record takes a free lane whose slot differs from its three predecessors' modulo 16.
        std::vector<int> slot_of(n, -1);
        std::vector<uint8_t> used((size_t)cells, 0);
        for (int i = 0; i < n; ++i) {
          const int c = cell_of[i];
          int best = -1;
          for (int pass = 0; pass < 2 && best < 0; ++pass)
            for (int p = 0; p < count_in[c]; ++p) {
              if ((used[c] >> p) & 1) continue;
              const int sl = 8 * c + p;
              bool clash = false;  // against the records of the three preceding runs (a half-warp of rows spans about that many)
              for (int back = 1; back <= 3 && back <= i; ++back) clash = clash || ((slot_of[i - back] ^ sl) & 15) == 0;
              if (pass == 1 || !clash) { best = p; break; }
            }
          used[c] |= (uint8_t)(1u << best);
          slot_of[i] = 8 * c + best;
        }
        std::vector<int> ordered(n);
        for (int i = 0; i < n; ++i) ordered[slot_of[i]] = blends[i];
        blends.swap(ordered);
      };

      auto close_group = [&](int t_end) {
        pack_group(open);
        const int n = (int)open.size();
        for (int p = 0; p < n; ++p) {
          const int b = open[p];
          slot[b] = p;
          rec_blend.push_back(b);
          for (int e = d->blend_offsets[b]; e < d->blend_offsets[b + 1]; ++e) {
            uint32_t wbits;
            memcpy(&wbits, &d->blend_weight[e], 4);
            ent.push_back(make_uint2((uint32_t)d->blend_transform[e], wbits));
          }
          ent_off.push_back((int)ent.size());
        }
        const int r0 = tile_off.back() * kTileRows, r1 = std::min(N, t_end * kTileRows);
        for (int r = r0; r < r1; ++r)
          if (row_live(r)) local[r] = (uint16_t)slot[row_blend(r)];
        rec_off.push_back(rec_off.back() + n);
        max_rec = std::max(max_rec, n);
        tile_off.push_back(t_end);
        ++ng;
        open.clear();
      };

      for (int t = 0; t < nt; ++t) {
        const int r0 = t * kTileRows, r1 = std::min(N, r0 + kTileRows);
        fresh.clear();  // distinct blends this tile would add to the open group
        for (int r = r0; r < r1; ++r) {
          if (!row_live(r)) continue;
          const int b = row_blend(r);
          if (in_group[b] != ng && seen_in_tile[b] != t) {
            seen_in_tile[b] = t;
            fresh.push_back(b);
          }
        }
        if (tile_off.back() != t && (int)open.size() + (int)fresh.size() > cap) {  // close the group before this tile
          close_group(t);
          for (int b : fresh) seen_in_tile[b] = -1;
          --t;  // redo this tile in the new group
          continue;
        }
        for (int b : fresh) {
          in_group[b] = ng;
          open.push_back(b);
        }
      }
      if (nt > 0) close_group(nt);
      if ((rc = upload<int>(mesh, tile_off.data(), tile_off.size(), &m.group_tile_offsets)) != P3CS_OK) return bail(rc);
      m.num_groups = ng;
      m.max_group_records = max_rec;
      if ((rc = upload<uint16_t>(mesh, local.data(), local.size(), &m.row_local)) != P3CS_OK) return bail(rc);
      if ((rc = upload<int>(mesh, rec_off.data(), rec_off.size(), &m.group_rec_offsets)) != P3CS_OK) return bail(rc);
      if ((rc = upload<int>(mesh, rec_blend.data(), rec_blend.size(), &m.group_rec_blend)) != P3CS_OK) return bail(rc);
      if ((rc = upload<int>(mesh, ent_off.data(), ent_off.size(), &m.grec_entry_offsets)) != P3CS_OK) return bail(rc);
      if ((rc = upload<uint2>(mesh, ent.data(), ent.size(), &m.grec_entries)) != P3CS_OK) return bail(rc);
      // fixed-width blocks: the first four entries of every group record, count in bits 16..31
      const size_t nrec_total = rec_blend.size();
      // padded by one CTA's worth of records: the kernel prefetches block (group start + tid)
      std::vector<uint4> blocks((nrec_total + kStripThreads) * 2, make_uint4(0, 0, 0, 0));
      for (size_t r = 0; r < nrec_total; ++r) {
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int beg = ent_off[r], cnt = ent_off[r + 1] - ent_off[r];
        for (int k = 0; k < std::min(cnt, 4); ++k) {
          w[2 * k] = ent[beg + k].x;
          w[2 * k + 1] = ent[beg + k].y;
        }
        w[0] |= (uint32_t)std::min(cnt, 0xFFFF) << 16;
        blocks[2 * r] = make_uint4(w[0], w[1], w[2], w[3]);
        blocks[2 * r + 1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
      if ((rc = upload<uint4>(mesh, blocks.data(), blocks.size(), &m.grec_blocks)) != P3CS_OK) return bail(rc);
    }
    m.index_bytes = wide ? 4 : 2;
    const void *dev_index = nullptr;
    if (wide) { const uint32_t *p; rc = upload<uint32_t>(mesh, idx32.data(), idx32.size(), &p); dev_index = p; }
    else { const uint16_t *p; rc = upload<uint16_t>(mesh, idx16.data(), idx16.size(), &p); dev_index = p; }
    if (rc != P3CS_OK) return bail(rc);
    m.index = dev_index;
    if (!full && (rc = upload<uint32_t>(mesh, mask.data(), mask.size(), &m.rows_mask)) != P3CS_OK) return bail(rc);
    if ((rc = upload<int>(mesh, d->blend_offsets, (size_t)d->num_blends + 1, &m.blend_offsets)) != P3CS_OK) return bail(rc);
    if ((rc = upload<int>(mesh, d->blend_transform, (size_t)nnz, &m.blend_transform)) != P3CS_OK) return bail(rc);
    if ((rc = upload<float>(mesh, d->blend_weight, (size_t)nnz, &m.blend_weight)) != P3CS_OK) return bail(rc);
    (void)has_blend;
  } else {
    m.num_blends = 0;
  }

  // ---- morphs: dense delta columns + slider rows -> per-row CSR, in the reference's
  // iteration order so rows hit by several morphs add up in the same order
  if (d->num_morphs > 0) {
    if (d->morphs == nullptr || d->num_sliders <= 0) return bail(fail(P3CS_ERR_INVALID, "morphs without sliders"));
    if (d->num_sliders > 65536) return bail(fail(P3CS_ERR_UNSUPPORTED, "more than 65536 sliders"));
    std::vector<int> counts((size_t)N + 1, 0);
    std::vector<int> morph_dyn(d->num_morphs, -1);
    for (int mi = 0; mi < d->num_morphs; ++mi) {
      const p3cs_morph_desc &mo = d->morphs[mi];
      if (mo.slider < 0 || mo.slider >= d->num_sliders) return bail(fail(P3CS_ERR_INVALID, "morph %d: slider %d out of range", mi, mo.slider));
      if ((rc = check_column(d, mo.base, "morph base column")) != P3CS_OK) return bail(rc);
      if ((rc = check_column(d, mo.delta, "morph delta column")) != P3CS_OK) return bail(rc);
      if ((rc = check_ranges(mo.row_ranges, mo.num_row_ranges, N, "slider rows")) != P3CS_OK) return bail(rc);
      if (out_of_array[mo.base.array] < 0) return bail(fail(P3CS_ERR_INVALID, "morph %d: base column in a non-output array", mi));
      auto is_float = [](int nt) { return nt == P3CS_NT_FLOAT32 || nt == P3CS_NT_FLOAT64; };
      if (!is_float(mo.base.numeric_type) || !is_float(mo.delta.numeric_type) || mo.base.num_values > 4 ||
          mo.base.start % 4 != 0 || mo.delta.start % 4 != 0)
        return bail(fail(P3CS_ERR_UNSUPPORTED, "morph %d: only 4-byte aligned float32 / float64 base and delta columns are supported", mi));
      int c = find_dyn(out_of_array[mo.base.array], mo.base.start);
      if (c < 0) {
        if (m.num_dyn >= kMaxDynColumns) return bail(fail(P3CS_ERR_UNSUPPORTED, "more than %d animated columns", kMaxDynColumns));
        c = m.num_dyn++;
        m.dyn[c].output = (int16_t)out_of_array[mo.base.array];
        m.dyn[c].start = (int16_t)mo.base.start;
        m.dyn[c].num_values = (int8_t)mo.base.num_values;
        m.dyn[c].skin = 0;
        m.dyn[c].f64 = mo.base.numeric_type == P3CS_NT_FLOAT64 ? 1 : 0;
      }
      m.dyn[c].homogeneous = (mo.base.num_values == 4 && (mo.base.contents == P3CS_C_POINT || mo.base.contents == P3CS_C_TEXCOORD)) ? 1 : 0;
      morph_dyn[mi] = c;
      for (int i = 0; i < mo.num_row_ranges; ++i)
        for (int r = mo.row_ranges[2 * i]; r < mo.row_ranges[2 * i + 1]; ++r) counts[r + 1]++;
    }
    for (int r = 0; r < N; ++r) counts[r + 1] += counts[r];
    const size_t total = counts[N];
    std::vector<uint32_t> keys(std::max<size_t>(total, 1));
    std::vector<float4> deltas(std::max<size_t>(total, 1));
    std::vector<int> cursor(counts.begin(), counts.end() - 1);
    for (int mi = 0; mi < d->num_morphs; ++mi) {
      const p3cs_morph_desc &mo = d->morphs[mi];
      const uint8_t *dbase = static_cast<const uint8_t *>(d->arrays[mo.delta.array].data) + mo.delta.start;
      const int dstride = d->arrays[mo.delta.array].stride;
      const int dn = std::min(mo.delta.num_values, mo.base.num_values == 4 && !m.dyn[morph_dyn[mi]].homogeneous ? 4 : 3);
      for (int i = 0; i < mo.num_row_ranges; ++i)
        for (int r = mo.row_ranges[2 * i]; r < mo.row_ranges[2 * i + 1]; ++r) {
          float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // Packer::get_data3f/4f pad with 0 (geomVertexColumn.cxx:718-737)
          if (mo.delta.numeric_type == P3CS_NT_FLOAT64) {  // read as float, like GeomVertexReader::get_data3/4
            double dv[4];
            memcpy(dv, dbase + (size_t)r * dstride, (size_t)dn * 8);
            for (int k = 0; k < dn; ++k) v[k] = (float)dv[k];
          } else {
            memcpy(v, dbase + (size_t)r * dstride, (size_t)dn * 4);
          }
          const int e = cursor[r]++;
          keys[e] = ((uint32_t)morph_dyn[mi] << 16) | (uint32_t)mo.slider;
          // a delta whose 4th component is never read carries its key there (one load per entry
          // in the specialised strip modes)
          if (!(mo.base.num_values == 4 && !m.dyn[morph_dyn[mi]].homogeneous)) memcpy(&v[3], &keys[e], 4);
          deltas[e] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    m.has_morphs = 1;
    if ((rc = upload<int>(mesh, counts.data(), counts.size(), &m.morph_row_offsets)) != P3CS_OK) return bail(rc);
    if ((rc = upload<uint32_t>(mesh, keys.data(), keys.size(), &m.morph_keys)) != P3CS_OK) return bail(rc);
    if ((rc = upload<float4>(mesh, deltas.data(), deltas.size(), &m.morph_deltas)) != P3CS_OK) return bail(rc);
  }
  // ---- slice geometry of the strip kernel (a slice = the 32 rows one warp stages at a time)
  if (m.rows_per_thread == 0) m.rows_per_thread = 1;
  if (m.tile_bufs == 0) m.tile_bufs = 3;
  if (m.strip_rec_floats == 0) m.strip_rec_floats = m.full_records ? kRecFull : kRecCompact;
  m.num_tiles = (N + kStripThreads * m.rows_per_thread - 1) / (kStripThreads * m.rows_per_thread);
  m.slice_bytes = 0;
  for (int o = 0; o < m.num_outputs; ++o) {
    m.slice_out_offset[o] = m.slice_bytes;
    m.slice_bytes += (int)round_up((size_t)32 * m.stride[o], 128);
  }
  if (m.group_rec_offsets == nullptr) {  // no blend table: one tile per group, zero records each
    m.num_groups = m.num_tiles;
    std::vector<int> zeros((size_t)m.num_groups + 1, 0), iota((size_t)m.num_groups + 1);
    for (size_t i = 0; i < iota.size(); ++i) iota[i] = (int)i;
    if ((rc = upload<int>(mesh, zeros.data(), zeros.size(), &m.group_rec_offsets)) != P3CS_OK) return bail(rc);
    if ((rc = upload<int>(mesh, iota.data(), iota.size(), &m.group_tile_offsets)) != P3CS_OK) return bail(rc);
    std::vector<uint4> blocks((size_t)kStripThreads * 2, make_uint4(0, 0, 0, 0));  // the kernel prefetches block [tid]
    if ((rc = upload<uint4>(mesh, blocks.data(), blocks.size(), &m.grec_blocks)) != P3CS_OK) return bail(rc);
  }
  {
    const char *path = getenv("P3CS_PATH");
    mesh->use_strip = strip_smem_bytes(m) <= 200 * 1024 && m.num_transforms < 65536 &&
                      !(path != nullptr && strcmp(path, "wide") == 0);
  }
  // uploads read from host vectors that die at scope exit: finish them now
  cudaError_t err = cudaStreamSynchronize(ctx->stream);
  if (err != cudaSuccess) return bail(fail(P3CS_ERR_CUDA, "mesh upload failed: %s", cudaGetErrorString(err)));
  mesh->out_of_array = out_of_array;
  *out = mesh;
  return P3CS_OK;
}

extern "C" int p3cs_mesh_destroy(p3cs_mesh mesh) {
  if (mesh == nullptr) return P3CS_OK;
  DeviceGuard g(mesh->ctx->device);
  for (void *p : mesh->allocations) cudaFree(p);
  delete mesh;
  return P3CS_OK;
}

extern "C" int p3cs_mesh_num_outputs(p3cs_mesh mesh) { return mesh ? mesh->dev.num_outputs : 0; }
extern "C" int p3cs_mesh_output_stride(p3cs_mesh mesh, int o) {
  return (mesh && o >= 0 && o < mesh->dev.num_outputs) ? mesh->dev.stride[o] : 0;
}
extern "C" int p3cs_mesh_num_rows(p3cs_mesh mesh) { return mesh ? mesh->dev.num_rows : 0; }

// ------------------------------------------------------------------ group
static size_t record_floats(const MeshDev &m) { return m.full_records ? kRecFull : kRecCompact; }

extern "C" int p3cs_group_create(p3cs_mesh mesh, int n, const p3cs_group_buffers *buffers, p3cs_group *out) {
  if (out == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_create: out is NULL");
  *out = nullptr;
  if (mesh == nullptr || n <= 0) return fail(P3CS_ERR_INVALID, "p3cs_group_create: bad mesh / num_instances");
  DeviceGuard g(mesh->ctx->device);
  const MeshDev &m = mesh->dev;
  p3cs_group grp = new (std::nothrow) p3cs_group_s;
  if (grp == nullptr) return fail(P3CS_ERR_NOMEM, "out of host memory");
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
    if (err != cudaSuccess) return bail(fail(P3CS_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(err)));
  }
  *out = grp;
  return P3CS_OK;
}

extern "C" int p3cs_group_destroy(p3cs_group grp) {
  if (grp == nullptr) return P3CS_OK;
  DeviceGuard g(grp->mesh->ctx->device);
  cudaStreamSynchronize(grp->mesh->ctx->stream);
  for (void *p : grp->allocations) cudaFree(p);
  for (auto e : grp->ev)
    if (e) cudaEventDestroy(e);
  for (auto e : grp->prof) cudaEventDestroy(e);
  if (grp->frames_dev != nullptr) cudaFree(grp->frames_dev);
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
  if (grp == nullptr) return fail(P3CS_ERR_INVALID, "%s: group is NULL", what);
  if (first < 0 || count < 0 || first + count > grp->n)
    return fail(P3CS_ERR_INVALID, "%s: instances [%d,%d) outside the group of %d", what, first, first + count, grp->n);
  return P3CS_OK;
}

extern "C" int p3cs_group_set_palettes(p3cs_group grp, int first, int count, const float *host) {
  int rc = check_span(grp, first, count, "p3cs_group_set_palettes");
  if (rc != P3CS_OK) return rc;
  if (host == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_set_palettes: host pointer is NULL");
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
  if (host == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_set_sliders: host pointer is NULL");
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
  if (grp == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_animate: group is NULL");
  DeviceGuard g(grp->mesh->ctx->device);
  return animate_range(grp, 0, grp->n, true);
}

static int copy_output_rows(p3cs_group grp, int o, int first, int count, void *host_dst, cudaStream_t stream);

static int ensure_side_streams(p3cs_context ctx) {
  if (ctx->fork_ev != nullptr) return P3CS_OK;
  CUDA_TRY(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
  for (int k = 0; k < p3cs_context_s::kSideStreams; ++k) {
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->side[k], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->join_ev[k], cudaEventDisableTiming));
  }
  return P3CS_OK;
}

// The crowd form of p3cs_animate_vertices: host palettes / sliders in, host arrays out, for ALL instances of the
// group, software-pipelined in chunks of instances over three streams (upload, kernels, download) so that the
// PCIe download -- the bound of this call -- never waits for an upload or a kernel after the first chunk.
extern "C" int p3cs_group_animate_host(p3cs_group grp, const float *host_palettes, const float *host_sliders,
                                       void *const *host_outputs, int chunk_instances) {
  if (grp == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_animate_host: group is NULL");
  const MeshDev &m = grp->mesh->dev;
  p3cs_context ctx = grp->mesh->ctx;
  if (m.num_transforms > 0 && host_palettes == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_animate_host: palettes are NULL");
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
  for (int c = 0; c < nchunks; ++c) {
    const int first = c * chunk, count = std::min(chunk, grp->n - first);
    if (pal_per > 0)
      CUDA_TRY(cudaMemcpyAsync(const_cast<float *>(grp->dev.palettes) + pal_per * first, host_palettes + pal_per * first,
                               pal_per * count * 4, cudaMemcpyHostToDevice, up));
    if (sl_per > 0 && host_sliders != nullptr)
      CUDA_TRY(cudaMemcpyAsync(const_cast<float *>(grp->dev.sliders) + sl_per * first, host_sliders + sl_per * first,
                               sl_per * count * 4, cudaMemcpyHostToDevice, up));
    CUDA_TRY(cudaEventRecord(ctx->pipe_ev[2 * c], up));
    CUDA_TRY(cudaStreamWaitEvent(run, ctx->pipe_ev[2 * c], 0));
    if ((rc = animate_range(grp, first, count, false)) != P3CS_OK) break;
    total_launches += grp->last.launches;
    CUDA_TRY(cudaEventRecord(ctx->pipe_ev[2 * c + 1], run));
    if (host_outputs != nullptr) {
      CUDA_TRY(cudaStreamWaitEvent(down, ctx->pipe_ev[2 * c + 1], 0));
      for (int o = 0; o < m.num_outputs; ++o)
        if (host_outputs[o] != nullptr) {
          uint8_t *dst = static_cast<uint8_t *>(host_outputs[o]) + (size_t)m.num_rows * m.stride[o] * first;
          if ((rc = copy_output_rows(grp, o, first, count, dst, down)) != P3CS_OK) break;
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
  if (num_groups < 0 || (num_groups > 0 && groups == nullptr)) return fail(P3CS_ERR_INVALID, "p3cs_animate_groups: bad arguments");
  if (num_groups == 0) return P3CS_OK;
  p3cs_context ctx = nullptr;
  for (int i = 0; i < num_groups; ++i) {
    if (groups[i] == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_animate_groups: group %d is NULL", i);
    if (ctx == nullptr) ctx = groups[i]->mesh->ctx;
    else if (groups[i]->mesh->ctx != ctx) return fail(P3CS_ERR_INVALID, "p3cs_animate_groups: group %d belongs to another context", i);
  }
  DeviceGuard g(ctx->device);
  if (num_groups == 1) return animate_range(groups[0], 0, groups[0]->n, true);
  {
    const int rc0 = ensure_side_streams(ctx);
    if (rc0 != P3CS_OK) return rc0;
  }
  // fork: group i runs on lane i % (1 + kSideStreams); lane 0 is the context's own stream
  const int lanes = std::min(num_groups, 1 + p3cs_context_s::kSideStreams);
  CUDA_TRY(cudaEventRecord(ctx->fork_ev, ctx->stream));
  for (int k = 1; k < lanes; ++k) CUDA_TRY(cudaStreamWaitEvent(ctx->side[k - 1], ctx->fork_ev, 0));
  int rc = P3CS_OK;
  for (int i = 0; i < num_groups && rc == P3CS_OK; ++i) {
    const int lane = i % lanes;
    rc = animate_range(groups[i], 0, groups[i]->n, true, lane == 0 ? ctx->stream : ctx->side[lane - 1]);
  }
  // join (also on failure, so the side streams never run ahead of the context's stream)
  for (int k = 1; k < lanes; ++k) {
    cudaEventRecord(ctx->join_ev[k - 1], ctx->side[k - 1]);
    cudaStreamWaitEvent(ctx->stream, ctx->join_ev[k - 1], 0);
  }
  return rc;
}

// Only the instances that changed: what the UpdateSeq test of GeomVertexData::animate_vertices
// (geomVertexData.cxx:937-946) does per character, for a crowd.
extern "C" int p3cs_group_animate_instances(p3cs_group grp, const int32_t *host_instances, int count) {
  if (grp == nullptr || count < 0 || (count > 0 && host_instances == nullptr)) return fail(P3CS_ERR_INVALID, "p3cs_group_animate_instances: bad arguments");
  for (int i = 0; i < count; ++i)
    if (host_instances[i] < 0 || host_instances[i] >= grp->n)
      return fail(P3CS_ERR_INVALID, "p3cs_group_animate_instances: entry %d = %d outside the group of %d", i, host_instances[i], grp->n);
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

extern "C" int p3cs_group_set_profiling(p3cs_group grp, int enable) {
  if (grp == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_set_profiling: group is NULL");
  grp->profiling = enable != 0;
  return P3CS_OK;
}

extern "C" int p3cs_group_last_timings(p3cs_group grp, p3cs_timings *out) {
  if (grp == nullptr || out == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_last_timings: NULL argument");
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
  if (grp == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_enable_bounds: group is NULL");
  if (vertex_column == nullptr) {  // off (the buffer stays allocated until the group is destroyed)
    grp->dev.bounds = nullptr;
    return P3CS_OK;
  }
  const p3cs_mesh_s *mesh = grp->mesh;
  const MeshDev &m = mesh->dev;
  const p3cs_column_desc &c = *vertex_column;
  if (c.array < 0 || c.array >= (int)mesh->out_of_array.size() || mesh->out_of_array[c.array] < 0)
    return fail(P3CS_ERR_INVALID, "p3cs_group_enable_bounds: column is not in an output array");
  const int o = mesh->out_of_array[c.array];
  // GeomVertexReader::get_data3 on anything else divides by w or converts (geomVertexColumn.cxx:2800-2808)
  if (c.numeric_type != P3CS_NT_FLOAT32 || c.num_values != 3 || c.start < 0 || c.start % 4 != 0 || c.start + 12 > m.stride[o])
    return fail(P3CS_ERR_UNSUPPORTED, "p3cs_group_enable_bounds: only a 4-byte aligned float32 x 3 column is supported");
  DeviceGuard g(mesh->ctx->device);
  if (grp->bounds_buf == nullptr) {
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, std::max<size_t>((size_t)grp->n * 32, 256)));
    grp->allocations.push_back(p);
    grp->bounds_buf = static_cast<float *>(p);
    launch_bounds_init(GroupDev{nullptr, nullptr, nullptr, {}, {}, nullptr, grp->bounds_buf, 0, 0, nullptr}, 0, grp->n, mesh->ctx->stream);
  }
  grp->dev.bounds = grp->bounds_buf;
  grp->dev.bounds_output = o;
  grp->dev.bounds_start = c.start;
  return P3CS_OK;
}

extern "C" int p3cs_group_read_bounds(p3cs_group grp, int first, int count, float *host_dst) {
  int rc = check_span(grp, first, count, "p3cs_group_read_bounds");
  if (rc != P3CS_OK) return rc;
  if (host_dst == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_read_bounds: destination is NULL");
  if (grp->dev.bounds == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_read_bounds: bounds are not enabled (p3cs_group_enable_bounds)");
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
  if (o < 0 || o >= m.num_outputs || host_dst == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_read_output: bad output / destination");
  DeviceGuard g(grp->mesh->ctx->device);
  const size_t row_bytes = (size_t)m.num_rows * m.stride[o];
  if (row_bytes == 0 || count == 0) return P3CS_OK;
  return copy_output_rows(grp, o, first, count, host_dst, grp->mesh->ctx->stream);
}

extern "C" int p3cs_group_read_selection(p3cs_group grp, int32_t *host_dst) {
  if (grp == nullptr || host_dst == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_read_selection: NULL argument");
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
  if (err != cudaSuccess) return fail(P3CS_ERR_CUDA, "p3cs_group_read_selection: %s", cudaGetErrorString(err));
  return P3CS_OK;
}

extern "C" int p3cs_group_read_blends(p3cs_group grp, int instance, float *host_dst) {
  int rc = check_span(grp, instance, 1, "p3cs_group_read_blends");
  if (rc != P3CS_OK) return rc;
  if (host_dst == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_read_blends: destination is NULL");
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

extern "C" int p3cs_animate_vertices(p3cs_group grp, int instance, const float *host_palette, const float *host_sliders,
                                     void *const *host_outputs) {
  int rc = check_span(grp, instance, 1, "p3cs_animate_vertices");
  if (rc != P3CS_OK) return rc;
  const MeshDev &m = grp->mesh->dev;
  if (m.num_transforms > 0 && (rc = p3cs_group_set_palettes(grp, instance, 1, host_palette)) != P3CS_OK) return rc;
  if (m.num_sliders > 0 && (rc = p3cs_group_set_sliders(grp, instance, 1, host_sliders)) != P3CS_OK) return rc;
  DeviceGuard g(grp->mesh->ctx->device);
  // Page-locked, device-mapped destinations (p3cs_host_register / cudaHostRegister): the kernel's row stores go
  // straight to host memory, no device->host copy afterwards.  The bulk stores move 16-byte units, hence the
  // alignment and size conditions; anything else takes the copy below.
  if (host_outputs != nullptr && m.num_outputs > 0) {
    GroupDev direct = grp->dev;
    bool ok = true;
    for (int o = 0; o < m.num_outputs && ok; ++o) {
      cudaPointerAttributes attr{};
      const size_t bytes = (size_t)m.num_rows * m.stride[o];
      ok = host_outputs[o] != nullptr && bytes % 16 == 0 && reinterpret_cast<uintptr_t>(host_outputs[o]) % 16 == 0 &&
           cudaPointerGetAttributes(&attr, host_outputs[o]) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
           attr.devicePointer != nullptr;
      if (ok) direct.out[o] = reinterpret_cast<uint8_t *>(reinterpret_cast<uintptr_t>(attr.devicePointer) - (uintptr_t)instance * grp->dev.pitch[o]);
    }
    cudaGetLastError();  // cudaPointerGetAttributes on plain malloc memory may leave a sticky-free error behind
    if (ok) {
      if ((rc = animate_range(grp, instance, 1, false, nullptr, &direct)) != P3CS_OK) return rc;
      return p3cs_context_sync(grp->mesh->ctx);
    }
  }
  if ((rc = animate_range(grp, instance, 1, false)) != P3CS_OK) return rc;
  if (host_outputs != nullptr)
    for (int o = 0; o < m.num_outputs; ++o)
      if (host_outputs[o] != nullptr && (rc = p3cs_group_read_output(grp, o, instance, 1, host_outputs[o])) != P3CS_OK) return rc;
  return p3cs_context_sync(grp->mesh->ctx);
}

extern "C" int p3cs_host_register(p3cs_context ctx, void *host_ptr, size_t bytes) {
  if (ctx == nullptr || host_ptr == nullptr || bytes == 0) return fail(P3CS_ERR_INVALID, "p3cs_host_register: bad arguments");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaHostRegister(host_ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
  return P3CS_OK;
}

extern "C" int p3cs_host_unregister(p3cs_context ctx, void *host_ptr) {
  if (ctx == nullptr || host_ptr == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_host_unregister: bad arguments");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(cudaHostUnregister(host_ptr));
  return P3CS_OK;
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
  if (out == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_skeleton_create: out is NULL");
  *out = nullptr;
  if (ctx == nullptr || d == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_skeleton_create: NULL argument");
  if (d->struct_size != sizeof(p3cs_skeleton_desc)) return fail(P3CS_ERR_INVALID, "p3cs_skeleton_desc::struct_size mismatch");
  const int J = d->num_joints;
  if (J < 0 || J > 3000) return fail(P3CS_ERR_UNSUPPORTED, "num_joints %d outside 0..3000", J);
  int cs = d->coordinate_system == 0 ? P3CS_CS_ZUP_RIGHT : d->coordinate_system;
  if (cs < P3CS_CS_ZUP_RIGHT || cs > P3CS_CS_YUP_LEFT) return fail(P3CS_ERR_INVALID, "bad coordinate_system %d", cs);
  std::vector<int> level(std::max(J, 1), 0);
  int max_level = 0;
  for (int j = 0; j < J; ++j) {
    const int p = d->parent[j];
    if (p >= j) return fail(P3CS_ERR_INVALID, "joint %d: parent %d does not precede it (joints must be listed parents-first)", j, p);
    level[j] = p < 0 ? 0 : level[p] + 1;
    max_level = std::max(max_level, level[j]);
    for (int i = 0; i < 12; ++i) {
      const int len = d->table_length[j * 12 + i], off = d->table_offset[j * 12 + i];
      if (len < 0 || off < 0 || off + len > d->num_table_values) return fail(P3CS_ERR_INVALID, "joint %d: table %d out of range", j, i);
    }
  }
  for (int t = 0; t < d->num_palette; ++t)
    if (d->palette_joint[t] < 0 || d->palette_joint[t] >= J) return fail(P3CS_ERR_INVALID, "palette slot %d has no joint", t);
  DeviceGuard g(ctx->device);
  p3cs_skeleton sk = new (std::nothrow) p3cs_skeleton_s;
  if (sk == nullptr) return fail(P3CS_ERR_NOMEM, "out of host memory");
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
  if ((rc = upload_skel<int>(sk, d->has_channel, J, &s.has_channel)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<int>(sk, d->table_length, (size_t)J * 12, &s.table_length)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<int>(sk, d->table_offset, (size_t)J * 12, &s.table_offset)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<float>(sk, d->tables, d->num_table_values, &s.tables)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<float>(sk, d->default_value, (size_t)J * 16, &s.default_value)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<float>(sk, d->initial_inverse, (size_t)J * 16, &s.initial_inverse)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<float>(sk, d->root_xform, 16, &s.root_xform)) != P3CS_OK) return bail(rc);
  if ((rc = upload_skel<int>(sk, d->palette_joint, d->num_palette, &s.palette_joint)) != P3CS_OK) return bail(rc);
  cudaError_t err = cudaStreamSynchronize(ctx->stream);
  if (err != cudaSuccess) return bail(fail(P3CS_ERR_CUDA, "skeleton upload failed: %s", cudaGetErrorString(err)));
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

static int pose_common(p3cs_group grp, p3cs_skeleton sk, int first, int count, const int32_t *host_frames,
                       const int32_t *host_next_frames, const float *host_fracs, int blend_type, const char *what) {
  int rc = check_span(grp, first, count, what);
  if (rc != P3CS_OK) return rc;
  if (sk == nullptr || host_frames == nullptr) return fail(P3CS_ERR_INVALID, "%s: NULL argument", what);
  if (sk->num_palette != grp->mesh->dev.num_transforms)
    return fail(P3CS_ERR_INVALID, "skeleton feeds %d palette slots, the mesh has %d transforms", sk->num_palette, grp->mesh->dev.num_transforms);
  if (count == 0) return P3CS_OK;
  DeviceGuard g(grp->mesh->ctx->device);
  cudaStream_t s = grp->mesh->ctx->stream;
  if (grp->frames_capacity < count) {  // frames, next frames, fractions: three arrays of `capacity` words
    if (grp->frames_dev != nullptr) {
      CUDA_TRY(cudaStreamSynchronize(s));
      CUDA_TRY(cudaFree(grp->frames_dev));
    }
    grp->frames_dev = nullptr;
    CUDA_TRY(cudaMalloc(&grp->frames_dev, (size_t)count * 3 * sizeof(int)));
    grp->frames_capacity = count;
  }
  int *frames_dev = grp->frames_dev, *next_dev = nullptr;
  float *fracs_dev = nullptr;
  CUDA_TRY(cudaMemcpyAsync(frames_dev, host_frames, (size_t)count * sizeof(int), cudaMemcpyHostToDevice, s));
  if (host_next_frames != nullptr) {
    next_dev = grp->frames_dev + grp->frames_capacity;
    fracs_dev = reinterpret_cast<float *>(grp->frames_dev + 2 * (size_t)grp->frames_capacity);
    CUDA_TRY(cudaMemcpyAsync(next_dev, host_next_frames, (size_t)count * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(fracs_dev, host_fracs, (size_t)count * sizeof(float), cudaMemcpyHostToDevice, s));
  }
  launch_pose(sk->dev, const_cast<float *>(grp->dev.palettes), sk->num_palette, frames_dev, next_dev, fracs_dev, blend_type, first, count, s);
  CUDA_TRY(cudaGetLastError());
  return P3CS_OK;
}

extern "C" int p3cs_group_pose(p3cs_group grp, p3cs_skeleton sk, int first, int count, const int32_t *host_frames) {
  return pose_common(grp, sk, first, count, host_frames, nullptr, nullptr, P3CS_BT_NORMALIZED_LINEAR, "p3cs_group_pose");
}

extern "C" int p3cs_group_pose_blend(p3cs_group grp, p3cs_skeleton sk, int first, int count, const int32_t *host_frames,
                                     const int32_t *host_next_frames, const float *host_fracs, int blend_type) {
  if (host_next_frames == nullptr || host_fracs == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_pose_blend: NULL argument");
  if (blend_type != P3CS_BT_LINEAR && blend_type != P3CS_BT_NORMALIZED_LINEAR)
    return fail(P3CS_ERR_UNSUPPORTED, "p3cs_group_pose_blend: blend type %d (only BT_linear and BT_normalized_linear)", blend_type);
  return pose_common(grp, sk, first, count, host_frames, host_next_frames, host_fracs, blend_type, "p3cs_group_pose_blend");
}

extern "C" int p3cs_group_read_palettes(p3cs_group grp, int first, int count, float *host_dst) {
  int rc = check_span(grp, first, count, "p3cs_group_read_palettes");
  if (rc != P3CS_OK) return rc;
  if (host_dst == nullptr) return fail(P3CS_ERR_INVALID, "p3cs_group_read_palettes: destination is NULL");
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
      p3cs_host_register,      p3cs_host_unregister, p3cs_animate_groups};
  return &api;
}
