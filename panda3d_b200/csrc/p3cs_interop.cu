// p3cs_interop.cu -- sharing memory with other GPUs, processes, renderers and the host: peer access and CUDA IPC
// (fused skin + gather over NVLink), exportable allocations over the driver's virtual memory management API
// ("next" row f2), page-locked host buffers the kernel stores into directly.  Host C++.
#include "p3cs_internal.h"

#include <cuda.h>  // driver API types only; entry points come from cudaGetDriverEntryPoint
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>

using namespace p3cs;

extern "C" int p3cs_context_enable_peer_access(p3cs_context ctx, int peer_device) {
  if (ctx == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "context is NULL");
  if (peer_device == ctx->device) return P3CS_OK;
  DeviceGuard g(ctx->device);
  int can = 0;
  CUDA_TRY(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
  if (!can) return p3cs_fail(P3CS_ERR_UNSUPPORTED, "device %d cannot access device %d", ctx->device, peer_device);
  cudaError_t err = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (err == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return P3CS_OK;
  }
  if (err != cudaSuccess) return p3cs_fail(P3CS_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(err));
  return P3CS_OK;
}

extern "C" int p3cs_device_alloc(p3cs_context ctx, size_t bytes, void **out) {
  if (ctx == nullptr || out == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_device_alloc: NULL argument");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaMalloc(out, bytes > 0 ? bytes : 256));
  return P3CS_OK;
}

extern "C" int p3cs_device_free(p3cs_context ctx, void *ptr) {
  if (ctx == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "context is NULL");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaFree(ptr));
  return P3CS_OK;
}

extern "C" int p3cs_device_read(p3cs_context ctx, const void *device_src, void *host_dst, size_t bytes) {
  if (ctx == nullptr || device_src == nullptr || host_dst == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_device_read: NULL argument");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return P3CS_OK;
}

extern "C" int p3cs_ipc_export(p3cs_context ctx, void *ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (ctx == nullptr || ptr == nullptr || handle == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_ipc_export: NULL argument");
  DeviceGuard g(ctx->device);
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, 64);
  return P3CS_OK;
}

extern "C" int p3cs_ipc_open(p3cs_context ctx, const unsigned char handle[64], void **out) {
  if (ctx == nullptr || handle == nullptr || out == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_ipc_open: NULL argument");
  DeviceGuard g(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return P3CS_OK;
}

extern "C" int p3cs_ipc_close(p3cs_context ctx, void *ptr) {
  if (ctx == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "context is NULL");
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
  if (r != CUDA_SUCCESS) return p3cs_fail(P3CS_ERR_CUDA, "cuMemAddressReserve failed (%d)", (int)r);
  r = a.map(ptr, bytes, 0, handle, 0);
  if (r != CUDA_SUCCESS) {
    a.address_free(ptr, bytes);
    return p3cs_fail(P3CS_ERR_CUDA, "cuMemMap failed (%d)", (int)r);
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
    return p3cs_fail(P3CS_ERR_CUDA, "cuMemSetAccess failed (%d)", (int)r);
  }
  *out = reinterpret_cast<void *>(ptr);
  std::lock_guard<std::mutex> lock(g_shareable_lock);
  g_shareable[*out] = Shareable{handle, bytes};
  return P3CS_OK;
}

}  // namespace

extern "C" int p3cs_shareable_alloc(p3cs_context ctx, size_t bytes, void **out_ptr, int *out_fd, size_t *out_bytes) {
  if (ctx == nullptr || out_ptr == nullptr || out_fd == nullptr || bytes == 0) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_shareable_alloc: bad arguments");
  *out_ptr = nullptr;
  *out_fd = -1;
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaFree(nullptr));  // make sure the primary context exists
  const VmmApi &a = vmm();
  if (!a.ok) return p3cs_fail(P3CS_ERR_UNSUPPORTED, "the CUDA driver does not expose the virtual memory management API");
  const CUmemAllocationProp prop = shareable_prop(ctx->device);
  size_t gran = 0;
  CUresult r = a.get_granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
  if (r != CUDA_SUCCESS || gran == 0) return p3cs_fail(P3CS_ERR_CUDA, "cuMemGetAllocationGranularity failed (%d)", (int)r);
  const size_t size = round_up(bytes, gran);
  CUmemGenericAllocationHandle handle;
  r = a.create(&handle, size, &prop, 0);
  if (r != CUDA_SUCCESS) return p3cs_fail(r == CUDA_ERROR_OUT_OF_MEMORY ? P3CS_ERR_NOMEM : P3CS_ERR_CUDA, "cuMemCreate(%zu bytes) failed (%d)", size, (int)r);
  int fd = -1;
  r = a.export_handle(&fd, handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    a.release(handle);
    return p3cs_fail(P3CS_ERR_CUDA, "cuMemExportToShareableHandle failed (%d)", (int)r);
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
  if (ctx == nullptr || out_ptr == nullptr || fd < 0 || bytes == 0) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_shareable_import: bad arguments");
  *out_ptr = nullptr;
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaFree(nullptr));
  const VmmApi &a = vmm();
  if (!a.ok) return p3cs_fail(P3CS_ERR_UNSUPPORTED, "the CUDA driver does not expose the virtual memory management API");
  CUmemGenericAllocationHandle handle;
  CUresult r = a.import_handle(&handle, reinterpret_cast<void *>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (r != CUDA_SUCCESS) return p3cs_fail(P3CS_ERR_CUDA, "cuMemImportFromShareableHandle failed (%d)", (int)r);
  int rc = map_shareable(ctx, handle, bytes, out_ptr);
  if (rc != P3CS_OK) a.release(handle);
  return rc;
}

extern "C" int p3cs_shareable_free(p3cs_context ctx, void *ptr) {
  if (ctx == nullptr || ptr == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_shareable_free: bad arguments");
  DeviceGuard g(ctx->device);
  Shareable sh;
  {
    std::lock_guard<std::mutex> lock(g_shareable_lock);
    auto it = g_shareable.find(ptr);
    if (it == g_shareable.end()) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_shareable_free: not a shareable allocation");
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


extern "C" int p3cs_host_register(p3cs_context ctx, void *host_ptr, size_t bytes) {
  if (ctx == nullptr || host_ptr == nullptr || bytes == 0) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_host_register: bad arguments");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaHostRegister(host_ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
  return P3CS_OK;
}

extern "C" int p3cs_host_unregister(p3cs_context ctx, void *host_ptr) {
  if (ctx == nullptr || host_ptr == nullptr) return p3cs_fail(P3CS_ERR_INVALID, "p3cs_host_unregister: bad arguments");
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  for (int o = 0; o < 4; ++o)  // p3cs_animate_vertices remembers the destinations it has resolved
    if (ctx->seen_host[o] == host_ptr) ctx->seen_host[o] = nullptr;
  CUDA_TRY(cudaHostUnregister(host_ptr));
  return P3CS_OK;
}
