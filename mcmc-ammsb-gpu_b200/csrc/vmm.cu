// vmm.cu -- device allocations that another process can map with full-size (2 MB) pages.
//
// Why not cudaIpcOpenMemHandle: measured on B200, a row gather over NVLink from a cudaIpc-imported
// 8 GB shard runs at 195 GB/s, the same gather through a direct peer mapping at 735 GB/s (tools/
// peer_probe.py) -- the legacy IPC import maps the peer memory with small pages and the gather
// becomes bound by address translation.  Memory created with the virtual-memory-management
// driver API and shared as a POSIX file descriptor is mapped by the importer at the allocation
// granularity.  The driver entry points are resolved through the runtime
// (cudaGetDriverEntryPoint), so the library does not link libcuda.
#include <cuda.h>
#include <unistd.h>

#include "common.cuh"
#include "vmm.h"

namespace {
struct Driver {
  CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*memUnmap)(CUdeviceptr, size_t);
  CUresult (*memAddressFree)(CUdeviceptr, size_t);
  CUresult (*memRelease)(CUmemGenericAllocationHandle);
  CUresult (*memExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*memImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*memGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  bool ok = false;
};

template <class F>
bool Resolve(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess ||
      p == nullptr)
    return false;
  *fn = reinterpret_cast<F>(p);
  return true;
}

Driver* GetDriver() {
  static Driver d;
  static bool tried = false;
  if (!tried) {
    tried = true;
    d.ok = Resolve("cuMemCreate", &d.memCreate) && Resolve("cuMemAddressReserve", &d.memAddressReserve) &&
           Resolve("cuMemMap", &d.memMap) && Resolve("cuMemSetAccess", &d.memSetAccess) &&
           Resolve("cuMemUnmap", &d.memUnmap) && Resolve("cuMemAddressFree", &d.memAddressFree) &&
           Resolve("cuMemRelease", &d.memRelease) && Resolve("cuMemExportToShareableHandle", &d.memExport) &&
           Resolve("cuMemImportFromShareableHandle", &d.memImport) &&
           Resolve("cuMemGetAllocationGranularity", &d.memGranularity);
  }
  return d.ok ? &d : nullptr;
}

CUmemAllocationProp PropFor(int device) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof prop);
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

#define VMM_CHECK(expr)                                                                      \
  do {                                                                                       \
    CUresult _r = (expr);                                                                    \
    if (_r != CUDA_SUCCESS) {                                                                \
      ammsb_set_error(std::string(#expr) + ": CUresult " + std::to_string((int)_r) + " (" + \
                      __FILE__ + ":" + std::to_string(__LINE__) + ")");                      \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

int MapForDevice(Driver* d, int device, VmmAlloc* a) {
  VMM_CHECK(d->memAddressReserve(&a->ptr, a->size, a->granularity, 0, 0));
  VMM_CHECK(d->memMap(a->ptr, a->size, 0, a->handle, 0));
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof acc);
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  VMM_CHECK(d->memSetAccess(a->ptr, a->size, &acc, 1));
  return 0;
}
}  // namespace

size_t vmm_rounded_size(int device, size_t bytes, size_t* granularity) {
  Driver* d = GetDriver();
  size_t g = size_t(2) << 20;
  if (d) {
    CUmemAllocationProp prop = PropFor(device);
    size_t q = 0;
    if (d->memGranularity(&q, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && q) g = q;
  }
  if (granularity) *granularity = g;
  return (bytes + g - 1) / g * g;
}

int vmm_alloc(int device, size_t bytes, VmmAlloc* out) {
  Driver* d = GetDriver();
  AMMSB_REQUIRE(d != nullptr, "CUDA virtual-memory-management driver API is not available");
  AMMSB_CHECK_CUDA(cudaSetDevice(device));
  AMMSB_CHECK_CUDA(cudaFree(0));  // make sure the primary context exists and is current
  VmmAlloc a;
  a.size = vmm_rounded_size(device, bytes ? bytes : 1, &a.granularity);
  a.imported = false;
  CUmemAllocationProp prop = PropFor(device);
  VMM_CHECK(d->memCreate(&a.handle, a.size, &prop, 0));
  if (MapForDevice(d, device, &a)) return 1;
  *out = a;
  return 0;
}

int vmm_export_fd(const VmmAlloc& a, int* fd) {
  Driver* d = GetDriver();
  AMMSB_REQUIRE(d != nullptr, "CUDA virtual-memory-management driver API is not available");
  int f = -1;
  VMM_CHECK(d->memExport(&f, a.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *fd = f;
  return 0;
}

int vmm_import_fd(int device, int fd, size_t bytes, VmmAlloc* out) {
  Driver* d = GetDriver();
  AMMSB_REQUIRE(d != nullptr, "CUDA virtual-memory-management driver API is not available");
  AMMSB_CHECK_CUDA(cudaSetDevice(device));
  AMMSB_CHECK_CUDA(cudaFree(0));
  VmmAlloc a;
  a.size = vmm_rounded_size(device, bytes ? bytes : 1, &a.granularity);
  a.imported = true;
  VMM_CHECK(d->memImport(&a.handle, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  if (MapForDevice(d, device, &a)) return 1;
  *out = a;
  return 0;
}

int vmm_grant(int device, const VmmAlloc& a) {
  Driver* d = GetDriver();
  AMMSB_REQUIRE(d != nullptr, "CUDA virtual-memory-management driver API is not available");
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof acc);
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  VMM_CHECK(d->memSetAccess(a.ptr, a.size, &acc, 1));
  return 0;
}

int vmm_free(VmmAlloc* a) {
  Driver* d = GetDriver();
  if (!d || !a->ptr) return 0;
  d->memUnmap(a->ptr, a->size);
  d->memAddressFree(a->ptr, a->size);
  d->memRelease(a->handle);
  a->ptr = 0;
  return 0;
}
