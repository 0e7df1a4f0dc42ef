// vmm.h -- shareable device allocations (see vmm.cu)
#ifndef AMMSB_VMM_H_
#define AMMSB_VMM_H_
#include <stddef.h>
#include <stdint.h>

struct VmmAlloc {
  unsigned long long ptr = 0;     // CUdeviceptr
  size_t size = 0, granularity = 0;
  unsigned long long handle = 0;  // CUmemGenericAllocationHandle
  bool imported = false;
};

size_t vmm_rounded_size(int device, size_t bytes, size_t* granularity);
int vmm_alloc(int device, size_t bytes, VmmAlloc* out);
int vmm_export_fd(const VmmAlloc& a, int* fd);
int vmm_import_fd(int device, int fd, size_t bytes, VmmAlloc* out);
// same process: let another device of this process access the mapping (direct peer access)
int vmm_grant(int device, const VmmAlloc& a);
int vmm_free(VmmAlloc* a);
#endif
