// Error convention of the C ABI (include/drvae_b200.h): entry points return 0 on success and a
// non-zero status otherwise; the message is kept per thread and read with drvae_last_error().
// Nothing in the library throws or aborts.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

namespace drvae {
char* error_buffer();  // thread-local, 512 bytes
inline int set_error(const char* msg) {
  char* b = error_buffer();
  strncpy(b, msg, 511);
  b[511] = 0;
  return 1;
}
inline int set_cuda_error(const char* what, cudaError_t err) {
  char* b = error_buffer();
  snprintf(b, 512, "%s: %s", what, cudaGetErrorString(err));
  return 2;
}
}  // namespace drvae

#define DRVAE_CUDA_OK(expr)                                         \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) return drvae::set_cuda_error(#expr, _e); \
  } while (0)
