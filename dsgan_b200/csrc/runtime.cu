// Runtime plumbing of the C ABI: error string, launch counter, device check.
#include "common.cuh"
#include "../../include/dsgan_b200.h"
#include <stdarg.h>
#include <string.h>

namespace dsgan {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
}  // namespace dsgan

extern "C" {
int dsgan_abi_version(void) { return DSGAN_ABI_VERSION; }
const char* dsgan_last_error(void) { return dsgan::g_err; }
unsigned long long dsgan_launch_count(void) { return dsgan::g_launches; }

int dsgan_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { dsgan::set_error("no CUDA device: %s", cudaGetErrorString(e)); return 1; }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) { dsgan::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return 1; }
  if (p.major != 10) {
    dsgan::set_error("dsgan_b200 needs an sm_100 (B200) device, found sm_%d%d (%s); there is no fallback path",
                     p.major, p.minor, p.name);
    return 1;
  }
  return 0;
}

int dsgan_memset(void* p, int byte_value, size_t bytes, void* stream) {
  cudaError_t e = cudaMemsetAsync(p, byte_value, bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) { dsgan::set_error("memset: %s", cudaGetErrorString(e)); return 1; }
  ++dsgan::g_launches;
  return 0;
}
}
