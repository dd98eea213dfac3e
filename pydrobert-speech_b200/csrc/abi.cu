// abi.cu -- library-wide pieces of the C ABI: error string, version, device probing.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace pds {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list args;
  va_start(args, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, args);
  va_end(args);
}

int sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

}  // namespace pds

extern "C" const char* pds_last_error(void) { return pds::g_error; }

extern "C" int pds_version(void) { return 100; }  // 0.1.0

extern "C" int pds_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
