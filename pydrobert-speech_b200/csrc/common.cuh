// common.cuh -- error plumbing and small device helpers shared by the translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/pds_b200.h"

namespace pds {

// thread-local message behind pds_last_error()
void set_error(const char* fmt, ...);

#define PDS_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t err__ = (expr);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      ::pds::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,               \
                       cudaGetErrorString(err__));                                        \
      return PDS_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define PDS_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::pds::set_error(__VA_ARGS__);  \
      return PDS_ERR_INVALID;         \
    }                                 \
  } while (0)

// Switches to `device` for the lifetime of the guard and restores the caller's current device
// afterwards: the C ABI never leaves the calling thread on another GPU than it came with.
class DeviceGuard {
 public:
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev_) != cudaSuccess) prev_ = -1;
    status_ = prev_ == device ? cudaSuccess : cudaSetDevice(device);
    switched_ = status_ == cudaSuccess && prev_ != device;
  }
  ~DeviceGuard() {
    if (switched_ && prev_ >= 0) cudaSetDevice(prev_);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
  cudaError_t status() const { return status_; }

 private:
  int prev_ = -1;
  bool switched_ = false;
  cudaError_t status_ = cudaSuccess;
};

// Number of SMs of the current device (cached per device).
int sm_count(int device);

#ifdef __CUDACC__

// ---- Philox4x32-10 counter-based generator: the dither stream -----------------------------
// Four standard normals per Philox call: samples 4*group .. 4*group+3 of utterance `utt`.
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t utt, uint64_t group) {
  uint32_t c0 = (uint32_t)group, c1 = (uint32_t)(group >> 32), c2 = utt, c3 = 0x5eed5eedu;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  // two Box-Muller pairs on four 32-bit uniforms in (0, 1)
  const float u1 = ((float)c0 + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)c1 + 0.5f) * 2.3283064365386963e-10f;
  const float u3 = ((float)c2 + 0.5f) * 2.3283064365386963e-10f;
  const float u4 = ((float)c3 + 0.5f) * 2.3283064365386963e-10f;
  const float ra = sqrtf(-2.0f * __logf(fmaxf(u1, 1e-12f)));
  const float rb = sqrtf(-2.0f * __logf(fmaxf(u3, 1e-12f)));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * u2, &sa, &ca);
  __sincosf(6.283185307179586f * u4, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

// One standard normal per (seed, utterance, sample) triple, independent of how utterances are
// batched, tiled or sharded over GPUs (reference only pins mean/std).
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t utt, uint64_t sample) {
  const float4 n = philox_normal4(seed, utt, sample >> 2);
  const uint32_t which = (uint32_t)sample & 3u;
  return which == 0 ? n.x : (which == 1 ? n.y : (which == 2 ? n.z : n.w));
}

template <typename T>
__device__ __forceinline__ float load_sample(const T* p, long long i) {
  return (float)p[i];
}

// symmetric ("half-sample") reflection of an out-of-range index, numpy.pad(..., 'symmetric')
__device__ __forceinline__ long long reflect_index(long long g, long long len) {
  if (g >= 0 && g < len) return g;
  const long long period = 2 * len;
  long long m = g % period;
  if (m < 0) m += period;
  return m < len ? m : period - 1 - m;
}

#endif  // __CUDACC__

}  // namespace pds
