// Micro-benchmark: issue throughput of packed FP32x2 (FFMA2/FADD2) vs scalar FFMA/FADD on sm_100a,
// and whether ptxas folds half-swaps / broadcasts of the 64-bit operands into operand modifiers.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

template <int MODE>
__global__ void bench(float* out, int iters, float seed) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
  const float m = 1.0001f, c = 0.5f;
  if (MODE == 0) {  // 16 independent scalar FFMA chains
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
  } else if (MODE == 1) {  // 8 independent FFMA2 chains (same flops)
    uint64_t v[8], mm = pack(m, m), cc = pack(c, c);
    for (int i = 0; i < 8; ++i) v[i] = pack(a[2 * i], a[2 * i + 1]);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fma2(v[i], mm, cc);
    for (int i = 0; i < 8; ++i) unpack(v[i], a[2 * i], a[2 * i + 1]);
  } else if (MODE == 2) {  // 16 scalar FADD chains
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = a[i] + c;
  } else if (MODE == 3) {  // 8 FADD2 chains
    uint64_t v[8], cc = pack(c, c);
    for (int i = 0; i < 8; ++i) v[i] = pack(a[2 * i], a[2 * i + 1]);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = add2(v[i], cc);
    for (int i = 0; i < 8; ++i) unpack(v[i], a[2 * i], a[2 * i + 1]);
  } else if (MODE == 4) {  // complex multiply by a constant twiddle with swapped operand, packed
    uint64_t v[8];
    const float wr = 0.9238795f, wi = -0.3826834f;
    const uint64_t wrr = pack(wr, wr), wii = pack(-wi, wi);
    for (int i = 0; i < 8; ++i) v[i] = pack(a[2 * i], a[2 * i + 1]);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float lo, hi;
        unpack(v[i], lo, hi);
        const uint64_t sw = pack(hi, lo);            // (im, re)
        uint64_t t;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(sw), "l"(wii));  // (-wi*im, wi*re)
        v[i] = fma2(v[i], wrr, t);                   // (re*wr - im*wi, im*wr + re*wi)
      }
    for (int i = 0; i < 8; ++i) unpack(v[i], a[2 * i], a[2 * i + 1]);
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(int iters) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  bench<MODE><<<148 * 8, 256>>>(out, 10, 1.f);
  cudaEventRecord(e0);
  bench<MODE><<<148 * 8, 256>>>(out, iters, 1.f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(out);
  return ms;
}

int main() {
  const int iters = 20000;
  const double lanes = 148.0 * 8 * 256 * 16.0 * iters;  // scalar-equivalent ops
  float t0 = run<0>(iters), t1 = run<1>(iters), t2 = run<2>(iters), t3 = run<3>(iters), t4 = run<4>(iters);
  printf("FFMA  x16 : %.3f ms  %.1f Gop/s (scalar fma lanes)\n", t0, lanes / t0 / 1e6);
  printf("FFMA2 x8  : %.3f ms  %.1f Gop/s\n", t1, lanes / t1 / 1e6);
  printf("FADD  x16 : %.3f ms  %.1f Gop/s\n", t2, lanes / t2 / 1e6);
  printf("FADD2 x8  : %.3f ms  %.1f Gop/s\n", t3, lanes / t3 / 1e6);
  printf("cmul packed (mul2+fma2 per complex, 8 per iter): %.3f ms  %.1f G cmul/s\n", t4,
         148.0 * 8 * 256 * 8.0 * iters / t4 / 1e6);
  return 0;
}
