// Throughput of the legacy warp-level tensor-core path (mma.sync, SASS HMMA) on sm_100a, per SM, as a
// function of the number of resident warps -- the measurement behind the filter-bank design notes in
// DESIGN.md.  Every warp runs ILP independent accumulator chains; build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__device__ __forceinline__ void mma(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  if (KIND == 0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else if (KIND == 1) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
  } else if (KIND == 2) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}

template <int KIND, int ILP>
__global__ void bench(float* out, int iters, long long* cycles) {
  float c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  const unsigned a = threadIdx.x * 0x3c00u + 0x3f800000u, b = 0x3f800000u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) mma<KIND>(c[i], a, a + 1, a + 2, a + 3, b, b + 1);
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// dependent FFMA2 / FFMA chains for comparison (one warp-instruction = 32 or 64 lane-FMAs)
template <int PACKED, int ILP>
__global__ void bench_fma(float* out, int iters, long long* cycles) {
  unsigned long long v[ILP];
  float f[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x + i, f[i] = threadIdx.x + i;
  const unsigned long long m = 0x3f8000003f800000ull;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (PACKED) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v[i]) : "l"(m));
      else asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(1.0001f));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += f[i] + (float)v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMallocManaged(&cyc, sizeof(long long));
  const int iters = 4096;
  const char* names[4] = {"tf32 m16n8k8", "tf32 m16n8k4", "bf16 m16n8k16", "f16 m16n8k16"};
  const double flops[4] = {2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 16, 2.0 * 16 * 8 * 16};
  for (int kind = 0; kind < 4; ++kind)
    for (int warps = 1; warps <= 16; warps *= 2) {
      constexpr int ILP = 4;
      switch (kind) {
        case 0: bench<0, ILP><<<148, 32 * warps>>>(out, iters, cyc); break;
        case 1: bench<1, ILP><<<148, 32 * warps>>>(out, iters, cyc); break;
        case 2: bench<2, ILP><<<148, 32 * warps>>>(out, iters, cyc); break;
        default: bench<3, ILP><<<148, 32 * warps>>>(out, iters, cyc); break;
      }
      cudaDeviceSynchronize();
      const double per_sm = (double)*cyc / ((double)iters * ILP * warps);
      printf("%-14s warps/SM=%2d ILP=%d : %.2f cycles per MMA per SM, %.1f flop/cycle/SM, single chain latency n/a\n", names[kind],
             warps, ILP, per_sm, flops[kind] / per_sm);
    }
  for (int kind = 0; kind < 4; ++kind) {  // latency: one warp, one dependent chain
    switch (kind) {
      case 0: bench<0, 1><<<1, 32>>>(out, iters, cyc); break;
      case 1: bench<1, 1><<<1, 32>>>(out, iters, cyc); break;
      case 2: bench<2, 1><<<1, 32>>>(out, iters, cyc); break;
      default: bench<3, 1><<<1, 32>>>(out, iters, cyc); break;
    }
    cudaDeviceSynchronize();
    printf("%-14s dependent-chain latency: %.1f cycles\n", names[kind], (double)*cyc / iters);
  }
  for (int packed = 0; packed < 2; ++packed)
    for (int warps = 4; warps <= 16; warps *= 2) {
      if (packed) bench_fma<1, 8><<<148, 32 * warps>>>(out, iters, cyc);
      else bench_fma<0, 8><<<148, 32 * warps>>>(out, iters, cyc);
      cudaDeviceSynchronize();
      const double per_sm = (double)*cyc / ((double)iters * 8 * warps);
      printf("%-14s warps/SM=%2d ILP=8 : %.3f cycles per warp-instruction per SM, %.1f lane-FMA/cycle/SM\n",
             packed ? "fma.f32x2" : "fma.f32", warps, per_sm, (packed ? 64 : 32) / per_sm);
    }
  return 0;
}
