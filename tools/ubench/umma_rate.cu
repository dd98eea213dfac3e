// Issue rate of the Blackwell tensor-core path (tcgen05.mma, SASS UTCxMMA) with both operands in
// shared memory, at the tile shapes a batched-DFT formulation of the frame transform would use
// (M = 128, N = 32 ... 256, kind::f16 on bf16 / kind::tf32), the TMEM read-back rate (tcgen05.ld),
// and how much shared-memory bandwidth is left for the other warps while the MMAs stream their
// operands.  The numbers behind DESIGN.md section 9 ("tensor-core transform").  One CTA per SM; the
// products are checked against integers computed on the host before anything is timed.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int kM = 128;
constexpr int kKSteps = 8;            // MMAs per accumulation chain: K = 128 (bf16) or 64 (tf32)
constexpr int kThreads = 256;         // warps 0-3 read TMEM, warps 4-7 hammer shared memory on request
constexpr long long kSpinLimit = 400000000LL;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// K-major, no swizzle: 8-row x 16-byte core matrices; LBO = step between the two 16-byte halves
// of one MMA's K extent, SBO = step between 8-row groups (both in bytes, encoded >> 4)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version of sm_100
  return d;
}

template <bool TF32>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  if (TF32) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
  }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > kSpinLimit) return false;
  }
}

__host__ __device__ inline int a_value(int m, int k) { return (m * 3 + k * 5) % 7 - 3; }
__host__ __device__ inline int b_value(int n, int k) { return (n * 2 + k) % 5 - 2; }

struct Result {
  long long mma_cycles;     // first issue -> completion of `iters` chains of kKSteps MMAs
  long long issue_cycles;   // first issue -> last issue
  long long ld_cycles;      // `ld_iters` rounds of tcgen05.ld.32x32b.x32 by four warps, one load in flight
  long long ld4_cycles;     // the same with four loads in flight per warp
  long long ld8_cycles;     // ... and two warps per sub-partition
  long long hammer_loads;   // 16-byte shared-memory loads per hammering thread while the MMAs ran
  int error;
};

template <bool TF32, int N>
__global__ void __launch_bounds__(kThreads, 1) umma_bench(int iters, int ld_iters, int hammer, float* out, Result* results) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int kEsize = TF32 ? 4 : 2;
  constexpr int kPerChunk = 16 / kEsize;                    // elements in a 16-byte row of a core matrix
  constexpr int kK = kKSteps * 2 * kPerChunk;               // K of one accumulation chain
  constexpr uint32_t kLboA = kM / 8 * 128, kLboB = N / 8 * 128, kSbo = 128;
  constexpr uint32_t kBytesA = kM * kK * kEsize, kBytesB = N * kK * kEsize;
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + kBytesA;
  uint4* s_h = reinterpret_cast<uint4*>(smem + kBytesA + kBytesB);  // 16 KB for the hammering warps
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ volatile int s_done;
  __shared__ long long s_hammer;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar)));
    s_done = 0;
  }
  for (int i = tid; i < kM * kK; i += kThreads) {
    const int m = i / kK, k = i - m * kK;
    const uint32_t at = (k / kPerChunk) * kLboA + (m / 8) * 128 + (m % 8) * 16 + (k % kPerChunk) * kEsize;
    if (TF32) *reinterpret_cast<float*>(s_a + at) = static_cast<float>(a_value(m, k));
    else *reinterpret_cast<__nv_bfloat16*>(s_a + at) = __float2bfloat16(static_cast<float>(a_value(m, k)));
  }
  for (int i = tid; i < N * kK; i += kThreads) {
    const int n = i / kK, k = i - n * kK;
    const uint32_t at = (k / kPerChunk) * kLboB + (n / 8) * 128 + (n % 8) * 16 + (k % kPerChunk) * kEsize;
    if (TF32) *reinterpret_cast<float*>(s_b + at) = static_cast<float>(b_value(n, k));
    else *reinterpret_cast<__nv_bfloat16*>(s_b + at) = __float2bfloat16(static_cast<float>(b_value(n, k)));
  }
  for (int i = tid; i < 1024; i += kThreads) s_h[i] = make_uint4(i, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (MMA) reads
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = (1u << 4) | ((TF32 ? 2u : 1u) << 7) | ((TF32 ? 2u : 1u) << 10) |
                         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(kM >> 4) << 24);
  const uint32_t a0 = smem_u32(s_a), b0 = smem_u32(s_b), bar = smem_u32(&s_bar);
  Result res = {0, 0, 0, 0, 0, 0, 0};

  // ---- pass 1: one chain, checked against the host --------------------------------------------
  if (tid == 0) {
    for (int ks = 0; ks < kKSteps; ++ks)
      umma<TF32>(tmem, smem_desc(a0 + ks * 2 * kLboA, kLboA, kSbo), smem_desc(b0 + ks * 2 * kLboB, kLboB, kSbo), idesc, ks > 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
    if (!mbar_wait(bar, 0)) res.error = 1;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4 && blockIdx.x == 0) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();

  // ---- pass 2: issue rate -------------------------------------------------------------------
  if (tid == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int ks = 0; ks < kKSteps; ++ks)
        umma<TF32>(tmem + (it & 1) * N, smem_desc(a0 + ks * 2 * kLboA, kLboA, kSbo), smem_desc(b0 + ks * 2 * kLboB, kLboB, kSbo),
                   idesc, 1);
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
    if (!mbar_wait(bar, 1)) res.error = 2;
    const long long t2 = clock64();
    res.mma_cycles = t2 - t0;
    res.issue_cycles = t1 - t0;
    s_done = 1;
  } else if (hammer && warp >= 4) {
    // conflict-free 16-byte loads (512 B per warp instruction) until the MMAs have drained
    uint4 acc = make_uint4(0, 0, 0, 0);
    long long loads = 0;
    while (!s_done) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint4 x = s_h[((warp - 4) * 256 + u * 32 + lane) & 1023];
        acc.x ^= x.x; acc.y += x.y;
      }
      loads += 8;
    }
    if (acc.x == 0x12345678u && acc.y == 77u) out[0] = 1.f;  // keep the loads alive
    if (tid == 128) s_hammer = loads;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- pass 3: TMEM read-back: latency (one load in flight per warp) and rate (four in flight,
  //      by one or two warps per 32-lane sub-partition) ------------------------------------
  {
    uint32_t keep = 0;
    const uint32_t lanes = static_cast<uint32_t>((warp & 3) * 32) << 16;
    if (warp < 4) {
      const long long t0 = clock64();
      for (int it = 0; it < ld_iters; ++it) {
        uint32_t v[32];
        tmem_ld32(tmem + lanes + ((it * 32) & (N - 1) & ~31), v);
        tmem_ld_wait();
        keep ^= v[0] ^ v[31];
      }
      if (tid == 0) res.ld_cycles = clock64() - t0;
    }
    for (int nwarps = 4; nwarps <= 8; nwarps += 4) {
      __syncthreads();
      if (warp < nwarps) {
        const long long t0 = clock64();
        for (int it = 0; it < ld_iters; it += 4) {
          uint32_t v0[32], v1[32], v2[32], v3[32];
          tmem_ld32(tmem + lanes + 0, v0);
          tmem_ld32(tmem + lanes + (32 & (N - 1)), v1);
          tmem_ld32(tmem + lanes + (64 & (N - 1)), v2);
          tmem_ld32(tmem + lanes + (96 & (N - 1)), v3);
          tmem_ld_wait();
          keep ^= v0[0] ^ v1[7] ^ v2[19] ^ v3[31] ^ v0[31] ^ v1[0] ^ v2[1] ^ v3[2];
        }
        if (tid == 0) (nwarps == 4 ? res.ld4_cycles : res.ld8_cycles) = clock64() - t0;
      }
    }
    if (keep == 0x12345678u) out[1] = 2.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    res.hammer_loads = hammer ? s_hammer : 0;
    results[blockIdx.x] = res;
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

template <bool TF32, int N>
static int run(int sms, int iters, int ld_iters) {
  constexpr int kEsize = TF32 ? 4 : 2;
  constexpr int kK = kKSteps * 32 / kEsize;
  const size_t smem = static_cast<size_t>(kM + N) * kK * kEsize + 16384;
  auto kernel = umma_bench<TF32, N>;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  float* d_out;
  Result* d_res;
  cudaMalloc(&d_out, kM * N * sizeof(float));
  cudaMalloc(&d_res, sms * sizeof(Result));
  std::vector<float> out(kM * N);
  std::vector<Result> res(sms);
  int bad = 0;
  for (int hammer = 0; hammer < 2; ++hammer) {
    cudaMemset(d_res, 0, sms * sizeof(Result));
    kernel<<<sms, kThreads, smem>>>(iters, ld_iters, hammer, d_out, d_res);
    const cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      std::printf("%s N=%d: %s\n", TF32 ? "tf32" : "bf16", N, cudaGetErrorString(err));
      return 1;
    }
    cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(res.data(), d_res, sms * sizeof(Result), cudaMemcpyDeviceToHost);
    int wrong = 0;
    for (int m = 0; m < kM; ++m)
      for (int n = 0; n < N; ++n) {
        int want = 0;
        for (int k = 0; k < kK; ++k) want += a_value(m, k) * b_value(n, k);
        if (out[m * N + n] != static_cast<float>(want) && wrong++ < 4)
          std::printf("  mismatch D[%d][%d] = %g, expected %d\n", m, n, out[m * N + n], want);
      }
    double mma = 0, issue = 0, ld = 0, ld4 = 0, ld8 = 0, ham = 0;
    int errors = 0;
    for (const Result& r : res) {
      mma += r.mma_cycles; issue += r.issue_cycles; ld += r.ld_cycles; ld4 += r.ld4_cycles; ld8 += r.ld8_cycles; ham += r.hammer_loads; errors += r.error != 0;
    }
    mma /= sms; issue /= sms; ld /= sms; ld4 /= sms; ld8 /= sms; ham /= sms;
    const double n_mma = static_cast<double>(iters) * kKSteps;
    const double per_mma = mma / n_mma;
    const double flops_per_cycle = 2.0 * kM * N * (kK / kKSteps) / per_mma;
    const double operand_bytes = static_cast<double>(kM + N) * (kK / kKSteps) * kEsize;  // read per MMA
    std::printf("%s M=128 N=%3d K=%2d%s: %s, %7.2f cycles/MMA (issue %5.2f), %6.0f flop/cycle/SM, operands %5.1f B/cycle",
                TF32 ? "tf32" : "bf16", N, kK / kKSteps, hammer ? " +smem loads" : "            ",
                wrong ? "WRONG" : "exact", per_mma, issue / n_mma, flops_per_cycle, operand_bytes / per_mma);
    if (hammer) std::printf(", other warps' loads %5.1f B/cycle", ham * 128 * 16 / mma);
    else
      std::printf(", tcgen05.ld.32x32b.x32: latency %5.1f cycles; 4 warps x 4 in flight %5.1f B/cycle/SM, 8 warps %5.1f",
                  ld / ld_iters, 16384.0 / (ld4 / ld_iters), 32768.0 / (ld8 / ld_iters));
    std::printf("%s\n", errors ? "  [mbarrier wait timed out]" : "");
    bad += wrong != 0 || errors != 0;
  }
  cudaFree(d_out);
  cudaFree(d_res);
  return bad;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  std::printf("%s, %d SMs, sm_%d%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
  const int sms = prop.multiProcessorCount, iters = 256, ld_iters = 2048;
  int bad = 0;
  bad += run<false, 32>(sms, iters, ld_iters);
  bad += run<false, 64>(sms, iters, ld_iters);
  bad += run<false, 128>(sms, iters, ld_iters);
  bad += run<false, 256>(sms, iters, ld_iters);
  bad += run<true, 32>(sms, iters, ld_iters);
  bad += run<true, 64>(sms, iters, ld_iters);
  bad += run<true, 128>(sms, iters, ld_iters);
  bad += run<true, 256>(sms, iters, ld_iters);
  return bad != 0;
}
