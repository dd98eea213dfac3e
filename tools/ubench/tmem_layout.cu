// Which thread gets which accumulator element?  One tcgen05.mma (kind::f16 on fp16 operands, M = 128,
// N = 128, K = 16, both operands K-major without swizzle) writes D[m][n] = 256 m + n exactly; the warps
// read it back with tcgen05.ld.16x256b.x4 and the kernel dumps every register, then transposes an 8 x 8
// block of packed bf16 pairs with movmatrix.  The host prints the mapping the STFT kernel relies on:
//   register 4 i + {0, 1} of thread T = D[lane0 + T / 4][col0 + 8 i + 2 (T % 4) + {0, 1}],
//   register 4 i + {2, 3}             = D[lane0 + T / 4 + 8][same columns].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_layout tmem_layout.cu
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

__global__ void __launch_bounds__(256, 1) layout_kernel(float* out, uint32_t* moved, int* status) {
  constexpr int M = 128, N = 128, K = 16;
  constexpr uint32_t kLboA = 2048, kLboB = 2064;  // the second one is padded like the kernel's sample operand
  __shared__ __align__(1024) uint8_t s_a[2 * kLboA];
  __shared__ __align__(1024) uint8_t s_b[2 * kLboB];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar)));
  for (int i = tid; i < M * K; i += 256) {
    const int m = i / K, k = i % K;
    const float v = k == 0 ? (float)m : (k == 9 ? 0.f : 0.f);
    *reinterpret_cast<__half*>(s_a + (k / 8) * kLboA + (m / 8) * 128 + (m % 8) * 16 + (k % 8) * 2) = __float2half(v);
  }
  for (int i = tid; i < N * K; i += 256) {
    const int n = i / K, k = i % K;
    const float v = k == 0 ? 256.f : 0.f;
    *reinterpret_cast<__half*>(s_b + (k / 8) * kLboB + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) = __float2half(v);
  }
  __syncthreads();
  // second K column: A[m][9] = 1, B[n][9] = n
  for (int i = tid; i < M; i += 256)
    *reinterpret_cast<__half*>(s_a + 1 * kLboA + (i / 8) * 128 + (i % 8) * 16 + 1 * 2) = __float2half(1.f);
  for (int i = tid; i < N; i += 256)
    *reinterpret_cast<__half*>(s_b + 1 * kLboB + (i / 8) * 128 + (i % 8) * 16 + 1 * 2) = __float2half((float)i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
                           (static_cast<uint32_t>(M >> 4) << 24);
    const uint64_t a = smem_desc(smem_u32(s_a), kLboA, 128), b = smem_desc(smem_u32(s_b), kLboB, 128);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(0) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&s_bar)) : "memory");
    const long long t0 = clock64();
    for (;;) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&s_bar)), "r"(0) : "memory");
      if (ok) break;
      if (clock64() - t0 > 200000000LL) { *status = 1; break; }
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warps 0-3: lanes 32 w + 0 .. 15, warps 4-7: lanes 32 (w - 4) + 16 .. 31; columns 32 .. 63
  {
    const uint32_t lane0 = 32 * (warp & 3) + 16 * (warp >> 2);
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + (lane0 << 16) + 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
    // movmatrix: thread T holds the pair (row T / 4, columns 2 (T % 4), 2 (T % 4) + 1) as 16-bit values
    const uint32_t src = (uint32_t)((lane / 4) * 8 + 2 * (lane % 4)) | ((uint32_t)((lane / 4) * 8 + 2 * (lane % 4) + 1) << 16);
    uint32_t dst;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(dst) : "r"(src));
    if (warp == 0) moved[lane] = dst;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128));
}

int main() {
  float* d_out; uint32_t* d_moved; int* d_status;
  cudaMalloc(&d_out, 256 * 16 * sizeof(float));
  cudaMalloc(&d_moved, 32 * sizeof(uint32_t));
  cudaMalloc(&d_status, sizeof(int));
  cudaMemset(d_status, 0, sizeof(int));
  layout_kernel<<<1, 256>>>(d_out, d_moved, d_status);
  const cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { std::printf("error: %s\n", cudaGetErrorString(err)); return 1; }
  std::vector<float> out(256 * 16);
  std::vector<uint32_t> moved(32);
  int status = 0;
  cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
  cudaMemcpy(moved.data(), d_moved, moved.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  cudaMemcpy(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost);
  if (status) { std::printf("mbarrier wait timed out\n"); return 1; }
  int bad = 0;
  for (int w = 0; w < 8; ++w)
    for (int t = 0; t < 32; ++t)
      for (int j = 0; j < 16; ++j) {
        const int lane0 = 32 * (w & 3) + 16 * (w >> 2);
        const int row = lane0 + t / 4 + ((j & 2) ? 8 : 0), col = 32 + 8 * (j / 4) + 2 * (t % 4) + (j & 1);
        const float want = 256.f * row + col, got = out[(w * 32 + t) * 16 + j];
        if (got != want && bad++ < 16)
          std::printf("warp %d thread %2d reg %2d: D[%d][%d] (value %g), expected D[%d][%d]\n", w, t, j,
                      (int)got / 256, (int)got % 256, got, row, col);
      }
  std::printf("tcgen05.ld.16x256b.x4 (fp16 operands): %s\n", bad ? "mapping differs" : "mapping as assumed, values exact");
  int bad_mov = 0;
  for (int t = 0; t < 32; ++t) {
    // transposed: thread T holds (row T / 4 of the transpose) = source column T / 4, source rows 2 (T % 4), + 1
    const uint32_t want = (uint32_t)((2 * (t % 4)) * 8 + t / 4) | ((uint32_t)((2 * (t % 4) + 1) * 8 + t / 4) << 16);
    if (moved[t] != want && bad_mov++ < 8) std::printf("movmatrix thread %2d: got %08x expected %08x\n", t, moved[t], want);
  }
  std::printf("movmatrix.m8n8.trans.b16: %s\n", bad_mov ? "mapping differs" : "mapping as assumed");
  return bad != 0 || bad_mov != 0;
}
