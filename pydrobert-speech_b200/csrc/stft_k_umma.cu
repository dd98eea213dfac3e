// One slice of the kernel instantiations: the tcgen05 transform kernel (stft_umma.cuh).
#include "stft_umma.cuh"

namespace pds {
template <int NT>
static KernelFn pick_umma_nt(bool power, int dtype) {
  if (power) return dtype == PDS_I16 ? stft_umma_kernel<true, short, NT> : stft_umma_kernel<true, float, NT>;
  return dtype == PDS_I16 ? stft_umma_kernel<false, short, NT> : stft_umma_kernel<false, float, NT>;
}
KernelFn pick_umma(bool power, int dtype, int nt) {
  if (nt <= 3) return pick_umma_nt<3>(power, dtype);
  if (nt <= 5) return pick_umma_nt<5>(power, dtype);
  return pick_umma_nt<8>(power, dtype);
}
}  // namespace pds
