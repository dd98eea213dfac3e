// One slice of the kernel instantiations: the tcgen05 transform kernel (stft_umma.cuh).
#include "stft_umma.cuh"

namespace pds {
template <int MT>
static KernelFn pick_umma_mt(bool power, int dtype) {
  if (power) return dtype == PDS_I16 ? stft_umma_kernel<true, short, MT> : stft_umma_kernel<true, float, MT>;
  return dtype == PDS_I16 ? stft_umma_kernel<false, short, MT> : stft_umma_kernel<false, float, MT>;
}
KernelFn pick_umma(bool power, int dtype, int mt) {
  if (mt <= 2) return pick_umma_mt<2>(power, dtype);
  if (mt <= 3) return pick_umma_mt<3>(power, dtype);
  return pick_umma_mt<4>(power, dtype);
}
}  // namespace pds
