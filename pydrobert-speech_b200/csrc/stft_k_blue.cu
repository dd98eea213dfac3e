// One slice of the kernel instantiations of stft_device.cuh (see the pickers declared there).
#include "stft_device.cuh"

namespace pds {
KernelFn pick_bluestein(bool power, int dtype) {
  if (power) return dtype == PDS_I16 ? stft_bluestein_kernel<true, short> : stft_bluestein_kernel<true, float>;
  return dtype == PDS_I16 ? stft_bluestein_kernel<false, short> : stft_bluestein_kernel<false, float>;
}
}  // namespace pds
