// One slice of the kernel instantiations of stft_device.cuh (see the pickers declared there).
#include "stft_device.cuh"

namespace pds {

template <int MODE>
static KernelFn w_mode(bool power, int nt) {
  if (power) return nt <= 5 ? stft_w_kernel<true, MODE, 5> : stft_w_kernel<true, MODE, 8>;
  return nt <= 5 ? stft_w_kernel<false, MODE, 5> : stft_w_kernel<false, MODE, 8>;
}

KernelFn pick_w512(bool power, int mode, int nt) {
  switch (mode) {
    case kRows13: return w_mode<kRows13>(power, nt);
    case kRows16: return w_mode<kRows16>(power, nt);
    default: return w_mode<kRowsAny>(power, nt);
  }
}
}  // namespace pds
