// One slice of the kernel instantiations of stft_device.cuh (see the pickers declared there).
#include "stft_device.cuh"

namespace pds {

template <int MODE>
static KernelFn fused_mode(bool power, int dtype) {
  if (power) return dtype == PDS_I16 ? stft_fused_kernel<512, true, short, MODE> : stft_fused_kernel<512, true, float, MODE>;
  return dtype == PDS_I16 ? stft_fused_kernel<512, false, short, MODE> : stft_fused_kernel<512, false, float, MODE>;
}

KernelFn pick_fused512(bool power, int dtype, int mode) {
  switch (mode) {
    case kRows13: return fused_mode<kRows13>(power, dtype);
    case kRows16: return fused_mode<kRows16>(power, dtype);
    default: return fused_mode<kRowsAny>(power, dtype);
  }
}

KernelFn pick_ws512(bool power, int mode) {
  switch (mode) {
    case kRows13: return power ? stft_ws_kernel<512, true, kRows13> : stft_ws_kernel<512, false, kRows13>;
    case kRows16: return power ? stft_ws_kernel<512, true, kRows16> : stft_ws_kernel<512, false, kRows16>;
    default: return power ? stft_ws_kernel<512, true, kRowsAny> : stft_ws_kernel<512, false, kRowsAny>;
  }
}

KernelFn pick_direct(bool power, int dtype) {
  if (power) return dtype == PDS_I16 ? stft_direct_kernel<true, short> : stft_direct_kernel<true, float>;
  return dtype == PDS_I16 ? stft_direct_kernel<false, short> : stft_direct_kernel<false, float>;
}
}  // namespace pds
