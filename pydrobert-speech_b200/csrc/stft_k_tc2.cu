// One slice of the kernel instantiations of stft_device.cuh (see the pickers declared there).
#include "stft_device.cuh"

namespace pds {
KernelFn pick_tc2_512(bool power, int dtype, int mode) { return pick_tc2_n<512>(power, dtype, mode); }
KernelFn pick_tc2_probe(int which) {
  if (which == 1) return stft_tc2_kernel<512, true, float, kRows13, 1>;
  if (which == 2) return stft_tc2_kernel<512, true, float, kRows13, 2>;
  return nullptr;
}
}  // namespace pds
