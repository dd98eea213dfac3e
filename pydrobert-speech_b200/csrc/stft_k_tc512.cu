// One slice of the kernel instantiations of stft_device.cuh (see the pickers declared there).
#include "stft_device.cuh"

namespace pds {
KernelFn pick_tc_512(bool power, int dtype, int mode) { return pick_tc_n<512>(power, dtype, mode); }
}  // namespace pds
