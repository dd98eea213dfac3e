// One slice of the kernel instantiations of stft_device.cuh (see the pickers declared there).
#include "stft_device.cuh"

namespace pds {
KernelFn pick_tc_256(bool power, int dtype, int mode) { return pick_tc_n<256>(power, dtype, mode); }
}  // namespace pds
