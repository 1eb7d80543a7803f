// Explicit instantiations: float, MODE_C2C, forward=true, both block shapes.  One TU per
// group keeps the build parallel (each holds ~2x15 fully unrolled kernels).
#include <utility>
#include "fft_dispatch.cuh"
namespace dscfft {
DSC_DEFINE_TABLE(float, true, MODE_C2C, false)
DSC_DEFINE_TABLE(float, true, MODE_C2C, true)
}
