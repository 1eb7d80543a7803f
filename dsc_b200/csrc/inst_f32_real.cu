// Explicit instantiations: float packed-real transforms (rfft forward, irfft inverse).
#include <utility>
#include "fft_dispatch.cuh"
namespace dscfft {
DSC_DEFINE_TABLE(float, true, MODE_R2C, false)
DSC_DEFINE_TABLE(float, false, MODE_C2R, false)
DSC_DEFINE_TABLE(float, true, MODE_FILTER, false)
}
