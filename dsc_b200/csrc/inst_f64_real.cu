// Explicit instantiations: double packed-real transforms (rfft forward, irfft inverse).
#include <utility>
#include "fft_dispatch.cuh"
namespace dscfft {
DSC_DEFINE_TABLE(double, true, MODE_R2C, false)
DSC_DEFINE_TABLE(double, false, MODE_C2R, false)
DSC_DEFINE_TABLE(double, true, MODE_FILTER, false)
}
