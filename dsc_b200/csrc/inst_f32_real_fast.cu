// Explicit instantiations: float packed-real transforms on dense last-axis lines (bandwidth path).
#include <utility>
#include "fft_dispatch.cuh"
namespace dscfft {
DSC_DEFINE_TABLE(float, true, MODE_R2C_FAST, false)
DSC_DEFINE_TABLE(float, false, MODE_C2R_FAST, false)
}
