// Explicit instantiations: double packed-real transforms on dense last-axis lines (bandwidth path).
#include <utility>
#include "fft_dispatch.cuh"
namespace dscfft {
DSC_DEFINE_TABLE(double, true, MODE_R2C_FAST, false)
DSC_DEFINE_TABLE(double, false, MODE_C2R_FAST, false)
}
