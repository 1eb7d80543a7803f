// Explicit instantiations: double, MODE_FAST (dense last-axis lines), both directions.
#include <utility>
#include "fft_dispatch.cuh"
namespace dscfft {
DSC_DEFINE_TABLE(double, true, MODE_FAST, false)
DSC_DEFINE_TABLE(double, false, MODE_FAST, false)
}
