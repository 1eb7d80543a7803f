// Explicit instantiations: float, MODE_FAST (dense last-axis lines), both directions.
#include <utility>
#include "fft_dispatch.cuh"
namespace dscfft {
DSC_DEFINE_TABLE(float, true, MODE_FAST, false)
DSC_DEFINE_TABLE(float, false, MODE_FAST, false)
}
