// Explicit instantiations: double two-pass transforms along a non-last axis (both passes as column tiles).
#include <utility>
#include "fft_dispatch.cuh"
namespace dscfft {
DSC_DEFINE_COLUMNS(double, true)
DSC_DEFINE_COLUMNS(double, false)
}
