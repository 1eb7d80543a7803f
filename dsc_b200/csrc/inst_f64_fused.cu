// Explicit instantiations: double fused four-step launches (first + second pass in one grid).
#include <utility>
#include "fft_dispatch.cuh"
namespace dscfft {
DSC_DEFINE_FUSED(double, true)
DSC_DEFINE_FUSED(double, false)
}
