// Explicit instantiations: double TMA-fed four-step launches (fft_tma.cuh).
#include <utility>
#include "fft_dispatch.cuh"
#if !defined(DSC_EMUL)
namespace dscfft {
DSC_DEFINE_TMA(double, true)
DSC_DEFINE_TMA(double, false)
}
#endif
