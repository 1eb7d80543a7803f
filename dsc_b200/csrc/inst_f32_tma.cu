// Explicit instantiations: float TMA-fed four-step launches (fft_tma.cuh).
#define DSC_F32X2 1      // issue-bound kernels: packed fp32x2 arithmetic (fft_math.cuh)
#include <utility>
#include "fft_dispatch.cuh"
#if !defined(DSC_EMUL)
namespace dscfft {
DSC_DEFINE_TMA(float, true)
DSC_DEFINE_TMA(float, false)
}
#endif
