// Explicit instantiations: complex64 lines transformed by one thread-block cluster each (fft_cluster.cuh).
#include <utility>
#include "fft_dispatch.cuh"
#if !defined(DSC_EMUL)
namespace dscfft {
DSC_DEFINE_CLUSTER(float, true)
DSC_DEFINE_CLUSTER(float, false)
}
#endif
