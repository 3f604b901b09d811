// Engine instantiation: two even power spectra per transform -> their two ACFs (ragged dipolar pairs).
#include "engine_run.cuh"
#include "kernels.h"
namespace dynpy {
cudaError_t forward_spectrum_real2(const SpectrumPairLoader& ld, const EpiParams& ep, i64 n_fft, i64 M, cd* T,
                                   i64 T_cap, const cd* tw, int n_sm, cudaStream_t st) {
    return run_forward<SpectrumPairLoader, EPI_WRITE_REAL2>(ld, ep, n_fft, M, 1, T, T_cap, tw, n_sm, st);
}
}  // namespace dynpy
