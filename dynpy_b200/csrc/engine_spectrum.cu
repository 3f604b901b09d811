// Engine instantiation: accumulated spectrum -> real part of its transform (the summed ACF).
#include "engine_run.cuh"
#include "kernels.h"
namespace dynpy {
cudaError_t forward_spectrum_real(const SpectrumLoader& ld, const EpiParams& ep, i64 n_fft, i64 M, int nslots,
                                  cd* T, i64 T_cap, const cd* tw, int n_sm, cudaStream_t st) {
    return run_forward<SpectrumLoader, EPI_WRITE_REAL>(ld, ep, n_fft, M, nslots, T, T_cap, tw, n_sm, st);
}
}  // namespace dynpy
