// Engine instantiation: planar series -> per-series power spectrum (ragged dipolar pairs).
#include "engine_run.cuh"
#include "kernels.h"
namespace dynpy {
cudaError_t forward_planar_power(const PlanarLoader& ld, const EpiParams& ep, i64 n_fft, i64 M, int nslots, cd* T,
                                 i64 T_cap, const cd* tw, int n_sm, cudaStream_t st) {
    return run_forward<PlanarLoader, EPI_WRITE_POWER>(ld, ep, n_fft, M, nslots, T, T_cap, tw, n_sm, st);
}
}  // namespace dynpy
