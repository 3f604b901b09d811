// Engine instantiation: planar series -> accumulated power spectrum (dipolar, spin-rotation,
// generic wiener_khinchin batches).
#include "engine_run.cuh"
#include "kernels.h"
namespace dynpy {
cudaError_t forward_planar_acc(const PlanarLoader& ld, const EpiParams& ep, i64 n_fft, i64 M, int nslots, cd* T,
                               i64 T_cap, const cd* tw, int n_sm, cudaStream_t st) {
    return run_forward<PlanarLoader, EPI_ACC_POWER>(ld, ep, n_fft, M, nslots, T, T_cap, tw, n_sm, st);
}
}  // namespace dynpy
