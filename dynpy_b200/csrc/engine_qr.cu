// Engine instantiation: EFG tensors -> R2,m spherical components -> accumulated power spectrum.
#include "engine_run.cuh"
#include "kernels.h"
namespace dynpy {
cudaError_t forward_qr_acc(const QrLoader& ld, const EpiParams& ep, i64 n_fft, i64 M, int nslots, cd* T, i64 T_cap,
                           const cd* tw, int n_sm, cudaStream_t st) {
    return run_forward<QrLoader, EPI_ACC_POWER>(ld, ep, n_fft, M, nslots, T, T_cap, tw, n_sm, st);
}
}  // namespace dynpy
