// Engine instantiation: planar series -> accumulated power spectrum (dipolar, spin-rotation,
// generic wiener_khinchin batches).
#include "engine_run.cuh"
#include "kernels.h"
namespace dynpy {
cudaError_t forward_planar_acc(const PlanarLoader& ld, const EpiParams& ep, i64 n_fft, i64 M, int nslots, cd* T,
                               i64 T_cap, const cd* tw, int n_sm, cudaStream_t st) {
    return run_forward<PlanarLoader, EPI_ACC_POWER>(ld, ep, n_fft, M, nslots, T, T_cap, tw, n_sm, st);
}
// second pass only (T already filled by a fused column kernel)
cudaError_t rows_acc_two_pass(const EpiParams& ep, cd* T, i64 n_fft, int nslots, int M1, const cd* tw, int n_sm,
                              cudaStream_t st) {
    const int target = 4 * n_sm;
    i64 items = n_fft / nslots;
    int ys = ceil_div_i(target, (i64)M1 * nslots);
    if (ys > items) ys = (int)items;
    if (ys < 1) ys = 1;
    TLoader tl;
    tl.T = T; tl.M1 = M1; tl.f_base = 0;
    return launch_rows_L<DYNPY_ROW_LEN, true, TLoader, EPI_ACC_POWER>(tl, ep, tw, 0, n_fft, ys * nslots, M1, st);
}
}  // namespace dynpy
