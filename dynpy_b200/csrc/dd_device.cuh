// Device helpers shared by the dipolar kernels: numpy-compatible wrap, the minimum-image
// searches (literal 27-image loop and the per-axis form) and the Y2m / r^-3 words.
#pragma once
#include <cuda_runtime.h>

namespace dynpy {

// np.mod(x, L) (universe.py:301-304 via distances.modv): C fmod, then the sign fix-up numpy
// applies (npy_divmod): a non-zero remainder with the wrong sign gets +L (may round to L).
// The C fmod itself is computed without the library's iterative routine when the quotient is small:
// q = trunc(x / L) is off by at most one, r = fma(-q, L, x) is EXACT for the right q (the remainder of
// fmod is always representable), and a wrong q shows as |r| >= |L| or a sign opposite to x.
__device__ __forceinline__ double fmod_exact(double x, double L, double inv_L) {
    const double t = x * inv_L;  // approximate quotient (inv_L = 1/L): q below may be off by one either way
    if (!(fabs(t) < 1.0e12)) return fmod(x, L);  // huge quotients, inf, nan: library routine
    double q = trunc(t);
    double r = fma(-q, L, x);
    const double aL = fabs(L);
    if (fabs(r) >= aL) {  // |q| one too small
        q += (t < 0.0) ? -1.0 : 1.0;
        r = fma(-q, L, x);
    } else if (r != 0.0 && ((r < 0.0) != (x < 0.0))) {  // |q| one too large
        q -= (t < 0.0) ? -1.0 : 1.0;
        r = fma(-q, L, x);
    }
    return r == 0.0 ? copysign(0.0, x) : r;
}
__device__ __forceinline__ double np_mod(double x, double L, double inv_L) {
    double m = fmod_exact(x, L, inv_L);
    if (m != 0.0) {
        if ((L < 0.0) != (m < 0.0)) m = __dadd_rn(m, L);
    } else {
        m = copysign(0.0, L);
    }
    return m;
}
__device__ __forceinline__ double np_mod(double x, double L) { return np_mod(x, L, 1.0 / L); }
// ------------------------------------------------------------------ minimum image
// Literal restatement of the candidate loop of pdist_ortho (distances.py:203-233): 27 images
// in (aa,bb,cc) order, pxi = xi + aa*a, d = pxi - xj, r2 = dx*dx + dy*dy + dz*dz left to
// right, strict '<' against a running best that starts at dmax^2.  __d*_rn blocks FMA
// contraction so every comparison sees the same bits numba/LLVM produces.
struct MinImage {
    double dx, dy, dz, r2;
    int prj;
    bool found;
};
__device__ __forceinline__ MinImage min_image27(double xi, double yi, double zi, double xj, double yj, double zj,
                                                double a, double b, double c, double dmax2) {
    MinImage r;
    r.dx = r.dy = r.dz = 0.0;
    r.r2 = dmax2;
    r.prj = 0;
    r.found = false;
    int prj = 0;
#pragma unroll
    for (int aa = -1; aa <= 1; ++aa) {
        const double dpx = __dsub_rn(__dadd_rn(xi, __dmul_rn((double)aa, a)), xj);
        const double sx = __dmul_rn(dpx, dpx);
#pragma unroll
        for (int bb = -1; bb <= 1; ++bb) {
            const double dpy = __dsub_rn(__dadd_rn(yi, __dmul_rn((double)bb, b)), yj);
            const double sxy = __dadd_rn(sx, __dmul_rn(dpy, dpy));
#pragma unroll
            for (int cc = -1; cc <= 1; ++cc) {
                const double dpz = __dsub_rn(__dadd_rn(zi, __dmul_rn((double)cc, c)), zj);
                const double r2 = __dadd_rn(sxy, __dmul_rn(dpz, dpz));
                if (r2 < r.r2) {
                    r.dx = dpx; r.dy = dpy; r.dz = dpz; r.r2 = r2; r.prj = prj; r.found = true;
                }
                ++prj;
            }
        }
    }
    return r;
}

// Per-axis form used inside the fused dipolar generator: the minimum over the 27 sums equals
// the sum of the per-axis minima bit for bit (fp add is monotone), so presence (r2 < rmax^2)
// and dr are identical to min_image27; only the image picked under an exact rounding tie
// of the 27 sums could differ.
__device__ __forceinline__ void axis_min(double xi, double xj, double a, double& d, double& d2) {
    const double dm = __dsub_rn(__dsub_rn(xi, a), xj);
    const double d0 = __dsub_rn(xi, xj);
    const double dp = __dsub_rn(__dadd_rn(xi, a), xj);
    const double sm = __dmul_rn(dm, dm), s0 = __dmul_rn(d0, d0), sp = __dmul_rn(dp, dp);
    d = dm; d2 = sm;
    if (s0 < d2) { d = d0; d2 = s0; }
    if (sp < d2) { d = dp; d2 = sp; }
}

// ------------------------------------------------------------------ dipolar words
// ddrax.py:72-126 in Cartesian closed form (no arccos/atan2/sincos):
//   Y20 = 1/4 sqrt(5/pi) (3 z^2/r^2 - 1)
//   Y21 = -1/2 sqrt(15/2pi) z (x + i y)/r^2        Y22 = 1/4 sqrt(15/2pi) (x + i y)^2/r^2
//   r^-3 ;  F2m = sqrt(4 pi/5) Y2m r^-3
struct DDWords {
    double y20, y21r, y21i, y22r, y22i, r3, f20, f21r, f21i, f22r, f22i;
};
__device__ __forceinline__ DDWords dd_words(double dx, double dy, double dz, double r2) {
    const double C20 = 0.31539156525252000603;   // (1/4) sqrt(5/pi)
    const double C21 = -0.77254840404638362;     // -(1/2) sqrt(15/(2 pi))
    const double C22 = 0.38627420202319181;      // (1/4) sqrt(15/(2 pi))
    const double KF = 1.5853309190424043;        // sqrt(4 pi/5)
    DDWords w;
    double rinv = rsqrt(r2);
    rinv = rinv * fma(-0.5 * r2, rinv * rinv, 1.5);  // one Newton step: r^-1 to ~1 ulp
    const double ir2 = rinv * rinv;
    w.r3 = ir2 * rinv;
    w.y20 = C20 * (3.0 * dz * dz * ir2 - 1.0);
    const double zr = dz * ir2;
    w.y21r = C21 * zr * dx;
    w.y21i = C21 * zr * dy;
    w.y22r = C22 * (dx * dx - dy * dy) * ir2;
    w.y22i = C22 * 2.0 * dx * dy * ir2;
    const double k = KF * w.r3;
    w.f20 = k * w.y20; w.f21r = k * w.y21r; w.f21i = k * w.y21i; w.f22r = k * w.y22r; w.f22i = k * w.y22i;
    return w;
}


}  // namespace dynpy
