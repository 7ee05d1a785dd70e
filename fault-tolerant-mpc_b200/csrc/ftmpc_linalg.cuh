// ftmpc_linalg.cuh -- block-cooperative dense kernels on row-major matrices with leading dimension ld:
// in-place Cholesky of the condensed Hessian and J = L^-T.  (MUMPS does the equivalent factorisation
// for IPOPT inside the reference's nlpsol call, ft_mpc/controllers/spiraling_mpc.py:230,346.)
#pragma once
#include "ftmpc_block.cuh"

namespace ftmpc {

// A (n x n, lower triangle significant) -> L in the lower triangle (diagonal included).
// Returns 0 on success, k+1 if pivot k is not safely positive.  Upper triangle is left untouched.
template <class Blk>
FT_HD int chol_lower(Blk& blk, int n, int ld, double* A, double piv_tol) {
    const int tid = blk.tid(), nt = blk.nthreads();
    for (int k = 0; k < n; ++k) {
        const double akk = A[(size_t)k * ld + k];
        if (!(akk > piv_tol)) return k + 1;
        const double lkk = sqrt(akk), inv = 1.0 / lkk;
        blk.sync();                                   // every thread has read A[k][k]
        for (int i = k + 1 + tid; i < n; i += nt) A[(size_t)i * ld + k] *= inv;
        if (tid == 0) A[(size_t)k * ld + k] = lkk;
        blk.sync();
        const int w = n - k - 1;
        for (int idx = tid; idx < w * w; idx += nt) {
            const int i = k + 1 + idx / w, j = k + 1 + idx % w;
            if (j <= i) A[(size_t)i * ld + j] -= A[(size_t)i * ld + k] * A[(size_t)j * ld + k];
        }
        blk.sync();
    }
    return 0;
}

// L in the lower triangle of A  ->  A := L^-T (upper triangular), strict lower triangle zeroed.
// dg: scratch of n doubles.
template <class Blk>
FT_HD void tri_inv_transpose(Blk& blk, int n, int ld, double* A, double* dg) {
    const int tid = blk.tid(), nt = blk.nthreads();
    for (int i = tid; i < n; i += nt) dg[i] = A[(size_t)i * ld + i];
    blk.sync();
    // thread j: column j of L^-1, written as row j of the upper triangle
    for (int j = tid; j < n; j += nt) {
        double* xr = A + (size_t)j * ld;
        xr[j] = 1.0 / dg[j];
        for (int i = j + 1; i < n; ++i) {
            const double* li = A + (size_t)i * ld;
            double acc = 0.0;
            for (int k = j; k < i; ++k) acc += li[k] * xr[k];
            xr[i] = -acc / dg[i];
        }
    }
    blk.sync();
    for (int idx = tid; idx < n * n; idx += nt) {
        const int i = idx / n, j = idx % n;
        if (j < i) A[(size_t)i * ld + j] = 0.0;
    }
    blk.sync();
}

}  // namespace ftmpc
