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

#if defined(__CUDACC__)
namespace ftmpc {
// ---------------------------------------------------------------------------------------------------------
// CUDA-block specialisations (same results up to rounding; chosen by overload resolution for CudaBlock).
// ---------------------------------------------------------------------------------------------------------

// Blocked right-looking Cholesky, panel width 8, matrix resident in shared memory.
//   per panel:  (1) every thread factors the 8x8 diagonal block redundantly in registers (no barrier, no
//                   single-thread critical section), the thread owning row r solves its 8 panel entries;
//               (2) trailing update A22 -= P P' with 4x4 register tiles.
// Two barriers per panel (30 at n = 120) instead of three per column.
#define FTMPC_CHOL_NB 8
__device__ __forceinline__ int chol_lower(CudaBlock& blk, int n, int ld, double* A, double piv_tol) {
    constexpr int NB = FTMPC_CHOL_NB;
    const int tid = blk.tid(), nt = blk.nthreads();
    for (int c0 = 0; c0 < n; c0 += NB) {
        const int nb = (n - c0 < NB) ? (n - c0) : NB;
        // (1a) diagonal block -> registers (lower triangle; rows >= nb are padded with the identity)
        double l[NB][NB];
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if (j <= i) l[i][j] = (i < nb) ? A[(size_t)(c0 + i) * ld + c0 + j] : ((i == j) ? 1.0 : 0.0);
        double inv[NB];
        int bad = 0;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            double d = l[j][j];
#pragma unroll
            for (int k = 0; k < NB; ++k) if (k < j) d -= l[j][k] * l[j][k];
            if (j < nb && !(d > piv_tol) && bad == 0) bad = c0 + j + 1;
            const double rs = rsqrt(d);
            inv[j] = rs;
            l[j][j] = d * rs;
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                if (i > j) {
                    double v = l[i][j];
#pragma unroll
                    for (int k = 0; k < NB; ++k) if (k < j) v -= l[i][k] * l[j][k];
                    l[i][j] = v * rs;
                }
            }
        }
        if (bad) return bad;                 // identical in every thread
        // (1b) panel rows
        for (int r = c0 + tid; r < n; r += nt) {
            double* ar = A + (size_t)r * ld + c0;
            const int ri = r - c0;
            if (ri >= NB) {
                double x[NB];
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    double v = (j < nb) ? ar[j] : 0.0;
#pragma unroll
                    for (int k = 0; k < NB; ++k) if (k < j) v -= x[k] * l[j][k];
                    x[j] = v * inv[j];
                }
#pragma unroll
                for (int j = 0; j < NB; ++j) if (j < nb) ar[j] = x[j];
            }
        }
        blk.sync();
        blk.mark(PH_CHOL_PANEL);
        // the factored diagonal block is stored only now: slower warps may still have been reading it above
        if (tid < nb) {
            double* ar = A + (size_t)(c0 + tid) * ld + c0;
#pragma unroll
            for (int i = 0; i < NB; ++i)
                if (i == tid) {
#pragma unroll
                    for (int j = 0; j < NB; ++j) if (j <= i) ar[j] = l[i][j];
                }
        }
        // (2) trailing update A22 -= P P' with 4x4 register tiles.  A tile takes the INTERLEAVED row sets
        //     I(ti) = {r0 + ti + T a}, I(tj) = {r0 + tj + T b}: neighbouring lanes (consecutive tj) then read
        //     neighbouring rows (odd leading dimension -> no bank conflicts; contiguous 4-row tiles would put a
        //     warp's loads on 4 of the 16 banks).  Pairs ti >= tj cover the lower triangle once: an entry with
        //     i < j is stored at its mirror (j, i), which no other tile owns.
        const int r0 = c0 + nb;
        const int w = n - r0;
        if (w > 0) {
            const int T = (w + 3) >> 2;
            const int ntiles = T * (T + 1) / 2;
            for (int tile = tid; tile < ntiles; tile += nt) {
                int ti = (int)((sqrt(8.0 * tile + 1.0) - 1.0) * 0.5);
                while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
                while (ti * (ti + 1) / 2 > tile) --ti;
                const int tj = tile - ti * (ti + 1) / 2;
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
                const double* pi[4];
                const double* pj[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int ri = r0 + ti + T * a, rj = r0 + tj + T * a;
                    pi[a] = A + (size_t)((ri < n) ? ri : n - 1) * ld + c0;
                    pj[a] = A + (size_t)((rj < n) ? rj : n - 1) * ld + c0;
                }
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    if (k < nb) {
                        double ai[4], aj[4];
#pragma unroll
                        for (int a = 0; a < 4; ++a) { ai[a] = pi[a][k]; aj[a] = pj[a][k]; }
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) acc[a][b] += ai[a] * aj[b];
                    }
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int i = r0 + ti + T * a, j = r0 + tj + T * b;
                        if (i < n && j < n && (ti != tj || a >= b)) {
                            if (j <= i) A[(size_t)i * ld + j] -= acc[a][b];
                            else A[(size_t)j * ld + i] -= acc[a][b];
                        }
                    }
            }
        }
        blk.sync();
        blk.mark(PH_CHOL_SYRK);
    }
    return 0;
}

// L (lower triangle of A)  ->  A := L^-T in the upper triangle, strict lower triangle zeroed.
// Column c of X = L^-1 is independent of the others and is written transposed into ROW c of the upper
// triangle, which never overlaps L: no staging buffer and no barrier until the end.  Each column is handled
// by two adjacent lanes that split the k-range of the 8-row block dot products and combine with a shuffle.
__device__ __forceinline__ void tri_inv_transpose(CudaBlock& blk, int n, int ld, double* A, double* dg) {
    const int tid = blk.tid(), nt = blk.nthreads();
    for (int i = tid; i < n; i += nt) dg[i] = 1.0 / A[(size_t)i * ld + i];
    blk.sync();
    const int half = tid & 1;
    const unsigned pair_mask = 3u << (tid & 30);       // the two lanes of a pair run identical trip counts
    for (int cbase = 0; cbase < n; cbase += nt / 2) {
        const int c = cbase + (tid >> 1);
        const bool live = c < n;                       // pairs stay together: the shuffle below needs both lanes
        const int cc = live ? c : n - 1;
        double* xr = A + (size_t)cc * ld;              // X[k][c] lives at xr[k], k >= c
        for (int i0 = cc; i0 < n; i0 += 8) {
            // acc[a] = sum_{k=c}^{i0-1} L[i0+a][k] X[k][c]   (k-range split between the two lanes)
            double acc[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) acc[a] = 0.0;
            const int len = i0 - cc, mid = cc + ((len + 1) >> 1);
            const int k0 = half ? mid : cc, k1 = half ? i0 : mid;
            const double* lrow[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) lrow[a] = A + (size_t)((i0 + a < n) ? i0 + a : n - 1) * ld;
            int k = k0;
            for (; k + 3 < k1; k += 4) {              // 36 loads in flight before the first FMA
                double xk[4], lv[8][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) xk[u] = xr[k + u];
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int u = 0; u < 4; ++u) lv[a][u] = lrow[a][k + u];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int a = 0; a < 8; ++a) acc[a] += lv[a][u] * xk[u];
            }
            for (; k < k1; ++k) {
                const double xk = xr[k];
#pragma unroll
                for (int a = 0; a < 8; ++a) acc[a] += lrow[a][k] * xk;
            }
#pragma unroll
            for (int a = 0; a < 8; ++a) acc[a] += __shfl_xor_sync(pair_mask, acc[a], 1);
            // 8x8 triangular solve within the block (both lanes compute it, lane `half == 0` stores)
            double x[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int i = i0 + a;
                double v = ((i == cc) ? 1.0 : 0.0) - acc[a];
#pragma unroll
                for (int b = 0; b < 8; ++b) if (b < a) v -= ((i < n) ? lrow[a][i0 + b] : 0.0) * x[b];
                x[a] = (i < n) ? v * dg[i] : 0.0;
            }
            if (live && half == 0) {
#pragma unroll
                for (int a = 0; a < 8; ++a) if (i0 + a < n) xr[i0 + a] = x[a];
            }
            __syncwarp(pair_mask);
        }
    }
    blk.sync();
    for (int i = 1 + (tid >> 5); i < n; i += (nt >> 5))          // strict lower triangle <- 0, one warp per row
        for (int j = tid & 31; j < i; j += 32) A[(size_t)i * ld + j] = 0.0;
    blk.sync();
}
}  // namespace ftmpc
#endif
