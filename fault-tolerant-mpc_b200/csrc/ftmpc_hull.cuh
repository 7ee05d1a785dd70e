// ftmpc_hull.cuh -- input-bound polytope of a fault set on the device.
//
// The reference builds  A_h u <= b_h  as the Qhull convex hull of the 2^(#healthy) corner wrenches
// (ft_mpc/controllers/tools/input_bounds.py:43-76, ~1 s per fault set).  The set is the zonotope
//     { D f_fault + sum_{i healthy} u_i D[:, i],  0 <= u_i <= max_thrust },
// so every facet normal is orthogonal to five independent generators: one CTA per fault set enumerates the C(m,5) <= 4368
// generator subsets, takes the null vector of each 5x6 matrix by cofactors, merges duplicate hyperplanes, and evaluates the
// support function  b = n . D f_fault + max_thrust sum_i max(0, n . D[:, i]).  Same facets and the same canonical row
// order (lexicographic on [A | -b] rounded to 1e-9) as the host routine `zonotope_facets` (input_bounds.py), which is the
// checker of this kernel.  Only full-rank generator sets are handled here (status 1 otherwise; the host routine treats the
// flat cases).  The reference's own row order is an artefact of Qhull's rounding noise and is NOT reproduced.
#pragma once
#include "ftmpc_block.cuh"
#include "ftmpc.h"

namespace ftmpc {

#if defined(__CUDACC__)
#define FTMPC_HULL_MAXCAND 512      /* appended candidates (with duplicates from concurrent appends) */
#define FTMPC_HULL_MAXHYP 64        /* distinct hyperplanes */

__device__ __forceinline__ double det3(const double a[3][3]) {
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}
__device__ __forceinline__ double det4(const double a[4][4]) {
    double d = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double m[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) m[r][c] = a[r + 1][c + (c >= j)];
        d += ((j & 1) ? -1.0 : 1.0) * a[0][j] * det3(m);
    }
    return d;
}
__device__ __forceinline__ double det5(const double a[5][5]) {
    double d = 0.0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        double m[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) m[r][c] = a[r + 1][c + (c >= j)];
        d += ((j & 1) ? -1.0 : 1.0) * a[0][j] * det4(m);
    }
    return d;
}

// one CTA per fault set
__global__ void k_hull_facets(const ftmpc_config* __restrict__ cfg, int n_sets, const uint16_t* __restrict__ fault_mask,
                              const double* __restrict__ fault_force, double* __restrict__ table, int32_t* __restrict__ nh_out,
                              int32_t* __restrict__ status_out) {
    const int set = blockIdx.x;
    if (set >= n_sets) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ double G[FTMPC_NTHR][FTMPC_NU];        // healthy generators D[:, i]
    __shared__ double off[FTMPC_NU];                  // D f_fault
    __shared__ double cand[FTMPC_HULL_MAXCAND][FTMPC_NU];
    __shared__ double rows[2 * FTMPC_HULL_MAXHYP][FTMPC_NU + 1];
    __shared__ int order[2 * FTMPC_HULL_MAXHYP];
    __shared__ int m_s, ncand, overflow;
    const uint16_t mask = fault_mask[set];
    if (tid == 0) {
        int m = 0;
        for (int i = 0; i < FTMPC_NTHR; ++i)
            if (!((mask >> i) & 1)) {
                for (int r = 0; r < FTMPC_NU; ++r) G[m][r] = cfg->D[r * FTMPC_NTHR + i];
                ++m;
            }
        m_s = m; ncand = 0; overflow = 0;
        for (int r = 0; r < FTMPC_NU; ++r) {
            double a = 0.0;
            for (int i = 0; i < FTMPC_NTHR; ++i)
                if ((mask >> i) & 1) a += cfg->D[r * FTMPC_NTHR + i] * fault_force[(size_t)set * FTMPC_NTHR + i];
            off[r] = a;
        }
    }
    __syncthreads();
    const int m = m_s;
    // number of 5-subsets and the scale of a null vector (product of five generator norms)
    int ncomb = 0;
    if (m >= 5) ncomb = m * (m - 1) * (m - 2) * (m - 3) / 24 * (m - 4) / 5;       // C(m,5), exact for m <= 16
    double gmax = 0.0;
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int r = 0; r < FTMPC_NU; ++r) s += G[i][r] * G[i][r];
        gmax = fmax(gmax, sqrt(s));
    }
    const double scale = gmax * gmax * gmax * gmax * gmax;
    // one pass of nt subsets at a time with a barrier in between: what a pass appends is visible to the next one, so only
    // the hyperplanes first met inside the same pass can be appended more than once
    for (int base = 0; base < ncomb; base += nt) {
      const int idx = base + tid;
      if (idx < ncomb) {
      do {
        // unrank idx -> c0 < c1 < c2 < c3 < c4 (lexicographic)
        int c[5], rem = idx, lo = 0;
        for (int pos = 0; pos < 5; ++pos) {
            for (int v = lo; v < m; ++v) {
                const int left = m - 1 - v, need = 4 - pos;       // C(left, need) combinations start with v at this position
                int cnt = 1;
                for (int q = 0; q < need; ++q) cnt = cnt * (left - q) / (q + 1);
                if (left < need) cnt = 0;
                if (rem < cnt) { c[pos] = v; lo = v + 1; break; }
                rem -= cnt;
            }
        }
        double nrm[FTMPC_NU], n2 = 0.0;
#pragma unroll
        for (int j = 0; j < FTMPC_NU; ++j) {
            double a[5][5];
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int cc = 0; cc < 5; ++cc) a[r][cc] = G[c[r]][cc + (cc >= j)];
            nrm[j] = ((j & 1) ? -1.0 : 1.0) * det5(a);
            n2 += nrm[j] * nrm[j];
        }
        const double nn = sqrt(n2);
        if (!(nn > 1e-9 * scale)) break;                            // dependent generators
        double sgn = 0.0;
#pragma unroll
        for (int j = 0; j < FTMPC_NU; ++j) {
            nrm[j] /= nn;
            if (sgn == 0.0 && fabs(nrm[j]) > 1e-6) sgn = (nrm[j] > 0.0) ? 1.0 : -1.0;
        }
#pragma unroll
        for (int j = 0; j < FTMPC_NU; ++j) nrm[j] *= sgn;
        // append unless an equal hyperplane is already listed (duplicates from concurrent appends are removed below)
        bool found = false;
        const int cur = *(volatile int*)&ncand;
        for (int k = 0; k < cur && k < FTMPC_HULL_MAXCAND && !found; ++k) {
            double dmx = 0.0;
            for (int j = 0; j < FTMPC_NU; ++j) dmx = fmax(dmx, fabs(((volatile double*)cand[k])[j] - nrm[j]));
            found = dmx < 1e-7;
        }
        if (!found) {
            const int slot = atomicAdd(&ncand, 1);
            if (slot < FTMPC_HULL_MAXCAND) {
                for (int j = 0; j < FTMPC_NU; ++j) cand[slot][j] = nrm[j];
                __threadfence_block();
            } else {
                overflow = 1;
            }
        }
      } while (false);
      }
      __syncthreads();
    }
    if (tid == 0) {
        // serial clean-up of the short list: drop duplicates, build the +/- rows, sort canonically
        int nc = ncand < FTMPC_HULL_MAXCAND ? ncand : FTMPC_HULL_MAXCAND, nu = 0;
        for (int k = 0; k < nc && !overflow; ++k) {
            bool dup = false;
            for (int l = 0; l < nu && !dup; ++l) {
                double dmx = 0.0;
                for (int j = 0; j < FTMPC_NU; ++j) dmx = fmax(dmx, fabs(cand[l][j] - cand[k][j]));
                dup = dmx < 1e-7;
            }
            if (!dup) {
                if (nu >= FTMPC_HULL_MAXHYP) { overflow = 1; break; }
                for (int j = 0; j < FTMPC_NU; ++j) cand[nu][j] = cand[k][j];
                ++nu;
            }
        }
        const double tol = 1e-9;
        int nr = 0;
        for (int k = 0; k < nu; ++k)
            for (int s = 0; s < 2; ++s) {
                const double sg = s ? -1.0 : 1.0;
                double b = 0.0;
                for (int j = 0; j < FTMPC_NU; ++j) {
                    double v = sg * cand[k][j];
                    if (fabs(v) < tol) v = 0.0;
                    rows[nr][j] = v;
                    b += v * off[j];
                }
                for (int i = 0; i < m; ++i) {
                    double d = 0.0;
                    for (int j = 0; j < FTMPC_NU; ++j) d += rows[nr][j] * G[i][j];
                    if (d > 0.0) b += cfg->max_thrust * d;
                }
                rows[nr][FTMPC_NU] = b;
                order[nr] = nr;
                ++nr;
            }
        // insertion sort, lexicographic on [A | -b] rounded to tol
        for (int i = 1; i < nr; ++i) {
            const int oi = order[i];
            int p = i - 1;
            while (p >= 0) {
                const int op = order[p];
                int cmp = 0;
                for (int j = 0; j <= FTMPC_NU && cmp == 0; ++j) {
                    const double a = (j < FTMPC_NU) ? rows[op][j] : -rows[op][FTMPC_NU];
                    const double bb = (j < FTMPC_NU) ? rows[oi][j] : -rows[oi][FTMPC_NU];
                    const double ra = rint(a / tol), rb = rint(bb / tol);
                    cmp = (ra > rb) - (ra < rb);
                }
                if (cmp <= 0) break;
                order[p + 1] = op;
                --p;
            }
            order[p + 1] = oi;
        }
        double* e = table + (size_t)set * FTMPC_HULL_STRIDE;
        const bool ok = m >= FTMPC_NU && nr > 0 && nr <= FTMPC_NH && !overflow;
        for (int r = 0; r < FTMPC_NH; ++r) {
            const bool live = ok && r < nr;
            for (int j = 0; j < FTMPC_NU; ++j) e[r * FTMPC_NU + j] = live ? rows[order[r]][j] : 0.0;
            e[FTMPC_NH * FTMPC_NU + r] = live ? rows[order[r]][FTMPC_NU] : 1e30;       // padding rows are never active
        }
        nh_out[set] = ok ? nr : 0;
        // a bounded zonotope in R^6 needs six independent generators: fewer than 12 rows means a flat (rank-deficient) set
        status_out[set] = (ok && nr >= 2 * FTMPC_NU) ? 0 : 1;
    }
}
#endif  // __CUDACC__

}  // namespace ftmpc
