// ftmpc_riccati.cuh -- the factor J of the condensed Hessian WITHOUT forming the condensed Hessian.
//
// The QP of an SQP iteration is the condensed problem of SpiralingController.build_solver
// (ft_mpc/controllers/spiraling_mpc.py:87-214 in its reduced form): H = R_bar + Gamma' Q_bar Gamma over U = (u_0..u_{N-1}),
// and the dual active-set solver (ftmpc_gi.cuh) needs a matrix J with J J' = H^-1 plus the rows X J (X = d x_N / d U).
// Building H (O(N^3) rank-13 block updates), its Cholesky factor (n^3/3) and the explicit triangular inverse (n^3/3)
// were 40 % of a solve at N = 20 and 85 % at N = 100.  H is the Hessian of an optimal-control problem, so it factors
// stage by stage instead:
//
//   backward  P_N = H_term;   F_t = [A_t B_t]' P_{t+1} [A_t B_t] + [[M_t, S_t'], [S_t, R_t]]        (19 x 19)
//             Lam_t = F_uu = C_t C_t'  (6 x 6 Cholesky),  Kh_t = C_t^-1 F_ux,  K_t = C_t^-T Kh_t,
//             P_t = F_xx - Kh_t' Kh_t                                                                 (Riccati recursion)
//   identity  with  v_t = u_t + K_t dx_t  (dx = the state perturbation driven by u):   1/2 U' H U = 1/2 sum_t v_t' Lam_t v_t,
//             i.e.  H = Phi' Lam Phi  with Phi unit block lower triangular, so  J = Phi^-1 blkdiag(C_t^-T)  satisfies
//             J' H J = I.  H is positive definite  <=>  every Lam_t is (same inertia), which is the only thing the
//             Hessian schedule of phase_qp asks of a factorisation attempt.
//   forward   column a = 6 s + j of J is a closed-loop rollout:  u_s = C_s^-T e_j,  dx_{s+1} = B_s u_s,
//             u_t = -K_t dx_t,  dx_{t+1} = A_t dx_t + B_t u_t  (t > s);  rows 6t..6t+5 of the column are u_t, the column of
//             X J is dx_N[0:9], and d = J' ga accumulates on the way.  One thread per column, no barrier.
//
// Cost: 20 k multiply-adds per stage backward (serial over stages, 4 barrier intervals each) + 250 per (column, later
// stage) forward: 0.4 M multiply-adds at N = 20 against 1.7 M for condensing + Cholesky + inverse, and O(N^2) instead of
// O(N^3) -- 8 M against 110 M at N = 100.  The cost gradient g, its augmented-Lagrangian twin ga (two costate
// recursions) ride in the same backward sweep.
//
// Written once against the Block abstraction (ftmpc_block.cuh): the CUDA block and the CPU port run the same code.
#pragma once
#include "ftmpc.h"
#include "ftmpc_block.cuh"

namespace ftmpc {

// FTMPC_RIC_WORK (ftmpc_sqp.cuh): doubles of scratch riccati_factor needs, see the carve below
#define FTMPC_RIC_REC 114        /* per-stage record left in the Wz slot: K_t [6][13], then C_t^-1 [6][6] (lower) */

// 6x6 Cholesky  A = C C'  (lower; A's lower triangle significant) and Ci = C^-1 (lower).  Returns the index + 1 of the
// first pivot that is not safely positive, else 0.  A is overwritten by C.
FT_HD int ric_chol6(double (&A)[6][6], double (&Ci)[6][6], double piv_tol) {
    int bad = 0;
    double inv[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = A[j][j];
#pragma unroll
        for (int m = 0; m < 6; ++m) if (m < j) d -= A[j][m] * A[j][m];
        if (!(d > piv_tol) && !bad) bad = j + 1;
#if defined(__CUDA_ARCH__)
        const double rs = rsqrt(d);
#else
        const double rs = 1.0 / sqrt(d);
#endif
        inv[j] = rs;
        A[j][j] = d * rs;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i > j) {
                double v = A[i][j];
#pragma unroll
                for (int m = 0; m < 6; ++m) if (m < j) v -= A[i][m] * A[j][m];
                A[i][j] = v * rs;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i < j) Ci[i][j] = 0.0;
            else if (i == j) Ci[i][j] = inv[j];
            else {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < 6; ++m) if (m >= j && m < i) v += A[i][m] * Ci[m][j];
                Ci[i][j] = -v * inv[i];
            }
        }
    }
    return bad;
}

// (row, col) of the idx-th entry of a lower triangle stored by rows: idx = row (row + 1) / 2 + col
FT_HD void ric_tri(int idx, int& row, int& col) {
    int r = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
    while ((r + 1) * (r + 2) / 2 <= idx) ++r;
    while (r * (r + 1) / 2 > idx) --r;
    row = r;
    col = idx - r * (r + 1) / 2;
}

#if defined(__CUDACC__)
// ---- backward sweep, CUDA block of 256 threads -------------------------------------------------------------------
// The generic sweep below pays four block barriers per stage, each with a fixed cost (hand-over through shared memory,
// task decode, the slowest warp) that is several times the 13-long multiply-add chains it separates, and during the 6x6
// Cholesky seven of the eight warps wait.  Here a stage is TWO barrier intervals:
//   AB  one group of 10 lanes per column c2 of [A B] (3 groups per warp, warps 0-6): the group forms P [A B][:, c2]
//       (13 entries over 10 lanes), __syncwarp, then its ten entries F[(c2 + j) mod 19][c2], j = 0..9, of the symmetric
//       19 x 19 matrix (the circulant assignment gives every column the same load and covers each pair once).  Warp 7
//       advances the costates / gradients.
//   CD  thread (r, c), c <= r, of the 91 entries of P_t factors Lam = F_uu ITSELF (the 6x6 Cholesky is a serial chain
//       either way; 91 redundant copies cost nothing but otherwise idle issue slots), forward-substitutes the two
//       columns r and c of F_ux alongside, and writes P_t[r][c] = F[r][c] - kh_r . kh_c.  Threads (r, 0) leave Kh[:, r] in
//       the stage record, thread (0, 0) the factor C.
// K_t = C^-T Kh_t and the columns of C_t^-T (what the forward rollouts read) are not on the critical path: one parallel
// pass over all stages after the sweep turns (C, Kh) into them by back substitution.
__device__ __forceinline__ int ric_backward_cuda(CudaBlock& blk, const ftmpc_config& cfg, int N, const double* Jz, double* Wz,
                                                 double theta, bool aug, const double* Cq, double* P, double* PABs, double* F,
                                                 double* cst, double* flg, const double* pre, const double* hau, double* g_out,
                                                 double* ga_out, double* dscale_out) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- AB roles
    const int grp = 3 * warp + lane / 10, j10 = lane % 10;
    const bool ab_on = warp < 7 && lane < 30 && grp < 19;
    const int c2 = ab_on ? grp : 0;
    const int c1 = (c2 + j10) % 19;
    const int hi = c1 > c2 ? c1 : c2, lo = c1 > c2 ? c2 : c1;
    // stage-Hessian entry (hi, lo): constant diagonal, offsets of the two W entries, optional Cq / hull-augmentation slots
    double h_diag = 0.0;
    int h_o1 = -1, h_o2 = -1, h_cq = -1, h_au = -1;
    if (ab_on) {
        if (hi == lo) h_diag = (hi < FTMPC_NE) ? 2.0 * cfg.Q[hi < FTMPC_NE ? hi : 0] : ((hi >= 13) ? 2.0 * cfg.R[hi - 13] : 0.0);
        if (lo >= 6) {
            const int k1 = hi - 6, k2 = lo - 6;
            h_o1 = k1 * 13 + k2; h_o2 = k2 * 13 + k1;
            if (k1 >= 3 && k1 < 7 && k2 >= 3) h_cq = 16 + (k1 - 3) * 4 + (k2 - 3);
            else if (k1 >= 7 && k1 < 10 && k2 >= 3 && k2 < 7) h_cq = 4 + (k1 - 7) * 4 + (k2 - 3);
            if (k2 >= 7) h_au = (k1 - 7) * (k1 - 6) / 2 + k2 - 7;
        }
    }
    // ---- CD roles
    int d_r = 0, d_c = 0;
    ric_tri(tid < 91 ? tid : 0, d_r, d_c);
    const bool cd_on = tid < 91;
    for (int t = N - 1; t >= 0; --t) {
        const double* jz = Jz + (size_t)t * 169;
        double* wz = Wz + (size_t)t * 169;
        // ================= AB
        if (ab_on) {
            const int ra = j10, rb = 10 + j10;                      // rows of P [A B][:, c2]: lanes 0-2 of the group take two
            double va, vb = 0.0;
            if (c2 < 6) {
                const int cc = (c2 < 3) ? c2 : c2 - 3;
                va = (c2 < 3) ? P[ra * 13 + c2] : cfg.dt * P[ra * 13 + cc] + P[ra * 13 + c2];
                if (j10 < 3) vb = (c2 < 3) ? P[rb * 13 + c2] : cfg.dt * P[rb * 13 + cc] + P[rb * 13 + c2];
            } else {
                const double* col = jz + (c2 - 6) * 13;
                const double* pa = P + ra * 13;
                const double* pb = P + ((j10 < 3) ? rb : ra) * 13;
                double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
                for (int k = 0; k < 12; k += 2) {
                    const double x0 = col[k], x1 = col[k + 1];
                    a0 += pa[k] * x0; a1 += pa[k + 1] * x1;
                    b0 += pb[k] * x0; b1 += pb[k + 1] * x1;
                }
                va = a0 + a1 + pa[12] * col[12];
                vb = b0 + b1 + pb[12] * col[12];
            }
            PABs[c2 * 13 + ra] = va;
            if (j10 < 3) PABs[c2 * 13 + rb] = vb;
        } else if (warp == 7 && lane < 19) {
            // costates of stage t and the gradient entries of stage t (kind 0: cost, kind 1: augmentation, only when sigma > 0)
            const double* pr = pre + (size_t)t * 25;
            const int c = lane;
            for (int kd = 0; kd < (aug ? 2 : 1); ++kd) {
                const double* p = cst + ((N - 1 - t) & 1) * 26 + kd * 13;
                double v;
                if (c < 3) v = p[c];
                else if (c < 6) v = p[c] + cfg.dt * p[c - 3];
                else {
                    const double* col = jz + (c - 6) * 13;
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int k = 0; k < 12; k += 2) { a0 += col[k] * p[k]; a1 += col[k + 1] * p[k + 1]; }
                    v = a0 + a1 + col[12] * p[12];
                }
                double* cn = cst + ((N - t) & 1) * 26 + kd * 13;
                if (c < 13) cn[c] = (kd == 0) ? v + pr[c] : v;
                else if (kd == 0) g_out[t * FTMPC_NU + c - 13] = v + pr[c];
                else ga_out[t * FTMPC_NU + c - 13] = v + pr[c + 6];
            }
            if (!aug && c >= 13) ga_out[t * FTMPC_NU + c - 13] = 0.0;
        }
        __syncwarp();
        blk.mark(PH_RIC_AB1);
        if (ab_on) {
            const double* pc = PABs + c2 * 13;
            double v;
            if (c1 < 3) v = pc[c1];
            else if (c1 < 6) v = cfg.dt * pc[c1 - 3] + pc[c1];
            else {
                const double* col = jz + (c1 - 6) * 13;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int k = 0; k < 12; k += 2) { a0 += col[k] * pc[k]; a1 += col[k + 1] * pc[k + 1]; }
                v = a0 + a1 + col[12] * pc[12];
            }
            v += h_diag;
            if (h_o1 >= 0) {
                double hsv = (theta != 0.0) ? theta * 0.5 * (wz[h_o1] + wz[h_o2]) : 0.0;     // theta = 0: W_t was not computed
                if (Cq && h_cq >= 0) hsv += Cq[(size_t)t * FTMPC_CQ + h_cq];
                if (aug && h_au >= 0) hsv += hau[t * 21 + h_au];
                v += hsv;
            }
            F[hi * 19 + lo] = v;
            F[lo * 19 + hi] = v;
        }
        blk.mark(PH_RIC_AB2);
        blk.sync();
        blk.mark(PH_RIC_ABW);
        // ================= CD
        const int par = t & 1;
        if (cd_on) {
            double Lm[6][6], fr[6], fc[6], yr[6], yc[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int j = 0; j < 6; ++j) Lm[i][j] = (j <= i) ? F[(13 + i) * 19 + 13 + j] : 0.0;
                fr[i] = F[(13 + i) * 19 + d_r];
                fc[i] = F[(13 + i) * 19 + d_c];
            }
            double pv = F[d_r * 19 + d_c];
            double dmx = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) dmx = fmax(dmx, fabs(Lm[i][i]));
            const double run = fmax(flg[1 + par], dmx);
            const double piv_tol = 1e-10 * fmax(1.0, run);
            int bad = 0;
            double rsv[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                double d = Lm[j][j];
#pragma unroll
                for (int m = 0; m < 6; ++m) if (m < j) d -= Lm[j][m] * Lm[j][m];
                if (!(d > piv_tol) && !bad) bad = j + 1;
                const double rs = rsqrt(d);
                rsv[j] = rs;
                Lm[j][j] = d * rs;
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    if (i > j) {
                        double v = Lm[i][j];
#pragma unroll
                        for (int m = 0; m < 6; ++m) if (m < j) v -= Lm[i][m] * Lm[j][m];
                        Lm[i][j] = v * rs;
                    }
                }
                double vr = fr[j], vc = fc[j];
#pragma unroll
                for (int m = 0; m < 6; ++m) if (m < j) { vr -= Lm[j][m] * yr[m]; vc -= Lm[j][m] * yc[m]; }
                yr[j] = vr * rs;
                yc[j] = vc * rs;
                pv -= yr[j] * yc[j];
            }
            P[d_r * 13 + d_c] = pv;
            P[d_c * 13 + d_r] = pv;
            if (d_c == 0) {
#pragma unroll
                for (int i = 0; i < 6; ++i) wz[i * 13 + d_r] = yr[i];
                if (d_r == 0) {
                    int o = 120;
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int j = 0; j < 6; ++j) if (j <= i) wz[o++] = Lm[i][j];
#pragma unroll
                    for (int i = 0; i < 6; ++i) wz[141 + i] = rsv[i];
                    if (bad) flg[0] = (double)(FTMPC_NU * t + bad);
                    flg[1 + (par ^ 1)] = run;
                }
            }
        }
        blk.mark(PH_RIC_CD);
        blk.sync();
        blk.mark(PH_RIC_CDW);
        if (flg[0] != 0.0) { *dscale_out = flg[1 + (par ^ 1)]; return (int)flg[0]; }
    }
    *dscale_out = flg[2];                                           // written at t = 0
    // ---- (C, Kh) -> (C^-T columns, K = C^-T Kh): back substitution, one task per (stage, column)
    for (int idx = tid; idx < N * 19; idx += blockDim.x) {
        const int t = idx / 19, e = idx - t * 19;
        double* wz = Wz + (size_t)t * 169;
        double Lm[6][6], rs[6], x[6];
        {
            int o = 120;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) Lm[i][j] = (j <= i) ? wz[o++] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            rs[i] = wz[141 + i];
            x[i] = (e < 13) ? wz[i * 13 + e] : ((e - 13 == i) ? 1.0 : 0.0);
        }
#pragma unroll
        for (int i = 5; i >= 0; --i) {
            double v = x[i];
#pragma unroll
            for (int m = 0; m < 6; ++m) if (m > i) v -= Lm[m][i] * x[m];
            x[i] = v * rs[i];
        }
        if (e < 13) {
#pragma unroll
            for (int i = 0; i < 6; ++i) wz[i * 13 + e] = x[i];
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) wz[78 + (e - 13) * 6 + i] = x[i];        // row j of C^-1 = column j of C^-T
        }
    }
    blk.sync();
    blk.mark(PH_RIC_POST);
    return 0;
}
template <class Blk> struct RicCuda { static constexpr bool value = false; };
template <> struct RicCuda<CudaBlock> { static constexpr bool value = true; };
#endif

// Inputs: Jz [N][13][13] ([col][row]: columns 0-6 = d x+ / d (omega, q), 7-12 = d x+ / d u; the (p, v) columns of A_t are
// [[I, dt I], [0, I], 0]), Wz [N][13][13] stage Hessians of the Lagrangian over (omega, q, u) -- OVERWRITTEN by the stage
// records (K_t, C_t^-1) --, X, U, xref, gradV / hessV of the terminal cost, blend theta, augmentation sigma on the rows
// with lam_prev > 0 (see `condense`), Cq (accelerating references) or nullptr.
// Outputs: rows 0..n-1 and nv..nv+8 of E (columns 0..n-1), s.g, s.ga, s.gi.d = J' ga, *dscale_out = largest diagonal
// entry of the Lam_t.  `work`: FTMPC_RIC_WORK doubles; `lam_stage`: mc doubles of fast scratch for the previous
// multipliers (only touched when sigma > 0).
// Returns 0, or 6 t + j + 1 when pivot j of Lam_t is not safely positive (uniform over the block; outputs invalid).
template <class Blk, class QpS, class Lay>
FT_HD int riccati_factor(Blk& blk, const ftmpc_config& cfg, const Lay& L, const QpS& s, const double* Jz, double* Wz,
                         const double* X, const double* U, const double* xref, const double* gradV, const double* hessV,
                         double theta, double sigma, const double* lam_prev_g, const double* Cq, double* work,
                         double* lam_stage, double* dscale_out, bool want_columns = true) {
    const int N = L.N, n = L.n, ld = L.nv, nv = L.nv, tid = blk.tid(), nt = blk.nthreads();
    double* P = work;                 // [13][13]  cost-to-go Hessian P_{t+1}
    double* PAB = P + 169;            // [13][19]  P_{t+1} [A_t B_t]
    double* F = PAB + 247;            // [19][19]  lower triangle
    double* Kh = F + 361;             // [6][13]   C^-1 F_ux
    double* Cs = Kh + 78;             // [6][6]    C^-1 (lower)
    double* cst = Cs + 36;            // costates: [2 buffers][2 kinds (g, aug)][13]
    double* Ht = cst + 52;            // [9][9] terminal Hessian model (+ augmentation), [9] augmentation of the terminal gradient
    double* tgv = Ht + 81;
    double* flg = tgv + 9;            // [0] failing pivot + 1, [1] running max of diag(Lam), [2] its candidate of the current stage
    // scratch of the set-up in the (still free) E region, behind the staged multipliers:
    //   pre [N][25]  per-stage vectors of the costate / gradient recursion (see below)
    //   hau [N][21]  lower triangle of sigma sum_A a_i a_i' (hull rows of stage t), only when sigma > 0
    //   lists        active rows (lam_prev > 0) of every stage [N][27] and of the terminal set [73] (count first), ints
    // (a horizon of 1 has no room there: it uses the tail of `work`)
    const size_t es_need = (size_t)L.mc + (size_t)60 * N + 40;
    double* es = (es_need <= (size_t)(nv + FTMPC_NE) * nv) ? lam_stage + L.mc : work + 1040;
    double* pre = es;
    double* hau = pre + (size_t)25 * N;
    int* hlist = reinterpret_cast<int*>(hau + (size_t)21 * N);
    int* tlist = hlist + (size_t)27 * N + (N & 1);
    const double* Ah = s.hull;
    const double* lam_prev = lam_prev_g;
    const bool aug = sigma > 0.0;
    // ---- set-up, first interval: everything that comes from global memory is requested at once -- the multipliers (when the
    //      augmentation is on), the per-stage vectors of the costate / gradient recursion
    //        pre[t][0:13]  running-cost gradient joining the costate of stage t (t >= 1)
    //        pre[t][13:19] 2 R (u~_t - rho_t)        pre[t][19:25] sigma sum_A c_i a_i  (hull rows of stage t; below)
    //      (two tasks per thread and trip, loads first: one memory round trip instead of one per pass), and the
    //      un-augmented terminal model  Ht = term_quad + theta (hessV - term_quad)
    if (aug) {
        for (int i = tid; i < L.mc; i += nt) lam_stage[i] = lam_prev_g[i];
        lam_prev = lam_stage;
    }
    if (tid == 0) { flg[0] = 0.0; flg[1] = 0.0; flg[2] = 0.0; s.g[n] = 0.0; s.ga[n] = 0.0; }
    {
        const int ntask = N * 19;
        for (int i0 = tid; i0 < ntask; i0 += 2 * nt) {
            double a[2], b[2], w[2];
            int tt[2], cc[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int idx = i0 + h * nt;
                const bool on = idx < ntask;
                const int t = on ? idx / 19 : 0, c = on ? idx - t * 19 : 0;
                tt[h] = on ? t : -1; cc[h] = c;
                a[h] = 0.0; b[h] = 0.0; w[h] = 0.0;
                if (!on) continue;
                if (c < FTMPC_NE) {
                    if (t > 0) { a[h] = X[t * FTMPC_NX + c]; b[h] = xref[t * FTMPC_NE + c]; w[h] = 2.0 * cfg.Q[c]; }
                } else if (c < 13) {
                    if (t > 0 && Cq) { a[h] = Cq[(size_t)t * FTMPC_CQ + c - FTMPC_NE]; w[h] = 1.0; }
                } else {
                    a[h] = U[t * FTMPC_NU + c - 13];
                    if (Cq) b[h] = Cq[(size_t)t * FTMPC_CQ + 32 + c - 13];
                    w[h] = 2.0 * cfg.R[c - 13];
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (tt[h] >= 0) pre[tt[h] * 25 + cc[h]] = w[h] * (a[h] - b[h]);
        }
        for (int idx = tid; idx < 81; idx += nt) {
            const double q0 = cfg.term_quad[idx];
            Ht[idx] = q0 + theta * (hessV[idx] - q0);
        }
        for (int idx = tid; idx < 9; idx += nt) tgv[idx] = 0.0;
        if (!aug) for (int idx = tid; idx < N * FTMPC_NU; idx += nt) pre[(idx / FTMPC_NU) * 25 + 19 + idx % FTMPC_NU] = 0.0;
    }
    blk.sync();
    if (aug) {
        // active-row lists (one thread per stage, the last thread takes the terminal set)
        for (int t = tid; t <= N; t += nt) {
            if (t < N) {
                int c = 0;
                for (int k = 0; k < FTMPC_NH; ++k)
                    if (lam_prev[t * FTMPC_NH + k] > 0.0) hlist[t * 27 + 1 + c++] = k;
                hlist[t * 27] = c;
            } else {
                int c = 0;
                for (int i = 0; i < FTMPC_NF; ++i)
                    if (lam_prev[FTMPC_NH * N + i] > 0.0) tlist[1 + c++] = i;
                tlist[0] = c;
            }
        }
        blk.sync();
        // augmentation: Ht += sigma sum_A a a',  tgv = sigma sum_A c a  (terminal rows);  per stage sigma sum_A c a and the
        // lower triangle of sigma sum_A a a' (hull rows)
        for (int idx = tid; idx < 90 + N * 27; idx += nt) {
            double v = 0.0;
            if (idx < 90) {
                const int kk = idx / 9, l = idx - kk * 9;
                const int na = tlist[0];
                if (s.tf_val) {                        // <= 2 non-zeros per row of A_f
                    for (int e = 0; e < na; ++e) {
                        const int i = tlist[1 + e];
                        const int k0 = s.tf_idx[2 * i], k1 = s.tf_idx[2 * i + 1];
                        const double v0 = s.tf_val[2 * i], v1 = s.tf_val[2 * i + 1];
                        const double al = (l == k0) ? v0 : ((l == k1) ? v1 : 0.0);
                        const double ak = (kk == k0) ? v0 : ((kk == k1) ? v1 : 0.0);
                        v += (idx < 81) ? ak * al : s.cv[FTMPC_NH * N + i] * al;
                    }
                } else {
                    const double* Af = s.cg ? s.cg->Af : cfg.Af;
                    for (int e = 0; e < na; ++e) {
                        const int i = tlist[1 + e];
                        const double al = Af[i * FTMPC_NE + l];
                        v += (idx < 81) ? Af[i * FTMPC_NE + kk] * al : s.cv[FTMPC_NH * N + i] * al;
                    }
                }
                if (idx < 81) Ht[idx] += sigma * v;
                else tgv[l] = sigma * v;
            } else {
                const int e0 = idx - 90, t = e0 / 27, c = e0 - t * 27;
                const int na = hlist[t * 27];
                if (c < 6) {
                    for (int e = 0; e < na; ++e) {
                        const int k = hlist[t * 27 + 1 + e];
                        v += s.cv[t * FTMPC_NH + k] * Ah[k * FTMPC_NU + c];
                    }
                    pre[t * 25 + 19 + c] = sigma * v;
                } else {
                    int i, j;
                    ric_tri(c - 6, i, j);
                    for (int e = 0; e < na; ++e) {
                        const int k = hlist[t * 27 + 1 + e];
                        v += Ah[k * FTMPC_NU + i] * Ah[k * FTMPC_NU + j];
                    }
                    hau[t * 21 + c - 6] = sigma * v;
                }
            }
        }
        blk.sync();
    }
    for (int idx = tid; idx < 169 + 26; idx += nt) {
        if (idx < 169) {
            const int r = idx / 13, c = idx - r * 13;
            P[idx] = (r < FTMPC_NE && c < FTMPC_NE) ? Ht[r * FTMPC_NE + c] : 0.0;
        } else {
            const int kd = (idx - 169) / 13, r = (idx - 169) - kd * 13;
            cst[kd * 13 + r] = (r < FTMPC_NE) ? (kd == 0 ? gradV[r] : tgv[r]) : 0.0;
        }
    }
    // task decode of this thread's first pass (the triangular index needs a square root: once, not per stage)
    int b_c1, b_c2, d_r, d_c;
    ric_tri(tid, b_c1, b_c2);
    d_r = b_c1; d_c = b_c2;
    double b_diag = 0.0;                   // constant diagonal of the stage Hessian: 2Q on (p, v, omega), 2R on u
    if (tid < 190 && b_c1 == b_c2) b_diag = (b_c1 < FTMPC_NE) ? 2.0 * cfg.Q[b_c1 < FTMPC_NE ? b_c1 : 0] : ((b_c1 >= 13) ? 2.0 * cfg.R[b_c1 - 13] : 0.0);
    blk.sync();
    blk.mark(PH_RIC_PRE);
    bool swept = false;
#if defined(__CUDA_ARCH__)
    if (RicCuda<Blk>::value && nt == 256) {
        // PAB is kept by columns ([19][13]) and F full symmetric in this path; same buffers
        const int bad = ric_backward_cuda(reinterpret_cast<CudaBlock&>(blk), cfg, N, Jz, Wz, theta, aug, Cq, P, PAB, F, cst, flg, pre, hau,
                                          s.g, s.ga, dscale_out);
        if (bad) return bad;
        swept = true;
    }
#endif
    // ---- backward sweep
    for (int t = swept ? -1 : N - 1; t >= 0; --t) {
        const double* jz = Jz + (size_t)t * 169;
        double* wz = Wz + (size_t)t * 169;
        const double* co = cst + ((N - 1 - t) & 1) * 26;            // costates of stage t + 1
        double* cn = cst + ((N - t) & 1) * 26;                      // costates of stage t
        const double* pr = pre + (size_t)t * 25;
        // (A) PAB = P [A B]  (13 x 19; the (p, v) columns of A are trivial: they ride with the first 78 tasks), and the two
        //     costate / gradient products [A B]' lambda
        //     (task ids 169..191 are left empty so that no warp runs both kinds of task one after the other)
        for (int idx = tid; idx < 192 + 38; idx += nt) {
            if (idx >= 169 && idx < 192) continue;
            if (idx < 169) {
                const int cm = idx / 13, r = idx - cm * 13;        // consecutive threads: consecutive rows of one column
                const double* col = jz + cm * 13;
                const double* pr_ = P + r * 13;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int k = 0; k < 12; k += 2) { a0 += pr_[k] * col[k]; a1 += pr_[k + 1] * col[k + 1]; }
                PAB[r * 19 + 6 + cm] = a0 + a1 + pr_[12] * col[12];
                if (cm < 6) PAB[r * 19 + cm] = (cm < 3) ? pr_[cm] : cfg.dt * pr_[cm - 3] + pr_[cm];
            } else {
                const int kd = (idx - 192) / 19, c = (idx - 192) - kd * 19;
                const double* p = co + kd * 13;
                double v;
                if (c < 3) v = p[c];
                else if (c < 6) v = p[c] + cfg.dt * p[c - 3];
                else {
                    const double* col = jz + (c - 6) * 13;
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int k = 0; k < 12; k += 2) { a0 += col[k] * p[k]; a1 += col[k + 1] * p[k + 1]; }
                    v = a0 + a1 + col[12] * p[12];
                }
                if (c < 13) {                                      // costate of stage t (the running-cost gradient joins for t > 0)
                    cn[kd * 13 + c] = (kd == 0) ? v + pr[c] : v;
                } else {
                    const int a = t * FTMPC_NU + c - 13;
                    if (kd == 0) s.g[a] = v + pr[c];
                    else s.ga[a] = v + pr[c + 6];                  // ga - g; combined below
                }
            }
        }
        blk.sync();
        blk.mark(PH_RIC_AB1);
        // (B) F = [A B]' PAB + stage Hessian, lower triangle (c1 >= c2)
        for (int idx = tid; idx < 190; idx += nt) {
            int c1 = b_c1, c2 = b_c2;
            double dg = b_diag;
            if (idx != tid) {
                ric_tri(idx, c1, c2);
                dg = (c1 != c2) ? 0.0 : ((c1 < FTMPC_NE) ? 2.0 * cfg.Q[c1 < FTMPC_NE ? c1 : 0] : ((c1 >= 13) ? 2.0 * cfg.R[c1 - 13] : 0.0));
            }
            double v;
            if (c1 < 3) v = PAB[c1 * 19 + c2];
            else if (c1 < 6) v = cfg.dt * PAB[(c1 - 3) * 19 + c2] + PAB[c1 * 19 + c2];
            else {
                const double* col = jz + (c1 - 6) * 13;
                const double* pc = PAB + c2;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int k = 0; k < 12; k += 2) { a0 += col[k] * pc[k * 19]; a1 += col[k + 1] * pc[(k + 1) * 19]; }
                v = a0 + a1 + col[12] * pc[12 * 19];
            }
            // stage Hessian: blkdiag(2 Q[0:6], theta sym(W) + 2Q on omega + 2R on u + Cq Gauss-Newton + sigma hull rows)
            v += dg;
            if (c2 >= 6) {
                const int k1 = c1 - 6, k2 = c2 - 6;
                double hsv = (theta != 0.0) ? theta * 0.5 * (wz[k1 * 13 + k2] + wz[k2 * 13 + k1]) : 0.0;   // theta = 0: W_t was not computed
                if (Cq) {
                    const double* cq = Cq + (size_t)t * FTMPC_CQ;
                    if (k1 >= 3 && k1 < 7 && k2 >= 3) hsv += cq[16 + (k1 - 3) * 4 + (k2 - 3)];                 // (q, q)
                    else if (k1 >= 7 && k1 < 10 && k2 >= 3 && k2 < 7) hsv += cq[4 + (k1 - 7) * 4 + (k2 - 3)];  // (u~_F, q)
                }
                if (aug && k2 >= 7) hsv += hau[t * 21 + (k1 - 7) * (k1 - 6) / 2 + k2 - 7];
                v += hsv;
            }
            F[c1 * 19 + c2] = v;
        }
        blk.sync();
        blk.mark(PH_RIC_AB2);
        // (C) Lam = F_uu = C C', C^-1; Kh = C^-1 F_ux: the 13 threads of the columns of F_ux factor Lam redundantly
        //     (a 6x6 Cholesky is a serial chain either way) and go on to their forward substitution without a barrier
        for (int c = tid; c < FTMPC_NX; c += nt) {
            double A6[6][6], Ci[6][6];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) A6[i][j] = (j <= i) ? F[(13 + i) * 19 + 13 + j] : 0.0;
            double fx[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) fx[i] = F[(13 + i) * 19 + c];
            double dmx = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) dmx = fmax(dmx, fabs(A6[i][i]));
            const double run = fmax(flg[1], dmx);
            const int bad = ric_chol6(A6, Ci, 1e-10 * fmax(1.0, run));
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < 6; ++m) if (m <= i) v += Ci[i][m] * fx[m];
                Kh[i * 13 + c] = v;
            }
            if (c == 0) {
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = 0; j < 6; ++j) wz[78 + i * 6 + j] = Ci[i][j];
                if (bad) flg[0] = (double)(FTMPC_NU * t + bad);
                flg[2] = run;
            }
        }
        blk.sync();
        blk.mark(PH_RIC_CD);
        if (flg[0] != 0.0) { *dscale_out = flg[2]; return (int)flg[0]; }
        // (D) P_t = F_xx - Kh' Kh  (symmetric; both halves written);  K_t = C^-T Kh -> stage record (off the critical path)
        for (int idx = tid; idx < 96 + 78 + 1; idx += nt) {
            if (idx >= 91 && idx < 96) continue;                   // (warp-aligned task kinds, as in (A))
            if (idx < 91) {
                int r = d_r, c = d_c;
                if (idx != tid) ric_tri(idx, r, c);
                double v = F[r * 19 + c];
#pragma unroll
                for (int i = 0; i < 6; ++i) v -= Kh[i * 13 + r] * Kh[i * 13 + c];
                P[r * 13 + c] = v;
                P[c * 13 + r] = v;
            } else if (idx < 96 + 78) {
                const int e = idx - 96, i = e / 13, c = e - i * 13;
                const double* ci = wz + 78;
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < 6; ++m) if (m >= i) v += ci[m * 6 + i] * Kh[m * 13 + c];
                wz[i * 13 + c] = v;
            } else {
                flg[1] = flg[2];
            }
        }
        blk.sync();
        blk.mark(PH_RIC_CDW);
    }
    if (!swept) *dscale_out = flg[1];
    for (int a = tid; a < n; a += nt) s.ga[a] += s.g[a];
    blk.sync();
    blk.mark(PH_RIC_POST);
    if (!want_columns) return 0;                   // operator form (ric_apply): the stage records are all the QP needs
    // ---- forward: one closed-loop rollout per column of J
    // (a lane pair per column with shuffled halves was measured slower: the rollouts are bound by the shared-memory
    //  instruction rate -- one broadcast load per multiply-add -- and the shuffles add to it; profiles/README.md r02u)
    for (int a = tid; a < n; a += nt) {
        const int st = a / FTMPC_NU, j = a - st * FTMPC_NU;
        double u[FTMPC_NU], dx[FTMPC_NX], dn[FTMPC_NX];
        double dacc = 0.0;
        for (int r = 0; r < FTMPC_NU * st; ++r) s.E[(size_t)r * ld + a] = 0.0;
        {
            const double* ci = Wz + (size_t)st * 169 + 78;          // C^-1 (lower): column j of C^-T = row j of C^-1
            const double* jz = Jz + (size_t)st * 169;
#pragma unroll
            for (int i = 0; i < FTMPC_NU; ++i) {
                u[i] = (i <= j) ? ci[j * 6 + i] : 0.0;
                s.E[(size_t)(FTMPC_NU * st + i) * ld + a] = u[i];
                dacc += u[i] * s.ga[FTMPC_NU * st + i];
            }
#pragma unroll
            for (int r = 0; r < FTMPC_NX; ++r) {
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < FTMPC_NU; ++i) v += jz[(7 + i) * 13 + r] * u[i];
                dx[r] = v;
            }
        }
        for (int t = st + 1; t < N; ++t) {
            const double* kt = Wz + (size_t)t * 169;
            const double* jz = Jz + (size_t)t * 169;
#pragma unroll
            for (int i = 0; i < FTMPC_NU; ++i) {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int k = 0; k < 12; k += 2) { a0 += kt[i * 13 + k] * dx[k]; a1 += kt[i * 13 + k + 1] * dx[k + 1]; }
                u[i] = -(a0 + a1 + kt[i * 13 + 12] * dx[12]);
            }
#pragma unroll
            for (int i = 0; i < FTMPC_NU; ++i) {
                s.E[(size_t)(FTMPC_NU * t + i) * ld + a] = u[i];
                dacc += u[i] * s.ga[FTMPC_NU * t + i];
            }
#pragma unroll
            for (int r = 0; r < FTMPC_NX; ++r) {
                double v = (r < 3) ? dx[r] + cfg.dt * dx[r + 3] : ((r < 6) ? dx[r] : 0.0);
#pragma unroll
                for (int l = 0; l < 7; ++l) v += jz[l * 13 + r] * dx[6 + l];
#pragma unroll
                for (int i = 0; i < FTMPC_NU; ++i) v += jz[(7 + i) * 13 + r] * u[i];
                dn[r] = v;
            }
#pragma unroll
            for (int r = 0; r < FTMPC_NX; ++r) dx[r] = dn[r];
        }
#pragma unroll
        for (int kk = 0; kk < FTMPC_NE; ++kk) s.E[(size_t)(nv + kk) * ld + a] = dx[kk];
        s.gi.d[a] = dacc;
    }
    blk.sync();
    blk.mark(PH_RIC_FWD);
    return 0;
}


// ---- K = E E' as an OPERATOR (long horizons) -------------------------------------------------------------------------
// E = [J ; X J] never has to exist: with the stage records of riccati_factor (K_t, C_t^-1) and the stage Jacobians,
//   out = K v,   v = [v_x (n) ; v_delta ; v_e (9)]   in the extended coordinates of the active-set solver,
// is one adjoint sweep and one closed-loop rollout:
//   backward  mu_N = [v_e ; 0];   y_t = v_x,t + B_t' mu_{t+1},   z_t = C_t^-1 y_t,   mu_t = A_t' mu_{t+1} - K_t' y_t
//   forward   dx_0 = 0;   u_t = C_t^-T z_t - K_t dx_t,   dx_{t+1} = A_t dx_t + B_t u_t;   out_x,t = u_t,  out_e = dx_N[0:9]
// (E' v = blkdiag(C_t^-1) Phi^-T [v_x + X' v_e], E w = Phi^-1 blkdiag(C_t^-T) w), 540 multiply-adds per stage instead of
// the 2 (6N + 10)(6N + 1) of the dense products -- and nothing of size N^2 in memory.  The elastic variable is decoupled:
// out_delta = v_delta / rho_slack.  Written against a block of ANY width (one warp on the device: the sweeps are serial
// over the stages and a warp barrier is all a stage needs).
struct RicOp {
    int N, n, nv;
    const double* Jz;      // [N][169]
    const double* Rec;     // [N][169]  K_t [6][13] at 0, C_t^-1 [6][6] (lower) at 78
    double dt, inv_rho;
    double* scr;           // 72 doubles of (fast) scratch
};
// one stage of the adjoint sweep: y = vx + B' mu, z = C^-1 y -> zout, mu_t = A' mu - K' y -> mn   (two barrier intervals)
template <class WB>
FT_HD void ric_back_step(WB& wb, const double* jz, const double* rec, double dt, const double* vx, const double* mu, double* mn,
                         double* y, double* zout) {
    const int tid = wb.tid(), nt = wb.nthreads();
    for (int i = tid; i < FTMPC_NU; i += nt) {
        const double* col = jz + (7 + i) * 13;
        double a0 = 0.0, a1 = 0.0;
        for (int r = 0; r < 12; r += 2) { a0 += col[r] * mu[r]; a1 += col[r + 1] * mu[r + 1]; }
        y[i] = vx[i] + (a0 + a1 + col[12] * mu[12]);
    }
    wb.sync();
    for (int idx = tid; idx < FTMPC_NU + FTMPC_NX; idx += nt) {
        if (idx < FTMPC_NU) {
            const double* ci = rec + 78 + idx * 6;
            double a = 0.0;
            for (int m = 0; m <= idx; ++m) a += ci[m] * y[m];
            zout[idx] = a;
        } else {
            const int c = idx - FTMPC_NU;
            double a;
            if (c < 3) a = mu[c];
            else if (c < 6) a = mu[c] + dt * mu[c - 3];
            else {
                const double* col = jz + (c - 6) * 13;
                double a0 = 0.0, a1 = 0.0;
                for (int r = 0; r < 12; r += 2) { a0 += col[r] * mu[r]; a1 += col[r + 1] * mu[r + 1]; }
                a = a0 + a1 + col[12] * mu[12];
            }
            for (int i = 0; i < FTMPC_NU; ++i) a -= rec[i * 13 + c] * y[i];
            mn[c] = a;
        }
    }
    wb.sync();
}
// one stage of the closed-loop rollout: u = C^-T z - K dx -> zu (in place of z), dx+ = A dx + B u -> dn
template <class WB>
FT_HD void ric_fwd_step(WB& wb, const double* jz, const double* rec, double dt, double* zu, const double* dx, double* dn, double* ub) {
    const int tid = wb.tid(), nt = wb.nthreads();
    for (int i = tid; i < FTMPC_NU; i += nt) {
        double a = 0.0;
        for (int m = i; m < FTMPC_NU; ++m) a += rec[78 + m * 6 + i] * zu[m];
        double a0 = 0.0, a1 = 0.0;
        for (int k = 0; k < 12; k += 2) { a0 += rec[i * 13 + k] * dx[k]; a1 += rec[i * 13 + k + 1] * dx[k + 1]; }
        ub[i] = a - (a0 + a1 + rec[i * 13 + 12] * dx[12]);
    }
    wb.sync();
    for (int r = tid; r < FTMPC_NX; r += nt) {
        double a = (r < 3) ? dx[r] + dt * dx[r + 3] : ((r < 6) ? dx[r] : 0.0);
        for (int l = 0; l < 7; ++l) a += jz[l * 13 + r] * dx[6 + l];
        for (int i = 0; i < FTMPC_NU; ++i) a += jz[(7 + i) * 13 + r] * ub[i];
        dn[r] = a;
        if (r < FTMPC_NU) zu[r] = ub[r];
    }
    wb.sync();
}
// t_top: highest stage with a non-zero v_x (N - 1 when v_e != 0 or unknown): the adjoint sweep starts there
template <class WB>
FT_HD void ric_apply(WB& wb, const RicOp& op, const double* v, double* out, int t_top, bool has_e) {
    const int N = op.N, n = op.n, nv = op.nv, tid = wb.tid(), nt = wb.nthreads();
    double* mu0 = op.scr;          // [2][13]
    double* y = op.scr + 26;       // [6]
    double* dxb = op.scr + 32;     // [2][13]
    double* ub = op.scr + 58;      // [6]
    if (has_e) t_top = N - 1;
    for (int i = tid; i < 13; i += nt) mu0[((t_top + 1) & 1) * 13 + i] = (has_e && i < FTMPC_NE) ? v[nv + i] : 0.0;
    for (int i = tid; i < n; i += nt) if (i >= FTMPC_NU * (t_top + 1)) out[i] = 0.0;       // z_t = 0 above the top stage
    wb.sync();
    for (int t = t_top; t >= 0; --t)
        ric_back_step(wb, op.Jz + (size_t)t * 169, op.Rec + (size_t)t * 169, op.dt, v + FTMPC_NU * t, mu0 + ((t + 1) & 1) * 13,
                      mu0 + (t & 1) * 13, y, out + FTMPC_NU * t);
    for (int i = tid; i < 13; i += nt) dxb[i] = 0.0;
    wb.sync();
    for (int t = 0; t < N; ++t)
        ric_fwd_step(wb, op.Jz + (size_t)t * 169, op.Rec + (size_t)t * 169, op.dt, out + FTMPC_NU * t, dxb + (t & 1) * 13,
                     dxb + ((t + 1) & 1) * 13, ub);
    const double* dx = dxb + (N & 1) * 13;
    for (int i = tid; i <= FTMPC_NE; i += nt) {
        if (i < FTMPC_NE) out[nv + i] = dx[i];
        else out[n] = v[n] * op.inv_rho;
    }
    wb.sync();
}

// ---- the operator in G-form ----------------------------------------------------------------------------------------------
// With the closed-loop matrix A~_t = A_t - B_t K_t both sweeps are products with ONE 19 x 19 matrix per stage,
//   G_t = [[A~_t, B_t], [-K_t, I]]:    backward  [mu_t ; y_t]    = G_t' [mu_{t+1} ; v_x,t]
//                                      forward   [dx_{t+1} ; u_t] = G_t  [dx_t ; w_t],   w_t = Lam_t^-1 y_t,
// i.e. one barrier interval per stage and direction, 19 lanes running the same 19-long dot product (no divergent
// branches; the two-interval form above serialises four of them per stage).  G_t and Lam_t^-1 are built once per
// factorisation, all stages in parallel (ric_build_g), FTMPC_RIC_GSTG doubles per stage.
#define FTMPC_RIC_GSTG 398          /* G [19][19] row-major, then Lam^-1 [6][6], one pad: a multiple of 16 bytes (bulk copies) */
template <class Blk>
FT_HD void ric_build_g(Blk& blk, int N, double dt, const double* Jz, const double* Rec, double* G) {
    const int tid = blk.tid(), nt = blk.nthreads();
    for (int idx = tid; idx < N * FTMPC_RIC_GSTG; idx += nt) {
        const int t = idx / FTMPC_RIC_GSTG, e = idx - t * FTMPC_RIC_GSTG;
        const double* jz = Jz + (size_t)t * 169;
        const double* rec = Rec + (size_t)t * 169;
        double v;
        if (e < 361) {
            const int r = e / 19, c = e - r * 19;
            if (r < 13 && c < 13) {                         // A~ = A - B K
                v = (c < 6) ? ((r == c) ? 1.0 : ((c >= 3 && r == c - 3) ? dt : 0.0)) : jz[(c - 6) * 13 + r];
                for (int i = 0; i < FTMPC_NU; ++i) v -= jz[(7 + i) * 13 + r] * rec[i * 13 + c];
            } else if (r < 13) v = jz[(7 + c - 13) * 13 + r];     // B
            else if (c < 13) v = -rec[(r - 13) * 13 + c];         // -K
            else v = (r == c) ? 1.0 : 0.0;
        } else if (e < 397) {
            const int i = (e - 361) / 6, j = (e - 361) - i * 6;   // Lam^-1 = C^-T C^-1
            v = 0.0;
            for (int m = (i > j ? i : j); m < FTMPC_NU; ++m) v += rec[78 + m * 6 + i] * rec[78 + m * 6 + j];
        } else v = 0.0;
        G[idx] = v;
    }
    blk.sync();
}
// out[lane] = sum_k G(lane, k) in[k]  (forward, TR = false)  or  sum_k G(k, lane) in[k]  (backward, TR = true), 19 lanes
// (cp_dst / cp_src: six values copied by the first lanes on the way -- the forward sweep parks u_t in a two-slot buffer and
//  moves it to the result one step later, so that no lane overwrites a w_t entry another lane is still reading)
template <bool TR, class WB>
FT_HD void ric_g_step(WB& wb, const double* G, const double* in13, const double* in6, double* out13, double* out6,
                      double* cp_dst = nullptr, const double* cp_src = nullptr) {
    if (cp_dst) for (int l = wb.tid(); l < FTMPC_NU; l += wb.nthreads()) cp_dst[l] = cp_src[l];
    for (int l = wb.tid(); l < 19; l += wb.nthreads()) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int k = 0; k < 12; k += 3) {
            a0 += (TR ? G[k * 19 + l] : G[l * 19 + k]) * in13[k];
            a1 += (TR ? G[(k + 1) * 19 + l] : G[l * 19 + k + 1]) * in13[k + 1];
            a2 += (TR ? G[(k + 2) * 19 + l] : G[l * 19 + k + 2]) * in13[k + 2];
        }
        a0 += (TR ? G[12 * 19 + l] : G[l * 19 + 12]) * in13[12];
#pragma unroll
        for (int k = 0; k < 6; k += 3) {
            a0 += (TR ? G[(13 + k) * 19 + l] : G[l * 19 + 13 + k]) * in6[k];
            a1 += (TR ? G[(14 + k) * 19 + l] : G[l * 19 + 14 + k]) * in6[k + 1];
            a2 += (TR ? G[(15 + k) * 19 + l] : G[l * 19 + 15 + k]) * in6[k + 2];
        }
        const double v = (a0 + a1) + a2;
        if (l < 13) out13[l] = v; else out6[l - 13] = v;
    }
    wb.sync();
}
// w = Lam^-1 y for the stages t0 .. t0 + cnt - 1 (in place in zu), Gs = the G records of those stages
template <class WB>
FT_HD void ric_g_scale(WB& wb, const double* Gs, double* zu, int cnt) {
    for (int k = wb.tid(); k < cnt; k += wb.nthreads()) {          // a thread per stage: nothing shared inside
        const double* li = Gs + (size_t)k * FTMPC_RIC_GSTG + 361;
        double yk[FTMPC_NU], wk[FTMPC_NU];
#pragma unroll
        for (int m = 0; m < FTMPC_NU; ++m) yk[m] = zu[k * FTMPC_NU + m];
#pragma unroll
        for (int i = 0; i < FTMPC_NU; ++i) {
            double a = 0.0;
#pragma unroll
            for (int m = 0; m < FTMPC_NU; ++m) a += li[i * 6 + m] * yk[m];
            wk[i] = a;
        }
#pragma unroll
        for (int i = 0; i < FTMPC_NU; ++i) zu[k * FTMPC_NU + i] = wk[i];
    }
    wb.sync();
}
template <class WB>
FT_HD void ric_apply_g(WB& wb, const RicOp& op, const double* G, const double* v, double* out, int t_top, bool has_e) {
    const int N = op.N, n = op.n, nv = op.nv, tid = wb.tid(), nt = wb.nthreads();
    double* mu0 = op.scr;          // [2][13]
    double* dxb = op.scr + 32;     // [2][13]
    if (has_e) t_top = N - 1;
    for (int i = tid; i < 13; i += nt) mu0[((t_top + 1) & 1) * 13 + i] = (has_e && i < FTMPC_NE) ? v[nv + i] : 0.0;
    for (int i = tid; i < n; i += nt) if (i >= FTMPC_NU * (t_top + 1)) out[i] = 0.0;
    wb.sync();
    for (int t = t_top; t >= 0; --t)
        ric_g_step<true>(wb, G + (size_t)t * FTMPC_RIC_GSTG, mu0 + ((t + 1) & 1) * 13, v + FTMPC_NU * t, mu0 + (t & 1) * 13,
                         out + FTMPC_NU * t);
    for (int t0 = 0; t0 <= t_top; t0 += 16)
        ric_g_scale(wb, G + (size_t)t0 * FTMPC_RIC_GSTG, out + FTMPC_NU * t0, (t_top + 1 - t0 < 16) ? t_top + 1 - t0 : 16);
    for (int i = tid; i < 13; i += nt) dxb[i] = 0.0;
    wb.sync();
    double* ub = op.scr + 58;      // [2][6]  u_t of the last two stages
    for (int t = 0; t < N; ++t)
        ric_g_step<false>(wb, G + (size_t)t * FTMPC_RIC_GSTG, dxb + (t & 1) * 13, out + FTMPC_NU * t, dxb + ((t + 1) & 1) * 13,
                          ub + (t & 1) * 6, t > 0 ? out + FTMPC_NU * (t - 1) : nullptr, ub + ((t - 1) & 1) * 6);
    const double* dx = dxb + (N & 1) * 13;
    for (int i = tid; i <= FTMPC_NE + FTMPC_NU; i += nt) {
        if (i < FTMPC_NE) out[nv + i] = dx[i];
        else if (i == FTMPC_NE) out[n] = v[n] * op.inv_rho;
        else out[FTMPC_NU * (N - 1) + i - FTMPC_NE - 1] = ub[((N - 1) & 1) * 6 + i - FTMPC_NE - 1];
    }
    wb.sync();
}

#if defined(__CUDACC__)
// ---- G-form with the stage matrices STAGED through shared memory (scratch in global memory: long horizons) -----------------
// Every interval of the sweeps would otherwise start with a round trip to L2 for the stage matrix (measured with the
// two-interval form: 0.35 ms per active-set iteration at N = 100).  Warp 0 sweeps a chunk of RIC_CH stages out of one half of
// a double buffer while the NEXT chunk arrives in the other half as ONE bulk copy (cp.async.bulk global -> shared, 50 KB,
// issued by lane 0, completion on an mbarrier): no load / store instruction of any warp is spent on the staging, and the
// other seven warps simply wait for the product at the block barrier behind it.  v and the result live in shared memory for
// the duration of the product.
#define FTMPC_RIC_CH 16
struct RicStage {
    double* buf;       // [2][RIC_CH][RIC_GSTG], 16-byte aligned
    double* sv;        // [nv + 9]
    double* so;        // [nv + 9]
    unsigned mbar;     // shared address of the mbarrier of the bulk copies (initialised once per kernel, count 1)
    unsigned* par;     // its phase parity, carried from product to product (shared memory, written by lane 0 of warp 0)
};
// shared-memory accessors on 32-bit shared addresses: the staged sweeps run entirely out of shared memory, but their pointers
// reach them through run-time selected fields, so the compiler would emit generic LD / ST (longer latency, long scoreboard)
__device__ __forceinline__ double ric_lds(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void ric_sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
template <bool TR>
__device__ __forceinline__ void ric_g_step_sh(unsigned G, unsigned in13, unsigned in6, unsigned out13, unsigned out6, unsigned cp_dst,
                                              unsigned cp_src, bool cp) {
    const int l = threadIdx.x & 31;
    if (cp && l < FTMPC_NU) ric_sts(cp_dst + 8u * l, ric_lds(cp_src + 8u * l));
    if (l < 19) {
        double g[19], x[19];
#pragma unroll
        for (int k = 0; k < 19; ++k) g[k] = ric_lds(G + 8u * (unsigned)(TR ? k * 19 + l : l * 19 + k));
#pragma unroll
        for (int k = 0; k < 13; ++k) x[k] = ric_lds(in13 + 8u * k);
#pragma unroll
        for (int k = 0; k < 6; ++k) x[13 + k] = ric_lds(in6 + 8u * k);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int k = 0; k < 16; k += 4) { a0 += g[k] * x[k]; a1 += g[k + 1] * x[k + 1]; a2 += g[k + 2] * x[k + 2]; a3 += g[k + 3] * x[k + 3]; }
        a0 += g[16] * x[16]; a1 += g[17] * x[17]; a2 += g[18] * x[18];
        const double v = (a0 + a1) + (a2 + a3);
        ric_sts((l < 13) ? out13 + 8u * l : out6 + 8u * (l - 13), v);
    }
    __syncwarp();
}
// w = Lam^-1 y for cnt <= 32 stages, a lane per stage (Gs, zu in shared memory)
__device__ __forceinline__ void ric_g_scale_sh(unsigned Gs, unsigned zu, int cnt) {
    const int k = threadIdx.x & 31;
    if (k < cnt) {
        const unsigned li = Gs + 8u * (unsigned)(k * FTMPC_RIC_GSTG + 361), z = zu + 8u * (unsigned)(k * FTMPC_NU);
        double yk[FTMPC_NU], wk[FTMPC_NU];
#pragma unroll
        for (int m = 0; m < FTMPC_NU; ++m) yk[m] = ric_lds(z + 8u * m);
#pragma unroll
        for (int i = 0; i < FTMPC_NU; ++i) {
            double a = 0.0;
#pragma unroll
            for (int m = 0; m < FTMPC_NU; ++m) a += ric_lds(li + 8u * (i * 6 + m)) * yk[m];
            wk[i] = a;
        }
#pragma unroll
        for (int i = 0; i < FTMPC_NU; ++i) ric_sts(z + 8u * i, wk[i]);
    }
    __syncwarp();
}
// lane 0 of warp 0: request stages t0 .. t0 + cnt - 1 as one bulk copy; every lane of warp 0 then waits with ric_chunk_wait
__device__ __forceinline__ void ric_chunk_request(const RicStage& sg, const double* G, double* dst, int t0, int cnt) {
    const unsigned bytes = (unsigned)cnt * FTMPC_RIC_GSTG * 8u;
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sg.mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                 "l"(G + (size_t)t0 * FTMPC_RIC_GSTG), "r"(bytes), "r"(sg.mbar)
                 : "memory");
}
__device__ __forceinline__ void ric_chunk_wait(const RicStage& sg, unsigned& par) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(sg.mbar), "r"(par)
                     : "memory");
    }
    par ^= 1u;
}
__device__ __forceinline__ void ric_apply_staged(CudaBlock& blk, const RicOp& op, const double* G, const RicStage& sg, const double* v,
                                                 double* out, int t_top, bool has_e) {
    const int N = op.N, n = op.n, nv = op.nv, tid = threadIdx.x, nt = blockDim.x, ne = nv + FTMPC_NE;
    double* mu0 = op.scr;
    double* dxb = op.scr + 32;
    double* ub = op.scr + 58;
    const unsigned mu_sh = (unsigned)__cvta_generic_to_shared(mu0), dx_sh = (unsigned)__cvta_generic_to_shared(dxb);
    const unsigned ub_sh = (unsigned)__cvta_generic_to_shared(ub), sv_sh = (unsigned)__cvta_generic_to_shared(sg.sv);
    const unsigned so_sh = (unsigned)__cvta_generic_to_shared(sg.so);
    if (has_e) t_top = N - 1;
    const size_t half = (size_t)FTMPC_RIC_CH * FTMPC_RIC_GSTG;
    if (tid == 0) {                                                   // first chunk of the adjoint sweep: in flight during the set-up
        const int lo = (t_top - FTMPC_RIC_CH + 1 > 0) ? t_top - FTMPC_RIC_CH + 1 : 0;
        ric_chunk_request(sg, G, sg.buf, lo, t_top - lo + 1);
    }
    for (int i = tid; i < ne; i += nt) { sg.sv[i] = v[i]; sg.so[i] = 0.0; }
    if (tid < 13) mu0[((t_top + 1) & 1) * 13 + tid] = (has_e && tid < FTMPC_NE) ? v[nv + tid] : 0.0;
    if (tid >= 32 && tid < 45) dxb[tid - 32] = 0.0;
    blk.sync();
    if (tid < 32) {
        unsigned par = *sg.par;
        // ---- backward, chunks from the top stage down (w = Lam^-1 y of a chunk right behind its sweep)
        int hi = t_top, b = 0;
        while (hi >= 0) {
            const int lo = (hi - FTMPC_RIC_CH + 1 > 0) ? hi - FTMPC_RIC_CH + 1 : 0;
            ric_chunk_wait(sg, par);
            if (tid == 0) {                                           // next chunk (or the first one of the rollout) into the other half
                if (lo > 0) {
                    const int nhi = lo - 1, nlo = (nhi - FTMPC_RIC_CH + 1 > 0) ? nhi - FTMPC_RIC_CH + 1 : 0;
                    ric_chunk_request(sg, G, sg.buf + (size_t)(b ^ 1) * half, nlo, nhi - nlo + 1);
                } else {
                    ric_chunk_request(sg, G, sg.buf + (size_t)(b ^ 1) * half, 0, (N < FTMPC_RIC_CH) ? N : FTMPC_RIC_CH);
                }
            }
            const unsigned c_sh = (unsigned)__cvta_generic_to_shared(sg.buf + (size_t)b * half);
            for (int t = hi; t >= lo; --t)
                ric_g_step_sh<true>(c_sh + 8u * (unsigned)((t - lo) * FTMPC_RIC_GSTG), mu_sh + 8u * (((t + 1) & 1) * 13),
                                    sv_sh + 8u * (FTMPC_NU * t), mu_sh + 8u * ((t & 1) * 13), so_sh + 8u * (FTMPC_NU * t), 0u, 0u, false);
            ric_g_scale_sh(c_sh, so_sh + 8u * (FTMPC_NU * lo), hi - lo + 1);
            hi = lo - 1;
            b ^= 1;
        }
        // ---- forward, chunks from stage 0 up (the first one was requested behind the last chunk of the adjoint sweep)
        int lo = 0;
        while (lo < N) {
            const int cnt = (N - lo < FTMPC_RIC_CH) ? N - lo : FTMPC_RIC_CH;
            ric_chunk_wait(sg, par);
            if (tid == 0 && lo + cnt < N) {
                const int nlo = lo + cnt, ncnt = (N - nlo < FTMPC_RIC_CH) ? N - nlo : FTMPC_RIC_CH;
                ric_chunk_request(sg, G, sg.buf + (size_t)(b ^ 1) * half, nlo, ncnt);
            }
            const unsigned c_sh = (unsigned)__cvta_generic_to_shared(sg.buf + (size_t)b * half);
            for (int t = lo; t < lo + cnt; ++t)
                ric_g_step_sh<false>(c_sh + 8u * (unsigned)((t - lo) * FTMPC_RIC_GSTG), dx_sh + 8u * ((t & 1) * 13), so_sh + 8u * (FTMPC_NU * t),
                                     dx_sh + 8u * (((t + 1) & 1) * 13), ub_sh + 8u * ((t & 1) * 6), so_sh + 8u * (FTMPC_NU * (t > 0 ? t - 1 : 0)),
                                     ub_sh + 8u * (((t - 1) & 1) * 6), t > 0);
            lo += cnt;
            b ^= 1;
        }
        if (tid == 0) *sg.par = par;
    }
    blk.sync();
    const double* dx = dxb + (N & 1) * 13;
    for (int i = tid; i < ne; i += nt)
        out[i] = (i < n) ? ((i >= FTMPC_NU * (N - 1)) ? ub[((N - 1) & 1) * 6 + i - FTMPC_NU * (N - 1)] : sg.so[i])
                         : ((i == n) ? v[n] * op.inv_rho : dx[i - nv]);
    blk.sync();
}
#endif

}  // namespace ftmpc
