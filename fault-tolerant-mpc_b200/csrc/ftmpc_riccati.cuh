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
        const double rs = 1.0 / sqrt(d);
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
                         double* lam_stage, double* dscale_out) {
    const int N = L.N, n = L.n, ld = L.nv, nv = L.nv, tid = blk.tid(), nt = blk.nthreads();
    double* P = work;                 // [13][13]  cost-to-go Hessian P_{t+1}
    double* PAB = P + 169;            // [13][19]  P_{t+1} [A_t B_t]
    double* F = PAB + 247;            // [19][19]  lower triangle
    double* Kh = F + 361;             // [6][13]   C^-1 F_ux
    double* Cs = Kh + 78;             // [6][6]    C^-1 (lower)
    double* cst = Cs + 36;            // costates: [2 buffers][2 kinds (g, aug)][13]
    double* Ht = cst + 52;            // [9][9] terminal Hessian model (+ augmentation), [9] augmentation of the terminal gradient
    double* tgv = Ht + 81;
    double* flg = tgv + 9;            // [0] failing pivot + 1, [1] running max of diag(Lam)
    const double* Ah = s.hull;
    const double* lam_prev = lam_prev_g;
    if (sigma > 0.0) {
        for (int i = tid; i < L.mc; i += nt) lam_stage[i] = lam_prev_g[i];
        lam_prev = lam_stage;
    }
    if (tid == 0) { flg[0] = 0.0; flg[1] = 0.0; s.g[n] = 0.0; s.ga[n] = 0.0; }
    blk.sync();
    // ---- terminal model: Ht = term_quad + theta (hessV - term_quad) + sigma sum_A a a',  tgv = sigma sum_A c a
    for (int idx = tid; idx < 90; idx += nt) {
        double v = 0.0;
        const int kk = idx / 9, l = idx - kk * 9;
        if (sigma > 0.0) {
            if (s.tf_val) {                        // <= 2 non-zeros per row of A_f
                for (int i = 0; i < FTMPC_NF; ++i) {
                    if (lam_prev[FTMPC_NH * N + i] > 0.0) {
                        const int k0 = s.tf_idx[2 * i], k1 = s.tf_idx[2 * i + 1];
                        const double v0 = s.tf_val[2 * i], v1 = s.tf_val[2 * i + 1];
                        const double al = (l == k0) ? v0 : ((l == k1) ? v1 : 0.0);
                        const double ak = (kk == k0) ? v0 : ((kk == k1) ? v1 : 0.0);
                        v += (idx < 81) ? ak * al : s.cv[FTMPC_NH * N + i] * al;
                    }
                }
            } else {
                const double* Af = s.cg ? s.cg->Af : cfg.Af;
                for (int i = 0; i < FTMPC_NF; ++i) {
                    if (lam_prev[FTMPC_NH * N + i] > 0.0) {
                        const double al = Af[i * FTMPC_NE + l];
                        v += (idx < 81) ? Af[i * FTMPC_NE + kk] * al : s.cv[FTMPC_NH * N + i] * al;
                    }
                }
            }
            v *= sigma;
        }
        if (idx < 81) {
            const double q0 = cfg.term_quad[idx];
            Ht[idx] = q0 + theta * (hessV[idx] - q0) + v;
        } else {
            tgv[l] = v;
        }
    }
    blk.sync();
    for (int idx = tid; idx < 169 + 26; idx += nt) {
        if (idx < 169) {
            const int r = idx / 13, c = idx - r * 13;
            P[idx] = (r < FTMPC_NE && c < FTMPC_NE) ? Ht[r * FTMPC_NE + c] : 0.0;
        } else {
            const int kd = (idx - 169) / 13, r = (idx - 169) - kd * 13;
            cst[kd * 13 + r] = (r < FTMPC_NE) ? (kd == 0 ? gradV[r] : tgv[r]) : 0.0;
        }
    }
    blk.sync();
    // ---- backward sweep
    for (int t = N - 1; t >= 0; --t) {
        const double* jz = Jz + (size_t)t * 169;
        double* wz = Wz + (size_t)t * 169;
        const double* co = cst + ((N - 1 - t) & 1) * 26;            // costates of stage t + 1
        double* cn = cst + ((N - t) & 1) * 26;                      // costates of stage t
        // (A) PAB = P [A B]  (13 x 19; the (p, v) columns of A are trivial), and the two costate / gradient products
        for (int idx = tid; idx < 247 + 38; idx += nt) {
            if (idx < 247) {
                const int c = idx / 13, r = idx - c * 13;          // consecutive threads: consecutive rows of one column
                double v;
                if (c < 3) v = P[r * 13 + c];
                else if (c < 6) v = cfg.dt * P[r * 13 + c - 3] + P[r * 13 + c];
                else {
                    const double* col = jz + (c - 6) * 13;
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int k = 0; k < 12; k += 2) { a0 += P[r * 13 + k] * col[k]; a1 += P[r * 13 + k + 1] * col[k + 1]; }
                    v = a0 + a1 + P[r * 13 + 12] * col[12];
                }
                PAB[r * 19 + c] = v;
            } else {
                const int kd = (idx - 247) / 19, c = (idx - 247) - kd * 19;
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
                    if (kd == 0 && t > 0) {
                        if (c < FTMPC_NE) v += 2.0 * cfg.Q[c] * (X[t * FTMPC_NX + c] - xref[t * FTMPC_NE + c]);
                        else if (Cq) v += Cq[(size_t)t * FTMPC_CQ + c - FTMPC_NE];
                    }
                    cn[kd * 13 + c] = v;
                } else {
                    const int i = c - 13, a = t * FTMPC_NU + i;
                    if (kd == 0) {
                        s.g[a] = v + 2.0 * cfg.R[i] * (U[a] - (Cq ? Cq[(size_t)t * FTMPC_CQ + 32 + i] : 0.0));
                    } else {
                        double av = 0.0;
                        if (sigma > 0.0)
                            for (int k = 0; k < FTMPC_NH; ++k)
                                if (lam_prev[t * FTMPC_NH + k] > 0.0) av += s.cv[t * FTMPC_NH + k] * Ah[k * FTMPC_NU + i];
                        s.ga[a] = v + sigma * av;                  // ga - g; combined below
                    }
                }
            }
        }
        blk.sync();
        // (B) F = [A B]' PAB + stage Hessian, lower triangle (c1 >= c2)
        for (int idx = tid; idx < 190; idx += nt) {
            int c1 = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
            while ((c1 + 1) * (c1 + 2) / 2 <= idx) ++c1;
            while (c1 * (c1 + 1) / 2 > idx) --c1;
            const int c2 = idx - c1 * (c1 + 1) / 2;
            double v;
            if (c1 < 3) v = PAB[c1 * 19 + c2];
            else if (c1 < 6) v = cfg.dt * PAB[(c1 - 3) * 19 + c2] + PAB[c1 * 19 + c2];
            else {
                const double* col = jz + (c1 - 6) * 13;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int k = 0; k < 12; k += 2) { a0 += col[k] * PAB[k * 19 + c2]; a1 += col[k + 1] * PAB[(k + 1) * 19 + c2]; }
                v = a0 + a1 + col[12] * PAB[12 * 19 + c2];
            }
            // stage Hessian: blkdiag(2 Q[0:6], theta sym(W) + 2Q on omega + 2R on u + Cq Gauss-Newton + sigma hull rows)
            if (c1 < 6) {
                if (c1 == c2) v += 2.0 * cfg.Q[c1];
            } else if (c2 >= 6) {
                const int k1 = c1 - 6, k2 = c2 - 6;
                double hsv = theta * 0.5 * (wz[k1 * 13 + k2] + wz[k2 * 13 + k1]);
                if (k1 == k2) {
                    if (k1 < 3) hsv += 2.0 * cfg.Q[6 + k1];
                    else if (k1 >= 7) hsv += 2.0 * cfg.R[k1 - 7];
                }
                if (Cq) {
                    const double* cq = Cq + (size_t)t * FTMPC_CQ;
                    if (k1 >= 3 && k1 < 7 && k2 >= 3) hsv += cq[16 + (k1 - 3) * 4 + (k2 - 3)];                 // (q, q)
                    else if (k1 >= 7 && k1 < 10 && k2 >= 3 && k2 < 7) hsv += cq[4 + (k1 - 7) * 4 + (k2 - 3)];  // (u~_F, q)
                }
                if (sigma > 0.0 && k2 >= 7) {
                    double av = 0.0;
                    for (int k = 0; k < FTMPC_NH; ++k)
                        if (lam_prev[t * FTMPC_NH + k] > 0.0) av += Ah[k * FTMPC_NU + k1 - 7] * Ah[k * FTMPC_NU + k2 - 7];
                    hsv += sigma * av;
                }
                v += hsv;
            }
            F[c1 * 19 + c2] = v;
        }
        blk.sync();
        // (C) Lam = F_uu = C C', C^-1; Kh = C^-1 F_ux: the 13 threads of the columns of F_ux factor Lam redundantly
        //     (a 6x6 Cholesky is a serial chain either way) and go on to their forward substitution without a barrier
        for (int c = tid; c < FTMPC_NX; c += nt) {
            double A6[6][6], Ci[6][6];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) A6[i][j] = (j <= i) ? F[(13 + i) * 19 + 13 + j] : 0.0;
            double dmx = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) dmx = fmax(dmx, fabs(A6[i][i]));
            const double run = fmax(flg[1], dmx);
            const int bad = ric_chol6(A6, Ci, 1e-10 * fmax(1.0, run));
            double kh[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < 6; ++m) if (m <= i) v += Ci[i][m] * F[(13 + m) * 19 + c];
                kh[i] = v;
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) Kh[i * 13 + c] = kh[i];
            // K[:, c] = C^-T kh  -> stage record (the Wz slot of this stage is dead: F has absorbed it)
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < 6; ++m) if (m >= i) v += Ci[m][i] * kh[m];
                wz[i * 13 + c] = v;
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
        if (flg[0] != 0.0) { *dscale_out = flg[2]; return (int)flg[0]; }
        // (D) P_t = F_xx - Kh' Kh   (symmetric; both halves written)
        for (int idx = tid; idx < 91 + 1; idx += nt) {
            if (idx == 91) { flg[1] = flg[2]; continue; }
            int r = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
            while ((r + 1) * (r + 2) / 2 <= idx) ++r;
            while (r * (r + 1) / 2 > idx) --r;
            const int c = idx - r * (r + 1) / 2;
            double v = F[r * 19 + c];
#pragma unroll
            for (int i = 0; i < 6; ++i) v -= Kh[i * 13 + r] * Kh[i * 13 + c];
            P[r * 13 + c] = v;
            P[c * 13 + r] = v;
        }
        blk.sync();
    }
    *dscale_out = flg[1];
    for (int a = tid; a < n; a += nt) s.ga[a] += s.g[a];
    blk.sync();
    // ---- forward: one closed-loop rollout per column of J
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
    return 0;
}

}  // namespace ftmpc
