// ftmpc_qp2.cuh -- the QP phase of k_solve2: TWO instances resident per SM (<= 113 KB of shared memory and <= 128
// registers per thread for a 256-thread CTA), CUDA only.
//
// What changed against ftmpc_sqp.cuh::phase_qp (one CTA per SM, 224 KB):
//   * the dual active-set iteration runs in range-space form (ftmpc_gis.cuh) on the packed, read-only "extended inverse"
//         K = [ H^-1     H^-1 X' ]        X = d x_N[0:9] / d U      (6x6 blocks, lower block triangle, 72.9 KB at N = 20)
//             [ X H^-1   X H^-1 X' ]
//     instead of rotating the dense E = [J ; X J] (126 KB);
//   * K comes out of ONE register-resident block sweep over the (N+2) x (N+2) block matrix [H X' ; X 0]:
//       thread (i,j) owns block (i,j), i >= j, in 36 registers and per eliminated block column k does exactly one
//       6x6x6 product -- Cholesky trailing update (j > k), inverse recurrence S_ij += L_ik X_kj (j <= k < i), or, once
//       its own X_ij is final (k >= i), the accumulation (H^-1)_ij += X_ki' X_kj.  The two extension block rows take part
//       like ordinary rows of the panel: their L blocks are (X J), their recurrence ends in X H^-1 and the trailing
//       updates of their diagonal corner in -X H^-1 X'.  Only block column k of L and block row k of X = L^-1 are ever
//       shared (2 x 6.3 KB), nothing of size n^2 is read-modify-written in shared memory;
//   * condensing uses ONE panel (column phase and block phase of a stage separated by a barrier -- with a second CTA on
//     the SM the barrier wait is no longer dead time) and carries the sensitivity columns in the panel instead of in
//     registers.
// The mathematics (Hessian schedule, augmented-Lagrangian convexification, pivoting rules of the active-set method,
// tolerances) is that of phase_qp; tests compare the two paths instance by instance.
#pragma once
#include "ftmpc_sqp.cuh"

namespace ftmpc {

#define FTMPC_Q2_BS 37           /* stride of the 6x6 blocks of the shared block column / row: odd, so that the lanes of a
                                     warp (different blocks, same entry) hit different banks */
#define FTMPC_Q2_QCAP 60          /* working-set capacity of the shared-memory R^-1; beyond it the QP is re-run with R^-1 in
                                     the CTA's global slot (7 of 40 000 QPs of the bench workload exceed 64 rows) */

// ---- block bookkeeping ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void q2_block_of(int tid, int& bi, int& bj) {
    bi = (int)((sqrt(8.0 * tid + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= tid) ++bi;
    while (bi * (bi + 1) / 2 > tid) --bi;
    bj = tid - bi * (bi + 1) / 2;
}
struct Qp2Scratch {
    // live for the whole phase
    double *cv, *hull, *g, *ga;
    // condensing
    double *panel, *Jz, *Wz, *qe, *Ht, *tgv, *lam_prev, *hv, *cqs;
    // block sweep
    double *Lcol, *Xrow, *linv, *flag;
    // active-set iteration
    double *K, *xe, *s, *ye, *ze;
    // per member of the working set: aval[8 k + 0..5] normal values, [6] coefficient on the elastic variable, [7] = the six
    // K coordinates of the values as bytes
    struct QVecs { double *Ui, *u, *w, *v, *r, *cs, *tmp, *sub, *aval; int *act, *itmp; int qcap; } qv;
    short* pos;
    const double* tf_val;
    const int* tf_idx;
    const ftmpc_config* cg;
    size_t total;
};
__host__ __device__ inline size_t qp2_fixed_doubles(int N) {
    const WsLayout L = ws_layout(N);
    return (size_t)((L.mc + 1) & ~1) + ((FTMPC_HULL_STRIDE + 1) & ~1) + 2 * (size_t)((L.nv + 1) & ~1);
}
#define FTMPC_Q2_PROWS 26         /* panel rows: 0-12 G_t, 13-19 (W' G_t)[omega, q], 20-25 theta W_ux G_t; terminal stage: 0-8 G_N, 13-21 Ht G_N */
// offset (doubles, from the start of the scratch) of the staged stage Hessians Wz: behind everything the condensing and the
// linearisation (work 326 N, its shared-memory copy of the Jacobians 169 N, costates + states) keep in front of it
__host__ __device__ inline size_t qp2_wz_offset(int N) {
    const WsLayout L = ws_layout(N);
    const int ldp = 7 * N + 1;
    const size_t cond = qp2_fixed_doubles(N) + 2 * (size_t)FTMPC_Q2_PROWS * ldp + (size_t)N * FTMPC_NE + 90 + L.mc + 90 + (size_t)N * 10 + 8;
    const size_t lin = (size_t)N * 326 + (size_t)N * 169 + 2 * (size_t)(N + 1) * FTMPC_NX + 8;
    return ((cond > lin ? cond : lin) + 1) & ~(size_t)1;
}
__host__ __device__ inline size_t qp2_scratch_doubles(int N) {
    const WsLayout L = ws_layout(N);
    const int NB = N + 2, ne = L.nv + FTMPC_NE, qc = FTMPC_Q2_QCAP;
    const size_t cond = qp2_wz_offset(N) + (size_t)N * 169;
    const size_t kblk = (size_t)(ne - 1) * ne / 2 + 1;                 // row-packed K over the n + 9 K coordinates (ne - 1 = n + 9)
    const size_t chol = qp2_fixed_doubles(N) + kblk + 3 * (size_t)NB * FTMPC_Q2_BS + 8;
    const size_t ints = ((size_t)2 * (qc + 2) + (L.m + 1) / 2 + 1) / 2 + 1;
    const size_t gi = qp2_fixed_doubles(N) + kblk + (size_t)qc * (qc + 1) / 2 + 3 * (size_t)(ne + 1) + (L.m + 2) + 6 * (size_t)(qc + 2) +
                      2 * (size_t)(qc + 2) + 8 * (size_t)qc + ints + 8;
    size_t r = cond > chol ? cond : chol;
    if (gi > r) r = gi;
    return r;
}
__device__ __forceinline__ Qp2Scratch qp2_carve(double* buf, int N, const StepIO& io) {
    const WsLayout L = ws_layout(N);
    const int NB = N + 2, ldp = 7 * N + 1, ne = L.nv + FTMPC_NE, qc = FTMPC_Q2_QCAP;
    Qp2Scratch s;
    s.cg = io.cfg_g; s.tf_val = io.tf_val; s.tf_idx = io.tf_idx;
    double* p = buf;
    s.cv = p; p += (L.mc + 1) & ~1;
    s.hull = p; p += (FTMPC_HULL_STRIDE + 1) & ~1;
    s.g = p; p += (L.nv + 1) & ~1;
    s.ga = p; p += (L.nv + 1) & ~1;
    double* R = p;
    // condensing
    s.panel = p; p += 2 * (size_t)FTMPC_Q2_PROWS * ldp;       // double-buffered
    s.qe = p; p += (size_t)N * FTMPC_NE;
    s.Ht = p; p += 81;
    s.tgv = p; p += 9;
    s.lam_prev = p; p += L.mc;
    s.hv = p; p += 90;
    s.cqs = p; p += (size_t)N * 10;
    s.Wz = buf + qp2_wz_offset(N);            // the linearisation leaves the stage Hessians here (the Jacobians are read
    s.Jz = nullptr;                           // from the CTA's global slot: only the column role touches them)
    // block sweep: K first (written when the sweep has succeeded), the shared block column / row behind it
    const size_t kblk = (size_t)(ne - 1) * ne / 2 + 1;
    p = R;
    s.K = p; p += kblk;
    s.Lcol = p; p += (size_t)NB * FTMPC_Q2_BS;
    s.Xrow = p; p += (size_t)NB * FTMPC_Q2_BS;
    s.linv = p; p += (size_t)NB * FTMPC_Q2_BS;
    s.flag = p; p += 8;
    // active-set iteration
    p = R + kblk;
    s.qv.qcap = qc;
    s.qv.Ui = p; p += (size_t)qc * (qc + 1) / 2;
    s.xe = p; p += ne + 1;
    s.ye = p; p += ne + 1;
    s.ze = p; p += ne + 1;
    s.s = p; p += L.m + 2;
    s.qv.u = p; p += qc + 2;
    s.qv.w = p; p += qc + 2;
    s.qv.v = p; p += qc + 2;
    s.qv.r = p; p += qc + 2;
    s.qv.tmp = p; p += qc + 2;
    s.qv.sub = p; p += qc + 2;
    s.qv.cs = p; p += 2 * (qc + 2);
    s.qv.aval = p; p += 8 * (size_t)qc;
    int* ip = reinterpret_cast<int*>(p);
    s.qv.act = ip; ip += qc + 2;
    s.qv.itmp = ip; ip += qc + 2;
    s.pos = reinterpret_cast<short*>(ip);
    s.total = qp2_scratch_doubles(N);
    return s;
}

// =====================================================================================================================
// condensing: H blocks -> registers of the block threads, g / ga -> shared memory, X = d x_N[0:9]/dU -> the extension
// block rows.  Same arithmetic as the register-tiled path of ftmpc_sqp.cuh::condense (see there for the derivation).
// =====================================================================================================================
__device__ __forceinline__ void condense2(CudaBlock& blk, const ftmpc_config& cfg, const WsLayout& L, const Qp2Scratch& s,
                                          const double* Jz /* global: [N][13 cols][13] */, const double* X, const double* U,
                                          const double* xref, double theta, double sigma, const double* Cq, double (&acc)[6][6],
                                          int bi, int bj) {
    const int N = L.N, n = L.n, tid = blk.tid(), nt = blk.nthreads();
    const int ldp = 7 * N + 1;
    double* Wp = s.Wz;
    double *qe = s.qe, *Ht = s.Ht, *tgv = s.tgv, *lam_prev = s.lam_prev, *hv = s.hv, *cqs = s.cqs;
    const double* Ah = s.hull;
    // ---- pre-pass: W' = theta * sym(W) (+ 2Q on the omega diagonal), qe, Ht           (hv, lam_prev staged by the caller)
    for (int idx = tid; idx < N * 169; idx += nt) {
        const int t = idx / 169, e = idx - t * 169;
        const int c = e / 13, r = e - c * 13;
        if (c < r) continue;
        double* wz = Wp + (size_t)t * 169;
        double v = theta * 0.5 * (wz[c * 13 + r] + wz[r * 13 + c]);
        if (c == r && c < 3) v += 2.0 * cfg.Q[6 + c];
        wz[c * 13 + r] = v;
        wz[r * 13 + c] = v;
    }
    for (int idx = tid; idx < N * FTMPC_NE; idx += nt) {
        const int t = idx / FTMPC_NE, kk = idx - t * FTMPC_NE;
        qe[idx] = 2.0 * cfg.Q[kk] * (X[t * FTMPC_NX + kk] - xref[t * FTMPC_NE + kk]);
    }
    if (Cq)
        for (int i = tid; i < N * 10; i += nt) {
            const int t = i / 10, e = i - t * 10;
            cqs[i] = Cq[(size_t)t * FTMPC_CQ + (e < 4 ? e : 28 + e)];
        }
    blk.sync();
    if (Cq) {       // Gauss-Newton part of the input cost's attitude coupling: NOT blended (added after the theta scaling)
        for (int idx = tid; idx < N * 28; idx += nt) {
            const int t = idx / 28, e = idx - t * 28;
            double* wz = Wp + (size_t)t * 169;
            if (e < 16) {
                wz[(3 + (e >> 2)) * 13 + 3 + (e & 3)] += Cq[(size_t)t * FTMPC_CQ + 16 + e];
            } else {
                const int i = (e - 16) >> 2, b = (e - 16) & 3;
                const double v = Cq[(size_t)t * FTMPC_CQ + 4 + i * 4 + b];
                wz[(7 + i) * 13 + 3 + b] += v;
                wz[(3 + b) * 13 + 7 + i] += v;
            }
        }
    }
    for (int idx = tid; idx < 90; idx += nt) {
        double v = 0.0;
        const int kk = idx / 9, l = idx - kk * 9;
        if (sigma > 0.0) {
            for (int i = 0; i < FTMPC_NF; ++i) {
                if (lam_prev[FTMPC_NH * N + i] > 0.0) {
                    const int k0 = s.tf_idx[2 * i], k1 = s.tf_idx[2 * i + 1];
                    const double v0 = s.tf_val[2 * i], v1 = s.tf_val[2 * i + 1];
                    const double al = (l == k0) ? v0 : ((l == k1) ? v1 : 0.0);
                    const double ak = (kk == k0) ? v0 : ((kk == k1) ? v1 : 0.0);
                    v += (idx < 81) ? ak * al : s.cv[FTMPC_NH * N + i] * al;
                }
            }
            v *= sigma;
        }
        if (idx < 81) {
            const double q0 = s.cg->term_quad[idx];
            Ht[idx] = q0 + theta * (hv[idx] - q0) + v;
        } else {
            tgv[l] = v;
        }
    }
    if (tid == 0) { s.g[n] = 0.0; s.ga[n] = 0.0; }
    // ---- roles: block (bi, bj) from the caller; column role on the LAST n threads (their blocks have the least work)
    const int a = (nt - 1 - tid < n) ? nt - 1 - tid : -1;
    const int ta = (a >= 0) ? a / FTMPC_NU : 0, ja = a - ta * FTMPC_NU;
    const int pa_ = 7 * ta + ja;
    double gs = 0.0, gaug = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = 0.0;
    blk.sync();
    blk.mark(PH_COND_PRE);
    // software pipeline over two panels: interval tt runs the column phase of stage tt + 1 (reads its own column of G_tt from
    // panel tt & 1, writes panel (tt + 1) & 1) and the block phase of stage tt (reads panel tt & 1); one barrier per stage
    for (int tt = -1; tt <= N; ++tt) {
        // ---------------- column phase of stage t = tt + 1
        const int t = tt + 1;
        if (a >= 0 && t <= N && ta < t) {
            double* buf = s.panel + (size_t)(t & 1) * FTMPC_Q2_PROWS * ldp;
            const double* jz = Jz + (size_t)(t - 1) * 169;
            double g[FTMPC_NX];
            if (ta == t - 1) {                     // birth: G_t[:, a] = B_{t-1} e_ja
#pragma unroll
                for (int r = 0; r < FTMPC_NX; ++r) g[r] = jz[(7 + ja) * 13 + r];
                gs = 2.0 * cfg.R[ja] * (U[a] - (Cq ? cqs[(t - 1) * 10 + 4 + ja] : 0.0));
                if (sigma > 0.0) {
                    double av = 0.0;
                    for (int i = 0; i < FTMPC_NH; ++i)
                        if (lam_prev[(t - 1) * FTMPC_NH + i] > 0.0) av += s.cv[(t - 1) * FTMPC_NH + i] * Ah[i * FTMPC_NU + ja];
                    gaug = sigma * av;
                }
            } else {                               // G_t[:, a] = A_{t-1} G_{t-1}[:, a]
                const double* prev = s.panel + (size_t)((t - 1) & 1) * FTMPC_Q2_PROWS * ldp;
                double gp[FTMPC_NX];
#pragma unroll
                for (int r = 0; r < FTMPC_NX; ++r) gp[r] = prev[r * ldp + pa_];
#pragma unroll
                for (int r = 0; r < FTMPC_NX; ++r) {
                    double v = (r < 3) ? gp[r] + cfg.dt * gp[r + 3] : ((r < 6) ? gp[r] : 0.0);
#pragma unroll
                    for (int l = 0; l < 7; ++l) v += jz[l * 13 + r] * gp[6 + l];
                    g[r] = v;
                }
            }
            if (t < N) {
                const double* wp = Wp + (size_t)t * 169;
#pragma unroll
                for (int r = 0; r < FTMPC_NX; ++r) buf[r * ldp + pa_] = g[r];
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    double v = 0.0;
#pragma unroll
                    for (int l = 0; l < 7; ++l) v += wp[k * 13 + l] * g[6 + l];
                    buf[(13 + k) * ldp + pa_] = v;
                }
#pragma unroll
                for (int i = 0; i < FTMPC_NU; ++i) {
                    double v = 0.0;
#pragma unroll
                    for (int l = 0; l < 7; ++l) v += wp[(7 + i) * 13 + l] * g[6 + l];
                    buf[(20 + i) * ldp + pa_] = v;
                }
#pragma unroll
                for (int kk = 0; kk < FTMPC_NE; ++kk) gs += qe[t * FTMPC_NE + kk] * g[kk];
                if (Cq) {
#pragma unroll
                    for (int l = 0; l < 4; ++l) gs += cqs[t * 10 + l] * g[9 + l];
                }
            } else {
                // terminal stage: publish G_N and Ht G_N, finish the gradient
                double va = 0.0, vg = 0.0;
#pragma unroll
                for (int kk = 0; kk < FTMPC_NE; ++kk) {
                    double v = 0.0;
#pragma unroll
                    for (int l = 0; l < FTMPC_NE; ++l) v += Ht[kk * FTMPC_NE + l] * g[l];
                    buf[kk * ldp + pa_] = g[kk];
                    buf[(13 + kk) * ldp + pa_] = v;
                    vg += hv[81 + kk] * g[kk];
                    va += tgv[kk] * g[kk];
                }
                gs += vg;
                s.g[a] = gs;
                s.ga[a] = gaug + gs + va;
            }
        }
        // ---------------- block phase of stage tt
        if (tt >= 0 && bi >= 0) {
            const double* buf = s.panel + (size_t)(tt & 1) * FTMPC_Q2_PROWS * ldp;
            if (bi < N) {
                if (bi < tt) {
                    // rank-13 update (9 rows at the terminal stage): G_t[:, bi]' (M_t G_t)[:, bj].  Rows 0-5 of M_t G_t are the
                    // rows of G_t scaled by 2 Q_r (scaled on the fly: six panel rows less).  The row loop stays ROLLED:
                    // unrolled it is 10 KB of straight-line code, and two CTAs share the instruction cache
                    const double* Pa = buf + 7 * bi;
                    const double* Pb = buf + 7 * bj;
                    const bool term = (tt == N);
                    const int nr = term ? FTMPC_NE : FTMPC_NX;
#pragma unroll 1
                    for (int r = 0; r < nr; ++r) {
                        const int rb = term ? 13 + r : ((r < 6) ? r : r + 7);
                        const double sc = (!term && r < 6) ? 2.0 * cfg.Q[r] : 1.0;
                        double pa[6], tb[6];
#pragma unroll
                        for (int i = 0; i < 6; ++i) { pa[i] = Pa[(size_t)r * ldp + i]; tb[i] = sc * Pb[(size_t)rb * ldp + i]; }
#pragma unroll
                        for (int i = 0; i < 6; ++i)
#pragma unroll
                            for (int j = 0; j < 6; ++j) acc[i][j] += pa[i] * tb[j];
                    }
                } else if (bi == tt) {
                    if (bj < tt) {
#pragma unroll
                        for (int i = 0; i < 6; ++i)
#pragma unroll
                            for (int j = 0; j < 6; ++j) acc[i][j] = buf[(size_t)(20 + i) * ldp + 7 * bj + j];
                    } else {
                        const double* wp = Wp + (size_t)tt * 169;
#pragma unroll
                        for (int i = 0; i < 6; ++i)
#pragma unroll
                            for (int j = 0; j < 6; ++j) {
                                double v = wp[(7 + i) * 13 + 7 + j];
                                if (i == j) v += 2.0 * cfg.R[i];
                                acc[i][j] = v;
                            }
                        if (sigma > 0.0) {
                            for (int r = 0; r < FTMPC_NH; ++r) {
                                if (lam_prev[tt * FTMPC_NH + r] > 0.0) {
#pragma unroll
                                    for (int i = 0; i < 6; ++i)
#pragma unroll
                                        for (int j = 0; j < 6; ++j) acc[i][j] += sigma * Ah[r * FTMPC_NU + i] * Ah[r * FTMPC_NU + j];
                                }
                            }
                        }
                    }
                }
            } else if (tt == N && bj < N) {
                // extension block rows: X = G_N[0:9, :] (rows 6..8 of the second one are padding)
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const int row = 6 * (bi - N) + i;
#pragma unroll
                    for (int j = 0; j < 6; ++j) acc[i][j] = (row < FTMPC_NE) ? buf[(size_t)row * ldp + 7 * bj + j] : 0.0;
                }
            }
        }
        blk.sync();
        blk.mark(PH_COND_BLK);
    }
}

// =====================================================================================================================
// the unified block sweep (see the header of this file).  On success K is in shared memory, returns 0; otherwise the
// failing pivot index + 1 (uniform over the block), K untouched.
// =====================================================================================================================
// factor the 6x6 diagonal block in place (lower triangle of acc <- L_kk) and publish Y = L_kk^-1 (lower) to linv
__device__ __forceinline__ void q2_diag_block(double (&acc)[6][6], double* lk, double* flag, int k, double piv_tol) {
    int bad = 0;
    double inv[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = acc[j][j];
#pragma unroll
        for (int m = 0; m < 6; ++m) if (m < j) d -= acc[j][m] * acc[j][m];
        if (!(d > piv_tol)) bad = 1;
        const double rs = rsqrt(d);
        inv[j] = rs;
        acc[j][j] = d * rs;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i > j) {
                double v = acc[i][j];
#pragma unroll
                for (int m = 0; m < 6; ++m) if (m < j) v -= acc[i][m] * acc[j][m];
                acc[i][j] = v * rs;
            }
        }
    }
    // Y = L_kk^-1, column by column, written straight to shared memory (upper part zero)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double y[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i < j) y[i] = 0.0;
            else if (i == j) y[i] = inv[j];
            else {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < 6; ++m) if (m >= j && m < i) v += acc[i][m] * y[m];
                y[i] = -v * inv[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) lk[i * 6 + j] = y[i];
    }
    if (bad) *flag = (double)(6 * k + 1);
}

__device__ __forceinline__ int chol_k_blocks(CudaBlock& blk, int N, const Qp2Scratch& s, double (&acc)[6][6], int bi, int bj,
                                             double piv_tol) {
    const int NB = N + 2;
    double* flag = s.flag;
    if (bi == 0 && bj == 0) {
        *flag = 0.0;
        q2_diag_block(acc, s.linv, flag, 0, piv_tol);
    }
    blk.sync();
    for (int k = 0; k < N; ++k) {
        if (*flag != 0.0) return (int)*flag;
        const double* lk = s.linv + (size_t)k * FTMPC_Q2_BS;          // Y = L_kk^-1 (lower triangular, full 6x6 storage)
        // ---------------- panel
        if (bi >= 0) {
            if (bj == k && bi > k) {                          // L_ik = A_ik Y'  (row by row, in place), publish
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    double x[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        double v = 0.0;
#pragma unroll
                        for (int m = 0; m < 6; ++m) if (m <= j) v += acc[i][m] * lk[j * 6 + m];
                        x[j] = v;
                    }
                    double* o = s.Lcol + (size_t)bi * FTMPC_Q2_BS + i * 6;
#pragma unroll
                    for (int j = 0; j < 6; ++j) { acc[i][j] = x[j]; o[j] = x[j]; }
                }
            } else if (bi == k && bj < k) {                   // X_kj = -Y S_kj  (column by column, in place), publish
#pragma unroll
                for (int b = 0; b < 6; ++b) {
                    double x[6];
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        double v = 0.0;
#pragma unroll
                        for (int m = 0; m < 6; ++m) if (m <= a) v += lk[a * 6 + m] * acc[m][b];
                        x[a] = -v;
                    }
#pragma unroll
                    for (int a = 0; a < 6; ++a) { acc[a][b] = x[a]; s.Xrow[(size_t)bj * FTMPC_Q2_BS + a * 6 + b] = x[a]; }
                }
            } else if (bi == k && bj == k) {                  // X_kk = Y
#pragma unroll
                for (int i = 0; i < 36; ++i) s.Xrow[(size_t)k * FTMPC_Q2_BS + i] = lk[i];
            }
        }
        blk.sync();
        // ---------------- update: exactly one 6x6x6 product per thread
        if (bi > k) {
            if (bj == k) {                                    // S_ik = L_ik Y  (first term of the inverse recurrence)
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    double x[6];
#pragma unroll
                    for (int b = 0; b < 6; ++b) {
                        double v = 0.0;
#pragma unroll
                        for (int m = 0; m < 6; ++m) if (m >= b) v += acc[a][m] * lk[m * 6 + b];
                        x[b] = v;
                    }
#pragma unroll
                    for (int b = 0; b < 6; ++b) acc[a][b] = x[b];
                }
            } else if (bj > k) {                              // A_ij -= L_ik L_jk'
                const double* La = s.Lcol + (size_t)bi * FTMPC_Q2_BS;
                const double* Lb = s.Lcol + (size_t)bj * FTMPC_Q2_BS;
#pragma unroll 1
                for (int m = 0; m < 6; ++m) {
                    double a6[6], b6[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) { a6[i] = La[i * 6 + m]; b6[i] = Lb[i * 6 + m]; }
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int j = 0; j < 6; ++j) acc[i][j] -= a6[i] * b6[j];
                }
                if (bi == k + 1 && bj == k + 1 && k + 1 < N) q2_diag_block(acc, s.linv + (size_t)(k + 1) * FTMPC_Q2_BS, flag, k + 1, piv_tol);
            } else {                                          // S_ij += L_ik X_kj
                const double* La = s.Lcol + (size_t)bi * FTMPC_Q2_BS;
                const double* Xb = s.Xrow + (size_t)bj * FTMPC_Q2_BS;
#pragma unroll 1
                for (int m = 0; m < 6; ++m) {
                    double a6[6], b6[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) { a6[i] = La[i * 6 + m]; b6[i] = Xb[m * 6 + i]; }
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int j = 0; j < 6; ++j) acc[i][j] += a6[i] * b6[j];
                }
            }
        } else if (bi == k) {                                 // own X_kj is final: (H^-1)_kj starts as Y' X_kj, in place
            if (bj == k) {
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int b = 0; b < 6; ++b) {
                        double v = 0.0;
#pragma unroll
                        for (int m = 0; m < 6; ++m) if (m >= a && m >= b) v += lk[m * 6 + a] * lk[m * 6 + b];
                        acc[a][b] = v;
                    }
            } else {
#pragma unroll
                for (int b = 0; b < 6; ++b) {
                    double x[6];
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        double v = 0.0;
#pragma unroll
                        for (int m = 0; m < 6; ++m) if (m >= a) v += lk[m * 6 + a] * acc[m][b];
                        x[a] = v;
                    }
#pragma unroll
                    for (int a = 0; a < 6; ++a) acc[a][b] = x[a];
                }
            }
        } else if (bi >= 0) {                                 // bi < k: (H^-1)_ij += X_ki' X_kj
            const double* Xa = s.Xrow + (size_t)bi * FTMPC_Q2_BS;
            const double* Xb = s.Xrow + (size_t)bj * FTMPC_Q2_BS;
#pragma unroll 1
            for (int m = 0; m < 6; ++m) {
                double a6[6], b6[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) { a6[i] = Xa[m * 6 + i]; b6[i] = Xb[m * 6 + i]; }
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = 0; j < 6; ++j) acc[i][j] += a6[i] * b6[j];
            }
        }
        blk.sync();
    }
    if (*flag != 0.0) return (int)*flag;
    blk.mark(PH_CHOL);
    // K blocks: (H^-1)_ij for i < N; X H^-1 = S for the extension rows; X H^-1 X' = -(trailing corner)
    if (bi >= 0) {
        const double sg = (bi >= N && bj >= N) ? -1.0 : 1.0;
        const int nk = 6 * N + FTMPC_NE;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int r = 6 * bi + i;
            if (r < nk) {                                     // (the last three rows of the second extension block are padding)
                double* o = s.K + ((r * (r + 1)) >> 1) + 6 * bj;
#pragma unroll
                for (int j = 0; j < 6; ++j) if (bi > bj || j <= i) o[j] = sg * acc[i][j];
            }
        }
    }
    (void)NB;
    blk.sync();
    blk.mark(PH_INV);
    return 0;
}

// =====================================================================================================================
// range-space dual active-set iteration on the packed K (CUDA specialisation of ftmpc_gis.cuh::gis_solve for the MPC
// constraint structure).  K is row-packed (lower triangle, K(i,j) at i (i + 1) / 2 + j) over the K coordinates
// [d (n) ; X d (9)]; the vectors (xe, ye, ze) use the same coordinates with the elastic variable LAST (index n + 9) -- it
// is decoupled in the Hessian, K_delta = 1 / rho_slack.  Per added constraint, seven barrier intervals:
//   A  ye = K n_p                                  B  w_k = n_k . ye (cached normals of the working set)
//   C  v = R^-T w      D  r = R^-1 v               E  ze = ye - sum_k r_k (K n_k)      (lane pair per row, k split by parity)
//   F  e_k = n_k . ze (refinement test), dual step length              G  primal/dual step, slacks, next candidate, update
// Every helper has ONE call site (the refinement re-enters C-E through a loop): with two CTAs per SM the instruction
// cache is shared, and the first version of this routine (15 k instructions, two instantiations) stalled on fetches.
// =====================================================================================================================
__device__ __forceinline__ int q2_tri(int i) { return (i * (i + 1)) >> 1; }
__device__ __forceinline__ double q2_Kel(const double* K, int i, int c) {
    const int hi = i > c ? i : c, lo = i > c ? c : i;
    return K[q2_tri(hi) + lo];
}
// n_p . v for an arbitrary constraint row (v in K coordinates, elastic variable at index sl)
__device__ __forceinline__ double q2_rowdot(const MpcCons& cons, int p, const double* v, int sl) {
    const int N = cons.N;
    if (p < FTMPC_NH * N) {
        const int t = p / FTMPC_NH, i = p - t * FTMPC_NH;
        const double* a = cons.Ah + i * FTMPC_NU;
        const double* x = v + t * FTMPC_NU;
        double sv = 0.0;
#pragma unroll
        for (int j = 0; j < FTMPC_NU; ++j) sv -= a[j] * x[j];
        const double c = cons.cv[p];
        if (c > 0.0) sv += c * v[sl];
        return sv;
    }
    if (p < cons.mc) {
        const int i = p - FTMPC_NH * N;
        double sv = -cons.tf_val[2 * i] * v[cons.n + cons.tf_idx[2 * i]] - cons.tf_val[2 * i + 1] * v[cons.n + cons.tf_idx[2 * i + 1]];
        const double c = cons.cv[p];
        if (c > 0.0) sv += c * v[sl];
        return sv;
    }
    return (p == cons.mc) ? v[sl] : -v[sl];
}

// returns GI_OK / GI_MAXIT / GI_INFEASIBLE, or 5 when the working set outgrew qcap (the caller hands the instance to the
// null-space kernel).  On entry s.ze holds the unconstrained minimiser (K coordinates), on exit s.xe the solution.
__device__ __forceinline__ int gis2_solve(CudaBlock& blk, const MpcCons& cons, const Qp2Scratch& s, double kslack, double* lam,
                                          int maxit, double tol, int* iters_out, int* nact_out) {
    const int tid = blk.tid(), nt = blk.nthreads(), lane = tid & 31, warp = tid >> 5;
    const int N = cons.N, n = cons.n, NK = n + FTMPC_NE, SL = NK, m = cons.mc + 2, qcap = s.qv.qcap;
    const double dep_tol = 1e-14, refine_tol = 1e-13;
    const Qp2Scratch::QVecs& qv = s.qv;
    double* const Ui = qv.Ui;
    double* gsc = blk.scratch + 128;
    int q = 0, iters = 0, status = GI_OK;
    // ---- primal step + slack update + most violated row not in the working set (two call sites: start-up, step G)
    double nb;
    int nbi;
    auto apply_step = [&](double t, int skip) {
        for (int i = tid; i <= NK; i += nt) s.xe[i] += t * s.ze[i];
        nb = 0.0;
        nbi = 0x7fffffff;
        for (int i = tid; i < m; i += nt) {
            const double v = s.s[i] + t * q2_rowdot(cons, i, s.ze, SL);
            s.s[i] = v;
            if (i != skip && s.pos[i] < 0 && (v < nb || (v == nb && i < nbi))) { nb = v; nbi = i; }
        }
    };
    // start: x = 0, slacks -beta, then the step x += 1 * x_unc
    for (int i = tid; i < m; i += nt) {
        s.s[i] = (i < cons.mc) ? -cons.cv[i] : ((i == cons.mc) ? 0.0 : 1.0);
        s.pos[i] = (short)-1;
    }
    for (int i = tid; i <= NK; i += nt) s.xe[i] = 0.0;
    blk.sync();
    apply_step(1.0, -1);
    blk.argmin(nb, nbi);
    for (;;) {
        if (nbi == 0x7fffffff || nb >= -tol) break;      // primal feasible -> optimal
        blk.mark(PH_GI_SELECT);
        const int p = nbi;
        double sp = nb;
        // the normal of p as (index, value) pairs in K coordinates + the coefficient on the elastic variable
        int pidx[6];
        double pval[6], pel;
        {
            if (p < FTMPC_NH * N) {
                const int t = p / FTMPC_NH, i = p - t * FTMPC_NH;
#pragma unroll
                for (int j = 0; j < 6; ++j) { pidx[j] = t * FTMPC_NU + j; pval[j] = -cons.Ah[i * FTMPC_NU + j]; }
                const double c = cons.cv[p];
                pel = (c > 0.0) ? c : 0.0;
            } else if (p < cons.mc) {
                const int i = p - FTMPC_NH * N;
#pragma unroll
                for (int j = 0; j < 6; ++j) { pidx[j] = n; pval[j] = 0.0; }
                pidx[0] = n + cons.tf_idx[2 * i]; pval[0] = -cons.tf_val[2 * i];
                pidx[1] = n + cons.tf_idx[2 * i + 1]; pval[1] = -cons.tf_val[2 * i + 1];
                const double c = cons.cv[p];
                pel = (c > 0.0) ? c : 0.0;
            } else {
#pragma unroll
                for (int j = 0; j < 6; ++j) { pidx[j] = 0; pval[j] = 0.0; }
                pel = (p == cons.mc) ? 1.0 : -1.0;
            }
        }
        if (tid == 0) qv.u[q] = 0.0;
        // ---- A: ye = K n_p
        for (int i = tid; i <= NK; i += nt) {
            double a = pel * kslack;
            if (i < NK) {
                a = 0.0;
#pragma unroll
                for (int j = 0; j < 6; ++j) a += pval[j] * q2_Kel(s.K, i, pidx[j]);
            }
            s.ye[i] = a;
        }
        blk.sync();
        double dn = pel * s.ye[SL];
#pragma unroll
        for (int j = 0; j < 6; ++j) dn += pval[j] * s.ye[pidx[j]];
        blk.mark(PH_GI_D);
        bool added = false;
        while (!added) {
            ++iters;
            if (iters > maxit) { status = GI_MAXIT; break; }
            double d2n = dn, t1 = INFINITY;
            int l = 0x7fffffff;
            if (q == 0) {
                for (int i = tid; i <= NK; i += nt) s.ze[i] = s.ye[i];
                blk.sync();
            } else {
                // ---- B: w = N_W' ye
                for (int k = tid; k < q; k += nt) {
                    const double* av = qv.aval + 8 * k;
                    const unsigned char* ai = reinterpret_cast<const unsigned char*>(av + 7);
                    double a = av[6] * s.ye[SL];
#pragma unroll
                    for (int j = 0; j < 6; ++j) a += av[j] * s.ye[ai[j]];
                    qv.w[k] = a;
                }
                blk.sync();
                for (int pass = 0; pass < 2; ++pass) {
                    // ---- C: v = R^-T w,  D: r = R^-1 v   (four lanes per entry)
                    {
                        const int part = tid & 3, per = nt >> 2;
                        for (int k0 = 0; k0 < q; k0 += per) {
                            const int k = k0 + (tid >> 2);
                            double a = 0.0;
                            if (k < q) {
                                const double* col = Ui + gi_tri(k);
                                for (int j = part; j <= k; j += 4) a += col[j] * qv.w[j];
                            }
                            a += __shfl_xor_sync(0xffffffffu, a, 1);
                            a += __shfl_xor_sync(0xffffffffu, a, 2);
                            if (k < q && part == 0) qv.v[k] = a;
                        }
                        blk.sync();
                        for (int k0 = 0; k0 < q; k0 += per) {
                            const int k = k0 + (tid >> 2);
                            double a = 0.0;
                            if (k < q)
                                for (int kk = k + part; kk < q; kk += 4) a += Ui[gi_tri(kk) + k] * qv.v[kk];
                            a += __shfl_xor_sync(0xffffffffu, a, 1);
                            a += __shfl_xor_sync(0xffffffffu, a, 2);
                            if (k < q && part == 0) qv.tmp[k] = a;         // r of this pass (delta r in the refinement pass)
                        }
                        blk.sync();
                    }
                    // ---- E: ze = (ye | ze) - sum_k r_k (K n_k): lane pair per row (k split by parity), last row on warp 0
                    {
                        const double* src = pass ? s.ze : s.ye;
                        const int pair = tid >> 1, part = tid & 1;
                        const bool on = pair < NK && (pair < NK - 1 || 2 * NK <= nt);
                        const int row = on ? pair : 0;
                        double sv = 0.0;
                        for (int k = part; on && k < q; k += 2) {
                            const double* av = qv.aval + 8 * k;
                            const unsigned char* ai = reinterpret_cast<const unsigned char*>(av + 7);
                            double a = 0.0;
#pragma unroll
                            for (int j = 0; j < 6; ++j) a += av[j] * q2_Kel(s.K, row, ai[j]);
                            sv += qv.tmp[k] * a;
                        }
                        sv += __shfl_xor_sync(0xffffffffu, sv, 1);
                        if (on && part == 0) s.ze[row] = src[row] - sv;
                        if (2 * NK > nt && warp == 0) {
                            double tv = 0.0;
                            for (int k = lane; k < q; k += 32) {
                                const double* av = qv.aval + 8 * k;
                                const unsigned char* ai = reinterpret_cast<const unsigned char*>(av + 7);
                                double a = 0.0;
#pragma unroll
                                for (int j = 0; j < 6; ++j) a += av[j] * q2_Kel(s.K, NK - 1, ai[j]);
                                tv += qv.tmp[k] * a;
                            }
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) tv += __shfl_xor_sync(0xffffffffu, tv, o);
                            if (lane == 0) s.ze[NK - 1] = src[NK - 1] - tv;
                        }
                        if (warp == (nt >> 5) - 1) {               // elastic coordinate
                            double tv = 0.0;
                            for (int k = lane; k < q; k += 32) tv += qv.tmp[k] * qv.aval[8 * k + 6];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) tv += __shfl_xor_sync(0xffffffffu, tv, o);
                            if (lane == 0) s.ze[SL] = src[SL] - kslack * tv;
                        }
                    }
                    if (pass == 0) { for (int k = tid; k < q; k += nt) qv.r[k] = qv.tmp[k]; }
                    else { for (int k = tid; k < q; k += nt) qv.r[k] += qv.tmp[k]; }
                    blk.sync();
                    // ---- F: e = N_W' ze (vanishes in exact arithmetic), dual step length
                    double emax = 0.0;
                    t1 = INFINITY;
                    l = 0x7fffffff;
                    for (int k = tid; k < q; k += nt) {
                        const double* av = qv.aval + 8 * k;
                        const unsigned char* ai = reinterpret_cast<const unsigned char*>(av + 7);
                        double e = av[6] * s.ze[SL];
#pragma unroll
                        for (int j = 0; j < 6; ++j) e += av[j] * s.ze[ai[j]];
                        qv.w[k] = e;
                        emax = fmax(emax, fabs(e));
                        const double rk = qv.r[k];
                        if (rk > 1e-13) {
                            const double tk = qv.u[k] / rk;
                            if (tk < t1 || (tk == t1 && k < l)) { t1 = tk; l = k; }
                        }
                    }
                    // one barrier for both reductions: the refinement flag rides on the arg-min exchange
                    for (int o = 16; o > 0; o >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
                    if (lane == 0) gsc[32 + 8 * pass + warp] = emax;
                    blk.argmin(t1, l);
                    emax = 0.0;
                    for (int i = 0; i < (nt >> 5); ++i) emax = fmax(emax, gsc[32 + 8 * pass + i]);
                    if (pass == 1 || !(emax > refine_tol * fabs(dn))) break;
                    blk.count(CT_GI_REFINE);
                }
                d2n = pel * s.ze[SL];
#pragma unroll
                for (int j = 0; j < 6; ++j) d2n += pval[j] * s.ze[pidx[j]];
            }
            blk.mark(PH_GI_Z);
            const bool dep = (q >= qcap) || !(d2n > dep_tol * fmax(1.0, dn)) || (d2n <= 1e-28);
            const double t2 = dep ? INFINITY : (-sp / d2n);
            const double t = fmin(t1, t2);
            if (t == INFINITY) { status = (q >= qcap) ? 5 : GI_INFEASIBLE; break; }
            // ---- G: step
            for (int j = tid; j <= q; j += nt) qv.u[j] += t * ((j < q) ? -qv.r[j] : 1.0);
            if (t2 != INFINITY) {
                apply_step(t, p);
                sp += t * d2n;
                blk.mark(PH_GI_STEP);
                if (t == t2) {
                    // full step: p joins the working set
                    const double rho = sqrt(d2n);
                    double* col = Ui + gi_tri(q);
                    for (int j = tid; j < q; j += nt) col[j] = -qv.r[j] / rho;
                    if (tid == nt - 1) {
                        col[q] = 1.0 / rho;
                        qv.act[q] = p;
                        s.pos[p] = (short)q;
                        double* av = qv.aval + 8 * q;
                        unsigned char* ai = reinterpret_cast<unsigned char*>(av + 7);
#pragma unroll
                        for (int j = 0; j < 6; ++j) { av[j] = pval[j]; ai[j] = (unsigned char)pidx[j]; }
                        av[6] = pel;
                    }
                    blk.argmin(nb, nbi);           // barrier of the update + next selection in one
                    q += 1;
                    added = true;
                    blk.mark(PH_GI_UPD);
                    continue;
                }
            }
            blk.sync();
            // ---- drop the l-th member of the working set (partial step or dual step): Givens on R^-1 only
            {
                for (int k = l + tid; k <= q - 2; k += nt) {
                    double ss = 0.0;
                    for (int j = l; j <= k; ++j) {
                        const double a = Ui[gi_tri(j) + l];
                        ss += a * a;
                    }
                    const double b = Ui[gi_tri(k + 1) + l];
                    const double carry = (k == l) ? Ui[gi_tri(l) + l] : sqrt(ss);
                    const double h = sqrt(ss + b * b);
                    double c = 1.0, sn = 0.0;
                    if (h > 0.0) { c = b / h; sn = carry / h; }
                    qv.cs[2 * k] = c;
                    qv.cs[2 * k + 1] = sn;
                }
                blk.sync();
                for (int j = tid; j < q; j += nt) {
                    if (j < l) {
                        double carry = Ui[gi_tri(l) + j];
                        for (int k = l; k <= q - 2; ++k) {
                            const double c = qv.cs[2 * k], sn = qv.cs[2 * k + 1], b = Ui[gi_tri(k + 1) + j];
                            Ui[gi_tri(k) + j] = c * carry - sn * b;
                            carry = sn * carry + c * b;
                        }
                    } else if (j > l) {
                        double carry = 0.0;
                        for (int k = j - 1; k <= q - 2; ++k) {
                            const double c = qv.cs[2 * k], sn = qv.cs[2 * k + 1], b = Ui[gi_tri(k + 1) + j];
                            const double nk = c * carry - sn * b;
                            if (k == j - 1) qv.sub[j] = nk; else Ui[gi_tri(k) + j] = nk;
                            carry = sn * carry + c * b;
                        }
                    }
                }
                // shifted copies of the per-member data (multiplier, id, cached normal)
                for (int i = l + tid; i < q - 1; i += nt) {
                    qv.tmp[i] = qv.u[i + 1];
                    qv.itmp[i] = qv.act[i + 1];
                }
                if (tid == nt - 1) { qv.tmp[q - 1] = qv.u[q]; s.pos[qv.act[l]] = (short)-1; }
                const int nsh = (q - 1 - l) * 8;                       // cached normals of the members behind l move up by one
                double keep0 = 0.0, keep1 = 0.0;
                if (tid < nsh) keep0 = qv.aval[8 * (l + 1) + tid];
                if (tid + nt < nsh) keep1 = qv.aval[8 * (l + 1) + tid + nt];
                blk.sync();
                for (int k = l + tid; k <= q - 2; k += nt) {
                    double* col = Ui + gi_tri(k);
                    for (int j = l; j < k; ++j) col[j] = col[j + 1];
                    col[k] = qv.sub[k + 1];
                }
                for (int i = l + tid; i < q; i += nt) {
                    qv.u[i] = qv.tmp[i];
                    if (i < q - 1) { qv.act[i] = qv.itmp[i]; s.pos[qv.itmp[i]] = (short)i; }
                }
                if (tid < nsh) qv.aval[8 * l + tid] = keep0;
                if (tid + nt < nsh) qv.aval[8 * l + tid + nt] = keep1;
                blk.sync();
                q -= 1;
                blk.mark(PH_GI_DROP);
                blk.count(CT_GI_DROP);
            }
        }
        if (status != GI_OK) break;
    }
    for (int i = tid; i < m; i += nt) lam[i] = 0.0;
    blk.sync();
    for (int j = tid; j < q; j += nt) lam[qv.act[j]] = qv.u[j];
    blk.sync();
    blk.count(CT_GI_ITER, iters);
    *iters_out = iters;
    *nact_out = q;
    return status;
}

// =====================================================================================================================
// phase_qp2: Hessian schedule + condensing + block sweep + active-set QP for one SQP iteration (CTA-cooperative).
// `staged`: the linearisation left Jz / Wz in this phase's scratch (s.Jz, s.Wz).
// =====================================================================================================================
__device__ __forceinline__ void phase_qp2(CudaBlock& blk, const ftmpc_config& cfg, const WsLayout& L, const StepIO& io, int inst,
                                          int slot, double* scratch, bool staged) {
    double* w = ws_slot(io, L, slot);
    double* sc = w + L.oSc;
    if (sc[SC_STATUS] != FTMPC_ST_RUNNING) return;
    const int N = L.N, n = L.n, nv = L.nv, tid = blk.tid(), nt = blk.nthreads(), NB = N + 2;
    const Qp2Scratch s = qp2_carve(scratch, N, io);
    const double* xref = io.xref + (size_t)inst * io.xref_stride;
    const double* hull_g = io.hull_table + (size_t)io.hull_idx[inst] * FTMPC_HULL_STRIDE;
    for (int i = tid; i < L.mc; i += nt) s.cv[i] = w[L.oC + i];
    for (int i = tid; i < FTMPC_HULL_STRIDE; i += nt) s.hull[i] = hull_g[i];
    blk.sync();
    const double* lam_prev = w + L.oLam;
    const double feas_aug = 1e-2;
    const bool can_aug = (sc[SC_ITER] > 0.0 || sc[SC_THETA] >= 0.0) && sc[SC_CSUM] <= feas_aug;
    const double dprev_keep = (sc[SC_ITER] > 0.0 && sc[SC_THETA] == 1.0 && sc[SC_ALPHA] == 1.0) ? sc[SC_DMAX] : 0.0;
    double theta = sc[SC_THETA], sigma = 0.0;
    theta = (theta < 0.0) ? 0.0 : ((theta == 0.0) ? cfg.theta_first : fmin(1.0, cfg.theta_growth * theta));
    if (can_aug && sc[SC_SIGMA] > 0.0) { theta = 1.0; sigma = sc[SC_SIGMA]; }
    const bool far = sc[SC_ITER] > 0.0 && sc[SC_THETA] <= 0.0 && sc[SC_DMAX] > cfg.blend_dmax;
    const bool skip_exact = (sc[SC_HFAIL] != 0.0 && sc[SC_CSUM] > feas_aug) || far;
    if (skip_exact) theta = 0.0;
    bool aug_allowed = can_aug;
    int fails = 0, qit = 0, nact = 0, st = GI_OK, aug_retry = 0;
    bool have_w = staged;
    // block role of this thread
    int bi = -1, bj = 0;
    if (tid < NB * (NB + 1) / 2) q2_block_of(tid, bi, bj);
    const double* Cq = io.uref ? w + L.oCq : nullptr;
    for (;;) {        // QP attempts (re-solved without augmentation if a predicted-active row came out inactive)
        double sig0 = 0.0;
        for (;;) {
            // the linearisation leaves Wz in place (shared memory); the condensing scales it in place, so a second attempt
            // re-reads it from the global backing copy.  The stage Jacobians are read from the CTA's global slot.
            if (!have_w) for (int i = tid; i < N * 169; i += nt) s.Wz[i] = w[L.oWz + i];
            if (sigma > 0.0) for (int i = tid; i < L.mc; i += nt) s.lam_prev[i] = lam_prev[i];
            for (int i = tid; i < 90; i += nt) s.hv[i] = (i < 81) ? w[L.oHV + i] : w[L.oGV + i - 81];
            blk.sync();
            double acc[6][6];
            condense2(blk, cfg, L, s, w + L.oJz, w + L.oX, w + L.oU, xref, theta, sigma, Cq, acc, bi, bj);
            blk.mark(PH_COND);
            blk.count(CT_CONDENSE);
            double dmaxl = 0.0;
            if (bi >= 0 && bi < N && bi == bj) {
#pragma unroll
                for (int i = 0; i < 6; ++i) dmaxl = fmax(dmaxl, fabs(acc[i][i]));
            }
            const double dscale = blk.max(dmaxl);
            const int bad = chol_k_blocks(blk, N, s, acc, bi, bj, 1e-10 * fmax(1.0, dscale));
            have_w = false;                 // the condensing scaled Wz in place: reload on a retry
            if (!bad) break;
            blk.count(CT_CHOL_FAIL);
            ++fails;
            blk.sync();
            if (sigma == 0.0 && theta == 1.0 && aug_allowed) { sig0 = 10.0 * dscale; sigma = sig0; }
            else if (sigma > 0.0 && sig0 > 0.0 && sigma < 5.0 * sig0) { sigma *= 10.0; }
            else if (sigma > 0.0) { sigma = 0.0; theta = 0.5; aug_allowed = false; }
            else if (theta <= 0.0) {
                if (tid == 0) sc[SC_QPST] = 3.0;
                blk.sync();
                return;
            }
            else { theta = (theta > cfg.theta_first) ? 0.5 * theta : 0.0; }
        }
        // unconstrained minimiser  x = -K [ga ; 0]  (K coordinates; the elastic variable has no gradient) -> s.ze
        {
            const int NK = n + FTMPC_NE, pair = tid >> 1, part = tid & 1;
            for (int r0 = 0; r0 < NK; r0 += nt >> 1) {
                const int row = r0 + pair;
                double sv = 0.0;
                if (row < NK) {
                    const int tr = (row * (row + 1)) >> 1;
                    for (int j = part; j < n; j += 2) sv += ((j <= row) ? s.K[tr + j] : s.K[((j * (j + 1)) >> 1) + row]) * s.ga[j];
                }
                sv += __shfl_xor_sync(0xffffffffu, sv, 1);
                if (row < NK && part == 0) s.ze[row] = -sv;
            }
            if (tid == nt - 1) s.ze[NK] = 0.0;
        }
        blk.sync();
        MpcCons cons{N, n, nv, L.mc, s.hull, io.cfg_g->Af, s.cv, io.tf_val, io.tf_idx};
        blk.mark(PH_QPSETUP);
        blk.count(CT_QP);
        int qit1 = 0;
        const double kslack = 1.0 / cfg.rho_slack;
        st = gis2_solve(blk, cons, s, kslack, w + L.oLam + L.m, cfg.max_qp_iter, cfg.qp_tol, &qit1, &nact);
        blk.mark(PH_GI);
        qit += qit1;
        if (sigma > 0.0 && st == GI_OK) {       // every predicted-active row must be active in the QP solution
            int viol = 0;
            for (int i = tid; i < L.mc; i += nt)
                if (lam_prev[i] > 0.0 && s.pos[i] < 0 && s.s[i] > 1e-9) viol = 1;
            if (blk.any(viol)) {
                ++fails;
                if (aug_retry < 2) {
                    ++aug_retry;
                    double* lp = w + L.oLam;
                    for (int i = tid; i < L.mc; i += nt)
                        if (lp[i] > 0.0 && s.pos[i] < 0 && s.s[i] > 1e-9) lp[i] = 0.0;
                    blk.sync();
                    continue;
                }
                sigma = 0.0; theta = 0.5; aug_allowed = false;
                continue;
            }
        }
        break;
    }
    for (int i = tid; i < L.m; i += nt) w[L.oLam + i] = w[L.oLam + L.m + i];
    blk.sync();
    double gd = 0.0, dmx = 0.0, lmx = 0.0;
    for (int i = tid; i < n; i += nt) {
        const double di = s.xe[i];
        w[L.oD + i] = di;
        gd += s.g[i] * di;
        dmx = fmax(dmx, fabs(di));
    }
    for (int i = tid; i < L.mc; i += nt) lmx = fmax(lmx, w[L.oLam + i]);
    gd = blk.sum(gd);
    dmx = blk.max(dmx);
    lmx = blk.max(lmx);
    if (tid == 0) {
        const double delta = s.xe[n + FTMPC_NE];
        w[L.oD + n] = delta;
        sc[SC_DPREV] = dprev_keep;
        sc[SC_GD] = gd; sc[SC_DMAX] = dmx; sc[SC_LAMMAX] = lmx; sc[SC_DELTA] = delta;
        sc[SC_HFAIL] = (skip_exact || (fails > 0 && theta == 0.0)) ? 1.0 : 0.0;
        sc[SC_THETA] = theta; sc[SC_SIGMA] = sigma; sc[SC_QPIT] += qit; sc[SC_NACT] = nact; sc[SC_CHOLFAIL] += fails;
        sc[SC_QPST] = (st == GI_OK && dmx == dmx) ? 0.0 : (double)(st ? st : 4);
    }
    blk.sync();
    blk.mark(PH_POST);
}

}  // namespace ftmpc
