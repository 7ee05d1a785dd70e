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
__device__ __forceinline__ int q2_blk(int bi, int bj) { return (bi * (bi + 1) / 2 + bj) * 36; }
// K(r, c) in K coordinates (0..6N-1 decision variables, 6N..6N+8 extension rows); the diagonal blocks are stored full
__device__ __forceinline__ double q2_K(const double* K, int r, int c) {
    const int br = r / 6, bc = c / 6;
    return (br >= bc) ? K[q2_blk(br, bc) + (r - 6 * br) * 6 + (c - 6 * bc)] : K[q2_blk(bc, br) + (c - 6 * bc) * 6 + (r - 6 * br)];
}

struct Qp2Scratch {
    // live for the whole phase
    double *cv, *hull, *g, *ga;
    // condensing
    double *panel, *Jz, *Wz, *qe, *Ht, *tgv, *lam_prev, *hv, *cqs;
    // block sweep
    double *Lcol, *Xrow, *linv, *flag;
    // active-set iteration
    double *K, *xe, *s, *ye, *ze, *c;
    struct QVecs { double *Ui, *u, *w, *v, *r, *cs, *tmp, *sub; int *act, *itmp; int qcap; } qv;
    short* pos;
    unsigned* smask;      // [N] hull rows in the working set per stage, [3] terminal rows, [1] the two rows of the elastic variable
    int* blist;           // blocks of K coordinates with a non-zero entry of c (uniform list), [NB + 1]
    const double* tf_val;
    const int* tf_idx;
    const ftmpc_config* cg;
    size_t total;
};
__host__ __device__ inline size_t qp2_fixed_doubles(int N) {
    const WsLayout L = ws_layout(N);
    return (size_t)((L.mc + 1) & ~1) + ((FTMPC_HULL_STRIDE + 1) & ~1) + 2 * (size_t)((L.nv + 1) & ~1);
}
__host__ __device__ inline size_t qp2_scratch_doubles(int N) {
    const WsLayout L = ws_layout(N);
    const int NB = N + 2, ldp = 7 * N + 1, ne = L.nv + FTMPC_NE, qc = FTMPC_Q2_QCAP;
    const size_t cond = (size_t)32 * ldp + 2 * (size_t)N * 169 + (size_t)N * FTMPC_NE + 90 + L.mc + 90 + (size_t)N * 10 + 8;
    const size_t kblk = (size_t)NB * (NB + 1) / 2 * 36;
    const size_t chol = kblk + 3 * (size_t)NB * FTMPC_Q2_BS + 8;
    const size_t ints = ((size_t)2 * (qc + 2) + (L.m + 1) / 2 + (N + 4) + (NB + 2) + 1) / 2 + 1;
    const size_t gi = kblk + (size_t)qc * (qc + 1) / 2 + 4 * (size_t)(ne + 1) + (L.m + 2) + 6 * (size_t)(qc + 2) + 2 * (size_t)(qc + 2) + ints + 8;
    size_t r = cond > chol ? cond : chol;
    if (gi > r) r = gi;
    return qp2_fixed_doubles(N) + r;
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
    s.panel = p; p += (size_t)32 * ldp;
    s.qe = p; p += (size_t)N * FTMPC_NE;
    s.Ht = p; p += 81;
    s.tgv = p; p += 9;
    s.lam_prev = p; p += L.mc;
    s.hv = p; p += 90;
    s.cqs = p; p += (size_t)N * 10;
    s.Jz = p; p += (size_t)N * 169;          // last: phase_lin2 leaves the stage Jacobians / Hessians here
    s.Wz = p; p += (size_t)N * 169;
    // block sweep: K first (written when the sweep has succeeded), the shared block column / row behind it
    const size_t kblk = (size_t)NB * (NB + 1) / 2 * 36;
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
    s.c = p; p += ne + 1;
    s.s = p; p += L.m + 2;
    s.qv.u = p; p += qc + 2;
    s.qv.w = p; p += qc + 2;
    s.qv.v = p; p += qc + 2;
    s.qv.r = p; p += qc + 2;
    s.qv.tmp = p; p += qc + 2;
    s.qv.sub = p; p += qc + 2;
    s.qv.cs = p; p += 2 * (qc + 2);
    int* ip = reinterpret_cast<int*>(p);
    s.qv.act = ip; ip += qc + 2;
    s.qv.itmp = ip; ip += qc + 2;
    s.smask = reinterpret_cast<unsigned*>(ip); ip += N + 4;
    s.blist = ip; ip += NB + 2;
    s.pos = reinterpret_cast<short*>(ip);
    s.total = qp2_scratch_doubles(N);
    return s;
}

// q-sized vectors of the overflow path (working sets beyond FTMPC_Q2_QCAP rows): capacity nv, global memory
__host__ __device__ inline size_t qp2_overflow_doubles(int N) {
    const size_t nv = 6 * (size_t)N + 1;
    return nv * (nv + 1) / 2 + 8 * (nv + 2) + (nv + 2) + 8;
}
__device__ __forceinline__ Qp2Scratch::QVecs qp2_overflow_carve(double* g, int N) {
    const int nv = 6 * N + 1;
    Qp2Scratch::QVecs q;
    double* p = g;
    q.qcap = nv;
    q.Ui = p; p += (size_t)nv * (nv + 1) / 2;
    q.u = p; p += nv + 2;
    q.w = p; p += nv + 2;
    q.v = p; p += nv + 2;
    q.r = p; p += nv + 2;
    q.tmp = p; p += nv + 2;
    q.sub = p; p += nv + 2;
    q.cs = p; p += 2 * (nv + 2);
    int* ip = reinterpret_cast<int*>(p);
    q.act = ip; ip += nv + 2;
    q.itmp = ip;
    return q;
}

// =====================================================================================================================
// condensing: H blocks -> registers of the block threads, g / ga -> shared memory, X = d x_N[0:9]/dU -> the extension
// block rows.  Same arithmetic as the register-tiled path of ftmpc_sqp.cuh::condense (see there for the derivation).
// =====================================================================================================================
__device__ __forceinline__ void condense2(CudaBlock& blk, const ftmpc_config& cfg, const WsLayout& L, const Qp2Scratch& s,
                                          const double* X, const double* U, const double* xref, double theta, double sigma,
                                          const double* Cq, double (&acc)[6][6], int bi, int bj) {
    const int N = L.N, n = L.n, tid = blk.tid(), nt = blk.nthreads();
    const int ldp = 7 * N + 1;
    double* buf = s.panel;                         // rows 0-12 G_t, 13-25 M_t G_t, 26-31 theta W_ux G_t
    double* Wp = s.Wz;
    const double* Jz = s.Jz;
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
    for (int t = 0; t <= N; ++t) {
        // ---------------- column phase of stage t: G_t[:, a] from G_{t-1}[:, a] (held in the panel), then publish
        if (a >= 0 && ta < t) {
            double g[FTMPC_NX];
            if (ta == t - 1) {                     // birth: G_t[:, a] = B_{t-1} e_ja
                const double* jz = Jz + (size_t)(t - 1) * 169;
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
                const double* jz = Jz + (size_t)(t - 1) * 169;
                double gp[FTMPC_NX];
#pragma unroll
                for (int r = 0; r < FTMPC_NX; ++r) gp[r] = buf[r * ldp + pa_];
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
                for (int r = 0; r < 6; ++r) buf[(13 + r) * ldp + pa_] = 2.0 * cfg.Q[r] * g[r];
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    double v = 0.0;
#pragma unroll
                    for (int l = 0; l < 7; ++l) v += wp[k * 13 + l] * g[6 + l];
                    buf[(19 + k) * ldp + pa_] = v;
                }
#pragma unroll
                for (int i = 0; i < FTMPC_NU; ++i) {
                    double v = 0.0;
#pragma unroll
                    for (int l = 0; l < 7; ++l) v += wp[(7 + i) * 13 + l] * g[6 + l];
                    buf[(26 + i) * ldp + pa_] = v;
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
        blk.sync();
        // ---------------- block phase of stage t
        if (bi >= 0 && bi < N) {
            if (bi < t) {
                // rank-13 update (9 rows at the terminal stage).  The row loop stays ROLLED: unrolled it is 10 KB of
                // straight-line code per stage, and the capture of the unrolled version showed the FMA lines stalled on
                // instruction fetch (no_inst 45 %) with two CTAs sharing the instruction cache
                const double* Pa = buf + 7 * bi;
                const double* Tb = buf + (size_t)13 * ldp + 7 * bj;
                const int nr = (t < N) ? FTMPC_NX : FTMPC_NE;
                double pa[6], tb[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) { pa[i] = Pa[i]; tb[i] = Tb[i]; }
#pragma unroll 1
                for (int r = 0; r < nr; ++r) {
                    double pn[6], tn[6];
                    const int rn = (r + 1 < nr) ? r + 1 : r;       // operands of the next row in flight during the FMAs
#pragma unroll
                    for (int i = 0; i < 6; ++i) { pn[i] = Pa[(size_t)rn * ldp + i]; tn[i] = Tb[(size_t)rn * ldp + i]; }
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int j = 0; j < 6; ++j) acc[i][j] += pa[i] * tb[j];
#pragma unroll
                    for (int i = 0; i < 6; ++i) { pa[i] = pn[i]; tb[i] = tn[i]; }
                }
            } else if (bi == t) {
                if (bj < t) {
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int j = 0; j < 6; ++j) acc[i][j] = buf[(size_t)(26 + i) * ldp + 7 * bj + j];
                } else {
                    const double* wp = Wp + (size_t)t * 169;
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
                            if (lam_prev[t * FTMPC_NH + r] > 0.0) {
#pragma unroll
                                for (int i = 0; i < 6; ++i)
#pragma unroll
                                    for (int j = 0; j < 6; ++j) acc[i][j] += sigma * Ah[r * FTMPC_NU + i] * Ah[r * FTMPC_NU + j];
                            }
                        }
                    }
                }
            }
        } else if (bi >= N && t == N) {
            // extension block rows: X = G_N[0:9, :] (rows 6..8 of the second one are padding)
            if (bj < N) {
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
        double* o = s.K + q2_blk(bi, bj);
        const double sg = (bi >= N && bj >= N) ? -1.0 : 1.0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) o[i * 6 + j] = sg * acc[i][j];
        if (bi == bj) {                                       // diagonal blocks are kept full: mirror the lower triangle
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) if (j > i) o[i * 6 + j] = sg * acc[j][i];
        }
    }
    (void)NB;
    blk.sync();
    blk.mark(PH_INV);
    return 0;
}

// =====================================================================================================================
// range-space dual active-set iteration on the block-packed K (CUDA specialisation of ftmpc_gis.cuh::gis_solve for the
// MPC constraint structure).  Vectors (xe, ye, ze, c) use the extended layout of MpcCons: [d (n) ; delta ; X d (9)].
// =====================================================================================================================
struct Q2Row {           // constraint normal in K coordinates
    int kind;            // 0 hull row, 1 terminal row, 2 delta >= 0, 3 delta <= 1
    int base;            // first K coordinate of the (up to 6) consecutive entries of a hull row
    double val[6];       // hull: -A_h[i][0..5];  terminal: val[0], val[1] at ext coordinates e0, e1
    int e0, e1;
    double el;           // coefficient on the elastic variable
};
__device__ __forceinline__ void q2_row(const MpcCons& cons, int p, Q2Row& r) {
    const int N = cons.N;
    if (p < FTMPC_NH * N) {
        const int t = p / FTMPC_NH, i = p - t * FTMPC_NH;
        r.kind = 0; r.base = t * FTMPC_NU;
#pragma unroll
        for (int j = 0; j < 6; ++j) r.val[j] = -cons.Ah[i * FTMPC_NU + j];
        const double c = cons.cv[p];
        r.el = (c > 0.0) ? c : 0.0;
        r.e0 = r.e1 = 0;
    } else if (p < cons.mc) {
        const int i = p - FTMPC_NH * N;
        r.kind = 1; r.base = 0;
        r.e0 = cons.tf_idx[2 * i]; r.e1 = cons.tf_idx[2 * i + 1];
        r.val[0] = -cons.tf_val[2 * i]; r.val[1] = -cons.tf_val[2 * i + 1];
#pragma unroll
        for (int j = 2; j < 6; ++j) r.val[j] = 0.0;
        const double c = cons.cv[p];
        r.el = (c > 0.0) ? c : 0.0;
    } else {
        r.kind = (p == cons.mc) ? 2 : 3; r.base = 0; r.e0 = r.e1 = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) r.val[j] = 0.0;
        r.el = (p == cons.mc) ? 1.0 : -1.0;
    }
}
// n_p . v for a vector in the extended layout
__device__ __forceinline__ double q2_dot(const Q2Row& r, const double* v, int n, int nv) {
    double a = r.el * v[n];
    if (r.kind == 0) {
#pragma unroll
        for (int j = 0; j < 6; ++j) a += r.val[j] * v[r.base + j];
    } else if (r.kind == 1) {
        a += r.val[0] * v[nv + r.e0] + r.val[1] * v[nv + r.e1];
    }
    return a;
}

// out[row] (+)= sign * sum_b K(row, block b) c_b  over the blocks in `bl`; rows in K coordinates, vectors in the extended
// layout.  Lane pairs: rows 0 .. 6N-1 on threads 0 .. 12N-1 (even / odd entries of the block list), the first 8
// extension rows on the next 16 threads, the last extension row on warp 0 afterwards.
template <bool SUB>
__device__ __forceinline__ void q2_matvec(const double* K, const double* c, const double* base_vec, double* out, const int* bl,
                                          int nbl, int N, int tid, int nt) {
    const int n = 6 * N, nv = n + 1, nrows = n + FTMPC_NE;
    const int pair = tid >> 1, part = tid & 1;
    {
        const bool on = pair < nrows - 1 || (pair == nrows - 1 && 2 * nrows <= nt);
        const int row = on ? pair : 0, br = row / 6, a = row - 6 * br;
        double s0 = 0.0, s1 = 0.0;
        for (int ii = part; on && ii < nbl; ii += 2) {
            const int b = bl[ii];
            const double* cb = c + ((b < N) ? 6 * b : 6 * b + 1);          // extension coordinates sit behind the elastic variable
            const int len = (b <= N) ? 6 : FTMPC_NE - 6;                    // the last block holds 3 extension rows
            if (b <= br) {
                const double* kb = K + q2_blk(br, b) + a * 6;
                s0 += kb[0] * cb[0] + kb[2] * cb[2];
                s1 += kb[1] * cb[1];
                if (len == 6) { s1 += kb[3] * cb[3] + kb[5] * cb[5]; s0 += kb[4] * cb[4]; }
            } else {
                const double* kb = K + q2_blk(b, br) + a;
                s0 += kb[0] * cb[0] + kb[12] * cb[2];
                s1 += kb[6] * cb[1];
                if (len == 6) { s1 += kb[18] * cb[3] + kb[30] * cb[5]; s0 += kb[24] * cb[4]; }
            }
        }
        double sv = s0 + s1;
        sv += __shfl_xor_sync(0xffffffffu, sv, 1);
        if (on && part == 0) {
            const int o = (row < n) ? row : row + 1;
            out[o] = SUB ? base_vec[o] - sv : -sv;
        }
    }
    if (2 * nrows > nt && tid < 32) {              // last extension row: warp 0, lanes over (block, entry)
        const int row = nrows - 1, br = row / 6, a = row - 6 * br;
        double sv = 0.0;
        for (int e = tid; e < 6 * nbl; e += 32) {
            const int ii = e / 6, j = e - 6 * ii, b = bl[ii];
            const int len = (b <= N) ? 6 : FTMPC_NE - 6;
            if (j < len) {
                const double cj = c[((b < N) ? 6 * b : 6 * b + 1) + j];
                sv += ((b <= br) ? K[q2_blk(br, b) + a * 6 + j] : K[q2_blk(b, br) + j * 6 + a]) * cj;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
        if (tid == 0) out[row + 1] = SUB ? base_vec[row + 1] - sv : -sv;
    }
    (void)nv;
}

// Ui access: shared memory (UIG = false) or the CTA's global slot (UIG = true, capacity nv)
template <bool UIG>
__device__ __forceinline__ int gis2_solve(CudaBlock& blk, const MpcCons& cons, const Qp2Scratch& s, const Qp2Scratch::QVecs& qv,
                                          double kslack, double* lam, int maxit, double tol, int* iters_out, int* nact_out) {
    double* const Ui = qv.Ui;
    const int qcap = qv.qcap;
    const int tid = blk.tid(), nt = blk.nthreads(), lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int N = cons.N, n = cons.n, nv = cons.nv, ne = nv + FTMPC_NE, m = cons.mc + 2, NB = N + 2;
    const double dep_tol = 1e-14, refine_tol = 1e-13;
    double* gsc = blk.scratch + 128;
    int q = 0, iters = 0, status = GI_OK;
    for (int i = tid; i < m; i += nt) {
        s.s[i] = cons.slack(i, s.xe, 1.0);
        s.pos[i] = (short)-1;
    }
    for (int i = tid; i < N + 4; i += nt) s.smask[i] = 0u;
    if (tid == 0) s.blist[NB + 1] = 0;
    blk.sync();
    bool have_next = false;
    double next_best = 0.0;
    int next_bi = 0x7fffffff;
    // c = N_W r in the extended layout, gathered per coordinate from the per-stage masks of the working set
    auto gather_c = [&]() {
        for (int i = tid; i < ne; i += nt) {
            double a = 0.0;
            if (i < n) {
                const int t = i / FTMPC_NU, j = i - t * FTMPC_NU;
                unsigned mk = s.smask[t];
                while (mk) {
                    const int row = __ffs(mk) - 1;
                    mk &= mk - 1;
                    a -= qv.r[s.pos[t * FTMPC_NH + row]] * cons.Ah[row * FTMPC_NU + j];
                }
            } else if (i > n) {
                const int e = i - nv;
                for (int wd = 0; wd < 3; ++wd) {
                    unsigned mk = s.smask[N + wd];
                    while (mk) {
                        const int row = 32 * wd + __ffs(mk) - 1;
                        mk &= mk - 1;
                        const double av = (s.tf_idx[2 * row] == e) ? s.tf_val[2 * row] : ((s.tf_idx[2 * row + 1] == e) ? s.tf_val[2 * row + 1] : 0.0);
                        a -= qv.r[s.pos[FTMPC_NH * N + row]] * av;
                    }
                }
            }
            if (i != n) s.c[i] = a;
        }
        if (warp == nw - 1) {                      // elastic coordinate: every member of the working set may touch it
            double a = 0.0;
            for (int k = lane; k < q; k += 32) {
                const int p = qv.act[k];
                const double cv = (p < cons.mc) ? cons.cv[p] : 0.0;
                const double el = (p < cons.mc) ? ((cv > 0.0) ? cv : 0.0) : ((p == cons.mc) ? 1.0 : -1.0);
                a += qv.r[k] * el;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) s.c[n] = a;
        }
    };
    // list of K blocks with a non-zero c (thread 0; called when the working set changed)
    auto rebuild_blist = [&]() {
        if (tid == 0) {
            int c = 0;
            for (int t = 0; t < N; ++t) if (s.smask[t]) s.blist[c++] = t;
            if (s.smask[N] | s.smask[N + 1] | s.smask[N + 2]) { s.blist[c++] = N; s.blist[c++] = N + 1; }
            s.blist[NB + 1] = c;
        }
    };
    // r = R^-1 R^-T w  (w in s.w): four lanes per entry
    auto schur_solve = [&]() {
        const int part = tid & 3, per = nt >> 2;
        for (int k0 = 0; k0 < q; k0 += per) {         // (uniform trip count: the shuffles run in every lane)
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
            if (k < q && part == 0) qv.r[k] = a;
        }
        blk.sync();
    };
    for (;;) {
        // ---- most violated row
        double best = next_best;
        int bi = next_bi;
        if (!have_next) {
            best = 0.0;
            bi = 0x7fffffff;
            for (int i = tid; i < m; i += nt) {
                if (s.pos[i] < 0) {
                    const double v = s.s[i];
                    if (v < best || (v == best && i < bi)) { best = v; bi = i; }
                }
            }
            blk.argmin(best, bi);
        }
        if (bi == 0x7fffffff || best >= -tol) break;
        have_next = false;
        blk.mark(PH_GI_SELECT);
        const int p = bi;
        double sp = best;
        Q2Row np;
        q2_row(cons, p, np);
        if (tid == 0) qv.u[q] = 0.0;
        // ye = K n_p (extended layout); the elastic variable is decoupled: K_delta = 1 / rho_slack
        for (int i = tid; i < ne; i += nt) {
            double a = 0.0;
            if (i == n) a = np.el * kslack;
            else {
                const int r = (i < n) ? i : i - 1;
                if (np.kind == 0) {
#pragma unroll
                    for (int j = 0; j < 6; ++j) a += np.val[j] * q2_K(s.K, r, np.base + j);
                } else if (np.kind == 1) {
                    a = np.val[0] * q2_K(s.K, r, n + np.e0) + np.val[1] * q2_K(s.K, r, n + np.e1);
                }
            }
            s.ye[i] = a;
        }
        blk.sync();
        const double dn = q2_dot(np, s.ye, n, nv);
        blk.mark(PH_GI_D);
        bool added = false;
        while (!added) {
            ++iters;
            if (iters > maxit) { status = GI_MAXIT; break; }
            double d2n, t1 = INFINITY;
            int l = 0x7fffffff;
            if (q > 0) {
                for (int k = tid; k < q; k += nt) qv.w[k] = cons.slack(qv.act[k], s.ye, 0.0);
                blk.sync();
                schur_solve();
                gather_c();
                blk.sync();
                q2_matvec<true>(s.K, s.c, s.ye, s.ze, s.blist, s.blist[NB + 1], N, tid, nt);
                if (tid == nt - 1) s.ze[n] = s.ye[n] - kslack * s.c[n];
                blk.sync();
                // refinement on the semi-normal equations: e = N_W' ze vanishes in exact arithmetic
                double emax = 0.0;
                for (int k = tid; k < q; k += nt) {
                    const double e = cons.slack(qv.act[k], s.ze, 0.0);
                    qv.w[k] = e;
                    qv.tmp[k] = qv.r[k];
                    emax = fmax(emax, fabs(e));
                }
                emax = blk.max(emax);
                if (emax > refine_tol * fabs(dn)) {
                    schur_solve();                 // delta r
                    gather_c();
                    blk.sync();
                    q2_matvec<true>(s.K, s.c, s.ze, s.ze, s.blist, s.blist[NB + 1], N, tid, nt);
                    if (tid == nt - 1) s.ze[n] -= kslack * s.c[n];
                    for (int k = tid; k < q; k += nt) qv.r[k] += qv.tmp[k];
                    blk.sync();
                    blk.count(CT_GI_REFINE);
                }
                d2n = q2_dot(np, s.ze, n, nv);
                for (int j = tid; j < q; j += nt) {
                    const double rj = qv.r[j];
                    if (rj > 1e-13) {
                        const double tj = qv.u[j] / rj;
                        if (tj < t1 || (tj == t1 && j < l)) { t1 = tj; l = j; }
                    }
                }
                blk.argmin(t1, l);
            } else {
                for (int i = tid; i < ne; i += nt) s.ze[i] = s.ye[i];
                blk.sync();
                d2n = dn;
            }
            blk.mark(PH_GI_Z);
            const bool dep = (q >= qcap) || !(d2n > dep_tol * fmax(1.0, dn)) || (d2n <= 1e-28);
            const double t2 = dep ? INFINITY : (-sp / d2n);
            const double t = fmin(t1, t2);
            if (t == INFINITY) { status = (q >= qcap) ? 5 : GI_INFEASIBLE; break; }
            if (t2 == INFINITY) {
                for (int j = tid; j <= q; j += nt) qv.u[j] += t * ((j < q) ? -qv.r[j] : 1.0);
                blk.sync();
            } else {
                const bool full = (t == t2);
                for (int i = tid; i < ne; i += nt) s.xe[i] += t * s.ze[i];
                for (int j = tid; j <= q; j += nt) qv.u[j] += t * ((j < q) ? -qv.r[j] : 1.0);
                double nb = 0.0;
                int nbi = 0x7fffffff;
                for (int i = tid; i < m; i += nt) {
                    const double v = s.s[i] + t * cons.slack(i, s.ze, 0.0);
                    s.s[i] = v;
                    if (i != p && s.pos[i] < 0 && (v < nb || (v == nb && i < nbi))) { nb = v; nbi = i; }
                }
                sp += t * d2n;
                blk.mark(PH_GI_STEP);
                if (full) {
                    const double rho = sqrt(d2n);
                    double* col = Ui + gi_tri(q);
                    for (int j = tid; j < q; j += nt) col[j] = -qv.r[j] / rho;
                    if (tid == nt - 1) {
                        col[q] = 1.0 / rho;
                        qv.act[q] = p;
                        s.pos[p] = (short)q;
                        if (p < FTMPC_NH * N) s.smask[p / FTMPC_NH] |= 1u << (p % FTMPC_NH);
                        else if (p < cons.mc) s.smask[N + (p - FTMPC_NH * N) / 32] |= 1u << ((p - FTMPC_NH * N) % 32);
                    }
                    blk.argmin(nb, nbi);           // barrier of the update + next selection in one
                    next_best = nb;
                    next_bi = nbi;
                    have_next = true;
                    q += 1;
                    added = true;
                    rebuild_blist();
                    if (tid == 0) s.s[p] = 0.0;    // on the constraint by construction
                    blk.sync();
                    blk.mark(PH_GI_UPD);
                    continue;
                }
                blk.sync();
            }
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
                for (int i = l + tid; i < q; i += nt) { qv.tmp[i] = qv.u[i + 1]; qv.itmp[i] = (i + 1 < q) ? qv.act[i + 1] : -1; }
                if (tid == 0) {
                    const int pd = qv.act[l];
                    s.pos[pd] = (short)-1;
                    if (pd < FTMPC_NH * N) s.smask[pd / FTMPC_NH] &= ~(1u << (pd % FTMPC_NH));
                    else if (pd < cons.mc) s.smask[N + (pd - FTMPC_NH * N) / 32] &= ~(1u << ((pd - FTMPC_NH * N) % 32));
                }
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
                rebuild_blist();
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
                                          int slot, double* scratch, bool staged, double* ui_global) {
    double* w = ws_slot(io, L, slot);
    double* sc = w + L.oSc;
    if (sc[SC_STATUS] != FTMPC_ST_RUNNING) return;
    const int N = L.N, n = L.n, nv = L.nv, ne = nv + FTMPC_NE, tid = blk.tid(), nt = blk.nthreads(), NB = N + 2;
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
    bool have_j = staged, have_w = staged;
    // block role of this thread
    int bi = -1, bj = 0;
    if (tid < NB * (NB + 1) / 2) q2_block_of(tid, bi, bj);
    const double* Cq = io.uref ? w + L.oCq : nullptr;
    for (;;) {        // QP attempts (re-solved without augmentation if a predicted-active row came out inactive)
        double sig0 = 0.0;
        for (;;) {
            // the linearisation leaves Jz / Wz in place; Wz is scaled in place by the condensing, so a second attempt
            // re-reads it (and, once the active-set solver has reused the region, Jz too) from the global backing copy
            if (!have_j) for (int i = tid; i < N * 169; i += nt) s.Jz[i] = w[L.oJz + i];
            if (!have_w) for (int i = tid; i < N * 169; i += nt) s.Wz[i] = w[L.oWz + i];
            if (sigma > 0.0) for (int i = tid; i < L.mc; i += nt) s.lam_prev[i] = lam_prev[i];
            for (int i = tid; i < 90; i += nt) s.hv[i] = (i < 81) ? w[L.oHV + i] : w[L.oGV + i - 81];
            blk.sync();
            double acc[6][6];
            condense2(blk, cfg, L, s, w + L.oX, w + L.oU, xref, theta, sigma, Cq, acc, bi, bj);
            blk.mark(PH_COND);
            blk.count(CT_CONDENSE);
            double dmaxl = 0.0;
            if (bi >= 0 && bi < N && bi == bj) {
#pragma unroll
                for (int i = 0; i < 6; ++i) dmaxl = fmax(dmaxl, fabs(acc[i][i]));
            }
            const double dscale = blk.max(dmaxl);
            const int bad = chol_k_blocks(blk, N, s, acc, bi, bj, 1e-10 * fmax(1.0, dscale));
            have_j = false;                 // the shared block column / row of the sweep may lie over the staged Jacobians
            have_w = false;                 // (short horizons), and the condensing scaled Wz in place: reload both on a retry
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
        have_j = false;                     // K has overwritten the staged Jacobians
        // unconstrained minimiser  x = -H^-1 ga  in the extended layout (the elastic variable has no gradient)
        if (tid == 0) {
            for (int b = 0; b < N; ++b) s.blist[b] = b;
            s.blist[NB + 1] = N;
        }
        for (int i = tid; i < ne + 1; i += nt) s.c[i] = (i < n) ? s.ga[i] : 0.0;
        blk.sync();
        q2_matvec<false>(s.K, s.c, s.c, s.xe, s.blist, N, N, tid, nt);
        if (tid == nt - 1) s.xe[n] = 0.0;
        blk.sync();
        MpcCons cons{N, n, nv, L.mc, s.hull, io.cfg_g->Af, s.cv, io.tf_val, io.tf_idx};
        blk.mark(PH_QPSETUP);
        blk.count(CT_QP);
        int qit1 = 0;
        const double kslack = 1.0 / cfg.rho_slack;
        st = gis2_solve<false>(blk, cons, s, s.qv, kslack, w + L.oLam + L.m, cfg.max_qp_iter, cfg.qp_tol, &qit1, &nact);
        if (st == 5) {
            // the working set outgrew the shared-memory R^-1: same QP again with R^-1 in the CTA's global slot
            qit += qit1;
            for (int i = tid; i < ne + 1; i += nt) s.c[i] = (i < n) ? s.ga[i] : 0.0;
            if (tid == 0) {
                for (int b = 0; b < N; ++b) s.blist[b] = b;
                s.blist[NB + 1] = N;
            }
            blk.sync();
            q2_matvec<false>(s.K, s.c, s.c, s.xe, s.blist, N, N, tid, nt);
            if (tid == nt - 1) s.xe[n] = 0.0;
            blk.sync();
            st = gis2_solve<true>(blk, cons, s, qp2_overflow_carve(ui_global, N), kslack, w + L.oLam + L.m, cfg.max_qp_iter, cfg.qp_tol, &qit1, &nact);
            if (st == 5) st = GI_MAXIT;
        }
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
        w[L.oD + n] = s.xe[n];
        sc[SC_DPREV] = dprev_keep;
        sc[SC_GD] = gd; sc[SC_DMAX] = dmx; sc[SC_LAMMAX] = lmx; sc[SC_DELTA] = s.xe[n];
        sc[SC_HFAIL] = (skip_exact || (fails > 0 && theta == 0.0)) ? 1.0 : 0.0;
        sc[SC_THETA] = theta; sc[SC_SIGMA] = sigma; sc[SC_QPIT] += qit; sc[SC_NACT] = nact; sc[SC_CHOLFAIL] += fails;
        sc[SC_QPST] = (st == GI_OK && dmx == dmx) ? 0.0 : (double)(st ? st : 4);
    }
    blk.sync();
    blk.mark(PH_POST);
}

}  // namespace ftmpc
