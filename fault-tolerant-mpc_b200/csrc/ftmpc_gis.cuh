// ftmpc_gis.cuh -- dual active-set QP (Goldfarb-Idnani) in RANGE-SPACE form: the iteration only needs the fixed matrix
//
//     K = [ G^-1      G^-1 X' ]      (symmetric, ne x ne; "extended inverse":  K = E E' for E = [J ; X J], J J' = G^-1)
//         [ X G^-1   X G^-1 X' ]
//
// and the inverse R^-1 of the Cholesky factor of the Schur complement  S = N_W' K N_W = R'R  of the working set W.
// Nothing of size ne x nv is rotated per iteration (ftmpc_gi.cuh rotates J = L^-T Q, a dense 130 x 121 matrix at N = 20,
// twice per added constraint) -- K is read-only and symmetric, so it is stored PACKED (66.5 KB instead of 126 KB at
// N = 20), which is what lets two instances share one SM.  With every constraint normal n_p sparse in the extended
// coordinates [x ; X x] (hull rows 6 + 1 entries, terminal rows <= 2 + 1):
//
//     ye  = K n_p                      sparse combination of <= 10 columns of K            (O(ne) per thread-row)
//     w_k = n_k . ye,  k in W          sparse dots
//     v   = R^-T w,   r = R^-1 v       two small triangular mat-vecs (q <= qcap)
//     c   = N_W r                      scattered into a dense ne-vector
//     ze  = ye - K c                   ONE symmetric mat-vec: the only O(ne^2) work of an iteration
//     |d2|^2 = n_p . ze,  |d|^2 = n_p . ye
//
// after which the step lengths, the primal/dual step, the slack update and the add / drop bookkeeping are those of the
// QR form (same pivoting rules, same tie breaks): add appends [-r/rho ; 1/rho] to R^-1, drop removes a column of R by
// Givens rotations applied to R^-1 alone.
//
// Numerics: the QR form gets |d2|^2 as a sum of squares; here it is a difference (ye - K c), so a linearly dependent
// constraint shows up as |d2|^2 ~ eps kappa |d|^2 instead of ~ eps^2 |d|^2.  One step of iterative refinement on the
// semi-normal equations (delta r = S^-1 N_W' ze, which would vanish in exact arithmetic) is applied whenever the active
// rows are not orthogonal to ze to working accuracy, and the dependency test uses a tolerance that matches:
// |d2|^2 <= dep_tol |d|^2  with dep_tol = 1e-14.
#pragma once
#include "ftmpc_gi.cuh"

namespace ftmpc {
#if defined(FTMPC_DEBUG_COUNTERS) && !defined(__CUDACC__)
extern long g_ftmpc_qmax_hist[16];      // histogram of the largest working set per QP, bins of 8
#endif

struct GisWork {
    const double* K;   // packed symmetric ne x ne, lower triangle by rows: K(i,j), j <= i, at i (i + 1) / 2 + j
    double* Ui;        // packed upper triangular R^-1 by columns, capacity qcap (qcap + 1) / 2
    double* xe;        // ne    extended primal [x ; X x]
    double* s;         // m     slack n_i . xe - beta_i
    double* u;         // qcap + 1 multipliers of the working set (+ candidate)
    double* ye;        // ne
    double* ze;        // ne
    double* c;         // ne
    double* w;         // qcap
    double* v;         // qcap
    double* r;         // qcap
    double* cs;        // 2 qcap rotation coefficients
    double* tmp;       // qcap + 1
    double* sub;       // qcap
    int* act;          // qcap + 1
    int* pos;          // m
    int* itmp;         // qcap + 1
    int qcap;          // capacity of the working set
};

FT_HD double gis_K(const double* K, int i, int j) { return (j <= i) ? K[(size_t)i * (i + 1) / 2 + j] : K[(size_t)j * (j + 1) / 2 + i]; }

// drop the l-th member of the working set: Givens rotations on R^-1 only (same coefficients as gi_drop)
template <class Blk>
FT_HD void gis_drop(Blk& blk, const GisWork& w, int& q, int l) {
    const int tid = blk.tid(), nt = blk.nthreads();
    for (int k = l + tid; k <= q - 2; k += nt) {
        double ss = 0.0;
        for (int j = l; j <= k; ++j) {
            const double a = w.Ui[gi_tri(j) + l];
            ss += a * a;
        }
        const double b = w.Ui[gi_tri(k + 1) + l];
        const double carry = (k == l) ? w.Ui[gi_tri(l) + l] : sqrt(ss);
        const double h = sqrt(ss + b * b);
        double c = 1.0, s = 0.0;
        if (h > 0.0) { c = b / h; s = carry / h; }
        w.cs[2 * k] = c;
        w.cs[2 * k + 1] = s;
    }
    blk.sync();
    for (int j = tid; j < q; j += nt) {
        if (j < l) {
            double carry = w.Ui[gi_tri(l) + j];
            for (int k = l; k <= q - 2; ++k) {
                const double c = w.cs[2 * k], s = w.cs[2 * k + 1], b = w.Ui[gi_tri(k + 1) + j];
                w.Ui[gi_tri(k) + j] = c * carry - s * b;
                carry = s * carry + c * b;
            }
        } else if (j > l) {
            double carry = 0.0;
            for (int k = j - 1; k <= q - 2; ++k) {
                const double c = w.cs[2 * k], s = w.cs[2 * k + 1], b = w.Ui[gi_tri(k + 1) + j];
                const double nk = c * carry - s * b;
                if (k == j - 1) w.sub[j] = nk; else w.Ui[gi_tri(k) + j] = nk;
                carry = s * carry + c * b;
            }
        }
    }
    for (int i = l + tid; i < q; i += nt) { w.tmp[i] = w.u[i + 1]; w.itmp[i] = (i + 1 < q) ? w.act[i + 1] : -1; }
    if (tid == 0) w.pos[w.act[l]] = -1;
    blk.sync();
    for (int k = l + tid; k <= q - 2; k += nt) {
        double* col = w.Ui + gi_tri(k);
        for (int j = l; j < k; ++j) col[j] = col[j + 1];
        col[k] = w.sub[k + 1];
    }
    for (int i = l + tid; i < q; i += nt) {
        w.u[i] = w.tmp[i];
        if (i < q - 1) { w.act[i] = w.itmp[i]; w.pos[w.itmp[i]] = i; }
    }
    blk.sync();
    q -= 1;
}

// c = N_W r as a dense vector in extended coordinates (generic form: one thread scatters; the CUDA specialisation
// gathers per coordinate from the rows bucketed by stage)
template <class Blk, class Cons>
FT_HD void gis_scatter(Blk& blk, const Cons& cons, const GisWork& w, int ne, int q) {
    const int tid = blk.tid(), nt = blk.nthreads();
    for (int i = tid; i < ne; i += nt) w.c[i] = 0.0;
    blk.sync();
    if (tid == 0) {
        for (int k = 0; k < q; ++k) {
            SparseRow nk;
            cons.row(w.act[k], nk);
            for (int j = 0; j < nk.nnz; ++j) w.c[nk.idx[j]] += w.r[k] * nk.val[j];
        }
    }
    blk.sync();
}

// r = S^-1 rhs = R^-1 R^-T rhs   (rhs in w.w, result in w.r; w.v = R^-T rhs)
template <class Blk>
FT_HD void gis_schur_solve(Blk& blk, const GisWork& w, int q) {
    const int tid = blk.tid(), nt = blk.nthreads();
    for (int k = tid; k < q; k += nt) {          // v_k = sum_{j <= k} Ui(j,k) w_j   (column k of R^-1: contiguous)
        const double* col = w.Ui + gi_tri(k);
        double a = 0.0;
        for (int j = 0; j <= k; ++j) a += col[j] * w.w[j];
        w.v[k] = a;
    }
    blk.sync();
    for (int j = tid; j < q; j += nt) {          // r_j = sum_{k >= j} Ui(j,k) v_k
        double a = 0.0;
        for (int k = j; k < q; ++k) a += w.Ui[gi_tri(k) + j] * w.v[k];
        w.r[j] = a;
    }
    blk.sync();
}

// On entry: xe = [x ; X x] the unconstrained minimiser, K as above.  On exit: xe the solution, lam[m] multipliers.
// Returns GI_OK / GI_MAXIT / GI_INFEASIBLE; GI_MAXIT is also returned when the working set would exceed qcap.
template <class Blk, class Cons>
FT_HD int gis_solve(Blk& blk, const Cons& cons, const GisWork& w, int ne, int m, double* lam, int maxit, double tol,
                    int* iters_out, int* nact_out) {
    const int tid = blk.tid(), nt = blk.nthreads();
    const double dep_tol = 1e-14, refine_tol = 1e-13;
    int q = 0, iters = 0, status = GI_OK, qmax = 0;
    for (int i = tid; i < m; i += nt) {
        w.s[i] = cons.slack(i, w.xe, 1.0);
        w.pos[i] = -1;
    }
    blk.sync();
    for (;;) {
        if (q > qmax) qmax = q;
        // ---- most violated row
        double best = 0.0;
        int bi = 0x7fffffff;
        for (int i = tid; i < m; i += nt) {
            if (w.pos[i] < 0) {
                const double v = w.s[i];
                if (v < best || (v == best && i < bi)) { best = v; bi = i; }
            }
        }
        blk.argmin(best, bi);
        if (bi == 0x7fffffff || best >= -tol) break;
        const int p = bi;
        double sp = best;
        SparseRow np;
        cons.row(p, np);
        if (tid == 0) w.u[q] = 0.0;
        // ye = K n_p  (the same for every pass of the inner loop: K and n_p are fixed)
        for (int i = tid; i < ne; i += nt) {
            double a = 0.0;
            for (int k = 0; k < np.nnz; ++k) a += np.val[k] * gis_K(w.K, i, np.idx[k]);
            w.ye[i] = a;
        }
        blk.sync();
        double dn = 0.0;
        for (int k = 0; k < np.nnz; ++k) dn += np.val[k] * w.ye[np.idx[k]];
        bool added = false;
        while (!added) {
            ++iters;
            if (iters > maxit) { status = GI_MAXIT; break; }
            // w = N_W' ye,  r = S^-1 w
            for (int k = tid; k < q; k += nt) w.w[k] = cons.slack(w.act[k], w.ye, 0.0);
            blk.sync();
            gis_schur_solve(blk, w, q);
            // c = N_W r (dense, extended coordinates), ze = ye - K c
            gis_scatter(blk, cons, w, ne, q);
            for (int i = tid; i < ne; i += nt) {
                double a = 0.0;
                for (int j = 0; j < ne; ++j) a += gis_K(w.K, i, j) * w.c[j];
                w.ze[i] = w.ye[i] - a;
            }
            blk.sync();
            // refinement on the semi-normal equations: e = N_W' ze vanishes in exact arithmetic
            if (q > 0) {
                double emax = 0.0;
                for (int k = tid; k < q; k += nt) {
                    const double e = cons.slack(w.act[k], w.ze, 0.0);
                    w.w[k] = e;
                    w.tmp[k] = w.r[k];
                    emax = fmax(emax, fabs(e));
                }
                emax = blk.max(emax);
                if (emax > refine_tol * fabs(dn)) {
                    gis_schur_solve(blk, w, q);   // delta r in w.r
                    gis_scatter(blk, cons, w, ne, q);
                    for (int i = tid; i < ne; i += nt) {
                        double a = 0.0;
                        for (int j = 0; j < ne; ++j) a += gis_K(w.K, i, j) * w.c[j];
                        w.ze[i] -= a;
                    }
                    for (int k = tid; k < q; k += nt) w.r[k] += w.tmp[k];
                    blk.sync();
                    blk.count(CT_GI_REFINE);
                }
            }
            double d2n = 0.0;
            for (int k = 0; k < np.nnz; ++k) d2n += np.val[k] * w.ze[np.idx[k]];
            // dual step length
            double t1 = INFINITY;
            int l = 0x7fffffff;
            for (int j = tid; j < q; j += nt) {
                if (w.r[j] > 1e-13) {
                    const double tj = w.u[j] / w.r[j];
                    if (tj < t1 || (tj == t1 && j < l)) { t1 = tj; l = j; }
                }
            }
            blk.argmin(t1, l);
            const bool dep = (q >= w.qcap) || !(d2n > dep_tol * fmax(1.0, dn)) || (d2n <= 1e-28);
            const double t2 = dep ? INFINITY : (-sp / d2n);
            const double t = fmin(t1, t2);
            if (t == INFINITY) { status = (q >= w.qcap) ? GI_MAXIT : GI_INFEASIBLE; break; }
            if (t2 == INFINITY) {
                for (int j = tid; j <= q; j += nt) w.u[j] += t * ((j < q) ? -w.r[j] : 1.0);
                blk.sync();
                gis_drop(blk, w, q, l);
                continue;
            }
            for (int i = tid; i < ne; i += nt) w.xe[i] += t * w.ze[i];
            for (int j = tid; j <= q; j += nt) w.u[j] += t * ((j < q) ? -w.r[j] : 1.0);
            for (int i = tid; i < m; i += nt) w.s[i] += t * cons.slack(i, w.ze, 0.0);
            sp += t * d2n;
            blk.sync();
            if (t == t2) {
                const double rho = sqrt(d2n);
                double* col = w.Ui + gi_tri(q);
                for (int j = tid; j < q; j += nt) col[j] = -w.r[j] / rho;
                if (tid == 0) {
                    col[q] = 1.0 / rho;
                    w.act[q] = p;
                    w.pos[p] = q;
                    w.s[p] = 0.0;                 // on the constraint by construction (keeps the active slacks from drifting)
                }
                blk.sync();
                q += 1;
                added = true;
            } else {
                gis_drop(blk, w, q, l);
            }
        }
        if (status != GI_OK) break;
    }
    for (int i = tid; i < m; i += nt) lam[i] = 0.0;
    blk.sync();
    for (int j = tid; j < q; j += nt) lam[w.act[j]] = w.u[j];
    blk.sync();
    *iters_out = iters;
    *nact_out = q;
#if defined(FTMPC_DEBUG_COUNTERS) && !defined(__CUDACC__)
    if (q > qmax) qmax = q;
    __sync_fetch_and_add(&g_ftmpc_qmax_hist[qmax / 8 < 15 ? qmax / 8 : 15], 1L);
#endif
    return status;
}

// ---- the same iteration with K as an OPERATOR (long horizons: ric_apply, ftmpc_riccati.cuh) ---------------------------------
// KOp::apply(blk, v, out, t_top, has_e) computes out = K v for a dense v in extended coordinates (called by every thread of
// the block; t_top = highest stage with a non-zero input entry, -1 for none; has_e = the terminal part of v is non-zero).
// Work vectors: ye, ze, c as above plus vin (ne).  c = N_W r is GATHERED per coordinate (fixed summation order: results do
// not depend on the thread count or on timing), which needs the structure of the MPC rows (MpcCons).
template <class Blk, class Cons>
FT_HD void gis_gather(Blk& blk, const Cons& cons, const GisWork& w, int ne, int q) {
    const int tid = blk.tid(), nt = blk.nthreads();
    const int n = cons.n, nv = cons.nv, nh = FTMPC_NH * cons.N;
    for (int i = tid; i < ne; i += nt) {
        double a = 0.0;
        if (i < n) {
            const int t = i / FTMPC_NU, j = i - t * FTMPC_NU;
            for (int k = 0; k < q; ++k) {
                const int p = w.act[k];
                if (p < nh && p / FTMPC_NH == t) a -= w.r[k] * cons.Ah[(p - t * FTMPC_NH) * FTMPC_NU + j];
            }
        } else if (i == n) {
            for (int k = 0; k < q; ++k) {
                const int p = w.act[k];
                if (p < cons.mc) { const double c = cons.cv[p]; if (c > 0.0) a += w.r[k] * c; }
                else a += (p == cons.mc) ? w.r[k] : -w.r[k];
            }
        } else {
            const int e = i - nv;
            for (int k = 0; k < q; ++k) {
                const int p = w.act[k];
                if (p >= nh && p < cons.mc) a -= w.r[k] * cons.Af[(p - nh) * FTMPC_NE + e];
            }
        }
        w.c[i] = a;
    }
    blk.sync();
}

// the row table of Yw follows the working set: member l leaves, its row goes to the free end (call BEFORE gis_drop, q = old size)
template <class Blk>
FT_HD void gis_yslot_drop(Blk& blk, int* yslot, bool on, int q, int l) {
    if (!on) return;
    if (blk.tid() == 0) {
        const int freed = yslot[l];
        for (int i = l; i < q - 1; ++i) yslot[i] = yslot[i + 1];
        yslot[q - 1] = freed;
    }
    blk.sync();
}
// Yw (optional, [qcap][ne]) keeps  K n_k  of every member of the working set (it is the `ye` computed when the member was
// added; yslot[k] = its row).  With it  ze = ye - K N_W r = ye - sum_k r_k Yw[k]  is a dense combination of q stored vectors
// instead of a second operator product per iteration (the product is still used by the refinement step).
template <class Blk, class Cons, class KOp>
FT_HD int gis_solve_op(Blk& blk, const Cons& cons, const GisWork& w, KOp& kop, double* vin, int ne, int m, double* lam, int maxit,
                       double tol, int* iters_out, int* nact_out, double* Yw = nullptr, int* yslot = nullptr, int ycap = 0) {
    const int tid = blk.tid(), nt = blk.nthreads();
    // ycap rows of Yw: once the working set outgrows them the rest of this QP goes back to two products per iteration
    if (Yw) {
        for (int k = tid; k < ycap; k += nt) yslot[k] = k;
        blk.sync();
    }
    const double dep_tol = 1e-14, refine_tol = 1e-13;
    const int nh = FTMPC_NH * cons.N;
    int q = 0, iters = 0, status = GI_OK;
    for (int i = tid; i < m; i += nt) {
        w.s[i] = cons.slack(i, w.xe, 1.0);
        w.pos[i] = -1;
    }
    blk.sync();
    for (;;) {
        double best = 0.0;
        int bi = 0x7fffffff;
        for (int i = tid; i < m; i += nt) {
            if (w.pos[i] < 0) {
                const double v = w.s[i];
                if (v < best || (v == best && i < bi)) { best = v; bi = i; }
            }
        }
        blk.argmin(best, bi);
        blk.mark(PH_GI_SELECT);
        if (bi == 0x7fffffff || best >= -tol) break;
        const int p = bi;
        double sp = best;
        SparseRow np;
        cons.row(p, np);
        if (tid == 0) w.u[q] = 0.0;
        // ye = K n_p
        for (int i = tid; i < ne; i += nt) vin[i] = 0.0;
        blk.sync();
        if (tid == 0) for (int k = 0; k < np.nnz; ++k) vin[np.idx[k]] = np.val[k];
        blk.sync();
        kop.apply(blk, vin, w.ye, (p < nh) ? p / FTMPC_NH : ((p < cons.mc) ? cons.N - 1 : -1), p >= nh && p < cons.mc);
        blk.mark(PH_GI_D);
        double dn = 0.0;
        for (int k = 0; k < np.nnz; ++k) dn += np.val[k] * w.ye[np.idx[k]];
        bool added = false;
        while (!added) {
            ++iters;
            if (iters > maxit) { status = GI_MAXIT; break; }
            for (int k = tid; k < q; k += nt) w.w[k] = cons.slack(w.act[k], w.ye, 0.0);
            blk.sync();
            if (q > 0) {
                gis_schur_solve(blk, w, q);
                if (Yw) {
                    blk.mark(PH_GI_UPD);
                    for (int i = tid; i < ne; i += nt) {
                        double a0 = 0.0, a1 = 0.0;
                        int k = 0;
                        for (; k + 1 < q; k += 2) {
                            a0 += w.r[k] * Yw[(size_t)yslot[k] * ne + i];
                            a1 += w.r[k + 1] * Yw[(size_t)yslot[k + 1] * ne + i];
                        }
                        if (k < q) a0 += w.r[k] * Yw[(size_t)yslot[k] * ne + i];
                        w.ze[i] = w.ye[i] - (a0 + a1);
                    }
                } else {
                    gis_gather(blk, cons, w, ne, q);
                    blk.mark(PH_GI_UPD);
                    kop.apply(blk, w.c, vin, cons.N - 1, true);
                    for (int i = tid; i < ne; i += nt) w.ze[i] = w.ye[i] - vin[i];
                }
                blk.sync();
                blk.mark(PH_GI_Z);
                double emax = 0.0;
                for (int k = tid; k < q; k += nt) {
                    const double e = cons.slack(w.act[k], w.ze, 0.0);
                    w.w[k] = e;
                    w.tmp[k] = w.r[k];
                    emax = fmax(emax, fabs(e));
                }
                emax = blk.max(emax);
                if (emax > refine_tol * fabs(dn)) {
                    gis_schur_solve(blk, w, q);
                    gis_gather(blk, cons, w, ne, q);
                    kop.apply(blk, w.c, vin, cons.N - 1, true);
                    for (int i = tid; i < ne; i += nt) w.ze[i] -= vin[i];
                    for (int k = tid; k < q; k += nt) w.r[k] += w.tmp[k];
                    blk.sync();
                    blk.count(CT_GI_REFINE);
                }
            } else {
                for (int i = tid; i < ne; i += nt) w.ze[i] = w.ye[i];
                blk.sync();
            }
            double d2n = 0.0;
            for (int k = 0; k < np.nnz; ++k) d2n += np.val[k] * w.ze[np.idx[k]];
            double t1 = INFINITY;
            int l = 0x7fffffff;
            for (int j = tid; j < q; j += nt) {
                if (w.r[j] > 1e-13) {
                    const double tj = w.u[j] / w.r[j];
                    if (tj < t1 || (tj == t1 && j < l)) { t1 = tj; l = j; }
                }
            }
            blk.argmin(t1, l);
            const bool dep = (q >= w.qcap) || !(d2n > dep_tol * fmax(1.0, dn)) || (d2n <= 1e-28);
            const double t2 = dep ? INFINITY : (-sp / d2n);
            const double t = fmin(t1, t2);
            if (t == INFINITY) { status = (q >= w.qcap) ? GI_MAXIT : GI_INFEASIBLE; break; }
            if (t2 == INFINITY) {
                for (int j = tid; j <= q; j += nt) w.u[j] += t * ((j < q) ? -w.r[j] : 1.0);
                blk.sync();
                gis_yslot_drop(blk, yslot, Yw != nullptr, q, l);
                gis_drop(blk, w, q, l);
                blk.count(CT_GI_DROP);
                continue;
            }
            for (int i = tid; i < ne; i += nt) w.xe[i] += t * w.ze[i];
            for (int j = tid; j <= q; j += nt) w.u[j] += t * ((j < q) ? -w.r[j] : 1.0);
            for (int i = tid; i < m; i += nt) w.s[i] += t * cons.slack(i, w.ze, 0.0);
            sp += t * d2n;
            blk.sync();
            blk.mark(PH_GI_STEP);
            if (t == t2) {
                const double rho = sqrt(d2n);
                double* col = w.Ui + gi_tri(q);
                for (int j = tid; j < q; j += nt) col[j] = -w.r[j] / rho;
                if (tid == 0) {
                    col[q] = 1.0 / rho;
                    w.act[q] = p;
                    w.pos[p] = q;
                    w.s[p] = 0.0;
                }
                if (Yw) {
                    if (q < ycap) {
                        double* yrow = Yw + (size_t)yslot[q] * ne;
                        for (int i = tid; i < ne; i += nt) yrow[i] = w.ye[i];
                    } else {
                        Yw = nullptr;                             // (uniform over the block)
                    }
                }
                blk.sync();
                q += 1;
                added = true;
            } else {
                gis_yslot_drop(blk, yslot, Yw != nullptr, q, l);
                gis_drop(blk, w, q, l);
                blk.count(CT_GI_DROP);
            }
        }
        if (status != GI_OK) break;
    }
    for (int i = tid; i < m; i += nt) lam[i] = 0.0;
    blk.sync();
    for (int j = tid; j < q; j += nt) lam[w.act[j]] = w.u[j];
    blk.sync();
    blk.count(CT_GI_ITER, iters);
    *iters_out = iters;
    *nact_out = q;
    return status;
}

}  // namespace ftmpc
