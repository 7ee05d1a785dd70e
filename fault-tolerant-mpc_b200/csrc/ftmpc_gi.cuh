// ftmpc_gi.cuh -- dense dual active-set QP (Goldfarb-Idnani) written for one cooperative block.
//
//     min 1/2 x'Gx + a'x    s.t.   n_i' x >= beta_i,  i = 0..m-1   (first `meq` rows are equalities)
//
// Replaces the QP machinery the reference reaches through third-party solvers:
//   IPOPT/MUMPS inside  self.solver(**args)      ft_mpc/controllers/spiraling_mpc.py:346
//   OSQP via CVXPY in   ControlAllocator          ft_mpc/controllers/tools/control_allocator.py:28-40,86
//
// GPU-specific formulation (nothing here needs a triangular solve, the latency killer at n~120):
//   * E  (ne x nv, row-major, odd ld) holds J = L^-T Q in its first nv rows and  X*J  in its last
//     ne-nv rows, where X maps the primal to "extended" coordinates in which every constraint normal
//     is SPARSE (for the MPC QP: X = d x_N / d U, so the 72 terminal rows have <= 10 non-zeros).
//   * Ui = R^-1 (packed upper triangular, by columns) is maintained instead of R:
//       add    : new column [-r/rho ; 1/rho]      (r = R^-1 d1 is needed by the step anyway)
//       drop l : the Givens coefficients come from prefix norms of row l of R^-1 (no recurrence
//                across threads), applied to columns of E and of R^-1 row-parallel.
//   * the "add" update of J is ONE Householder reflector:  E2 -= (2/v'v)(E2 v)v',  E2 v = z + s*alpha*E[:,q].
// Every inner iteration is therefore a handful of mat-vecs over shared memory.
#pragma once
#include "ftmpc_block.cuh"

namespace ftmpc {

#define FTMPC_GI_MAXNNZ 12
struct SparseRow {
    int nnz;
    int idx[FTMPC_GI_MAXNNZ];
    double val[FTMPC_GI_MAXNNZ];
    double beta;
};

enum { GI_OK = 0, GI_MAXIT = 1, GI_INFEASIBLE = 2 };
#if defined(FTMPC_DEBUG_COUNTERS) && !defined(__CUDACC__)
extern thread_local int g_ftmpc_warm_hit;
#endif

// n_p . v - beta_p through the generic row() interface (constraint types with structure override `slack`)
template <class Cons>
FT_HD double cons_slack_generic(const Cons& cons, int p, const double* v, double scale_beta) {
    SparseRow row;
    cons.row(p, row);
    double a = -scale_beta * row.beta;
    for (int k = 0; k < row.nnz; ++k) a += row.val[k] * v[row.idx[k]];
    return a;
}

struct GiWork {
    double* E;      // ne x ld
    double* Ui;     // packed upper triangular, capacity nv*(nv+1)/2
    double* xe;     // ne   extended primal  [x ; X x]
    double* s;      // m    slack  n_i' x - beta_i
    double* u;      // nv+1 multipliers of the working set (+ candidate)
    double* d;      // nv
    double* ze;     // ne
    double* r;      // nv
    double* cs;     // 2*nv rotation coefficients
    double* tmp;    // nv+1
    double* sub;    // nv   sub-diagonal fill during a drop
    int* act;       // nv+1 working set (constraint ids)
    int* pos;       // m    position in working set or -1
    int* itmp;      // nv+1
    double* esign;  // meq  orientation used for each equality row
};

FT_HD int gi_tri(int j) { return (j * (j + 1)) >> 1; }

// drop the l-th member of the working set
template <class Blk>
FT_HD void gi_drop(Blk& blk, const GiWork& w, int nv, int ne, int ld, int& q, int l) {
    const int tid = blk.tid(), nt = blk.nthreads();
    // rotation coefficients from prefix norms of row l of R^-1
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
    // rotate columns l..q-1 of E (all rows) and of R^-1 (rows != l)
    for (int row = tid; row < ne + q; row += nt) {
        if (row < ne) {
            double* e = w.E + (size_t)row * ld;
            double carry = e[l];
            for (int k = l; k <= q - 2; ++k) {
                const double c = w.cs[2 * k], s = w.cs[2 * k + 1], b = e[k + 1];
                e[k] = c * carry - s * b;
                carry = s * carry + c * b;
            }
            e[q - 1] = carry;
        } else {
            const int j = row - ne;
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
    }
    // save shifted multipliers / ids
    for (int i = l + tid; i < q; i += nt) { w.tmp[i] = w.u[i + 1]; w.itmp[i] = (i + 1 < q) ? w.act[i + 1] : -1; }
    if (tid == 0) w.pos[w.act[l]] = -1;
    blk.sync();
    // shift rows l+1.. of R^-1 up by one (thread per column) and commit the shifted lists
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

// ---- warm start of the working set ---------------------------------------------------------------------
// Consecutive SQP iterations mostly share their active set, yet the dual method has to re-add it one constraint
// (one O(ne nv) sweep over E plus bookkeeping) at a time.  gi_warm_start puts a PREDICTED set W0 (rows with a
// positive multiplier in the previous QP) into the working set in one go:
//   D0 = J' N0  ->  Householder QR (the same reflectors, in the same order, a one-by-one add would build)
//   E <- E Q,  R^-1,  x = x_unc - J1 R^-T s0,  u = -R^-1 R^-T s0     (s0 = slacks of W0 at x_unc)
// If every u_k >= 0 the pair (x, W0) is exactly the S-pair the dual algorithm would have reached after adding W0 with
// full steps, and gi_solve continues from there.  Otherwise (wrong prediction, dependent rows) the start is
// abandoned: x <- x_unc, empty working set -- E Q is still a valid J, so nothing has to be undone.
// Returns the size of the working set (0 = cold start).  xe_save: scratch of ne doubles.
#define FTMPC_GI_WARM_MAX 40
#define FTMPC_GI_WARM_OFF 2048        /* D0 lives behind the first 2048 entries of the packed R^-1 (q <= 63) */
template <class Blk, class Cons>
FT_HD int gi_warm_start(Blk& blk, const Cons& cons, const GiWork& w, int nv, int ne, int ld, int m, int meq,
                        const double* lam_warm, int m_warm, double* xe_save) {
    const int tid = blk.tid(), nt = blk.nthreads();
    if (!lam_warm || meq != 0) return 0;
    const int cap = (nv * (nv + 1) / 2 + 1 - FTMPC_GI_WARM_OFF) / nv;
    if (cap < 1) return 0;
    const int qmax = cap < FTMPC_GI_WARM_MAX ? cap : FTMPC_GI_WARM_MAX;
    double* Dm = w.Ui + FTMPC_GI_WARM_OFF;       // [q0][nv]
    double* vk0 = w.tmp;                         // v_k[0]
    double* fk = w.sub;                          // 2 / v_k'v_k
    double* rho = w.r;                           // R[k][k]
    // predicted rows, in index order
    for (int i = tid; i < m_warm; i += nt) if (lam_warm[i] > 0.0) w.pos[i] = -2;
    blk.sync();
    if (tid == 0) {
        int c = 0;
        for (int i = 0; i < m_warm; ++i)
            if (w.pos[i] == -2) {
                w.pos[i] = -1;
                if (c <= qmax) w.act[c < qmax ? c : qmax] = i;      // (the count is what matters beyond qmax)
                ++c;
            }
        w.itmp[0] = c;
    }
    blk.sync();
    const int q0 = w.itmp[0];
    if (q0 == 0 || q0 > qmax) return 0;
    for (int i = tid; i < ne; i += nt) xe_save[i] = w.xe[i];
    // D0 = J' N0
    for (int k = 0; k < q0; ++k) {
        SparseRow np;
        cons.row(w.act[k], np);
        for (int i = tid; i < nv; i += nt) {
            double v = 0.0;
            for (int j = 0; j < np.nnz; ++j) v += np.val[j] * w.E[(size_t)np.idx[j] * ld + i];
            Dm[(size_t)k * nv + i] = v;
        }
    }
    blk.sync();
    // Householder QR of D0, reflector k built from column k, rows k..nv-1
    bool ok = true;
    for (int k = 0; k < q0 && ok; ++k) {
        const double* dk = Dm + (size_t)k * nv;
        double p2 = 0.0, pa = 0.0;
        for (int i = tid; i < nv; i += nt) {
            const double v = dk[i] * dk[i];
            pa += v;
            if (i >= k) p2 += v;
        }
        const double d2n = blk.sum(p2);
        const double dn = blk.sum(pa);
        if (d2n <= 1e-22 * fmax(1.0, dn) || d2n <= 1e-28) { ok = false; break; }
        const double alpha = sqrt(d2n), d0 = dk[k];
        const double sg = (d0 >= 0.0) ? 1.0 : -1.0;
        const double v0 = d0 + sg * alpha, f = 2.0 / (2.0 * alpha * (alpha + fabs(d0)));
        blk.sync();                               // everybody has read dk[k]
        if (tid == 0) { vk0[k] = v0; fk[k] = f; rho[k] = -sg * alpha; }
        for (int c = k + 1 + tid; c < q0; c += nt) {          // remaining columns, one thread each
            double* dc = Dm + (size_t)c * nv;
            double a = v0 * dc[k];
            for (int i = k + 1; i < nv; ++i) a += dk[i] * dc[i];
            const double wv = f * a;
            dc[k] -= wv * v0;
            for (int i = k + 1; i < nv; ++i) dc[i] -= wv * dk[i];
        }
        blk.sync();
    }
    if (!ok) return 0;                            // dependent rows: cold start (E untouched so far)
    // R^-1 (packed upper, by columns): R[j][c] = Dm[c][j] (j < c), R[c][c] = rho[c]
    for (int c = tid; c < q0; c += nt) {
        double* col = w.Ui + gi_tri(c);
        col[c] = 1.0 / rho[c];
        for (int j = c - 1; j >= 0; --j) {
            double a = 0.0;
            for (int l = j + 1; l <= c; ++l) a += Dm[(size_t)l * nv + j] * col[l];
            col[j] = -a / rho[j];
        }
    }
    // y = R^-T s0 (serial, q0^2 / 2 operations)
    if (tid == 0) {
        for (int k = 0; k < q0; ++k) {
            double a = w.s[w.act[k]];
            for (int j = 0; j < k; ++j) a -= Dm[(size_t)k * nv + j] * w.d[j];
            w.d[k] = a / rho[k];
        }
    }
    blk.sync();
    // u = -R^-1 y
    double umin = 0.0, umax = 0.0;
    for (int j = tid; j < q0; j += nt) {
        double a = 0.0;
        for (int k = j; k < q0; ++k) a += w.Ui[gi_tri(k) + j] * w.d[k];
        w.u[j] = -a;
        umin = fmin(umin, -a);
        umax = fmax(umax, fabs(a));
    }
    umin = -blk.max(-umin);
    umax = blk.max(umax);
    if (umin < -1e-9 * fmax(1.0, umax)) return 0; // a predicted row wants a negative multiplier: cold start (E untouched)
    // E <- E Q  (rows in extended coordinates), then x = x_unc - E[:, :q0] y
    for (int row = tid; row < ne; row += nt) {
        double* e = w.E + (size_t)row * ld;
        for (int k = 0; k < q0; ++k) {
            const double* dk = Dm + (size_t)k * nv;
            double a = vk0[k] * e[k];
            for (int i = k + 1; i < nv; ++i) a += dk[i] * e[i];
            const double wv = fk[k] * a;
            e[k] -= wv * vk0[k];
            for (int i = k + 1; i < nv; ++i) e[i] -= wv * dk[i];
        }
        double a = 0.0;
        for (int k = 0; k < q0; ++k) a += e[k] * w.d[k];
        w.xe[row] -= a;
    }
    for (int j = tid; j < q0; j += nt) {
        if (w.u[j] < 0.0) w.u[j] = 0.0;
        w.pos[w.act[j]] = j;
    }
    blk.sync();
    for (int i = tid; i < m; i += nt) w.s[i] = cons.slack(i, w.xe, 1.0);
    blk.sync();
    return q0;
}

// On entry: E = [J ; X J] with J J' = G^-1, xe = [x ; X x] the unconstrained minimiser.
// On exit : xe the solution, lam[m] multipliers (>=0 for inequalities), returns GI_* status.
template <class Blk, class Cons>
FT_HD int gi_solve(Blk& blk, const Cons& cons, const GiWork& w, int nv, int ne, int ld, int m, int meq,
                   double* lam, int maxit, double tol, int* iters_out, int* nact_out,
                   const double* lam_warm = nullptr, int m_warm = 0) {
    const int tid = blk.tid(), nt = blk.nthreads();
    int q = 0, iters = 0, status = GI_OK;
    // slack of every constraint at the unconstrained minimiser
    for (int i = tid; i < m; i += nt) {
        w.s[i] = cons.slack(i, w.xe, 1.0);
        w.pos[i] = -1;
    }
    blk.sync();
    q = gi_warm_start(blk, cons, w, nv, ne, ld, m, meq, lam_warm, m_warm, w.ze);
#if defined(FTMPC_DEBUG_COUNTERS) && !defined(__CUDACC__)
    g_ftmpc_warm_hit = q > 0;
#endif
    int eq_next = 0;
    for (;;) {
        // ---- choose the constraint to add: pending equalities first, then the most violated row
        int p = -1;
        double sp = 0.0;
        bool is_eq = false;
        if (eq_next < meq) {
            p = eq_next++;
            sp = w.s[p];
            is_eq = true;
        } else {
            double best = 0.0;
            int bi = 0x7fffffff;
            for (int i = meq + tid; i < m; i += nt) {
                if (w.pos[i] < 0) {
                    const double v = w.s[i];
                    if (v < best || (v == best && i < bi)) { best = v; bi = i; }
                }
            }
            blk.argmin(best, bi);
            if (bi == 0x7fffffff || best >= -tol) break;      // primal feasible -> optimal
            p = bi;
            sp = best;
        }
        SparseRow np;
        cons.row(p, np);
        if (is_eq) {
            const bool rev = sp > 0.0;      // equality violated from above: use the reversed normal
            if (rev) {
                for (int k = 0; k < np.nnz; ++k) np.val[k] = -np.val[k];
                sp = -sp;
            }
            if (tid == 0) w.esign[p] = rev ? -1.0 : 1.0;
        }
        if (tid == 0) w.u[q] = 0.0;
        bool added = false;
        while (!added) {
            ++iters;
            if (iters > maxit) { status = GI_MAXIT; break; }
            // d = J' n_p  (sparse combination of rows of E)
            for (int i = tid; i < nv; i += nt) {
                double v = 0.0;
                for (int k = 0; k < np.nnz; ++k) v += np.val[k] * w.E[(size_t)np.idx[k] * ld + i];
                w.d[i] = v;
            }
            blk.sync();
            // ze = E[:, q:] d[q:] ;  r = R^-1 d[:q]
            for (int row = tid; row < ne + q; row += nt) {
                if (row < ne) {
                    const double* e = w.E + (size_t)row * ld;
                    double v = 0.0;
                    for (int k = q; k < nv; ++k) v += e[k] * w.d[k];
                    w.ze[row] = v;
                } else {
                    const int j = row - ne;
                    double v = 0.0;
                    for (int k = j; k < q; ++k) v += w.Ui[gi_tri(k) + j] * w.d[k];
                    w.r[j] = v;
                }
            }
            // |d2|^2, |d|^2
            double p2 = 0.0, pa = 0.0;
            for (int k = tid; k < nv; k += nt) {
                const double v = w.d[k] * w.d[k];
                pa += v;
                if (k >= q) p2 += v;
            }
            blk.sync();      // ze, r visible
            const double d2n = blk.sum(p2);
            const double dn = blk.sum(pa);
            // dual step length t1 (inequalities in the working set only)
            double t1 = INFINITY;
            int l = 0x7fffffff;
            for (int j = tid; j < q; j += nt) {
                if (w.act[j] >= meq && w.r[j] > 1e-13) {
                    const double tj = w.u[j] / w.r[j];
                    if (tj < t1 || (tj == t1 && j < l)) { t1 = tj; l = j; }
                }
            }
            blk.argmin(t1, l);
            const bool dep = (q >= nv) || (d2n <= 1e-22 * fmax(1.0, dn)) || (d2n <= 1e-28);
            const double t2 = dep ? INFINITY : (-sp / d2n);
            const double t = fmin(t1, t2);
            if (t == INFINITY) {
                if (is_eq && fabs(sp) <= 1e-9) break;      // redundant (dependent, consistent) equality: skip it
                status = GI_INFEASIBLE;
                break;
            }
            if (t2 == INFINITY) {
                // step in dual space only, then drop l
                for (int j = tid; j <= q; j += nt) w.u[j] += t * ((j < q) ? -w.r[j] : 1.0);
                blk.sync();
                gi_drop(blk, w, nv, ne, ld, q, l);
                continue;
            }
            // primal + dual step
            for (int i = tid; i < ne; i += nt) w.xe[i] += t * w.ze[i];
            for (int j = tid; j <= q; j += nt) w.u[j] += t * ((j < q) ? -w.r[j] : 1.0);
            for (int i = tid; i < m; i += nt) w.s[i] += t * cons.slack(i, w.ze, 0.0);
            sp += t * d2n;
            blk.sync();
            if (t == t2) {
                // full step: add p.  Householder on d[q:]
                const double alpha = sqrt(d2n);
                const double d0 = w.d[q];
                const double sg = (d0 >= 0.0) ? 1.0 : -1.0;
                const double vv = 2.0 * alpha * (alpha + fabs(d0));
                const double rho = -sg * alpha;
                if (vv > 0.0) {
                    const double f = 2.0 / vv;
                    for (int row = tid; row < ne; row += nt) {
                        double* e = w.E + (size_t)row * ld;
                        const double wv = f * (w.ze[row] + sg * alpha * e[q]);
                        e[q] -= wv * (d0 + sg * alpha);
                        for (int k = q + 1; k < nv; ++k) e[k] -= wv * w.d[k];
                    }
                }
                double* col = w.Ui + gi_tri(q);
                for (int j = tid; j < q; j += nt) col[j] = -w.r[j] / rho;
                if (tid == 0) {
                    col[q] = 1.0 / rho;
                    w.act[q] = p;
                    w.pos[p] = q;
                }
                blk.sync();
                q += 1;
                added = true;
            } else {
                gi_drop(blk, w, nv, ne, ld, q, l);
            }
        }
        if (status != GI_OK) break;
    }
    // multipliers
    for (int i = tid; i < m; i += nt) lam[i] = 0.0;
    blk.sync();
    for (int j = tid; j < q; j += nt) lam[w.act[j]] = (w.act[j] < meq) ? w.esign[w.act[j]] * w.u[j] : w.u[j];
    blk.sync();
    *iters_out = iters;
    *nact_out = q;
    return status;
}

#if defined(__CUDACC__)
// Register-blocked row kernels.  A CTA runs 8 warps (2 per scheduler), so a dependent load -> FMA -> store chain
// is bound by latency; these load eight operands of each stream before touching them.
// Row kernels for a LANE PAIR per row: lane `part` of the pair owns the columns of its parity (k, k + 2, ...).
// With rows 2 apart inside a half-warp (see gi_pair_row) the 16 lanes of a shared-memory phase hit 16 distinct 8-byte
// banks for any odd row stride, so the pair split costs no extra wavefronts, every thread of the block works, and the
// dependent chain per thread is half as long.
__device__ __forceinline__ double dot_stride2(const double* __restrict__ a, const double* __restrict__ b, int k0, int k1) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int k = k0;
    for (; k + 14 < k1; k += 16) {
        double x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { x[j] = a[k + 2 * j]; y[j] = b[k + 2 * j]; }
        s0 += x[0] * y[0]; s1 += x[1] * y[1]; s2 += x[2] * y[2]; s3 += x[3] * y[3];
        s0 += x[4] * y[4]; s1 += x[5] * y[5]; s2 += x[6] * y[6]; s3 += x[7] * y[7];
    }
    for (; k + 6 < k1; k += 8) {
        double x[4], y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { x[j] = a[k + 2 * j]; y[j] = b[k + 2 * j]; }
        s0 += x[0] * y[0]; s1 += x[1] * y[1]; s2 += x[2] * y[2]; s3 += x[3] * y[3];
    }
    for (; k < k1; k += 2) s0 += a[k] * b[k];
    return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ void axpy_stride2(double* __restrict__ e, const double* __restrict__ d, double wv, int k0, int k1) {
    int k = k0;
    for (; k + 14 < k1; k += 16) {
        double x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { x[j] = e[k + 2 * j]; y[j] = d[k + 2 * j]; }
#pragma unroll
        for (int j = 0; j < 8; ++j) e[k + 2 * j] = x[j] - wv * y[j];
    }
    for (; k + 6 < k1; k += 8) {
        double x[4], y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { x[j] = e[k + 2 * j]; y[j] = d[k + 2 * j]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) e[k + 2 * j] = x[j] - wv * y[j];
    }
    for (; k < k1; k += 2) e[k] -= wv * d[k];
}
// row of a lane pair inside a group of nt / 2 rows: half-warp hw holds the rows  16 (hw / 2) + (hw & 1) + 2 j,  j = 0..7
__device__ __forceinline__ int gi_pair_row(int tid) {
    const int hw = tid >> 4, j = (tid & 15) >> 1;
    return ((hw >> 1) << 4) + (hw & 1) + (j << 1);
}
#endif

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------------------
// CUDA-block specialisation of gi_solve: identical pivoting rules, five barriers per "add" iteration.
//   * |d|^2 and |d2|^2 are reduced with warp shuffles while d is produced (no extra block reduction);
//   * z = E[:, q:] d[q:] runs thread-per-row with four independent accumulators (the dependent-FMA chain of the
//     straightforward loop is what bounds a 121-long dot product), r = R^-1 d[:q] runs on the threads at the other
//     end of the block at the same time, and the dual step length falls out of the same barrier (argmin);
//   * constraint slacks use Cons::slack (no sparse-row materialisation in the sweeps over all m rows).
// ---------------------------------------------------------------------------------------------------------
// CUDA-block specialisation of gi_warm_start (same mathematics; the order of W0 follows the thread mapping instead
// of the row index, which only changes rounding):
//   * ordered compaction by a block prefix sum, D0 one warp per predicted row;
//   * Householder QR with one warp per remaining column (every warp recomputes the pivot column's norm, so a step
//     needs a single barrier);
//   * E <- E Q with HALF A ROW OF E IN REGISTERS per thread: a lane pair loads its row once, applies all q0
//     reflectors (partial dot, one shuffle, update) and stores it once -- two sweeps over E instead of three per
//     added constraint.
#define FTMPC_GI_HALF 61
#define FTMPC_GI_QUARTER 31
__device__ __forceinline__ double lds_f64(unsigned addr) {       // shared-memory load from a 32-bit shared address
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __noinline__ void gi_warm_apply_rows(const double* Dm, const double* vk0, const double* fk, const double* y,
                                                double* E, double* xe, int nv, int ne, int ld, int q0, int tid, int nt) {
    // A QUARTER row (31 columns) per thread, rows in rounds of nt/4: half rows (61 registers + 61 operands in flight), or
    // two quarter rows, do not fit the register file.  The function is out of line, so the operand pointer is generic;
    // when it points to shared memory the loads are issued as ld.shared explicitly.
    const int quarter = tid & 3;
    const int c0 = quarter * FTMPC_GI_QUARTER;
    const int rows_per_round = nt >> 2, rows_cov = (nt >> 1) < ne ? (nt >> 1) : ne;
    const bool in_smem = __isShared(Dm);
    const unsigned dm_s = in_smem ? (unsigned)__cvta_generic_to_shared(Dm) : 0u;
    for (int r0 = 0; r0 < rows_cov; r0 += rows_per_round) {
        const int row = r0 + (tid >> 2);
        const bool live = row < rows_cov;
        double e[FTMPC_GI_QUARTER];
        const double* er = E + (size_t)(live ? row : 0) * ld;
#pragma unroll
        for (int c = 0; c < FTMPC_GI_QUARTER; ++c) e[c] = (live && c0 + c < nv) ? er[c0 + c] : 0.0;
        for (int k = 0; k < q0; ++k) {
            const double v0 = vk0[k], f = fk[k];
            double vv[FTMPC_GI_QUARTER];
            double a4[4] = {0.0, 0.0, 0.0, 0.0};          // four partial sums: the 31-long dependent FMA chain was the bottleneck
#pragma unroll
            for (int c = 0; c < FTMPC_GI_QUARTER; ++c) {
                const int col = c0 + c;
                double v = 0.0;
                if (col > k && col < nv) v = in_smem ? lds_f64(dm_s + (unsigned)((k * nv + col) * 8)) : Dm[(size_t)k * nv + col];
                else if (col == k) v = v0;
                vv[c] = v;
                a4[c & 3] += v * e[c];
            }
            double a = (a4[0] + a4[1]) + (a4[2] + a4[3]);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            const double wv = f * a;
#pragma unroll
            for (int c = 0; c < FTMPC_GI_QUARTER; ++c) e[c] -= wv * vv[c];
        }
        if (live) {
            double* ew = E + (size_t)row * ld;
#pragma unroll
            for (int c = 0; c < FTMPC_GI_QUARTER; ++c) if (c0 + c < nv) ew[c0 + c] = e[c];
        }
        // x = x_unc - E[:, :q0] y : the first q0 <= 40 columns sit in quarters 0 and 1
        double a = 0.0;
#pragma unroll
        for (int c = 0; c < FTMPC_GI_QUARTER; ++c) if (c0 + c < q0) a += e[c] * y[c0 + c];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (live && quarter == 0) xe[row] -= a;
    }
}

template <class Cons>
__device__ __forceinline__ int gi_warm_start(CudaBlock& blk, const Cons& cons, const GiWork& w, int nv, int ne, int ld, int m,
                                             int meq, const double* lam_warm, int m_warm, double* xe_save) {
    const int tid = blk.tid(), nt = blk.nthreads(), lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    if (!lam_warm || meq != 0) return 0;
    if (nv > 4 * FTMPC_GI_QUARTER || ne > (nt >> 1) + nw)                  // shapes the register path does not cover
        return gi_warm_start<CudaBlock, Cons>(blk, cons, w, nv, ne, ld, m, meq, lam_warm, m_warm, xe_save);
    const int cap = (nv * (nv + 1) / 2 + 1 - FTMPC_GI_WARM_OFF) / nv;
    if (cap < 1) return 0;
    const int qmax = cap < FTMPC_GI_WARM_MAX ? cap : FTMPC_GI_WARM_MAX;
    double* Dm = w.Ui + FTMPC_GI_WARM_OFF;       // [q0][nv]
    double* vk0 = w.tmp;
    double* fk = w.sub;
    double* rho = w.r;
    double* gsc = blk.scratch + 128;
    int* isc = reinterpret_cast<int*>(gsc + 32);
    // ---- ordered compaction of the predicted rows (thread t owns rows t, t + nt, ...)
    int mine = 0;
    for (int i = tid; i < m_warm; i += nt) mine += (lam_warm[i] > 0.0) ? 1 : 0;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) isc[warp] = incl;
    blk.sync();
    int base = 0, q0 = 0;
    for (int i = 0; i < nw; ++i) { if (i < warp) base += isc[i]; q0 += isc[i]; }
    if (q0 == 0 || q0 > qmax) return 0;
    {
        int o = base + incl - mine;
        for (int i = tid; i < m_warm; i += nt) if (lam_warm[i] > 0.0) w.act[o++] = i;
    }
    for (int i = tid; i < ne; i += nt) xe_save[i] = w.xe[i];
    blk.sync();
    // ---- D0 = J' N0, one warp per predicted row
    for (int k = warp; k < q0; k += nw) {
        SparseRow np;
        cons.row(w.act[k], np);
        for (int i = lane; i < nv; i += 32) {
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < FTMPC_GI_MAXNNZ; ++j)
                if (j < np.nnz) v += np.val[j] * w.E[(size_t)np.idx[j] * ld + i];
            Dm[(size_t)k * nv + i] = v;
        }
    }
    blk.sync();
    blk.mark(PH_WS_D0);
    // ---- Householder QR: step k = reflector from column k (rows k..nv-1), applied to columns k+1..q0-1
    for (int k = 0; k < q0; ++k) {
        const double* dk = Dm + (size_t)k * nv;
        double p2 = 0.0, pa = 0.0;
        for (int i = lane; i < nv; i += 32) {
            const double v = dk[i] * dk[i];
            pa += v;
            if (i >= k) p2 += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pa += __shfl_xor_sync(0xffffffffu, pa, o);
            p2 += __shfl_xor_sync(0xffffffffu, p2, o);
        }
        if (p2 <= 1e-22 * fmax(1.0, pa) || p2 <= 1e-28) return 0;          // dependent rows (same verdict in every warp)
        const double alpha = sqrt(p2), d0 = dk[k];
        const double sg = (d0 >= 0.0) ? 1.0 : -1.0;
        const double v0 = d0 + sg * alpha, f = 2.0 / (2.0 * alpha * (alpha + fabs(d0)));
        if (tid == 0) { vk0[k] = v0; fk[k] = f; rho[k] = -sg * alpha; }
        for (int c = k + 1 + warp; c < q0; c += nw) {
            double* dc = Dm + (size_t)c * nv;
            double a = 0.0;
            for (int i = k + lane; i < nv; i += 32) a += ((i == k) ? v0 : dk[i]) * dc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            const double wv = f * a;
            for (int i = k + lane; i < nv; i += 32) dc[i] -= wv * ((i == k) ? v0 : dk[i]);
        }
        blk.sync();
    }
    blk.mark(PH_WS_QR);
    // ---- R^-1 (packed upper, by columns): R[j][c] = Dm[c][j] (j < c), R[c][c] = rho[c]
    for (int c = tid; c < q0; c += nt) {
        double* col = w.Ui + gi_tri(c);
        col[c] = 1.0 / rho[c];
        for (int j = c - 1; j >= 0; --j) {
            double a = 0.0;
            for (int l = j + 1; l <= c; ++l) a += Dm[(size_t)l * nv + j] * col[l];
            col[j] = -a / rho[j];
        }
    }
    blk.sync();
    // ---- y = R^-T s0 = (R^-1)' s0,  u = -R^-1 y
    for (int k = tid; k < q0; k += nt) {
        double a = 0.0;
        for (int j = 0; j <= k; ++j) a += w.Ui[gi_tri(k) + j] * w.s[w.act[j]];
        w.d[k] = a;
    }
    blk.sync();
    int bad = 0;
    {
        double umax = 0.0;
        for (int k = 0; k < q0; ++k) umax = fmax(umax, fabs(w.d[k]));      // |y| bounds the scale; cheap and uniform
        for (int j = tid; j < q0; j += nt) {
            double a = 0.0;
            for (int k = j; k < q0; ++k) a += w.Ui[gi_tri(k) + j] * w.d[k];
            w.u[j] = fmax(-a, 0.0);
            if (-a < -1e-9 * fmax(1.0, umax)) bad = 1;
        }
    }
    const int anybad = blk.any(bad);
    blk.mark(PH_WS_SOLVE);
    if (anybad) return 0;                         // a predicted row wants a negative multiplier: cold start, E untouched
    // ---- E <- E Q and x = x_unc - E[:, :q0] y, half a row per thread (out of line: inside this kernel the 61 row
    //      registers would be spilled to local memory, which is what made a first version slower than the adds it replaces)
    gi_warm_apply_rows(Dm, vk0, fk, w.d, w.E, w.xe, nv, ne, ld, q0, tid, nt);
    // rows beyond the lane-pair range: one warp each, in shared memory
    for (int row = (nt >> 1) + warp; row < ne; row += nw) {
        double* e = w.E + (size_t)row * ld;
        for (int k = 0; k < q0; ++k) {
            const double* dk = Dm + (size_t)k * nv;
            const double v0 = vk0[k];
            double a = 0.0;
            for (int i = k + lane; i < nv; i += 32) a += ((i == k) ? v0 : dk[i]) * e[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            const double wv = fk[k] * a;
            for (int i = k + lane; i < nv; i += 32) e[i] -= wv * ((i == k) ? v0 : dk[i]);
            __syncwarp();
        }
        double a = 0.0;
        for (int k = lane; k < q0; k += 32) a += e[k] * w.d[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) w.xe[row] -= a;
    }
    for (int j = tid; j < q0; j += nt) w.pos[w.act[j]] = j;
    blk.sync();
    for (int i = tid; i < m; i += nt) w.s[i] = cons.slack(i, w.xe, 1.0);
    blk.sync();
    blk.mark(PH_WS_E);
    return q0;
}

template <class Cons>
__device__ __forceinline__ int gi_solve(CudaBlock& blk, const Cons& cons, const GiWork& w, int nv, int ne, int ld, int m,
                                        int meq, double* lam, int maxit, double tol, int* iters_out, int* nact_out,
                                        const double* lam_warm = nullptr, int m_warm = 0) {
    const int tid = blk.tid(), nt = blk.nthreads(), lane = tid & 31, warp = tid >> 5, nw = (nt + 31) >> 5;
    double* gsc = blk.scratch + 128;              // 64 doubles of block scratch reserved for this routine
    // row sweeps over E: a lane pair per row (gi_pair_row), rpp = nt / 2 rows per pass; ne_main rows are covered by full
    // passes, up to 2 nw rows left over after the last full pass are swept one per warp instead of paying a whole pass
    const int rpp = nt >> 1, prow = gi_pair_row(tid), part = tid & 1;
    // (with E in global memory -- long horizons -- a whole warp per row, lanes along the row, was measured 2x slower than
    //  this lane-pair mapping: 16 rows in flight per warp hide the latency, and the pair shares every 32-byte sector)
    const int ne_main = ((nt & 31) != 0) ? 0 : ((ne % rpp <= 2 * nw) ? ne - ne % rpp : ne);
    int q = 0, iters = 0, status = GI_OK;
    for (int i = tid; i < m; i += nt) {
        w.s[i] = cons.slack(i, w.xe, 1.0);
        w.pos[i] = -1;
    }
    blk.sync();
    q = gi_warm_start(blk, cons, w, nv, ne, ld, m, meq, lam_warm, m_warm, w.ze);
    blk.count(q > 0 ? CT_GI_WARM_OK : CT_GI_WARM_MISS);
    int eq_next = 0;
    bool have_next = false;                       // most violated row already known from the previous full step
    double next_best = 0.0;
    int next_bi = 0x7fffffff;
    for (;;) {
        // ---- choose the constraint to add: pending equalities first, then the most violated row
        int p = -1;
        double sp = 0.0;
        bool is_eq = false;
        if (eq_next < meq) {
            p = eq_next++;
            sp = w.s[p];
            is_eq = true;
        } else {
            double best = next_best;
            int bi = next_bi;
            if (!have_next) {
                best = 0.0;
                bi = 0x7fffffff;
                for (int i = meq + tid; i < m; i += nt) {
                    if (w.pos[i] < 0) {
                        const double v = w.s[i];
                        if (v < best || (v == best && i < bi)) { best = v; bi = i; }
                    }
                }
                blk.argmin(best, bi);
            }
            if (bi == 0x7fffffff || best >= -tol) break;      // primal feasible -> optimal
            p = bi;
            sp = best;
        }
        have_next = false;
        blk.mark(PH_GI_SELECT);
        SparseRow np;
        cons.row(p, np);
        if (is_eq) {
            const bool rev = sp > 0.0;
            if (rev) {
#pragma unroll
                for (int k = 0; k < FTMPC_GI_MAXNNZ; ++k) np.val[k] = -np.val[k];
                sp = -sp;
            }
            if (tid == 0) w.esign[p] = rev ? -1.0 : 1.0;
        }
        if (tid == 0) w.u[q] = 0.0;
        bool added = false;
        while (!added) {
            ++iters;
            if (iters > maxit) { status = GI_MAXIT; break; }
            // d = J' n_p, with the two squared norms reduced on the fly
            double pa = 0.0, p2 = 0.0;
            for (int i = tid; i < nv; i += nt) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < FTMPC_GI_MAXNNZ; ++k)
                    if (k < np.nnz) v += np.val[k] * w.E[(size_t)np.idx[k] * ld + i];
                w.d[i] = v;
                pa += v * v;
                if (i >= q) p2 += v * v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                pa += __shfl_xor_sync(0xffffffffu, pa, o);
                p2 += __shfl_xor_sync(0xffffffffu, p2, o);
            }
            if (lane == 0) { gsc[warp] = pa; gsc[32 + warp] = p2; }
            blk.sync();
            blk.mark(PH_GI_D);
            double dn = 0.0, d2n = 0.0;
            for (int i = 0; i < nw; ++i) { dn += gsc[i]; d2n += gsc[32 + i]; }
            // ze = E[:, q:] d[q:]  (thread per row: a lane-pair split was measured slower -- the two halves of a row
            // land on the same banks as their neighbours and double the shared-memory wavefronts)
            // -- lane pair per row, nt / 2 rows per pass; the few rows left over after the last full pass go one per warp
            for (int rb = 0; rb < ne_main; rb += rpp) {
                const int row = rb + prow;
                double sv = 0.0;
                if (row < ne_main) sv = dot_stride2(w.E + (size_t)row * ld, w.d, q + ((q ^ part) & 1), nv);
                sv += __shfl_xor_sync(0xffffffffu, sv, 1);
                if (part == 0 && row < ne_main) w.ze[row] = sv;
            }
            for (int row = ne_main + warp; row < ne; row += nw) {
                const double* e = w.E + (size_t)row * ld;
                double sv = 0.0;
                for (int k = q + lane; k < nv; k += 32) sv += e[k] * w.d[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
                if (lane == 0) w.ze[row] = sv;
            }
            // r = R^-1 d[:q] and the dual step length
            double t1 = INFINITY;
            int l = 0x7fffffff;
            for (int j = nt - 1 - tid; j < q; j += nt) {
                double a0 = 0.0, a1 = 0.0;
                int k = j;
                for (; k + 1 < q; k += 2) {
                    a0 += w.Ui[gi_tri(k) + j] * w.d[k];
                    a1 += w.Ui[gi_tri(k + 1) + j] * w.d[k + 1];
                }
                if (k < q) a0 += w.Ui[gi_tri(k) + j] * w.d[k];
                const double rj = a0 + a1;
                w.r[j] = rj;
                if (w.act[j] >= meq && rj > 1e-13) {
                    const double tj = w.u[j] / rj;
                    if (tj < t1 || (tj == t1 && j < l)) { t1 = tj; l = j; }
                }
            }
            blk.argmin(t1, l);                    // its barrier also publishes ze and r
            blk.mark(PH_GI_Z);
            const bool dep = (q >= nv) || (d2n <= 1e-22 * fmax(1.0, dn)) || (d2n <= 1e-28);
            const double t2 = dep ? INFINITY : (-sp / d2n);
            const double t = fmin(t1, t2);
            if (t == INFINITY) {
                if (is_eq && fabs(sp) <= 1e-9) break;
                status = GI_INFEASIBLE;
                break;
            }
            if (t2 == INFINITY) {
                for (int j = tid; j <= q; j += nt) w.u[j] += t * ((j < q) ? -w.r[j] : 1.0);
                blk.sync();
                gi_drop(blk, w, nv, ne, ld, q, l);
                blk.mark(PH_GI_DROP);
                blk.count(CT_GI_DROP);
                continue;
            }
            const bool full = (t == t2);
            for (int i = tid; i < ne; i += nt) w.xe[i] += t * w.ze[i];
            for (int j = tid; j <= q; j += nt) w.u[j] += t * ((j < q) ? -w.r[j] : 1.0);
            // slack update; on a full step the same sweep finds the most violated remaining row for the next pass
            double nb = 0.0;
            int nbi = 0x7fffffff;
            for (int i = tid; i < m; i += nt) {
                const double v = w.s[i] + t * cons.slack(i, w.ze, 0.0);
                w.s[i] = v;
                if (i >= meq && i != p && w.pos[i] < 0 && (v < nb || (v == nb && i < nbi))) { nb = v; nbi = i; }
            }
            sp += t * d2n;
            // a full step goes straight on to the update: it reads ze, d, r and writes E, R^-1, act/pos[p] -- nothing the
            // sweep above writes (xe, u, s) or reads (pos[i], i != p) -- so the barrier is only needed before a drop
            if (!full) blk.sync();
            blk.mark(PH_GI_STEP);
            if (full) {
                const double alpha = sqrt(d2n);
                const double d0 = w.d[q];
                const double sg = (d0 >= 0.0) ? 1.0 : -1.0;
                const double vv = 2.0 * alpha * (alpha + fabs(d0));
                const double rho = -sg * alpha;
                if (vv > 0.0) {
                    const double f = 2.0 / vv;
                    for (int rb = 0; rb < ne_main; rb += rpp) {
                        const int row = rb + prow;
                        const bool on = row < ne_main;
                        double* e = w.E + (size_t)(on ? row : 0) * ld;
                        const double eq = e[q];
                        const double wv = f * (w.ze[on ? row : 0] + sg * alpha * eq);
                        __syncwarp();                               // both lanes of the pair have read e[q]
                        if (on) {
                            if (((q ^ part) & 1) == 0) e[q] = eq - wv * (d0 + sg * alpha);
                            axpy_stride2(e, w.d, wv, q + 1 + (((q + 1) ^ part) & 1), nv);
                        }
                    }
                    for (int row = ne_main + warp; row < ne; row += nw) {
                        double* e = w.E + (size_t)row * ld;
                        const double eq = e[q];
                        const double wv = f * (w.ze[row] + sg * alpha * eq);
                        __syncwarp();
                        if (lane == 0) e[q] = eq - wv * (d0 + sg * alpha);
                        for (int k = q + 1 + lane; k < nv; k += 32) e[k] -= wv * w.d[k];
                    }
                }
                double* col = w.Ui + gi_tri(q);
                for (int j = nt - 1 - tid; j < q; j += nt) col[j] = -w.r[j] / rho;
                if (tid == nt - 1) {
                    col[q] = 1.0 / rho;
                    w.act[q] = p;
                    w.pos[p] = q;
                }
                blk.argmin(nb, nbi);                               // barrier of the update + next selection in one
                next_best = nb;
                next_bi = nbi;
                have_next = eq_next >= meq;
                blk.mark(PH_GI_UPD);
                q += 1;
                added = true;
            } else {
                gi_drop(blk, w, nv, ne, ld, q, l);
                blk.mark(PH_GI_DROP);
                blk.count(CT_GI_DROP);
            }
        }
        if (status != GI_OK) break;
    }
    for (int i = tid; i < m; i += nt) lam[i] = 0.0;
    blk.sync();
    for (int j = tid; j < q; j += nt) lam[w.act[j]] = (w.act[j] < meq) ? w.esign[w.act[j]] * w.u[j] : w.u[j];
    blk.sync();
    blk.count(CT_GI_ITER, iters);
    *iters_out = iters;
    *nact_out = q;
    return status;
}
#endif  // __CUDACC__

}  // namespace ftmpc
