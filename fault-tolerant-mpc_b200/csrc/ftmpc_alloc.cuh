// ftmpc_alloc.cuh -- output assembly and thrust allocation, one thread per instance.
//
// Replaces (reference):
//   get_control post-processing                ft_mpc/controllers/spiraling_mpc.py:301-307
//   ControlAllocator.clip_generalized_input    ft_mpc/controllers/tools/control_allocator.py:42-63
//   ControlAllocator.get_physical_input        ft_mpc/controllers/tools/control_allocator.py:65-95
//       min |u|^2  s.t.  D u = u_des, 0 <= u <= ub          (problem definition :28-40, CVXPY/OSQP)
// Both small QPs run through the same exact dual active-set solver as the MPC QP (SerialBlock
// instantiation: n = 6 / n = 16 fit in thread-local memory).
#pragma once
#include "ftmpc_sqp.cuh"

namespace ftmpc {

struct DenseCons16 {      // rows: 6 equalities D u = u_des, then u_i >= 0, then -u_i >= -ub_i
    const double* D;
    const double* udes;
    const double* ub;
    FT_HD void row(int p, SparseRow& r) const {
        if (p < FTMPC_NU) {
            int k = 0;
            for (int j = 0; j < FTMPC_NTHR; ++j) {
                const double a = D[p * FTMPC_NTHR + j];
                if (a != 0.0 && k < FTMPC_GI_MAXNNZ) { r.idx[k] = j; r.val[k] = a; ++k; }
            }
            r.nnz = k; r.beta = udes[p];
        } else if (p < FTMPC_NU + FTMPC_NTHR) {
            r.nnz = 1; r.idx[0] = p - FTMPC_NU; r.val[0] = 1.0; r.beta = 0.0;
        } else {
            const int j = p - FTMPC_NU - FTMPC_NTHR;
            r.nnz = 1; r.idx[0] = j; r.val[0] = -1.0; r.beta = -ub[j];
        }
    }
    FT_HD double slack(int p, const double* v, double sb) const { return cons_slack_generic(*this, p, v, sb); }
};

struct HullCons6 {        // -A_h v >= -b_h
    const double* Ah;
    const double* bh;
    FT_HD void row(int p, SparseRow& r) const {
        int k = 0;
        for (int j = 0; j < FTMPC_NU; ++j) {
            const double a = Ah[p * FTMPC_NU + j];
            if (a != 0.0) { r.idx[k] = j; r.val[k] = -a; ++k; }
        }
        r.nnz = k; r.beta = -bh[p];
    }
    FT_HD double slack(int p, const double* v, double sb) const { return cons_slack_generic(*this, p, v, sb); }
};

// generic tiny QP  min 1/2 h |x - x0|^2  (G = h I)  with the rows of `cons`; NV <= 16, M <= 38
template <int NV, int M, class Cons>
FT_HD int tiny_qp(const Cons& cons, double h, const double* x0, int meq, double tol, double* x) {
    constexpr int LD = NV | 1;
    double E[NV * LD], Ui[NV * (NV + 1) / 2 + 1], xe[NV], s[M], u[NV + 2], d[NV], ze[NV], r[NV], cs[2 * NV + 2],
        tmp[NV + 2], sub[NV + 2], esign[M], lam[M];
    int act[NV + 2], pos[M], itmp[NV + 2];
    const double jd = 1.0 / sqrt(h);
    for (int i = 0; i < NV * LD; ++i) E[i] = 0.0;
    for (int i = 0; i < NV; ++i) { E[i * LD + i] = jd; xe[i] = x0[i]; }
    GiWork w{E, Ui, xe, s, u, d, ze, r, cs, tmp, sub, act, pos, itmp, esign};
    SerialBlock blk;
    int it = 0, na = 0;
    const int st = gi_solve(blk, cons, w, NV, NV, LD, M, meq, lam, 50 * (NV + M), tol, &it, &na);
    for (int i = 0; i < NV; ++i) x[i] = xe[i];
    return st;
}

// control allocation; returns 0 ok / nonzero infeasible
FT_HD int allocate_thrust(const ftmpc_config& cfg, const double* udes, const double* ub, double* thrust) {
    DenseCons16 cons{cfg.D, udes, ub};
    double x0[FTMPC_NTHR];
    for (int i = 0; i < FTMPC_NTHR; ++i) x0[i] = 0.0;
    const int st = tiny_qp<FTMPC_NTHR, FTMPC_NU + 2 * FTMPC_NTHR>(cons, 2.0, x0, FTMPC_NU, 1e-9, thrust);
    for (int i = 0; i < FTMPC_NTHR; ++i) thrust[i] = fmin(fmax(thrust[i], 0.0), ub[i]);
    return st;
}

// clip_generalized_input: identity inside the hull (tolerance clip_tol), Euclidean projection otherwise
FT_HD int clip_to_hull(const ftmpc_config& cfg, const double* hull, const double* v, double* out) {
    const double* Ah = hull;
    const double* bh = hull + FTMPC_NH * FTMPC_NU;
    bool inside = true;
    for (int i = 0; i < FTMPC_NH; ++i) {
        double a = -bh[i];
        for (int j = 0; j < FTMPC_NU; ++j) a += Ah[i * FTMPC_NU + j] * v[j];
        if (a > cfg.clip_tol) inside = false;
    }
    if (inside) {
        for (int j = 0; j < FTMPC_NU; ++j) out[j] = v[j];
        return 0;
    }
    HullCons6 cons{Ah, bh};
    return tiny_qp<FTMPC_NU, FTMPC_NH>(cons, 1.0, v, 0, 1e-12, out);
}

// ---- phase_out: write the per-instance results, then clip + allocate ---------------------------------
// (a) phase_out_write -- block-cooperative: decision vector [u | x], u0, active set, status, counters, cost
template <class Blk>
FT_HD void phase_out_write(Blk& blk, const ftmpc_config& cfg, const WsLayout& L, const StepIO& io, int inst, int slot) {
    const double* w = ws_slot(io, L, slot);
    const double* sc = w + L.oSc;
    const int N = L.N, tid = blk.tid(), nt = blk.nthreads();
    const double* U = w + L.oU;
    const double* X = w + L.oX;
    const double* C = w + L.oC;
    if ((int)sc[SC_STATUS] == FTMPC_ST_REDO && io.redo) {
        // handed over to the null-space kernel: nothing is written (the warm-start input must stay intact), the id is queued
        if (tid == 0) {
#if defined(__CUDA_ARCH__)
            const int k = atomicAdd(io.redo, 1);
            io.redo[1 + k] = inst;
#endif
            io.status[inst] = FTMPC_ST_REDO;
        }
        blk.sync();
        return;
    }
    // optimal decision vector, reference layout [u | x]   (spiraling_mpc.py:110-114)
    double* zw = io.z_warm + (size_t)inst * (L.n + (size_t)(N + 1) * FTMPC_NX);
    const double* uref = io.uref ? io.uref + (size_t)inst * io.uref_stride : nullptr;
    if (uref) {
        // the solver iterated on u~ = u + rho_t(q_t) (ftmpc_sqp.cuh, FTMPC_CQ): hand back the reference's variable u
        for (int t = tid; t < N; t += nt) {
            double rho[FTMPC_NU];
            nominal_rot(X + t * FTMPC_NX + 9, uref + t * FTMPC_NU, rho);
            for (int j = 0; j < FTMPC_NU; ++j) {
                const double u = U[t * FTMPC_NU + j] - rho[j];
                zw[t * FTMPC_NU + j] = u;
                if (t == 0) io.u0[(size_t)inst * FTMPC_NU + j] = u;
            }
        }
    } else {
        for (int i = tid; i < L.n; i += nt) zw[i] = U[i];
        for (int j = tid; j < FTMPC_NU; j += nt) io.u0[(size_t)inst * FTMPC_NU + j] = U[j];
    }
    for (int i = tid; i < (N + 1) * FTMPC_NX; i += nt) zw[L.n + i] = X[i];
    // active set: row i active <=> b_i - g_i <= act_tol  <=>  c_i >= -act_tol
    const int nw = (L.mc + 31) / 32;
    for (int wd = tid; wd < nw; wd += nt) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int i = wd * 32 + b;
            if (i < L.mc && C[i] >= -cfg.act_tol) bits |= (1u << b);
        }
        io.active_set[(size_t)inst * nw + wd] = bits;
    }
    if (tid == 0) {
        int status = (int)sc[SC_STATUS];
        if (status == FTMPC_ST_RUNNING) status = FTMPC_ST_MAXITER;
        if (status == FTMPC_ST_REDO) status = FTMPC_ST_QPFAIL;          // (no second pass configured)
        io.status[inst] = status;
        io.iters[2 * inst] = (int)sc[SC_ITER];
        io.iters[2 * inst + 1] = (int)sc[SC_QPIT];
        if (io.cost) io.cost[inst] = sc[SC_F];
    }
    blk.sync();
    blk.mark(PH_OUT);
}

// (b) phase_alloc -- one thread per instance, from the written results: u_res, clip, thrust allocation
FT_HD void phase_alloc(const ftmpc_config& cfg, const WsLayout& L, const StepIO& io, int inst) {
    const int N = L.N;
    int status = io.status[inst];
    if (status == FTMPC_ST_BADINPUT) return;               // rejected inputs (phase_out_invalid wrote a zero command)
    const double* zw = io.z_warm + (size_t)inst * (L.n + (size_t)(N + 1) * FTMPC_NX);
    const double* U = zw;
    const double* X = zw + L.n;
    const double* uref = io.uref ? io.uref + (size_t)inst * io.uref_stride : nullptr;
    const double* hull = io.hull_table + (size_t)io.hull_idx[inst] * FTMPC_HULL_STRIDE;
    // u_res = u*_0 + RotFullInv(q_0) ur_0 + u_comp                     spiraling_mpc.py:301-306
    const double* ff = io.fault_force + (size_t)inst * FTMPC_NTHR;
    double Df[FTMPC_NU], v[FTMPC_NU], vc[FTMPC_NU], ub[FTMPC_NTHR];
    for (int i = 0; i < FTMPC_NU; ++i) {
        double a = 0.0;
        for (int j = 0; j < FTMPC_NTHR; ++j) a += cfg.D[i * FTMPC_NTHR + j] * ff[j];      // sys_model.py:241
        Df[i] = a;
    }
    double Wr[FTMPC_NU];
    stage_wrench(cfg, U, uref, X + 9, Wr);                 // u_0 + u_ref_rot + [f_virt;0]
    for (int i = 0; i < FTMPC_NU; ++i) v[i] = Wr[i];        // = u_res + D f   (u_comp = [f_virt;0] - D f)
    // u_des = clip(u_res + D f) - D f                                  control_allocator.py:79
    int st_clip = clip_to_hull(cfg, hull, v, vc);
    double udes[FTMPC_NU];
    for (int i = 0; i < FTMPC_NU; ++i) udes[i] = vc[i] - Df[i];
    const uint16_t mask = io.fault_mask[inst];
    for (int j = 0; j < FTMPC_NTHR; ++j) ub[j] = ((mask >> j) & 1) ? 0.0 : cfg.max_thrust;  // sys_model.py:240
    double th[FTMPC_NTHR];
    int st_alloc = allocate_thrust(cfg, udes, ub, th);
    // residual check of the allocation (the reference exits on a non-optimal status, :88-93)
    double res = 0.0;
    for (int i = 0; i < FTMPC_NU; ++i) {
        double a = -udes[i];
        for (int j = 0; j < FTMPC_NTHR; ++j) a += cfg.D[i * FTMPC_NTHR + j] * th[j];
        res = fmax(res, fabs(a));
    }
    if ((st_alloc != 0 || st_clip != 0 || !(res <= 1e-6)) && status == FTMPC_ST_OK) io.status[inst] = FTMPC_ST_ALLOC;
    for (int j = 0; j < FTMPC_NTHR; ++j) io.thrust[(size_t)inst * FTMPC_NTHR + j] = th[j];
}

}  // namespace ftmpc
