// ftmpc_sqp.cuh -- the per-instance phases of the MPC solve (SQP on the reference NLP).
//
// The reference builds the NLP once with CasADi and lets IPOPT iterate on it
//   SpiralingController.build_solver   ft_mpc/controllers/spiraling_mpc.py:87-238
//   SpiralingController.solve_mpc      ft_mpc/controllers/spiraling_mpc.py:319-354
// Here the same NLP is solved in its reduced (single-shooting) form by an SQP whose sub-problem is
// the condensed QP the north star names:
//   phase_ls   : step acceptance (l1 merit), forward RK4 rollout, cost/constraint values
//   phase_lin  : RK4 Jacobians; costates + exact stage Hessians of the Lagrangian when the QP will blend them in (qp_start)
//   phase_qp   : Hessian schedule, stage-wise (Riccati) factorisation of the condensed QP (ftmpc_riccati.cuh), dual active-set
//                QP -- null-space form on E = [J ; X J] (ftmpc_gi.cuh) up to N = 20, operator form (ftmpc_gis.cuh) above;
//                the round-1 factorisation (condense + Cholesky + L^-T below) stays in host builds as a cross-check
//   phase_out  : u0, active set, thrust allocation
// One CUDA block per instance on the device (k_solve, ftmpc_kernels.cu), one host thread per instance in the CPU port.
// Decision vector and constraint order follow the reference: z = [u_0..u_{N-1} | x_0..x_N]
// (:110-114), inequalities [hull_0 .. hull_{N-1} | terminal] (:206-214).
#pragma once
#include "ftmpc.h"
#include "ftmpc_block.cuh"
#include "ftmpc_dyn.cuh"
#include "ftmpc_gi.cuh"
#include "ftmpc_gis.cuh"
#if !defined(__CUDACC__)
#include <vector>
#endif
#include "ftmpc_linalg.cuh"
#include "ftmpc_terminal.cuh"

namespace ftmpc {

// ---- per-instance workspace (global memory), offsets in doubles ------------------------------
// Accelerating references (non-zero nominal wrench u_r, spiraling_mpc.py:156-172, 279-286).  The reference NLP applies
//   W_t = u_t + rho_t(q_t) + u_comp,   rho_t(q) = [Rot(q)' u_r,F ; u_r,tau]
// to the dynamics AND to the hull rows, so in the variable u both depend on the attitude q_t.  The solver therefore
// iterates on  u~_t = u_t + rho_t(q_t)  (a bijection for fixed states): dynamics and hull rows are then exactly those of
// the hover problem, and the attitude enters only through the input cost  sum_j R_j (u~_j - rho_j(q_t))^2 .  Per stage
// that cost contributes (record of FTMPC_CQ doubles, written by phase_lin):
//   [0..3]   gq   = d l / d q              = -2 Jrho' r,            r = R o (u~ - rho)
//   [4..15]  cx   = d2 l / d u~_i d q_b     = -2 R_i Jrho[i][b]      (i < 3: only the force rotates)
//   [16..31] cqq  = 2 Jrho' R Jrho                                   (Gauss-Newton part of d2 l / d q2)
//   [32..37] rho
//   [40..55] sx   = -2 sum_i r_i d2 rho_i / d q2   (exact remainder, added to the (q,q) block of W_t)
// and the exact remainder  -2 sum_i r_i d2 rho_i / d q2  goes into the (q,q) block of W_t next to the dynamics' second-
// order terms (blended by theta like them).  u = u~ - rho is restored when results are written (phase_out_write).
#define FTMPC_CQ 56
#define FTMPC_ST_REDO 6          /* internal: k_solve2 hands the instance to the null-space kernel (never returned to the caller) */
struct WsLayout {
    int N, n, nv, mc, m;
    size_t oU, oD, oX, oC, oLam, oMu, oJz, oWz, oGV, oHV, oSc, oCq, stride;
};
enum {      // scalar slots
    SC_F = 0, SC_CSUM, SC_NU, SC_GD, SC_THETA, SC_DMAX, SC_DELTA, SC_LAMMAX, SC_STATUS, SC_ITER, SC_QPIT,
    SC_NACT, SC_CHOLFAIL, SC_QPST, SC_ALPHA, SC_CMAX, SC_SIGMA, SC_HFAIL, SC_DPREV, SC_DREF, SC_FREF, SC_COUNT
};
FT_HD WsLayout ws_layout(int N) {
    WsLayout L;
    L.N = N; L.n = FTMPC_NU * N; L.nv = L.n + 1; L.mc = FTMPC_NH * N + FTMPC_NF; L.m = L.mc + 2;
    size_t o = 0;
    L.oU = o; o += L.n;
    L.oD = o; o += L.nv;
    L.oX = o; o += (size_t)(N + 1) * FTMPC_NX;
    L.oC = o; o += L.mc;
    L.oLam = o; o += 2 * (size_t)L.m;       // multipliers + the QP's output copy
    L.oMu = o; o += (size_t)(N + 1) * FTMPC_NX;
    L.oJz = o; o += (size_t)N * 169;
    L.oWz = o; o += (size_t)N * 169;
    L.oGV = o; o += FTMPC_NE;
    L.oHV = o; o += FTMPC_NE * FTMPC_NE;
    L.oSc = o; o += SC_COUNT;
    L.oCq = o; o += (size_t)N * FTMPC_CQ;      // input-cost coupling records (accelerating references only)
    L.stride = (o + 7) & ~(size_t)7;
    return L;
}

// ---- batch I/O ---------------------------------------------------------------------------------
struct StepIO {
    int batch;
    const double* state;        // [B,13] robot state
    const double* xref;         // [B,N+1,9]
    const double* uref;         // [B,N+1,6] or null
    const uint16_t* fault_mask; // [B]
    const double* fault_force;  // [B,16]
    const int32_t* hull_idx;    // [B]
    const double* hull_table;   // device/host table
    int warm;
    double* z_warm;             // [B, 6N+13(N+1)]
    double* thrust;             // [B,16]
    double* u0;                 // [B,6]
    uint32_t* active_set;       // [B, ceil(mc/32)]
    int32_t* status;            // [B]
    int32_t* iters;             // [B,2]
    double* cost;               // [B]
    const double* tf_val;       // compact terminal-set rows (shared memory, built once per CTA) or nullptr
    const int* tf_idx;
    const ftmpc_config* cfg_g;  // copy of the configuration in global memory (tables indexed per thread); the kernels
                                // also receive it by value as a __grid_constant__ parameter for uniform accesses
    double* ws;                 // workspace: one slot of L.stride doubles per instance (CPU port) or per CTA (k_solve)
    int* redo;                  // k_solve2 only: redo[0] = number of instances handed over to the null-space kernel (their working set
                                // outgrew the shared-memory capacity of the range-space QP), redo[1..] = their ids; else nullptr
    const int* inst_list;       // second pass: instance ids to solve (batch_dev[0] of them) instead of 0 .. batch-1; else nullptr
    const int* batch_dev;
    size_t xref_stride;         // doubles between the reference windows of consecutive instances: (N+1)*9, or 0 when every
    size_t uref_stride;         // instance tracks the same window (closed-loop driver: one table shared by the batch)
};
FT_HD void stepio_default_strides(StepIO& io, int N) {
    io.xref_stride = (size_t)(N + 1) * FTMPC_NE;
    io.uref_stride = (size_t)(N + 1) * FTMPC_NU;
}

FT_HD double* ws_slot(const StepIO& io, const WsLayout& L, int slot) { return io.ws + (size_t)slot * L.stride; }

// per-instance input validation: the hull table is indexed with hull_idx straight from the caller's tensor
FT_HD bool instance_input_ok(const ftmpc_config& cfg, const StepIO& io, int inst) {
    const int h = io.hull_idx[inst];
    return h >= 0 && h < cfg.n_hull_sets;
}
// results of an instance whose inputs were rejected: status FTMPC_ST_BADINPUT, zero command, empty active set
template <class Blk>
FT_HD void phase_out_invalid(Blk& blk, const WsLayout& L, const StepIO& io, int inst) {
    const int tid = blk.tid(), nt = blk.nthreads(), nw = (L.mc + 31) / 32;
    for (int j = tid; j < FTMPC_NU; j += nt) io.u0[(size_t)inst * FTMPC_NU + j] = 0.0;
    for (int j = tid; j < FTMPC_NTHR; j += nt) io.thrust[(size_t)inst * FTMPC_NTHR + j] = 0.0;
    for (int wd = tid; wd < nw; wd += nt) io.active_set[(size_t)inst * nw + wd] = 0u;
    if (tid == 0) {
        io.status[inst] = FTMPC_ST_BADINPUT;
        io.iters[2 * inst] = 0;
        io.iters[2 * inst + 1] = 0;
        if (io.cost) io.cost[inst] = 0.0;
    }
    blk.sync();
}

FT_HD DynConsts dyn_consts(const ftmpc_config& c) {
    DynConsts k;
    k.dt = c.dt; k.mass = c.mass;
    for (int i = 0; i < 3; ++i) { k.Jd[i] = c.inertia[i]; k.r[i] = c.r[i]; k.iJ[i] = 1.0 / c.inertia[i]; }
    k.im = 1.0 / c.mass;
    return k;
}

// total body wrench of stage t:  u_t + [Rot(q_t)^T ur_F ; ur_tau] + [f_virt ; 0]
//   spiraling_mpc.py:156-172 (u_t + u_ref_rot + u_comp) + spiral_model.py:61 (+ D f_fault)  ==> fault-independent
FT_HD void stage_wrench(const ftmpc_config& c, const double* u, const double* ur, const double* q, double* Wr) {
    for (int i = 0; i < 3; ++i) { Wr[i] = u[i] + c.f_virt[i]; Wr[3 + i] = u[3 + i]; }
    if (ur) {
        double R[9], a[3];
        rot_mat(q, R);
        mat3_tmul(R, ur, a);
        for (int i = 0; i < 3; ++i) { Wr[i] += a[i]; Wr[3 + i] += ur[3 + i]; }
    }
}

// rho_t(q) = [Rot(q)' u_r,F ; u_r,tau]
FT_HD void nominal_rot(const double* q, const double* ur, double* rho) {
    double R[9];
    rot_mat(q, R);
    mat3_tmul(R, ur, rho);
    for (int i = 0; i < 3; ++i) rho[3 + i] = ur[3 + i];
}
// record described at FTMPC_CQ
FT_HD void stage_cost_coupling(const ftmpc_config& c, const double* ut, const double* ur, const double* q, double* rec) {
    double* Sx = rec + 40;
    double rho[6], r[3], Jr[3][4];
    nominal_rot(q, ur, rho);
    for (int i = 0; i < 3; ++i) r[i] = c.R[i] * (ut[i] - rho[i]);
    for (int j = 0; j < 4; ++j) {
        double e[4] = {0.0, 0.0, 0.0, 0.0}, B[9], col[3];
        e[j] = 1.0;
        rot_bilinear(q, e, B);                       // d Rot / d q_j
        mat3_tmul(B, ur, col);
        for (int i = 0; i < 3; ++i) Jr[i][j] = col[i];
    }
    for (int j = 0; j < 4; ++j) {
        rec[j] = -2.0 * (r[0] * Jr[0][j] + r[1] * Jr[1][j] + r[2] * Jr[2][j]);
        for (int i = 0; i < 3; ++i) rec[4 + i * 4 + j] = -2.0 * c.R[i] * Jr[i][j];
        for (int l = 0; l < 4; ++l)
            rec[16 + j * 4 + l] = 2.0 * (c.R[0] * Jr[0][j] * Jr[0][l] + c.R[1] * Jr[1][j] * Jr[1][l] + c.R[2] * Jr[2][j] * Jr[2][l]);
    }
    for (int i = 0; i < 6; ++i) rec[32 + i] = rho[i];
    rec[38] = rec[39] = 0.0;
    {
        for (int j = 0; j < 4; ++j)
            for (int l = 0; l <= j; ++l) {
                double ej[4] = {0.0, 0.0, 0.0, 0.0}, el[4] = {0.0, 0.0, 0.0, 0.0}, B[9], col[3];
                ej[j] = 1.0; el[l] = 1.0;
                rot_bilinear(ej, el, B);             // d2 Rot / d q_j d q_l
                mat3_tmul(B, ur, col);
                const double v = -2.0 * (r[0] * col[0] + r[1] * col[1] + r[2] * col[2]);
                Sx[j * 4 + l] = v; Sx[l * 4 + j] = v;
            }
    }
}

// SQP termination: the QP step is below sqp_tol, or it is within 100 sqp_tol AND the decrease it predicts is at the
// rounding level of the objective (|g'd| <= 1e-13 max(1,|f|)): at that point the step is numerical noise of the
// gradient divided by a small curvature and never shrinks further, although the iterate no longer moves.
FT_HD bool sqp_step_converged(const ftmpc_config& cfg, double dmax, double gd, double f) {
    return dmax <= cfg.sqp_tol || (dmax <= 100.0 * cfg.sqp_tol && fabs(gd) <= 1e-13 * fmax(1.0, fabs(f)));
}
// Early stop in the quadratic regime: the full step d_k just taken and the full step d_{k-1} before it both came from
// the exact-Hessian QP (theta = 1; the augmented-Lagrangian convexification leaves the QP solution unchanged), so
// |d_{k+1}| ~ C |d_k|^2 with C estimated by |d_k| / |d_{k-1}|^2.  When that prediction is a decade below sqp_tol the QP
// that would only confirm it is not solved.  dprev = 0 disables the rule (phase_qp stores |d_{k-1}| only when iteration
// k-1 qualified); fast_dmax = 0 in the configuration switches it off altogether.
FT_HD bool sqp_fast_converged(const ftmpc_config& cfg, double dmax, double dprev, double alpha, double theta) {
    return cfg.fast_dmax > 0.0 && alpha == 1.0 && theta == 1.0 && dprev > 0.0 && dmax < dprev && dmax <= cfg.fast_dmax &&
           dmax * dmax * dmax <= 0.1 * cfg.sqp_tol * dprev * dprev;
}

// Stall detector: every stall_window iterations (from 2 stall_window on) step size and objective are compared with those a
// window earlier.  An iterate whose steps have not halved AND whose objective has moved by less than 1e-7 relative over a
// whole window is creeping along a flat, non-convex valley (the exact Hessian is indefinite there and every attempt falls
// back to the damped blend): it will not reach sqp_tol before the iteration cap and is reported as FTMPC_ST_MAXITER right
// away instead of occupying its SM for the remaining iterations (each with several failed factorisations).
// sc_ref = {step, objective} a window ago.  Returns true when stalled.
FT_HD bool sqp_stalled(const ftmpc_config& cfg, double iter, double dmax, double f, double* sc_dref, double* sc_fref) {
    const int w = cfg.stall_window;
    if (w <= 0) return false;
    const int it = (int)iter;
    if (it % w != 0) return false;
    const double dref = *sc_dref, fref = *sc_fref;
    *sc_dref = dmax;
    *sc_fref = f;
    return it >= 2 * w && dref > 0.0 && dmax > 0.5 * dref && fabs(fref - f) <= 1e-7 * fmax(1.0, fabs(f));
}

// forward rollout at U + alpha*d: states, cost, constraint values (c <= 0 feasible).
FT_HD void rollout_eval(const ftmpc_config& cfg, const WsLayout& L, const double* hull, const double* xref,
                        const double* uref, const double* U, const double* d, double alpha, double* X, double* C,
                        double& f, double& csum, double& cmax, double* Uconv = nullptr) {
    // U holds u~ (see FTMPC_CQ); with Uconv != nullptr it still holds u (warm start / zeros) and is converted on the fly
    const DynConsts k = dyn_consts(cfg);
    const int N = L.N;
    double x[FTMPC_NX], xn[FTMPC_NX], u[FTMPC_NU], Wr[FTMPC_NU], rho[FTMPC_NU] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < FTMPC_NX; ++i) x[i] = X[i];
    f = 0.0; csum = 0.0; cmax = 0.0;
    const double* Ah = hull;
    const double* bh = hull + FTMPC_NH * FTMPC_NU;
    for (int t = 0; t < N; ++t) {
        for (int j = 0; j < FTMPC_NU; ++j) u[j] = U[t * FTMPC_NU + j] + alpha * d[t * FTMPC_NU + j];
        if (uref) {
            nominal_rot(x + 9, uref + t * FTMPC_NU, rho);
            if (Uconv) for (int j = 0; j < FTMPC_NU; ++j) { u[j] += rho[j]; Uconv[t * FTMPC_NU + j] = u[j]; }
        }
        stage_wrench(cfg, u, nullptr, x + 9, Wr);
        for (int j = 0; j < FTMPC_NE; ++j) {                            // running cost, spiraling_mpc.py:188
            const double e = x[j] - xref[t * FTMPC_NE + j];
            f += cfg.Q[j] * e * e;
        }
        for (int j = 0; j < FTMPC_NU; ++j) f += cfg.R[j] * (u[j] - rho[j]) * (u[j] - rho[j]);
        for (int i = 0; i < FTMPC_NH; ++i) {                            // hull rows, spiraling_mpc.py:175-177
            double v = -bh[i];
            for (int j = 0; j < FTMPC_NU; ++j) v += Ah[i * FTMPC_NU + j] * Wr[j];
            C[t * FTMPC_NH + i] = v;
            if (v > 0.0) { csum += v; cmax = fmax(cmax, v); }
        }
        rk4_step(k, x, Wr, xn);                                        // spiraling_mpc.py:171
        for (int i = 0; i < FTMPC_NX; ++i) { x[i] = xn[i]; X[(t + 1) * FTMPC_NX + i] = xn[i]; }
    }
    double e[FTMPC_NE];
    for (int j = 0; j < FTMPC_NE; ++j) e[j] = x[j] - xref[N * FTMPC_NE + j];
    f += terminal_value(cfg, e);                                        // spiraling_mpc.py:195-196
    for (int i = 0; i < FTMPC_NF; ++i) {                                // terminal set, spiraling_mpc.py:199-202
        double v = -cfg.bf[i];
        for (int j = 0; j < FTMPC_NE; ++j) v += cfg.Af[i * FTMPC_NE + j] * e[j];
        C[FTMPC_NH * N + i] = v;
        if (v > 0.0) { csum += v; cmax = fmax(cmax, v); }
    }
    if (!(f == f) || !(csum == csum)) { f = INFINITY; csum = INFINITY; }
}

// ---- phase_ls: initialise (first != 0) or accept the QP step, then re-evaluate -----------------------
FT_HD void phase_ls(const ftmpc_config& cfg, const WsLayout& L, const StepIO& io, int inst, int slot, int first) {
    double* w = ws_slot(io, L, slot);
    double* sc = w + L.oSc;
    const int N = L.N;
    const double* xref = io.xref + (size_t)inst * io.xref_stride;
    const double* uref = io.uref ? io.uref + (size_t)inst * io.uref_stride : nullptr;
    const double* hull = io.hull_table + (size_t)io.hull_idx[inst] * FTMPC_HULL_STRIDE;
    double* U = w + L.oU;
    double* D = w + L.oD;
    double* X = w + L.oX;
    double* C = w + L.oC;
    double f, csum, cmax;
    if (first) {
        const DynConsts k = dyn_consts(cfg);
        robot_to_center(k, io.state + (size_t)inst * FTMPC_NX, X);          // spiraling_mpc.py:290
        const double* zw = io.z_warm + (size_t)inst * (L.n + (size_t)(N + 1) * FTMPC_NX);
        for (int t = 0; t < N; ++t)                                          // warm start shift, :324-331
            for (int j = 0; j < FTMPC_NU; ++j)
                U[t * FTMPC_NU + j] = (io.warm && t + 1 < N) ? zw[(t + 1) * FTMPC_NU + j] : 0.0;
        for (int i = 0; i < L.nv; ++i) D[i] = 0.0;
        for (int i = 0; i < L.m; ++i) w[L.oLam + i] = 0.0;
        for (int i = 0; i < SC_COUNT; ++i) sc[i] = 0.0;
        sc[SC_NU] = 1.0;
        sc[SC_THETA] = -1.0;                       // first QP uses the Gauss-Newton model
        sc[SC_STATUS] = FTMPC_ST_RUNNING;
        rollout_eval(cfg, L, hull, xref, uref, U, D, 0.0, X, C, f, csum, cmax, uref ? U : nullptr);
        if (!(f < INFINITY)) sc[SC_STATUS] = FTMPC_ST_QPFAIL;
    } else {
        if (sc[SC_STATUS] != FTMPC_ST_RUNNING) return;
        if (sc[SC_QPST] != 0.0) { sc[SC_STATUS] = FTMPC_ST_QPFAIL; return; }
        sc[SC_ITER] += 1.0;
        const double nu = fmax(sc[SC_NU], 1.1 * sc[SC_LAMMAX]);
        sc[SC_NU] = nu;
        const double phi0 = sc[SC_F] + nu * sc[SC_CSUM];
        const double dphi = sc[SC_GD] - nu * sc[SC_CSUM];
        double alpha = 1.0;
        for (;;) {
            rollout_eval(cfg, L, hull, xref, uref, U, D, alpha, X, C, f, csum, cmax);
            const double phi = f + nu * csum;
            if (phi <= phi0 + 1e-4 * alpha * dphi || alpha < 1e-8) break;
            alpha *= 0.5;
        }
        sc[SC_ALPHA] = alpha;
        if (!(f < INFINITY)) { sc[SC_STATUS] = FTMPC_ST_QPFAIL; return; }
        for (int i = 0; i < L.n; ++i) U[i] += alpha * D[i];
        if (sqp_step_converged(cfg, sc[SC_DMAX], sc[SC_GD], sc[SC_F]) ||
            sqp_fast_converged(cfg, sc[SC_DMAX], sc[SC_DPREV], alpha, sc[SC_THETA])) {
            sc[SC_STATUS] = (cmax <= cfg.feas_tol) ? FTMPC_ST_OK : FTMPC_ST_INFEASIBLE;
        } else if (sc[SC_ITER] >= cfg.max_sqp_iter || sqp_stalled(cfg, sc[SC_ITER], sc[SC_DMAX], f, sc + SC_DREF, sc + SC_FREF)) {
            sc[SC_STATUS] = FTMPC_ST_MAXITER;
        }
    }
    sc[SC_F] = f; sc[SC_CSUM] = csum; sc[SC_CMAX] = cmax;
    // terminal gradient / Hessian at the accepted point
    double e[FTMPC_NE];
    for (int j = 0; j < FTMPC_NE; ++j) e[j] = X[N * FTMPC_NX + j] - xref[N * FTMPC_NE + j];
    terminal_eval(cfg, e, w + L.oGV, w + L.oHV);
}

#if defined(__CUDACC__)
// ---- block-cooperative step acceptance (device only) ------------------------------------------------
// Same merit test as phase_ls, organised for one CTA:
//   A. the lanes of warp 0 roll the dynamics out for ALL backtracking step lengths at once
//      (lane i <-> alpha = 2^-i, i < 28: the serial loop of phase_ls stops at alpha < 1e-8 = lane 27),
//      states / stage wrenches go to shared scratch, the cost stays in a register;
//   B. constraint rows are evaluated row-parallel: first for alpha = 1 by the whole block (the common
//      case ends here), otherwise for the remaining step lengths warp-per-alpha;
//   C. the first step length passing the Armijo test is committed.
#define FTMPC_LS_NALPHA 28
// The rollouts of the step lengths are kept in shared memory FTMPC_LS_CHUNK at a time (alpha = 2^-a lives in slot a % CHUNK):
// the first chunk holds 2^0 .. 2^-11, which covers every accepted step but a handful per million; later chunks are only
// rolled out when all of the previous one were rejected.  Same acceptance rule (first = largest passing step length).
#define FTMPC_LS_CHUNK 12
struct LsScratch {
    double *Xs, *Ws, *fa, *csa, *cma, *hull, *U, *D, *xref, *uref, *rec, *out;
    int xs_stride, ws_stride;
};
__host__ __device__ inline size_t ls_scratch_doubles(int N) {
    return (size_t)FTMPC_LS_CHUNK * ((N + 1) * FTMPC_NX + 1 + N * FTMPC_NU + 1) + 96 + FTMPC_HULL_STRIDE +
           2 * (size_t)(FTMPC_NU * N + 1) + (size_t)(N + 1) * (FTMPC_NE + FTMPC_NU) + 8 +
           (size_t)(FTMPC_MAX_POLY + FTMPC_MAX_ROOT) * FTMPC_TERM_REC + FTMPC_TERM_REC + 8;
}
__device__ __forceinline__ LsScratch ls_carve(double* buf, int N) {
    LsScratch s;
    s.xs_stride = (N + 1) * FTMPC_NX + 1;          // odd strides: lanes of warp 0 hit different banks
    s.ws_stride = N * FTMPC_NU + 1;
    double* p = buf;
    s.Xs = p; p += (size_t)FTMPC_LS_CHUNK * s.xs_stride;
    s.Ws = p; p += (size_t)FTMPC_LS_CHUNK * s.ws_stride;
    s.fa = p; p += 32;
    s.csa = p; p += 32;
    s.cma = p; p += 32;
    p += ((size_t)(p - buf)) & 1;                  // 16-byte aligned staging buffers (bulk copies)
    s.hull = p; p += FTMPC_HULL_STRIDE;
    s.xref = p; p += (((size_t)(N + 1) * FTMPC_NE + 2) + 1) & ~(size_t)1;       // + 2: a row may start / end on an odd double
    s.uref = p; p += (((size_t)(N + 1) * FTMPC_NU + 2) + 1) & ~(size_t)1;
    s.U = p; p += FTMPC_NU * N + 1;
    s.D = p; p += FTMPC_NU * N + 1;
    s.rec = p; p += (size_t)(FTMPC_MAX_POLY + FTMPC_MAX_ROOT) * FTMPC_TERM_REC;
    s.out = p; p += FTMPC_TERM_REC;
    return s;
}

// ---- split rollout --------------------------------------------------------------------------------------
// The attitude (omega, q) of the centre state evolves on its own (torque in, no dependence on p, v), and the translation
// is a quadrature over the attitude stages:  v_{t+1} = v_t + dv_t,  p_{t+1} = p_t + dt v_t + cp_t  with
//   dv_t = dt/6 (kv1 + 2 kv2 + 2 kv3 + kv4),  cp_t = dt^2/6 (kv1 + kv2 + kv3)      (exactly the RK4 update of rk4_step),
// kv_s = dyn_v at the s-th attitude stage.  Only the 7-state attitude chain is serial (40 % of the instructions of the full
// step); the per-stage quadrature terms are independent tasks for the whole block, followed by a 2-FMA-per-stage scan.
__device__ __noinline__ void rollout_attitude(const ftmpc_config& cfg, int N, const double* U, const double* d, double alpha,
                                              const double* ur_conv /* u_ref when U still holds u (first rollout) */,
                                              const double* x0, double* Xs) {
    const DynConsts k = dyn_consts(cfg);
    const AttConsts ac = att_consts(k);
    double w[3], q[4];
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = x0[6 + i];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = x0[9 + i];
#pragma unroll
    for (int i = 0; i < FTMPC_NX; ++i) Xs[i] = x0[i];
    const double cs[4] = {0.0, 0.5 * k.dt, 0.5 * k.dt, k.dt};
    const double bs[4] = {k.dt / 6.0, k.dt / 3.0, k.dt / 3.0, k.dt / 6.0};
    for (int t = 0; t < N; ++t) {
        double tau[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            tau[j] = U[t * FTMPC_NU + 3 + j] + alpha * d[t * FTMPC_NU + 3 + j];
            if (ur_conv) tau[j] += ur_conv[t * FTMPC_NU + 3 + j];
        }
        double kw[3] = {0.0, 0.0, 0.0}, kq[4] = {0.0, 0.0, 0.0, 0.0}, aw[3], aq[4];
#pragma unroll
        for (int i = 0; i < 3; ++i) aw[i] = w[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) aq[i] = q[i];
        // the chain form of the attitude right-hand side (dyn_wq_chain): per-step constants outside the serial RK stages
        const double tq[3] = {tau[0] * k.iJ[0], tau[1] * k.iJ[1], tau[2] * k.iJ[2]};
#pragma unroll
        for (int st = 0; st < 4; ++st) {
            double sw[3], sq[4];
#pragma unroll
            for (int i = 0; i < 3; ++i) sw[i] = w[i] + cs[st] * kw[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) sq[i] = q[i] + cs[st] * kq[i];
            dyn_wq_chain(ac, tq, sw, sq, kw, kq);
#pragma unroll
            for (int i = 0; i < 3; ++i) aw[i] += bs[st] * kw[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) aq[i] += bs[st] * kq[i];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) { w[i] = aw[i]; Xs[(t + 1) * FTMPC_NX + 6 + i] = aw[i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) { q[i] = aq[i]; Xs[(t + 1) * FTMPC_NX + 9 + i] = aq[i]; }
    }
}
// One (step length, stage) task: stage wrench -> Ws, quadrature terms dv_t / cp_t -> the (v, p) slots of stage t + 1 of Xs
// (turned into states by rollout_scan), input + attitude-rate cost of the stage -> *cost.
__device__ __noinline__ void rollout_stage_task(const ftmpc_config& cfg, int t, const double* xref, const double* uref,
                                                const double* U, const double* d, double alpha, double* Xs, double* Ws,
                                                double* cost, double* Uconv_s, double* Uconv_g) {
    const DynConsts k = dyn_consts(cfg);
    const double* xt = Xs + t * FTMPC_NX;
    double u[FTMPC_NU], rho[FTMPC_NU] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, Wr[FTMPC_NU];
#pragma unroll
    for (int j = 0; j < FTMPC_NU; ++j) u[j] = U[t * FTMPC_NU + j] + alpha * d[t * FTMPC_NU + j];
    if (uref) {                                       // U holds u~ = u + rho(q_t), see FTMPC_CQ
        nominal_rot(xt + 9, uref + t * FTMPC_NU, rho);
        if (Uconv_s) {                                // first rollout: U still holds u
#pragma unroll
            for (int j = 0; j < FTMPC_NU; ++j) { u[j] += rho[j]; Uconv_s[t * FTMPC_NU + j] = u[j]; Uconv_g[t * FTMPC_NU + j] = u[j]; }
        }
    }
    stage_wrench(cfg, u, nullptr, xt + 9, Wr);
    double f = 0.0;
#pragma unroll
    for (int j = 0; j < FTMPC_NU; ++j) { f += cfg.R[j] * (u[j] - rho[j]) * (u[j] - rho[j]); Ws[t * FTMPC_NU + j] = Wr[j]; }
#pragma unroll
    for (int j = 6; j < FTMPC_NE; ++j) {
        const double e = xt[j] - xref[t * FTMPC_NE + j];
        f += cfg.Q[j] * e * e;
    }
    *cost = f;
    // attitude stages of this step again (cheaper than parking 4 x 10 doubles per stage and step length)
    double w[3], q[4], kw[3] = {0.0, 0.0, 0.0}, kq[4] = {0.0, 0.0, 0.0, 0.0}, dv[3] = {0.0, 0.0, 0.0}, cp[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = xt[6 + i];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = xt[9 + i];
    const double cs[4] = {0.0, 0.5 * k.dt, 0.5 * k.dt, k.dt};
    const double bs[4] = {k.dt / 6.0, k.dt / 3.0, k.dt / 3.0, k.dt / 6.0};
    const double cq = k.dt * k.dt / 6.0;
    const AttConsts ac = att_consts(k);
    const double tq[3] = {Wr[3] * k.iJ[0], Wr[4] * k.iJ[1], Wr[5] * k.iJ[2]};
#pragma unroll
    for (int st = 0; st < 4; ++st) {
        double sw[3], sq[4], kv[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) sw[i] = w[i] + cs[st] * kw[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) sq[i] = q[i] + cs[st] * kq[i];
        dyn_wq_chain(ac, tq, sw, sq, kw, kq);                       // (the same arithmetic as rollout_attitude: identical stage values)
        dyn_v(k, sw, sq, kw, Wr, kv);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            dv[i] += bs[st] * kv[i];
            if (st < 3) cp[i] += cq * kv[i];
        }
    }
    double* xn = Xs + (t + 1) * FTMPC_NX;
#pragma unroll
    for (int i = 0; i < 3; ++i) { xn[i] = cp[i]; xn[3 + i] = dv[i]; }
}
// component c of (p, v) along the horizon from the quadrature terms; returns the position + velocity cost of that component
__device__ __forceinline__ double rollout_scan(const ftmpc_config& cfg, int N, int c, const double* xref, double* Xs) {
    double p = Xs[c], v = Xs[3 + c], f = 0.0;
    for (int t = 0; t < N; ++t) {
        const double ep = p - xref[t * FTMPC_NE + c], ev = v - xref[t * FTMPC_NE + 3 + c];
        f += cfg.Q[c] * ep * ep + cfg.Q[3 + c] * ev * ev;
        double* xn = Xs + (t + 1) * FTMPC_NX;
        const double pn = p + cfg.dt * v + xn[c], vn = v + xn[3 + c];
        xn[c] = pn; xn[3 + c] = vn;
        p = pn; v = vn;
    }
    return f;
}

// value of constraint row p (c <= 0 feasible) from stored stage wrenches / terminal state
__device__ __forceinline__ double cons_value(const ftmpc_config& cg /* global copy: per-thread rows */, int N,
                                             const double* hull, const double* xrefN, const double* Xs,
                                             const double* Ws, int p) {
    if (p < FTMPC_NH * N) {
        const int t = p / FTMPC_NH, i = p - t * FTMPC_NH;
        double v = -hull[FTMPC_NH * FTMPC_NU + i];
#pragma unroll
        for (int j = 0; j < FTMPC_NU; ++j) v += hull[i * FTMPC_NU + j] * Ws[t * FTMPC_NU + j];
        return v;
    }
    const int i = p - FTMPC_NH * N;
    double v = -cg.bf[i];
#pragma unroll
    for (int j = 0; j < FTMPC_NE; ++j) v += cg.Af[i * FTMPC_NE + j] * (Xs[N * FTMPC_NX + j] - xrefN[j]);
    return v;
}

__device__ __forceinline__ void phase_ls_block(CudaBlock& blk, const ftmpc_config& cfg, const WsLayout& L,
                                               const StepIO& io, int inst, int slot, int first, double* scratch) {
    double* w = ws_slot(io, L, slot);
    double* sc = w + L.oSc;
    const int N = L.N, tid = blk.tid(), nt = blk.nthreads(), lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const double* xref_g = io.xref + (size_t)inst * io.xref_stride;
    const double* uref_g = io.uref ? io.uref + (size_t)inst * io.uref_stride : nullptr;
    const double* hull_g = io.hull_table + (size_t)io.hull_idx[inst] * FTMPC_HULL_STRIDE;
    const ftmpc_config& cg = *io.cfg_g;
    double* U = w + L.oU;
    double* D = w + L.oD;
    double* X = w + L.oX;
    double* C = w + L.oC;
    LsScratch s = ls_carve(scratch, N);
    const int nterm = cfg.n_poly + cfg.n_root;
    double nu = 1.0, phi0 = 0.0, dphi = 0.0, dmax = 0.0, iter = 0.0, gd_prev = 0.0, f_prev = 0.0, dprev = 0.0, theta_qp = 0.0;
    if (first) {
        if (tid == 0) robot_to_center(dyn_consts(cfg), io.state + (size_t)inst * FTMPC_NX, X);      // spiraling_mpc.py:290
        const double* zw = io.z_warm + (size_t)inst * (L.n + (size_t)(N + 1) * FTMPC_NX);
        for (int i = tid; i < L.n; i += nt) {                                                       // warm start shift, :324-331
            const int t = i / FTMPC_NU;
            const double v = (io.warm && t + 1 < N) ? zw[i + FTMPC_NU] : 0.0;
            U[i] = v; s.U[i] = v; s.D[i] = 0.0;
        }
        for (int i = tid; i < L.nv; i += nt) D[i] = 0.0;
        for (int i = tid; i < L.m; i += nt) w[L.oLam + i] = 0.0;
    } else {
        const double status = sc[SC_STATUS], qpst = sc[SC_QPST];
        nu = fmax(sc[SC_NU], 1.1 * sc[SC_LAMMAX]);
        phi0 = sc[SC_F] + nu * sc[SC_CSUM];
        dphi = sc[SC_GD] - nu * sc[SC_CSUM];
        dmax = sc[SC_DMAX];
        gd_prev = sc[SC_GD];
        f_prev = sc[SC_F];
        iter = sc[SC_ITER] + 1.0;
        dprev = sc[SC_DPREV];
        theta_qp = sc[SC_THETA];
        blk.sync();                                   // everybody has read the scalars
        if (status != FTMPC_ST_RUNNING) return;
        if (qpst != 0.0) {
            if (tid == 0) sc[SC_STATUS] = (qpst == 5.0) ? FTMPC_ST_REDO : FTMPC_ST_QPFAIL;
            blk.sync();
            return;
        }
        for (int i = tid; i < L.n; i += nt) { s.U[i] = U[i]; s.D[i] = D[i]; }
    }
    // per-instance data -> shared memory: input-bound hull (A_h, b_h), reference window, nominal wrench
    if (blk.tma.addr) {
        // TMA: thread 0 issues three bulk copies against the CTA's mbarrier, everybody waits on it
        double* xr_raw = s.xref;
        double* ur_raw = s.uref;
        s.xref = tma_row_ptr(xr_raw, xref_g);
        if (uref_g) s.uref = tma_row_ptr(ur_raw, uref_g);
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic-proxy accesses of these buffers are done (barrier above)
            unsigned bytes = FTMPC_HULL_STRIDE * 8u + tma_row_bytes(xref_g, (N + 1) * FTMPC_NE);
            if (uref_g) bytes += tma_row_bytes(uref_g, (N + 1) * FTMPC_NU);
            tma_expect(blk.tma, bytes);
            tma_bulk_g2s(blk.tma, s.hull, hull_g, FTMPC_HULL_STRIDE * 8u);
            tma_stage_row(blk.tma, xr_raw, xref_g, (N + 1) * FTMPC_NE);
            if (uref_g) tma_stage_row(blk.tma, ur_raw, uref_g, (N + 1) * FTMPC_NU);
        }
        tma_wait(blk.tma);
    } else {
        for (int i = tid; i < FTMPC_HULL_STRIDE; i += nt) s.hull[i] = hull_g[i];
        for (int i = tid; i < (N + 1) * FTMPC_NE; i += nt) s.xref[i] = xref_g[i];
        if (uref_g) for (int i = tid; i < (N + 1) * FTMPC_NU; i += nt) s.uref[i] = uref_g[i];
    }
    blk.sync();
    const double* xrefN = s.xref + N * FTMPC_NE;
    // A. rollouts, one step length per lane
    const int nalpha = first ? 1 : FTMPC_LS_NALPHA;
    const bool conv = first && uref_g;              // accelerating reference: U still holds u (warm start), see FTMPC_CQ
    auto slot_x = [&](int a) { return s.Xs + (size_t)(a % FTMPC_LS_CHUNK) * s.xs_stride; };
    auto slot_w = [&](int a) { return s.Ws + (size_t)(a % FTMPC_LS_CHUNK) * s.ws_stride; };
    // A1. attitude chains of the step lengths [base, base + CHUNK), one per lane of warp 0
    auto attitude_chunk = [&](int base) {
        const int a = base + lane;
        if (warp == 0 && lane < FTMPC_LS_CHUNK && a < nalpha)
            rollout_attitude(cfg, N, s.U, s.D, first ? 0.0 : ldexp(1.0, -a), conv ? s.uref : nullptr, X, slot_x(a));
        blk.sync();
    };
    attitude_chunk(0);
    // A2-A4 for the step lengths [a0, a1): stage tasks spread over the warps (task k runs on lane k / nw of warp k % nw, so a
    // handful of tasks costs one pass of one lane per warp), (p, v) scan per component, cost per step length
    double* stage_cost = s.rec;                     // [nalpha][N], free until the terminal records are built
    double* pv_cost = s.rec + (size_t)FTMPC_LS_NALPHA * N;    // [nalpha][3]
    auto rollout_rest = [&](int a0, int a1) {
        const int ntask = (a1 - a0) * N;
        for (int idx = lane * nw + warp; idx < ntask; idx += nt) {
            const int a = a0 + idx / N, t = idx - (a - a0) * N;
            rollout_stage_task(cfg, t, s.xref, uref_g ? s.uref : nullptr, s.U, s.D, first ? 0.0 : ldexp(1.0, -a),
                               slot_x(a), slot_w(a), stage_cost + a * N + t, conv ? s.U : nullptr, U);
        }
        blk.sync();
        for (int idx = tid; idx < 3 * (a1 - a0); idx += nt) {
            const int a = a0 + idx / 3, c = idx - (a - a0) * 3;
            pv_cost[a * 3 + c] = rollout_scan(cfg, N, c, s.xref, slot_x(a));
        }
        blk.sync();
        for (int a = a0 + tid; a < a1; a += nt) {
            double f = (pv_cost[a * 3] + pv_cost[a * 3 + 1]) + pv_cost[a * 3 + 2];
            for (int t = 0; t < N; ++t) f += stage_cost[a * N + t];
            s.fa[a] = f;
        }
        blk.sync();
    };
    rollout_rest(0, 1);
    blk.mark(PH_LS_ROLL);
    // B0. alpha index 0: constraint rows by the whole block (straight into C), terminal cost with its
    //     gradient / Hessian records one term per thread (alpha = 1 is accepted most of the time)
    int win = -1;
    {
        if (tid < nterm) {
            double e[FTMPC_NE];
#pragma unroll
            for (int j = 0; j < FTMPC_NE; ++j) e[j] = s.Xs[N * FTMPC_NX + j] - xrefN[j];
            double* rec = s.rec + (size_t)tid * FTMPC_TERM_REC;
            for (int i = 0; i < FTMPC_TERM_REC; ++i) rec[i] = 0.0;
            term_eval(term_desc(cg, tid), e, rec);
        }
        double cs = 0.0, cm = 0.0;
        for (int p = tid; p < L.mc; p += nt) {
            const double v = cons_value(cg, N, s.hull, xrefN, s.Xs, s.Ws, p);
            C[p] = v;
            if (v > 0.0) { cs += v; cm = fmax(cm, v); }
        }
        cs = blk.sum(cs);                             // barrier: the term records are complete
        cm = blk.max(cm);
        if (tid < FTMPC_TERM_REC) {
            double v = 0.0;
            for (int k = 0; k < nterm; ++k) v += s.rec[(size_t)k * FTMPC_TERM_REC + tid];
            s.out[tid] = v;
        }
        blk.sync();
        double f = s.fa[0] + cfg.term_const + s.out[0];
        if (!(f == f) || !(cs == cs)) { f = INFINITY; cs = INFINITY; }
        blk.sync();                                   // everybody has read fa[0] before it is overwritten
        if (tid == 0) { s.fa[0] = f; s.csa[0] = cs; s.cma[0] = cm; }
        if (first || f + nu * cs <= phi0 + 1e-4 * dphi) win = 0;
    }
    blk.sync();
    if (win < 0) {
        for (int base = 0; base < nalpha && win < 0; base += FTMPC_LS_CHUNK) {
            if (base > 0) attitude_chunk(base);       // (the attitude chains of the first chunk exist)
            const int a_lo = base > 1 ? base : 1, a_hi = (base + FTMPC_LS_CHUNK < nalpha) ? base + FTMPC_LS_CHUNK : nalpha;
            rollout_rest(a_lo, a_hi);                 // translations of these step lengths
            // B1. rows one warp per alpha, terminal cost one thread per (alpha, term)
            for (int idx = tid; idx < (a_hi - a_lo) * nterm; idx += nt) {
                const int a = a_lo + idx / nterm, k = idx - (a - a_lo) * nterm;
                const double* Xa = slot_x(a);
                double e[FTMPC_NE];
#pragma unroll
                for (int j = 0; j < FTMPC_NE; ++j) e[j] = Xa[N * FTMPC_NX + j] - xrefN[j];
                s.rec[(size_t)a * (FTMPC_MAX_POLY + FTMPC_MAX_ROOT) + k] = term_eval(term_desc(cg, k), e, nullptr);
            }
            for (int a = a_lo + warp; a < a_hi; a += nw) {
                const double* Xa = slot_x(a);
                const double* Wa = slot_w(a);
                double cs = 0.0, cm = 0.0;
                for (int p = lane; p < L.mc; p += 32) {
                    const double v = cons_value(cg, N, s.hull, xrefN, Xa, Wa, p);
                    if (v > 0.0) { cs += v; cm = fmax(cm, v); }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    cs += __shfl_xor_sync(0xffffffffu, cs, o);
                    cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, o));
                }
                if (lane == 0) { s.csa[a] = cs; s.cma[a] = cm; }
            }
            blk.sync();
            if (tid >= a_lo && tid < a_hi) {
                double v = 0.0;
                for (int k = 0; k < nterm; ++k) v += s.rec[(size_t)tid * (FTMPC_MAX_POLY + FTMPC_MAX_ROOT) + k];
                double f = s.fa[tid] + cfg.term_const + v, cs = s.csa[tid];
                if (!(f == f) || !(cs == cs)) { f = INFINITY; cs = INFINITY; }
                s.fa[tid] = f; s.csa[tid] = cs;
            }
            blk.sync();
            for (int a = a_lo; a < a_hi && win < 0; ++a) {
                const double alpha = ldexp(1.0, -a);
                if (s.fa[a] + nu * s.csa[a] <= phi0 + 1e-4 * alpha * dphi || alpha < 1e-8) win = a;
            }
            blk.sync();                               // the per-alpha term values have been consumed
        }
        // constraint values and terminal records of the accepted step
        const double* Xa = slot_x(win);
        const double* Wa = slot_w(win);
        for (int p = tid; p < L.mc; p += nt) C[p] = cons_value(cg, N, s.hull, xrefN, Xa, Wa, p);
        if (tid < nterm) {
            double e[FTMPC_NE];
#pragma unroll
            for (int j = 0; j < FTMPC_NE; ++j) e[j] = Xa[N * FTMPC_NX + j] - xrefN[j];
            double* rec = s.rec + (size_t)tid * FTMPC_TERM_REC;
            for (int i = 0; i < FTMPC_TERM_REC; ++i) rec[i] = 0.0;
            term_eval(term_desc(cg, tid), e, rec);
        }
        blk.sync();
        if (tid < FTMPC_TERM_REC) {
            double v = 0.0;
            for (int k = 0; k < nterm; ++k) v += s.rec[(size_t)k * FTMPC_TERM_REC + tid];
            s.out[tid] = v;
        }
        blk.sync();
    }
    blk.mark(PH_LS_EVAL);
    if (win > 0) blk.count(CT_LS_BACKTRACK);
    // C. commit
    const double alpha = first ? 0.0 : ldexp(1.0, -win);
    const double f = s.fa[win], csum = s.csa[win], cmax = s.cma[win];
    const bool finite = f < INFINITY;
    if (!first && finite) for (int i = tid; i < L.n; i += nt) U[i] = s.U[i] + alpha * s.D[i];
    const double* Xa = slot_x(win);
    if (finite) for (int i = tid; i < (N + 1) * FTMPC_NX; i += nt) X[i] = Xa[i];
    if (finite || first) {                            // terminal gradient / Hessian at the accepted point
        for (int i = tid; i < FTMPC_NE; i += nt) w[L.oGV + i] = s.out[1 + i];
        for (int i = tid; i < FTMPC_NE * FTMPC_NE; i += nt) w[L.oHV + i] = s.out[10 + i];
    }
    if (tid == 0) {
        if (first) {
            for (int i = 0; i < SC_COUNT; ++i) sc[i] = 0.0;
            sc[SC_NU] = 1.0;
            sc[SC_THETA] = -1.0;                   // first QP uses the Gauss-Newton model
            sc[SC_STATUS] = finite ? FTMPC_ST_RUNNING : FTMPC_ST_QPFAIL;
        } else {
            sc[SC_ITER] = iter;
            sc[SC_NU] = nu;
            sc[SC_ALPHA] = alpha;
            if (!finite) sc[SC_STATUS] = FTMPC_ST_QPFAIL;
            else if (sqp_step_converged(cfg, dmax, gd_prev, f_prev) || sqp_fast_converged(cfg, dmax, dprev, alpha, theta_qp)) sc[SC_STATUS] = (cmax <= cfg.feas_tol) ? FTMPC_ST_OK : FTMPC_ST_INFEASIBLE;
            else if (iter >= cfg.max_sqp_iter || sqp_stalled(cfg, iter, dmax, f, sc + SC_DREF, sc + SC_FREF)) sc[SC_STATUS] = FTMPC_ST_MAXITER;
        }
        if (finite || first) { sc[SC_F] = f; sc[SC_CSUM] = csum; sc[SC_CMAX] = cmax; }
    }
    blk.sync();
    blk.mark(PH_LS_TERM);
    blk.count(first ? CT_INST : CT_SQP);
}
#endif  // __CUDACC__

// On the device the (stage, column) sweep is kept out of line: inside the persistent kernel it would otherwise
// inherit the register pressure of everything around it and spill (local memory is L2-latency here, the
// shared-memory carve-out leaves almost no L1).
#if defined(__CUDA_ARCH__)
__device__ __noinline__ void rk4_column_call(const DynConsts& k, const double* x, const double* Wr, int col, const double* lam,
                                             double* jac_col, double* hess_col) {
    rk4_column(k, x, Wr, col, lam, jac_col, hess_col);
}
#else
FT_HD void rk4_column_call(const DynConsts& k, const double* x, const double* Wr, int col, const double* lam,
                           double* jac_col, double* hess_col) {
    rk4_column(k, x, Wr, col, lam, jac_col, hess_col);
}
#endif

// ---- Hessian schedule: the blend the next QP STARTS from --------------------------------------------------------------
// (phase_qp only ever lowers theta inside an iteration, so theta = 0 here means the exact second-order terms W_t are never
//  read in this SQP iteration: the linearisation then skips the costate recursion and the Hessian sweeps -- 38 % of the SQP
//  iterations of the bench workload, 9 % of the kernel's time.  ONE function for both phases, so they cannot disagree.)
struct QpStart {
    double theta, sigma;
    bool can_aug, skip_exact;
};
FT_HD QpStart qp_start(const ftmpc_config& cfg, const double* sc) {
    QpStart q;
    // the convexified QP equals the plain one whenever the predicted rows stay active, whatever the feasibility of the
    // iterate; it is only kept away from the first, wildly infeasible iterations (elastic variable far from zero)
    const double feas_aug = 1e-2;
    q.can_aug = (sc[SC_ITER] > 0.0 || sc[SC_THETA] >= 0.0) && sc[SC_CSUM] <= feas_aug;
    double theta = sc[SC_THETA];
    q.sigma = 0.0;
    theta = (theta < 0.0) ? 0.0 : ((theta == 0.0) ? cfg.theta_first : fmin(1.0, cfg.theta_growth * theta));
    if (q.can_aug && sc[SC_SIGMA] > 0.0) { theta = 1.0; q.sigma = sc[SC_SIGMA]; }
    // while the iterate is still infeasible the exact Hessian is almost always indefinite and no convexification is
    // allowed: once an attempt has fallen all the way back to Gauss-Newton, do not pay for the failing attempts again
    // until feasibility is reached
    // far from the solution (last QP step still large) the blended Hessian is almost always indefinite: start the
    // blend only once the steps have become moderate (|d|_inf <= blend_dmax; measured: 2x fewer failed attempts for
    // +0.5 % iterations)
    const bool far = sc[SC_ITER] > 0.0 && sc[SC_THETA] <= 0.0 && sc[SC_DMAX] > cfg.blend_dmax;
    q.skip_exact = (sc[SC_HFAIL] != 0.0 && sc[SC_CSUM] > feas_aug) || far;
    if (q.skip_exact) theta = 0.0;
    q.theta = theta;
    return q;
}

// ---- phase_lin: Jacobians, costates, stage Hessians (lanes of one warp / a serial loop) -------------
template <class Blk>
FT_HD void phase_lin(Blk& blk, const ftmpc_config& cfg, const WsLayout& L, const StepIO& io, int inst, int slot) {
    double* w = ws_slot(io, L, slot);
    const double* sc = w + L.oSc;
    if (sc[SC_STATUS] != FTMPC_ST_RUNNING) return;
    const int N = L.N, tid = blk.tid(), nt = blk.nthreads();
    const DynConsts k = dyn_consts(cfg);
    const double* xref = io.xref + (size_t)inst * io.xref_stride;
    const double* uref = io.uref ? io.uref + (size_t)inst * io.uref_stride : nullptr;
    const double* U = w + L.oU;
    const double* X = w + L.oX;
    double* Jz = w + L.oJz;
    double* Wz = w + L.oWz;
    double* Mu = w + L.oMu;
    const double* lam = w + L.oLam;
    double* Cq = w + L.oCq;
    // first-order columns (U holds u~: the wrench does not depend on the attitude, see FTMPC_CQ)
    for (int it = tid; it < N * 13; it += nt) {
        const int t = it / 13, col = it % 13;
        double Wr[FTMPC_NU];
        stage_wrench(cfg, U + t * FTMPC_NU, nullptr, X + t * FTMPC_NX + 9, Wr);
        double jc[13];
        rk4_column_call(k, X + t * FTMPC_NX, Wr, col, nullptr, jc, nullptr);
        for (int i = 0; i < 13; ++i) Jz[(size_t)it * 13 + i] = jc[i];
    }
    if (uref)
        for (int t = tid; t < N; t += nt)
            stage_cost_coupling(cfg, U + t * FTMPC_NU, uref + t * FTMPC_NU, X + t * FTMPC_NX + 9, Cq + (size_t)t * FTMPC_CQ);
    blk.sync();
    blk.mark(PH_LIN_JAC);
    if (qp_start(cfg, sc).theta == 0.0) { blk.mark(PH_LIN); return; }        // Gauss-Newton iteration: W_t is never read
    // costates  mu_N = [grad V_f + A_f' lam_term ; 0],  mu_t = [2Q e_t ; 0] + A_t' mu_{t+1}
    for (int i = tid; i < FTMPC_NX; i += nt) {
        double v = 0.0;
        if (i < FTMPC_NE) {
            v = w[L.oGV + i];
            for (int r = 0; r < FTMPC_NF; ++r) v += cfg.Af[r * FTMPC_NE + i] * lam[FTMPC_NH * N + r];
        }
        Mu[N * FTMPC_NX + i] = v;
    }
    blk.sync();
    for (int t = N - 1; t >= 1; --t) {
        const double* mn = Mu + (t + 1) * FTMPC_NX;
        for (int i = tid; i < FTMPC_NX; i += nt) {
            double v = (i < FTMPC_NE) ? 2.0 * cfg.Q[i] * (X[t * FTMPC_NX + i] - xref[t * FTMPC_NE + i])
                                      : (uref ? Cq[(size_t)t * FTMPC_CQ + i - FTMPC_NE] : 0.0);
            if (i < 3) v += mn[i];
            else if (i < 6) v += k.dt * mn[i - 3] + mn[i];
            else {
                const double* col = Jz + ((size_t)t * 13 + (i - 6)) * 13;
                for (int r = 0; r < 13; ++r) v += col[r] * mn[r];
            }
            Mu[t * FTMPC_NX + i] = v;
        }
        blk.sync();
    }
    blk.mark(PH_LIN_MU);
    // second-order columns: Hessian of mu_{t+1}' RK4(x_t, W_t) in z-space
    for (int it = tid; it < N * 13; it += nt) {
        const int t = it / 13, col = it % 13;
        double Wr[FTMPC_NU];
        stage_wrench(cfg, U + t * FTMPC_NU, nullptr, X + t * FTMPC_NX + 9, Wr);
        double jc[13], hc[13];
        rk4_column_call(k, X + t * FTMPC_NX, Wr, col, Mu + (t + 1) * FTMPC_NX, jc, hc);
        for (int i = 0; i < 13; ++i) Wz[(size_t)it * 13 + i] = hc[i];
    }
    blk.sync();
    if (uref) {                                     // exact remainder of the input cost's (q,q) Hessian
        for (int idx = tid; idx < N * 16; idx += nt) {
            const int t = idx >> 4, a = (idx >> 2) & 3, b = idx & 3;
            Wz[(size_t)t * 169 + (3 + a) * 13 + 3 + b] += Cq[(size_t)t * FTMPC_CQ + 40 + a * 4 + b];
        }
        blk.sync();
    }
    blk.mark(PH_LIN);
}

// ---- the MPC QP seen by the active-set solver -----------------------------------------------------
// variables (d[0..n), delta); extended rows: [d ; delta ; dx_N[0..9)].  Row i (i < mc):
//   J_i d - delta * max(c_i,0) <= -c_i        (Powell/Schittkowski elastic relaxation)
struct MpcCons {
    int N, n, nv, mc;
    const double* Ah;       // [26][6]
    const double* Af;       // [72][9]
    const double* cv;       // [mc] constraint values at the linearisation point
    const double* tf_val;   // optional compact terminal rows (<= 2 non-zeros per row of A_f): [72][2] values ...
    const int* tf_idx;      // ... and column indices; nullptr -> dense rows of Af
    // fixed layout (zeros kept) so that the row stays in registers: hull rows 6 + 1 entries, terminal rows 9 + 1
    FT_HD void row(int p, SparseRow& r) const {
        if (p < FTMPC_NH * N) {
            const int t = p / FTMPC_NH, i = p - t * FTMPC_NH;
            for (int j = 0; j < FTMPC_NU; ++j) { r.idx[j] = t * FTMPC_NU + j; r.val[j] = -Ah[i * FTMPC_NU + j]; }
            const double c = cv[p];
            r.idx[FTMPC_NU] = n; r.val[FTMPC_NU] = (c > 0.0) ? c : 0.0;
            r.nnz = FTMPC_NU + 1;
            r.beta = c;
        } else if (p < mc) {
            const int i = p - FTMPC_NH * N;
            for (int j = 0; j < FTMPC_NE; ++j) { r.idx[j] = nv + j; r.val[j] = -Af[i * FTMPC_NE + j]; }
            const double c = cv[p];
            r.idx[FTMPC_NE] = n; r.val[FTMPC_NE] = (c > 0.0) ? c : 0.0;
            r.nnz = FTMPC_NE + 1;
            r.beta = c;
        } else if (p == mc) {          // delta >= 0
            r.idx[0] = n; r.val[0] = 1.0; r.nnz = 1; r.beta = 0.0;
        } else {                       // delta <= 1
            r.idx[0] = n; r.val[0] = -1.0; r.nnz = 1; r.beta = -1.0;
        }
    }
    // n_p . v - sb * beta_p without materialising the sparse row (v in extended coordinates [d ; delta ; dx_N])
    FT_HD double slack(int p, const double* v, double sb) const {
        if (p < FTMPC_NH * N) {
            const int t = p / FTMPC_NH, i = p - t * FTMPC_NH;
            const double* a = Ah + i * FTMPC_NU;
            const double* x = v + t * FTMPC_NU;
            const double c = cv[p];
            double s = -sb * c;
            for (int j = 0; j < FTMPC_NU; ++j) s -= a[j] * x[j];
            if (c > 0.0) s += c * v[n];
            return s;
        }
        if (p < mc) {
            const int i = p - FTMPC_NH * N;
            const double c = cv[p];
            double s = -sb * c;
            if (tf_val) {                      // skipping exact zeros leaves the sum bit-identical
                s -= tf_val[2 * i] * v[nv + tf_idx[2 * i]];
                s -= tf_val[2 * i + 1] * v[nv + tf_idx[2 * i + 1]];
            } else {
                const double* a = Af + i * FTMPC_NE;
                for (int j = 0; j < FTMPC_NE; ++j) s -= a[j] * v[nv + j];
            }
            if (c > 0.0) s += c * v[n];
            return s;
        }
        return (p == mc) ? v[n] : (sb - v[n]);
    }
};

// per-CTA scratch carved out of one buffer of doubles (shared memory on the device)
// horizons above this use the O(N^2) condensing (condense_long) in the generic path
#define FTMPC_LONG_N 20
#define FTMPC_RIC_WORK 1160      /* doubles of scratch riccati_factor needs (ftmpc_riccati.cuh) */
struct QpScratch {
    double* GL;                  // condense_long: sensitivity columns G_t[:, a], t = j+1 .. N, per column a = 6 j + ja
    const ftmpc_config* cg;      // configuration copy addressable per thread (global memory on the device)
    const double* tf_val;        // compact terminal-set rows (see StepIO), nullptr -> dense rows of cg->Af
    const int* tf_idx;
    double *E, *RS, *G, *T, *g, *ga, *taug, *cv, *hull;
    GiWork gi;
    double* dg;
    size_t total;       // doubles
};
FT_HD size_t qp_scratch_doubles(int N) {
    const WsLayout L = ws_layout(N);
    const size_t nv = L.nv, ne = nv + FTMPC_NE, n = L.n;
    size_t rs = nv * (nv + 1) / 2;
    if ((size_t)N * 338 > rs) rs = (size_t)N * 338;
    size_t gi_vec = ne /*xe*/ + L.m /*s*/ + (nv + 1) /*u*/ + nv /*d*/ + ne /*ze*/ + nv /*r*/ + 2 * nv /*cs*/ + (nv + 1) /*tmp*/ + nv /*sub*/ + nv /*dg*/ + 2;
    size_t tt = (size_t)FTMPC_NE * n;
    if (tt > gi_vec) gi_vec = tt;
    size_t ints = ((nv + 1) + L.m + (nv + 1) + 1) / 2 + 1;
    const size_t gl = (N > FTMPC_LONG_N) ? (size_t)FTMPC_NX * FTMPC_NU * N * (N + 1) / 2 : 0;     // condense_long: G_t columns
    size_t gw = (size_t)FTMPC_NX * nv;                 // G = d x_N / d U of the dense path; workspace of riccati_factor
    if (gw < FTMPC_RIC_WORK) gw = FTMPC_RIC_WORK;
    return ne * nv + rs + gw + gi_vec + 2 * nv /*g, ga*/ + 90 /*taug*/ + L.mc /*cv*/ + FTMPC_HULL_STRIDE + ints + 8 + gl;
}
FT_HD QpScratch qp_carve(double* buf, int N, const ftmpc_config* cg = nullptr) {
    const WsLayout L = ws_layout(N);
    const size_t nv = L.nv, ne = nv + FTMPC_NE, n = L.n;
    QpScratch s;
    s.cg = cg;
    s.tf_val = nullptr;
    s.tf_idx = nullptr;
    double* p = buf;
    s.E = p; p += ne * nv;
    size_t rs = nv * (nv + 1) / 2;
    if ((size_t)N * 338 > rs) rs = (size_t)N * 338;
    s.RS = p; p += rs;
    s.G = p; p += ((size_t)FTMPC_NX * nv < FTMPC_RIC_WORK) ? (size_t)FTMPC_RIC_WORK : (size_t)FTMPC_NX * nv;
    s.g = p; p += nv;
    s.ga = p; p += nv;
    s.taug = p; p += 90;
    s.cv = p; p += L.mc;
    s.hull = p; p += FTMPC_HULL_STRIDE;
    double* v = p;
    s.T = v;                      // condensing scratch aliases the active-set vectors
    s.gi.E = s.E; s.gi.Ui = s.RS;
    s.gi.xe = v; v += ne;
    s.gi.s = v; v += L.m;
    s.gi.u = v; v += nv + 1;
    s.gi.d = v; v += nv;
    s.gi.ze = v; v += ne;
    s.gi.r = v; v += nv;
    s.gi.cs = v; v += 2 * nv;
    s.gi.tmp = v; v += nv + 1;
    s.gi.sub = v; v += nv;
    s.dg = v; v += nv;
    s.gi.esign = v; v += 2;
    size_t gi_vec = (size_t)(v - p), tt = (size_t)FTMPC_NE * n;
    p += (tt > gi_vec) ? tt : gi_vec;
    int* ip = reinterpret_cast<int*>(p);
    s.gi.act = ip; ip += nv + 1;
    s.gi.pos = ip; ip += L.m;
    s.gi.itmp = ip; ip += nv + 1;
    s.total = qp_scratch_doubles(N);
    s.GL = (N > FTMPC_LONG_N) ? buf + s.total - (size_t)FTMPC_NX * FTMPC_NU * N * (N + 1) / 2 : nullptr;
    return s;
}

// ---- O(N^2) condensing for long horizons -------------------------------------------------------------------------
// Same H, g, ga, G_N as `condense` below, which accumulates the rank-13 stage updates G_t' M_t G_t over the whole live
// block of H at every stage: O(N^3) operations and read-modify-write traffic on H (96 M multiply-adds at N = 100, half the
// time of a long-horizon solve).  Here every COLUMN a = 6 j + ja of the sensitivities is independent:
//   forward   G_{j+1}[:, a] = B_j e_ja,   G_{t+1}[:, a] = A_t G_t[:, a]                      (stored, t = j+1 .. N)
//   backward  P_N = Ht G_N[:, a];   for t = N-1 .. j+1:   H[6t.., a] = B_t' P_{t+1} + W_ux,t G_t[:, a],
//                                                          P_t = M_t G_t[:, a] + A_t' P_{t+1};
//             H[6j.., a] = B_j' P_{j+1} + W_uu,j[:, ja]
// so each entry of H is written once, 17 M multiply-adds at N = 100, one thread per column, no barrier inside.  The
// gradient (and its augmented-Lagrangian twin) are two more costate recursions.
// M_t = diag(2Q, 0) + [theta sym(W_t) + Cq_t^GN on (omega, q)],  W_ux / W_uu = the input rows of the same stage matrix.
template <class Blk>
FT_HD void condense_long(Blk& blk, const ftmpc_config& cfg, const WsLayout& L, const QpScratch& s, const double* Jz,
                         const double* Wz, const double* X, const double* U, const double* xref, const double* gradV,
                         const double* hessV, double theta, double sigma, const double* lam_prev, const double* Cq) {
    const int N = L.N, n = L.n, ld = L.nv, tid = blk.tid(), nt = blk.nthreads();
    double* H = s.E;
    double* GL = s.GL;
    const double* Ah = s.hull;
    // terminal rows of the augmentation: 9x9 matrix + 9-vector (as in `condense`)
    for (int idx = tid; idx < 90; idx += nt) {
        double v = 0.0;
        if (sigma > 0.0) {
            const int kk = idx / 9, l = idx % 9;
            for (int i = 0; i < FTMPC_NF; ++i) {
                if (lam_prev[FTMPC_NH * N + i] > 0.0) {
                    const double a = cfg.Af[i * FTMPC_NE + l];
                    v += (idx < 81) ? cfg.Af[i * FTMPC_NE + kk] * a : s.cv[FTMPC_NH * N + i] * a;
                }
            }
        }
        s.taug[idx] = sigma * v;
    }
    if (tid == 0) { s.g[n] = 0.0; s.ga[n] = 0.0; }
    for (int r = tid; r < FTMPC_NX; r += nt) s.G[r * ld + n] = 0.0;
    blk.sync();
    auto gl_base = [&](int j, int ja) { return GL + (size_t)FTMPC_NX * ((size_t)FTMPC_NU * ((size_t)j * N - (size_t)j * (j - 1) / 2) + (size_t)ja * (N - j)); };
    auto symw = [&](const double* wz, int r, int c) { return (theta != 0.0) ? 0.5 * (wz[r * 13 + c] + wz[c * 13 + r]) : 0.0; };   // theta = 0: W_t was not computed
    // A_t' p  (A_t = [[I, dt I, *], [0, I, *], [0, 0, *]] with the (omega, q) columns in jz[l*13 + r], l < 7)
    auto at_times = [&](const double* jz, const double* p, double* out) {
        for (int c = 0; c < 3; ++c) out[c] = p[c];
        for (int c = 3; c < 6; ++c) out[c] = p[c] + cfg.dt * p[c - 3];
        for (int l = 0; l < 7; ++l) {
            double v = 0.0;
            for (int r = 0; r < FTMPC_NX; ++r) v += jz[l * 13 + r] * p[r];
            out[6 + l] = v;
        }
    };
    for (int a = tid; a < n + 2; a += nt) {
        if (a >= n) {
            // ---- costates of the gradient (a == n) and of the terminal augmentation (a == n + 1)
            const bool aug = (a == n + 1);
            double p[FTMPC_NX], q[FTMPC_NX];
            for (int r = 0; r < FTMPC_NX; ++r) p[r] = 0.0;
            for (int kk = 0; kk < FTMPC_NE; ++kk) p[kk] = aug ? s.taug[81 + kk] : gradV[kk];
            for (int t = N - 1; t >= 0; --t) {
                const double* jz = Jz + (size_t)t * 169;
                for (int i = 0; i < FTMPC_NU; ++i) {
                    double v = 0.0;
                    for (int r = 0; r < FTMPC_NX; ++r) v += jz[(7 + i) * 13 + r] * p[r];
                    if (aug) {
                        double av = 0.0;
                        if (sigma > 0.0)
                            for (int k = 0; k < FTMPC_NH; ++k)
                                if (lam_prev[t * FTMPC_NH + k] > 0.0) av += s.cv[t * FTMPC_NH + k] * Ah[k * FTMPC_NU + i];
                        s.T[t * FTMPC_NU + i] = v + sigma * av;              // ga - g, combined below
                    } else {
                        s.g[t * FTMPC_NU + i] = v + 2.0 * cfg.R[i] * (U[t * FTMPC_NU + i] - (Cq ? Cq[(size_t)t * FTMPC_CQ + 32 + i] : 0.0));
                    }
                }
                at_times(jz, p, q);
                for (int r = 0; r < FTMPC_NX; ++r) p[r] = q[r];
                if (!aug && t > 0) {
                    for (int kk = 0; kk < FTMPC_NE; ++kk) p[kk] += 2.0 * cfg.Q[kk] * (X[t * FTMPC_NX + kk] - xref[t * FTMPC_NE + kk]);
                    if (Cq) for (int l = 0; l < 4; ++l) p[9 + l] += Cq[(size_t)t * FTMPC_CQ + l];
                }
            }
            continue;
        }
        const int j = a / FTMPC_NU, ja = a - j * FTMPC_NU;
        double* gcol = gl_base(j, ja);                      // G_t[:, a] at gcol + 13 (t - j - 1)
        // ---- forward
        double g[FTMPC_NX], gn[FTMPC_NX];
        {
            const double* jz = Jz + (size_t)j * 169;
            for (int r = 0; r < FTMPC_NX; ++r) { g[r] = jz[(7 + ja) * 13 + r]; gcol[r] = g[r]; }
        }
        for (int t = j + 1; t < N; ++t) {                   // G_{t+1} = A_t G_t
            const double* jz = Jz + (size_t)t * 169;
            for (int r = 0; r < FTMPC_NX; ++r) {
                double v = (r < 3) ? g[r] + cfg.dt * g[r + 3] : ((r < 6) ? g[r] : 0.0);
                for (int l = 0; l < 7; ++l) v += jz[l * 13 + r] * g[6 + l];
                gn[r] = v;
            }
            double* o = gcol + (size_t)FTMPC_NX * (t - j);
            for (int r = 0; r < FTMPC_NX; ++r) { g[r] = gn[r]; o[r] = gn[r]; }
        }
        for (int r = 0; r < FTMPC_NX; ++r) s.G[r * ld + a] = g[r];     // G_N[:, a]
        // ---- backward
        double p[FTMPC_NX], q[FTMPC_NX];
        for (int kk = 0; kk < FTMPC_NE; ++kk) {
            double v = 0.0;
            for (int l = 0; l < FTMPC_NE; ++l) {
                const double q0 = cfg.term_quad[kk * FTMPC_NE + l];
                v += (q0 + theta * (hessV[kk * FTMPC_NE + l] - q0) + s.taug[kk * FTMPC_NE + l]) * g[l];
            }
            p[kk] = v;
        }
        for (int r = FTMPC_NE; r < FTMPC_NX; ++r) p[r] = 0.0;
        for (int t = N - 1; t > j; --t) {
            const double* jz = Jz + (size_t)t * 169;
            const double* wz = Wz + (size_t)t * 169;
            const double* gt = gcol + (size_t)FTMPC_NX * (t - j - 1);
            const double* cq = Cq ? Cq + (size_t)t * FTMPC_CQ : nullptr;
            for (int i = 0; i < FTMPC_NU; ++i) {            // H[6t + i][a] = B_t' P_{t+1} + W_ux,t G_t[:, a]
                double v = 0.0, wv = 0.0;
                for (int r = 0; r < FTMPC_NX; ++r) v += jz[(7 + i) * 13 + r] * p[r];
                for (int l = 0; l < 7; ++l) wv += symw(wz, 7 + i, l) * gt[6 + l];
                wv *= theta;
                if (cq && i < 3) for (int l = 0; l < 4; ++l) wv += cq[4 + i * 4 + l] * gt[9 + l];
                H[(size_t)(FTMPC_NU * t + i) * ld + a] = v + wv;
            }
            at_times(jz, p, q);                              // P_t = M_t G_t + A_t' P_{t+1}
            for (int kk = 0; kk < FTMPC_NE; ++kk) q[kk] += 2.0 * cfg.Q[kk] * gt[kk];
            for (int kk = 0; kk < 7; ++kk) {
                double v = 0.0;
                for (int l = 0; l < 7; ++l) v += symw(wz, kk, l) * gt[6 + l];
                v *= theta;
                if (cq && kk >= 3) for (int l = 0; l < 4; ++l) v += cq[16 + (kk - 3) * 4 + l] * gt[9 + l];
                q[6 + kk] += v;
            }
            for (int r = 0; r < FTMPC_NX; ++r) p[r] = q[r];
        }
        {
            const double* jz = Jz + (size_t)j * 169;
            const double* wz = Wz + (size_t)j * 169;
            for (int i = ja; i < FTMPC_NU; ++i) {           // diagonal block, lower triangle: rows i >= ja
                double v = 0.0;
                for (int r = 0; r < FTMPC_NX; ++r) v += jz[(7 + i) * 13 + r] * p[r];
                v += theta * symw(wz, 7 + i, 7 + ja);
                if (i == ja) v += 2.0 * cfg.R[i];
                if (sigma > 0.0) {
                    double av = 0.0;
                    for (int k = 0; k < FTMPC_NH; ++k)
                        if (lam_prev[j * FTMPC_NH + k] > 0.0) av += Ah[k * FTMPC_NU + i] * Ah[k * FTMPC_NU + ja];
                    v += sigma * av;
                }
                H[(size_t)(FTMPC_NU * j + i) * ld + a] = v;
            }
        }
    }
    blk.sync();
    for (int a = tid; a < n; a += nt) s.ga[a] = s.g[a] + s.T[a];
    blk.sync();
}

// condensed Hessian (lower triangle of the n x n block of E, ld = nv) and gradient at blend theta
template <class Blk>
FT_HD void condense(Blk& blk, const ftmpc_config& cfg, const WsLayout& L, const QpScratch& s, const double* Jz,
                    const double* Wz, const double* X, const double* U, const double* xref, const double* gradV,
                    const double* hessV, double theta, double sigma, const double* lam_prev, const double* Cq = nullptr) {
    // Cq: input-cost coupling records of an accelerating reference (FTMPC_CQ) or nullptr
    const int N = L.N, n = L.n, ld = L.nv, tid = blk.tid(), nt = blk.nthreads();
    if (s.GL && (cfg.qp_method & 8) == 0) {            // long horizons: O(N^2) column-wise condensing (bit 3 keeps the O(N^3) form)
        condense_long(blk, cfg, L, s, Jz, Wz, X, U, xref, gradV, hessV, theta, sigma, lam_prev, Cq);
        return;
    }
    double* H = s.E;
    double* G = s.G;          // 13 x ld, current d x_t / d U
    double* T = s.T;
    // Augmented-Lagrangian convexification on the predicted active set A (rows with lam_prev > 0):
    //   H += sigma * sum_A a_i a_i',  ga = g + sigma * sum_A c_i a_i.   If every row of A is active in the QP
    //   solution this leaves the QP solution and its multipliers unchanged, but makes H positive definite
    //   whenever the second-order sufficient condition holds (the exact Hessian alone need not be).
    const double* Ah = s.hull;
    for (int i = tid; i < FTMPC_NX * ld; i += nt) G[i] = 0.0;
    for (int i = tid; i < L.nv; i += nt) { s.g[i] = 0.0; s.ga[i] = 0.0; }
    for (int idx = tid; idx < 90; idx += nt) {       // terminal rows: 9x9 matrix + 9-vector
        double v = 0.0;
        if (sigma > 0.0) {
            const int kk = idx / 9, l = idx % 9;
            for (int i = 0; i < FTMPC_NF; ++i) {
                if (lam_prev[FTMPC_NH * N + i] > 0.0) {
                    const double a = cfg.Af[i * FTMPC_NE + l];
                    v += (idx < 81) ? cfg.Af[i * FTMPC_NE + kk] * a : s.cv[FTMPC_NH * N + i] * a;
                }
            }
        }
        s.taug[idx] = sigma * v;
    }
    blk.sync();
    for (int t = 0; t < N; ++t) {
        const double* jz = Jz + (size_t)t * 169;     // [col][row]
        const double* wz = Wz + (size_t)t * 169;     // [col][row], symmetrised on use
        const int nc = FTMPC_NU * t;                 // columns of G that are live
        // T[k][b] = theta * sum_l Wz[k][l] G[6+l][b],  k,l in (w,q)
        for (int idx = tid; idx < 7 * nc; idx += nt) {
            const int kk = idx / nc, b = idx % nc;
            double v = 0.0;
            if (theta != 0.0) for (int l = 0; l < 7; ++l) v += 0.5 * (wz[kk * 13 + l] + wz[l * 13 + kk]) * G[(6 + l) * ld + b];
            v *= theta;
            if (Cq && kk >= 3)
                for (int l = 0; l < 4; ++l) v += Cq[(size_t)t * FTMPC_CQ + 16 + (kk - 3) * 4 + l] * G[(9 + l) * ld + b];
            T[kk * n + b] = v;
        }
        blk.sync();
        // accumulate into the live block  H[a][b], a,b < nc
        for (int bi = 0; bi < t; ++bi) {
            for (int idx = tid; idx < 36 * (bi + 1); idx += nt) {
                const int bj = idx / 36, rem = idx % 36;
                const int a = 6 * bi + rem / 6, b = 6 * bj + rem % 6;
                if (b > a) continue;
                double v = 0.0;
                for (int kk = 0; kk < FTMPC_NE; ++kk) v += 2.0 * cfg.Q[kk] * G[kk * ld + a] * G[kk * ld + b];
                for (int kk = 0; kk < 7; ++kk) v += G[(6 + kk) * ld + a] * T[kk * n + b];
                H[(size_t)a * ld + b] += v;
            }
        }
        // new block row t:  H[6t+j][b]
        for (int idx = tid; idx < 6 * (nc + 6); idx += nt) {
            const int j = idx / (nc + 6), b = idx % (nc + 6);
            double v = 0.0;
            if (b < nc) {
                if (theta != 0.0) for (int l = 0; l < 7; ++l) v += 0.5 * (wz[(7 + j) * 13 + l] + wz[l * 13 + 7 + j]) * G[(6 + l) * ld + b];
                v *= theta;
                if (Cq && j < 3)
                    for (int l = 0; l < 4; ++l) v += Cq[(size_t)t * FTMPC_CQ + 4 + j * 4 + l] * G[(9 + l) * ld + b];
            } else {
                const int j2 = b - nc;
                v = (theta != 0.0) ? theta * 0.5 * (wz[(7 + j) * 13 + 7 + j2] + wz[(7 + j2) * 13 + 7 + j]) : 0.0;
                if (j2 == j) v += 2.0 * cfg.R[j];
                if (sigma > 0.0) {
                    double av = 0.0;
                    for (int i = 0; i < FTMPC_NH; ++i)
                        if (lam_prev[t * FTMPC_NH + i] > 0.0) av += Ah[i * FTMPC_NU + j] * Ah[i * FTMPC_NU + j2];
                    v += sigma * av;
                }
            }
            H[(size_t)(nc + j) * ld + b] = v;
        }
        // gradient
        for (int a = tid; a < nc + 6; a += nt) {
            if (a < nc) {
                double v = 0.0;
                for (int kk = 0; kk < FTMPC_NE; ++kk)
                    v += 2.0 * cfg.Q[kk] * (X[t * FTMPC_NX + kk] - xref[t * FTMPC_NE + kk]) * G[kk * ld + a];
                if (Cq) for (int l = 0; l < 4; ++l) v += Cq[(size_t)t * FTMPC_CQ + l] * G[(9 + l) * ld + a];
                s.g[a] += v;
            } else {
                s.g[a] = 2.0 * cfg.R[a - nc] * (U[a] - (Cq ? Cq[(size_t)t * FTMPC_CQ + 32 + a - nc] : 0.0));
                if (sigma > 0.0) {
                    double av = 0.0;
                    for (int i = 0; i < FTMPC_NH; ++i)
                        if (lam_prev[t * FTMPC_NH + i] > 0.0) av += s.cv[t * FTMPC_NH + i] * Ah[i * FTMPC_NU + a - nc];
                    s.ga[a] = sigma * av;
                }
            }
        }
        blk.sync();
        // G_{t+1} = A_t G_t + B_t E_t   (thread per column)
        for (int a = tid; a < nc + 6; a += nt) {
            double o[13], nw[13];
            if (a < nc) {
                for (int r = 0; r < 13; ++r) o[r] = G[r * ld + a];
                for (int r = 0; r < 13; ++r) {
                    double v = (r < 3) ? o[r] + cfg.dt * o[r + 3] : ((r < 6) ? o[r] : 0.0);
                    for (int l = 0; l < 7; ++l) v += jz[l * 13 + r] * o[6 + l];
                    nw[r] = v;
                }
            } else {
                for (int r = 0; r < 13; ++r) nw[r] = jz[(7 + a - nc) * 13 + r];
            }
            for (int r = 0; r < 13; ++r) G[r * ld + a] = nw[r];
        }
        blk.sync();
    }
    // terminal cost:  H += G_N' Ht G_N,  g += G_N' grad V      Ht = term_quad + theta (hessV - term_quad)
    for (int idx = tid; idx < FTMPC_NE * n; idx += nt) {
        const int kk = idx / n, b = idx % n;
        double v = 0.0;
        for (int l = 0; l < FTMPC_NE; ++l) {
            const double q0 = cfg.term_quad[kk * FTMPC_NE + l];
            v += (q0 + theta * (hessV[kk * FTMPC_NE + l] - q0) + s.taug[kk * FTMPC_NE + l]) * G[l * ld + b];
        }
        T[kk * n + b] = v;
    }
    blk.sync();
    for (int bi = 0; bi < N; ++bi) {
        for (int idx = tid; idx < 36 * (bi + 1); idx += nt) {
            const int bj = idx / 36, rem = idx % 36;
            const int a = 6 * bi + rem / 6, b = 6 * bj + rem % 6;
            if (b > a) continue;
            double v = 0.0;
            for (int kk = 0; kk < FTMPC_NE; ++kk) v += G[kk * ld + a] * T[kk * n + b];
            H[(size_t)a * ld + b] += v;
        }
    }
    for (int a = tid; a < n; a += nt) {
        double v = 0.0, va = 0.0;
        for (int kk = 0; kk < FTMPC_NE; ++kk) {
            v += gradV[kk] * G[kk * ld + a];
            va += s.taug[81 + kk] * G[kk * ld + a];
        }
        s.g[a] += v;
        s.ga[a] += s.g[a] + va;
    }
    blk.sync();
}

#if defined(__CUDACC__)
// ---- linearisation, CUDA-block specialisation ------------------------------------------------------------
// Same Jz / Wz / Mu as phase_lin above.  The (stage, column) sweep of rk4_column is split so that nothing has to be
// held in registers across the costate recursion:
//   pass 1  thread (t, c), c = one of the 10 columns in (w, q, tau): forward tangent through the RK4 stages; parks the
//           stage tangents (and, for c = 0, the nominal stage data) in shared memory; the three force columns of the
//           Jacobian come in closed form from the nominal stage rotations (their (w, q) tangents vanish);
//   costate one warp, shared-memory Jacobians, no block barrier inside the recursion;
//   pass 2  thread (t, c): reverse sweep from the parked data -> Hessian column; the force rows/columns of W_t are
//           the transposes of what the other columns produce (the dynamics are linear in F).
struct LinScratch {
    double *nom, *wr, *tang, *Jz, *mu, *X;
};
__host__ __device__ inline size_t lin_scratch_doubles(int N) {
    return (size_t)N * 40 + (size_t)N * 6 + (size_t)28 * 10 * N + (size_t)N * 169 + 2 * (size_t)(N + 1) * FTMPC_NX + 8;
}
__device__ __forceinline__ LinScratch lin_carve(double* buf, int N) {
    LinScratch s;
    double* p = buf;
    s.nom = p; p += (size_t)N * 40;
    s.wr = p; p += (size_t)N * 6;
    s.tang = p; p += (size_t)28 * 10 * N;
    s.Jz = p; p += (size_t)N * 169;
    s.mu = p; p += (size_t)(N + 1) * FTMPC_NX;
    s.X = p; p += (size_t)(N + 1) * FTMPC_NX;
    return s;
}
__device__ __noinline__ void lin_forward_task(const DynConsts& k, const double* x, const double* Wr, int col, double* jac_col,
                                              double* nom, double* tang, int tstride) {
    rk4_col_forward(k, x, Wr, col, jac_col, nom, tang, tstride);
}
__device__ __noinline__ void lin_reverse_task(const DynConsts& k, const double* nom, const double* tang, int tstride, int col,
                                              const double* lam, double* hess_col) {
    rk4_col_reverse(k, nom, tang, tstride, col, lam, hess_col);
}

// LinPlace: where the phase keeps its own scratch and where it leaves the stage Jacobians / Hessians for the condensing
// that follows (JzS / WzS = nullptr: only the global backing copies are written)
struct LinPlace {
    double* work;      // nom, wr, tang  (N * 326 doubles)
    double* tail;      // mu, X          (2 (N + 1) * 13 doubles)
    double* JzS;       // N * 169, always in shared memory (the costate recursion reads it)
    double* WzS;       // N * 169 or nullptr
};
__device__ __forceinline__ void phase_lin(CudaBlock& blk, const ftmpc_config& cfg, const WsLayout& L, const StepIO& io,
                                          int inst, int slot, const LinPlace& place,
                                          bool skip_unused_w = false /* the caller's QP phase starts from qp_start() */) {
    double* w = ws_slot(io, L, slot);
    const double* sc = w + L.oSc;
    if (sc[SC_STATUS] != FTMPC_ST_RUNNING) return;
    const int N = L.N, tid = blk.tid(), nt = blk.nthreads();
    const DynConsts k = dyn_consts(cfg);
    const double* xref = io.xref + (size_t)inst * io.xref_stride;
    const double* uref = io.uref ? io.uref + (size_t)inst * io.uref_stride : nullptr;
    const double* U = w + L.oU;
    const double* X = w + L.oX;
    double* Jz = w + L.oJz;
    double* Wz = w + L.oWz;
    double* Mu = w + L.oMu;
    const double* lam = w + L.oLam;
    LinScratch s;
    s.nom = place.work;
    s.wr = s.nom + (size_t)N * 40;
    s.tang = s.wr + (size_t)N * 6;
    s.mu = place.tail;
    s.X = s.mu + (size_t)(N + 1) * FTMPC_NX;
    s.Jz = place.JzS;
    double* WzS = place.WzS;
    const int ntask = 10 * N;
    // stage states and wrenches -> shared memory
    for (int i = tid; i < (N + 1) * FTMPC_NX; i += nt) s.X[i] = X[i];
    blk.sync();
    double* Cq = w + L.oCq;
    for (int t = tid; t < N; t += nt) {
        stage_wrench(cfg, U + t * FTMPC_NU, nullptr, s.X + t * FTMPC_NX + 9, s.wr + t * 6);      // U holds u~ (FTMPC_CQ)
        if (uref) stage_cost_coupling(cfg, U + t * FTMPC_NU, uref + t * FTMPC_NU, s.X + t * FTMPC_NX + 9, Cq + (size_t)t * FTMPC_CQ);
    }
    blk.sync();
    // pass 1: first-order columns
    for (int it = tid; it < ntask; it += nt) {
        const int t = it / 10, c = it - t * 10;
        const int col = (c < 7) ? c : c + 3;
        double jc[13];
        lin_forward_task(k, s.X + t * FTMPC_NX, s.wr + t * 6, col, jc, (c == 0) ? s.nom + (size_t)t * 40 : nullptr,
                         s.tang + it, ntask);
        double* js = s.Jz + ((size_t)t * 13 + col) * 13;
        double* jg = Jz + ((size_t)t * 13 + col) * 13;
#pragma unroll
        for (int i = 0; i < 13; ++i) { js[i] = jc[i]; jg[i] = jc[i]; }
    }
    // running-cost gradient 2Q (x_t - xr_t) of every stage, parked in the (not yet used) costate slots
    for (int i = tid; i < N * FTMPC_NX; i += nt) {
        const int t = i / FTMPC_NX, r = i - t * FTMPC_NX;
        s.mu[i] = (r < FTMPC_NE) ? 2.0 * cfg.Q[r] * (s.X[i] - xref[t * FTMPC_NE + r])
                                 : (uref ? Cq[(size_t)t * FTMPC_CQ + r - FTMPC_NE] : 0.0);
    }
    // terminal costate  mu_N = [grad V_f + A_f' lam_term ; 0]   (threads from the other end of the block)
    for (int i = nt - 1 - tid; i < FTMPC_NX; i += nt) {
        double v = 0.0;
        if (i < FTMPC_NE) {
            v = w[L.oGV + i];
            const double* Af = io.cfg_g->Af;
            for (int r = 0; r < FTMPC_NF; ++r) v += Af[r * FTMPC_NE + i] * lam[FTMPC_NH * N + r];
        }
        s.mu[N * FTMPC_NX + i] = v;
    }
    blk.sync();
    for (int t = tid; t < N; t += nt) {                       // force columns need the parked nominal stages
        double jf[39];
        rk4_force_columns(k, s.nom + (size_t)t * 40, jf);
        for (int i = 0; i < 39; ++i) { s.Jz[((size_t)t * 13 + 7) * 13 + i] = jf[i]; Jz[((size_t)t * 13 + 7) * 13 + i] = jf[i]; }
    }
    blk.sync();
    blk.mark(PH_LIN_JAC);
    // a QP that starts from the Gauss-Newton model never reads W_t (qp_start): no costates, no Hessian sweeps
    if (skip_unused_w && qp_start(cfg, sc).theta == 0.0) { blk.mark(PH_LIN); return; }
    // costates  mu_t = [2Q e_t ; 0] + A_t' mu_{t+1}   (one warp, 13 lanes)
    if (tid < 32) {
        for (int t = N - 1; t >= 1; --t) {
            const double* mn = s.mu + (t + 1) * FTMPC_NX;
            if (tid < FTMPC_NX) {
                const int i = tid;
                double v = s.mu[t * FTMPC_NX + i];
                if (i < 3) v += mn[i];
                else if (i < 6) v += k.dt * mn[i - 3] + mn[i];
                else {
                    const double* col = s.Jz + ((size_t)t * 13 + (i - 6)) * 13;
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int r = 0; r < 12; r += 2) { a0 += col[r] * mn[r]; a1 += col[r + 1] * mn[r + 1]; }
                    v += a0 + a1 + col[12] * mn[12];
                }
                s.mu[t * FTMPC_NX + i] = v;
            }
            __syncwarp();
        }
    }
    blk.sync();
    for (int i = tid; i < (N + 1) * FTMPC_NX; i += nt) Mu[i] = s.mu[i];
    blk.mark(PH_LIN_MU);
    // pass 2: second-order columns
    for (int it = tid; it < ntask; it += nt) {
        const int t = it / 10, c = it - t * 10;
        const int col = (c < 7) ? c : c + 3;
        double hc[13];
        lin_reverse_task(k, s.nom + (size_t)t * 40, s.tang + it, ntask, col, s.mu + (t + 1) * FTMPC_NX, hc);
        double* wg = Wz + (size_t)t * 169;
        double* wsm = WzS ? WzS + (size_t)t * 169 : wg;
#pragma unroll
        for (int i = 0; i < 13; ++i) { wg[col * 13 + i] = hc[i]; wsm[col * 13 + i] = hc[i]; }
#pragma unroll
        for (int j = 0; j < 3; ++j) { wg[(7 + j) * 13 + col] = hc[7 + j]; wsm[(7 + j) * 13 + col] = hc[7 + j]; }   // W[F_j][col] = W[col][F_j]
        if (c == 0) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int l = 0; l < 3; ++l) { wg[(7 + j) * 13 + 7 + l] = 0.0; wsm[(7 + j) * 13 + 7 + l] = 0.0; }   // the dynamics are linear in F
        }
    }
    blk.sync();
    if (uref) {                                     // exact remainder of the input cost's (q,q) Hessian
        for (int idx = tid; idx < N * 16; idx += nt) {
            const int t = idx >> 4, a = (idx >> 2) & 3, b = idx & 3;
            const double v = Cq[(size_t)t * FTMPC_CQ + 40 + a * 4 + b];
            const size_t o = (size_t)t * 169 + (3 + a) * 13 + 3 + b;
            Wz[o] += v;
            if (WzS) WzS[o] += v;
        }
        blk.sync();
    }
    blk.mark(PH_LIN);
}
#endif  // __CUDACC__

#if defined(__CUDACC__)
// scratch placement of the linearisation inside k_solve (one CTA per SM): own layout, Jz / Wz optionally left in the RS
// region of the QP scratch
__device__ __forceinline__ LinPlace lin_place_v1(double* scratch, int N, bool stage_rs) {
    const LinScratch l = lin_carve(scratch, N);
    LinPlace p;
    p.work = l.nom; p.tail = l.mu; p.JzS = l.Jz; p.WzS = nullptr;
    if (stage_rs) {
        const QpScratch q = qp_carve(scratch, N);
        p.JzS = q.RS;
        p.WzS = q.RS + (size_t)N * 169;
    }
    return p;
}
#endif

#if defined(__CUDACC__)
// ---- condensing, CUDA-block specialisation -----------------------------------------------------------------
// Same H, g, ga, G_N as the generic routine above, organised so that nothing is read-modify-written in
// shared memory:
//   * thread a < 6N owns COLUMN a of the sensitivity G_t = d x_t / d U in registers and advances it stage by
//     stage (G_{t+1}[:,a] = A_t G_t[:,a] is thread-local); per stage it publishes G_t[:,a], (M_t G_t)[:,a] and
//     (theta W_ux G_t)[:,a] to a double-buffered panel and accumulates its gradient entry;
//   * thread (bi >= bj) owns the 6x6 BLOCK (bi, bj) of H in 36 registers and adds  G_t[:,bi]' (M_t G_t)[:,bj]
//     (rank 13) at every stage t > bi; H is written once at the end.
// One barrier per stage.  The panels live in the E region (free until H is stored).
// When `hb` is given and the register-tiled path is taken, H is NOT stored: the 6x6 blocks stay in hb->acc (owner
// thread (hb->bi, hb->bj)) for chol_inv_blocks, hb->fast = true.
struct HBlocks {
    double acc[6][6];
    int bi, bj;
    bool fast;
};
__device__ __forceinline__ bool condense_fast_path(const WsLayout& L, int nt) {
    const int N = L.N, ld = L.nv;
    const int ldp = 7 * N + 1;
    const size_t panel_doubles = (size_t)64 * ldp + (size_t)N * FTMPC_NE + 90 + (size_t)L.mc + 90 + (size_t)N * 10;
    return !(N * (N + 1) / 2 > nt || L.n > nt || panel_doubles > (size_t)(ld + FTMPC_NE) * ld);
}
__device__ __forceinline__ void condense(CudaBlock& blk, const ftmpc_config& cfg, const WsLayout& L, const QpScratch& s,
                                         const double* Jz, const double* Wz_in, const double* X, const double* U,
                                         const double* xref, const double* gradV, const double* hessV, double theta,
                                         double sigma, const double* lam_prev_g, HBlocks* hb = nullptr, const double* Cq = nullptr) {
    const int N = L.N, n = L.n, ld = L.nv, tid = blk.tid(), nt = blk.nthreads();
    const int nblk = N * (N + 1) / 2;
    if (hb) hb->fast = false;
    const int ldp = 7 * N + 1;                     // panel row length: block column b starts at 7 b (6 + 1 pad -> lanes of
                                                   // neighbouring blocks are an odd number of doubles apart: no bank conflicts)
    if (!condense_fast_path(L, nt)) {              // very short / long horizons: generic path
        condense<CudaBlock>(blk, cfg, L, s, Jz, Wz_in, X, U, xref, gradV, hessV, theta, sigma, lam_prev_g, Cq);
        return;
    }
    double* Wp = const_cast<double*>(Wz_in);       // scratch copy owned by the caller: symmetrised / scaled in place
    double* buf0 = s.E;                            // panel b: rows 0-12 G_t, 13-25 M_t G_t, 26-31 theta W_ux G_t
    double* qe = s.E + (size_t)64 * ldp;           // [N][9]   2 Q (x_t - xr_t)
    double* Ht = qe + (size_t)N * FTMPC_NE;        // [9][9]   terminal Hessian model (+ augmentation)
    double* tgv = Ht + 81;                         // [9]      augmentation of the terminal gradient
    double* lam_prev = tgv + 9;                    // [mc]     shared-memory copy of the previous multipliers: the tests
                                                   //          `lam_prev[i] > 0` sit in serial loops, one L2 round trip each otherwise
    double* hv = lam_prev + L.mc;                  // [81 + 9] hessV, gradV
    double* cqs = hv + 90;                         // [N][10]  accelerating references (FTMPC_CQ): gq (4), rho (6)
    const double* Ah = s.hull;
    // ---- pre-pass: W' = theta * sym(W) (+ 2Q on the omega diagonal), qe, Ht
    for (int idx = tid; idx < N * 169; idx += nt) {
        const int t = idx / 169, e = idx - t * 169;
        const int c = e / 13, r = e - c * 13;
        if (c < r) continue;                       // the pair (c >= r) is written by one thread
        double* wz = Wp + (size_t)t * 169;
        double v = (theta != 0.0) ? theta * 0.5 * (wz[c * 13 + r] + wz[r * 13 + c]) : 0.0;
        if (c == r && c < 3) v += 2.0 * cfg.Q[6 + c];
        wz[c * 13 + r] = v;
        wz[r * 13 + c] = v;
    }
    for (int idx = tid; idx < N * FTMPC_NE; idx += nt) {
        const int t = idx / FTMPC_NE, kk = idx - t * FTMPC_NE;
        qe[idx] = 2.0 * cfg.Q[kk] * (X[t * FTMPC_NX + kk] - xref[t * FTMPC_NE + kk]);
    }
    if (sigma > 0.0) for (int i = tid; i < L.mc; i += nt) lam_prev[i] = lam_prev_g[i];
    for (int i = tid; i < 90; i += nt) hv[i] = (i < 81) ? hessV[i] : gradV[i - 81];
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
            if (s.tf_val) {                        // <= 2 non-zeros per row, tables in shared memory
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
                const double* Af = s.cg->Af;
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
            const double q0 = s.cg->term_quad[idx];
            Ht[idx] = q0 + theta * (hv[idx] - q0) + v;
        } else {
            tgv[l] = v;
        }
    }
    if (tid == 0) { s.g[n] = 0.0; s.ga[n] = 0.0; }
    for (int r = tid; r < FTMPC_NX; r += nt) s.G[r * ld + n] = 0.0;
    // ---- roles
    // column role: taken by the LAST n threads of the block -- their block role (large bi) has the least work, and
    // the column phase of stage t + 1 runs in the same barrier interval as the block phase of stage t
    // (reversed: the first columns -- the only ones alive in the early stages -- sit in the LAST warp, which shares its
    // scheduler with block warp 3, idle until stage 14; warp 0 owns the blocks that are busy from stage 1 on)
    const int a = (nt - 1 - tid < n) ? nt - 1 - tid : -1;   // column index (0 <= a < n) or negative
    const int ta = (a >= 0) ? a / FTMPC_NU : 0, ja = a - ta * FTMPC_NU;
    const int pa_ = 7 * ta + ja;                   // column a inside a panel
    int bi = -1, bj = 0;                           // block role (tid < nblk)
    if (tid < nblk) {
        bi = (int)((sqrt(8.0 * tid + 1.0) - 1.0) * 0.5);
        while ((bi + 1) * (bi + 2) / 2 <= tid) ++bi;
        while (bi * (bi + 1) / 2 > tid) --bi;
        bj = tid - bi * (bi + 1) / 2;
    }
    double g[FTMPC_NX], gs = 0.0, gaug = 0.0;
    double acc[6][6];
#pragma unroll
    for (int i = 0; i < FTMPC_NX; ++i) g[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = 0.0;
    blk.sync();
    blk.mark(PH_COND_PRE);
    // software pipeline: interval tt runs the column phase of stage tt + 1 (writes panel (tt + 1) & 1) and the block
    // phase of stage tt (reads panel tt & 1, published one barrier earlier); one barrier per stage
    for (int tt = -1; tt <= N; ++tt) {
        // ---------------- column phase of stage t = tt + 1
        if (a >= 0 && tt < N) {
            const int t = tt + 1;
            double* buf = buf0 + (size_t)(t & 1) * 32 * ldp;
            if (ta < t) {
                if (t < N) {
                    const double* jz = Jz + (size_t)t * 169;
                    const double* wp = Wp + (size_t)t * 169;
                    double tp[FTMPC_NX], sx[FTMPC_NU];
#pragma unroll
                    for (int r = 0; r < 6; ++r) tp[r] = 2.0 * cfg.Q[r] * g[r];
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        double v = 0.0;
#pragma unroll
                        for (int l = 0; l < 7; ++l) v += wp[k * 13 + l] * g[6 + l];
                        tp[6 + k] = v;
                    }
#pragma unroll
                    for (int i = 0; i < FTMPC_NU; ++i) {
                        double v = 0.0;
#pragma unroll
                        for (int l = 0; l < 7; ++l) v += wp[(7 + i) * 13 + l] * g[6 + l];
                        sx[i] = v;
                    }
#pragma unroll
                    for (int r = 0; r < FTMPC_NX; ++r) { buf[r * ldp + pa_] = g[r]; buf[(13 + r) * ldp + pa_] = tp[r]; }
#pragma unroll
                    for (int i = 0; i < FTMPC_NU; ++i) buf[(26 + i) * ldp + pa_] = sx[i];
#pragma unroll
                    for (int kk = 0; kk < FTMPC_NE; ++kk) gs += qe[t * FTMPC_NE + kk] * g[kk];
                    if (Cq) {
#pragma unroll
                        for (int l = 0; l < 4; ++l) gs += cqs[t * 10 + l] * g[9 + l];
                    }
                    double gn[FTMPC_NX];
#pragma unroll
                    for (int r = 0; r < FTMPC_NX; ++r) {
                        double v = (r < 3) ? g[r] + cfg.dt * g[r + 3] : ((r < 6) ? g[r] : 0.0);
#pragma unroll
                        for (int l = 0; l < 7; ++l) v += jz[l * 13 + r] * g[6 + l];
                        gn[r] = v;
                    }
#pragma unroll
                    for (int r = 0; r < FTMPC_NX; ++r) g[r] = gn[r];
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
#pragma unroll
                    for (int r = 0; r < FTMPC_NX; ++r) s.G[r * ld + a] = g[r];
                }
            } else if (ta == t) {
                // birth of column a = 6t + ja:  G_{t+1}[:, a] = B_t e_ja,  g_a = 2 R u
                const double* jz = Jz + (size_t)t * 169;
#pragma unroll
                for (int r = 0; r < FTMPC_NX; ++r) g[r] = jz[(7 + ja) * 13 + r];
                gs = 2.0 * cfg.R[ja] * (U[a] - (Cq ? cqs[t * 10 + 4 + ja] : 0.0));
                if (sigma > 0.0) {
                    double av = 0.0;
                    for (int i = 0; i < FTMPC_NH; ++i)
                        if (lam_prev[t * FTMPC_NH + i] > 0.0) av += s.cv[t * FTMPC_NH + i] * Ah[i * FTMPC_NU + ja];
                    gaug = sigma * av;
                }
            }
        }
        // ---------------- block phase of stage t = tt
        if (bi >= 0 && tt >= 0) {
            const int t = tt;
            const double* buf = buf0 + (size_t)(t & 1) * 32 * ldp;
            if (bi < t) {
                const double* Pa = buf + 7 * bi;
                const double* Tb = buf + (size_t)13 * ldp + 7 * bj;
                // rank-13 update, software-pipelined by hand: the operands of row r + 1 are in flight while the 36 FMAs of
                // row r issue (with one block warp on a scheduler nothing else hides the shared-memory latency)
                double pa[6], tb[6], pn[6], tn[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) { pa[i] = Pa[i]; tb[i] = Tb[i]; }
#if !defined(FTMPC_UNROLL_COND)
                // the row loop is ROLLED (800 B instead of 10 KB of straight-line code per stage): +4.7 % solves/s on the B200 --
                // the unrolled form stalled on instruction fetch (profiles/README.md, r02k)
                const int nr = (t < N) ? FTMPC_NX : FTMPC_NE;
#pragma unroll 1
                for (int r = 0; r < nr; ++r) {
                    const int rn = (r + 1 < nr) ? r + 1 : r;
#pragma unroll
                    for (int i = 0; i < 6; ++i) { pn[i] = Pa[(size_t)rn * ldp + i]; tn[i] = Tb[(size_t)rn * ldp + i]; }
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int j = 0; j < 6; ++j) acc[i][j] += pa[i] * tb[j];
#pragma unroll
                    for (int i = 0; i < 6; ++i) { pa[i] = pn[i]; tb[i] = tn[i]; }
                }
#else
#pragma unroll
                for (int r = 0; r < FTMPC_NX; ++r) {
                    if (r < FTMPC_NE || t < N) {               // the terminal stage has 9 rows (uniform branch)
                        if (r + 1 < FTMPC_NX) {
#pragma unroll
                            for (int i = 0; i < 6; ++i) { pn[i] = Pa[(size_t)(r + 1) * ldp + i]; tn[i] = Tb[(size_t)(r + 1) * ldp + i]; }
                        }
#pragma unroll
                        for (int i = 0; i < 6; ++i)
#pragma unroll
                            for (int j = 0; j < 6; ++j) acc[i][j] += pa[i] * tb[j];
#pragma unroll
                        for (int i = 0; i < 6; ++i) { pa[i] = pn[i]; tb[i] = tn[i]; }
                    }
                }
#endif
            } else if (bi == t) {
                if (bj < t) {
                    // new block row: theta W_ux G_t   (row i of the block <- input i, column j <- column 6 bj + j)
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
        }
        blk.sync();
        blk.mark(PH_COND_BLK);                     // thread 0 owns block (0,0), the longest-lived accumulator
    }
    if (hb) {                                      // hand the blocks over in registers
        hb->fast = true;
        hb->bi = bi;
        hb->bj = bj;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) hb->acc[i][j] = acc[i][j];
        return;
    }
    if (bi >= 0) {                                 // store H (lower triangle)
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int ra = 6 * bi + i, cb = 6 * bj + j;
                if (cb <= ra) s.E[(size_t)ra * ld + cb] = acc[i][j];
            }
    }
    blk.sync();
}

// ---- register-resident block Cholesky + L^-T, ONE sweep -------------------------------------------------
// H arrives as 6x6 blocks in the registers of their owner threads (from `condense`), so the factorisation never
// read-modify-writes shared memory.  The triangular inverse X = L^-1 is built inside the same sweep: once block (i,j)
// has become L_ij (step j) its accumulator is dead, so it is reused for  S_ij = sum_{m=j}^{i-1} L_im X_mj , and
// X_ij = -L_ii^-1 S_ij falls out at step i.  Per block column k, two barrier intervals:
//   panel   owners of (i,k), i > k:  L_ik = A_ik L_kk^-T            -> lower triangle of E (kept in registers too);
//           owners of (k,j), j < k:  X_kj = -L_kk^-1 S_kj           -> stored TRANSPOSED in the upper triangle (J = L^-T);
//   update  owners of (i,j), i > k:  j > k:  A_ij -= L_ik L_jk'     (Cholesky trailing update)
//                                    j = k:  S_ik  = L_ik X_kk      (first term of the inverse recurrence, from registers)
//                                    j < k:  S_ij += L_ik X_kj
//           every thread below block row k does exactly one 6x6x6 product per step (the two kinds of update are
//           complementary), the active threads are a contiguous suffix of the block, and the separate wavefront pass of
//           the inverse (19 more barrier intervals) is gone.  The owner of (k+1,k+1) factors and inverts its block at
//           the end of the same interval, overlapped with the other warps' updates.
// Returns 0, or 6k+1.. when a pivot of block column k is not safely positive (uniform over the block).
__device__ __forceinline__ void chol_diag_block(double (&acc)[6][6], int k, int ld, double* E, double* linv, double* flag,
                                                double piv_tol) {
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
    // Y = L_kk^-1 (lower)
    double Y[6][6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i < j) Y[i][j] = 0.0;
            else if (i == j) Y[i][j] = inv[j];
            else {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < 6; ++m) if (m >= j && m < i) v += acc[i][m] * Y[m][j];
                Y[i][j] = -v * inv[i];
            }
        }
    }
    double* lk = linv + (size_t)k * 36;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            lk[i * 6 + j] = Y[i][j];
            // diagonal block of E: X_kk^T in the upper triangle (incl. diagonal), zeros strictly below
            E[(size_t)(6 * k + j) * ld + 6 * k + i] = Y[i][j];      // (j,i) <- Y[i][j]; for i < j this writes the zeros
        }
    if (bad) *flag = (double)(6 * k + 1);
}

__device__ __forceinline__ int chol_inv_blocks(CudaBlock& blk, int Nb, int ld, double* E, double* linv, HBlocks& hb,
                                               double piv_tol) {
    const int bi = hb.bi, bj = hb.bj;
    double (&acc)[6][6] = hb.acc;
    double* flag = linv + (size_t)Nb * 36;
    if (bi == 0 && bj == 0) {                      // thread 0
        *flag = 0.0;
        chol_diag_block(acc, 0, ld, E, linv, flag, piv_tol);
    }
    blk.sync();
    for (int k = 0; k < Nb; ++k) {
        if (*flag != 0.0) return (int)*flag;
        // ---------------- panel
        if (bj == k && bi > k) {                   // L_ik = A_ik L_kk^-T
            const double* lk = linv + (size_t)k * 36;
            double Y[6][6];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) Y[i][j] = (j <= i) ? lk[i * 6 + j] : 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double x[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    double v = 0.0;
#pragma unroll
                    for (int m = 0; m < 6; ++m) if (m <= j) v += acc[i][m] * Y[j][m];
                    x[j] = v;
                }
                double* er = E + (size_t)(6 * bi + i) * ld + 6 * k;
#pragma unroll
                for (int j = 0; j < 6; ++j) { acc[i][j] = x[j]; er[j] = x[j]; }
            }
        } else if (bi == k && bj >= 0 && bj < k) { // X_kj = -L_kk^-1 S_kj, stored transposed
            const double* lk = linv + (size_t)k * 36;
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
                double* er = E + (size_t)(6 * bj + b) * ld + 6 * k;
#pragma unroll
                for (int a = 0; a < 6; ++a) er[a] = x[a];
            }
        }
        blk.sync();
        // ---------------- update
        if (bi > k) {
            if (bj == k) {                         // S_ik = L_ik X_kk  (L_ik is still in the registers)
                const double* lk = linv + (size_t)k * 36;
                double Y[6][6];
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = 0; j < 6; ++j) Y[i][j] = (j <= i) ? lk[i * 6 + j] : 0.0;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    double x[6];
#pragma unroll
                    for (int b = 0; b < 6; ++b) {
                        double v = 0.0;
#pragma unroll
                        for (int m = 0; m < 6; ++m) if (m >= b) v += acc[a][m] * Y[m][b];
                        x[b] = v;
                    }
#pragma unroll
                    for (int b = 0; b < 6; ++b) acc[a][b] = x[b];
                }
            } else {
                // second operand: rows 6 bj .. of E at columns 6k..: L_jk (j > k, lower triangle) or X_kj^T (j < k, upper
                // triangle) -- the same addressing, only the sign of the product differs
                const double sgn = (bj > k) ? -1.0 : 1.0;
#if defined(FTMPC_ROLL_CHOL)
                // rank-1 form with the inner-dimension loop ROLLED (12 operand loads + 36 FMAs per trip instead of 216 FMAs of
                // straight-line code): the instruction footprint of the sweep, not its arithmetic, was what stalled it
                const double* ea = E + (size_t)(6 * bi) * ld + 6 * k;
                const double* eb = E + (size_t)(6 * bj) * ld + 6 * k;
#pragma unroll 1
                for (int m = 0; m < 6; ++m) {
                    double a6[6], b6[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) { a6[i] = sgn * ea[(size_t)i * ld + m]; b6[i] = eb[(size_t)i * ld + m]; }
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int j = 0; j < 6; ++j) acc[i][j] += a6[i] * b6[j];
                }
#else
                double Ob[6][6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const double* er = E + (size_t)(6 * bj + j) * ld + 6 * k;
#pragma unroll
                    for (int m = 0; m < 6; ++m) Ob[j][m] = er[m];
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const double* er = E + (size_t)(6 * bi + i) * ld + 6 * k;
                    double La[6];
#pragma unroll
                    for (int m = 0; m < 6; ++m) La[m] = sgn * er[m];
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        double v = acc[i][j];
#pragma unroll
                        for (int m = 0; m < 6; ++m) v += La[m] * Ob[j][m];
                        acc[i][j] = v;
                    }
                }
#endif
                if (bi == k + 1 && bj == k + 1) chol_diag_block(acc, k + 1, ld, E, linv, flag, piv_tol);
            }
        }
        blk.sync();
    }
    blk.mark(PH_CHOL);
    // the strictly lower blocks (L) are dead: J is upper triangular
    if (bi > bj) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double* er = E + (size_t)(6 * bi + i) * ld + 6 * bj;
#pragma unroll
            for (int j = 0; j < 6; ++j) er[j] = 0.0;
        }
    }
    blk.sync();
    blk.mark(PH_INV);
    return 0;
}
#endif  // __CUDACC__

// ---- condensed Hessian -> J = L^-T (rows 0..n-1 of E); returns 0 or the failing pivot + 1 -------------------
template <class Blk>
FT_HD int factor_hessian(Blk& blk, const ftmpc_config& cfg, const WsLayout& L, const QpScratch& s, double* Jz, double* Wz,
                         const double* Jz_src, const double* Wz_src, const double* X, const double* U, const double* xref,
                         const double* gradV, const double* hessV, double theta, double sigma, const double* lam_prev,
                         double* dscale_out, bool copy_j = true, bool copy_w = true, const double* Cq = nullptr) {
    const int N = L.N, n = L.n, ld = L.nv, tid = blk.tid(), nt = blk.nthreads();
    (void)copy_j; (void)copy_w;
    for (int i = tid; i < N * 169; i += nt) { Jz[i] = Jz_src[i]; Wz[i] = Wz_src[i]; }     // (the QP reuses this region)
    blk.sync();
    condense(blk, cfg, L, s, Jz, Wz, X, U, xref, gradV, hessV, theta, sigma, lam_prev, Cq);
    blk.mark(PH_COND);
    blk.count(CT_CONDENSE);
    double dmaxl = 0.0;
    for (int i = tid; i < n; i += nt) dmaxl = fmax(dmaxl, fabs(s.E[(size_t)i * ld + i]));
    const double dscale = blk.max(dmaxl);
    *dscale_out = dscale;
    const int bad = chol_lower(blk, n, ld, s.E, 1e-10 * fmax(1.0, dscale));
    blk.mark(PH_CHOL);
    if (bad) return bad;
    tri_inv_transpose(blk, n, ld, s.E, s.dg);
    blk.mark(PH_INV);
    return 0;
}
#if defined(__CUDACC__)
__device__ __forceinline__ int factor_hessian(CudaBlock& blk, const ftmpc_config& cfg, const WsLayout& L, const QpScratch& s,
                                              double* Jz, double* Wz, const double* Jz_src, const double* Wz_src,
                                              const double* X, const double* U, const double* xref, const double* gradV,
                                              const double* hessV, double theta, double sigma, const double* lam_prev,
                                              double* dscale_out, bool copy_j = true, bool copy_w = true, const double* Cq = nullptr) {
    const int N = L.N, n = L.n, ld = L.nv, tid = blk.tid(), nt = blk.nthreads();
    if (!condense_fast_path(L, nt))
        return factor_hessian<CudaBlock>(blk, cfg, L, s, Jz, Wz, Jz_src, Wz_src, X, U, xref, gradV, hessV, theta, sigma,
                                         lam_prev, dscale_out, true, true, Cq);
    // the linearisation leaves Jz / Wz in place (shared memory); Wz is scaled in place below, so a second attempt
    // re-reads Wz (and, once the active-set solver has reused the region, Jz too) from the global backing copy
    if (copy_j) for (int i = tid; i < N * 169; i += nt) Jz[i] = Jz_src[i];
    if (copy_w) for (int i = tid; i < N * 169; i += nt) Wz[i] = Wz_src[i];
    blk.sync();
    HBlocks hb;
    condense(blk, cfg, L, s, Jz, Wz, X, U, xref, gradV, hessV, theta, sigma, lam_prev, &hb, Cq);
    blk.mark(PH_COND);
    blk.count(CT_CONDENSE);
    double dmaxl = 0.0;
    if (hb.bi >= 0 && hb.bi == hb.bj) {
#pragma unroll
        for (int i = 0; i < 6; ++i) dmaxl = fmax(dmaxl, fabs(hb.acc[i][i]));
    }
    const double dscale = blk.max(dmaxl);
    *dscale_out = dscale;
    (void)n;
    return chol_inv_blocks(blk, N, ld, s.E, s.T, hb, 1e-10 * fmax(1.0, dscale));
}
#endif

// development aid of the CPU port: counts which kind of Hessian attempt failed (compiled out everywhere else)
#if defined(FTMPC_DEBUG_COUNTERS) && !defined(__CUDACC__)
extern long g_ftmpc_dbg[8];
#define FT_DBG_COUNT(k) (__sync_fetch_and_add(&g_ftmpc_dbg[k], 1L))
extern long g_ftmpc_dbg2[8];
extern thread_local int g_ftmpc_warm_hit;
#define FT_DBG_COUNT2(bin, hit) (__sync_fetch_and_add(&g_ftmpc_dbg2[2 * (bin) + (hit)], 1L))
#else
#define FT_DBG_COUNT(k) ((void)0)
#endif

}  // namespace ftmpc
#include "ftmpc_riccati.cuh"      // needs FTMPC_CQ and the scratch structs above
namespace ftmpc {

// the dense path (condense + Cholesky + L^-T) stays selectable on the host (qp_method bit 4) and in CUDA builds made
// with -DFTMPC_DENSE_FACTOR (A/B measurements); the product kernel carries only the Riccati factorisation
#if !defined(__CUDACC__) || defined(FTMPC_DENSE_FACTOR)
#define FTMPC_HAVE_DENSE_FACTOR 1
#else
#define FTMPC_HAVE_DENSE_FACTOR 0
#endif

// ---- J, X J, g, ga, J' ga by the Riccati factorisation (ftmpc_riccati.cuh); same contract as factor_hessian ---------
template <class Blk>
FT_HD int factor_riccati(Blk& blk, const ftmpc_config& cfg, const WsLayout& L, const QpScratch& s, double* Jz, double* Wz,
                         const double* Jz_src, const double* Wz_src, const double* X, const double* U, const double* xref,
                         const double* gradV, const double* hessV, double theta, double sigma, const double* lam_prev,
                         double* dscale_out, bool copy_j, bool copy_w, const double* Cq, bool want_columns = true) {
    const int N = L.N, tid = blk.tid(), nt = blk.nthreads();
    // the linearisation may have left Jz / Wz in place; the stage records overwrite Wz, so another attempt re-reads it
    if (copy_j) for (int i = tid; i < N * 169; i += nt) Jz[i] = Jz_src[i];
    if (copy_w) for (int i = tid; i < N * 169; i += nt) Wz[i] = Wz_src[i];
    blk.sync();
    const int bad = riccati_factor(blk, cfg, L, s, Jz, Wz, X, U, xref, gradV, hessV, theta, sigma, lam_prev, Cq, s.G, s.E,
                                   dscale_out, want_columns);
    blk.mark(PH_CHOL);
    blk.count(CT_CONDENSE);
    return bad;
}

// ---- K as an operator for gis_solve_op: the whole block calls, one warp (device) or the caller (host) sweeps -------------
template <class Blk>
struct RicKOp {
    RicOp op;
    const double* G;        // stage matrices of the G-form (ric_build_g) or nullptr: two-interval form on (Jz, records)
    FT_HD void apply(Blk& blk, const double* v, double* out, int t_top, bool has_e) {
        if (t_top < 0 && !has_e) {                  // only the elastic variable: K is diagonal there
            for (int i = blk.tid(); i < op.nv + FTMPC_NE; i += blk.nthreads()) out[i] = (i == op.n) ? v[op.n] * op.inv_rho : 0.0;
            blk.sync();
            return;
        }
        if (G) ric_apply_g(blk, op, G, v, out, t_top, has_e);
        else ric_apply(blk, op, v, out, t_top, has_e);
    }
};
#if defined(__CUDACC__)
// on-chip memory the global-scratch kernel hands to phase_qp (fast_work): Riccati workspace / operator scratch, the double
// buffer of staged stages, v and K v
FT_HD size_t gs_fast_doubles(int N) {
    return (size_t)FTMPC_RIC_WORK + 2 * (size_t)FTMPC_RIC_CH * FTMPC_RIC_GSTG + 2 * (size_t)(FTMPC_NU * N + 1 + FTMPC_NE);
}
template <>
struct RicKOp<CudaBlock> {
    RicOp op;
    const double* G;
    RicStage sg;
    bool staged;
    __device__ __forceinline__ void apply(CudaBlock& blk, const double* v, double* out, int t_top, bool has_e) {
        if (t_top < 0 && !has_e) {
            for (int i = blk.tid(); i < op.nv + FTMPC_NE; i += blk.nthreads()) out[i] = (i == op.n) ? v[op.n] * op.inv_rho : 0.0;
            blk.sync();
            return;
        }
        if (staged && G) {
            ric_apply_staged(blk, op, G, sg, v, out, t_top, has_e);
            return;
        }
        if (threadIdx.x < 32) {
            WarpBlock wb;
            if (G) ric_apply_g(wb, op, G, v, out, t_top, has_e);
            else ric_apply(wb, op, v, out, t_top, has_e);
        }
        blk.sync();
    }
};
#endif
template <class Blk>
FT_HD void qp_op_stage(RicKOp<Blk>&, double*, int, unsigned, unsigned*) {}
#if defined(__CUDACC__)
__device__ __forceinline__ void qp_op_stage(RicKOp<CudaBlock>& kop, double* fast_work, int ne, unsigned mbar, unsigned* par) {
    kop.staged = fast_work != nullptr && mbar != 0u;
    kop.sg.buf = fast_work ? fast_work + FTMPC_RIC_WORK : nullptr;
    kop.sg.sv = fast_work ? kop.sg.buf + 2 * (size_t)FTMPC_RIC_CH * FTMPC_RIC_GSTG : nullptr;
    kop.sg.so = fast_work ? kop.sg.sv + ne : nullptr;
    kop.sg.mbar = mbar;
    kop.sg.par = par;
}
#endif
// horizons above FTMPC_LONG_N run the QP in operator form (qp_method bit 5 forces it at any horizon, bit 6 forbids it)
FT_HD bool qp_operator_form(const ftmpc_config& cfg, int N) {
    if (cfg.qp_method & 64) return false;
    return N > FTMPC_LONG_N || (cfg.qp_method & 32) != 0;
}
#define FTMPC_OP_QCAP 640        /* capacity of the working set in operator form (R^-1 packed: 1.6 MB of the per-CTA global slice at most) */

// ---- phase_qp ------------------------------------------------------------------------------------------
template <class Blk>
FT_HD void phase_qp(Blk& blk, const ftmpc_config& cfg, const WsLayout& L, const StepIO& io, int inst, int slot, double* scratch,
                   bool staged = false /* Jz, Wz already sit in the scratch (CUDA linearisation) */,
                   double* fast_work = nullptr /* gs_fast_doubles() of on-chip memory when the scratch itself is not */,
                   unsigned op_mbar = 0u, unsigned* op_par = nullptr /* mbarrier + parity word of the operator's bulk copies */) {
    double* w = ws_slot(io, L, slot);
    double* sc = w + L.oSc;
    if (sc[SC_STATUS] != FTMPC_ST_RUNNING) return;
    const int N = L.N, n = L.n, nv = L.nv, ne = nv + FTMPC_NE, ld = nv, tid = blk.tid(), nt = blk.nthreads();
    QpScratch s = qp_carve(scratch, N, io.cfg_g);
    s.tf_val = io.tf_val;
    s.tf_idx = io.tf_idx;
    if (fast_work) s.G = fast_work;              // workspace of riccati_factor (the G region is only read by the dense path)
    const double* xref = io.xref + (size_t)inst * io.xref_stride;
    const double* hull_g = io.hull_table + (size_t)io.hull_idx[inst] * FTMPC_HULL_STRIDE;
    // stage data -> scratch (the R^-1 region is free until the active-set solve starts)
    double* Jz = s.RS;
    double* Wz = s.RS + (size_t)N * 169;
    for (int i = tid; i < L.mc; i += nt) s.cv[i] = w[L.oC + i];
    for (int i = tid; i < FTMPC_HULL_STRIDE; i += nt) s.hull[i] = hull_g[i];
    blk.sync();
    // Hessian schedule: exact second-order terms blended by theta; when the exact Hessian is indefinite,
    // first try the augmented-Lagrangian convexification (sigma > 0), then fall back to smaller theta.
    const double* lam_prev = w + L.oLam;
    const QpStart qs = qp_start(cfg, sc);
    const bool can_aug = qs.can_aug, skip_exact = qs.skip_exact;
    // |d| of the previous iteration is remembered for sqp_fast_converged when that iteration was an exact-Hessian full step
    const double dprev_keep = (sc[SC_ITER] > 0.0 && sc[SC_THETA] == 1.0 && sc[SC_ALPHA] == 1.0) ? sc[SC_DMAX] : 0.0;
    double theta = qs.theta, sigma = qs.sigma;
    FT_DBG_COUNT(theta == 0.0 ? 6 : 7);            // (host development counter: pure Gauss-Newton iterations vs blended)
    bool aug_allowed = can_aug;
    int fails = 0, qit = 0, nact = 0, st = GI_OK, aug_retry = 0;
    bool have_j = staged, have_w = staged;
    const bool use_op = qp_operator_form(cfg, N);
#if FTMPC_HAVE_DENSE_FACTOR
    const bool dense_factor = (cfg.qp_method & 16) != 0;      // bit 4: condensed Hessian + Cholesky + triangular inverse (the round-1 path)
#endif
    for (;;) {        // QP attempts (re-solved without augmentation if a predicted-active row came out inactive)
    double sig0 = 0.0;
    for (;;) {
        double dscale = 0.0;
        int bad;
#if FTMPC_HAVE_DENSE_FACTOR
        if (dense_factor)
            bad = factor_hessian(blk, cfg, L, s, Jz, Wz, w + L.oJz, w + L.oWz, w + L.oX, w + L.oU, xref, w + L.oGV,
                                 w + L.oHV, theta, sigma, lam_prev, &dscale, !have_j, !have_w, io.uref ? w + L.oCq : nullptr);
        else
#endif
            bad = factor_riccati(blk, cfg, L, s, Jz, Wz, w + L.oJz, w + L.oWz, w + L.oX, w + L.oU, xref, w + L.oGV,
                                 w + L.oHV, theta, sigma, lam_prev, &dscale, !have_j, !have_w, io.uref ? w + L.oCq : nullptr,
                                 !use_op);
        have_j = true;                  // Jz survives a failed factorisation, the scaled Wz does not
        have_w = false;
        if (!bad) break;
        blk.count(CT_CHOL_FAIL);
        ++fails;
        blk.sync();
        if (sigma == 0.0 && theta == 1.0 && aug_allowed) { sig0 = 10.0 * dscale; sigma = sig0; FT_DBG_COUNT(0); }
        else if (sigma > 0.0 && sig0 > 0.0 && sigma < 5.0 * sig0) { sigma *= 10.0; FT_DBG_COUNT(1); }
        else if (sigma > 0.0) { sigma = 0.0; theta = 0.5; aug_allowed = false; FT_DBG_COUNT(2); }
        else if (theta <= 0.0) {
            // even the Gauss-Newton model could not be factorised (NaN / non-positive weights): every thread must see
            // SC_QPST before the step acceptance reads it, otherwise the warps of the block take different barrier paths
            if (tid == 0) sc[SC_QPST] = 3.0;
            blk.sync();
            return;
        }
        else { FT_DBG_COUNT(theta == 1.0 ? 3 : 4); theta = (theta > cfg.theta_first) ? 0.5 * theta : 0.0; }     // below the first blend level: Gauss-Newton
    }
    MpcCons cons{N, n, nv, L.mc, s.hull, io.cfg_g->Af, s.cv, io.tf_val, io.tf_idx};
    int qit1 = 0;
    if (use_op) {
        // ---- operator form: nothing of size n^2 exists; K = E E' is applied through the stage records (ric_apply)
        RicKOp<Blk> kop;
        kop.op.N = N; kop.op.n = n; kop.op.nv = nv; kop.op.Jz = Jz; kop.op.Rec = Wz; kop.op.dt = cfg.dt;
        kop.op.inv_rho = 1.0 / cfg.rho_slack;
        kop.op.scr = s.G;                           // on-chip when the caller provided fast_work
        qp_op_stage(kop, fast_work, ne, op_mbar, op_par);
        // G-form when the scratch has the room (the GL region of long horizons), else the two-interval form
        kop.G = nullptr;
        if (s.GL) {
            double* g16 = s.GL + ((reinterpret_cast<unsigned long long>(s.GL) >> 3) & 1ull);      // bulk copies: 16-byte aligned source
            ric_build_g(blk, N, cfg.dt, Jz, Wz, g16);
            kop.G = g16;
        }
        // work vectors carved from the (unused) E region; R^-1 goes behind them: the RS region holds the stage records
        double* p = s.E;
        GisWork gw;
        gw.qcap = (nv < FTMPC_OP_QCAP) ? nv : FTMPC_OP_QCAP;
        gw.K = nullptr;
        gw.xe = s.gi.xe; gw.s = s.gi.s; gw.pos = s.gi.pos;
        gw.ye = p; p += ne; gw.ze = p; p += ne; gw.c = p; p += ne;
        double* vin = p; p += ne;
        gw.u = p; p += gw.qcap + 2; gw.w = p; p += gw.qcap + 2; gw.v = p; p += gw.qcap + 2; gw.r = p; p += gw.qcap + 2;
        gw.cs = p; p += 2 * (gw.qcap + 2); gw.tmp = p; p += gw.qcap + 2; gw.sub = p; p += gw.qcap + 2;
        gw.act = reinterpret_cast<int*>(p); p += (gw.qcap + 3) / 2 + 1;
        gw.itmp = reinterpret_cast<int*>(p); p += (gw.qcap + 3) / 2 + 1;
        gw.Ui = p; p += (size_t)gw.qcap * (gw.qcap + 1) / 2 + 2;
        // K n_k of the working set, when the E region has the room (long horizons: it is the memory E itself would have taken)
        int* yslot = reinterpret_cast<int*>(p); p += (gw.qcap + 3) / 2 + 1;
        const size_t yroom = ((size_t)ne * nv > (size_t)(p - s.E)) ? ((size_t)ne * nv - (size_t)(p - s.E)) / ne : 0;
        const int ycap = (yroom > (size_t)gw.qcap + 1) ? gw.qcap + 1 : (int)yroom;
        double* Yw = (ycap >= 16) ? p : nullptr;
        // unconstrained minimiser  x = -K [ga ; 0 ; 0]
        for (int i = tid; i < ne; i += nt) vin[i] = (i < n) ? -s.ga[i] : 0.0;
        blk.sync();
        kop.apply(blk, vin, s.gi.xe, N - 1, false);
        blk.mark(PH_QPSETUP);
        blk.count(CT_QP);
        st = gis_solve_op(blk, cons, gw, kop, vin, ne, L.m, w + L.oLam + L.m, cfg.max_qp_iter, cfg.qp_tol, &qit1, &nact, Yw, yslot, ycap);
        have_j = false; have_w = false;
    } else {
    // slack variable column/row, extension rows  X J  (X = d x_N[0:9] / d U)
    for (int i = tid; i < nv; i += nt) {
        s.E[(size_t)i * ld + n] = 0.0;
        s.E[(size_t)n * ld + i] = (i == n) ? 1.0 / sqrt(cfg.rho_slack) : 0.0;
    }
    blk.sync();
#if FTMPC_HAVE_DENSE_FACTOR
    if (dense_factor) {
    // column i of [X J ; ga' J]: ten accumulators share one sweep down column i of J (J is upper triangular)
    for (int i = tid; i < nv; i += nt) {
        double acc[FTMPC_NE + 1];
        for (int kk = 0; kk <= FTMPC_NE; ++kk) acc[kk] = 0.0;
        if (i < n) {
            for (int r = 0; r <= i; ++r) {
                const double e = s.E[(size_t)r * ld + i];
                for (int kk = 0; kk < FTMPC_NE; ++kk) acc[kk] += s.G[kk * ld + r] * e;
                acc[FTMPC_NE] += s.ga[r] * e;
            }
        }
        for (int kk = 0; kk < FTMPC_NE; ++kk) s.E[(size_t)(nv + kk) * ld + i] = acc[kk];
        s.gi.d[i] = acc[FTMPC_NE];
    }
    } else
#endif
    {   // riccati_factor has written X J and J' ga; only the slack column is left
        for (int kk = tid; kk <= FTMPC_NE; kk += nt) {
            if (kk < FTMPC_NE) s.E[(size_t)(nv + kk) * ld + n] = 0.0;
            else s.gi.d[n] = 0.0;
        }
    }
    blk.sync();
    // unconstrained minimiser  x = -J J' ga  (extended coordinates)
    for (int row = tid; row < ne; row += nt) {
        const double* e = s.E + (size_t)row * ld;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int kk = 0;
        for (; kk + 3 < nv; kk += 4) {
            a0 += e[kk] * s.gi.d[kk];
            a1 += e[kk + 1] * s.gi.d[kk + 1];
            a2 += e[kk + 2] * s.gi.d[kk + 2];
            a3 += e[kk + 3] * s.gi.d[kk + 3];
        }
        for (; kk < nv; ++kk) a0 += e[kk] * s.gi.d[kk];
        s.gi.xe[row] = -((a0 + a1) + (a2 + a3));
    }
    blk.sync();
    blk.mark(PH_QPSETUP);
    blk.count(CT_QP);
    // the previous multipliers are still needed if this attempt is rejected: the QP writes to the spare copy
#if !defined(__CUDACC__)
    if (cfg.qp_method == 2) {
        // range-space form, host prototype: K = E E' (packed) from the dense E the null-space set-up built
        const int qcap = 80;
        std::vector<double> Kp((size_t)ne * (ne + 1) / 2), Ui((size_t)qcap * (qcap + 1) / 2 + 1), vec((size_t)3 * ne + 8 * (qcap + 2));
        std::vector<int> iv(2 * (qcap + 2));
        for (int i = 0; i < ne; ++i)
            for (int j = 0; j <= i; ++j) {
                double a = 0.0;
                for (int k = 0; k < nv; ++k) a += s.E[(size_t)i * ld + k] * s.E[(size_t)j * ld + k];
                Kp[(size_t)i * (i + 1) / 2 + j] = a;
            }
        GisWork gw;
        double* v = vec.data();
        gw.K = Kp.data(); gw.Ui = Ui.data(); gw.xe = s.gi.xe; gw.s = s.gi.s;
        gw.ye = v; v += ne; gw.ze = v; v += ne; gw.c = v; v += ne;
        gw.u = v; v += qcap + 2; gw.w = v; v += qcap + 2; gw.v = v; v += qcap + 2; gw.r = v; v += qcap + 2;
        gw.cs = v; v += 2 * (qcap + 2); gw.tmp = v; v += qcap + 2; gw.sub = v; v += qcap + 2;
        gw.act = iv.data(); gw.itmp = iv.data() + qcap + 2; gw.pos = s.gi.pos; gw.qcap = qcap;
        st = gis_solve(blk, cons, gw, ne, L.m, w + L.oLam + L.m, cfg.max_qp_iter, cfg.qp_tol, &qit1, &nact);
    } else
#endif
    st = gi_solve(blk, cons, s.gi, nv, ne, ld, L.m, 0, w + L.oLam + L.m, cfg.max_qp_iter, cfg.qp_tol, &qit1, &nact,
                  (cfg.warm_qp != 0 && sc[SC_ITER] > 0.0 && sc[SC_DMAX] <= 1.0) ? lam_prev : nullptr, L.mc);    // hit rate 4 % above |d| = 1
    have_j = false;                     // R^-1 has overwritten the staged Jacobians
    }
#if defined(FTMPC_DEBUG_COUNTERS) && !defined(__CUDACC__)
    if (cfg.warm_qp != 0 && sc[SC_ITER] > 0.0) {       // warm-start outcome binned by the size of the previous step
        const double dm = sc[SC_DMAX];
        const int bin = dm > 1.0 ? 0 : (dm > 0.1 ? 1 : (dm > 1e-2 ? 2 : 3));
        FT_DBG_COUNT2(bin, g_ftmpc_warm_hit ? 1 : 0);
    }
#endif
    blk.mark(PH_GI);
    qit += qit1;
    if (sigma > 0.0 && st == GI_OK) {       // every predicted-active row must be active in the QP solution
        int viol = 0;
        for (int i = tid; i < L.mc; i += nt)
            if (lam_prev[i] > 0.0 && s.gi.pos[i] < 0 && s.gi.s[i] > 1e-9) viol = 1;
        if (blk.any(viol)) {
            ++fails;
            FT_DBG_COUNT(5);
            if (aug_retry < 2) {
                // drop the rows that came out inactive from the predicted set and convexify again: the exact Hessian
                // with the corrected set keeps the Newton-like rate, the theta = 1/2 fallback below does not
                ++aug_retry;
                double* lp = w + L.oLam;
                for (int i = tid; i < L.mc; i += nt)
                    if (lp[i] > 0.0 && s.gi.pos[i] < 0 && s.gi.s[i] > 1e-9) lp[i] = 0.0;
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
    // step, directional derivative, multiplier bound
    double gd = 0.0, dmx = 0.0, lmx = 0.0;
    for (int i = tid; i < n; i += nt) {
        const double di = s.gi.xe[i];
        w[L.oD + i] = di;
        gd += s.g[i] * di;
        dmx = fmax(dmx, fabs(di));
    }
    for (int i = tid; i < L.mc; i += nt) lmx = fmax(lmx, w[L.oLam + i]);
    gd = blk.sum(gd);
    dmx = blk.max(dmx);
    lmx = blk.max(lmx);
    if (tid == 0) {
        w[L.oD + n] = s.gi.xe[n];
        sc[SC_DPREV] = dprev_keep;
        sc[SC_GD] = gd; sc[SC_DMAX] = dmx; sc[SC_LAMMAX] = lmx; sc[SC_DELTA] = s.gi.xe[n];
        sc[SC_HFAIL] = (skip_exact || (fails > 0 && theta == 0.0)) ? 1.0 : 0.0;
        sc[SC_THETA] = theta; sc[SC_SIGMA] = sigma; sc[SC_QPIT] += qit; sc[SC_NACT] = nact; sc[SC_CHOLFAIL] += fails;
        sc[SC_QPST] = (st == GI_OK && dmx == dmx) ? 0.0 : (double)(st ? st : 4);
    }
    blk.sync();
    blk.mark(PH_POST);
}

}  // namespace ftmpc
