// ftmpc_dyn.cuh -- orbit-centre prediction model, RK4, and its first/second-order derivatives.
//
// Replaces (reference, paths relative to /root/reference):
//   SpiralModel.dx_dt            ft_mpc/models/spiral_model.py:44-76
//   OmegaOperator                ft_mpc/models/sys_model.py:8-29
//   RotCasadi                    ft_mpc/util/utils.py:33-55
//   SystemModel.rk4_integrator   ft_mpc/models/sys_model.py:138-162
//   CasADi AD of the above inside nlpsol (spiraling_mpc.py:171,230)
//
// The model is a polynomial map, so every derivative below is written with three symmetric
// bilinear forms (R(q)=1/2 B(q,q), c(w)=1/2 C(w,w), e(w)=1/2 E(w,w)); first and second
// directional derivatives then follow from the product rule with no further algebra.
//
// State x = [p(3) v(3) w(3) q(4)], q = [x y z w] scalar last, wrench W = [F(3) tau(3)].
// "z-space" = the 13 variables the dynamics are nonlinear in: [w(3) q(4) F(3) tau(3)].
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FT_HD __host__ __device__ __forceinline__
#else
#define FT_HD inline
#endif

namespace ftmpc {

struct DynConsts {
    double dt, mass, Jd[3], r[3];
    double im, iJ[3];        // reciprocals: an fp64 division costs ~10 multiplications on the GPU
};

FT_HD void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
FT_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// B(q,p): symmetric bilinear form with Rot(q) = 1/2 B(q,q)   (utils.py:12-18)
FT_HD void rot_bilinear(const double* q, const double* p, double M[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double a = p[0], b = p[1], c = p[2], d = p[3];
    const double xx = x * a, yy = y * b, zz = z * c, ww = w * d;
    M[0] = 2.0 * (xx - yy - zz + ww);
    M[1] = 2.0 * (x * b + a * y + z * d + c * w);
    M[2] = 2.0 * (x * c + a * z - y * d - b * w);
    M[3] = 2.0 * (x * b + a * y - z * d - c * w);
    M[4] = 2.0 * (-xx + yy - zz + ww);
    M[5] = 2.0 * (y * c + b * z + x * d + a * w);
    M[6] = 2.0 * (x * c + a * z + y * d + b * w);
    M[7] = 2.0 * (y * c + b * z - x * d - a * w);
    M[8] = 2.0 * (-xx - yy + zz + ww);
}
FT_HD void rot_mat(const double* q, double M[9]) {
    rot_bilinear(q, q, M);
    for (int i = 0; i < 9; ++i) M[i] *= 0.5;
}
FT_HD void mat3_mul(const double M[9], const double* v, double* o) {          // o = M v
    o[0] = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    o[1] = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    o[2] = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
}
FT_HD void mat3_tmul(const double M[9], const double* v, double* o) {         // o = M^T v
    o[0] = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    o[1] = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    o[2] = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
}
// out[j] = g^T B(q,e_j) a   (gradient w.r.t. q of  g . (Rot(q) a)).
// g' Rot(q) a is the quadratic form q' S q with the symmetric 4x4 matrix S(g, a) below, and B is its
// polarisation, so the gradient is 2 S q: 9 products for S instead of four bilinear-form evaluations.
FT_HD void quat_grad(const double* q, const double* a, const double* g, double out[4]) {
    const double P00 = g[0] * a[0], P01 = g[0] * a[1], P02 = g[0] * a[2];
    const double P10 = g[1] * a[0], P11 = g[1] * a[1], P12 = g[1] * a[2];
    const double P20 = g[2] * a[0], P21 = g[2] * a[1], P22 = g[2] * a[2];
    const double Sxx = P00 - P11 - P22, Syy = -P00 + P11 - P22, Szz = -P00 - P11 + P22, Sww = P00 + P11 + P22;
    const double Sxy = P01 + P10, Szw = P01 - P10, Sxz = P02 + P20, Syw = P20 - P02, Syz = P12 + P21, Sxw = P12 - P21;
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    out[0] = 2.0 * (Sxx * x + Sxy * y + Sxz * z + Sxw * w);
    out[1] = 2.0 * (Sxy * x + Syy * y + Syz * z + Syw * w);
    out[2] = 2.0 * (Sxz * x + Syz * y + Szz * z + Szw * w);
    out[3] = 2.0 * (Sxw * x + Syw * y + Szw * z + Sww * w);
}
// OmegaOperator(w) @ q                                                   (sys_model.py:8-29)
FT_HD void omega_apply(const double* w, const double* q, double* o) {
    o[0] = w[2] * q[1] - w[1] * q[2] + w[0] * q[3];
    o[1] = -w[2] * q[0] + w[0] * q[2] + w[1] * q[3];
    o[2] = w[1] * q[0] - w[0] * q[1] + w[2] * q[3];
    o[3] = -w[0] * q[0] - w[1] * q[1] - w[2] * q[2];
}
// C(w,n) = w x (J n) + n x (J w);  c(w) = w x (J w) = 1/2 C(w,w)
FT_HD void gyro_bilinear(const DynConsts& k, const double* w, const double* n, double* o) {
    const double Jn[3] = {k.Jd[0] * n[0], k.Jd[1] * n[1], k.Jd[2] * n[2]};
    const double Jw[3] = {k.Jd[0] * w[0], k.Jd[1] * w[1], k.Jd[2] * w[2]};
    double a[3], b[3];
    cross3(w, Jn, a);
    cross3(n, Jw, b);
    o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
}
// C(w,.)^T beta = J (beta x w) + (J w) x beta
FT_HD void gyro_bilinear_T(const DynConsts& k, const double* w, const double* beta, double* o) {
    double a[3], b[3];
    cross3(beta, w, a);
    const double Jw[3] = {k.Jd[0] * w[0], k.Jd[1] * w[1], k.Jd[2] * w[2]};
    cross3(Jw, beta, b);
    o[0] = k.Jd[0] * a[0] + b[0]; o[1] = k.Jd[1] * a[1] + b[1]; o[2] = k.Jd[2] * a[2] + b[2];
}
// E(w,n) = w (n.r) + n (w.r) - 2 r (w.n);  w x (w x r) = 1/2 E(w,w)
FT_HD void centri_bilinear(const DynConsts& k, const double* w, const double* n, double* o) {
    const double nr = dot3(n, k.r), wr = dot3(w, k.r), wn = dot3(w, n);
    for (int i = 0; i < 3; ++i) o[i] = w[i] * nr + n[i] * wr - 2.0 * k.r[i] * wn;
}
// E(w,.)^T h = r (h.w) + h (w.r) - 2 w (h.r)
FT_HD void centri_bilinear_T(const DynConsts& k, const double* w, const double* h, double* o) {
    const double hw = dot3(h, w), wr = dot3(w, k.r), hr = dot3(h, k.r);
    for (int i = 0; i < 3; ++i) o[i] = k.r[i] * hw + h[i] * wr - 2.0 * w[i] * hr;
}

// ---- continuous dynamics --------------------------------------------- spiral_model.py:44-76
// xs = [v(3) w(3) q(4)] is enough (p never enters the right-hand side).  Returns also g (body accel).
FT_HD void dyn_f(const DynConsts& k, const double* v, const double* w, const double* q, const double* Wr,
                 double* dp, double* dv, double* dw, double* dq, double* g_out) {
    double c[3], e[3], g[3], t[3];
    gyro_bilinear(k, w, w, c);
    centri_bilinear(k, w, w, e);
    for (int i = 0; i < 3; ++i) dw[i] = (Wr[3 + i] - 0.5 * c[i]) * k.iJ[i];       // :63-67
    cross3(dw, k.r, t);
    for (int i = 0; i < 3; ++i) g[i] = Wr[i] * k.im + t[i] + 0.5 * e[i];         // :70-72
    double R[9];
    rot_mat(q, R);
    mat3_tmul(R, g, dv);                                                             // :69 RotCasadi(q).T @ (...)
    double oq[4];
    omega_apply(w, q, oq);
    for (int i = 0; i < 4; ++i) dq[i] = 0.5 * oq[i];                                 // :75
    for (int i = 0; i < 3; ++i) dp[i] = v[i];                                        // :57
    if (g_out) for (int i = 0; i < 3; ++i) g_out[i] = g[i];
}

// The right-hand side splits into an attitude part that does not see (p, v) and a translational part that is a pure
// function of the attitude stages: the same arithmetic as dyn_f, in two pieces (used by the split rollout of the step
// acceptance, where only the attitude chain is serial).
FT_HD void dyn_wq(const DynConsts& k, const double* w, const double* q, const double* tau, double* dw, double* dq) {
    double c[3];
    gyro_bilinear(k, w, w, c);
    for (int i = 0; i < 3; ++i) dw[i] = (tau[i] - 0.5 * c[i]) * k.iJ[i];
    double oq[4];
    omega_apply(w, q, oq);
    for (int i = 0; i < 4; ++i) dq[i] = 0.5 * oq[i];
}
// The same attitude right-hand side arranged for a SERIAL chain (the RK stages of the step-acceptance rollout): Euler's
// equations for the diagonal inertia,  dw_x = tau_x / J_x - (J_z - J_y) / J_x * w_y w_z  (w x J w = [(J_z - J_y) w_y w_z, ...]),
// with  tq = tau / J  and  gy = [(J_z - J_y) / J_x, (J_x - J_z) / J_y, (J_y - J_x) / J_z]  prepared once per step: two
// dependent operations per stage for dw and three for dq instead of six / four (a dependent FP64 operation costs 23
// cycles on the B200).  Same value as dyn_wq up to rounding.  (Measured: the rollout's share of the kernel 7.5 % -> 7.2 %.)
struct AttConsts { double gy[3]; };
FT_HD AttConsts att_consts(const DynConsts& k) {
    AttConsts a;
    a.gy[0] = (k.Jd[2] - k.Jd[1]) * k.iJ[0];
    a.gy[1] = (k.Jd[0] - k.Jd[2]) * k.iJ[1];
    a.gy[2] = (k.Jd[1] - k.Jd[0]) * k.iJ[2];
    return a;
}
FT_HD void dyn_wq_chain(const AttConsts& a, const double* tq, const double* w, const double* q, double* dw, double* dq) {
    const double h0 = 0.5 * w[0], h1 = 0.5 * w[1], h2 = 0.5 * w[2];
    dw[0] = tq[0] - (a.gy[0] * w[1]) * w[2];
    dw[1] = tq[1] - (a.gy[1] * w[2]) * w[0];
    dw[2] = tq[2] - (a.gy[2] * w[0]) * w[1];
    dq[0] = h2 * q[1] - h1 * q[2] + h0 * q[3];
    dq[1] = -h2 * q[0] + h0 * q[2] + h1 * q[3];
    dq[2] = h1 * q[0] - h0 * q[1] + h2 * q[3];
    dq[3] = -h0 * q[0] - h1 * q[1] - h2 * q[2];
}
FT_HD void dyn_v(const DynConsts& k, const double* w, const double* q, const double* dw, const double* F, double* dv) {
    double e[3], g[3], t[3];
    centri_bilinear(k, w, w, e);
    cross3(dw, k.r, t);
    for (int i = 0; i < 3; ++i) g[i] = F[i] * k.im + t[i] + 0.5 * e[i];
    double R[9];
    rot_mat(q, R);
    mat3_tmul(R, g, dv);
}

// One RK4 step of the full 13-state.                                   sys_model.py:150-158
FT_HD void rk4_step(const DynConsts& k, const double* x, const double* Wr, double* xn) {
    double s[13], kk[13], acc[13];
    const double cs[4] = {0.0, 0.5 * k.dt, 0.5 * k.dt, k.dt};
    const double bs[4] = {k.dt / 6.0, k.dt / 3.0, k.dt / 3.0, k.dt / 6.0};
    for (int i = 0; i < 13; ++i) { acc[i] = x[i]; kk[i] = 0.0; }
    for (int st = 0; st < 4; ++st) {
        for (int i = 0; i < 13; ++i) s[i] = x[i] + cs[st] * kk[i];
        dyn_f(k, s + 3, s + 6, s + 9, Wr, kk, kk + 3, kk + 6, kk + 9, nullptr);
        for (int i = 0; i < 13; ++i) acc[i] += bs[st] * kk[i];
    }
    for (int i = 0; i < 13; ++i) xn[i] = acc[i];
}

// ---- first-order tangent of f ------------------------------------------------------------
// direction: (dv_in, dw_in, dq_in, dW) -> (dp', dv', dw', dq')
FT_HD void dyn_jvp(const DynConsts& k, const double* w, const double* q, const double* g,
                   const double* tv, const double* tw, const double* tq, const double* tW,
                   double* op, double* ov, double* ow, double* oq) {
    double c[3], e[3], t[3], tg[3];
    gyro_bilinear(k, w, tw, c);
    for (int i = 0; i < 3; ++i) ow[i] = (tW[3 + i] - c[i]) * k.iJ[i];
    centri_bilinear(k, w, tw, e);
    cross3(ow, k.r, t);
    for (int i = 0; i < 3; ++i) tg[i] = tW[i] * k.im + t[i] + e[i];
    double R[9], Bm[9], a[3], b[3];
    rot_mat(q, R);
    rot_bilinear(q, tq, Bm);
    mat3_tmul(R, tg, a);
    mat3_tmul(Bm, g, b);
    for (int i = 0; i < 3; ++i) ov[i] = a[i] + b[i];
    double o1[4], o2[4];
    omega_apply(tw, q, o1);
    omega_apply(w, tq, o2);
    for (int i = 0; i < 4; ++i) oq[i] = 0.5 * (o1[i] + o2[i]);
    for (int i = 0; i < 3; ++i) op[i] = tv[i];
}

// ---- adjoint of f (vector-Jacobian product) -----------------------------------------------
// m = a_p.v + (Rot(q) a_v).g + a_w.wdot + 1/2 a_q.Omega(w) q
// outputs: m_v(3)=a_p, m_w(3), m_q(4), m_F(3), m_tau(3).  Also returns h=Rot(q)a_v and beta=J^-1(a_w + r x h).
FT_HD void dyn_vjp(const DynConsts& k, const double* w, const double* q, const double* g,
                   const double* av, const double* aw, const double* aq,
                   double* mw, double* mq, double* mF, double* mT, double* h_out, double* beta_out) {
    double R[9], h[3], rxh[3], beta[3];
    rot_mat(q, R);
    mat3_mul(R, av, h);
    cross3(k.r, h, rxh);
    for (int i = 0; i < 3; ++i) beta[i] = (aw[i] + rxh[i]) * k.iJ[i];
    for (int i = 0; i < 3; ++i) { mF[i] = h[i] * k.im; mT[i] = beta[i]; }
    double ct[3], et[3];
    gyro_bilinear_T(k, w, beta, ct);
    centri_bilinear_T(k, w, h, et);
    for (int j = 0; j < 3; ++j) {
        double ej[3] = {0, 0, 0};
        ej[j] = 1.0;
        double oq[4];
        omega_apply(ej, q, oq);
        mw[j] = -ct[j] + et[j] + 0.5 * (aq[0] * oq[0] + aq[1] * oq[1] + aq[2] * oq[2] + aq[3] * oq[3]);
    }
    double qg[4], oa[4];
    quat_grad(q, av, g, qg);
    omega_apply(w, aq, oa);
    for (int j = 0; j < 4; ++j) mq[j] = qg[j] - 0.5 * oa[j];
    if (h_out) for (int i = 0; i < 3; ++i) { h_out[i] = h[i]; beta_out[i] = beta[i]; }
}

// ---- directional derivative of dyn_vjp in direction (tw,tq,tW) with the adjoint held fixed --
// = (Hessian of m) * direction, restricted to z-space.  tg = directional derivative of g (from dyn_jvp algebra).
FT_HD void dyn_hvp(const DynConsts& k, const double* w, const double* q, const double* g,
                   const double* av, const double* aq, const double* h, const double* beta,
                   const double* tw, const double* tq, const double* tW,
                   double* ow, double* oq, double* oF, double* oT) {
    // tangents of the intermediate quantities
    double c[3], e[3], t[3], twd[3], tg[3];
    gyro_bilinear(k, w, tw, c);
    for (int i = 0; i < 3; ++i) twd[i] = (tW[3 + i] - c[i]) * k.iJ[i];
    centri_bilinear(k, w, tw, e);
    cross3(twd, k.r, t);
    for (int i = 0; i < 3; ++i) tg[i] = tW[i] * k.im + t[i] + e[i];
    double Bm[9], th[3], rxth[3], tbeta[3];
    rot_bilinear(q, tq, Bm);
    mat3_mul(Bm, av, th);
    cross3(k.r, th, rxth);
    for (int i = 0; i < 3; ++i) tbeta[i] = rxth[i] * k.iJ[i];
    for (int i = 0; i < 3; ++i) { oF[i] = th[i] * k.im; oT[i] = tbeta[i]; }
    double c1[3], c2[3], e1[3], e2[3];
    gyro_bilinear_T(k, w, tbeta, c1);
    gyro_bilinear_T(k, tw, beta, c2);
    centri_bilinear_T(k, w, th, e1);
    centri_bilinear_T(k, tw, h, e2);
    for (int j = 0; j < 3; ++j) {
        double ej[3] = {0, 0, 0};
        ej[j] = 1.0;
        double o1[4];
        omega_apply(ej, tq, o1);
        ow[j] = -c1[j] - c2[j] + e1[j] + e2[j]
              + 0.5 * (aq[0] * o1[0] + aq[1] * o1[1] + aq[2] * o1[2] + aq[3] * o1[3]);
    }
    double g1[4], g2[4], oa[4];
    quat_grad(q, av, tg, g1);
    quat_grad(tq, av, g, g2);
    omega_apply(tw, aq, oa);
    for (int j = 0; j < 4; ++j) oq[j] = g1[j] + g2[j] - 0.5 * oa[j];
}

// ------------------------------------------------------------------------------------------
// Per-(stage, column) sweep: column `col` (0..12, z-space order [w q F tau]) of
//   [A_t | B_t]  (first order; 13 rows)  and, when lam != nullptr, of
//   W_t = Hessian of lam^T RK4(x,W) w.r.t. z-space (13 entries).
// x = x_t (13), Wr = total wrench.  Outputs: jac_col[13] = d x_{t+1} / d z_col ;  hess_col[13].
// ------------------------------------------------------------------------------------------
FT_HD void rk4_column(const DynConsts& k, const double* x, const double* Wr, int col,
                      const double* lam, double* jac_col, double* hess_col) {
    const double cs[4] = {0.0, 0.5 * k.dt, 0.5 * k.dt, k.dt};
    const double bs[4] = {k.dt / 6.0, k.dt / 3.0, k.dt / 3.0, k.dt / 6.0};
    // nominal stage states (v,w,q) and body accelerations, tangents of (v,w,q)
    double sw[4][3], sq[4][4], sg[4][3];
    double tws[4][3], tqs[4][4];
    // unit tangent by comparison, not by a run-time index: a dynamically indexed local array lives in local memory
    double tW[6], tx0[13];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 13; ++i) tx0[i] = (i == 6 + col) ? 1.0 : 0.0;       // col >= 7 never matches
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) tW[i] = (i == col - 7) ? 1.0 : 0.0;
    double kn[13], kt[13], accn[13], acct[13];
    for (int i = 0; i < 13; ++i) { kn[i] = 0.0; kt[i] = 0.0; acct[i] = tx0[i]; accn[i] = x[i]; }
    for (int st = 0; st < 4; ++st) {
        double s[13], ts[13];
        for (int i = 0; i < 13; ++i) { s[i] = x[i] + cs[st] * kn[i]; ts[i] = tx0[i] + cs[st] * kt[i]; }
        for (int i = 0; i < 3; ++i) { sw[st][i] = s[6 + i]; tws[st][i] = ts[6 + i]; }
        for (int i = 0; i < 4; ++i) { sq[st][i] = s[9 + i]; tqs[st][i] = ts[9 + i]; }
        dyn_f(k, s + 3, s + 6, s + 9, Wr, kn, kn + 3, kn + 6, kn + 9, sg[st]);
        dyn_jvp(k, s + 6, s + 9, sg[st], ts + 3, ts + 6, ts + 9, tW, kt, kt + 3, kt + 6, kt + 9);
        for (int i = 0; i < 13; ++i) { accn[i] += bs[st] * kn[i]; acct[i] += bs[st] * kt[i]; }
    }
    for (int i = 0; i < 13; ++i) jac_col[i] = acct[i];
    if (!lam) return;
    // reverse sweeps: nominal adjoints a_i (v,w,q parts) and their tangents (w,q parts only; the
    // p and v adjoint components do not depend on the linearisation point).
    double hw[3] = {0, 0, 0}, hq[4] = {0, 0, 0, 0}, hF[3] = {0, 0, 0}, hT[3] = {0, 0, 0};
    double psi_v[3] = {0, 0, 0}, psi_w[3] = {0, 0, 0}, psi_q[4] = {0, 0, 0, 0};     // nominal psi_{i+1}
    double dps_w[3] = {0, 0, 0}, dps_q[4] = {0, 0, 0, 0};                           // tangent of psi_{i+1}
    for (int st = 3; st >= 0; --st) {
        const double cnext = (st == 3) ? 0.0 : cs[st + 1];
        double av[3], aw[3], aq[4], daw[3], daq[4];
        for (int i = 0; i < 3; ++i) {
            av[i] = bs[st] * lam[3 + i] + cnext * psi_v[i];
            aw[i] = bs[st] * lam[6 + i] + cnext * psi_w[i];
            daw[i] = cnext * dps_w[i];
        }
        for (int i = 0; i < 4; ++i) { aq[i] = bs[st] * lam[9 + i] + cnext * psi_q[i]; daq[i] = cnext * dps_q[i]; }
        const double ap[3] = {bs[st] * lam[0], bs[st] * lam[1], bs[st] * lam[2]};
        double mw[3], mq[4], mF[3], mT[3], h[3], beta[3];
        dyn_vjp(k, sw[st], sq[st], sg[st], av, aw, aq, mw, mq, mF, mT, h, beta);
        // second-order part: Hessian-vector product + adjoint-tangent propagated through f_x^T, f_W^T
        double ow[3], oq[4], oF[3], oT[3];
        dyn_hvp(k, sw[st], sq[st], sg[st], av, aq, h, beta, tws[st], tqs[st], tW, ow, oq, oF, oT);
        const double zero3[3] = {0, 0, 0};
        double nw[3], nq[4], nF[3], nT[3];
        dyn_vjp(k, sw[st], sq[st], sg[st], zero3, daw, daq, nw, nq, nF, nT, nullptr, nullptr);
        for (int i = 0; i < 3; ++i) {
            psi_v[i] = ap[i]; psi_w[i] = mw[i];
            dps_w[i] = ow[i] + nw[i];
            hw[i] += dps_w[i]; hF[i] += oF[i] + nF[i]; hT[i] += oT[i] + nT[i];
        }
        for (int i = 0; i < 4; ++i) { psi_q[i] = mq[i]; dps_q[i] = oq[i] + nq[i]; hq[i] += dps_q[i]; }
    }
    for (int i = 0; i < 3; ++i) { hess_col[i] = hw[i]; hess_col[7 + i] = hF[i]; hess_col[10 + i] = hT[i]; }
    for (int i = 0; i < 4; ++i) hess_col[3 + i] = hq[i];
}

// ------------------------------------------------------------------------------------------
// Split form of rk4_column for the CUDA block path: the forward sweep parks the per-stage data the
// reverse sweep needs in (shared) memory instead of registers, the nominal stage data once per stage
// t, the tangents once per (t, column).  `st` strides let the caller choose a conflict-free layout.
//   nom  : [4][10]   stage (w q g), contiguous per stage t
//   tang : [4*7]     tangent of (w q) at the four RK stages, element e at tang[e * tstride]
// ------------------------------------------------------------------------------------------
FT_HD void rk4_col_forward(const DynConsts& k, const double* x, const double* Wr, int col, double* jac_col,
                           double* nom /* or nullptr */, double* tang, int tstride) {
    const double cs[4] = {0.0, 0.5 * k.dt, 0.5 * k.dt, k.dt};
    const double bs[4] = {k.dt / 6.0, k.dt / 3.0, k.dt / 3.0, k.dt / 6.0};
    // unit tangent by comparison, not by a run-time index: a dynamically indexed local array lives in local memory
    double tW[6], tx0[13];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 13; ++i) tx0[i] = (i == 6 + col) ? 1.0 : 0.0;       // col >= 7 never matches
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) tW[i] = (i == col - 7) ? 1.0 : 0.0;
    double kn[13], kt[13], accn[13], acct[13];
    for (int i = 0; i < 13; ++i) { kn[i] = 0.0; kt[i] = 0.0; acct[i] = tx0[i]; accn[i] = x[i]; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int st = 0; st < 4; ++st) {
        double s[13], ts[13], g[3];
        for (int i = 0; i < 13; ++i) { s[i] = x[i] + cs[st] * kn[i]; ts[i] = tx0[i] + cs[st] * kt[i]; }
        for (int i = 0; i < 7; ++i) tang[(size_t)(st * 7 + i) * tstride] = ts[6 + i];
        dyn_f(k, s + 3, s + 6, s + 9, Wr, kn, kn + 3, kn + 6, kn + 9, g);
        if (nom) {
            for (int i = 0; i < 7; ++i) nom[st * 10 + i] = s[6 + i];
            for (int i = 0; i < 3; ++i) nom[st * 10 + 7 + i] = g[i];
        }
        dyn_jvp(k, s + 6, s + 9, g, ts + 3, ts + 6, ts + 9, tW, kt, kt + 3, kt + 6, kt + 9);
        for (int i = 0; i < 13; ++i) { accn[i] += bs[st] * kn[i]; acct[i] += bs[st] * kt[i]; }
    }
    for (int i = 0; i < 13; ++i) jac_col[i] = acct[i];
}

// Jacobian columns of the three force inputs in closed form (the (w, q) tangents stay zero):
//   d v+ / d F = sum_i b_i Rot(q_i)^T / m,   d p+ / d F = sum_i b_i c_i (d v-stage tangent),  jacF[j][13]
FT_HD void rk4_force_columns(const DynConsts& k, const double* nom, double* jacF) {
    const double cs[4] = {0.0, 0.5 * k.dt, 0.5 * k.dt, k.dt};
    const double bs[4] = {k.dt / 6.0, k.dt / 3.0, k.dt / 3.0, k.dt / 6.0};
    double kv_prev[9];
    for (int i = 0; i < 39; ++i) jacF[i] = 0.0;
    for (int i = 0; i < 9; ++i) kv_prev[i] = 0.0;
    for (int st = 0; st < 4; ++st) {
        double R[9];
        rot_mat(nom + st * 10 + 3, R);
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) {
                const double kv = R[3 * j + i] * k.im;         // (Rot^T e_j)_i = R[j][i]
                jacF[j * 13 + i] += bs[st] * cs[st] * kv_prev[3 * j + i];
                jacF[j * 13 + 3 + i] += bs[st] * kv;
                kv_prev[3 * j + i] = kv;
            }
    }
}

// Reverse sweep: column `col` of the Hessian of lam^T RK4(x, W) in z-space from the parked forward data.
FT_HD void rk4_col_reverse(const DynConsts& k, const double* nom, const double* tang, int tstride, int col,
                           const double* lam, double* hess_col) {
    double tW[6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) tW[i] = (i == col - 7) ? 1.0 : 0.0;
    double hw[3] = {0, 0, 0}, hq[4] = {0, 0, 0, 0}, hF[3] = {0, 0, 0}, hT[3] = {0, 0, 0};
    double psi_v[3] = {0, 0, 0}, psi_w[3] = {0, 0, 0}, psi_q[4] = {0, 0, 0, 0};
    double dps_w[3] = {0, 0, 0}, dps_q[4] = {0, 0, 0, 0};
    // the stage loop is deliberately NOT unrolled on the device: one stage already fills the register file
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int st = 3; st >= 0; --st) {
        const double b_st = (st == 0 || st == 3) ? k.dt / 6.0 : k.dt / 3.0;          // RK4 weights
        const double cnext = (st == 3) ? 0.0 : ((st == 2) ? k.dt : 0.5 * k.dt);      // node of the NEXT stage
        double sw[3], sq[4], sg[3], tws[3], tqs[4];
        for (int i = 0; i < 3; ++i) { sw[i] = nom[st * 10 + i]; sg[i] = nom[st * 10 + 7 + i]; tws[i] = tang[(size_t)(st * 7 + i) * tstride]; }
        for (int i = 0; i < 4; ++i) { sq[i] = nom[st * 10 + 3 + i]; tqs[i] = tang[(size_t)(st * 7 + 3 + i) * tstride]; }
        double av[3], aw[3], aq[4], daw[3], daq[4];
        for (int i = 0; i < 3; ++i) {
            av[i] = b_st * lam[3 + i] + cnext * psi_v[i];
            aw[i] = b_st * lam[6 + i] + cnext * psi_w[i];
            daw[i] = cnext * dps_w[i];
        }
        for (int i = 0; i < 4; ++i) { aq[i] = b_st * lam[9 + i] + cnext * psi_q[i]; daq[i] = cnext * dps_q[i]; }
        const double ap[3] = {b_st * lam[0], b_st * lam[1], b_st * lam[2]};
        double mw[3], mq[4], mF[3], mT[3], h[3], beta[3];
        dyn_vjp(k, sw, sq, sg, av, aw, aq, mw, mq, mF, mT, h, beta);
        double ow[3], oq[4], oF[3], oT[3];
        dyn_hvp(k, sw, sq, sg, av, aq, h, beta, tws, tqs, tW, ow, oq, oF, oT);
        const double zero3[3] = {0, 0, 0};
        double nw[3], nq[4], nF[3], nT[3];
        dyn_vjp(k, sw, sq, sg, zero3, daw, daq, nw, nq, nF, nT, nullptr, nullptr);
        for (int i = 0; i < 3; ++i) {
            psi_v[i] = ap[i]; psi_w[i] = mw[i];
            dps_w[i] = ow[i] + nw[i];
            hw[i] += dps_w[i]; hF[i] += oF[i] + nF[i]; hT[i] += oT[i] + nT[i];
        }
        for (int i = 0; i < 4; ++i) { psi_q[i] = mq[i]; dps_q[i] = oq[i] + nq[i]; hq[i] += dps_q[i]; }
    }
    for (int i = 0; i < 3; ++i) { hess_col[i] = hw[i]; hess_col[7 + i] = hF[i]; hess_col[10 + i] = hT[i]; }
    for (int i = 0; i < 4; ++i) hess_col[3 + i] = hq[i];
}

// Robot state [p v q w] -> orbit-centre state [p_c v_c w q]             spiral_model.py:91-109
FT_HD void robot_to_center(const DynConsts& k, const double* x, double* c) {
    const double* q = x + 6;
    const double* w = x + 10;
    double R[9], a[3], wxr[3], b[3];
    rot_mat(q, R);
    mat3_tmul(R, k.r, a);
    cross3(w, k.r, wxr);
    mat3_tmul(R, wxr, b);
    for (int i = 0; i < 3; ++i) { c[i] = x[i] + a[i]; c[3 + i] = x[3 + i] + b[i]; c[6 + i] = w[i]; }
    for (int i = 0; i < 4; ++i) c[9 + i] = q[i];
}

}  // namespace ftmpc
