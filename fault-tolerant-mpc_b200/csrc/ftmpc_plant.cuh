// ftmpc_plant.cuh -- the 16-thruster plant (robot state [p v q w]) used by the closed-loop driver.
//
// Replaces (reference):
//   SystemModel.dx_dt              ft_mpc/models/sys_model.py:177-226
//   SystemModel.rk4_integrator     ft_mpc/models/sys_model.py:138-162
//   SystemModel.normalize_quaternion   ft_mpc/models/sys_model.py:164-175
//   SimulationEnvironment.step     ft_mpc/simulation/sim_env.py:77-99  (noise is an INPUT tensor here;
//                                  the reference draws unseeded np.random.uniform(0,1e-3), :88-91)
#pragma once
#include "ftmpc_sqp.cuh"

namespace ftmpc {

FT_HD void plant_f(const ftmpc_config& c, const double* x, const double* gen, double* dx) {
    const double* v = x + 3;
    const double* q = x + 6;
    const double* w = x + 10;
    double R[9], a[3];
    rot_mat(q, R);
    mat3_tmul(R, gen, a);                                         // RotInv(q) F        sys_model.py:219
    for (int i = 0; i < 3; ++i) { dx[i] = v[i]; dx[3 + i] = a[i] / c.mass; }
    double oq[4];
    omega_apply(w, q, oq);
    for (int i = 0; i < 4; ++i) dx[6 + i] = 0.5 * oq[i];          // :222
    const double Jw[3] = {c.inertia[0] * w[0], c.inertia[1] * w[1], c.inertia[2] * w[2]};
    double cr[3];
    cross3(w, Jw, cr);
    for (int i = 0; i < 3; ++i) dx[10 + i] = (gen[3 + i] - cr[i]) / c.inertia[i];   // :225-227
}

// one closed-loop plant step: failed thrusters zeroed, stuck-on force added, RK4, + noise, renormalise
FT_HD void plant_step(const ftmpc_config& c, const double* x, const double* thrust, uint16_t mask,
                      const double* fault_force, const double* noise, int normalize, double* xn) {
    double gen[FTMPC_NU];
    for (int i = 0; i < FTMPC_NU; ++i) {
        double a = 0.0;
        for (int j = 0; j < FTMPC_NTHR; ++j) {
            const double uj = ((mask >> j) & 1) ? 0.0 : thrust[j];            // sys_model.py:198-206
            a += c.D[i * FTMPC_NTHR + j] * (uj + fault_force[j]);             // :211
        }
        gen[i] = a;
    }
    double s[13], k[13], acc[13];
    const double cs[4] = {0.0, 0.5 * c.dt, 0.5 * c.dt, c.dt};
    const double bs[4] = {c.dt / 6.0, c.dt / 3.0, c.dt / 3.0, c.dt / 6.0};
    for (int i = 0; i < 13; ++i) { acc[i] = x[i]; k[i] = 0.0; }
    for (int st = 0; st < 4; ++st) {
        for (int i = 0; i < 13; ++i) s[i] = x[i] + cs[st] * k[i];
        plant_f(c, s, gen, k);
        for (int i = 0; i < 13; ++i) acc[i] += bs[st] * k[i];
    }
    if (noise) for (int i = 0; i < 13; ++i) acc[i] += noise[i];               // sim_env.py:88-91
    if (normalize) {
        const double nq = sqrt(acc[6] * acc[6] + acc[7] * acc[7] + acc[8] * acc[8] + acc[9] * acc[9]);
        for (int i = 6; i < 10; ++i) acc[i] /= nq;                            // sim_env.py:93
    }
    for (int i = 0; i < 13; ++i) xn[i] = acc[i];
}

}  // namespace ftmpc
