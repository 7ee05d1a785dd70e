// CPU port of the GPU algorithm -- TEST / BASELINE INFRASTRUCTURE ONLY (lives under oracle/).
//
// Compiles the product's own per-instance headers (fault-tolerant-mpc_b200/csrc/*.cuh) with the
// SerialBlock instantiation: one host thread per MPC instance, OpenMP over the batch.  It exists so
// that (a) the algorithm can be debugged and checked against oracle/ftmpc_oracle.py without a GPU and
// (b) bench.py can time "the same solve on the host cores" (cpu_baseline.kind = "port").
// The product library (libftmpc.so) never links or loads this file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#define FTMPC_DEBUG_COUNTERS 1
namespace ftmpc { long g_ftmpc_dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long g_ftmpc_dbg2[8] = {0, 0, 0, 0, 0, 0, 0, 0}; thread_local int g_ftmpc_warm_hit = 0; long g_ftmpc_qmax_hist[16] = {0}; }
#include "ftmpc_alloc.cuh"
using namespace ftmpc;

extern "C" {

int ftmpc_cpu_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

size_t ftmpc_cpu_workspace_doubles(int N, int batch) { return ws_layout(N).stride * (size_t)batch; }

int ftmpc_cpu_step(const ftmpc_config* cfg, const double* hull_table, int batch, const double* state,
                   const double* xref, const double* uref, const uint16_t* fault_mask, const double* fault_force,
                   const int32_t* hull_idx, int warm, double* z_warm, double* thrust, double* u0,
                   uint32_t* active_set, int32_t* status, int32_t* iters, double* cost, double* ws, int nthreads) {
    const WsLayout L = ws_layout(cfg->horizon);
    std::vector<double> own;
    if (!ws) { own.resize(L.stride * (size_t)batch); ws = own.data(); }
    StepIO io{batch, state, xref, uref, fault_mask, fault_force, hull_idx, hull_table, warm, z_warm, thrust, u0,
              active_set, status, iters, cost, nullptr, nullptr, cfg, ws, nullptr, nullptr, nullptr, 0, 0};
    stepio_default_strides(io, cfg->horizon);
    const size_t sdoubles = qp_scratch_doubles(cfg->horizon);
    const bool trace = std::getenv("FTMPC_TRACE") != nullptr;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        std::vector<double> scratch(sdoubles);
        SerialBlock blk;
#pragma omp for schedule(dynamic, 1)
        for (int inst = 0; inst < batch; ++inst) {
            if (!instance_input_ok(*cfg, io, inst)) { phase_out_invalid(blk, L, io, inst); continue; }
            phase_ls(*cfg, L, io, inst, inst, 1);
            for (int it = 0; it < cfg->max_sqp_iter; ++it) {
                if (ws[(size_t)inst * L.stride + L.oSc + SC_STATUS] != FTMPC_ST_RUNNING) break;
                phase_lin(blk, *cfg, L, io, inst, inst);
                phase_qp(blk, *cfg, L, io, inst, inst, scratch.data());
                phase_ls(*cfg, L, io, inst, inst, 0);
                if (trace) {
                    const double* sc = ws + (size_t)inst * L.stride + L.oSc;
                    std::printf("inst %d it %2d f %.9f csum %.3e cmax %.3e nu %.3g theta %.3g dmax %.3e delta %.3e lammax %.3g alpha %.3g qpit %g nact %g cholfail %g qpst %g st %g\n",
                                inst, it, sc[SC_F], sc[SC_CSUM], sc[SC_CMAX], sc[SC_NU], sc[SC_THETA], sc[SC_DMAX], sc[SC_DELTA],
                                sc[SC_LAMMAX], sc[SC_ALPHA], sc[SC_QPIT], sc[SC_NACT], sc[SC_CHOLFAIL], sc[SC_QPST], sc[SC_STATUS]);
                }
            }
            phase_out_write(blk, *cfg, L, io, inst, inst);
            phase_alloc(*cfg, L, io, inst);
        }
    }
    return 0;
}

// ---- stage entry points (mirror include/ftmpc.h) ------------------------------------------------------
int ftmpc_cpu_rk4_jac(const ftmpc_config* cfg, int batch, double* x, const double* wrench, double* jac,
                      const double* lam, double* hess) {
    const int N = cfg->horizon;
    const DynConsts k = dyn_consts(*cfg);
    for (int b = 0; b < batch; ++b) {
        double* X = x + (size_t)b * (N + 1) * 13;
        for (int t = 0; t < N; ++t) {
            const double* Wr = wrench + ((size_t)b * N + t) * 6;
            rk4_step(k, X + t * 13, Wr, X + (t + 1) * 13);
        }
        for (int t = 0; t < N; ++t)
            for (int c = 0; c < 13; ++c) {
                const double* Wr = wrench + ((size_t)b * N + t) * 6;
                double jc[13], hc[13];
                rk4_column(k, X + t * 13, Wr, c, lam ? lam + ((size_t)b * (N + 1) + t + 1) * 13 : nullptr, jc, hc);
                for (int i = 0; i < 13; ++i) {
                    jac[(((size_t)b * N + t) * 13 + c) * 13 + i] = jc[i];
                    if (lam && hess) hess[(((size_t)b * N + t) * 13 + c) * 13 + i] = hc[i];
                }
            }
    }
    return 0;
}

int ftmpc_cpu_terminal(const ftmpc_config* cfg, int batch, const double* e, double* V, double* grad, double* hess) {
    for (int b = 0; b < batch; ++b) V[b] = terminal_eval(*cfg, e + b * 9, grad + b * 9, hess + b * 81);
    return 0;
}

int ftmpc_cpu_allocate(const ftmpc_config* cfg, int batch, const double* u_des, const double* ub, double* thrust,
                       int32_t* status) {
    for (int b = 0; b < batch; ++b) status[b] = allocate_thrust(*cfg, u_des + b * 6, ub + b * 16, thrust + b * 16);
    return 0;
}

int ftmpc_cpu_clip(const ftmpc_config* cfg, int batch, const double* hull_table, const int32_t* hull_idx, const double* u,
                   double* out, int32_t* status) {
    for (int b = 0; b < batch; ++b)
        status[b] = clip_to_hull(*cfg, hull_table + (size_t)hull_idx[b] * FTMPC_HULL_STRIDE, u + b * 6, out + b * 6);
    return 0;
}

void ftmpc_cpu_qmax_hist(long* out) { for (int i = 0; i < 16; ++i) out[i] = ftmpc::g_ftmpc_qmax_hist[i]; }

void ftmpc_cpu_debug_counters(long* out, int reset) {
    for (int i = 0; i < 8; ++i) { out[i] = ftmpc::g_ftmpc_dbg[i]; if (reset) ftmpc::g_ftmpc_dbg[i] = 0; }
    for (int i = 0; i < 8; ++i) { out[8 + i] = ftmpc::g_ftmpc_dbg2[i]; if (reset) ftmpc::g_ftmpc_dbg2[i] = 0; }
}

// read back per-instance scalars of the workspace (diagnostics)
void ftmpc_cpu_scalars(int N, const double* ws, int inst, double* out) {
    const WsLayout L = ws_layout(N);
    std::memcpy(out, ws + (size_t)inst * L.stride + L.oSc, SC_COUNT * sizeof(double));
}
}
