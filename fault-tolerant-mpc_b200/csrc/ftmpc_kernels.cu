// ftmpc_kernels.cu -- CUDA kernels (sm_100a) and the C ABI of include/ftmpc.h.
//
// Kernel map:
//   k_solve : ONE PERSISTENT CTA PER SM runs whole MPC solves back to back (dynamic instance queue):
//               K4 step acceptance (all backtracking step lengths rolled out at once by the lanes of a warp)
//               K1 RK4 Jacobians (+ costates and exact stage Hessians when the QP blends them in), one (stage, column) task per thread
//               K2+K3 Riccati factorisation of the condensed QP (J = Phi^-1 blkdiag(C_t^-T), ftmpc_riccati.cuh), dual active-set QP
//             every matrix of the QP lives in shared memory (224 KB at N = 20); the per-CTA iterate (U, X,
//             multipliers, stage Jacobians) sits in a per-CTA global slot that stays L1/L2 resident.
//             k_solve<true> (N > 20): per-CTA global scratch, operator-form QP (no matrix of size N^2), stage matrices
//             staged through shared memory by cp.async.bulk.
//   k_alloc : K5 clip + thrust allocation QP, one thread per instance
//   k_plant : K6 plant step for closed-loop rollouts
// There is no CPU fallback in this translation unit: without a CUDA device ftmpc_create fails.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "ftmpc.h"
#include "ftmpc_alloc.cuh"
#include "ftmpc_qp2.cuh"
#include "ftmpc_plant.cuh"
#include "ftmpc_hull.cuh"

using namespace ftmpc;


struct ftmpc_ctx {
    ftmpc_config cfg;          // host copy
    ftmpc_config* d_cfg;       // device copy
    double* d_hull;            // device hull table
    long long* d_prof;         // per-phase cycle accumulators (PH_COUNT) of the last profiled step
    int device, num_sms;
    size_t smem_optin, smem_dyn_max, smem_dyn2_max;
    int ctas_per_sm2;          // resident k_solve2 CTAs per SM for this horizon (occupancy query)
    WsLayout L;
    int profile, last_launches;
    cudaEvent_t ev[3];         // before k_solve / between / after k_alloc (profile mode)
};

// -------------------------------------------------------------------------------------------------
// kernels
// -------------------------------------------------------------------------------------------------
#define FTMPC_QP_THREADS 256
// GS = false: the CTA scratch is the dynamic shared memory.  The choice is a TEMPLATE parameter on purpose: with a
// run-time select the compiler cannot prove the address space and every scratch access becomes a generic LD/ST
// (longer latency than LDS/STS, and it lands on the long scoreboard).  GS = true (horizons whose QP does not fit in
// 227 KB) keeps the scratch in a per-CTA global slice.
template <bool GS>
__global__ void __launch_bounds__(FTMPC_QP_THREADS, 1)
    k_solve(const __grid_constant__ ftmpc_config cfg, WsLayout L, StepIO io, int* queue, double* gscratch, size_t sdoubles,
            long long* prof) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(16) double red[192];
    __shared__ int s_inst;
    __shared__ __align__(8) unsigned long long s_mbar;
    double* scratch = GS ? gscratch + (size_t)blockIdx.x * sdoubles : smem;
    __shared__ double s_tfv[2 * FTMPC_NF];
    __shared__ int s_tfi[2 * FTMPC_NF];
    CudaBlock blk(red, prof);
    __shared__ unsigned s_op_par;
    if (GS && threadIdx.x == 0) {                 // global-scratch kernel: the mbarrier belongs to the operator products (ric_apply_staged)
        s_op_par = 0u;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!GS) tma_bar_init(&s_mbar, blk.tma);      // per-instance data is staged by TMA bulk copies when the scratch is shared memory
    const int slot = blockIdx.x;
    const double* sc = ws_slot(io, L, slot) + L.oSc;
    // compact rows of the terminal set (A_f has at most two non-zeros per row; checked by ftmpc_create)
    for (int i = threadIdx.x; i < FTMPC_NF; i += blockDim.x) {
        int k = 0;
        s_tfv[2 * i] = s_tfv[2 * i + 1] = 0.0;
        s_tfi[2 * i] = s_tfi[2 * i + 1] = 0;
        for (int j = 0; j < FTMPC_NE; ++j) {
            const double a = io.cfg_g->Af[i * FTMPC_NE + j];
            if (a != 0.0 && k < 2) { s_tfv[2 * i + k] = a; s_tfi[2 * i + k] = j; ++k; }
        }
    }
    io.tf_val = s_tfv;
    io.tf_idx = s_tfi;
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) {
            const int k = atomicAdd(queue, 1);
            const int lim = io.batch_dev ? *io.batch_dev : io.batch;       // second pass behind k_solve2: device-side list
            s_inst = (k < lim) ? (io.inst_list ? io.inst_list[k] : k) : -1;
        }
        __syncthreads();
        const int inst = s_inst;
        __syncthreads();
        if (inst < 0) break;
        if (!instance_input_ok(cfg, io, inst)) {                   // uniform over the block
            phase_out_invalid(blk, L, io, inst);
            continue;
        }
        phase_ls_block(blk, cfg, L, io, inst, slot, 1, scratch);
        for (int it = 0; it < cfg.max_sqp_iter; ++it) {
            if (sc[SC_STATUS] != (double)FTMPC_ST_RUNNING) break;      // uniform: written before the last barrier
            // with the shared-memory scratch the linearisation leaves Jz / Wz where condensing reads them
            const bool staged = !GS && lin_scratch_doubles(L.N) <= (size_t)(L.nv + FTMPC_NE) * L.nv && condense_fast_path(L, blockDim.x);
            phase_lin(blk, cfg, L, io, inst, slot, lin_place_v1(scratch, L.N, staged), true);
            phase_qp(blk, cfg, L, io, inst, slot, scratch, staged, GS ? smem : nullptr,
                     GS ? (unsigned)__cvta_generic_to_shared(&s_mbar) : 0u, GS ? &s_op_par : nullptr);
            phase_ls_block(blk, cfg, L, io, inst, slot, 0, scratch);
        }
        phase_out_write(blk, cfg, L, io, inst, slot);
    }
}

// Two CTAs per SM (<= 113 KB of shared memory, <= 128 registers per thread): range-space QP on the packed extended inverse
// (ftmpc_qp2.cuh).  Horizons with (N + 2)(N + 3) / 2 <= 256 block threads, i.e. N <= 20.
__global__ void __launch_bounds__(FTMPC_QP_THREADS, 2)
    k_solve2(const __grid_constant__ ftmpc_config cfg, WsLayout L, StepIO io, int* queue, long long* prof) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(16) double red[192];
    __shared__ int s_inst;
    __shared__ double s_tfv[2 * FTMPC_NF];
    __shared__ int s_tfi[2 * FTMPC_NF];
    __shared__ __align__(8) unsigned long long s_mbar;
    double* scratch = smem;
    CudaBlock blk(red, prof);
    tma_bar_init(&s_mbar, blk.tma);
    const int slot = blockIdx.x;
    const double* sc = ws_slot(io, L, slot) + L.oSc;
    for (int i = threadIdx.x; i < FTMPC_NF; i += blockDim.x) {
        int k = 0;
        s_tfv[2 * i] = s_tfv[2 * i + 1] = 0.0;
        s_tfi[2 * i] = s_tfi[2 * i + 1] = 0;
        for (int j = 0; j < FTMPC_NE; ++j) {
            const double a = io.cfg_g->Af[i * FTMPC_NE + j];
            if (a != 0.0 && k < 2) { s_tfv[2 * i + k] = a; s_tfi[2 * i + k] = j; ++k; }
        }
    }
    io.tf_val = s_tfv;
    io.tf_idx = s_tfi;
    // the linearisation works in the front of the scratch and leaves Jz / Wz where condense2 reads them
    LinPlace lp;
    {
        const Qp2Scratch q = qp2_carve(scratch, L.N, io);
        lp.work = scratch;
        lp.JzS = scratch + (size_t)L.N * 326;
        lp.tail = lp.JzS + (size_t)L.N * 169;
        lp.WzS = q.Wz;
    }
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) s_inst = atomicAdd(queue, 1);
        __syncthreads();
        const int inst = s_inst;
        __syncthreads();
        if (inst >= io.batch) break;
        if (!instance_input_ok(cfg, io, inst)) {
            phase_out_invalid(blk, L, io, inst);
            continue;
        }
        phase_ls_block(blk, cfg, L, io, inst, slot, 1, scratch);
        for (int it = 0; it < cfg.max_sqp_iter; ++it) {
            if (sc[SC_STATUS] != (double)FTMPC_ST_RUNNING) break;
            phase_lin(blk, cfg, L, io, inst, slot, lp);
            phase_qp2(blk, cfg, L, io, inst, slot, scratch, true);
            phase_ls_block(blk, cfg, L, io, inst, slot, 0, scratch);
        }
        phase_out_write(blk, cfg, L, io, inst, slot);
    }
}

__global__ void __launch_bounds__(64) k_alloc(const __grid_constant__ ftmpc_config cfg, WsLayout L, StepIO io) {
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= io.batch) return;
    phase_alloc(cfg, L, io, inst);
}

// ---- stage kernels --------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_rk4_jac(const ftmpc_config* __restrict__ cfg, int batch, double* x,
                                                const double* __restrict__ wrench, double* jac,
                                                const double* __restrict__ lam, double* hess) {
    const int inst = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (inst >= batch) return;
    const int lane = threadIdx.x & 31, N = cfg->horizon;
    const DynConsts k = dyn_consts(*cfg);
    double* X = x + (size_t)inst * (N + 1) * 13;
    const double* W = wrench + (size_t)inst * N * 6;
    // rollout: every lane carries the state in registers, lane 0 stores it
    double xs[13], xn[13];
    for (int i = 0; i < 13; ++i) xs[i] = X[i];
    for (int t = 0; t < N; ++t) {
        rk4_step(k, xs, W + t * 6, xn);
        for (int i = 0; i < 13; ++i) xs[i] = xn[i];
        if (lane == 0) for (int i = 0; i < 13; ++i) X[(t + 1) * 13 + i] = xn[i];
    }
    __syncwarp();
    for (int it = lane; it < N * 13; it += 32) {
        const int t = it / 13, c = it % 13;
        double jc[13], hc[13];
        rk4_column(k, X + t * 13, W + t * 6, c, lam ? lam + ((size_t)inst * (N + 1) + t + 1) * 13 : nullptr, jc, hc);
        double* jo = jac + ((size_t)inst * N * 13 + it) * 13;
        for (int i = 0; i < 13; ++i) jo[i] = jc[i];
        if (lam && hess) {
            double* ho = hess + ((size_t)inst * N * 13 + it) * 13;
            for (int i = 0; i < 13; ++i) ho[i] = hc[i];
        }
    }
}

__global__ void k_r2c(const ftmpc_config* __restrict__ cfg, int batch, const double* state, double* center) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    robot_to_center(dyn_consts(*cfg), state + (size_t)i * 13, center + (size_t)i * 13);
}

__global__ void k_terminal(const ftmpc_config* __restrict__ cfg, int batch, const double* e, double* V, double* grad,
                           double* hess) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    double g[9], H[81];
    V[i] = terminal_eval(*cfg, e + (size_t)i * 9, g, H);
    for (int k = 0; k < 9; ++k) grad[(size_t)i * 9 + k] = g[k];
    for (int k = 0; k < 81; ++k) hess[(size_t)i * 81 + k] = H[k];
}

__global__ void __launch_bounds__(FTMPC_QP_THREADS, 1)
    k_condense(const ftmpc_config* __restrict__ cfg, WsLayout L, int batch, const double* jac, const double* hess,
               const double* x, const double* u, const double* xref, const double* gradV, const double* hessV,
               double theta, double* H, double* g, double* gscratch, size_t sdoubles) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(16) double red[192];
    double* scratch = smem;
    (void)gscratch; (void)sdoubles;
    CudaBlock blk(red);
    const int N = L.N, n = L.n, ld = L.nv;
    for (int inst = blockIdx.x; inst < batch; inst += gridDim.x) {
        const QpScratch s = qp_carve(scratch, N, cfg);
        double* Jz = s.RS;
        double* Wz = s.RS + (size_t)N * 169;
        for (int i = threadIdx.x; i < N * 169; i += blockDim.x) {
            Jz[i] = jac[(size_t)inst * N * 169 + i];
            Wz[i] = hess ? hess[(size_t)inst * N * 169 + i] : 0.0;
        }
        __syncthreads();
        condense(blk, *cfg, L, s, Jz, Wz, x + (size_t)inst * (N + 1) * 13, u + (size_t)inst * n,
                 xref + (size_t)inst * (N + 1) * 9, gradV + (size_t)inst * 9, hessV + (size_t)inst * 81, theta, 0.0, nullptr);
        for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
            const int a = idx / n, b = idx % n;
            H[(size_t)inst * n * n + idx] = (b <= a) ? s.E[(size_t)a * ld + b] : s.E[(size_t)b * ld + a];
        }
        for (int a = threadIdx.x; a < n; a += blockDim.x) g[(size_t)inst * n + a] = s.g[a];
        __syncthreads();
    }
}

struct CsrCons {
    const int32_t* ptr;
    const int32_t* idx;
    const double* val;
    const double* b;
    __device__ __forceinline__ void row(int p, SparseRow& r) const {      // C x <= b  ->  -C x >= -b
        const int s = ptr[p];
        int k = ptr[p + 1] - s;
        if (k > FTMPC_GI_MAXNNZ) k = FTMPC_GI_MAXNNZ;
        for (int j = 0; j < k; ++j) { r.idx[j] = idx[s + j]; r.val[j] = -val[s + j]; }
        r.nnz = k;
        r.beta = -b[p];
    }
    __device__ __forceinline__ double slack(int p, const double* v, double sb) const { return cons_slack_generic(*this, p, v, sb); }
};

__global__ void __launch_bounds__(FTMPC_QP_THREADS, 1)
    k_qp_generic(int batch, int n, int m, const double* H, const double* g, const int32_t* ptr, const int32_t* idx,
                 const double* val, const double* b, double* x, double* lam, int32_t* status, int maxit, double tol) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(16) double red[192];
    CudaBlock blk(red);
    const int ld = n | 1, tid = threadIdx.x, nt = blockDim.x;
    double* p = smem;
    GiWork w;
    w.E = p; p += (size_t)n * ld;
    w.Ui = p; p += (size_t)n * (n + 1) / 2 + 1;
    w.xe = p; p += n;
    w.s = p; p += m;
    w.u = p; p += n + 2;
    w.d = p; p += n;
    w.ze = p; p += n;
    w.r = p; p += n;
    w.cs = p; p += 2 * n + 2;
    w.tmp = p; p += n + 2;
    w.sub = p; p += n + 2;
    double* dg = p; p += n;
    w.esign = p; p += 2;
    int* ip = reinterpret_cast<int*>(p);
    w.act = ip; ip += n + 2;
    w.pos = ip; ip += m;
    w.itmp = ip; ip += n + 2;
    for (int inst = blockIdx.x; inst < batch; inst += gridDim.x) {
        const double* Hi = H + (size_t)inst * n * n;
        const double* gi = g + (size_t)inst * n;
        for (int i = tid; i < n * n; i += nt) w.E[(size_t)(i / n) * ld + (i % n)] = Hi[i];
        __syncthreads();
        int st;
        if (chol_lower(blk, n, ld, w.E, 1e-300)) {
            st = 3;
        } else {
            tri_inv_transpose(blk, n, ld, w.E, dg);
            for (int i = tid; i < n; i += nt) {
                double v = 0.0;
                for (int r = 0; r <= i; ++r) v += w.E[(size_t)r * ld + i] * gi[r];
                w.d[i] = v;
            }
            __syncthreads();
            for (int row = tid; row < n; row += nt) {
                double v = 0.0;
                for (int k = 0; k < n; ++k) v += w.E[(size_t)row * ld + k] * w.d[k];
                w.xe[row] = -v;
            }
            __syncthreads();
            CsrCons cons{ptr, idx, val, b + (size_t)inst * m};
            int it = 0, na = 0;
            st = gi_solve(blk, cons, w, n, n, ld, m, 0, lam + (size_t)inst * m, maxit, tol, &it, &na);
            for (int i = tid; i < n; i += nt) x[(size_t)inst * n + i] = w.xe[i];
        }
        if (tid == 0) status[inst] = st;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(64) k_allocate(const ftmpc_config* __restrict__ cfg, int batch, const double* udes,
                                                const double* ub, double* thrust, int32_t* status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    double th[16];
    status[i] = allocate_thrust(*cfg, udes + (size_t)i * 6, ub + (size_t)i * 16, th);
    for (int j = 0; j < 16; ++j) thrust[(size_t)i * 16 + j] = th[j];
}

__global__ void __launch_bounds__(64) k_clip(const ftmpc_config* __restrict__ cfg, int batch, const double* hull_table,
                                             const int32_t* hull_idx, const double* u, double* out, int32_t* status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const int hi = hull_idx[i];
    double v[FTMPC_NU];
    if (hi < 0 || hi >= cfg->n_hull_sets) {
        for (int j = 0; j < FTMPC_NU; ++j) out[(size_t)i * FTMPC_NU + j] = u[(size_t)i * FTMPC_NU + j];
        status[i] = FTMPC_ST_BADINPUT;
        return;
    }
    status[i] = clip_to_hull(*cfg, hull_table + (size_t)hi * FTMPC_HULL_STRIDE, u + (size_t)i * FTMPC_NU, v);
    for (int j = 0; j < FTMPC_NU; ++j) out[(size_t)i * FTMPC_NU + j] = v[j];
}

__global__ void k_plant(const ftmpc_config* __restrict__ cfg, int batch, const double* state, const double* thrust,
                        const uint16_t* mask, const double* ff, const double* noise, int normalize, double* next) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    double xn[13];
    plant_step(*cfg, state + (size_t)i * 13, thrust + (size_t)i * 16, mask[i], ff + (size_t)i * 16,
               noise ? noise + (size_t)i * 13 : nullptr, normalize, xn);
    for (int k = 0; k < 13; ++k) next[(size_t)i * 13 + k] = xn[k];
}

// closed-loop driver: plant step in place (SimulationEnvironment.step, sim_env.py:85-93) + running totals of the loop
__global__ void k_plant_loop(const ftmpc_config* __restrict__ cfg, int batch, double* state, const double* thrust,
                             const uint16_t* mask, const double* ff, const double* noise, const int32_t* status,
                             const double* cost, double* cost_sum, int32_t* worst, int first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    double xn[13];
    plant_step(*cfg, state + (size_t)i * 13, thrust + (size_t)i * 16, mask[i], ff + (size_t)i * 16,
               noise ? noise + (size_t)i * 13 : nullptr, 1, xn);
    for (int k = 0; k < 13; ++k) state[(size_t)i * 13 + k] = xn[k];
    cost_sum[i] = (first ? 0.0 : cost_sum[i]) + cost[i];
    const int w = first ? 0 : worst[i];
    worst[i] = status[i] > w ? status[i] : w;
}

// FP64 FMA throughput probe (roofline denominator for bench.py: MEASURED_PEAKS.json has no fp64 figure).
// 8 independent DFMA chains per thread, 8 resident warps per scheduler.
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; ++i) {
        r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
        r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
    const double s = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
    if (s == 12345.6789) out[0] = s;
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
#define CU(x)                                   \
    do {                                        \
        cudaError_t e_ = (x);                   \
        if (e_ != cudaSuccess) return FTMPC_ERR_CUDA; \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Every entry point runs on the device the handle was created on, whatever the caller's current device is
// (a handle bound to cuda:1 must not launch on cuda:0 with device-1 pointers); the caller's device is restored.
struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(true) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define GUARD(h) DeviceGuard guard_((h)->device); if (!guard_.ok) return FTMPC_ERR_CUDA

static size_t qp_smem_bytes(int N) { return qp_scratch_doubles(N) * sizeof(double); }

// k_solve2 (two CTAs per SM, range-space QP) covers the horizons whose block sweep fits one 256-thread CTA
static size_t solve2_smem_bytes(int N) {
    size_t a = qp2_scratch_doubles(N), b = ls_scratch_doubles(N);
    if (b > a) a = b;
    return a * sizeof(double);
}
static bool use_solve2(const ftmpc_ctx* h) {
    const int N = h->cfg.horizon;
    return (h->cfg.qp_method & 3) != 0 && (N + 2) * (N + 3) / 2 <= FTMPC_QP_THREADS && 6 * N <= FTMPC_QP_THREADS &&
           solve2_smem_bytes(N) <= h->smem_dyn2_max;
}


extern "C" {

const char* ftmpc_strerror(int code) {
    switch (code) {
        case FTMPC_OK: return "ok";
        case FTMPC_ERR_ARG: return "invalid argument";
        case FTMPC_ERR_CUDA: return "CUDA error";
        case FTMPC_ERR_WORKSPACE: return "workspace too small";
        case FTMPC_ERR_UNSUPPORTED: return "unsupported configuration";
        case FTMPC_ERR_NO_DEVICE: return "no CUDA device (ft_mpc_b200 has no CPU fallback)";
        case FTMPC_ERR_NOMEM: return "out of host memory";
        default: return "unknown error";
    }
}

int ftmpc_create(ftmpc_handle* out, const ftmpc_config* cfg, const double* hull_table) {
    if (!out || !cfg || !hull_table) return FTMPC_ERR_ARG;
    if (cfg->horizon < 1 || cfg->horizon > 512 || cfg->n_hull_sets < 1 || cfg->n_poly > FTMPC_MAX_POLY ||
        cfg->n_root > FTMPC_MAX_ROOT || cfg->max_sqp_iter < 1)
        return FTMPC_ERR_ARG;
    if (cfg->dtype != 0) return FTMPC_ERR_UNSUPPORTED;
    for (int i = 0; i < FTMPC_NF; ++i) {               // structure the kernels rely on (true for the reference's terminal.yaml)
        int nz = 0;
        for (int j = 0; j < FTMPC_NE; ++j) nz += cfg->Af[i * FTMPC_NE + j] != 0.0;
        if (nz > 2) return FTMPC_ERR_UNSUPPORTED;
    }
    for (int k = 0; k < cfg->n_poly + cfg->n_root; ++k) {
        const int8_t* p = k < cfg->n_poly ? cfg->poly_e[k] : cfg->root_e[k - cfg->n_poly];
        int nvv = 0;
        for (int j = 0; j < FTMPC_NE; ++j) nvv += p[j] > 0;
        if (nvv > 3) return FTMPC_ERR_UNSUPPORTED;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return FTMPC_ERR_NO_DEVICE;
    ftmpc_ctx* h = new (std::nothrow) ftmpc_ctx;
    if (!h) return FTMPC_ERR_NOMEM;
    std::memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->L = ws_layout(cfg->horizon);
    const size_t hb = (size_t)cfg->n_hull_sets * FTMPC_HULL_STRIDE * sizeof(double);
    cudaDeviceProp prop;
    // any failure below releases everything acquired so far (ftmpc_destroy copes with a half-built handle)
#define CREATE_TRY(x) do { if ((x) != cudaSuccess) { ftmpc_destroy(h); return FTMPC_ERR_CUDA; } } while (0)
    CREATE_TRY(cudaGetDevice(&h->device));
    CREATE_TRY(cudaGetDeviceProperties(&prop, h->device));
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    CREATE_TRY(cudaMalloc(&h->d_cfg, sizeof(ftmpc_config)));
    CREATE_TRY(cudaMemcpy(h->d_cfg, cfg, sizeof(ftmpc_config), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMalloc(&h->d_hull, hb));
    CREATE_TRY(cudaMemcpy(h->d_hull, hull_table, hb, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMalloc(&h->d_prof, PH_COUNT * sizeof(long long)));
    CREATE_TRY(cudaMemset(h->d_prof, 0, PH_COUNT * sizeof(long long)));
    for (int i = 0; i < 3; ++i) CREATE_TRY(cudaEventCreate(&h->ev[i]));
    // dynamic shared memory opt-in, once per handle: the attribute is per function and per device, so it is raised to the
    // device maximum (every horizon whose scratch fits uses the same kernel instantiation)
    {
        const void* fns[4] = {(const void*)k_solve<false>, (const void*)k_condense, (const void*)k_qp_generic, (const void*)k_solve<true>};
        for (int i = 0; i < 4; ++i) {               // the opt-in limit covers static + dynamic shared memory of a block
            cudaFuncAttributes fa;
            CREATE_TRY(cudaFuncGetAttributes(&fa, fns[i]));
            const int dyn = (int)h->smem_optin - (int)fa.sharedSizeBytes;
            CREATE_TRY(cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
            if (i == 0) h->smem_dyn_max = (size_t)dyn;
        }
        cudaFuncAttributes fa;
        CREATE_TRY(cudaFuncGetAttributes(&fa, (const void*)k_solve2));
        h->smem_dyn2_max = h->smem_optin - fa.sharedSizeBytes;
        CREATE_TRY(cudaFuncSetAttribute((const void*)k_solve2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_dyn2_max));
        CREATE_TRY(cudaFuncSetAttribute((const void*)k_solve2, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        h->ctas_per_sm2 = 1;
        if (use_solve2(h)) {
            int nb = 0;
            CREATE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_solve2, FTMPC_QP_THREADS, solve2_smem_bytes(cfg->horizon)));
            h->ctas_per_sm2 = nb < 1 ? 1 : nb;
            if (const char* e = std::getenv("FTMPC_CTAS_PER_SM")) {      // experiments: fewer resident CTAs than the SM would take
                const int v = std::atoi(e);
                if (v >= 1 && v < h->ctas_per_sm2) h->ctas_per_sm2 = v;
            }
        }
    }
#undef CREATE_TRY
    *out = h;
    return FTMPC_OK;
}

int ftmpc_profile_enable(ftmpc_handle h, int enable) {
    if (!h) return FTMPC_ERR_ARG;
    h->profile = enable;
    return FTMPC_OK;
}

int ftmpc_last_launches(ftmpc_handle h) { return h ? h->last_launches : FTMPC_ERR_ARG; }

int ftmpc_profile_read(ftmpc_handle h, void* stream, double* kernel_ms, int64_t* phase_cycles, int n_phase) {
    if (!h || !kernel_ms) return FTMPC_ERR_ARG;
    GUARD(h);
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    kernel_ms[0] = kernel_ms[1] = 0.0;
    if (h->profile) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, h->ev[0], h->ev[1]));
        kernel_ms[0] = t;
        CU(cudaEventElapsedTime(&t, h->ev[1], h->ev[2]));
        kernel_ms[1] = t;
    }
    if (phase_cycles && n_phase > 0) {
        long long tmp[PH_COUNT];
        CU(cudaMemcpy(tmp, h->d_prof, sizeof(tmp), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n_phase; ++i) phase_cycles[i] = i < PH_COUNT ? (int64_t)tmp[i] : 0;
    }
    return FTMPC_OK;
}

void ftmpc_destroy(ftmpc_handle h) {
    if (!h) return;
    DeviceGuard guard_(h->device);
    for (int i = 0; i < 3; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    cudaFree(h->d_cfg);
    cudaFree(h->d_hull);
    cudaFree(h->d_prof);
    delete h;
}

int ftmpc_num_var(ftmpc_handle h) { return h ? h->L.n + (h->L.N + 1) * FTMPC_NX : FTMPC_ERR_ARG; }
int ftmpc_num_ineq(ftmpc_handle h) { return h ? h->L.mc : FTMPC_ERR_ARG; }

// shared-memory scratch of one CTA: the QP matrices, or the line-search rollouts (they alternate)
static size_t solve_smem_bytes(int N) {
    size_t a = qp_scratch_doubles(N), b = ls_scratch_doubles(N), c = lin_scratch_doubles(N);
    if (b > a) a = b;
    if (c > a) a = c;
    return a * sizeof(double);
}

static int solve_grid(const ftmpc_ctx* h, int batch) {
    const int cap = use_solve2(h) ? h->ctas_per_sm2 * h->num_sms : h->num_sms;
    return batch < cap ? batch : cap;
}

int ftmpc_workspace_bytes(ftmpc_handle h, int batch, size_t* out) {
    if (!h || !out || batch < 1) return FTMPC_ERR_ARG;
    // one iterate slot per resident CTA (not per instance), plus the CTA scratch when it does not fit in shared memory
    const int grid = solve_grid(h, batch);
    size_t b = align_up(h->L.stride * sizeof(double) * (size_t)grid, 256);
    const size_t need = solve_smem_bytes(h->cfg.horizon);
    if (use_solve2(h)) b += align_up(sizeof(int) * ((size_t)batch + 2), 256);        // ids handed over to the null-space kernel
    else if (need > h->smem_optin) b += align_up(need, 256) * (size_t)grid;
    b += 256;      // work-queue head of the launch: it lives in the caller's workspace, so steps with different workspaces
                   // may be in flight on different streams at the same time (pipelined batches)
    *out = b;
    return FTMPC_OK;
}

// enqueue memset(queue) + k_solve + k_alloc for one batch (no synchronisation)
static int launch_step(ftmpc_ctx* h, const StepIO& io, void* workspace, size_t need_ws, cudaStream_t stream) {
    const WsLayout L = h->L;
    const int grid = solve_grid(h, io.batch);
    if (use_solve2(h)) {
        // pass 1: k_solve2 (two CTAs per SM).  pass 2: the few instances whose working set outgrew its shared-memory capacity
        // are solved from scratch by the null-space kernel (device-side list, usually empty: the launch then ends at once)
        int* redo = (int*)((char*)workspace + align_up(L.stride * sizeof(double) * (size_t)grid, 256));
        int* queue = (int*)((char*)workspace + need_ws - 256);
        CU(cudaMemsetAsync(queue, 0, 2 * sizeof(int), stream));
        CU(cudaMemsetAsync(redo, 0, sizeof(int), stream));
        if (h->profile) {
            CU(cudaMemsetAsync(h->d_prof, 0, PH_COUNT * sizeof(long long), stream));
            CU(cudaEventRecord(h->ev[0], stream));
        }
        StepIO io1 = io;
        io1.redo = redo;
        k_solve2<<<grid, FTMPC_QP_THREADS, solve2_smem_bytes(h->cfg.horizon), stream>>>(h->cfg, L, io1, queue, h->profile ? h->d_prof : nullptr);
        StepIO io2 = io;
        io2.inst_list = redo + 1;
        io2.batch_dev = redo;
        const int grid1 = io.batch < h->num_sms ? io.batch : h->num_sms;
        k_solve<false><<<grid1 < grid ? grid1 : grid, FTMPC_QP_THREADS, solve_smem_bytes(h->cfg.horizon), stream>>>(
            h->cfg, L, io2, queue + 1, nullptr, 0, h->profile ? h->d_prof : nullptr);
        if (h->profile) CU(cudaEventRecord(h->ev[1], stream));
        k_alloc<<<(io.batch + 63) / 64, 64, 0, stream>>>(h->cfg, L, io);
        if (h->profile) CU(cudaEventRecord(h->ev[2], stream));
        h->last_launches = 3;
        CU(cudaGetLastError());
        return FTMPC_OK;
    }
    const size_t smem = solve_smem_bytes(h->cfg.horizon);
    const bool use_global = smem > h->smem_optin;
    double* gscratch = use_global ? (double*)((char*)workspace + align_up(L.stride * sizeof(double) * (size_t)grid, 256)) : nullptr;
    const size_t sdoubles = align_up(smem, 256) / sizeof(double);
    int* queue = (int*)((char*)workspace + need_ws - 256);
    CU(cudaMemsetAsync(queue, 0, sizeof(int), stream));
    if (h->profile) {
        CU(cudaMemsetAsync(h->d_prof, 0, PH_COUNT * sizeof(long long), stream));
        CU(cudaEventRecord(h->ev[0], stream));
    }
    if (use_global)
        k_solve<true><<<grid, FTMPC_QP_THREADS, gs_fast_doubles(h->cfg.horizon) * sizeof(double), stream>>>(h->cfg, L, io, queue, gscratch, sdoubles,
                                                             h->profile ? h->d_prof : nullptr);
    else
        k_solve<false><<<grid, FTMPC_QP_THREADS, smem, stream>>>(h->cfg, L, io, queue, nullptr, 0,
                                                                 h->profile ? h->d_prof : nullptr);
    if (h->profile) CU(cudaEventRecord(h->ev[1], stream));
    k_alloc<<<(io.batch + 63) / 64, 64, 0, stream>>>(h->cfg, L, io);
    if (h->profile) CU(cudaEventRecord(h->ev[2], stream));
    h->last_launches = 2;
    CU(cudaGetLastError());
    return FTMPC_OK;
}

int ftmpc_step(ftmpc_handle h, int batch, const double* state, const double* xref, const double* uref,
               const uint16_t* fault_mask, const double* fault_force, const int32_t* hull_idx, int warm,
               double* z_warm, double* thrust, double* u0, uint32_t* active_set, int32_t* status, int32_t* iters,
               double* cost, void* workspace, size_t workspace_bytes, void* stream_) {
    if (!h || batch < 1 || !state || !xref || !fault_mask || !fault_force || !hull_idx || !z_warm || !thrust || !u0 ||
        !active_set || !status || !iters || !workspace)
        return FTMPC_ERR_ARG;
    GUARD(h);
    size_t need_ws = 0;
    ftmpc_workspace_bytes(h, batch, &need_ws);
    if (workspace_bytes < need_ws) return FTMPC_ERR_WORKSPACE;
    StepIO io{batch, state, xref, uref, fault_mask, fault_force, hull_idx, h->d_hull, warm, z_warm, thrust, u0,
              active_set, status, iters, cost, nullptr, nullptr, h->d_cfg, (double*)workspace, nullptr, nullptr, nullptr, 0, 0};
    stepio_default_strides(io, h->L.N);
    return launch_step(h, io, workspace, need_ws, (cudaStream_t)stream_);
}

int ftmpc_closed_loop(ftmpc_handle h, int batch, int steps, int start_step, int table_rows, double* state,
                      const double* trajectory, const double* nominal, const uint16_t* fault_mask,
                      const double* fault_force, const int32_t* hull_idx, const double* noise, int warm_first,
                      double* z_warm, double* thrust, double* u0, uint32_t* active_set, int32_t* status, int32_t* iters,
                      double* cost, double* cost_sum, int32_t* worst_status, void* workspace, size_t workspace_bytes,
                      void* stream_) {
    if (!h || batch < 1 || steps < 1 || start_step < 0 || !state || !trajectory || !fault_mask || !fault_force || !hull_idx ||
        !z_warm || !thrust || !u0 || !active_set || !status || !iters || !cost || !cost_sum || !worst_status || !workspace)
        return FTMPC_ERR_ARG;
    if (start_step + steps + h->L.N > table_rows) return FTMPC_ERR_ARG;      // the last window must lie inside the table
    GUARD(h);
    size_t need_ws = 0;
    ftmpc_workspace_bytes(h, batch, &need_ws);
    if (workspace_bytes < need_ws) return FTMPC_ERR_WORKSPACE;
    cudaStream_t stream = (cudaStream_t)stream_;
    int launches = 0;
    for (int k = 0; k < steps; ++k) {
        const int row = start_step + k;
        StepIO io{batch, state, trajectory + (size_t)row * FTMPC_NE, nominal ? nominal + (size_t)row * FTMPC_NU : nullptr,
                  fault_mask, fault_force, hull_idx, h->d_hull, (k > 0 || warm_first) ? 1 : 0, z_warm, thrust, u0, active_set,
                  status, iters, cost, nullptr, nullptr, h->d_cfg, (double*)workspace, nullptr, nullptr, nullptr, 0, 0};      // strides 0: shared window
        const int rc = launch_step(h, io, workspace, need_ws, stream);
        if (rc != FTMPC_OK) return rc;
        k_plant_loop<<<(batch + 127) / 128, 128, 0, stream>>>(h->d_cfg, batch, state, thrust, fault_mask, fault_force,
                                                              noise ? noise + (size_t)k * batch * FTMPC_NX : nullptr, status,
                                                              cost, cost_sum, worst_status, k == 0);
        launches += 3;
    }
    h->last_launches = launches;
    CU(cudaGetLastError());
    return FTMPC_OK;
}

int ftmpc_hull_facets(ftmpc_handle h, int n_sets, const uint16_t* fault_mask, const double* fault_force, double* table,
                      int32_t* n_rows, int32_t* status, void* stream) {
    if (!h || n_sets < 1 || !fault_mask || !fault_force || !table || !n_rows || !status) return FTMPC_ERR_ARG;
    GUARD(h);
    k_hull_facets<<<n_sets, 256, 0, (cudaStream_t)stream>>>(h->d_cfg, n_sets, fault_mask, fault_force, table, n_rows, status);
    CU(cudaGetLastError());
    return FTMPC_OK;
}

int ftmpc_rk4_jac(ftmpc_handle h, int batch, double* x, const double* wrench, double* jac, const double* lam,
                  double* hess, void* stream) {
    if (!h || batch < 1 || !x || !wrench || !jac) return FTMPC_ERR_ARG;
    GUARD(h);
    k_rk4_jac<<<(batch + 3) / 4, 128, 0, (cudaStream_t)stream>>>(h->d_cfg, batch, x, wrench, jac, lam, hess);
    CU(cudaGetLastError());
    return FTMPC_OK;
}

int ftmpc_robot_to_center(ftmpc_handle h, int batch, const double* state, double* center, void* stream) {
    if (!h || batch < 1 || !state || !center) return FTMPC_ERR_ARG;
    GUARD(h);
    k_r2c<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->d_cfg, batch, state, center);
    CU(cudaGetLastError());
    return FTMPC_OK;
}

int ftmpc_terminal(ftmpc_handle h, int batch, const double* e, double* V, double* grad, double* hess, void* stream) {
    if (!h || batch < 1 || !e || !V || !grad || !hess) return FTMPC_ERR_ARG;
    GUARD(h);
    k_terminal<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->d_cfg, batch, e, V, grad, hess);
    CU(cudaGetLastError());
    return FTMPC_OK;
}

int ftmpc_condense(ftmpc_handle h, int batch, const double* jac, const double* hess, const double* x, const double* u,
                   const double* xref, const double* gradV, const double* hessV, double theta, double* H, double* g,
                   void* stream) {
    if (!h || batch < 1 || !jac || !x || !u || !xref || !gradV || !hessV || !H || !g) return FTMPC_ERR_ARG;
    GUARD(h);
    const size_t smem = qp_smem_bytes(h->cfg.horizon);
    if (smem > h->smem_optin) return FTMPC_ERR_UNSUPPORTED;
    const int grid = batch < h->num_sms ? batch : h->num_sms;
    k_condense<<<grid, FTMPC_QP_THREADS, smem, (cudaStream_t)stream>>>(h->d_cfg, h->L, batch, jac, hess, x, u, xref,
                                                                         gradV, hessV, theta, H, g, nullptr, 0);
    CU(cudaGetLastError());
    return FTMPC_OK;
}

int ftmpc_qp_solve(ftmpc_handle h, int batch, int n, int m, const double* H, const double* g, const int32_t* row_ptr,
                   const int32_t* col_idx, const double* val, const double* b, double* x, double* lam,
                   int32_t* status, void* stream) {
    if (!h || batch < 1 || n < 1 || m < 0 || !H || !g || !row_ptr || !col_idx || !val || !b || !x || !lam || !status)
        return FTMPC_ERR_ARG;
    GUARD(h);
    const int ld = n | 1;
    const size_t doubles = (size_t)n * ld + (size_t)n * (n + 1) / 2 + 1 + 12 * (size_t)n + m + 16;
    const size_t bytes = doubles * 8 + ((size_t)2 * n + m + 8) * 4;
    if (bytes > h->smem_optin) return FTMPC_ERR_UNSUPPORTED;
    const int grid = batch < h->num_sms ? batch : h->num_sms;
    k_qp_generic<<<grid, FTMPC_QP_THREADS, bytes, (cudaStream_t)stream>>>(batch, n, m, H, g, row_ptr, col_idx, val, b, x,
                                                                            lam, status, 20 * (n + m), 1e-11);
    CU(cudaGetLastError());
    return FTMPC_OK;
}

int ftmpc_allocate(ftmpc_handle h, int batch, const double* u_des, const double* ub, double* thrust, int32_t* status,
                   void* stream) {
    if (!h || batch < 1 || !u_des || !ub || !thrust || !status) return FTMPC_ERR_ARG;
    GUARD(h);
    k_allocate<<<(batch + 63) / 64, 64, 0, (cudaStream_t)stream>>>(h->d_cfg, batch, u_des, ub, thrust, status);
    CU(cudaGetLastError());
    return FTMPC_OK;
}

int ftmpc_clip(ftmpc_handle h, int batch, const int32_t* hull_idx, const double* u, double* u_clipped, int32_t* status,
               void* stream) {
    if (!h || batch < 1 || !hull_idx || !u || !u_clipped || !status) return FTMPC_ERR_ARG;
    GUARD(h);
    k_clip<<<(batch + 63) / 64, 64, 0, (cudaStream_t)stream>>>(h->d_cfg, batch, h->d_hull, hull_idx, u, u_clipped, status);
    CU(cudaGetLastError());
    return FTMPC_OK;
}

int ftmpc_plant_step(ftmpc_handle h, int batch, const double* state, const double* thrust, const uint16_t* fault_mask,
                     const double* fault_force, const double* noise, int normalize, double* next, void* stream) {
    if (!h || batch < 1 || !state || !thrust || !fault_mask || !fault_force || !next) return FTMPC_ERR_ARG;
    GUARD(h);
    k_plant<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->d_cfg, batch, state, thrust, fault_mask,
                                                                     fault_force, noise, normalize, next);
    CU(cudaGetLastError());
    return FTMPC_OK;
}

int ftmpc_fp64_peak(ftmpc_handle h, double* tflops, void* stream_) {
    if (!h || !tflops) return FTMPC_ERR_ARG;
    GUARD(h);
    cudaStream_t stream = (cudaStream_t)stream_;
    double* d = nullptr;
    CU(cudaMalloc(&d, sizeof(double)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int iters = 1 << 16, grid = h->num_sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(e0, stream));
        k_fp64_peak<<<grid, 256, 0, stream>>>(d, iters, 0.999999, 1e-9);
        CU(cudaEventRecord(e1, stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 8.0 * iters * 256.0 * grid / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return FTMPC_OK;
}

}  // extern "C"
