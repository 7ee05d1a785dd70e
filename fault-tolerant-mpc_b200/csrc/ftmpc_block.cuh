// ftmpc_block.cuh -- the cooperative-group abstraction every per-instance routine is written against.
//
// All solver routines (condensing, Cholesky, the dual active-set QP) are written once, in
// "strided loop + barrier" form, against a Block type:
//
//     for (int i = blk.tid(); i < n; i += blk.nthreads()) { ... }   blk.sync();
//
//   * CudaBlock  : one CUDA thread block per MPC instance (the product path).
//   * SerialBlock: one host thread per instance.  Used ONLY by the CPU port that tests and
//                  bench.py's cpu_baseline build from these same headers (oracle/cpu_port); the
//                  product library never instantiates it.
//
// Reductions return the block-wide result to every thread, in a fixed order, so that all threads
// take identical branches (the solver control flow is uniform across the block).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FT_HD __host__ __device__ __forceinline__
#define FT_D __device__ __forceinline__
#else
#define FT_HD inline
#endif

namespace ftmpc {

struct SerialBlock {
    FT_HD int tid() const { return 0; }
    FT_HD int nthreads() const { return 1; }
    FT_HD void sync() const {}
    FT_HD double sum(double v) { return v; }
    FT_HD double max(double v) { return v; }
    // (value, index) pair with the smallest value; ties -> smallest index
    FT_HD void argmin(double& v, int& idx) { (void)v; (void)idx; }
    FT_HD int any(int pred) { return pred; }
    FT_HD void mark(int) {}
    FT_HD void count(int, int = 1) {}
};

// phase ids of the in-kernel cycle profile (ftmpc_profile_read)
enum { PH_LS = 0, PH_LIN, PH_COND, PH_CHOL, PH_INV, PH_QPSETUP, PH_GI, PH_POST, PH_OUT,
       // finer attribution inside the phases above (the coarse ids then only receive the remainder)
       PH_LS_ROLL, PH_LS_EVAL, PH_LS_TERM, PH_CHOL_PANEL, PH_CHOL_SYRK, PH_GI_SELECT, PH_GI_D, PH_GI_Z, PH_GI_STEP, PH_GI_UPD,
       PH_GI_DROP, PH_LIN_JAC, PH_LIN_MU, PH_COND_PRE, PH_COND_COL, PH_COND_BLK, PH_WS_D0, PH_WS_QR, PH_WS_SOLVE, PH_WS_E,
       // Riccati factorisation (ftmpc_riccati.cuh): set-up, the two steps of the AB interval as warp 0 sees them + its wait at
       // the barrier, the CD interval (thread 0's Cholesky chain) + wait, the back-substitution pass, the forward rollouts
       PH_RIC_PRE, PH_RIC_AB1, PH_RIC_AB2, PH_RIC_ABW, PH_RIC_CD, PH_RIC_CDW, PH_RIC_POST, PH_RIC_FWD,
       // event counters (not cycles)
       CT_INST, CT_SQP, CT_CONDENSE, CT_CHOL_FAIL, CT_QP, CT_GI_ITER, CT_GI_DROP, CT_LS_BACKTRACK, CT_GI_WARM_OK, CT_GI_WARM_MISS, CT_GI_REFINE,
       PH_COUNT };

#if defined(__CUDACC__)
// ---- TMA bulk staging (cp.async.bulk + mbarrier): per-instance problem data (input-bound hull, reference windows,
// constraint values) travel from HBM / L2 to shared memory as bulk copies issued by ONE thread; the other threads
// only wait on the mbarrier.  Addresses must be 16-byte aligned and sizes multiples of 16 bytes: a row of doubles
// that starts or ends on an odd double gets that element moved by a plain load (tma_stage_row).
struct TmaBar {
    unsigned addr;       // shared-memory address of the mbarrier
    unsigned parity;     // phase parity the next wait expects (per thread, all threads in step)
};
__device__ __forceinline__ void tma_bar_init(unsigned long long* bar_smem, TmaBar& b) {
    b.addr = (unsigned)__cvta_generic_to_shared(bar_smem);
    b.parity = 0u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b.addr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
}
// one thread: announce `bytes` of bulk traffic on the barrier (also its single arrival)
__device__ __forceinline__ void tma_expect(const TmaBar& b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b.addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(const TmaBar& b, void* dst_smem, const void* src_gmem, unsigned bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src_gmem),
                 "r"(bytes), "r"(b.addr)
                 : "memory");
}
// all threads: wait for the current phase of the barrier, then flip the expected parity
__device__ __forceinline__ void tma_wait(TmaBar& b) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(b.addr), "r"(b.parity)
                     : "memory");
    }
    b.parity ^= 1u;
}
// Where a row of n doubles starting at `src` lands in a 16-byte aligned staging buffer `raw` (n + 2 doubles): the data
// pointer is raw + (src on an odd double ? 1 : 0), so that the 16-byte aligned body of the row is 16-byte aligned there too.
__device__ __forceinline__ double* tma_row_ptr(double* raw, const double* src) {
    return raw + ((reinterpret_cast<unsigned long long>(src) >> 3) & 1ull);
}
// one thread: odd head / tail element by plain loads, the aligned body as one bulk copy; returns the bytes announced
__device__ __forceinline__ unsigned tma_stage_row(const TmaBar& b, double* raw, const double* src, int n) {
    const int mis = (int)((reinterpret_cast<unsigned long long>(src) >> 3) & 1ull);
    double* dst = raw + mis;
    if (mis) dst[0] = src[0];
    const int body = (n - mis) & ~1;
    if ((n - mis) & 1) dst[n - 1] = src[n - 1];
    if (body > 0) tma_bulk_g2s(b, dst + mis, src + mis, (unsigned)body * 8u);
    return (unsigned)body * 8u;
}
__device__ __forceinline__ unsigned tma_row_bytes(const double* src, int n) {
    const int mis = (int)((reinterpret_cast<unsigned long long>(src) >> 3) & 1ull);
    return (unsigned)((n - mis) & ~1) * 8u;
}

// one warp as a block (serial sweeps that need nothing wider: ric_apply)
struct WarpBlock {
    __device__ __forceinline__ int tid() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int nthreads() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};

// scratch: >= 2 * 32 * 2 doubles of shared memory (double-buffered so one barrier per reduction suffices)
struct CudaBlock {
    double* scratch;
    int phase;
    long long* prof;         // optional per-phase cycle accumulators (global memory), thread 0 only
    long long t_last;
    TmaBar tma;              // bulk-copy barrier of this CTA (addr = 0: the kernel stages with plain loads)
    __device__ __forceinline__ explicit CudaBlock(double* s, long long* p = nullptr) : scratch(s), phase(0), prof(p), t_last(0) {
        if (prof) t_last = clock64();
        tma.addr = 0u;
        tma.parity = 0u;
    }
    // attribute the cycles since the previous mark to phase `id` (call right after a block-wide barrier)
    __device__ __forceinline__ void mark(int id) {
        if (prof && threadIdx.x == 0) {
            const long long t = clock64();
            atomicAdd(reinterpret_cast<unsigned long long*>(prof + id), (unsigned long long)(t - t_last));
            t_last = t;
        }
    }
    __device__ __forceinline__ void count(int id, int n = 1) {
        if (prof && threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(prof + id), (unsigned long long)n);
    }
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }

    __device__ __forceinline__ double* buf() {
        double* b = scratch + phase * 64;
        phase ^= 1;
        return b;
    }
    __device__ __forceinline__ double sum(double v) {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        double* b = buf();
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if ((threadIdx.x & 31) == 0) b[w] = v;
        __syncthreads();
        if (nw == 8) {                               // the product configuration: independent loads, fixed order
            const double2 p0 = *reinterpret_cast<const double2*>(b), p1 = *reinterpret_cast<const double2*>(b + 2);
            const double2 p2 = *reinterpret_cast<const double2*>(b + 4), p3 = *reinterpret_cast<const double2*>(b + 6);
            return ((p0.x + p0.y) + (p1.x + p1.y)) + ((p2.x + p2.y) + (p3.x + p3.y));
        }
        double r = 0.0;
        for (int i = 0; i < nw; ++i) r += b[i];
        return r;
    }
    __device__ __forceinline__ double max(double v) {
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        double* b = buf();
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if ((threadIdx.x & 31) == 0) b[w] = v;
        __syncthreads();
        if (nw == 8) {
            const double2 p0 = *reinterpret_cast<const double2*>(b), p1 = *reinterpret_cast<const double2*>(b + 2);
            const double2 p2 = *reinterpret_cast<const double2*>(b + 4), p3 = *reinterpret_cast<const double2*>(b + 6);
            return fmax(fmax(fmax(p0.x, p0.y), fmax(p1.x, p1.y)), fmax(fmax(p2.x, p2.y), fmax(p3.x, p3.y)));
        }
        double r = b[0];
        for (int i = 1; i < nw; ++i) r = fmax(r, b[i]);
        return r;
    }
    __device__ __forceinline__ void argmin(double& v, int& idx) {
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        double* b = buf();
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if ((threadIdx.x & 31) == 0) { b[w] = v; b[32 + w] = (double)idx; }
        __syncthreads();
        if (nw == 8) {                               // all sixteen loads in flight at once, then a compare tree
            double vv[8], ii[8];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const double2 pv = *reinterpret_cast<const double2*>(b + i), pi = *reinterpret_cast<const double2*>(b + 32 + i);
                vv[i] = pv.x; vv[i + 1] = pv.y; ii[i] = pi.x; ii[i + 1] = pi.y;
            }
#pragma unroll
            for (int st = 1; st < 8; st <<= 1)
#pragma unroll
                for (int i = 0; i < 8; i += 2 * st)
                    if (vv[i + st] < vv[i] || (vv[i + st] == vv[i] && ii[i + st] < ii[i])) { vv[i] = vv[i + st]; ii[i] = ii[i + st]; }
            v = vv[0];
            idx = (int)ii[0];
            return;
        }
        double rv = b[0];
        int ri = (int)b[32];
        for (int i = 1; i < nw; ++i) {
            const double ov = b[i];
            const int oi = (int)b[32 + i];
            if (ov < rv || (ov == rv && oi < ri)) { rv = ov; ri = oi; }
        }
        v = rv;
        idx = ri;
    }
    __device__ __forceinline__ int any(int pred) { return __syncthreads_or(pred); }
};
#endif

}  // namespace ftmpc
