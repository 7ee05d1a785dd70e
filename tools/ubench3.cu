// ubench3.cu -- the stage step of the G-form operator (ric_g_step_sh) in isolation: one warp, shared memory only.
#include <cstdio>
#include <cuda_runtime.h>
#define REP 400
__device__ __forceinline__ double lds(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

template <bool TR, int VAR>
__global__ void k_step(double* out, long long* cyc) {
    __shared__ double G[8 * 397], vec[2][13], w[16 * 6], o6[16 * 6];
    for (int i = threadIdx.x; i < 8 * 397; i += blockDim.x) G[i] = 1e-3 * (i % 17) - 5e-3;
    for (int i = threadIdx.x; i < 26; i += blockDim.x) (&vec[0][0])[i] = 0.1;
    for (int i = threadIdx.x; i < 96; i += blockDim.x) { w[i] = 0.01; o6[i] = 0.0; }
    __syncthreads();
    if (threadIdx.x >= 32) { __syncthreads(); return; }
    const int l = threadIdx.x;
    const unsigned Gs = (unsigned)__cvta_generic_to_shared(G), vs = (unsigned)__cvta_generic_to_shared(&vec[0][0]);
    const unsigned ws = (unsigned)__cvta_generic_to_shared(w), os = (unsigned)__cvta_generic_to_shared(o6);
    const long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < REP; ++rep) {
        const int t = rep & 7;
        const unsigned g0 = Gs + 8u * (t * 397), in13 = vs + 8u * ((rep & 1) * 13), out13 = vs + 8u * (((rep + 1) & 1) * 13);
        const unsigned in6 = ws + 8u * (t * 6), out6 = os + 8u * (t * 6);
        if (VAR == 0) {
            if (l < 19) {
                double g[19], x[19];
#pragma unroll
                for (int k = 0; k < 19; ++k) g[k] = lds(g0 + 8u * (unsigned)(TR ? k * 19 + l : l * 19 + k));
#pragma unroll
                for (int k = 0; k < 13; ++k) x[k] = lds(in13 + 8u * k);
#pragma unroll
                for (int k = 0; k < 6; ++k) x[13 + k] = lds(in6 + 8u * k);
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                for (int k = 0; k < 16; k += 4) { a0 += g[k] * x[k]; a1 += g[k + 1] * x[k + 1]; a2 += g[k + 2] * x[k + 2]; a3 += g[k + 3] * x[k + 3]; }
                a0 += g[16] * x[16]; a1 += g[17] * x[17]; a2 += g[18] * x[18];
                sts((l < 13) ? out13 + 8u * l : out6 + 8u * (l - 13), (a0 + a1) + (a2 + a3));
            }
            __syncwarp();
        } else {
            // the state vector travels in REGISTERS: lane l < 13 holds entry l; broadcast by shuffles, no shared-memory hand-over
            static_assert(VAR == 0 || VAR == 1, "");
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = lds(vs + 8u * (threadIdx.x % 13));
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    __syncthreads();
}

// variant: state vector in registers, shuffle broadcast
template <bool TR>
__global__ void k_step_shfl(double* out, long long* cyc) {
    __shared__ double G[8 * 397], w[16 * 6], o6[16 * 6];
    for (int i = threadIdx.x; i < 8 * 397; i += blockDim.x) G[i] = 1e-3 * (i % 17) - 5e-3;
    for (int i = threadIdx.x; i < 96; i += blockDim.x) { w[i] = 0.01; o6[i] = 0.0; }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int l = threadIdx.x;
    const unsigned Gs = (unsigned)__cvta_generic_to_shared(G), ws = (unsigned)__cvta_generic_to_shared(w), os = (unsigned)__cvta_generic_to_shared(o6);
    double xl = 0.1;                 // lane l < 13: state entry l
    const long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < REP; ++rep) {
        const int t = rep & 7;
        const unsigned g0 = Gs + 8u * (t * 397), in6 = ws + 8u * (t * 6), out6 = os + 8u * (t * 6);
        double g[19];
        const int ll = l < 19 ? l : 18;
#pragma unroll
        for (int k = 0; k < 19; ++k) g[k] = lds(g0 + 8u * (unsigned)(TR ? k * 19 + ll : ll * 19 + k));
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) { const double xk = lds(in6 + 8u * k); if (k & 1) a1 += g[13 + k] * xk; else a0 += g[13 + k] * xk; }
#pragma unroll
        for (int k = 0; k < 13; ++k) {
            const double xk = __shfl_sync(0xffffffffu, xl, k);
            if ((k & 3) == 0) a0 += g[k] * xk; else if ((k & 3) == 1) a1 += g[k] * xk; else if ((k & 3) == 2) a2 += g[k] * xk; else a3 += g[k] * xk;
        }
        const double v = (a0 + a1) + (a2 + a3);
        if (l >= 13 && l < 19) sts(out6 + 8u * (l - 13), v);
        xl = v;
    }
    const long long t1 = clock64();
    out[threadIdx.x] = xl;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
static void rep(const char* n, long long* d) {
    long long h; cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-64s %9lld cycles  %8.1f per stage step\n", n, h, (double)h / REP);
}
int main() {
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 8192); cudaMalloc(&d_cyc, 64);
    for (int pass = 0; pass < 2; ++pass) {
        k_step<true, 0><<<1, 256>>>(d_out, d_cyc); rep("G' step (backward), smem hand-over, 256 thr CTA", d_cyc);
        k_step<false, 0><<<1, 256>>>(d_out, d_cyc); rep("G step (forward), smem hand-over, 256 thr CTA", d_cyc);
        k_step_shfl<true><<<1, 32>>>(d_out, d_cyc); rep("G' step, state in registers (shuffle broadcast)", d_cyc);
        k_step_shfl<false><<<1, 32>>>(d_out, d_cyc); rep("G step, state in registers (shuffle broadcast)", d_cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
