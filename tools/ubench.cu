// ubench.cu -- latency / issue-rate probes that size the per-instance serial chains of k_solve on sm_100a.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench tools/ubench.cu ; run on one GPU.
// Every probe runs ONE CTA (optionally one per SM) and reports cycles per operation from clock64().
#include <cstdio>
#include <cuda_runtime.h>

#define REP 512

template <int CH>
__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
    double x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = a + i + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP; ++r) {
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fma(x[i], b, a);
    }
    const long long t1 = clock64();
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_barrier(long long* cyc) {
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP; ++r) __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// barrier + one dependent shared-memory round trip per interval (the shape of a short solver phase)
__global__ void k_barrier_lds(long long* cyc, double* out) {
    __shared__ double sm[1024];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double v = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP; ++r) {
        v += sm[(threadIdx.x + r) & 255];
        sm[threadIdx.x] = v;
        __syncthreads();
    }
    const long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lds_chase(long long* cyc, int* out) {
    __shared__ int sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 33 + 7) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP; ++r) p = sm[p];
    const long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lds64_chase(long long* cyc, double* out) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 33 + 7) & 1023);
    __syncthreads();
    double p = threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP; ++r) p = sm[(int)p];          // includes an F2I
    const long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
__global__ void k_special(long long* cyc, double* out, double a) {
    double x = a + threadIdx.x * 1e-3;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP; ++r) {
        if (OP == 0) x = rsqrt(x) + a;
        if (OP == 1) x = 1.0 / x + a;
        if (OP == 2) x = sqrt(x) + a;
        if (OP == 3) x = __shfl_xor_sync(0xffffffffu, x, 1) + a;
        if (OP == 4) x = x + a;
        if (OP == 5) x = x * a;
        if (OP == 6) x = fmax(x, a) + 1e-9;
    }
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared-memory streaming: every thread sums `len` doubles of its own row (row stride ld), 4 accumulators
__global__ void k_lds_stream(long long* cyc, double* out, int ld, int len) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 130 * ld; i += blockDim.x) sm[i] = i * 1e-6;
    __syncthreads();
    const int row = threadIdx.x >> 1, part = threadIdx.x & 1;
    const double* e = sm + (size_t)(row % 130) * ld + part;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < 16; ++r) {
        for (int k = 0; k + 6 < len; k += 8) {
            s0 += e[k]; s1 += e[k + 2]; s2 += e[k + 4]; s3 += e[k + 6];
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = s0 + s1 + s2 + s3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static long long run_report(const char* name, long long* d_cyc, double per) {
    long long h = 0;
    cudaDeviceSynchronize();
    cudaMemcpy(&h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s %10lld cycles  %8.2f per op\n", name, h, (double)h / per);
    return h;
}

int main() {
    double* d_out;
    long long* d_cyc;
    int* d_iout;
    cudaMalloc(&d_out, 1 << 20);
    cudaMalloc(&d_iout, 1 << 20);
    cudaMalloc(&d_cyc, 4096);
    cudaError_t e;
    for (int threads : {32, 128, 256, 512, 1024}) {
        printf("--- %d threads per CTA, one CTA\n", threads);
        char nm[96];
        k_dfma<1><<<1, threads>>>(d_out, d_cyc, 1.0, 0.999);  snprintf(nm, 96, "DFMA 1 chain/thread");  run_report(nm, d_cyc, REP);
        k_dfma<2><<<1, threads>>>(d_out, d_cyc, 1.0, 0.999);  snprintf(nm, 96, "DFMA 2 chains/thread (per DFMA)");  run_report(nm, d_cyc, REP * 2);
        k_dfma<4><<<1, threads>>>(d_out, d_cyc, 1.0, 0.999);  snprintf(nm, 96, "DFMA 4 chains/thread (per DFMA)");  run_report(nm, d_cyc, REP * 4);
        k_dfma<8><<<1, threads>>>(d_out, d_cyc, 1.0, 0.999);  snprintf(nm, 96, "DFMA 8 chains/thread (per DFMA)");  run_report(nm, d_cyc, REP * 8);
        k_dfma<36><<<1, threads>>>(d_out, d_cyc, 1.0, 0.999); snprintf(nm, 96, "DFMA 36 chains/thread (per DFMA)"); run_report(nm, d_cyc, REP * 36);
        k_barrier<<<1, threads>>>(d_cyc);                     run_report("__syncthreads", d_cyc, REP);
        if (threads <= 256) { k_barrier_lds<<<1, threads>>>(d_cyc, d_out); run_report("LDS + STS + __syncthreads", d_cyc, REP); }
    }
    printf("--- one warp\n");
    k_lds_chase<<<1, 32>>>(d_cyc, d_iout);    run_report("LDS.32 pointer chase", d_cyc, REP);
    k_lds64_chase<<<1, 32>>>(d_cyc, d_out);   run_report("LDS.64 + F2I chase", d_cyc, REP);
    k_special<0><<<1, 32>>>(d_cyc, d_out, 1.5); run_report("rsqrt(double) + DADD", d_cyc, REP);
    k_special<1><<<1, 32>>>(d_cyc, d_out, 1.5); run_report("1.0 / x + DADD", d_cyc, REP);
    k_special<2><<<1, 32>>>(d_cyc, d_out, 1.5); run_report("sqrt(double) + DADD", d_cyc, REP);
    k_special<3><<<1, 32>>>(d_cyc, d_out, 1.5); run_report("shfl.f64 + DADD", d_cyc, REP);
    k_special<4><<<1, 32>>>(d_cyc, d_out, 1.5); run_report("DADD", d_cyc, REP);
    k_special<5><<<1, 32>>>(d_cyc, d_out, 1.000001); run_report("DMUL", d_cyc, REP);
    k_special<6><<<1, 32>>>(d_cyc, d_out, 1.5); run_report("fmax + DADD", d_cyc, REP);
    printf("--- 256 threads, lane pair per row sweep over 130 x ld doubles of shared memory (16 sweeps)\n");
    cudaFuncSetAttribute(k_lds_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int ld : {121, 122, 128}) {
        k_lds_stream<<<1, 256, 130 * ld * 8>>>(d_cyc, d_out, ld, 120);
        char nm[96];
        snprintf(nm, 96, "row sweep ld=%d (per sweep of 126 KB)", ld);
        run_report(nm, d_cyc, 16);
    }
    e = cudaGetLastError();
    printf("last error: %s\n", cudaGetErrorString(e));
    return 0;
}
