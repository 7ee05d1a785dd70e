// FP64 throughput probe for B200 (sm_100a): vector DFMA vs tensor-core DMMA shapes exposed by mma.sync.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/dmma_probe.cu -o build/dmma_probe ; run on the GPU box.
// Decides whether the rank-13 updates of the condensing step belong on the tensor cores (profiles/README.md).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a[8], x = 1.0 + threadIdx.x * 1e-9, y = 0.999999;
    for (int i = 0; i < 8; ++i) a[i] = i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma884(double* out, int iters) {
    double c[8][2], a = 1.0 + threadIdx.x * 1e-9, b = 0.999999;
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma16816(double* out, int iters) {
    double c[4][4], a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 4; ++i) { b[i] = 0.999999 + i; for (int j = 0; j < 4; ++j) c[i][j] = i + j; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> static double time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 4096;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    for (int warps = 1; warps <= 32; warps *= 2) {
        const int bs = 32 * (warps > 8 ? 8 : warps), blocks = sms * (warps > 8 ? warps / 8 : 1);
        double t0 = time_ms([&] { k_dfma<<<blocks, bs>>>(out, iters); });
        double t1 = time_ms([&] { k_dmma884<<<blocks, bs>>>(out, iters); });
        double t2 = time_ms([&] { k_dmma16816<<<blocks, bs>>>(out, iters); });
        const double thr = (double)blocks * bs;
        printf("warps/SM %2d : DFMA %7.2f TFLOP/s | DMMA m8n8k4 %7.2f TFLOP/s | DMMA m16n8k16 %7.2f TFLOP/s\n", warps,
               thr * iters * 8 * 2 / (t0 * 1e-3) * 1e-12, thr / 32 * iters * 8 * (8 * 8 * 4 * 2) / (t1 * 1e-3) * 1e-12,
               thr / 32 * iters * 4 * (16.0 * 8 * 16 * 2) / (t2 * 1e-3) * 1e-12);
    }
    return 0;
}
