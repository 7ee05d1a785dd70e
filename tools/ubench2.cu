// ubench2.cu -- cycle counts of the building blocks of the Riccati sweep on one warp / one CTA (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
#define REP 200

// (1) 6x6 Cholesky with two forward substitutions riding along, as in ric_backward_cuda (registers only)
template <int VAR>
__global__ void k_chol(const double* in, double* out, long long* cyc) {
    __shared__ double F[64];
    if (threadIdx.x < 64) F[threadIdx.x] = in[threadIdx.x];
    __syncthreads();
    double accum = 0.0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < REP; ++rep) {
        double Lm[6][6], fr[6], fc[6], yr[6], yc[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j < 6; ++j) Lm[i][j] = (j <= i) ? F[i * 6 + j] + accum : 0.0;
            fr[i] = F[36 + i]; fc[i] = F[42 + i];
        }
        double pv = F[50];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double d = Lm[j][j];
#pragma unroll
            for (int m = 0; m < 6; ++m) if (m < j) d -= Lm[j][m] * Lm[j][m];
            double rs;
            if (VAR == 0) rs = rsqrt(d);
            else if (VAR == 1) rs = 1.0 / sqrt(d);
            else {      // MUFU.RSQ64H seed + two Newton steps, no special cases (d is a safely positive pivot)
                double y;
                asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\trsqrt.approx.ftz.f64 %0, %1;\n\t}" : "=d"(y) : "d"(d));
                const double h = 0.5 * d;
                double e = fma(-h * y, y, 0.5);
                y = fma(y, e, y);
                e = fma(-h * y, y, 0.5);
                rs = fma(y, e, y);
            }
            Lm[j][j] = d * rs;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                if (i > j) {
                    double v = Lm[i][j];
#pragma unroll
                    for (int m = 0; m < 6; ++m) if (m < j) v -= Lm[i][m] * Lm[j][m];
                    Lm[i][j] = v * rs;
                }
            }
            double vr = fr[j], vc = fc[j];
#pragma unroll
            for (int m = 0; m < 6; ++m) if (m < j) { vr -= Lm[j][m] * yr[m]; vc -= Lm[j][m] * yc[m]; }
            yr[j] = vr * rs; yc[j] = vc * rs;
            pv -= yr[j] * yc[j];
        }
        accum = pv * 1e-30;      // loop-carried dependency through the whole factorisation
    }
    const long long t1 = clock64();
    out[threadIdx.x] = accum;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// (2) 13-long dot product from shared memory, two accumulators, result stored and re-read (loop-carried through smem)
template <int NACC>
__global__ void k_dot13(double* out, long long* cyc) {
    __shared__ double P[13 * 13], col[32 * 13], res[256];
    for (int i = threadIdx.x; i < 169; i += blockDim.x) P[i] = 1e-3 * i;
    for (int i = threadIdx.x; i < 32 * 13; i += blockDim.x) col[i] = 1e-3 * i;
    __syncthreads();
    const int r = threadIdx.x % 13, c = (threadIdx.x / 13) % 19;
    double carry = 0.0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < REP; ++rep) {
        const double* pa = P + r * 13;
        const double* cl = col + c * 13;
        double v;
        if (NACC == 2) {
            double a0 = carry, a1 = 0.0;
#pragma unroll
            for (int k = 0; k < 12; k += 2) { a0 += pa[k] * cl[k]; a1 += pa[k + 1] * cl[k + 1]; }
            v = a0 + a1 + pa[12] * cl[12];
        } else {
            double a0 = carry, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int k = 0; k < 12; k += 4) { a0 += pa[k] * cl[k]; a1 += pa[k + 1] * cl[k + 1]; a2 += pa[k + 2] * cl[k + 2]; a3 += pa[k + 3] * cl[k + 3]; }
            v = (a0 + a1) + (a2 + a3) + pa[12] * cl[12];
        }
        res[threadIdx.x] = v;
        __syncwarp();
        carry = res[threadIdx.x ^ 1] * 1e-30;
    }
    const long long t1 = clock64();
    out[threadIdx.x] = carry;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// (3) the same with a block barrier per repetition (the shape of one barrier interval)
__global__ void k_dot13_bar(double* out, long long* cyc) {
    __shared__ double P[13 * 13], col[32 * 13], res[256];
    for (int i = threadIdx.x; i < 169; i += blockDim.x) P[i] = 1e-3 * i;
    for (int i = threadIdx.x; i < 32 * 13; i += blockDim.x) col[i] = 1e-3 * i;
    __syncthreads();
    const int r = threadIdx.x % 13, c = (threadIdx.x / 13) % 19;
    double carry = 0.0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < REP; ++rep) {
        const double* pa = P + r * 13;
        const double* cl = col + c * 13;
        double a0 = carry, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < 12; k += 2) { a0 += pa[k] * cl[k]; a1 += pa[k + 1] * cl[k + 1]; }
        res[threadIdx.x] = a0 + a1 + pa[12] * cl[12];
        __syncthreads();
        carry = res[(threadIdx.x + 32) & 255] * 1e-30;
    }
    const long long t1 = clock64();
    out[threadIdx.x] = carry;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

static void rep(const char* n, long long* d) {
    long long h; cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-60s %9lld cycles  %8.1f per repetition\n", n, h, (double)h / REP);
}
int main() {
    double h_in[64]; for (int i = 0; i < 64; ++i) h_in[i] = 0.01 * (i % 7);
    for (int i = 0; i < 6; ++i) h_in[i * 6 + i] = 4.0 + i;
    double *d_in, *d_out; long long* d_cyc;
    cudaMalloc(&d_in, 512); cudaMalloc(&d_out, 8192); cudaMalloc(&d_cyc, 64);
    cudaMemcpy(d_in, h_in, 512, cudaMemcpyHostToDevice);
    for (int th : {32, 96, 256}) {
        printf("--- %d threads\n", th);
        k_chol<0><<<1, th>>>(d_in, d_out, d_cyc); rep("chol6 + 2 fwd subst, rsqrt()", d_cyc);
        k_chol<1><<<1, th>>>(d_in, d_out, d_cyc); rep("chol6 + 2 fwd subst, 1/sqrt()", d_cyc);
        k_chol<2><<<1, th>>>(d_in, d_out, d_cyc); rep("chol6 + 2 fwd subst, rsqrt.approx + 2 Newton", d_cyc);
        k_dot13<2><<<1, th>>>(d_out, d_cyc); rep("13-dot from smem, 2 acc, STS + syncwarp + LDS", d_cyc);
        k_dot13<4><<<1, th>>>(d_out, d_cyc); rep("13-dot from smem, 4 acc, STS + syncwarp + LDS", d_cyc);
    }
    k_dot13_bar<<<1, 256>>>(d_out, d_cyc); rep("256 thr: 13-dot, STS + __syncthreads + LDS", d_cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
