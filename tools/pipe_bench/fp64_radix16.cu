// fp64_radix16.cu -- the compute ceiling of the NTT's FP64 radix-16 block AS COMPILED: sixteen values in registers, four
// butterfly stages with twiddles from shared memory, one fold per value, repeated; no global or shared data traffic.
// Butterflies per clock and SM at 2 / 3 / 4 CTAs of 256 threads per SM (register caps 128 / 85 / 64).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ double modmul_fp(double a, double w, double q, double qinv) {
    const double h = __dmul_rn(a, w);
    const double l = __fma_rn(a, w, -h);
    const double c = rint(__dmul_rn(h, qinv));
    return __dadd_rn(__fma_rn(-c, q, h), l);
}
__device__ __forceinline__ double fold_fp(double x, double q, double qinv) { return __fma_rn(-rint(__dmul_rn(x, qinv)), q, x); }
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k(double* out, double q, int iters) {
    __shared__ double tw[256];
    tw[threadIdx.x] = 734623412345677.0 - threadIdx.x * 1024.0;
    __syncthreads();
    const double qinv = 1.0 / q;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = 1000.0 * threadIdx.x + i;
    const int xl = threadIdx.x >> 4;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int span = 8 >> s;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int g = i / span, k0 = g * 2 * span + (i % span), k1 = k0 + span;
                const double w = tw[(16 << s) - 1 + (xl << s) + g];
                const double u = x[k0];
                const double v = modmul_fp(x[k1], w, q, qinv);
                x[k0] = __dadd_rn(u, v);
                x[k1] = __dsub_rn(u, v);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fold_fp(x[i], q, qinv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the same block with the pass-A shared-memory exchange around it (store 16, barrier, load 16 transposed, barrier), still no
// global traffic: what the exchange and its two CTA barriers cost on top of the arithmetic
template <int MINB>
__global__ void __launch_bounds__(256, MINB) kx(double* out, double q, int iters) {
    __shared__ double tw[256];
    __shared__ double sd[4096];
    tw[threadIdx.x] = 734623412345677.0 - threadIdx.x * 1024.0;
    __syncthreads();
    const double qinv = 1.0 / q;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = 1000.0 * threadIdx.x + i;
    const int c = threadIdx.x & 15, rr = threadIdx.x >> 4;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int span = 8 >> s;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int g = i / span, k0 = g * 2 * span + (i % span), k1 = k0 + span;
                const double w = tw[(16 << s) - 1 + (rr << s) + g];
                const double u = x[k0];
                const double v = modmul_fp(x[k1], w, q, qinv);
                x[k0] = __dadd_rn(u, v);
                x[k1] = __dsub_rn(u, v);
            }
        }
#pragma unroll
        for (int k = 0; k < 16; k++) sd[(rr + 16 * k) * 16 + c] = fold_fp(x[k], q, qinv);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = sd[(16 * rr + k) * 16 + c];
        __syncthreads();
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MINB>
void runx() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * MINB, iters = 2048;
    double* out; cudaMalloc(&out, (size_t)blocks * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kx<MINB><<<blocks, 256>>>(out, 1125899906826241.0, 8);
    cudaEventRecord(e0);
    kx<MINB><<<blocks, 256>>>(out, 1125899906826241.0, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double bf = (double)blocks * 256 * iters * 32 / (ms * 1e-3);
    printf("radix-16 block + shared-memory exchange + 2 barriers, %d CTAs/SM: %8.3f ms  %6.2f butterflies/clk/SM  (NTT-equivalent %.3f us per limb)\n",
           MINB, ms, bf / ((double)clk * 1e3) / sms, 524288.0 / bf * 1e6);
    cudaFree(out);
}
template <int MINB>
void run() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * MINB, iters = 2048;
    double* out; cudaMalloc(&out, (size_t)blocks * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MINB><<<blocks, 256>>>(out, 1125899906826241.0, 8);
    cudaEventRecord(e0);
    k<MINB><<<blocks, 256>>>(out, 1125899906826241.0, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double bf = (double)blocks * 256 * iters * 32 / (ms * 1e-3);
    printf("radix-16 block, %d CTAs/SM: %8.3f ms  %6.2f butterflies/clk/SM  (NTT-equivalent %.3f us per 2^16-point limb)\n", MINB, ms,
           bf / ((double)clk * 1e3) / sms, 524288.0 / bf * 1e6);
    cudaFree(out);
}
int main() { run<2>(); run<3>(); run<4>(); run<5>(); runx<2>(); runx<3>(); runx<4>(); runx<5>(); return 0; }
