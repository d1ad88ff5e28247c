// fp64_modmul.cu -- throughput and exactness of an FP64 (DFMA) lazy modular multiplication for moduli below 2^50,
// measured against the integer Shoup multiplication (is the FP64 pipe a faster home for the 50-bit NTT limbs?).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;

__device__ __forceinline__ double modmul_fp(double a, double w, double wq, double q) {
    const double h = __dmul_rn(a, w);
    const double l = __fma_rn(a, w, -h);
    const double c = rint(__dmul_rn(a, wq));
    const double r = __fma_rn(-c, q, h);
    return __dadd_rn(r, l);                     // in (-q, 2q), exact integer
}
template <int OP>
__global__ void k(double* out, double a0, double w, double wq, double q, int iters) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a0 + threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) x[i] = __fma_rn(x[i], w, a0);
            if (OP == 1) { double r = modmul_fp(x[i], w, wq, q); x[i] = r < 0 ? r + q : r; }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void check(const u64* a, u64 w, u64 q, int n, int* bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r = modmul_fp((double)a[i], (double)w, (double)w / (double)q, (double)q);
    long long ri = (long long)r;
    ri %= (long long)q; if (ri < 0) ri += q;
    const u64 want = (u64)(((unsigned __int128)a[i] * w) % q);
    if ((u64)ri != want || r <= -(double)q || r >= 2.0 * (double)q) atomicAdd(bad, 1);
}
template <int OP>
void run(const char* name, double ops) {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double* out; cudaMalloc(&out, (size_t)blocks * threads * 8);
    const double q = 1125899906826241.0, w = 734623412345677.0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 12345.0, w, w / q, q, 16);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 12345.0, w, w / q, q, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double per_s = (double)blocks * threads * iters * 8 * ops / (ms * 1e-3);
    printf("%-34s %8.3f ms  %8.2f Gop/s  %6.2f op/clk/SM\n", name, ms, per_s * 1e-9, per_s / ((double)clk * 1e3) / sms);
}
int main() {
    run<0>("DFMA", 1);
    run<1>("fp64 lazy modmul + sign fix", 1);
    const int n = 1 << 22;
    u64* h = (u64*)malloc(n * 8);
    const u64 q = 1125899906826241ull, w = 734623412345677ull;          // q < 2^50, q = 1 mod 2^17
    srand(1);
    for (int i = 0; i < n; i++) h[i] = ((((u64)rand() << 31) ^ rand()) << 10 ^ rand()) % (8 * q);   // lazy inputs up to 8q < 2^53
    u64* d; int* bad; cudaMalloc(&d, n * 8); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
    check<<<n / 256, 256>>>(d, w, q, n, bad);
    int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("exactness: %d mismatches out of %d random lazy inputs (< 8q)\n", hb, n);
    return 0;
}
