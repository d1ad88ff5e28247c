// fp64_round.cu -- where does the rounding of the FP64 modular multiplication belong?  rint() on a double is FRND.F64,
// a conversion-pipe (XU) instruction; (t + M) - M with M = 3 * 2^51 is two FP64-pipe instructions (one when the product
// that forms t is fused into the addition).  Measures each alone and the two complete multiplications.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define MAGIC51 6755399441055744.0      /* 3 * 2^51: nearest integer for |t| < 2^51 */
#define MAGIC52 13510798882111488.0     /* 3 * 2^52: nearest even integer for |t| < 2^52 */

__device__ __forceinline__ double modmul_frnd(double a, double w, double wq, double q) {
    const double h = __dmul_rn(a, w);
    const double l = __fma_rn(a, w, -h);
    const double c = rint(__dmul_rn(a, wq));
    return __dadd_rn(__fma_rn(-c, q, h), l);
}
template <int WIDE>
__device__ __forceinline__ double modmul_magic(double a, double w, double wq, double q) {
    const double M = WIDE ? MAGIC52 : MAGIC51;
    const double h = __dmul_rn(a, w);
    const double l = __fma_rn(a, w, -h);
    const double c = __dadd_rn(__fma_rn(a, wq, M), -M);
    return __dadd_rn(__fma_rn(-c, q, h), l);
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
            if (OP == 1) x[i] = rint(x[i]) + 0.25;                                  // FRND + DADD
            if (OP == 2) x[i] = __dadd_rn(__dadd_rn(x[i], MAGIC51), -MAGIC51) + 0.25;   // 3 DADD
            if (OP == 3) x[i] = modmul_frnd(x[i], w, wq, q);
            if (OP == 4) x[i] = modmul_magic<0>(x[i], w, wq, q);
            if (OP == 5) x[i] = (double)(u64)(long long)x[i] + 0.25;                // F2I + I2F + DADD
            if (OP == 6) {                                                          // butterfly-shaped: modmul + add + sub
                const double v = modmul_frnd(x[i], w, wq, q);
                const double u = x[i ^ 1];
                x[i] = (i & 1) ? __dadd_rn(u, -v) : __dadd_rn(u, v);
            }
            if (OP == 7) {
                const double v = modmul_magic<0>(x[i], w, wq, q);
                const double u = x[i ^ 1];
                x[i] = (i & 1) ? __dadd_rn(u, -v) : __dadd_rn(u, v);
            }
        }
        if (OP >= 6) {                                                              // keep magnitudes bounded: fold
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const double c = OP == 6 ? rint(__dmul_rn(x[i], wq)) : __dadd_rn(__fma_rn(x[i], wq, MAGIC51), -MAGIC51);
                x[i] = __fma_rn(-c, q, x[i]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// exactness of the magic variants against 128-bit integer arithmetic, lazy inputs |a| < m q
template <int WIDE>
__global__ void check(const long long* a, u64 w, u64 q, int n, int* bad, double* worst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r = modmul_magic<WIDE>((double)a[i], (double)w, (double)w / (double)q, (double)q);
    __int128 want = ((__int128)a[i] * (__int128)w) % (__int128)q;
    __int128 got = (__int128)(long long)r % (__int128)q;
    if (want < 0) want += q;
    if (got < 0) got += q;
    if (want != got || r != (double)(long long)r) atomicAdd(bad, 1);
    const double m = fabs(r) / (double)q;
    if (m > 1.7) atomicAdd(bad + 1, 1);
    (void)worst;
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
    printf("%-44s %8.3f ms  %9.2f Gop/s  %6.2f op/clk/SM\n", name, ms, per_s * 1e-9, per_s / ((double)clk * 1e3) / sms);
    cudaFree(out);
}
int main() {
    run<0>("DFMA", 1);
    run<1>("FRND.F64 (+ DADD)", 1);
    run<2>("magic round (3 DADD)", 1);
    run<5>("F2I.S64.F64 + I2F.F64.U64 (+ DADD)", 1);
    run<3>("modmul, FRND quotient", 1);
    run<4>("modmul, magic quotient", 1);
    run<6>("butterfly + fold, FRND", 1);
    run<7>("butterfly + fold, magic", 1);
    const int n = 1 << 22;
    long long* h = (long long*)malloc(n * 8);
    const u64 q = 1125899906826241ull, w = 1125899906826240ull - 12345;          // w close to q: largest quotients
    for (int wide = 0; wide < 2; wide++) {
        const double lim = wide ? 3.9 : 1.9;                                     // |a w / q| < 2^52 / 2^51 at q ~ 2^50
        srand(1);
        for (int i = 0; i < n; i++) {
            const u64 v = ((((u64)rand() << 31) ^ rand()) << 10 ^ rand()) % (u64)(lim * (double)q);
            h[i] = (rand() & 1) ? (long long)v : -(long long)v;
        }
        long long* d; int* bad; cudaMalloc(&d, n * 8); cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
        cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
        if (wide) check<1><<<n / 256, 256>>>(d, w, q, n, bad, nullptr);
        else check<0><<<n / 256, 256>>>(d, w, q, n, bad, nullptr);
        int hb[2] = {-1, -1}; cudaMemcpy(hb, bad, 8, cudaMemcpyDeviceToHost);
        printf("magic%d exactness: %d mismatches, %d results beyond 1.7 q, out of %d lazy inputs |a| < %.1f q\n",
               wide ? 52 : 51, hb[0], hb[1], n, lim);
        cudaFree(d); cudaFree(bad);
    }
    return 0;
}
