// imad_peak.cu -- measures the integer-multiply throughput of the GPU the NTT runs on ("secondary bound
// integer pipe: measure IMAD peak on the box").  Each kernel keeps 8 independent dependency chains per thread.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a imad_peak.cu -o imad_peak && ./imad_peak
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

template <int OP>
__global__ void k(u64* out, u64 a, u64 b, int iters) {
    u64 x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a + threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) x[i] = (u64)((u32)x[i]) * (u32)b + x[i];                 // IMAD.WIDE.U32 (32x32+64)
            if (OP == 1) x[i] = (u32)((u32)x[i] * (u32)b + (u32)a);               // IMAD (32x32+32 low)
            if (OP == 2) x[i] = __umul64hi(x[i], b) + a;                          // 64x64 high
            if (OP == 3) x[i] = x[i] * b + a;                                     // 64x64 low
            if (OP == 4) { u64 h = __umul64hi(b, x[i]); x[i] = a * x[i] - h * b; } // Shoup lazy modmul
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
double run(const char* name, double ops_per_iter) {
    int dev = 0, sms = 0, clk = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    u64* out;
    cudaMalloc(&out, (size_t)blocks * threads * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 0x9E3779B97F4A7C15ull, 0x1FFFFFFFFFFFFF61ull, 16);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 0x9E3779B97F4A7C15ull, 0x1FFFFFFFFFFFFF61ull, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double total = (double)blocks * threads * iters * 8 * ops_per_iter;
    const double per_s = total / (ms * 1e-3);
    printf("%-34s %8.3f ms  %8.2f Gop/s  %6.2f op/clk/SM (at %d MHz max clock, %d SMs)\n", name, ms, per_s * 1e-9,
           per_s / ((double)clk * 1e3) / sms, clk / 1000, sms);
    cudaFree(out);
    return per_s;
}

int main() {
    run<0>("IMAD.WIDE.U32 (32x32+64)", 1);
    run<1>("IMAD (32x32+32)", 1);
    run<2>("mul.hi.u64", 1);
    run<3>("mul.lo.u64 + add", 1);
    const double shoup = run<4>("Shoup lazy modmul (hi + 2 lo)", 1);
    printf("NTT integer bound: %.3f us per 2^16-point limb transform (524288 butterflies at one Shoup modmul each)\n",
           524288.0 / shoup * 1e6);
    return 0;
}
