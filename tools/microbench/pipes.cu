// Issue-rate microbenchmark for the FP32 instructions the step kernel is made of (sm_100a): FFMA, FFMA2 (packed pair),
// FFMA2 with a broadcast scalar operand, FMNMX, MUFU.  Each thread runs 8 independent dependency chains so that the
// pipes, not the latencies, set the rate.  Prints warp-instructions per cycle per SM sub-partition (4 per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
template <int MODE> __global__ void k(float *out, float a, float b) {
    float x[8];
    float2 y[8];
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 1e-3f + i; y[i] = make_float2(x[i], x[i] + 0.5f); }
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);
            if (MODE == 1) y[i] = __ffma2_rn(y[i], a2, b2);
            if (MODE == 2) y[i] = __ffma2_rn(y[i], make_float2(a, a), b2);
            if (MODE == 3) x[i] = fminf(fmaxf(x[i], a), b + x[i]);           // 2 FMNMX + 1 FADD
            if (MODE == 4) x[i] = __sinf(x[i]);
            if (MODE == 5) { x[i] = fmaf(x[i], a, b); y[i] = __ffma2_rn(y[i], a2, b2); }   // mixed
        }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += x[i] + y[i].x + y[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int inst_per_iter) {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(out, 1.0001f, 1e-6f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(out, 1.0001f, 1e-6f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // per SM: 4 blocks x 8 warps = 32 warps = 8 per sub-partition
    double winst = 8.0 * ITER * inst_per_iter * 8;       // warp-instructions per sub-partition
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %.3f ms  %.3f warp-inst/clk/SMSP (at the nominal %d MHz)\n", name, ms, winst / cycles, clk / 1000);
    cudaFree(out);
}
int main() {
    run<0>("FFMA", 1); run<1>("FFMA2", 1); run<2>("FFMA2 broadcast", 1); run<3>("2 FMNMX + FADD", 3); run<4>("MUFU.SIN", 1);
    run<5>("FFMA + FFMA2", 2);
    return 0;
}
