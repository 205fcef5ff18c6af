// FP32 pipe throughput on sm_100a: 3-register FFMA, FFMA with a constant-bank operand, FFMA2, FADD.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float cw[16];
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float x[16];
    float2 y[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = make_float2(x[2 * i], x[2 * i + 1]);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);            // 3 register operands
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], cw[i & 7], x[(i + 1) & 15]);   // reg, const, reg
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = __ffma2_rn(y[i], a2, b2);     // packed, 3 register pairs
            } else if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = x[i] + a;                    // FADD
            } else if (MODE == 4) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, x[(i + 1) & 15]);  // 3 distinct registers, chained
            } else if (MODE == 5) {
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = __ffma2_rn(y[i], a2, y[(i + 1) & 7]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += y[i].x + y[i].y;
    if (s == 12345.678f) out[0] = s;
}
template <int MODE>
void run(const char* name, int per_iter_lane_fma) {
    float* d; cudaMalloc(&d, 4);
    const int iters = 4096, blocks = 148 * 4, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 16, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = (double)blocks * threads * iters * 4 * per_iter_lane_fma;
    printf("%-40s %8.3f ms  %7.2f T lane-ops/s  = %6.1f lane-ops/clk/SM at 1.9 GHz\n", name, ms, lane_ops / ms * 1e-9,
           lane_ops / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
    float h[16]; for (int i = 0; i < 16; ++i) h[i] = 1.0f + i * 1e-4f;
    cudaMemcpyToSymbol(cw, h, sizeof h);
    run<0>("FFMA R,R,R,R (2 loop-invariant)", 16);
    run<4>("FFMA R,R,R,R (3 distinct, chained)", 16);
    run<1>("FFMA R,c[],R", 16);
    run<2>("FFMA2 (2 invariant pairs)", 16);
    run<5>("FFMA2 (3 distinct pairs)", 16);
    run<3>("FADD R,R,R", 16);
    return 0;
}
