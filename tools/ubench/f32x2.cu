// Microbenchmark: issue/pipe throughput of packed f32x2 arithmetic (FFMA2/FMUL2/FADD2) against scalar
// FFMA on sm_100a, alone and mixed with integer ALU work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(256) k(float* sink, int iters) {
    float2 a[8];
    float s[8];
    unsigned u[4];
    for (int i = 0; i < 8; ++i) { a[i] = make_float2(threadIdx.x + i, threadIdx.x - i); s[i] = threadIdx.x * 0.5f + i; }
    for (int i = 0; i < 4; ++i) u[i] = threadIdx.x * 7 + i;
    const float2 m = make_float2(0.999999f, 1.000001f), c = make_float2(1e-6f, -1e-6f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { s[i] = __fmaf_rn(s[i], m.x, c.x); }                       // 8 FFMA
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, c); }                          // 8 FFMA2
            if (MODE == 2) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fmaf_rn(a[i].y, m.y, c.y); }  // 16 FFMA
            if (MODE == 3) { a[i] = __fmul2_rn(a[i], m); }                             // 8 FMUL2
            if (MODE == 4) { a[i] = __fadd2_rn(a[i], c); }                             // 8 FADD2
            if (MODE == 5) { a[i] = __ffma2_rn(a[i], m, c); u[i & 3] = (u[i & 3] ^ (u[(i + 1) & 3] >> 3)) + 0x9e3779b9u; }  // 8 FFMA2 + ~16 ALU
            if (MODE == 6) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fmaf_rn(a[i].y, m.y, c.y);
                             u[i & 3] = (u[i & 3] ^ (u[(i + 1) & 3] >> 3)) + 0x9e3779b9u; }                            // 16 FFMA + ALU
            if (MODE == 7) { s[i] = __fadd_rn(s[i], c.x); }                            // 8 FADD
        }
    }
    float t = 0;
    for (int i = 0; i < 8; ++i) t += a[i].x + a[i].y + s[i];
    for (int i = 0; i < 4; ++i) t += (float)u[i];
    if (t == -1.2345f) sink[0] = t;
}

template <int MODE> void run(const char* name, double lane_ops_per_iter) {
    float* sink; cudaMalloc(&sink, 4);
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int iters = 20000, grid = p.multiProcessorCount * 8;
    k<MODE><<<grid, 256>>>(sink, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(sink, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)grid * 256 * iters * lane_ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f T lane-flop-instr/s (FMA counted once)  err=%s\n", name, ms, ops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    run<0>("8 FFMA", 8);
    run<1>("8 FFMA2 (16 fma)", 16);
    run<2>("16 FFMA", 16);
    run<3>("8 FMUL2 (16 mul)", 16);
    run<4>("8 FADD2 (16 add)", 16);
    run<7>("8 FADD", 8);
    run<5>("8 FFMA2 + int ALU", 16);
    run<6>("16 FFMA + int ALU", 16);
    return 0;
}
