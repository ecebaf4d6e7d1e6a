// Microbenchmark: does the register-operand pattern limit FFMA2?  Same instruction, operands shared vs all distinct.
#include <cstdio>
#include <cuda_runtime.h>
#define R8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#define FFMA2_SH(i)  asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m2), "l"(c2));
#define FFMA2_D3(i)  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(d[i]));
#define FFMA2_D2(i)  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(m2));
#define FFMA2_BC(i)  asm volatile("{.reg .b64 t; mov.b64 t, {%2, %2}; fma.rn.f32x2 %0, %1, t, %0;}" : "+l"(a[i]) : "l"(b[i]), "f"(sc[i]));
#define FMUL2_D2(i)  asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b[i]));
#define LOP3(i)     asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(k1), "r"(k2));
#define FSELV(i)    asm volatile("selp.f32 %0, %0, %1, p;" : "+f"(s2[i]) : "f"(sd[i]));
#define IADDV(i)    asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(k1));
#define LDSV(i)     asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "+f"(qq[0]), "+f"(qq[1]), "+f"(qq[2]), "+f"(qq[3]) : "r"(saddr + 16 * i));
#define FFMA_SH(i)   asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(m), "f"(c));
#define FFMA_D3(i)   asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(s[i]) : "f"(sc[i]), "f"(sd[i]));
#define FFMA_D3b(i)  asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(s2[i]) : "f"(sd[i]), "f"(sc[i]));

template <int MODE> __global__ void __launch_bounds__(256) k(float* sink, int iters, float m, float c) {
    __shared__ float4 sm[64];
    float s[8], s2[8], sc[8], sd[8], qq[4] = {0, 0, 0, 0}; unsigned long long a[8], b[8], d[8]; unsigned u[8];
    const unsigned k1 = __float_as_uint(m), k2 = __float_as_uint(c);
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(1, 2, 3, 4);
    __syncthreads();
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
    for (int i = 0; i < 8; ++i) u[i] = threadIdx.x + i;
    asm volatile("{.reg .pred p; setp.ne.u32 p, %0, 0;" :: "r"(k1 & 1));
    const unsigned long long m2 = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
    const unsigned long long c2 = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int i = 0; i < 8; ++i) { s[i] = threadIdx.x * 0.5f + i; s2[i] = s[i] + 3; sc[i] = m + i * 1e-7f; sd[i] = c + i * 1e-7f; a[i] = m2 + i; b[i] = m2 + 16 * i; d[i] = c2 + 32 * i; }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { R8(FFMA2_SH) }
        if (MODE == 1) { R8(FFMA2_D3) }
        if (MODE == 2) { R8(FFMA2_D2) }
        if (MODE == 3) { R8(FFMA2_BC) }
        if (MODE == 4) { R8(FMUL2_D2) }
        if (MODE == 5) { R8(FFMA_SH) }
        if (MODE == 6) { R8(FFMA_D3) }
        if (MODE == 7) { R8(FFMA_D3) R8(FFMA_D3b) }
        if (MODE == 8) { R8(FFMA2_D3) R8(LOP3) }
        if (MODE == 9) { R8(FFMA2_SH) R8(LOP3) }
        if (MODE == 10) { R8(FFMA2_D3) R8(FSELV) }
        if (MODE == 11) { R8(FFMA2_D3) R8(IADDV) }
        if (MODE == 12) { R8(FFMA2_D3) R8(LDSV) }
        if (MODE == 13) { R8(FFMA2_D3) R8(FFMA2_D3) R8(FFMA2_D3) R8(LOP3) R8(FSELV) R8(IADDV) }
        if (MODE == 14) { R8(FFMA2_SH) R8(FFMA2_SH) R8(FFMA2_SH) R8(LOP3) R8(FSELV) R8(IADDV) }
        if (MODE == 15) { R8(FFMA2_BC) R8(LOP3) }
    }
    asm volatile("}");
    float t = qq[0] + qq[1] + qq[2] + qq[3];
    for (int i = 0; i < 8; ++i) t += (float)u[i] + s[i] + s2[i] + (float)a[i] + sc[i] + sd[i] + (float)b[i] + (float)d[i];
    if (t == -1.2345f) sink[0] = t;
}
template <int MODE> void run(const char* name, int n8) {
    float* sink; cudaMalloc(&sink, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 20000, grid = p.multiProcessorCount * 4;
    k<MODE><<<grid, 256>>>(sink, 100, 0.999999f, 1e-6f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<grid, 256>>>(sink, iters, 0.999999f, 1e-6f); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9;
    printf("%-40s %6.3f warp-inst/clk/SMSP (%.2f clk each)\n", name, 8.0 * iters * 8.0 * n8 / cyc, cyc / (8.0 * iters * 8.0 * n8));
    cudaFree(sink);
}
int main() {
    run<0>("FFMA2 shared multiplicand+addend", 1); run<2>("FFMA2 two distinct reg pairs", 1); run<1>("FFMA2 three distinct reg pairs", 1);
    run<3>("FFMA2 broadcast scalar operand", 1); run<4>("FMUL2 two distinct", 1);
    run<8>("FFMA2 (3 distinct) + LOP3", 2); run<9>("FFMA2 (shared operands) + LOP3", 2); run<15>("FFMA2 (broadcast) + LOP3", 2);
    run<10>("FFMA2 (3 distinct) + FSEL", 2); run<11>("FFMA2 (3 distinct) + IADD", 2); run<12>("FFMA2 (3 distinct) + LDS.128", 2);
    run<13>("3 FFMA2 (3 distinct) + LOP3 + FSEL + IADD", 6); run<14>("3 FFMA2 (shared) + LOP3 + FSEL + IADD", 6);
    run<5>("FFMA shared", 1); run<6>("FFMA three distinct", 1); run<7>("2 x FFMA three distinct", 2);
    return 0;
}
