// Microbenchmark: issue/pipe rates of the instruction kinds in k_quad's hot loop on sm_100a, alone and mixed.
// Bodies are volatile inline PTX over 8 independent register chains per kind, 8 warps per SM sub-partition (SMSP).
// Reported: warp-instructions per clock per SMSP from clock64() of one CTA.
#include <cstdio>
#include <cuda_runtime.h>

#define FFMA(i)   asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(m), "f"(c));
#define FADD(i)   asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(c));
#define FMUL(i)   asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(m));
#define FFMA2(i)  asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m2), "l"(c2));
#define FMUL2(i)  asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(m2));
#define FADD2(i)  asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(c2));
#define LOP3(i)   asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(k1), "r"(k2));
#define IADD(i)   asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(k1));
#define FMNMX(i)  asm volatile("max.f32 %0, %0, %1;" : "+f"(w[i]) : "f"(c));
#define FSEL(i)   asm volatile("selp.f32 %0, %0, %1, p;" : "+f"(w[i]) : "f"(c));
#define MUFU(i)   asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(w[i]));
#define LDS(i)    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q[0]), "=f"(q[1]), "=f"(q[2]), "=f"(q[3]) : "r"(saddr + 16 * i));
#define SEL(i)    asm volatile("selp.b32 %0, %0, %1, p;" : "+r"(u[i]) : "r"(k1));
#define IMAD(i)   asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(k1), "r"(k2));
#define SHL(i)    asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[i]));
#define PRMT(i)   asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(u[i]) : "r"(k1));
#define FSETPSEL(i) asm volatile("{.reg .pred q; setp.lt.f32 q, %0, %1; selp.f32 %0, %0, %1, q;}" : "+f"(w[i]) : "f"(c));
#define ISETPSEL(i) asm volatile("{.reg .pred q; setp.lt.u32 q, %0, %1; selp.b32 %0, %0, %1, q;}" : "+r"(u[i]) : "r"(k2));
#define SHFL(i)   asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(u[i]));
#define LDS4(i)   asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "+f"(q[0]), "+f"(q[1]), "+f"(q[2]), "+f"(q[3]) : "r"(saddr + 16 * i));
#define LDS2(i)   asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "+f"(q[0]), "+f"(q[1]) : "r"(saddr + 16 * i));
#define LDS1(i)   asm volatile("ld.shared.f32 %0, [%1];" : "+f"(q[0]) : "r"(saddr + 16 * i));
#define R8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int MODE> __global__ void __launch_bounds__(256) k(float* sink, int iters, long long* cycles, float m, float c, unsigned k1, unsigned k2) {
    __shared__ float4 sm[64];
    float s[8], w[8], q[4] = {0, 0, 0, 0}; unsigned u[8]; unsigned long long a[8];
    const unsigned long long m2 = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
    const unsigned long long c2 = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int i = 0; i < 8; ++i) { s[i] = threadIdx.x * 0.5f + i; w[i] = s[i] + 1; u[i] = threadIdx.x * 7 + i; a[i] = m2 + i; }
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(1, 2, 3, 4);
    __syncthreads();
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
    asm volatile("{.reg .pred p; setp.ne.u32 p, %0, 0;" :: "r"(k1 & 1));   // predicate p lives for the whole kernel body
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { R8(FFMA) }
        if (MODE == 1) { R8(FFMA2) }
        if (MODE == 2) { R8(FMUL2) }
        if (MODE == 3) { R8(FADD2) }
        if (MODE == 4) { R8(FADD) }
        if (MODE == 5) { R8(LOP3) }
        if (MODE == 6) { R8(IADD) }
        if (MODE == 7) { R8(FMNMX) }
        if (MODE == 8) { R8(FSEL) }
        if (MODE == 9) { R8(MUFU) }
        if (MODE == 10) { R8(FFMA) R8(LOP3) }
        if (MODE == 11) { R8(FFMA2) R8(LOP3) }
        if (MODE == 12) { R8(FFMA2) R8(FFMA) }
        if (MODE == 13) { R8(FFMA2) R8(FADD) }
        if (MODE == 14) { R8(FFMA2) R8(LOP3) R8(FSEL) }
        if (MODE == 15) { R8(FFMA) R8(FMUL) R8(LOP3) R8(FSEL) }
        if (MODE == 16) { R8(FFMA2) R8(FFMA2) R8(FFMA2) R8(FFMA2) R8(MUFU) }
        if (MODE == 17) { R8(FFMA) R8(FSEL) }
        if (MODE == 18) { R8(FFMA2) R8(FSEL) }
        if (MODE == 19) { R8(FFMA2) R8(FFMA2) R8(FFMA2) R8(LOP3) R8(FSEL) R8(FADD) R8(LDS4) }   // roughly the hot loop's mix
        if (MODE == 20) { R8(FFMA) R8(FMUL) R8(FFMA) R8(FMUL) R8(FFMA) R8(FMUL) R8(LOP3) R8(FSEL) R8(FADD) R8(LDS4) }   // the same mix, unpacked
        if (MODE == 21) { R8(LDS4) }
        if (MODE == 24) { R8(SEL) }
        if (MODE == 25) { R8(IMAD) }
        if (MODE == 26) { R8(SHL) }
        if (MODE == 27) { R8(PRMT) }
        if (MODE == 28) { R8(FSETPSEL) }
        if (MODE == 29) { R8(ISETPSEL) }
        if (MODE == 30) { R8(SHFL) }
        if (MODE == 31) { R8(LDS2) }
        if (MODE == 32) { R8(LDS1) }
        if (MODE == 33) { R8(FFMA2) R8(IMAD) }
        if (MODE == 34) { R8(FFMA2) R8(LDS4) }
        if (MODE == 35) { R8(FFMA2) R8(LDS2) }
        if (MODE == 36) { R8(FFMA) R8(LDS4) }
        if (MODE == 37) { R8(FFMA2) R8(IADD) }
        if (MODE == 38) { R8(FFMA2) R8(FMNMX) }
        if (MODE == 39) { R8(FFMA) R8(IADD) }
        if (MODE == 40) { R8(FFMA2) R8(FFMA2) R8(IADD) R8(FMNMX) }
        if (MODE == 41) { R8(FFMA2) R8(SHFL) }
        if (MODE == 22) { R8(FFMA2) R8(FFMA2) R8(FFMA2) R8(LOP3) }
        if (MODE == 23) { R8(FFMA2) R8(FFMA2) R8(FFMA2) R8(FADD) }
    }
    const long long t1 = clock64();
    asm volatile("}");
    float t = q[0] + q[1] + q[2] + q[3];
    for (int i = 0; i < 8; ++i) t += s[i] + w[i] + (float)u[i] + (float)a[i];
    if (t == -1.2345f) sink[0] = t;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE> void run(const char* name, int groups_of_8) {
    float* sink; long long* cyc; cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 4000, grid = p.multiProcessorCount * 4;   // 4 CTAs x 8 warps = 32 warps per SM (8 per SMSP)
    k<MODE><<<grid, 256>>>(sink, 100, cyc, 0.999999f, 1e-6f, 3u, 5u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(sink, iters, cyc, 0.999999f, 1e-6f, 3u, 5u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    const long long h = (long long)(ms * 1e-3 * khz * 1e3);   // cycles at the nominal SM clock from the event time
    static bool once = false; if (!once) { once = true; printf("clock64 cycles %lld vs event-time cycles %lld (clock %d kHz)\n", hc, h, khz); }
    const double n = 8.0 /*warps per SMSP*/ * iters * 8.0 * groups_of_8;
    printf("%-44s %2d inst: %6.3f warp-inst/clk/SMSP = %6.2f clk per pass  %s\n", name, 8 * groups_of_8, n / (double)h, (double)h / (8.0 * iters),
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    run<0>("FFMA", 1); run<1>("FFMA2", 1); run<2>("FMUL2", 1); run<3>("FADD2", 1); run<4>("FADD", 1);
    run<5>("LOP3", 1); run<6>("IADD", 1); run<7>("FMNMX", 1); run<8>("FSEL", 1); run<9>("MUFU.RSQ", 1); run<21>("LDS.128", 1); run<31>("LDS.64", 1); run<32>("LDS.32", 1);
    run<24>("SEL", 1); run<25>("IMAD", 1); run<26>("SHL", 1); run<27>("PRMT", 1); run<28>("FSETP+FSEL", 2); run<29>("ISETP+SEL", 2); run<30>("SHFL", 1);
    run<33>("FFMA2 + IMAD", 2); run<34>("FFMA2 + LDS.128", 2); run<35>("FFMA2 + LDS.64", 2); run<36>("FFMA + LDS.128", 2);
    run<37>("FFMA2 + IADD", 2); run<38>("FFMA2 + FMNMX", 2); run<39>("FFMA + IADD", 2); run<40>("2 FFMA2 + IADD + FMNMX", 4); run<41>("FFMA2 + SHFL", 2);
    run<10>("FFMA + LOP3", 2); run<11>("FFMA2 + LOP3", 2); run<12>("FFMA2 + FFMA", 2); run<13>("FFMA2 + FADD", 2);
    run<17>("FFMA + FSEL", 2); run<18>("FFMA2 + FSEL", 2);
    run<22>("3 FFMA2 + LOP3", 4); run<23>("3 FFMA2 + FADD", 4);
    run<14>("FFMA2 + LOP3 + FSEL", 3); run<15>("FFMA + FMUL + LOP3 + FSEL", 4);
    run<16>("4 FFMA2 + MUFU", 5);
    run<19>("hot-loop mix packed (3 FFMA2,LOP3,FSEL,FADD,LDS)", 7);
    run<20>("hot-loop mix scalar (6 FFMA/FMUL,LOP3,FSEL,FADD,LDS)", 10);
    return 0;
}
