// Issue cost of the instructions the exponential paths are made of, per SM sub-partition: W warps per SMSP each run
// ILP independent dependent-chains of one instruction; reports cycles per warp-instruction per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAIN(NAME, ASM, ...)                                                            \
    __global__ void NAME(uint32_t* out, int iters, uint32_t seed) {                     \
        uint32_t x[8];                                                                   \
        uint32_t a = seed | 0x3f800000u, b = seed ^ 0x3f000000u;                         \
        _Pragma("unroll") for (int i = 0; i < 8; ++i) x[i] = (seed + i * 77u + threadIdx.x) | 0x3c003c00u; \
        for (int it = 0; it < iters; ++it) {                                             \
            _Pragma("unroll") for (int r = 0; r < 4; ++r)                                \
            _Pragma("unroll") for (int i = 0; i < 8; ++i) asm volatile(ASM : "+r"(x[i]) : __VA_ARGS__); \
        }                                                                                \
        uint32_t s = 0;                                                                  \
        _Pragma("unroll") for (int i = 0; i < 8; ++i) s ^= x[i];                         \
        if (s == 0x12345u) out[0] = s;                                                   \
    }

CHAIN(k_fadd_imm, "add.f32 %0, %0, 0f4B400000;", "r"(a))
CHAIN(k_fadd_reg, "add.f32 %0, %0, %1;", "r"(a))
CHAIN(k_ffma_rrr, "fma.rn.f32 %0, %0, %1, %2;", "r"(a), "r"(b))
CHAIN(k_ffma_imm, "fma.rn.f32 %0, %0, 0f3F800001, %1;", "r"(a))
CHAIN(k_fmul_reg, "mul.f32 %0, %0, %1;", "r"(a))
CHAIN(k_hfma2_rrr, "fma.rn.f16x2 %0, %0, %1, %2;", "r"(a), "r"(b))
CHAIN(k_hmul2, "mul.f16x2 %0, %0, %1;", "r"(a))
CHAIN(k_hadd2, "add.f16x2 %0, %0, %1;", "r"(a))
CHAIN(k_f2fp, "cvt.rn.f16x2.f32 %0, %0, %1;", "r"(a))
CHAIN(k_prmt, "prmt.b32 %0, %0, %1, 0x5410;", "r"(a))
CHAIN(k_lop3, "lop3.b32 %0, %0, %1, %2, 0x96;", "r"(a), "r"(b))
CHAIN(k_shl, "shl.b32 %0, %0, 1;", "r"(a))
CHAIN(k_iadd, "add.u32 %0, %0, %1;", "r"(a))
CHAIN(k_imad, "mad.lo.u32 %0, %0, %1, %2;", "r"(a), "r"(b))
CHAIN(k_vimnmx, "max.s16x2 %0, %0, %1;", "r"(a))
CHAIN(k_fmnmx, "max.f32 %0, %0, %1;", "r"(a))
CHAIN(k_ex2, "ex2.approx.ftz.f32 %0, %0;", "r"(a))
CHAIN(k_ex2h, "ex2.approx.f16x2 %0, %0;", "r"(a))
CHAIN(k_lea, "{ .reg .b32 t; shl.b32 t, %1, 23; add.u32 %0, %0, t; }", "r"(a))

#define CHAIN64(NAME, ASM)                                                              \
    __global__ void NAME(uint32_t* out, int iters, uint32_t seed) {                     \
        unsigned long long x[8];                                                         \
        unsigned long long a = ((unsigned long long)(seed | 0x3f800000u) << 32) | (seed | 0x3f800000u), b = a ^ 0x0010000000100000ull; \
        _Pragma("unroll") for (int i = 0; i < 8; ++i) x[i] = a + i * 77u + threadIdx.x;  \
        for (int it = 0; it < iters; ++it) {                                             \
            _Pragma("unroll") for (int r = 0; r < 4; ++r)                                \
            _Pragma("unroll") for (int i = 0; i < 8; ++i) asm volatile(ASM : "+l"(x[i]) : "l"(a), "l"(b)); \
        }                                                                                \
        unsigned long long s = 0;                                                        \
        _Pragma("unroll") for (int i = 0; i < 8; ++i) s ^= x[i];                         \
        if (s == 0x12345u) out[0] = (uint32_t)s;                                         \
    }
CHAIN64(k_ffma2, "fma.rn.f32x2 %0, %0, %1, %2;")
CHAIN64(k_fadd2, "add.rn.f32x2 %0, %0, %1;")
CHAIN64(k_fmul2, "mul.rn.f32x2 %0, %0, %1;")

typedef void (*kern_t)(uint32_t*, int, uint32_t);
static void run(const char* name, kern_t k, int sms, uint32_t* d) {
    for (int wps : {2, 4}) {
        const int iters = 2000;
        k<<<sms, wps * 4 * 32>>>(d, 10, 1);
        cudaDeviceSynchronize();
        cudaEvent_t a, b;
        cudaEventCreate(&a), cudaEventCreate(&b);
        cudaEventRecord(a);
        k<<<sms, wps * 4 * 32>>>(d, iters, 1);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double instr_per_smsp = (double)iters * 32 * wps;     // warp-instructions issued per SMSP
        printf("%-12s %d warps/SMSP: %6.2f cycles per warp-instruction per SMSP\n", name, wps, ms * 1e-3 * 1.965e9 / instr_per_smsp);
    }
}
int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* d;
    cudaMalloc(&d, 4);
#define R(k) run(#k, k, sms, d)
    R(k_ffma2); R(k_fadd2); R(k_fmul2); R(k_fadd_imm); R(k_fadd_reg); R(k_ffma_rrr); R(k_ffma_imm); R(k_fmul_reg); R(k_hfma2_rrr); R(k_hmul2); R(k_hadd2);
    R(k_f2fp); R(k_prmt); R(k_lop3); R(k_shl); R(k_iadd); R(k_imad); R(k_vimnmx); R(k_fmnmx); R(k_ex2); R(k_ex2h); R(k_lea);
    return 0;
}
