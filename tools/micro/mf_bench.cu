// Exponential throughput of one SM when two row groups alternate "M units" (64 exponentials on the MUFU unit) and
// "F units" (P of 64 exponentials as a polynomial on the FMA pipe): the scheme of the fused ghost-point attention
// kernel (a3d_xattn6.cu).  8 warps per CTA, 1 CTA per SM: warps 0-3 = row group 0, warps 4-7 = row group 1; warp w
// and warp w+4 share one SM sub-partition.  A "unit" = 64 scores per thread (one 64-key tile of one head).
//   MODE 0: every unit on the MUFU unit (the xattn4 scheme) .................... floor 512 cycles per (128-row) unit
//   MODE 1: units alternate M / F, group 1 starts with F, no hand-shake
//   MODE 2: same, an M unit may only start when the partner warp has finished its previous M unit (mbarrier token)
//   MODE 3: every unit mixes P polynomial + (64-P) MUFU exponentials inside the warp (what ptxas clusters)
// POLY 0: fp32 Cody-Waite + degree-3 polynomial (8 instructions per score); POLY 1: fp32 range reduction, polynomial and
// scaling on packed halves (7 per score, no separate pack).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float poly32(float x) {
    x = fmaxf(x, -126.0f);
    const float t = __fadd_rn(x, 12582912.0f);
    const float f = __fsub_rn(x, __fsub_rn(t, 12582912.0f));
    float p = fmaf(0.055170901f, f, 0.24260952f);
    p = fmaf(p, f, 0.69326097f);
    p = fmaf(p, f, 0.99992818f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// two scores -> packed fp16 pair, polynomial on the packed halves
__device__ __forceinline__ uint32_t poly16x2(float x0, float x1) {
    constexpr float M = 12582912.0f + 15.0f;        // low mantissa bits of t hold rint(x) + 15 (the fp16 exponent bias)
    const float t0 = __fadd_rn(x0, M), t1 = __fadd_rn(x1, M);
    const float f0 = __fsub_rn(x0, __fsub_rn(t0, M)), f1 = __fsub_rn(x1, __fsub_rn(t1, M));
    const __half2 f = __floats2half2_rn(f0, f1);
    __half2 p = __hfma2(__float2half2_rn(0.055170901f), f, __float2half2_rn(0.24260952f));
    p = __hfma2(p, f, __float2half2_rn(0.69326097f));
    p = __hfma2(p, f, __float2half2_rn(0.99992818f));
    uint32_t n = __byte_perm(__float_as_uint(t0), __float_as_uint(t1), 0x5410);      // (n0 + 15) | (n1 + 15) << 16
    asm("max.s16x2 %0, %0, %1;" : "+r"(n) : "r"(0u));
    asm("min.s16x2 %0, %0, %1;" : "+r"(n) : "r"(0x001f001fu));
    n <<= 10;                                                                          // 2^n as packed halves (0 -> flush)
    const __half2 r = __hmul2(p, *reinterpret_cast<const __half2*>(&n));
    return *reinterpret_cast<const uint32_t*>(&r);
}

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}

template <int NP, int POLY>
__device__ __forceinline__ void unit(const float4* src, uint4* dst, int tid) {
    float x[64];
    uint32_t p[32];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(src + c * 256 + tid)));
        x[4 * c] = v.x, x[4 * c + 1] = v.y, x[4 * c + 2] = v.z, x[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (2 * i < NP) {
            if (POLY == 1) p[i] = poly16x2(x[2 * i], x[2 * i + 1]);
            else p[i] = pack(poly32(x[2 * i]), poly32(x[2 * i + 1]));
        } else {
            p[i] = pack(ex2(x[2 * i]), ex2(x[2 * i + 1]));
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(dst + c * 256 + tid)), "r"(p[4 * c]), "r"(p[4 * c + 1]), "r"(p[4 * c + 2]),
                     "r"(p[4 * c + 3])
                     : "memory");
}

template <int MODE, int NP, int POLY>
__global__ void __launch_bounds__(256, 1) k(const float* in, uint32_t* out, int units) {
    extern __shared__ __align__(16) unsigned char smem[];
    float4* src = reinterpret_cast<float4*>(smem);                    // [16][256] float4
    uint4* dst = reinterpret_cast<uint4*>(smem + 65536);              // [8][256] uint4
    uint64_t* tok = reinterpret_cast<uint64_t*>(smem + 65536 + 32768);   // [2][4]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = warp >> 2, w = warp & 3;
    for (int c = 0; c < 16; ++c) src[c * 256 + tid] = reinterpret_cast<const float4*>(in)[(c * 256 + tid) % 1024];
    if (tid < 8) mbar_init(tok + tid, 1);
    __syncthreads();
    int km = 0;
    for (int u = 0; u < units; ++u) {
        if (MODE == 0) {
            unit<0, POLY>(src, dst, tid);
        } else if (MODE == 3) {
            unit<NP, POLY>(src, dst, tid);
        } else {
            const bool m_unit = ((u + g) & 1) == 0;
            if (m_unit) {
                if (MODE == 2) {
                    if (g == 0) {
                        if (km > 0) mbar_wait(tok + 4 + w, (km - 1) & 1);
                    } else {
                        mbar_wait(tok + w, km & 1);
                    }
                }
                unit<0, POLY>(src, dst, tid);
                if (MODE == 2) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tok + g * 4 + w);
                }
                ++km;
            } else {
                unit<NP, POLY>(src, dst, tid);
            }
        }
    }
    __syncthreads();
    if (dst[tid].x == 0x12345678u) out[0] = dst[tid].y;
}

template <int MODE, int NP, int POLY>
static void run(const char* what, const float* d_in, uint32_t* d_out, int sms) {
    const int smem = 65536 + 32768 + 64, units = 4000;
    cudaFuncSetAttribute(k<MODE, NP, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<MODE, NP, POLY><<<sms, 256, smem>>>(d_in, d_out, 50);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE, NP, POLY><<<sms, 256, smem>>>(d_in, d_out, units);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    // one "128-row unit" = 4 warps x 1 unit; the SM runs 2 groups -> 2 * units of them
    const double cyc = ms * 1e-3 * 1.965e9 / (2.0 * units);
    printf("%-58s %8.3f ms  %7.1f cycles per 128-row unit per SM (MUFU floor 512)  %s\n", what, ms, cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* h = new float[4096];
    for (int i = 0; i < 4096; ++i) h[i] = -((i * 37) % 2000) * 0.01f;       // scores in [-20, 0]
    float* d_in;
    uint32_t* d_out;
    cudaMalloc(&d_in, 4096 * 4);
    cudaMalloc(&d_out, 4);
    cudaMemcpy(d_in, h, 4096 * 4, cudaMemcpyHostToDevice);
    run<0, 0, 0>("all MUFU", d_in, d_out, sms);
    run<3, 16, 0>("in-warp mix 16/64 poly32", d_in, d_out, sms);
    run<3, 24, 0>("in-warp mix 24/64 poly32", d_in, d_out, sms);
    run<3, 32, 0>("in-warp mix 32/64 poly32", d_in, d_out, sms);
    run<3, 24, 1>("in-warp mix 24/64 poly16x2", d_in, d_out, sms);
    run<3, 32, 1>("in-warp mix 32/64 poly16x2", d_in, d_out, sms);
    run<1, 64, 0>("alternate M/F(64 poly32), free", d_in, d_out, sms);
    run<2, 64, 0>("alternate M/F(64 poly32), token", d_in, d_out, sms);
    run<1, 48, 0>("alternate M/F(48 poly32), free", d_in, d_out, sms);
    run<2, 48, 0>("alternate M/F(48 poly32), token", d_in, d_out, sms);
    run<2, 56, 0>("alternate M/F(56 poly32), token", d_in, d_out, sms);
    run<2, 40, 0>("alternate M/F(40 poly32), token", d_in, d_out, sms);
    run<1, 64, 1>("alternate M/F(64 poly16x2), free", d_in, d_out, sms);
    run<2, 64, 1>("alternate M/F(64 poly16x2), token", d_in, d_out, sms);
    run<2, 56, 1>("alternate M/F(56 poly16x2), token", d_in, d_out, sms);
    run<2, 48, 1>("alternate M/F(48 poly16x2), token", d_in, d_out, sms);
    run<3, 64, 0>("all poly32 (FMA pipe only)", d_in, d_out, sms);
    run<3, 64, 1>("all poly16x2 (FMA pipe only)", d_in, d_out, sms);
    return 0;
}
