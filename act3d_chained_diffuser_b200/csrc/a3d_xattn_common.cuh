// Device helpers shared by the fused cross-attention kernels (a3d_xattn2.cu: mma.sync core,
// a3d_xattn4/5/6.cu: tcgen05 / TMEM cores): packed-weight layout, register-chained split GEMM,
// fragment-layout LayerNorm, exp2 variants.
#pragma once
#include "a3d_mma_gemm.cuh"

namespace a3d {

struct Xa2 {
    static constexpr int E = 60, H = 4, EP = 64, ROWS = 128;
    static constexpr int XP = 72;                     // floats per row of the parked residual tile
    static constexpr int QP = 72;                     // halfs per row of the fp16 Q tile
    static constexpr int TILE_BYTES = 2 * H * 2048;   // K image + V image of one 64-key tile
    static constexpr int STAGES = 3;
    static constexpr size_t X_BYTES = (size_t)ROWS * XP * 4;
    static constexpr size_t Q_BYTES = (size_t)ROWS * QP * 2;
    static constexpr size_t SMEM = X_BYTES + Q_BYTES + (size_t)STAGES * TILE_BYTES + 64;
    // packed weights of one layer: four fragment-ordered [K=64][N=64] matrices (uint4 units) ...
    static constexpr int MAT = 4 * 8 * 32;            // uint4 per matrix
    static constexpr int W_Q = 0, W_O = MAT, W_1 = 2 * MAT, W_2 = 3 * MAT;
    // ... followed by the same four matrices as tcgen05 operand images (a3d_xattn6.cu; packing.pack_umma_weight):
    // image m at W_IMG + m * MAT, 16 KiB each = [hi | lo][4 k-slabs][64 out][16 in] fp16, SWIZZLE_32B rows
    static constexpr int W_IMG = 4 * MAT, LAYER_W = 8 * MAT;
    // ... and eight fp32 vectors of 64
    static constexpr int B_Q = 0, B_O = 64, G_1 = 128, BE_1 = 192, B_1 = 256, B_2 = 320, G_2 = 384, BE_2 = 448,
                         LAYER_V = 512;
};

struct Xa2Args {
    const float* x0;
    long x0_sb, x0_sn;
    const float* qpos;
    int batch, nq, nk, ntiles, nlayers;
    const unsigned char* kv_base;
    size_t kv_layer_stride;
    const uint4* w;
    const float* v;
    float* feat_out;
    int feat_rows, feat_all;
    const float* qvec;
    int nqv;
    float* logits;
};

__device__ __forceinline__ void mma_16816_c(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1,
                                            const float (&c)[4]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

// 2^x on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial on [-0.5, 0.5], max relative
// error 7.5e-5 -- well below the fp16 rounding of P): used for a fraction of the scores so that the MUFU
// unit (16 results / clk / SM, the measured ceiling of this kernel) and the FMA pipe work in parallel.
__device__ __forceinline__ float exp2_poly(float x) {
    x = fminf(fmaxf(x, -126.0f), 126.0f);       // outside: the exponent insertion below would wrap
    const float t = __fadd_rn(x, 12582912.0f);               // 1.5 * 2^23: rint(x) lands in the low mantissa bits
    const float f = __fsub_rn(x, __fsub_rn(t, 12582912.0f));  // x - rint(x) in [-0.5, 0.5]
    float p = fmaf(0.055170901f, f, 0.24260952f);
    p = fmaf(p, f, 0.69326097f);
    p = fmaf(p, f, 0.99992818f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
template <int PM, int IDX>
__device__ __forceinline__ float exp2_sel(float x) {
    if (PM & (1 << IDX)) return exp2_poly(x);
    return exp2_fast(x);
}

// 16 x 64 (this warp's rows) times a fragment-ordered [64][64] weight; the A operand is taken straight from
// accumulator-layout registers: k step ks <- column tiles 2ks, 2ks+1 (split into fp16 hi/lo on the fly).
__device__ __forceinline__ void gemm_reg(const float (&src)[8][4], const uint4* __restrict__ wfrag, int lane,
                                         float (&acc)[8][4]) {
    float cor[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[n][e] = 0.f;
            cor[n][e] = 0.f;
        }
    const uint4* wp = wfrag + lane;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t ah[4], al[4];
        split_h2(src[2 * ks][0], src[2 * ks][1], ah[0], al[0]);
        split_h2(src[2 * ks][2], src[2 * ks][3], ah[1], al[1]);
        split_h2(src[2 * ks + 1][0], src[2 * ks + 1][1], ah[2], al[2]);
        split_h2(src[2 * ks + 1][2], src[2 * ks + 1][3], ah[3], al[3]);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const uint4 b = __ldg(wp + (ks * 8 + n) * 32);
            mma_16816(acc[n], ah, b.x, b.y);
            mma_16816(cor[n], ah, b.z, b.w);
            mma_16816(cor[n], al, b.x, b.y);
        }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[n][e] = fmaf(cor[n][e], kLoScaleInv, acc[n][e]);
}

// LayerNorm of the two rows (g, g+8) this thread shares with its quad; columns 8n + 2q4 + {0,1}; eps 1e-5
__device__ __forceinline__ void layernorm_frag(float (&x)[8][4], int q4, const float* __restrict__ gamma,
                                               const float* __restrict__ beta) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float s = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n)
            if (8 * n + 2 * q4 < Xa2::E) s += x[n][2 * r] + x[n][2 * r + 1];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const float mean = s * (1.0f / Xa2::E);
        float v = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n)
            if (8 * n + 2 * q4 < Xa2::E) {
                const float d0 = x[n][2 * r] - mean, d1 = x[n][2 * r + 1] - mean;
                v = fmaf(d0, d0, v);
                v = fmaf(d1, d1, v);
            }
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        const float rstd = 1.0f / sqrtf(v * (1.0f / Xa2::E) + 1e-5f);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int c = 8 * n + 2 * q4;
            if (c < Xa2::E) {
                x[n][2 * r] = (x[n][2 * r] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
                x[n][2 * r + 1] = (x[n][2 * r + 1] - mean) * rstd * __ldg(gamma + c + 1) + __ldg(beta + c + 1);
            } else {
                x[n][2 * r] = 0.f;
                x[n][2 * r + 1] = 0.f;
            }
        }
    }
}

}  // namespace a3d
