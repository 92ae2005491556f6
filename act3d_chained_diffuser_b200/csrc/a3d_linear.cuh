// Register-tiled fp32 mini-GEMM used by the projection / epilogue stages.
//
//   acc[r][c] = sum_k AT[k][2*rg + r] * WT[k][16*cg + c]        r < 2, c < 16
//
// AT is the activation tile stored K-MAJOR in shared memory ([k][row], row pitch RP floats) so
// that the two rows a thread owns are one 8-byte load and the 8 row-groups of a warp read 64
// contiguous bytes.  WT is the weight stored K-MAJOR in global memory ([k][NP], NP = padded
// output width); it is tiny (<= 60 KB/layer), shared by every CTA and is read through L1 with
// 16-byte __ldg loads (the 4 column groups of a warp hit 4 distinct 64-byte segments).
// Accumulation is plain fp32 FMA in k order: these layers carry the residual stream and stay
// out of the reduced-precision budget of the attention core.
//
// Thread mapping (256 threads): a warp covers 8 row groups x 4 column groups:
//   rg = (warp % (RG/8)) * 8 + lane % 8      rows 2rg, 2rg+1
//   cg = (warp / (RG/8)) * 4 + lane / 8      cols 16cg .. 16cg+15
// so LayerNorm row statistics reduce over lanes with xor 8, 16 (plus nothing else when the
// output is 64 wide: RG = 64, 4 column groups).
#pragma once
#include "a3d_common.cuh"

namespace a3d {

template <int KD, int RP, int NP>
__device__ __forceinline__ void gemm_2x16(const float* __restrict__ at, const float* __restrict__ wt, int rg, int cg,
                                          float (&acc)[2][16]) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[r][c] = 0.f;
    const float* ap = at + 2 * rg;
    const float4* wp = reinterpret_cast<const float4*>(wt + 16 * cg);
#pragma unroll 4
    for (int k = 0; k < KD; ++k) {
        const float2 a = *reinterpret_cast<const float2*>(ap + k * RP);
        float w[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(wp + k * (NP / 4) + q);
            w[4 * q + 0] = v.x;
            w[4 * q + 1] = v.y;
            w[4 * q + 2] = v.z;
            w[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            acc[0][c] = fmaf(a.x, w[c], acc[0][c]);
            acc[1][c] = fmaf(a.y, w[c], acc[1][c]);
        }
    }
}

// LayerNorm over the E valid columns of a 64-wide row held by the 4 column-group lanes
// (lane, lane^8, lane^16, lane^24).  eps = 1e-5, biased variance, two-pass (torch semantics).
template <int E>
__device__ __forceinline__ void layernorm_rows(float (&x)[2][16], int cg, const float* __restrict__ gamma,
                                               const float* __restrict__ beta) {
    const int c0 = 16 * cg;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c0 + c < E) s += x[r][c];
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        const float mean = s * (1.0f / E);
        float v = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c0 + c < E) {
                const float d = x[r][c] - mean;
                v = fmaf(d, d, v);
            }
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        const float rstd = 1.0f / sqrtf(v * (1.0f / E) + 1e-5f);
#pragma unroll
        for (int c = 0; c < 16; ++c)
            x[r][c] = (c0 + c < E) ? (x[r][c] - mean) * rstd * __ldg(gamma + c0 + c) + __ldg(beta + c0 + c) : 0.f;
    }
}

}  // namespace a3d
