// Building blocks of the ChainedDiffuser denoiser kernels: one CTA (256 threads) owns the 64-row
// (50 waypoints + padding) token tile of one sample; activations live K-MAJOR in shared memory
// ([channel][row], pitch 66 floats) so that the register-tiled GEMM of a3d_linear.cuh reads them
// with 8-byte loads.  E = 120, H = 8, head_dim 15, FFN 480 (diffusion_head.py / layers.py:7-218).
#pragma once
#include "a3d_linear.cuh"

namespace a3d {
namespace cd {

constexpr int E = 120;       // embedding_dim of the shipped planner
constexpr int EP = 128;      // padded
constexpr int H = 8;
constexpr int HD = 15;
constexpr int FF = 480;
constexpr int FFP = 512;
constexpr int ROWS = 64;     // padded trajectory length
constexpr int RP = 66;       // row pitch of K-major tiles
constexpr int TILE = EP * RP;   // floats per K-major activation tile
constexpr int NPAIR = E / 2;

struct Map {   // GEMM thread mapping for a 64 x 128 output tile (see a3d_linear.cuh)
    int rg, cg, lane, warp, tid;
    __device__ Map() {
        tid = threadIdx.x;
        lane = tid & 31;
        warp = tid >> 5;
        rg = (warp & 3) * 8 + (lane & 7);
        cg = (warp >> 2) * 4 + (lane >> 3);
    }
};

// dst[c][r] = src[r][c] for r < nrows (zero padded); src row-major with `ld` floats per row
__device__ __forceinline__ void load_tile(float* __restrict__ dst, const float* __restrict__ src, int nrows, int ld,
                                          int ncols = E) {
    for (int i = threadIdx.x; i < ROWS * ncols; i += blockDim.x) {
        const int r = i / ncols, c = i - r * ncols;
        dst[c * RP + r] = (r < nrows) ? __ldg(src + (long)r * ld + c) : 0.f;
    }
}
__device__ __forceinline__ void store_tile(float* __restrict__ dst, const float* __restrict__ src, int nrows, int ld,
                                           int ncols = E) {
    for (int i = threadIdx.x; i < ROWS * ncols; i += blockDim.x) {
        const int r = i / ncols, c = i - r * ncols;
        if (r < nrows) dst[(long)r * ld + c] = src[c * RP + r];
    }
}

// (FFMA path, used only for the K = 9 trajectory-encoder layer; threads >= 256 idle)
// out[c][r] = act( sum_k in[k][r] * W^T[k][c] + bias[c] )  for c < 128 (one pass), optional ReLU.
// wt: [KD][NP] K-major, `col0` selects a 128-column window.
template <int KD, int NP, bool RELU>
__device__ __forceinline__ void linear_to_smem(const Map& m, const float* __restrict__ in, const float* __restrict__ wt,
                                               const float* __restrict__ bias, int col0, float* __restrict__ out) {
    if (threadIdx.x >= 256) return;
    float acc[2][16];
    gemm_2x16<KD, RP, NP>(in, wt + col0, m.rg, m.cg, acc);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float b = bias ? __ldg(bias + col0 + 16 * m.cg + c) : 0.f;
        float v0 = acc[0][c] + b, v1 = acc[1][c] + b;
        if (RELU) {
            v0 = fmaxf(v0, 0.f);
            v1 = fmaxf(v1, 0.f);
        }
        if (16 * m.cg + c < E) *reinterpret_cast<float2*>(out + (16 * m.cg + c) * RP + 2 * m.rg) = make_float2(v0, v1);
    }
}

// same, but accumulates into registers that the caller keeps (used to sum FFN chunks)
template <int KD, int NP>
__device__ __forceinline__ void linear_accumulate(const Map& m, const float* __restrict__ in, const float* __restrict__ wt,
                                                  float (&sum)[2][16]) {
    float acc[2][16];
    gemm_2x16<KD, RP, NP>(in, wt, m.rg, m.cg, acc);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        sum[0][c] += acc[0][c];
        sum[1][c] += acc[1][c];
    }
}

// projection followed by the 3-D rotary rotation of channel pairs, computed on the accumulator
// registers (a thread owns 8 complete pairs of 2 rows); xyz: [64][3] in shared memory.
template <int KD, int NP>
__device__ __forceinline__ void linear_rope_to_smem(const Map& m, const float* __restrict__ in, const float* __restrict__ wt,
                                                    const float* __restrict__ bias, const float* __restrict__ xyz,
                                                    const float* __restrict__ freq, bool rope, float* __restrict__ out) {
    float acc[2][16];
    gemm_2x16<KD, RP, NP>(in, wt, m.rg, m.cg, acc);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float b = __ldg(bias + 16 * m.cg + c);
        acc[0][c] += b;
        acc[1][c] += b;
    }
    if (rope) {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int pi = 8 * m.cg + p;
            if (pi < NPAIR) {
                const int axis = pi / (E / 6), j = pi - axis * (E / 6);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const float ang = __fmul_rn(xyz[(2 * m.rg + r) * 3 + axis], freq[j]);
                    float sv, cv;
                    sincosf(ang, &sv, &cv);
                    const float ev = acc[r][2 * p], od = acc[r][2 * p + 1];
                    acc[r][2 * p] = ev * cv - od * sv;
                    acc[r][2 * p + 1] = od * cv + ev * sv;
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c)
        *reinterpret_cast<float2*>(out + (16 * m.cg + c) * RP + 2 * m.rg) = make_float2(acc[0][c], acc[1][c]);
}

// x[c][r] <- LayerNorm_c( x[c][r] + add[c][r] ) * g + b, rows 0..63 (4 threads per row).  add may be null.
// LPR = lanes cooperating on one row = blockDim.x / 64 (4 for 256 threads, 8 for 512)
template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int LPR>
__device__ __forceinline__ float row_max(float v) {
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int LPR>
__device__ __forceinline__ void residual_layernorm(float* __restrict__ x, const float* __restrict__ add,
                                                   const float* __restrict__ g, const float* __restrict__ b) {
    constexpr int NC = E / LPR;                       // columns per lane (E = 120, LPR = 4 or 8: exact)
    static_assert(E % LPR == 0, "row passes assume E divisible by the lanes per row");
    const int r = threadIdx.x / LPR, part = threadIdx.x % LPR;
    // LayerNorm parameters first: their L1/L2 latency overlaps the reductions below
    float gg[NC], bb[NC], v[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        gg[i] = __ldg(g + part + i * LPR);
        bb[i] = __ldg(b + part + i * LPR);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = part + i * LPR;
        v[i] = x[c * RP + r];
        if (add) v[i] += add[c * RP + r];
        s += v[i];
    }
    s = row_sum<LPR>(s);
    const float mean = s * (1.0f / E);
    float v2 = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const float d = v[i] - mean;
        v2 = fmaf(d, d, v2);
    }
    v2 = row_sum<LPR>(v2);
    const float rstd = 1.0f / sqrtf(v2 * (1.0f / E) + 1e-5f);
#pragma unroll
    for (int i = 0; i < NC; ++i) x[(part + i * LPR) * RP + r] = (v[i] - mean) * rstd * gg[i] + bb[i];
}

// dst[c][r] = (src[c][r] + pe[r][c]) * (1 + scale[c]) + shift[c]; pe / scale / shift may be null
// (AdaLN, layers.py:282-290; the waypoint-index embedding is the reference's seq1_sem_pos).
__device__ __forceinline__ void modulate(float* __restrict__ dst, const float* __restrict__ src, const float* __restrict__ pe,
                                         const float* __restrict__ scale, const float* __restrict__ shift, int nrows) {
    for (int i = threadIdx.x; i < ROWS * E; i += blockDim.x) {
        const int c = i / ROWS, r = i - c * ROWS;
        float v = src[c * RP + r];
        if (pe && r < nrows) v += __ldg(pe + r * E + c);
        if (scale) v = v * (1.0f + __ldg(scale + c)) + __ldg(shift + c);
        dst[c * RP + r] = v;
    }
}

// Small multi-head attention entirely in shared memory, fp32 (self-attention over <= 64 waypoints,
// or attention to the 53 instruction tokens).  q, k, v, out: K-major tiles; q carries
// hd^-1/2 * log2(e).  key_mask[j] != 0 => key j ignored (key_padding_mask, layers.py:178).
// scores: [64][65] scratch.  out may alias q (each head's q columns are consumed before its
// output columns are written).
template <int LPR>
__device__ __forceinline__ void small_mha(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                          int nk, const unsigned char* __restrict__ key_mask, float* __restrict__ scores,
                                          float* __restrict__ out) {
    const int i = threadIdx.x / LPR, part = threadIdx.x % LPR;
    for (int h = 0; h < H; ++h) {
        const int d0 = h * HD;
        float qv[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) qv[d] = q[(d0 + d) * RP + i];
        float mx = -INFINITY;
        for (int j = part; j < nk; j += LPR) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;   // three independent chains hide the FMA latency
#pragma unroll
            for (int d = 0; d < HD; d += 3) {
                s0 = fmaf(qv[d], k[(d0 + d) * RP + j], s0);
                s1 = fmaf(qv[d + 1], k[(d0 + d + 1) * RP + j], s1);
                s2 = fmaf(qv[d + 2], k[(d0 + d + 2) * RP + j], s2);
            }
            float s = (s0 + s1) + s2;
            if (key_mask && key_mask[j]) s = -INFINITY;
            scores[i * 65 + j] = s;
            mx = fmaxf(mx, s);
        }
        mx = row_max<LPR>(mx);
        float sum = 0.f;
        for (int j = part; j < nk; j += LPR) {
            const float p = exp2f(scores[i * 65 + j] - mx);
            scores[i * 65 + j] = p;
            sum += p;
        }
        sum = row_sum<LPR>(sum);
        const float inv = 1.0f / sum;
        __syncwarp();   // the LPR lanes of a row share its score row
        for (int d = part; d < HD; d += LPR) {
            float o0 = 0.f, o1 = 0.f;
            int j = 0;
            for (; j + 1 < nk; j += 2) {
                o0 = fmaf(scores[i * 65 + j], v[(d0 + d) * RP + j], o0);
                o1 = fmaf(scores[i * 65 + j + 1], v[(d0 + d) * RP + j + 1], o1);
            }
            if (j < nk) o0 = fmaf(scores[i * 65 + j], v[(d0 + d) * RP + j], o0);
            out[(d0 + d) * RP + i] = (o0 + o1) * inv;
        }
        __syncwarp();
    }
}

}  // namespace cd
}  // namespace a3d
