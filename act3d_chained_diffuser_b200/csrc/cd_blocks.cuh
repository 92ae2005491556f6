// Building blocks of the ChainedDiffuser denoiser kernels: one CTA owns the 64-row (50 waypoints +
// padding) token tile of one sample; the fp32 residual stream lives K-MAJOR in shared memory
// ([channel][row], pitch 66 floats).  E = 120, H = 8, head_dim 15, FFN 480
// (diffusion_head.py / layers.py:7-218).  The tensor-core GEMMs and attention are in cd_denoiser.cu.
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

constexpr int FRAG = 32;                   // uint4 per (k step, n tile)
constexpr int U = 8 * 16 * FRAG;           // uint4 of one [K=128][N=128] weight

// ------------------------------------------------------------------ packed weight layouts
// fragment-ordered fp16 (hi, lo) weights, offsets in uint4 (packing.py mirrors these tables)
struct LangW { static constexpr int WQ = 0, WO = U, W1 = 2 * U, W2 = W1 + 8 * 64 * FRAG, SIZE = W2 + 32 * 16 * FRAG; };
struct AdaW {
    static constexpr int C_WQ = 0, C_WO = U, S_WQ = 2 * U, S_WK = 3 * U, S_WV = 4 * U, S_WO = 5 * U, W1 = 6 * U,
                         W2 = W1 + 8 * 64 * FRAG, SIZE = W2 + 32 * 16 * FRAG;
};
struct MlpW { static constexpr int W1 = 0, W2 = U, SIZE = 2 * U; };
// fp32 vectors (biases, LayerNorm), offsets in floats
struct LangV { static constexpr int BQ = 0, BO = EP, G12 = 2 * EP, B12 = 3 * EP, B1 = 4 * EP, B2 = B1 + FFP, G122 = B2 + EP, B122 = G122 + EP, SIZE = B122 + EP; };
struct AdaV {
    static constexpr int C_BQ = 0, C_BO = EP, G12 = 2 * EP, B12 = 3 * EP, S_BQ = 4 * EP, S_BK = 5 * EP, S_BV = 6 * EP,
                         S_BO = 7 * EP, G1 = 8 * EP, B1N = 9 * EP, B1 = 10 * EP, B2 = B1 + FFP, G122 = B2 + EP,
                         B122 = G122 + EP, SIZE = B122 + EP;
};
struct MlpV { static constexpr int B1 = 0, B2 = EP, SIZE = 2 * EP; };
constexpr int ADA_ROW = 3 * 2 * EP;   // per (timestep, layer): {adaln_12, adaln_1, adaln_ff1} x {scale, shift}

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

// x[c][r] <- LayerNorm_c( x[c][r] + add[c][r] ) * g + b, rows 0..63 (4 threads per row).  add may be null.
// LPR = lanes cooperating on one row = blockDim.x / 64 (4 for 256 threads, 8 for 512)
template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
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

}  // namespace cd
}  // namespace a3d
