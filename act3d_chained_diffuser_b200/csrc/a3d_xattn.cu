// FIRST-GENERATION fused cross-attention stack (FFMA epilogue through shared memory).  Superseded by
// a3d_xattn2.cu; kept exported as a3d_xattn_stack_v1 for A/B profiling only (not in the public header).
// Fused cross-attention stack (Act3D ghost-point / query / vision-language stacks).
//
// One CTA owns 128 query rows of one sample for ALL layers of the stack; the residual stream
// never leaves shared memory.  Per layer:
//   (a) q = (x Wq^T + bq) * hd^-1/2 * log2(e), 3-D rotary from the query xyz, fp16, head-padded
//   (b) flash loop over the context K/V tile images (a3d_ctx_kv): a 2-stage mbarrier ring fed by
//       1-D bulk async copies (TMA engine), per warp 16 rows x 4 heads:
//       S = Q K^T (m16n8k16 tensor-core MMA, fp32 accumulate) -> running row max -> P = 2^(S-m)
//       as packed fp16 pairs (one MUFU op per two scores) -> O += P V, where slot 15 of every V
//       row is 1 so that the softmax denominator is accumulated by the same MMA in fp32
//   (c) out-proj + residual + LayerNorm, FFN(ReLU) + residual + LayerNorm as register-tiled fp32
//       GEMMs on the K-major shared tile
// and after the last layer the mask logits <qvec, x> (act3d.py:493-494) and/or the features.
// The score matrix, the attention weights and the rotary tables never exist in HBM.
#include <string.h>

#include "a3d_linear.cuh"

namespace a3d {

template <int E, int H, int FF>
struct XaCfg {
    static constexpr int EP = 16 * H;                 // 64
    static constexpr int ROWS = 128;
    static constexpr int RP = 130;                    // row pitch (floats) of the K-major tiles
    static constexpr int QP = 64;                     // row pitch (halfs) of the fp16 Q tile (hi | lo planes)
    static constexpr int TILE_BYTES = 2 * H * 2048;   // K image + V image of one 64-key tile
    static constexpr int STAGES = 2;
    static constexpr size_t XT_BYTES = (size_t)EP * RP * 4;
    static constexpr size_t SMEM = 2 * XT_BYTES + (size_t)STAGES * TILE_BYTES + 64;
    // packed weights of one layer (floats), all K-major and padded to EP columns:
    static constexpr int W_Q = 0;                     // [E][EP]   (pre-scaled by hd^-1/2 * log2 e)
    static constexpr int B_Q = W_Q + E * EP;          // [EP]
    static constexpr int W_O = B_Q + EP;              // [E][EP]
    static constexpr int B_O = W_O + E * EP;
    static constexpr int G_1 = B_O + EP;              // LayerNorm after attention
    static constexpr int BE_1 = G_1 + EP;
    static constexpr int W_1 = BE_1 + EP;             // [E][EP]  (FF == E here)
    static constexpr int B_1 = W_1 + E * EP;
    static constexpr int W_2 = B_1 + EP;              // [FF][EP]
    static constexpr int B_2 = W_2 + FF * EP;
    static constexpr int G_2 = B_2 + EP;
    static constexpr int BE_2 = G_2 + EP;
    static constexpr int LAYER_FLOATS = BE_2 + EP;
};

struct XaArgs {
    const float* x0;
    long x0_sb, x0_sn;
    const float* qpos;
    int batch, nq, nk, ntiles, nlayers;
    const unsigned char* kv_base;
    size_t kv_layer_stride;
    const float* w;
    float* feat_out;
    int feat_rows, feat_all;
    const float* qvec;
    int nqv;
    float* logits;
};

// EXP32: evaluate 2^(s-m) with the fp32 MUFU path per element (default: packed fp16x2, one op per two
// scores).  QSPLIT: carry Q as an fp16 hi + lo pair (two S MMAs) so that only K is rounded to fp16.
template <int E, int H, int FF, bool EXP32, bool QSPLIT>
__global__ void __launch_bounds__(256, 2) xattn_stack_kernel(const XaArgs a) {
    using C = XaCfg<E, H, FF>;
    static_assert(FF == E, "this instantiation keeps the FFN hidden tile in the 64-wide layout");
    extern __shared__ __align__(128) unsigned char smem[];
    float* xt = reinterpret_cast<float*>(smem);                                  // residual stream [EP][RP]
    float* at = reinterpret_cast<float*>(smem + C::XT_BYTES);                    // attention out / FFN hidden
    __half* qs = reinterpret_cast<__half*>(smem + C::XT_BYTES);                  // aliases `at` (phase a/b only)
    unsigned char* kvs = smem + 2 * C::XT_BYTES;                                 // STAGES x TILE_BYTES
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(kvs + C::STAGES * C::TILE_BYTES);
    uint64_t* bar_empty = bar_full + C::STAGES;
    __shared__ float freq[E / 6];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, row0 = blockIdx.x * C::ROWS;
    const int rg = warp * 8 + (lane & 7), cg = lane >> 3;     // GEMM mapping: rows 2rg,2rg+1 / cols 16cg..
    const int g = lane >> 2, q4 = lane & 3;                   // MMA mapping: rows g, g+8 / col pair q4

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, 8);
        }
        mbar_fence_init();
    }
    if (tid < E / 6) freq[tid] = rope_freq<E>(tid);

    // ---- residual stream init: xt[c][r] = x0[b, row0 + r, c]
    for (int i = tid; i < C::ROWS * E; i += 256) {
        const int r = i / E, c = i - r * E;
        float v = 0.f;
        if (row0 + r < a.nq) v = __ldg(a.x0 + (long)b * a.x0_sb + (long)(row0 + r) * a.x0_sn + c);
        xt[c * C::RP + r] = v;
    }
    // query positions of this thread's two GEMM rows
    float qx[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    if (a.qpos) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = row0 + 2 * rg + r;
            if (row < a.nq) {
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) qx[r][ax] = __ldg(a.qpos + ((long)b * a.nq + row) * 3 + ax);
            }
        }
    }
    __syncthreads();

    const unsigned char* kv_sample = a.kv_base + (size_t)b * a.ntiles * C::TILE_BYTES;
    const bool tail_mask = (a.nk % kTileKeys) != 0;
    uint32_t gtile = 0;   // global tile counter over all layers (drives stage / parity)

    for (int layer = 0; layer < a.nlayers; ++layer) {
        const float* w = a.w + (size_t)layer * C::LAYER_FLOATS;
        const unsigned char* kv_layer = kv_sample + (size_t)layer * a.kv_layer_stride;

        // producer: first tiles of this layer (stages are free: every warp passed the previous
        // layer's loop, and the empty barriers of those uses have completed)
        if (tid == 0) {
            for (int t = 0; t < C::STAGES && t < a.ntiles; ++t) {
                const uint32_t gt = gtile + t, s = gt % C::STAGES;
                if (gt >= (uint32_t)C::STAGES) mbar_wait(bar_empty + s, ((gt / C::STAGES) - 1) & 1);
                mbar_expect_tx(bar_full + s, C::TILE_BYTES);
                bulk_g2s(kvs + s * C::TILE_BYTES, kv_layer + (size_t)t * C::TILE_BYTES, C::TILE_BYTES, bar_full + s);
            }
        }

        // ---------------------------------------------------------------- (a) q projection + rotary
        {
            float acc[2][16];
            gemm_2x16<E, C::RP, C::EP>(xt, w + C::W_Q, rg, cg, acc);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float bias = __ldg(w + C::B_Q + 16 * cg + c);
                acc[0][c] += bias;
                acc[1][c] += bias;
            }
            if (a.qpos) {
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int pi = 8 * cg + p;          // pair index over the full embed vector
                    if (2 * pi < E) {
                        const int axis = pi / (E / 6), j = pi - axis * (E / 6);
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            const float ang = qx[r][axis] * freq[j];
                            float sv, cv;
                            if (fabsf(ang) < 3.0f) {
                                __sincosf(ang, &sv, &cv);
                            } else {
                                sincosf(ang, &sv, &cv);
                            }
                            const float ev = acc[r][2 * p], od = acc[r][2 * p + 1];
                            acc[r][2 * p] = ev * cv - od * sv;
                            acc[r][2 * p + 1] = od * cv + ev * sv;
                        }
                    }
                }
            }
            // fp16, head-padded: qs[row][h*16 + d]; padded embed dims own the pad slots (zero)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = 2 * rg + r;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int dim = 16 * cg + c;
                    int slot;
                    float val;
                    if (dim < E) {
                        const int h = dim / 15;
                        slot = h * 16 + (dim - 15 * h);
                        val = acc[r][c];
                    } else {
                        slot = (dim - E) * 16 + 15;
                        val = 0.f;
                    }
                    const __half hi = __float2half_rn(val);
                    qs[row * C::QP + slot] = hi;
                    if (QSPLIT) qs[C::ROWS * C::QP + row * C::QP + slot] = __float2half_rn(val - __half2float(hi));
                }
            }
        }
        __syncthreads();

        // ---------------------------------------------------------------- (b) attention core
        uint32_t qf[H][4];
        uint32_t ql[QSPLIT ? H : 1][4];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const int row = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
            ldmatrix_x4(qf[h], smem_u32(qs + row * C::QP + h * 16 + 8 * (lane >> 4)));
            if (QSPLIT) ldmatrix_x4(ql[h], smem_u32(qs + C::ROWS * C::QP + row * C::QP + h * 16 + 8 * (lane >> 4)));
        }
        __syncthreads();   // qs is dead from here on: `at` may be overwritten by the early warps

        float o[H][2][4];
        float mrow[H][2];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            mrow[h][0] = -INFINITY;
            mrow[h][1] = -INFINITY;
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) o[h][n][e] = 0.f;
        }

        for (int t = 0; t < a.ntiles; ++t) {
            const uint32_t gt = gtile + t, stage = gt % C::STAGES;
            if (tid == 0 && t >= 1 && t + 1 < a.ntiles) {   // prefetch tile t+1 into the stage tile t-1 used
                const uint32_t gn = gt + 1, sn = gn % C::STAGES;
                mbar_wait(bar_empty + sn, ((gn / C::STAGES) - 1) & 1);
                mbar_expect_tx(bar_full + sn, C::TILE_BYTES);
                bulk_g2s(kvs + sn * C::TILE_BYTES, kv_layer + (size_t)(t + 1) * C::TILE_BYTES, C::TILE_BYTES,
                         bar_full + sn);
            }
            mbar_wait(bar_full + stage, (gt / C::STAGES) & 1);

            const uint32_t kbase = smem_u32(kvs + stage * C::TILE_BYTES);
            const uint32_t vbase = kbase + H * 2048;
            const bool mask_this = tail_mask && (t == a.ntiles - 1);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float s[8][4];
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int key = kk * 16 + (lane & 7) + 8 * (lane >> 4);
                    const int chunk = (lane >> 3) & 1;
                    uint32_t r[4];
                    ldmatrix_x4(r, kbase + h * 2048 + key * 32 + ((chunk ^ ((key >> 2) & 1)) << 4));
                    mma_16816(s[2 * kk], qf[h], r[0], r[1]);
                    mma_16816(s[2 * kk + 1], qf[h], r[2], r[3]);
                    if (QSPLIT) {
                        mma_16816(s[2 * kk], ql[h], r[0], r[1]);
                        mma_16816(s[2 * kk + 1], ql[h], r[2], r[3]);
                    }
                }
                if (mask_this) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int key = t * kTileKeys + 8 * j + 2 * q4 + (e & 1);
                            if (key >= a.nk) s[j][e] = -INFINITY;
                        }
                }
                float mx0 = s[0][0], mx1 = s[0][2];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
                    mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
                }
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
                const float mn0 = fmaxf(mrow[h][0], mx0), mn1 = fmaxf(mrow[h][1], mx1);
                const float al0 = exp2_fast(mrow[h][0] - mn0), al1 = exp2_fast(mrow[h][1] - mn1);
                mrow[h][0] = mn0;
                mrow[h][1] = mn1;
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    o[h][n][0] *= al0;
                    o[h][n][1] *= al0;
                    o[h][n][2] *= al1;
                    o[h][n][3] *= al1;
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    uint32_t pa[4];
                    if (EXP32) {
                        pa[0] = pack_h2(exp2_fast(s[2 * kk][0] - mn0), exp2_fast(s[2 * kk][1] - mn0));
                        pa[1] = pack_h2(exp2_fast(s[2 * kk][2] - mn1), exp2_fast(s[2 * kk][3] - mn1));
                        pa[2] = pack_h2(exp2_fast(s[2 * kk + 1][0] - mn0), exp2_fast(s[2 * kk + 1][1] - mn0));
                        pa[3] = pack_h2(exp2_fast(s[2 * kk + 1][2] - mn1), exp2_fast(s[2 * kk + 1][3] - mn1));
                    } else {
                        pa[0] = exp2_pack_h2(s[2 * kk][0] - mn0, s[2 * kk][1] - mn0);
                        pa[1] = exp2_pack_h2(s[2 * kk][2] - mn1, s[2 * kk][3] - mn1);
                        pa[2] = exp2_pack_h2(s[2 * kk + 1][0] - mn0, s[2 * kk + 1][1] - mn0);
                        pa[3] = exp2_pack_h2(s[2 * kk + 1][2] - mn1, s[2 * kk + 1][3] - mn1);
                    }
                    const int key = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
                    const int chunk = lane >> 4;
                    uint32_t r[4];
                    ldmatrix_x4_trans(r, vbase + h * 2048 + key * 32 + ((chunk ^ ((key >> 2) & 1)) << 4));
                    mma_16816(o[h][0], pa, r[0], r[1]);
                    mma_16816(o[h][1], pa, r[2], r[3]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + stage);
        }
        gtile += a.ntiles;

        // normalise by the denominator carried in slot 15 and scatter to the K-major tile
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float l0 = __shfl_sync(0xffffffffu, o[h][1][1], (lane & ~3) | 3);
            const float l1 = __shfl_sync(0xffffffffu, o[h][1][3], (lane & ~3) | 3);
            const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int d = 8 * n + 2 * q4 + (e & 1);
                    const int row = warp * 16 + g + 8 * (e >> 1);
                    if (d < 15) at[(15 * h + d) * C::RP + row] = o[h][n][e] * ((e >> 1) ? i1 : i0);
                }
        }
        __syncthreads();

        // ---------------------------------------------------------------- (c) out-proj + res + LN
        float y[2][16];
        {
            gemm_2x16<E, C::RP, C::EP>(at, w + C::W_O, rg, cg, y);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float bias = __ldg(w + C::B_O + 16 * cg + c);
                const float2 res = *reinterpret_cast<const float2*>(xt + (16 * cg + c) * C::RP + 2 * rg);
                y[0][c] += bias + res.x;
                y[1][c] += bias + res.y;
            }
            layernorm_rows<E>(y, cg, w + C::G_1, w + C::BE_1);
#pragma unroll
            for (int c = 0; c < 16; ++c)
                *reinterpret_cast<float2*>(xt + (16 * cg + c) * C::RP + 2 * rg) = make_float2(y[0][c], y[1][c]);
        }
        __syncthreads();
        // ---------------------------------------------------------------- FFN: hidden = relu(x W1^T + b1)
        {
            float hdn[2][16];
            gemm_2x16<E, C::RP, C::EP>(xt, w + C::W_1, rg, cg, hdn);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float bias = __ldg(w + C::B_1 + 16 * cg + c);
                *reinterpret_cast<float2*>(at + (16 * cg + c) * C::RP + 2 * rg) =
                    make_float2(fmaxf(hdn[0][c] + bias, 0.f), fmaxf(hdn[1][c] + bias, 0.f));
            }
        }
        __syncthreads();
        {
            float z[2][16];
            gemm_2x16<FF, C::RP, C::EP>(at, w + C::W_2, rg, cg, z);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float bias = __ldg(w + C::B_2 + 16 * cg + c);
                y[0][c] += z[0][c] + bias;
                y[1][c] += z[1][c] + bias;
            }
            layernorm_rows<E>(y, cg, w + C::G_2, w + C::BE_2);
#pragma unroll
            for (int c = 0; c < 16; ++c)
                *reinterpret_cast<float2*>(xt + (16 * cg + c) * C::RP + 2 * rg) = make_float2(y[0][c], y[1][c]);
        }

        // ---------------------------------------------------------------- outputs of this layer
        const bool last = (layer == a.nlayers - 1);
        if (a.feat_out && (a.feat_all || last)) {
            float* fo = a.feat_out + ((size_t)(a.feat_all ? layer : 0) * a.batch + b) * (size_t)a.feat_rows * E;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = row0 + 2 * rg + r;
                if (row < a.nq) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const int col = 16 * cg + 4 * c4;
                        if (col < E)
                            *reinterpret_cast<float4*>(fo + (size_t)row * E + col) =
                                make_float4(y[r][4 * c4], y[r][4 * c4 + 1], y[r][4 * c4 + 2], y[r][4 * c4 + 3]);
                    }
                }
            }
        }
        if (last && a.logits) {
            for (int j = 0; j < a.nqv; ++j) {
                const float* qv = a.qvec + ((size_t)j * a.batch + b) * E;
                float p0 = 0.f, p1 = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int col = 16 * cg + c;
                    if (col < E) {
                        const float qc = __ldg(qv + col);
                        p0 = fmaf(y[0][c], qc, p0);
                        p1 = fmaf(y[1][c], qc, p1);
                    }
                }
                p0 += __shfl_xor_sync(0xffffffffu, p0, 8);
                p0 += __shfl_xor_sync(0xffffffffu, p0, 16);
                p1 += __shfl_xor_sync(0xffffffffu, p1, 8);
                p1 += __shfl_xor_sync(0xffffffffu, p1, 16);
                if (cg == 0) {
                    const int row = row0 + 2 * rg;
                    float* lo = a.logits + ((size_t)j * a.batch + b) * a.nq;
                    if (row < a.nq) lo[row] = p0;
                    if (row + 1 < a.nq) lo[row + 1] = p1;
                }
            }
        }
        __syncthreads();   // xt complete before the next layer's q projection reads it
    }
}

}  // namespace a3d

using namespace a3d;

// numerics variant of the attention core: bit0 = fp32 exp, bit1 = split-Q (see kernel comment)
static int g_xattn_variant = 0;
extern int g_xattn_poly;
extern int g_xattn_core;
extern "C" int a3d_set_option(const char* name, int value) {
    if (name && strcmp(name, "xattn_core") == 0) {
        A3D_REQUIRE(value == 0 || (value >= 2 && value <= 5),
                    "a3d_set_option: xattn_core must be 0 (auto), 2 (mma.sync), 3 (tcgen05, two-pass), 4 (tcgen05, single pass) "
                    "or 5 (tcgen05, single pass, warp-specialised exponentials)");
        g_xattn_core = value;
        return A3D_OK;
    }
    if (name && strcmp(name, "xattn_poly") == 0) {
        A3D_REQUIRE(value == 0 || value == 2 || value == 3 || value == 4, "a3d_set_option: xattn_poly must be 0, 2, 3 or 4");
        g_xattn_poly = value;
        return A3D_OK;
    }
    if (name && strcmp(name, "xattn_variant") == 0) {
        A3D_REQUIRE(value >= 0 && value <= 3, "a3d_set_option: xattn_variant must be 0..3");
        g_xattn_variant = value;
        return A3D_OK;
    }
    A3D_REQUIRE(false, "a3d_set_option: unknown option '%s'", name ? name : "(null)");
}

extern "C" size_t a3d_xattn_v1_layer_floats(int embed, int ffn) {
    if (embed == 60 && ffn == 60) return XaCfg<60, 4, 60>::LAYER_FLOATS;
    return 0;
}

extern "C" int a3d_xattn_stack_v1(const float* x0, long x0_stride_b, long x0_stride_n, const float* qpos, int batch,
                               int nq, int nk, int embed, int heads, int ffn, int nlayers, const void* kv_base,
                               size_t kv_layer_stride_bytes, const float* w, float* feat_out, int feat_rows,
                               int feat_all_layers, const float* qvec, int nqv, float* logits, void* stream) {
    A3D_REQUIRE(x0 && kv_base && w, "a3d_xattn_stack_v1: null pointer");
    A3D_REQUIRE(batch > 0 && nq > 0 && nk > 0 && nlayers > 0, "a3d_xattn_stack_v1: empty problem (B=%d nq=%d nk=%d L=%d)", batch, nq, nk, nlayers);
    A3D_REQUIRE(embed == 60 && heads == 4 && ffn == 60, "a3d_xattn_stack_v1: (embed, heads, ffn) = (%d,%d,%d) not supported; built for (60,4,60)", embed, heads, ffn);
    A3D_REQUIRE(!feat_out || feat_rows >= nq, "a3d_xattn_stack_v1: feat_rows %d < nq %d", feat_rows, nq);
    A3D_REQUIRE((logits == nullptr) == (qvec == nullptr || nqv == 0), "a3d_xattn_stack_v1: qvec/logits must be given together");
    A3D_REQUIRE(((uintptr_t)kv_base & 15) == 0 && (kv_layer_stride_bytes & 15) == 0, "a3d_xattn_stack_v1: K/V cache must be 16-byte aligned");
    A3D_REQUIRE(batch <= 65535, "a3d_xattn_stack_v1: batch %d exceeds grid.y", batch);
    using C = XaCfg<60, 4, 60>;
    XaArgs a;
    a.x0 = x0;
    a.x0_sb = x0_stride_b;
    a.x0_sn = x0_stride_n;
    a.qpos = qpos;
    a.batch = batch;
    a.nq = nq;
    a.nk = nk;
    a.ntiles = (nk + kTileKeys - 1) / kTileKeys;
    a.nlayers = nlayers;
    a.kv_base = (const unsigned char*)kv_base;
    a.kv_layer_stride = kv_layer_stride_bytes;
    a.w = w;
    a.feat_out = feat_out;
    a.feat_rows = feat_rows;
    a.feat_all = feat_all_layers;
    a.qvec = qvec;
    a.nqv = nqv;
    a.logits = logits;
    dim3 grid((nq + C::ROWS - 1) / C::ROWS, batch);
    const int variant = g_xattn_variant;
#define A3D_LAUNCH_XA(EXP32, QSPLIT)                                                                                  \
    do {                                                                                                               \
        cudaFuncSetAttribute(xattn_stack_kernel<60, 4, 60, EXP32, QSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             (int)C::SMEM);                                                                            \
        xattn_stack_kernel<60, 4, 60, EXP32, QSPLIT><<<grid, 256, C::SMEM, (cudaStream_t)stream>>>(a);                 \
    } while (0)
    switch (variant) {
        case 1: A3D_LAUNCH_XA(true, false); break;
        case 2: A3D_LAUNCH_XA(false, true); break;
        case 3: A3D_LAUNCH_XA(true, true); break;
        default: A3D_LAUNCH_XA(false, false); break;
    }
#undef A3D_LAUNCH_XA
    return check_launch("a3d_xattn_stack_v1");
}
