// ChainedDiffuser trajectory denoiser (diffusion_head.py:200-363, layers.py:7-290) as four kernels:
//
//   cd_ctx_lang    step-invariant: context tokens cross-attend to the instruction (+FFN), 2 layers
//   cd_step_begin  per step: trajectory MLP, trajectory->instruction attention, Q of layer 0
//   cd_cross       per layer: flash cross-attention of the 50 waypoints over the cached context
//                  K/V (one CTA per (head, sample); tensor-core QK^T / PV, TMA-fed tile ring)
//   cd_post        per layer: out-proj + LN, adaLN self-attention (rotary, padding mask) + LN,
//                  adaLN FFN + LN, optional regressor head, Q of the next layer, and on the last
//                  layer the DDPM posterior step (diffusion_model.py:100-117)
//
// Everything that does not depend on the trajectory or the timestep is hoisted out of the loop
// (SURVEY.md F6): context K/V for the 8 layers (a3d_ctx_kv), instruction K/V, adaLN tables.
#include "cd_blocks.cuh"

namespace a3d {
namespace cd {

// ------------------------------------------------------------------ packed weight layouts (floats)
// "lang" layer = ParallelAttentionLayer without adaLN / self-attention (vl_attention, traj_lang_attention)
struct LangPack {
    static constexpr int WQ = 0, BQ = WQ + E * EP, WO = BQ + EP, BO = WO + E * EP, G12 = BO + EP, B12 = G12 + EP;
    static constexpr int W1 = B12 + EP, B1 = W1 + E * FFP, W2 = B1 + FFP, B2 = W2 + FFP * EP, G122 = B2 + EP,
                         B122 = G122 + EP;
    static constexpr int SIZE = B122 + EP;
};
// "ada" layer = adaLN cross + self + FFN layer (traj_attention / pos_attention / rot_attention)
struct AdaPack {
    static constexpr int C_WQ = 0, C_BQ = C_WQ + E * EP, C_WO = C_BQ + EP, C_BO = C_WO + E * EP, G12 = C_BO + EP,
                         B12 = G12 + EP;
    static constexpr int S_WQ = B12 + EP, S_BQ = S_WQ + E * EP, S_WK = S_BQ + EP, S_BK = S_WK + E * EP,
                         S_WV = S_BK + EP, S_BV = S_WV + E * EP, S_WO = S_BV + EP, S_BO = S_WO + E * EP,
                         G1 = S_BO + EP, B1N = G1 + EP;
    static constexpr int W1 = B1N + EP, B1 = W1 + E * FFP, W2 = B1 + FFP, B2 = W2 + FFP * EP, G122 = B2 + EP,
                         B122 = G122 + EP;
    static constexpr int SIZE = B122 + EP;
};
// two-layer MLP head E -> E -> out (traj_encoder uses in = 9)
struct MlpPack {
    static constexpr int W1 = 0, B1 = W1 + E * EP, W2 = B1 + EP, B2 = W2 + E * EP;
    static constexpr int SIZE = B2 + EP;
};
constexpr int ADA_ROW = 3 * 2 * EP;   // per (timestep, layer): {adaln_12, adaln_1, adaln_ff1} x {scale, shift}

constexpr size_t SMEM_POST = (size_t)5 * TILE * 4 + 64 * 65 * 4 + 64 * 3 * 4 + 32 * 4 + 64;

struct Smem {
    float *xs, *t1, *t2, *t3, *t4, *scores, *xyz, *freq;
    unsigned char* mask;
    __device__ explicit Smem(unsigned char* base) {
        xs = reinterpret_cast<float*>(base);
        t1 = xs + TILE;
        t2 = t1 + TILE;
        t3 = t2 + TILE;
        t4 = t3 + TILE;
        scores = t4 + TILE;
        xyz = scores + 64 * 65;
        freq = xyz + 64 * 3;
        mask = reinterpret_cast<unsigned char*>(freq + 32);
    }
};

__device__ __forceinline__ void init_common(const Smem& s, const float* traj_b, int nrows, int traj_ld,
                                            const unsigned char* mask_b) {
    const int t = threadIdx.x;
    if (t < E / 6) s.freq[t] = rope_freq<E>(t);
    if (t < 64 * 3) {
        const int r = t / 3, a = t - 3 * r;
        s.xyz[t] = (traj_b && r < nrows) ? traj_b[r * traj_ld + a] : 0.f;
    }
    if (t < 64) s.mask[t] = (mask_b && t < nrows) ? mask_b[t] : 0;
}

// FFN with hidden 480 processed in four 128-wide chunks: out = W2 relu(W1 y + b1) + b2  -> dst tile
__device__ __forceinline__ void ffn_chunked(const Map& m, const float* y, const float* w1, const float* b1,
                                            const float* w2, const float* b2, float* hidden, float* dst) {
    float sum[2][16];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 16; ++c) sum[r][c] = 0.f;
    for (int ch = 0; ch < FFP / 128; ++ch) {
        linear_to_smem<E, FFP, true>(m, y, w1, b1, ch * 128, hidden);
        __syncthreads();
        linear_accumulate<128, EP>(m, hidden, w2 + (size_t)ch * 128 * EP, sum);
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float b = __ldg(b2 + 16 * m.cg + c);
        *reinterpret_cast<float2*>(dst + (16 * m.cg + c) * RP + 2 * m.rg) = make_float2(sum[0][c] + b, sum[1][c] + b);
    }
}

// Q (K-major fp32 tile, already rotated / scaled) -> global fp16 [H][64][16], pad slot zero
__device__ __forceinline__ void write_q_half(const float* q, __half* dst) {
    for (int i = threadIdx.x; i < H * ROWS * 16; i += blockDim.x) {
        const int d = i & 15, r = (i >> 4) & 63, h = i >> 10;
        dst[i] = __float2half_rn(d < HD ? q[(h * HD + d) * RP + r] : 0.f);
    }
}

// =============================================================== vision -> language (step-invariant)
struct CtxLangArgs {
    float* tok;            // [B][tok_rows][E], first nctx rows updated in place
    int tok_rows, nctx, n_instr, nlayers, batch;
    const float* kin;      // [nlayers][B][n_instr][E]  fp32 K projection of the instruction tokens
    const float* vin;      // same for V
    const float* w;        // [nlayers] LangPack
};

__global__ void __launch_bounds__(256, 1) cd_ctx_lang_kernel(const CtxLangArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem s(smem_raw);
    const Map m;
    const int b = blockIdx.y, r0 = blockIdx.x * ROWS;
    const int nrows = min(ROWS, a.nctx - r0);
    float* tok_b = a.tok + ((size_t)b * a.tok_rows + r0) * E;
    load_tile(s.xs, tok_b, nrows, E);
    __syncthreads();
    for (int l = 0; l < a.nlayers; ++l) {
        const float* w = a.w + (size_t)l * LangPack::SIZE;
        const size_t kvo = ((size_t)l * a.batch + b) * a.n_instr * E;
        linear_to_smem<E, EP, false>(m, s.xs, w + LangPack::WQ, w + LangPack::BQ, 0, s.t2);      // q (no pos, no adaLN)
        load_tile(s.t3, a.kin + kvo, a.n_instr, E);
        load_tile(s.t4, a.vin + kvo, a.n_instr, E);
        __syncthreads();
        small_mha(s.t2, s.t3, s.t4, a.n_instr, nullptr, s.scores, s.t2);
        __syncthreads();
        linear_to_smem<E, EP, false>(m, s.t2, w + LangPack::WO, w + LangPack::BO, 0, s.t1);
        __syncthreads();
        residual_layernorm(s.xs, s.t1, w + LangPack::G12, w + LangPack::B12);
        __syncthreads();
        ffn_chunked(m, s.xs, w + LangPack::W1, w + LangPack::B1, w + LangPack::W2, w + LangPack::B2, s.t2, s.t1);
        __syncthreads();
        residual_layernorm(s.xs, s.t1, w + LangPack::G122, w + LangPack::B122);
        __syncthreads();
    }
    store_tile(tok_b, s.xs, nrows, E);
}

// =============================================================== per-step kernels
struct StepArgs {
    int batch, nrows;               // nrows = trajectory length L (<= 64)
    const float* traj;              // [B][L][9] current noisy trajectory (normalised frame)
    const unsigned char* mask;      // [B][L] key padding mask or null
    const float* wp_pe;             // [L][E] sinusoidal waypoint-index embedding
    const int* t_idx;               // [B] timestep of each sample (row into the adaLN table)
    const float* ada;               // [T][nlayers_total][ADA_ROW]
    int ada_layers;                 // layers per timestep in the table
    // ---- begin kernel
    const float* traj_enc;          // MlpPack with W1 as [9][EP]
    const float* lang_w;            // LangPack (no FFN part used) or null
    const float* lang_k;            // [B][n_instr][E]
    const float* lang_v;
    int n_instr;
    // ---- post kernel
    const float* x_in;              // [B][64][E]
    const float* att;               // [B][64][E] cross-attention output (heads concatenated)
    const float* layer_w;           // AdaPack of this layer
    int ada_layer;                  // index of this layer in the adaLN table
    float* x_out;                   // [B][64][E]
    const float* reg_w;             // MlpPack or null
    float* reg_out;                 // [B][L][reg_dim]
    int reg_dim;
    // ---- next layer's Q
    const float* next_src;          // null: use the tile just computed; else [B][64][E]
    const float* next_wq;           // [E][EP] + bias [EP] contiguous (C_WQ, C_BQ of the next layer) or null
    int next_ada_layer;
    __half* q_out;                  // [B][H][64][16]
    // ---- DDPM update (last layer of a step)
    int do_update, last_step;
    float* traj_out;                // [B][L][9]
    const float* pos_upd;           // [B][L][3] output of the position regressor
    const float* cond_data;         // [B][L][9]
    const unsigned char* cond_mask; // [B][L][9]
    float coef[6];                  // pos: c_x0, c_xt, sigma ; rot: c_x0, c_xt, sigma
    const float* noise_pos;         // [B][L][3]
    const float* noise_rot;         // [B][L][6]
};

__device__ __forceinline__ void next_q(const Map& m, const Smem& s, const StepArgs& a, int b, const float* src_tile) {
    const float* ada = a.ada + ((size_t)a.t_idx[b] * a.ada_layers + a.next_ada_layer) * ADA_ROW;
    modulate(s.t1, src_tile, a.wp_pe, ada + 0, ada + EP, a.nrows);                    // adaln_12(x + wp_pe)
    __syncthreads();
    linear_rope_to_smem<E, EP>(m, s.t1, a.next_wq, a.next_wq + E * EP, s.xyz, s.freq, true, s.t2);
    __syncthreads();
    write_q_half(s.t2, a.q_out + (size_t)b * H * ROWS * 16);
}

__global__ void __launch_bounds__(256, 1) cd_step_begin_kernel(const StepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem s(smem_raw);
    const Map m;
    const int b = blockIdx.x;
    const float* traj_b = a.traj + (size_t)b * a.nrows * 9;
    init_common(s, traj_b, a.nrows, 9, nullptr);
    load_tile(s.t1, traj_b, a.nrows, 9, 9);
    __syncthreads();
    // trajectory encoder: Linear(9,E) -> ReLU -> Linear(E,E)   (diffusion_head.py:43-48, 215)
    linear_to_smem<9, EP, true>(m, s.t1, a.traj_enc + MlpPack::W1, a.traj_enc + MlpPack::B1, 0, s.t2);
    __syncthreads();
    linear_to_smem<E, EP, false>(m, s.t2, a.traj_enc + MlpPack::W2, a.traj_enc + MlpPack::B2, 0, s.xs);
    __syncthreads();
    if (a.lang_w) {   // trajectory tokens attend to the instruction (diffusion_head.py:330-336)
        const float* w = a.lang_w;
        modulate(s.t1, s.xs, a.wp_pe, nullptr, nullptr, a.nrows);
        load_tile(s.t3, a.lang_k + (size_t)b * a.n_instr * E, a.n_instr, E);
        load_tile(s.t4, a.lang_v + (size_t)b * a.n_instr * E, a.n_instr, E);
        __syncthreads();
        linear_to_smem<E, EP, false>(m, s.t1, w + LangPack::WQ, w + LangPack::BQ, 0, s.t2);
        __syncthreads();
        small_mha(s.t2, s.t3, s.t4, a.n_instr, nullptr, s.scores, s.t2);
        __syncthreads();
        linear_to_smem<E, EP, false>(m, s.t2, w + LangPack::WO, w + LangPack::BO, 0, s.t1);
        __syncthreads();
        residual_layernorm(s.xs, s.t1, w + LangPack::G12, w + LangPack::B12);
        __syncthreads();
    }
    store_tile(a.x_out + (size_t)b * ROWS * E, s.xs, ROWS, E);
    if (a.next_wq) next_q(m, s, a, b, s.xs);
}

__global__ void __launch_bounds__(256, 1) cd_post_kernel(const StepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem s(smem_raw);
    const Map m;
    const int b = blockIdx.x;
    const float* w = a.layer_w;
    const float* ada = a.ada + ((size_t)a.t_idx[b] * a.ada_layers + a.ada_layer) * ADA_ROW;
    const float* traj_b = a.traj + (size_t)b * a.nrows * 9;
    init_common(s, traj_b, a.nrows, 9, a.mask ? a.mask + (size_t)b * a.nrows : nullptr);
    load_tile(s.xs, a.x_in + (size_t)b * ROWS * E, ROWS, E);
    load_tile(s.t1, a.att + (size_t)b * ROWS * E, ROWS, E);
    __syncthreads();
    // ---- cross-attention epilogue: x = LN_12(x + att Wo^T + bo)          (layers.py:146-147)
    linear_to_smem<E, EP, false>(m, s.t1, w + AdaPack::C_WO, w + AdaPack::C_BO, 0, s.t2);
    __syncthreads();
    residual_layernorm(s.xs, s.t2, w + AdaPack::G12, w + AdaPack::B12);
    __syncthreads();
    // ---- self-attention: q = k = adaLN_1(x + pe), v = adaLN_1(x), rotary on q and k, padding mask (layers.py:165-182)
    modulate(s.t1, s.xs, a.wp_pe, ada + 2 * EP, ada + 3 * EP, a.nrows);
    modulate(s.t2, s.xs, nullptr, ada + 2 * EP, ada + 3 * EP, a.nrows);
    __syncthreads();
    linear_rope_to_smem<E, EP>(m, s.t1, w + AdaPack::S_WQ, w + AdaPack::S_BQ, s.xyz, s.freq, true, s.t3);
    linear_rope_to_smem<E, EP>(m, s.t1, w + AdaPack::S_WK, w + AdaPack::S_BK, s.xyz, s.freq, true, s.t4);
    __syncthreads();
    linear_to_smem<E, EP, false>(m, s.t2, w + AdaPack::S_WV, w + AdaPack::S_BV, 0, s.t1);
    __syncthreads();
    small_mha(s.t3, s.t4, s.t1, a.nrows, a.mask ? s.mask : nullptr, s.scores, s.t3);
    __syncthreads();
    linear_to_smem<E, EP, false>(m, s.t3, w + AdaPack::S_WO, w + AdaPack::S_BO, 0, s.t2);
    __syncthreads();
    residual_layernorm(s.xs, s.t2, w + AdaPack::G1, w + AdaPack::B1N);
    __syncthreads();
    // ---- FFN: y = adaLN_ff(x); x = LN_122(y + FFN(y))                      (layers.py:205-209)
    modulate(s.t1, s.xs, nullptr, ada + 4 * EP, ada + 5 * EP, a.nrows);
    __syncthreads();
    ffn_chunked(m, s.t1, w + AdaPack::W1, w + AdaPack::B1, w + AdaPack::W2, w + AdaPack::B2, s.t2, s.t3);
    __syncthreads();
    residual_layernorm(s.t1, s.t3, w + AdaPack::G122, w + AdaPack::B122);
    __syncthreads();
    float* x = s.t1;   // layer output
    store_tile(a.x_out + (size_t)b * ROWS * E, x, ROWS, E);

    // ---- regressor head: Linear(E,E) -> ReLU -> Linear(E,d)                (diffusion_head.py:179-198)
    if (a.reg_w) {
        linear_to_smem<E, EP, true>(m, x, a.reg_w + MlpPack::W1, a.reg_w + MlpPack::B1, 0, s.t2);
        __syncthreads();
        linear_to_smem<E, EP, false>(m, s.t2, a.reg_w + MlpPack::W2, a.reg_w + MlpPack::B2, 0, s.t3);
        __syncthreads();
        for (int i = threadIdx.x; i < a.nrows * a.reg_dim; i += blockDim.x) {
            const int r = i / a.reg_dim, d = i - r * a.reg_dim;
            a.reg_out[((size_t)b * a.nrows + r) * a.reg_dim + d] = s.t3[d * RP + r];
        }
    }
    // ---- Q of the next cross-attention layer
    if (a.next_wq) {
        const float* src = x;
        if (a.next_src) {
            __syncthreads();
            load_tile(s.xs, a.next_src + (size_t)b * ROWS * E, ROWS, E);
            __syncthreads();
            src = s.xs;
        }
        next_q(m, s, a, b, src);
    }
    // ---- denoiser output + DDPM posterior step (diffusion_head.py:271-274, diffusion_model.py:105-117)
    if (a.do_update) {
        __syncthreads();
        for (int i = threadIdx.x; i < a.nrows * 9; i += blockDim.x) {
            const int r = i / 9, d = i - 9 * r;
            const size_t gi = ((size_t)b * a.nrows + r) * 9 + d;
            const float cur = a.traj[gi];
            float out;
            if (d < 3) out = cur + a.pos_upd[((size_t)b * a.nrows + r) * 3 + d];      // position is residual
            else out = s.t3[(d - 3) * RP + r];                                          // rotation is direct
            if (a.cond_mask[gi]) out = a.cond_data[gi];
            float nxt = out;
            if (!a.last_step) {
                const float* cf = a.coef + (d < 3 ? 0 : 3);
                const float x0 = fminf(fmaxf(out, -1.f), 1.f);                          // clip_sample
                nxt = cf[0] * x0 + cf[1] * cur;
                const float nz = d < 3 ? a.noise_pos[((size_t)b * a.nrows + r) * 3 + d]
                                       : a.noise_rot[((size_t)b * a.nrows + r) * 6 + d - 3];
                nxt += cf[2] * nz;
            }
            a.traj_out[gi] = nxt;
        }
    }
}

// =============================================================== cross attention, one (head, sample) per CTA
struct CrossArgs {
    const __half* q;               // [B][H][64][16]
    const unsigned char* kv;       // tile images of this layer: [B][ntiles][2][H][64][16] fp16
    int nk, ntiles, batch;
    float* att;                    // [B][64][E]
};

__global__ void __launch_bounds__(128) cd_cross_kernel(const CrossArgs a) {
    constexpr int STAGES = 3;
    __shared__ __align__(128) unsigned char kvs[STAGES][4096];   // K image (2 KB) | V image (2 KB) of one head
    __shared__ __align__(16) __half qs[ROWS * 24];
    __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES];
    const int h = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q4 = lane & 3;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, 4);
        }
        mbar_fence_init();
    }
    const __half* qg = a.q + ((size_t)b * H + h) * ROWS * 16;
    for (int i = tid; i < ROWS * 16; i += 128) qs[(i >> 4) * 24 + (i & 15)] = qg[i];
    __syncthreads();
    const size_t tile_bytes = (size_t)2 * H * 2048;
    const unsigned char* kv_b = a.kv + (size_t)b * a.ntiles * tile_bytes;
    auto issue = [&](int t) {
        const int s = t % STAGES;
        if (t >= STAGES) mbar_wait(bar_empty + s, ((t / STAGES) - 1) & 1);
        mbar_expect_tx(bar_full + s, 4096);
        const unsigned char* src = kv_b + (size_t)t * tile_bytes;
        bulk_g2s(kvs[s], src + (size_t)h * 2048, 2048, bar_full + s);
        bulk_g2s(kvs[s] + 2048, src + (size_t)H * 2048 + (size_t)h * 2048, 2048, bar_full + s);
    };
    if (tid == 0)
        for (int t = 0; t < STAGES - 1 && t < a.ntiles; ++t) issue(t);

    uint32_t qf[4];
    {
        const int row = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        ldmatrix_x4(qf, smem_u32(qs + row * 24 + 8 * (lane >> 4)));
    }
    float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float m0 = -INFINITY, m1 = -INFINITY;
    const bool tail_mask = (a.nk % kTileKeys) != 0;
    for (int t = 0; t < a.ntiles; ++t) {
        const int stage = t % STAGES;
        if (tid == 0 && t + STAGES - 1 < a.ntiles) issue(t + STAGES - 1);
        mbar_wait(bar_full + stage, (t / STAGES) & 1);
        const uint32_t kbase = smem_u32(kvs[stage]), vbase = kbase + 2048;
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
            ldmatrix_x4(r, kbase + key * 32 + ((chunk ^ ((key >> 2) & 1)) << 4));
            mma_16816(s[2 * kk], qf, r[0], r[1]);
            mma_16816(s[2 * kk + 1], qf, r[2], r[3]);
        }
        if (tail_mask && t == a.ntiles - 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (t * kTileKeys + 8 * j + 2 * q4 + (e & 1) >= a.nk) s[j][e] = -INFINITY;
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
        const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
        const float al0 = exp2_fast(m0 - n0), al1 = exp2_fast(m1 - n1);
        m0 = n0;
        m1 = n1;
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            o[n][0] *= al0;
            o[n][1] *= al0;
            o[n][2] *= al1;
            o[n][3] *= al1;
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t pa[4];
            pa[0] = pack_h2(exp2_fast(s[2 * kk][0] - n0), exp2_fast(s[2 * kk][1] - n0));
            pa[1] = pack_h2(exp2_fast(s[2 * kk][2] - n1), exp2_fast(s[2 * kk][3] - n1));
            pa[2] = pack_h2(exp2_fast(s[2 * kk + 1][0] - n0), exp2_fast(s[2 * kk + 1][1] - n0));
            pa[3] = pack_h2(exp2_fast(s[2 * kk + 1][2] - n1), exp2_fast(s[2 * kk + 1][3] - n1));
            const int key = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
            const int chunk = lane >> 4;
            uint32_t r[4];
            ldmatrix_x4_trans(r, vbase + key * 32 + ((chunk ^ ((key >> 2) & 1)) << 4));
            mma_16816(o[0], pa, r[0], r[1]);
            mma_16816(o[1], pa, r[2], r[3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + stage);
    }
    const float l0 = __shfl_sync(0xffffffffu, o[1][1], (lane & ~3) | 3);
    const float l1 = __shfl_sync(0xffffffffu, o[1][3], (lane & ~3) | 3);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    float* out = a.att + (size_t)b * ROWS * E;
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int d = 8 * n + 2 * q4 + (e & 1);
            const int row = warp * 16 + g + 8 * (e >> 1);
            if (d < HD) out[(size_t)row * E + h * HD + d] = o[n][e] * ((e >> 1) ? i1 : i0);
        }
}

}  // namespace cd
}  // namespace a3d

// ===================================================================== C ABI
using namespace a3d;
using namespace a3d::cd;

extern "C" size_t cd_pack_floats(int which) {
    switch (which) {
        case 0: return LangPack::SIZE;
        case 1: return AdaPack::SIZE;
        case 2: return MlpPack::SIZE;
        case 3: return ADA_ROW;
        default: return 0;
    }
}

static int set_smem(const void* fn, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return A3D_ECUDA;
    }
    return A3D_OK;
}

extern "C" int cd_ctx_lang(float* tok, int batch, int tok_rows, int nctx, int embed, int heads, const float* kin,
                           const float* vin, int n_instr, const float* w, int nlayers, void* stream) {
    A3D_REQUIRE(tok && kin && vin && w, "cd_ctx_lang: null pointer");
    A3D_REQUIRE(embed == E && heads == H, "cd_ctx_lang: built for embedding_dim 120 / 8 heads (got %d / %d)", embed, heads);
    A3D_REQUIRE(batch > 0 && nctx > 0 && nctx <= tok_rows && n_instr > 0 && n_instr <= 64 && nlayers > 0,
                "cd_ctx_lang: bad sizes (nctx=%d rows=%d n_instr=%d)", nctx, tok_rows, n_instr);
    static bool once = false;
    if (!once) {
        if (int rc = set_smem((const void*)cd_ctx_lang_kernel, SMEM_POST)) return rc;
        once = true;
    }
    CtxLangArgs a{tok, tok_rows, nctx, n_instr, nlayers, batch, kin, vin, w};
    dim3 grid((nctx + ROWS - 1) / ROWS, batch);
    cd_ctx_lang_kernel<<<grid, 256, SMEM_POST, (cudaStream_t)stream>>>(a);
    return check_launch("cd_ctx_lang");
}

extern "C" int cd_step_begin(const float* traj, int batch, int length, const float* wp_pe, const int* t_idx,
                             const float* ada, int ada_layers, const float* traj_enc, const float* lang_w,
                             const float* lang_k, const float* lang_v, int n_instr, float* x_out,
                             const float* next_wq, int next_ada_layer, void* q_out, void* stream) {
    A3D_REQUIRE(traj && wp_pe && t_idx && ada && traj_enc && x_out, "cd_step_begin: null pointer");
    A3D_REQUIRE(batch > 0 && length > 0 && length <= ROWS, "cd_step_begin: trajectory length %d not in [1,64]", length);
    A3D_REQUIRE(!lang_w || (lang_k && lang_v && n_instr > 0 && n_instr <= 64), "cd_step_begin: instruction K/V missing");
    A3D_REQUIRE(!next_wq || q_out, "cd_step_begin: q_out missing");
    static bool once = false;
    if (!once) {
        if (int rc = set_smem((const void*)cd_step_begin_kernel, SMEM_POST)) return rc;
        once = true;
    }
    StepArgs a{};
    a.batch = batch;
    a.nrows = length;
    a.traj = traj;
    a.wp_pe = wp_pe;
    a.t_idx = t_idx;
    a.ada = ada;
    a.ada_layers = ada_layers;
    a.traj_enc = traj_enc;
    a.lang_w = lang_w;
    a.lang_k = lang_k;
    a.lang_v = lang_v;
    a.n_instr = n_instr;
    a.x_out = x_out;
    a.next_wq = next_wq;
    a.next_ada_layer = next_ada_layer;
    a.q_out = (__half*)q_out;
    cd_step_begin_kernel<<<batch, 256, SMEM_POST, (cudaStream_t)stream>>>(a);
    return check_launch("cd_step_begin");
}

extern "C" int cd_cross(const void* q, const void* kv, int batch, int nk, int heads, float* att, void* stream) {
    A3D_REQUIRE(q && kv && att, "cd_cross: null pointer");
    A3D_REQUIRE(heads == H && batch > 0 && nk > 0, "cd_cross: built for 8 heads (got %d)", heads);
    A3D_REQUIRE(((uintptr_t)kv & 15) == 0, "cd_cross: K/V cache must be 16-byte aligned");
    CrossArgs a{(const __half*)q, (const unsigned char*)kv, nk, (nk + kTileKeys - 1) / kTileKeys, batch, att};
    dim3 grid(H, batch);
    cd_cross_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    return check_launch("cd_cross");
}

extern "C" int cd_post(const float* traj, int batch, int length, const unsigned char* mask, const float* wp_pe,
                       const int* t_idx, const float* ada, int ada_layers, int ada_layer, const float* x_in,
                       const float* att, const float* layer_w, float* x_out, const float* reg_w, float* reg_out,
                       int reg_dim, const float* next_src, const float* next_wq, int next_ada_layer, void* q_out,
                       int do_update, int last_step, float* traj_out, const float* pos_upd, const float* cond_data,
                       const unsigned char* cond_mask, const float* coef_host, const float* noise_pos,
                       const float* noise_rot, void* stream) {
    A3D_REQUIRE(traj && wp_pe && t_idx && ada && x_in && att && layer_w && x_out, "cd_post: null pointer");
    A3D_REQUIRE(batch > 0 && length > 0 && length <= ROWS, "cd_post: trajectory length %d not in [1,64]", length);
    A3D_REQUIRE(!reg_w || (reg_out && reg_dim > 0 && reg_dim <= 16), "cd_post: regressor output missing");
    A3D_REQUIRE(!next_wq || q_out, "cd_post: q_out missing");
    A3D_REQUIRE(!do_update || (traj_out && pos_upd && cond_data && cond_mask && coef_host && reg_w && reg_dim == 6 &&
                               (last_step || (noise_pos && noise_rot))),
                "cd_post: DDPM update needs traj_out, pos_upd, cond_*, coef, the rotation regressor and noise");
    static bool once = false;
    if (!once) {
        if (int rc = set_smem((const void*)cd_post_kernel, SMEM_POST)) return rc;
        once = true;
    }
    StepArgs a{};
    a.batch = batch;
    a.nrows = length;
    a.traj = traj;
    a.mask = mask;
    a.wp_pe = wp_pe;
    a.t_idx = t_idx;
    a.ada = ada;
    a.ada_layers = ada_layers;
    a.ada_layer = ada_layer;
    a.x_in = x_in;
    a.att = att;
    a.layer_w = layer_w;
    a.x_out = x_out;
    a.reg_w = reg_w;
    a.reg_out = reg_out;
    a.reg_dim = reg_dim;
    a.next_src = next_src;
    a.next_wq = next_wq;
    a.next_ada_layer = next_ada_layer;
    a.q_out = (__half*)q_out;
    a.do_update = do_update;
    a.last_step = last_step;
    a.traj_out = traj_out;
    a.pos_upd = pos_upd;
    a.cond_data = cond_data;
    a.cond_mask = cond_mask;
    for (int i = 0; i < 6; ++i) a.coef[i] = coef_host ? coef_host[i] : 0.f;
    a.noise_pos = noise_pos;
    a.noise_rot = noise_rot;
    cd_post_kernel<<<batch, 256, SMEM_POST, (cudaStream_t)stream>>>(a);
    return check_launch("cd_post");
}
