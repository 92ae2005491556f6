// Fused cross-attention stack, second generation (Act3D ghost-point / query / vision-language stacks).
//
// One CTA = 128 query rows of one sample for ALL layers of the stack; one WARP = 16 rows, and a warp
// never exchanges data with another warp: its slice of the residual stream, of Q and of every
// intermediate activation lives in its own registers (MMA accumulator layout) or in its own rows of
// shared memory.  The only CTA-wide objects are the K/V tile ring (3 stages, 1-D bulk async copies
// completing on mbarriers) and the weights (read-only, L1/L2).  No __syncthreads after the prologue.
//
// Per layer and warp:
//   q      = rotary((x Wq^T + bq) * hd^-1/2 * log2 e)                 error-compensated fp16-split MMA
//   flash loop over 64-key tiles, 4 heads:
//       S' = Q K^T - m_stale       the stale row max enters as the accumulator's initial value
//       P  = 2^S'                  (fp32 MUFU, packed to fp16); the running max is only refreshed when
//                                  a score exceeds the stale max by more than 2^8 (conditional rescale)
//       O += P V                   slot 15 of V is 1: the softmax denominator rides in the same MMA
//   x = LN(x + (O / l) Wo^T + bo);  x = LN(x + W2 relu(W1 x + b1) + b2)  split MMAs chained in registers
// After the last layer: mask logits <qvec, x> (act3d.py:493-494) and/or the features.
#include <string.h>

#include "a3d_xattn_common.cuh"

namespace a3d {

// PM: bit i set -> the i-th of the 8 scores a thread exponentiates per 16-key step uses exp2_poly
template <int PM>
__global__ void __launch_bounds__(256, 2) xattn2_kernel(const Xa2Args a) {
    using C = Xa2;
    constexpr int E = C::E, H = C::H;
    extern __shared__ __align__(128) unsigned char smem[];
    float* xpark = reinterpret_cast<float*>(smem);                               // [128][XP] residual stream (parked)
    __half* qs = reinterpret_cast<__half*>(smem + C::X_BYTES);                   // [128][QP] fp16 Q, head-padded
    unsigned char* kvs = smem + C::X_BYTES + C::Q_BYTES;                         // STAGES x TILE_BYTES
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(kvs + C::STAGES * C::TILE_BYTES);
    uint64_t* bar_empty = bar_full + C::STAGES;
    __shared__ float freq[E / 6];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, row0 = blockIdx.x * C::ROWS;
    const int g = lane >> 2, q4 = lane & 3;
    const int lrow0 = warp * 16 + g, lrow1 = lrow0 + 8;       // this thread's two rows inside the CTA tile

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, 8);
        }
        mbar_fence_init();
    }
    if (tid < E / 6) freq[tid] = rope_freq<E>(tid);
    __syncthreads();

    // ---- residual stream in accumulator layout: xr[n][2r + j] = x[row_r][8n + 2 q4 + j]
    float xr[8][4];
    float qxyz[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = row0 + (r ? lrow1 : lrow0);
        const bool ok = row < a.nq;
        const float* xp = a.x0 + (long)b * a.x0_sb + (long)row * a.x0_sn;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int c = 8 * n + 2 * q4;
            float2 v = make_float2(0.f, 0.f);
            if (ok && c < E) v = __ldg(reinterpret_cast<const float2*>(xp + c));
            xr[n][2 * r] = v.x;
            xr[n][2 * r + 1] = v.y;
        }
        if (a.qpos && ok) {
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) qxyz[r][ax] = __ldg(a.qpos + ((long)b * a.nq + row) * 3 + ax);
        }
    }

    const unsigned char* kv_sample = a.kv_base + (size_t)b * a.ntiles * C::TILE_BYTES;
    const bool tail_mask = (a.nk % kTileKeys) != 0;
    uint32_t gtile = 0;   // global tile counter over all layers (drives stage / parity)

    for (int layer = 0; layer < a.nlayers; ++layer) {
        const uint4* w = a.w + (size_t)layer * C::LAYER_W;
        const float* vv = a.v + (size_t)layer * C::LAYER_V;
        const unsigned char* kv_layer = kv_sample + (size_t)layer * a.kv_layer_stride;

        auto issue_tile = [&](int t) {
            const uint32_t gt = gtile + t, s = gt % C::STAGES;
            if (gt >= (uint32_t)C::STAGES) mbar_wait(bar_empty + s, ((gt / C::STAGES) - 1) & 1);
            mbar_expect_tx(bar_full + s, C::TILE_BYTES);
            bulk_g2s(kvs + s * C::TILE_BYTES, kv_layer + (size_t)t * C::TILE_BYTES, C::TILE_BYTES, bar_full + s);
        };
        if (tid == 0)
            for (int t = 0; t < C::STAGES - 1 && t < a.ntiles; ++t) issue_tile(t);

        // ---- park x (residual for after the attention), then q = rotary(x Wq^T + bq) -> fp16 head-padded
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            *reinterpret_cast<float2*>(xpark + lrow0 * C::XP + 8 * n + 2 * q4) = make_float2(xr[n][0], xr[n][1]);
            *reinterpret_cast<float2*>(xpark + lrow1 * C::XP + 8 * n + 2 * q4) = make_float2(xr[n][2], xr[n][3]);
        }
        {
            float qa[8][4];
            gemm_reg(xr, w + C::W_Q, lane, qa);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int c = 8 * n + 2 * q4;
                const float b0 = __ldg(vv + C::B_Q + c), b1 = __ldg(vv + C::B_Q + c + 1);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    float v0 = qa[n][2 * r] + b0, v1 = qa[n][2 * r + 1] + b1;
                    if (a.qpos && c < E) {
                        const int pi = c >> 1, axis = pi / (E / 6), j = pi - axis * (E / 6);
                        const float ang = qxyz[r][axis] * freq[j];
                        float sv, cv;
                        if (fabsf(ang) < 3.0f) {
                            __sincosf(ang, &sv, &cv);
                        } else {
                            sincosf(ang, &sv, &cv);
                        }
                        const float ev = v0, od = v1;
                        v0 = ev * cv - od * sv;
                        v1 = od * cv + ev * sv;
                    }
                    __half* qrow = qs + (r ? lrow1 : lrow0) * C::QP;
                    if (c < E) {
                        const int h0 = c / 15, h1 = (c + 1) / 15;
                        qrow[c + h0] = __float2half_rn(v0);             // slot = e + e/15
                        qrow[c + 1 + h1] = __float2half_rn(v1);
                    } else {                                            // dims 60..63 own the pad slots of heads 0..3
                        qrow[(c - E) * 16 + 15] = __float2half_rn(0.f);
                        qrow[(c + 1 - E) * 16 + 15] = __float2half_rn(0.f);
                    }
                }
            }
        }
        __syncwarp();

        // ---------------------------------------------------------------- attention core
        uint32_t qf[H][4];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const int row = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
            ldmatrix_x4(qf[h], smem_u32(qs + row * C::QP + h * 16 + 8 * (lane >> 4)));
        }
        float o[H][2][4];
        float negm[H][2];   // minus the (stale) running max of rows g / g+8, per head
#pragma unroll
        for (int h = 0; h < H; ++h) {
            negm[h][0] = 0.f;
            negm[h][1] = 0.f;
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) o[h][n][e] = 0.f;
        }

        for (int t = 0; t < a.ntiles; ++t) {
            const uint32_t gt = gtile + t, stage = gt % C::STAGES;
            if (tid == 0 && t + C::STAGES - 1 < a.ntiles) issue_tile(t + C::STAGES - 1);
            mbar_wait(bar_full + stage, (gt / C::STAGES) & 1);

            const uint32_t kbase = smem_u32(kvs + stage * C::TILE_BYTES);
            const uint32_t vbase = kbase + H * 2048;
            const bool mask_this = tail_mask && (t == a.ntiles - 1);
            const bool first = (t == 0);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float cinit[4] = {negm[h][0], negm[h][0], negm[h][1], negm[h][1]};
                float s[8][4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int key = kk * 16 + (lane & 7) + 8 * (lane >> 4);
                    const int chunk = (lane >> 3) & 1;
                    uint32_t r[4];
                    ldmatrix_x4(r, kbase + h * 2048 + key * 32 + ((chunk ^ ((key >> 2) & 1)) << 4));
                    mma_16816_c(s[2 * kk], qf[h], r[0], r[1], cinit);
                    mma_16816_c(s[2 * kk + 1], qf[h], r[2], r[3], cinit);
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
                // conditional rescale: scores are already relative to the stale max; refresh it only when
                // some row of this warp overshoots it by more than 2^8 (always on the first tile)
                if (first || __any_sync(0xffffffffu, fmaxf(mx0, mx1) > 8.0f)) {
                    const float d0 = first ? mx0 : fmaxf(mx0, 0.f), d1 = first ? mx1 : fmaxf(mx1, 0.f);
                    const float al0 = exp2_fast(-d0), al1 = exp2_fast(-d1);
                    negm[h][0] -= d0;
                    negm[h][1] -= d1;
#pragma unroll
                    for (int n = 0; n < 2; ++n) {
                        o[h][n][0] *= al0;
                        o[h][n][1] *= al0;
                        o[h][n][2] *= al1;
                        o[h][n][3] *= al1;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        s[j][0] -= d0;
                        s[j][1] -= d0;
                        s[j][2] -= d1;
                        s[j][3] -= d1;
                    }
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    uint32_t pa[4];
                    pa[0] = pack_h2(exp2_sel<PM, 0>(s[2 * kk][0]), exp2_sel<PM, 1>(s[2 * kk][1]));
                    pa[1] = pack_h2(exp2_sel<PM, 2>(s[2 * kk][2]), exp2_sel<PM, 3>(s[2 * kk][3]));
                    pa[2] = pack_h2(exp2_sel<PM, 4>(s[2 * kk + 1][0]), exp2_sel<PM, 5>(s[2 * kk + 1][1]));
                    pa[3] = pack_h2(exp2_sel<PM, 6>(s[2 * kk + 1][2]), exp2_sel<PM, 7>(s[2 * kk + 1][3]));
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

        // normalise by the denominator carried in slot 15
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float l0 = __shfl_sync(0xffffffffu, o[h][1][1], (lane & ~3) | 3);
            const float l1 = __shfl_sync(0xffffffffu, o[h][1][3], (lane & ~3) | 3);
            const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                o[h][n][0] *= i0;
                o[h][n][1] *= i0;
                o[h][n][2] *= i1;
                o[h][n][3] *= i1;
            }
        }

        // ---------------------------------------------------------------- out-proj + residual + LN
        {
            float y[8][4];
            gemm_reg(reinterpret_cast<const float(&)[8][4]>(o), w + C::W_O, lane, y);   // k = 16 h + d (pad rows of Wo are 0)
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int c = 8 * n + 2 * q4;
                const float b0 = __ldg(vv + C::B_O + c), b1 = __ldg(vv + C::B_O + c + 1);
                const float2 r0 = *reinterpret_cast<const float2*>(xpark + lrow0 * C::XP + c);
                const float2 r1 = *reinterpret_cast<const float2*>(xpark + lrow1 * C::XP + c);
                xr[n][0] = y[n][0] + b0 + r0.x;
                xr[n][1] = y[n][1] + b1 + r0.y;
                xr[n][2] = y[n][2] + b0 + r1.x;
                xr[n][3] = y[n][3] + b1 + r1.y;
            }
            layernorm_frag(xr, q4, vv + C::G_1, vv + C::BE_1);
        }
        // ---------------------------------------------------------------- FFN + residual + LN
        {
#pragma unroll
            for (int n = 0; n < 8; ++n) {   // park x1: it is the residual of the FFN
                *reinterpret_cast<float2*>(xpark + lrow0 * C::XP + 8 * n + 2 * q4) = make_float2(xr[n][0], xr[n][1]);
                *reinterpret_cast<float2*>(xpark + lrow1 * C::XP + 8 * n + 2 * q4) = make_float2(xr[n][2], xr[n][3]);
            }
            float hid[8][4];
            gemm_reg(xr, w + C::W_1, lane, hid);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int c = 8 * n + 2 * q4;
                const float b0 = __ldg(vv + C::B_1 + c), b1 = __ldg(vv + C::B_1 + c + 1);
                hid[n][0] = fmaxf(hid[n][0] + b0, 0.f);
                hid[n][1] = fmaxf(hid[n][1] + b1, 0.f);
                hid[n][2] = fmaxf(hid[n][2] + b0, 0.f);
                hid[n][3] = fmaxf(hid[n][3] + b1, 0.f);
            }
            gemm_reg(hid, w + C::W_2, lane, xr);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int c = 8 * n + 2 * q4;
                const float b0 = __ldg(vv + C::B_2 + c), b1 = __ldg(vv + C::B_2 + c + 1);
                const float2 r0 = *reinterpret_cast<const float2*>(xpark + lrow0 * C::XP + c);
                const float2 r1 = *reinterpret_cast<const float2*>(xpark + lrow1 * C::XP + c);
                xr[n][0] += b0 + r0.x;
                xr[n][1] += b1 + r0.y;
                xr[n][2] += b0 + r1.x;
                xr[n][3] += b1 + r1.y;
            }
            layernorm_frag(xr, q4, vv + C::G_2, vv + C::BE_2);
        }

        // ---------------------------------------------------------------- outputs of this layer
        const bool last = (layer == a.nlayers - 1);
        if (a.feat_out && (a.feat_all || last)) {
            float* fo = a.feat_out + ((size_t)(a.feat_all ? layer : 0) * a.batch + b) * (size_t)a.feat_rows * E;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = row0 + (r ? lrow1 : lrow0);
                if (row < a.nq) {
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        const int c = 8 * n + 2 * q4;
                        if (c < E) *reinterpret_cast<float2*>(fo + (size_t)row * E + c) = make_float2(xr[n][2 * r], xr[n][2 * r + 1]);
                    }
                }
            }
        }
        if (last && a.logits) {
            for (int j = 0; j < a.nqv; ++j) {
                const float* qv = a.qvec + ((size_t)j * a.batch + b) * E;
                float p0 = 0.f, p1 = 0.f;
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int c = 8 * n + 2 * q4;
                    if (c < E) {
                        const float2 qc = __ldg(reinterpret_cast<const float2*>(qv + c));
                        p0 = fmaf(xr[n][0], qc.x, p0);
                        p0 = fmaf(xr[n][1], qc.y, p0);
                        p1 = fmaf(xr[n][2], qc.x, p1);
                        p1 = fmaf(xr[n][3], qc.y, p1);
                    }
                }
                p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
                p0 += __shfl_xor_sync(0xffffffffu, p0, 2);
                p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
                p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
                if (q4 == 0) {
                    float* lo = a.logits + ((size_t)j * a.batch + b) * a.nq;
                    if (row0 + lrow0 < a.nq) lo[row0 + lrow0] = p0;
                    if (row0 + lrow1 < a.nq) lo[row0 + lrow1] = p1;
                }
            }
        }
        __syncwarp();   // this warp's rows of qs / xpark are rewritten by the next layer
    }
}

}  // namespace a3d

using namespace a3d;

// 0 (default): pick per launch -- the tcgen05 / TMEM core (a3d_xattn6.cu) for launches that fill the GPU,
// the mma.sync core (this file) for the small ones (the 1-token query stack: 16 CTAs, runs on a side stream next to a
// ghost-point launch whose CTAs own all of an SM's tensor memory); 2 / 4 / 6 force mma.sync / the round-1 tcgen05 kernel (A/B reference) / a3d_xattn6.cu
int g_xattn_core = 0;
int a3d_launch_xattn4(const Xa2Args& a, dim3 grid, cudaStream_t stream, int poly);
int a3d_launch_xattn6(const Xa2Args& a, dim3 grid, cudaStream_t stream, int poly);
int g_xattn6_np = 6;    // a3d_set_option("xattn6_np", n): score pairs (of 16 per thread and unit) on the FMA-pipe polynomial
namespace a3d { extern int g_train_attn_core; extern int g_cd_prefetch_tiles; }   // a3d_train.cu, cd_loop.cu
int g_xattn_poly = 0;   // set through a3d_set_option("xattn_poly", 0|2|3|4); measured: 0 is fastest (issue-bound)

extern "C" int a3d_set_option(const char* name, int value) {
    if (name && strcmp(name, "xattn_core") == 0) {
        A3D_REQUIRE(value == 0 || value == 2 || value == 4 || value == 6,
                    "a3d_set_option: xattn_core must be 0 (auto), 2 (mma.sync), 4 (tcgen05 attention, round-1 kernel) "
                    "or 6 (tcgen05 attention + linear layers, FMA-pipe exponentials)");
        g_xattn_core = value;
        return A3D_OK;
    }
    if (name && strcmp(name, "xattn6_np") == 0) {
        A3D_REQUIRE(value == 0 || value == 4 || (value >= 6 && value <= 10), "a3d_set_option: xattn6_np must be 0, 4 or 6..10");
        g_xattn6_np = value;
        return A3D_OK;
    }
    if (name && strcmp(name, "xattn_poly") == 0) {
        A3D_REQUIRE(value == 0 || value == 2 || value == 3 || value == 4, "a3d_set_option: xattn_poly must be 0, 2, 3 or 4");
        g_xattn_poly = value;
        return A3D_OK;
    }
    if (name && strcmp(name, "train_attn_core") == 0) {
        A3D_REQUIRE(value == 0 || value == 1, "a3d_set_option: train_attn_core must be 0 (tensor cores) or 1 (fp32 CUDA cores)");
        g_train_attn_core = value;
        return A3D_OK;
    }
    if (name && strcmp(name, "cd_prefetch_tiles") == 0) {
        A3D_REQUIRE(value >= 0 && value <= 64, "a3d_set_option: cd_prefetch_tiles must be in [0, 64]");
        g_cd_prefetch_tiles = value;
        return A3D_OK;
    }
    A3D_REQUIRE(false, "a3d_set_option: unknown option '%s'", name ? name : "(null)");
}

int a3d_xattn4_replays(unsigned long long* value, int reset);
int a3d_xattn6_replays(unsigned long long* value, int reset);

extern "C" int a3d_debug_counter(const char* name, int reset, unsigned long long* value_host) {
    A3D_REQUIRE(name && value_host, "a3d_debug_counter: null pointer");
    if (strcmp(name, "xattn_replays") == 0) {
        unsigned long long v4 = 0, v6 = 0;
        if (a3d_xattn4_replays(&v4, reset) != A3D_OK || a3d_xattn6_replays(&v6, reset) != A3D_OK) {
            set_error("a3d_debug_counter: %s", cudaGetErrorString(cudaGetLastError()));
            return A3D_ECUDA;
        }
        *value_host = v4 + v6;
        return A3D_OK;
    }
    A3D_REQUIRE(false, "a3d_debug_counter: unknown counter '%s'", name);
}

extern "C" size_t a3d_xattn_layer_words(int embed, int ffn) {
    return (embed == 60 && ffn == 60) ? (size_t)Xa2::LAYER_W * 4 : 0;   // 32-bit words of fragment weights per layer
}
extern "C" size_t a3d_xattn_layer_floats(int embed, int ffn) {
    return (embed == 60 && ffn == 60) ? (size_t)Xa2::LAYER_V : 0;       // fp32 vector floats per layer
}

extern "C" int a3d_xattn_stack(const float* x0, long x0_stride_b, long x0_stride_n, const float* qpos, int batch,
                               int nq, int nk, int embed, int heads, int ffn, int nlayers, const void* kv_base,
                               size_t kv_layer_stride_bytes, const void* w, const float* v, float* feat_out,
                               int feat_rows, int feat_all_layers, const float* qvec, int nqv, float* logits,
                               void* stream) {
    A3D_REQUIRE(x0 && kv_base && w && v, "a3d_xattn_stack: null pointer");
    A3D_REQUIRE(batch > 0 && nq > 0 && nk > 0 && nlayers > 0, "a3d_xattn_stack: empty problem (B=%d nq=%d nk=%d L=%d)", batch, nq, nk, nlayers);
    A3D_REQUIRE(embed == 60 && heads == 4 && ffn == 60, "a3d_xattn_stack: (embed, heads, ffn) = (%d,%d,%d) not supported; built for (60,4,60)", embed, heads, ffn);
    A3D_REQUIRE(!feat_out || feat_rows >= nq, "a3d_xattn_stack: feat_rows %d < nq %d", feat_rows, nq);
    A3D_REQUIRE((logits == nullptr) == (qvec == nullptr || nqv == 0), "a3d_xattn_stack: qvec/logits must be given together");
    A3D_REQUIRE(((uintptr_t)kv_base & 15) == 0 && (kv_layer_stride_bytes & 15) == 0, "a3d_xattn_stack: K/V cache must be 16-byte aligned");
    A3D_REQUIRE(((uintptr_t)w & 15) == 0, "a3d_xattn_stack: weights must be 16-byte aligned");
    A3D_REQUIRE((x0_stride_b % 2 == 0) && (x0_stride_n % 2 == 0) && (((uintptr_t)x0 & 7) == 0), "a3d_xattn_stack: x0 must be 8-byte aligned with even strides");
    A3D_REQUIRE(batch <= 65535, "a3d_xattn_stack: batch %d exceeds grid.y", batch);
    Xa2Args a;
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
    a.w = (const uint4*)w;
    a.v = v;
    a.feat_out = feat_out;
    a.feat_rows = feat_rows;
    a.feat_all = feat_all_layers;
    a.qvec = qvec;
    a.nqv = nqv;
    a.logits = logits;
    dim3 grid((nq + Xa2::ROWS - 1) / Xa2::ROWS, batch);
    const int core = g_xattn_core ? g_xattn_core : ((long)grid.x * grid.y >= 296 ? 6 : 2);
    if (core == 4) return a3d_launch_xattn4(a, grid, (cudaStream_t)stream, g_xattn_poly);
    if (core == 6) return a3d_launch_xattn6(a, grid, (cudaStream_t)stream, g_xattn6_np);
#define A3D_XA2(PM)                                                                                                   \
    do {                                                                                                               \
        static PerDeviceOnce once_dev;                                                                                      \
        if (bool& once = once_dev.flag(); !once) {                                                                                                   \
            cudaError_t e = cudaFuncSetAttribute(xattn2_kernel<PM>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                                 (int)Xa2::SMEM);                                                      \
            if (e != cudaSuccess) {                                                                                    \
                set_error("a3d_xattn_stack: cudaFuncSetAttribute: %s", cudaGetErrorString(e));                       \
                return A3D_ECUDA;                                                                                      \
            }                                                                                                          \
            once = true;                                                                                               \
        }                                                                                                              \
        xattn2_kernel<PM><<<grid, 256, Xa2::SMEM, (cudaStream_t)stream>>>(a);                                          \
    } while (0)
    switch (g_xattn_poly) {   // fraction of exponentials evaluated on the FMA pipe: 0, 2/8, 3/8, 4/8
        case 2: A3D_XA2(0x22); break;
        case 3: A3D_XA2(0x2A); break;
        case 4: A3D_XA2(0xAA); break;
        default: A3D_XA2(0x00); break;
    }
#undef A3D_XA2
    return check_launch("a3d_xattn_stack");
}
