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
//
// All E x E / E x 4E linear layers run on the tensor cores through the error-compensated fp16-split
// GEMM of a3d_mma_gemm.cuh (fp32-class accuracy); one CTA owns the 64-row token tile of a sample:
// residual stream fp32 K-major in shared memory, GEMM inputs as fp16 (hi, lo) row-major planes.
#include "a3d_mma_gemm.cuh"
#include "cd_blocks.cuh"

namespace a3d {
namespace cd {

constexpr int kCrossSplits = 4;            // key tiles of one (sample, head) are split over this many cd_cross CTAs
constexpr int THREADS = 512;               // 16 warps: 4 m tiles x 4 n parts
constexpr int LPR = THREADS / ROWS;        // lanes per row in the row-wise passes
constexpr int NTW = 4;                     // n tiles per warp in the CTA GEMM
constexpr int PITCH = 136;                 // halfs per row of an A plane (272 B: conflict-free ldmatrix)
constexpr int PLANE = ROWS * PITCH;        // halfs per plane
// ------------------------------------------------------------------ shared memory map
constexpr int TILE_E = E * RP;             // floats of a K-major activation tile (only the E real channels are stored)
constexpr size_t SMEM_BYTES = (size_t)4 * TILE_E * 4 + (size_t)4 * PLANE * 2 + (size_t)kRingStages * kSlabBytes +
                              64 * 3 * 4 + 32 * 4 + 64 + 16;

struct Smem {
    float *xs, *t1, *t2, *t3, *xyz, *freq;
    __half *ah, *al, *hh, *hl;
    __half *kh, *kl, *vh, *vl;       // head-padded K / V planes of the in-CTA attention (alias t1..t3)
    unsigned char* mask;
    WeightRing ring;
    __device__ explicit Smem(unsigned char* base) {
        ring.buf = base;                                // 16-byte aligned weight ring first
        ring.head_of = nullptr;
        xs = reinterpret_cast<float*>(base + kRingStages * kSlabBytes);
        t1 = xs + TILE_E;
        t2 = t1 + TILE_E;
        t3 = t2 + TILE_E;
        ah = reinterpret_cast<__half*>(t3 + TILE_E);
        al = ah + PLANE;
        hh = al + PLANE;
        hl = hh + PLANE;
        kh = reinterpret_cast<__half*>(t1);             // 4 planes = 69632 B inside the 95040 B of t1..t3
        kl = kh + PLANE;
        vh = kl + PLANE;
        vl = vh + PLANE;
        xyz = reinterpret_cast<float*>(hl + PLANE);
        freq = xyz + 64 * 3;
        mask = reinterpret_cast<unsigned char*>(freq + 32);
    }
};

struct WarpMap {
    int tid, lane, warp, m0, nh, g, q4;
    __device__ WarpMap() {
        tid = threadIdx.x;
        lane = tid & 31;
        warp = tid >> 5;
        m0 = 16 * (warp & 3);
        nh = warp >> 2;                         // n part 0..3 (4 n tiles each)
        g = lane >> 2;
        q4 = lane & 3;
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

// ---- producers of GEMM input planes ----------------------------------------------------------
// planes[r][c] = split( (src[c][r] + pe[r][c]) * (1 + scale[c]) + shift[c] ), optional fp32 copy (K-major)
__device__ __forceinline__ void planes_from_tile(const float* __restrict__ src, const float* __restrict__ pe,
                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                 int nrows, __half* __restrict__ hi, __half* __restrict__ lo,
                                                 float* __restrict__ f32_out) {
#pragma unroll 4
    for (int i = threadIdx.x; i < ROWS * (EP / 2); i += blockDim.x) {
        const int r = i >> 6, c = (i & 63) * 2;
        float v0 = 0.f, v1 = 0.f;
        if (c < E) {
            v0 = src[c * RP + r];
            v1 = src[(c + 1) * RP + r];
            if (pe && r < nrows) {
                v0 += __ldg(pe + r * E + c);
                v1 += __ldg(pe + r * E + c + 1);
            }
            if (scale) {
                v0 = v0 * (1.0f + __ldg(scale + c)) + __ldg(shift + c);
                v1 = v1 * (1.0f + __ldg(scale + c + 1)) + __ldg(shift + c + 1);
            }
            if (f32_out) {
                f32_out[c * RP + r] = v0;
                f32_out[(c + 1) * RP + r] = v1;
            }
        }
        uint32_t h, l;
        split_h2(v0, v1, h, l);
        *reinterpret_cast<uint32_t*>(hi + r * PITCH + c) = h;
        *reinterpret_cast<uint32_t*>(lo + r * PITCH + c) = l;
    }
}
// planes of the cross-attention output: merge the kCrossSplits unnormalised partials of every (row, head)
//   att[r][15h + d] = sum_s O_s[d] 2^(m_s - M) / sum_s l_s 2^(m_s - M),  M = max_s m_s
__device__ __forceinline__ void planes_from_partials(const float* __restrict__ part_b, __half* __restrict__ hi,
                                                     __half* __restrict__ lo) {
    for (int i = threadIdx.x; i < ROWS * H; i += blockDim.x) {
        const int r = i & 63, h = i >> 6;
        float m[kCrossSplits], big = -INFINITY;
#pragma unroll
        for (int sp = 0; sp < kCrossSplits; ++sp) {
            m[sp] = __ldg(part_b + ((size_t)sp * H + h) * ROWS * 17 + 16 * ROWS + r);
            big = fmaxf(big, m[sp]);
        }
        float acc[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[d] = 0.f;
#pragma unroll
        for (int sp = 0; sp < kCrossSplits; ++sp) {
            const float wgt = (m[sp] == -INFINITY) ? 0.f : exp2f(m[sp] - big);
            const float* src = part_b + ((size_t)sp * H + h) * ROWS * 17 + r;   // [17][ROWS]: coalesced over rows
#pragma unroll
            for (int d = 0; d < 16; ++d) acc[d] = fmaf(wgt, __ldg(src + d * ROWS), acc[d]);
        }
        const float inv = 1.0f / acc[15];
        // columns 15h .. 15h+14 of row r (pairs may straddle the 4-byte packing: write element-wise)
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            __half vh, vl;
            split_h(acc[d] * inv, vh, vl);
            hi[r * PITCH + h * HD + d] = vh;
            lo[r * PITCH + h * HD + d] = vl;
        }
    }
    // zero the padded columns 120..127
    for (int i = threadIdx.x; i < ROWS * (EP - E); i += blockDim.x) {
        const int r = i / (EP - E), c = E + i % (EP - E);
        hi[r * PITCH + c] = __float2half_rn(0.f);
        lo[r * PITCH + c] = __float2half_rn(0.f);
    }
}
// planes from a row-major global matrix [nrows][ld] (zero padded to 64 x 128)
__device__ __forceinline__ void planes_from_global(const float* __restrict__ src, int nrows, int ld, int ncols,
                                                   __half* __restrict__ hi, __half* __restrict__ lo) {
    for (int i = threadIdx.x; i < ROWS * (EP / 2); i += blockDim.x) {
        const int r = i >> 6, c = (i & 63) * 2;
        float v0 = 0.f, v1 = 0.f;
        if (r < nrows && c < ncols) {
            v0 = __ldg(src + (size_t)r * ld + c);
            v1 = (c + 1 < ncols) ? __ldg(src + (size_t)r * ld + c + 1) : 0.f;
        }
        uint32_t h, l;
        split_h2(v0, v1, h, l);
        *reinterpret_cast<uint32_t*>(hi + r * PITCH + c) = h;
        *reinterpret_cast<uint32_t*>(lo + r * PITCH + c) = l;
    }
}

// ---- the 64 x 128 GEMM of one CTA: warp = (m tile, n half), eight n tiles per warp, weights via the smem ring
// epi(n, row, col, v0, v1) receives output elements (row, col) and (row, col + 1) of the warp's n-th tile, col relative to the window
template <class Epi>
__device__ __forceinline__ void gemm64(const WarpMap& w, Smem& s, const __half* ah, const __half* al, const uint4* wfrag,
                                       int ntiles_total, int nt_base, Epi epi) {
    float acc[NTW][4];
    mma_gemm_split_cta<8, PITCH, NTW>(ah, al, w.m0, w.nh, s.ring, wfrag, ntiles_total, nt_base, w.lane, acc);
#pragma unroll
    for (int n = 0; n < NTW; ++n) {
        const int col = 8 * (NTW * w.nh + n) + 2 * w.q4;
        epi(n, w.m0 + w.g, col, acc[n][0], acc[n][1]);
        epi(n, w.m0 + w.g + 8, col, acc[n][2], acc[n][3]);
    }
}
// this thread's 2 x NTW output-column biases, fetched BEFORE the GEMM so the load latency hides under the MMAs
struct BiasFrag {
    float v[NTW][2];
    __device__ __forceinline__ BiasFrag(const WarpMap& w, const float* __restrict__ bias) {
#pragma unroll
        for (int n = 0; n < NTW; ++n) {
            const int col = 8 * (NTW * w.nh + n) + 2 * w.q4;     // < 128: bias vectors are padded to 128 floats
            v[n][0] = __ldg(bias + col);
            v[n][1] = __ldg(bias + col + 1);
        }
    }
};

// early fetch of the first weight slabs of the next GEMM; legal right after a __syncthreads that follows the last GEMM
__device__ __forceinline__ void prefetch_w(Smem& s, const uint4* wfrag, int ntiles_total = 16, int nt_base = 0) {
    ring_prefetch_head<8>(s.ring, wfrag, ntiles_total, nt_base);
}

// out tile (K-major fp32) = A W^T + bias
__device__ __forceinline__ void linear_tile(const WarpMap& w, Smem& s, const __half* ah, const __half* al, const uint4* wf,
                                            const float* __restrict__ bias, float* __restrict__ out) {
    const BiasFrag bf(w, bias);
    gemm64(w, s, ah, al, wf, 16, 0, [&](int n, int r, int c, float v0, float v1) {
        if (c < E) {
            out[c * RP + r] = v0 + bf.v[n][0];
            out[(c + 1) * RP + r] = v1 + bf.v[n][1];
        }
    });
}
// rotary variant: the accumulator pair (c, c+1) is exactly rotary pair c/2 of row r
__device__ __forceinline__ void rotate_pair(const Smem& s, int r, int c, float& v0, float& v1) {
    const int pi = c >> 1;
    if (pi < NPAIR) {
        const int axis = pi / (E / 6), j = pi - axis * (E / 6);
        // |angle| is a few radians at most (normalised coordinates x frequency <= 1): one step of
        // 2*pi range reduction + the fast SFU sine/cosine (abs error ~5e-7) instead of libm's sincosf
        float ang = __fmul_rn(s.xyz[r * 3 + axis], s.freq[j]);
        ang = fmaf(-6.283185307179586f, rintf(ang * 0.15915494309189535f), ang);
        float sv, cv;
        __sincosf(ang, &sv, &cv);
        const float ev = v0, od = v1;
        v0 = ev * cv - od * sv;
        v1 = od * cv + ev * sv;
    }
}
// rotary projection written straight to the global fp16 Q of the next cross-attention [H][64][16]
__device__ __forceinline__ void linear_rope_q(const WarpMap& w, Smem& s, const __half* ah, const __half* al,
                                              const uint4* wf, const float* __restrict__ bias, __half* __restrict__ q) {
    const BiasFrag bf(w, bias);
    gemm64(w, s, ah, al, wf, 16, 0, [&](int n, int r, int c, float v0, float v1) {
        if (c < E) {
            v0 += bf.v[n][0];
            v1 += bf.v[n][1];
            rotate_pair(s, r, c, v0, v1);
            const int h0 = c / HD, h1 = (c + 1) / HD;
            q[(h0 * ROWS + r) * 16 + (c - h0 * HD)] = __float2half_rn(v0);
            q[(h1 * ROWS + r) * 16 + (c + 1 - h1 * HD)] = __float2half_rn(v1);
        } else {   // padded embed dims 120..127 own the pad slot of head c-120
            q[((c - E) * ROWS + r) * 16 + 15] = __float2half_rn(0.f);
            q[((c + 1 - E) * ROWS + r) * 16 + 15] = __float2half_rn(0.f);
        }
    });
}


// ---- in-CTA multi-head attention on the tensor cores -------------------------------------------
// Operands live as fp16 (hi, lo) planes [64 rows][PITCH] whose 128 columns are HEAD-PADDED: column
// 16 h + d holds channel 15 h + d, column 16 h + 15 is zero, so that one k16 step covers one head.
__device__ __forceinline__ int head_slot(int c) { return c + c / HD; }

// out planes (head-padded) = [rotary](A W^T + bias); the GEMM runs in the natural channel order (rotary
// pairs (2p, 2p+1) straddle heads) and the epilogue scatters to the padded columns
template <bool ROPE>
__device__ __forceinline__ void linear_heads(const WarpMap& w, Smem& s, const __half* ah, const __half* al, const uint4* wf,
                                             const float* __restrict__ bias, __half* __restrict__ oh, __half* __restrict__ ol) {
    const BiasFrag bf(w, bias);
    gemm64(w, s, ah, al, wf, 16, 0, [&](int n, int r, int c, float v0, float v1) {
        if (c < E) {
            v0 += bf.v[n][0];
            v1 += bf.v[n][1];
            if (ROPE) rotate_pair(s, r, c, v0, v1);
            __half h0, l0, h1, l1;
            split_h(v0, h0, l0);
            split_h(v1, h1, l1);
            const int s0 = r * PITCH + head_slot(c), s1 = r * PITCH + head_slot(c + 1);
            oh[s0] = h0;
            ol[s0] = l0;
            oh[s1] = h1;
            ol[s1] = l1;
        } else {   // GEMM columns 120..127 own the zero pad column of head c - 120
            const int s0 = r * PITCH + (c - E) * 16 + 15;
            oh[s0] = ol[s0] = oh[s0 + 16] = ol[s0 + 16] = __float2half_rn(0.f);
        }
    });
}
// head-padded planes of a row-major fp32 global matrix [nrows][E] (rows >= nrows zero)
__device__ __forceinline__ void head_planes_from_global(const float* __restrict__ src, int nrows, __half* __restrict__ hi,
                                                        __half* __restrict__ lo) {
    for (int i = threadIdx.x; i < ROWS * EP; i += blockDim.x) {
        const int r = i >> 7, slot = i & 127, h = slot >> 4, d = slot & 15;
        const float v = (r < nrows && d < HD) ? __ldg(src + (size_t)r * E + h * HD + d) : 0.f;
        __half vh, vl;
        split_h(v, vh, vl);
        hi[r * PITCH + slot] = vh;
        lo[r * PITCH + slot] = vl;
    }
}

// out planes (natural channel order, fp16 hi/lo) = softmax(q k^T) v per head; q carries hd^-1/2 log2(e).
// warp = (head, 32-row half); all three products use the error-compensated split (hi hi + 2^-11 (hi lo + lo hi)),
// the softmax is exact over the <= 64 keys (no online rescale), the denominator is summed in fp32.
// nk keys are valid; key_mask[j] != 0 => key j ignored (key_padding_mask, layers.py:178).
__device__ __forceinline__ void mma_mha(const WarpMap& w, const __half* __restrict__ qh, const __half* __restrict__ ql,
                                        const __half* __restrict__ kh, const __half* __restrict__ kl,
                                        const __half* __restrict__ vh, const __half* __restrict__ vl, int nk,
                                        const unsigned char* __restrict__ key_mask, __half* __restrict__ oh,
                                        __half* __restrict__ ol) {
    const int h = w.warp & 7, lane = w.lane, g = w.g, q4 = w.q4;
    // key validity of this thread's score columns (8 j + 2 q4 + {0, 1})
    uint32_t dead = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int key = 8 * j + 2 * q4 + e;
            if (key >= nk || (key_mask && key_mask[key])) dead |= 1u << (2 * j + e);
        }
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
        const int m0 = 32 * (w.warp >> 3) + 16 * mt;
        uint32_t fqh[4], fql[4];
        {
            const int row = m0 + (lane & 7) + 8 * ((lane >> 3) & 1);
            const int off = row * PITCH + 16 * h + 8 * (lane >> 4);
            ldmatrix_x4(fqh, smem_u32(qh + off));
            ldmatrix_x4(fql, smem_u32(ql + off));
        }
        float sc[8][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int key = kk * 16 + (lane & 7) + 8 * (lane >> 4);
            const int off = key * PITCH + 16 * h + 8 * ((lane >> 3) & 1);
            uint32_t bh[4], bl[4];
            ldmatrix_x4(bh, smem_u32(kh + off));
            ldmatrix_x4(bl, smem_u32(kl + off));
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f}, cor[4] = {0.f, 0.f, 0.f, 0.f};
                mma_16816(acc, fqh, bh[2 * t], bh[2 * t + 1]);
                mma_16816(cor, fqh, bl[2 * t], bl[2 * t + 1]);
                mma_16816(cor, fql, bh[2 * t], bh[2 * t + 1]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool off_key = (dead >> (2 * (2 * kk + t) + (e & 1))) & 1u;
                    sc[2 * kk + t][e] = off_key ? -INFINITY : fmaf(cor[e], kLoScaleInv, acc[e]);
                }
            }
        }
        float mx0 = sc[0][0], mx1 = sc[0][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx0 = fmaxf(mx0, fmaxf(sc[j][0], sc[j][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[j][2], sc[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j][0] = exp2f(sc[j][0] - mx0);
            sc[j][1] = exp2f(sc[j][1] - mx0);
            sc[j][2] = exp2f(sc[j][2] - mx1);
            sc[j][3] = exp2f(sc[j][3] - mx1);
            l0 += sc[j][0] + sc[j][1];
            l1 += sc[j][2] + sc[j][3];
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, oc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t ph[4], pl[4];
            split_h2(sc[2 * kk][0], sc[2 * kk][1], ph[0], pl[0]);
            split_h2(sc[2 * kk][2], sc[2 * kk][3], ph[1], pl[1]);
            split_h2(sc[2 * kk + 1][0], sc[2 * kk + 1][1], ph[2], pl[2]);
            split_h2(sc[2 * kk + 1][2], sc[2 * kk + 1][3], ph[3], pl[3]);
            const int key = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
            const int off = key * PITCH + 16 * h + 8 * (lane >> 4);
            uint32_t bh[4], bl[4];
            ldmatrix_x4_trans(bh, smem_u32(vh + off));
            ldmatrix_x4_trans(bl, smem_u32(vl + off));
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                mma_16816(o[n], ph, bh[2 * n], bh[2 * n + 1]);
                mma_16816(oc[n], ph, bl[2 * n], bl[2 * n + 1]);
                mma_16816(oc[n], pl, bh[2 * n], bh[2 * n + 1]);
            }
        }
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d = 8 * n + 2 * q4 + (e & 1);
                if (d < HD) {
                    const int row = m0 + g + 8 * (e >> 1);
                    __half vh_, vl_;
                    split_h(fmaf(oc[n][e], kLoScaleInv, o[n][e]) * ((e >> 1) ? i1 : i0), vh_, vl_);
                    oh[row * PITCH + h * HD + d] = vh_;
                    ol[row * PITCH + h * HD + d] = vl_;
                }
            }
    }
}

// FFN 120 -> 480 -> 120 in four 128-wide hidden chunks; y planes in (ah, al); result (+b2) -> out tile
__device__ __forceinline__ void ffn_tile(const WarpMap& w, Smem& s, const uint4* w1, const float* __restrict__ b1,
                                         const uint4* w2, const float* __restrict__ b2, float* __restrict__ out) {
    float sum[NTW][4];
#pragma unroll
    for (int n = 0; n < NTW; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) sum[n][e] = 0.f;
#pragma unroll 1
    for (int ch = 0; ch < FFP / 128; ++ch) {
        const BiasFrag bf(w, b1 + 128 * ch);
        gemm64(w, s, s.ah, s.al, w1, 64, 16 * ch, [&](int n, int r, int c, float v0, float v1) {
            v0 = fmaxf(v0 + bf.v[n][0], 0.f);
            v1 = fmaxf(v1 + bf.v[n][1], 0.f);
            uint32_t h, l;
            split_h2(v0, v1, h, l);
            *reinterpret_cast<uint32_t*>(s.hh + r * PITCH + c) = h;
            *reinterpret_cast<uint32_t*>(s.hl + r * PITCH + c) = l;
        });
        __syncthreads();
        prefetch_w(s, w2 + (size_t)ch * 8 * 16 * FRAG);
        {
            float acc[NTW][4];
            mma_gemm_split_cta<8, PITCH, NTW>(s.hh, s.hl, w.m0, w.nh, s.ring, w2 + (size_t)ch * 8 * 16 * FRAG, 16, 0, w.lane, acc);
#pragma unroll
            for (int n = 0; n < NTW; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) sum[n][e] += acc[n][e];
        }
        __syncthreads();
        if (ch + 1 < FFP / 128) prefetch_w(s, w1, 64, 16 * (ch + 1));
    }
#pragma unroll
    for (int n = 0; n < NTW; ++n) {
        const int c = 8 * (NTW * w.nh + n) + 2 * w.q4;
        if (c < E) {
            const float bb0 = __ldg(b2 + c), bb1 = __ldg(b2 + c + 1);
            out[c * RP + w.m0 + w.g] = sum[n][0] + bb0;
            out[(c + 1) * RP + w.m0 + w.g] = sum[n][1] + bb1;
            out[c * RP + w.m0 + w.g + 8] = sum[n][2] + bb0;
            out[(c + 1) * RP + w.m0 + w.g + 8] = sum[n][3] + bb1;
        }
    }
}

// attention to the <= 64 instruction tokens: x <- LN(x + Wo attn((x [+pe]) Wq, K, V))   (lang layers)
__device__ __forceinline__ void lang_attention(const WarpMap& w, Smem& s, const float* pe, int nrows,
                                               const uint4* wq, const float* bq, const uint4* wo, const float* bo,
                                               const float* g, const float* b, const float* kin, const float* vin,
                                               int n_instr) {
    planes_from_tile(s.xs, pe, nullptr, nullptr, nrows, s.ah, s.al, nullptr);
    head_planes_from_global(kin, n_instr, s.kh, s.kl);
    head_planes_from_global(vin, n_instr, s.vh, s.vl);
    __syncthreads();
    linear_heads<false>(w, s, s.ah, s.al, wq, bq, s.hh, s.hl);
    __syncthreads();
    prefetch_w(s, wo);
    mma_mha(w, s.hh, s.hl, s.kh, s.kl, s.vh, s.vl, n_instr, nullptr, s.ah, s.al);   // (ah, al): pad columns stay zero
    __syncthreads();
    linear_tile(w, s, s.ah, s.al, wo, bo, s.t2);
    __syncthreads();
    residual_layernorm<LPR>(s.xs, s.t2, g, b);
    __syncthreads();
}

// =============================================================== vision -> language (step-invariant)
struct CtxLangArgs {
    float* tok;            // [B][tok_rows][E], first nctx rows updated in place
    int tok_rows, nctx, n_instr, nlayers, batch;
    const float* kin;      // [nlayers][B][n_instr][E]  fp32 K projection of the instruction tokens
    const float* vin;      // same for V
    const uint4* w;        // [nlayers] LangW
    const float* v;        // [nlayers] LangV
};

__global__ void __launch_bounds__(THREADS, 1) cd_ctx_lang_kernel(const CtxLangArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem s(smem_raw);
    const WarpMap w;
    const int b = blockIdx.y, r0 = blockIdx.x * ROWS;
    const int nrows = min(ROWS, a.nctx - r0);
    float* tok_b = a.tok + ((size_t)b * a.tok_rows + r0) * E;
    load_tile(s.xs, tok_b, nrows, E);
    __syncthreads();
    for (int l = 0; l < a.nlayers; ++l) {
        const uint4* wl = a.w + (size_t)l * LangW::SIZE;
        const float* vl = a.v + (size_t)l * LangV::SIZE;
        const size_t kvo = ((size_t)l * a.batch + b) * a.n_instr * E;
        lang_attention(w, s, nullptr, 0, wl + LangW::WQ, vl + LangV::BQ, wl + LangW::WO, vl + LangV::BO, vl + LangV::G12,
                       vl + LangV::B12, a.kin + kvo, a.vin + kvo, a.n_instr);
        planes_from_tile(s.xs, nullptr, nullptr, nullptr, 0, s.ah, s.al, nullptr);
        __syncthreads();
        ffn_tile(w, s, wl + LangW::W1, vl + LangV::B1, wl + LangW::W2, vl + LangV::B2, s.t1);
        __syncthreads();
        residual_layernorm<LPR>(s.xs, s.t1, vl + LangV::G122, vl + LangV::B122);
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
    const float* traj_enc1;         // first trajectory-encoder layer, fp32 K-major [9][EP] + bias [EP]
    const uint4* traj_enc2;         // second layer, fragment order [8][16]
    const float* traj_enc2_b;       // [EP]
    const uint4* lang_w;            // LangW or null
    const float* lang_v;            // LangV
    const float* lang_k;            // [B][n_instr][E]
    const float* lang_vv;
    int n_instr;
    // ---- post kernel
    const float* x_in;              // [B][64][E]
    const float* att;               // [B][kCrossSplits][H][64][17] partial cross-attention outputs of cd_cross
    const uint4* layer_w;           // AdaW of this layer
    const float* layer_v;           // AdaV
    int ada_layer;                  // index of this layer in the adaLN table
    float* x_out;                   // [B][64][E]
    const uint4* reg_w;             // MlpW or null
    const float* reg_v;             // MlpV
    float* reg_out;                 // [B][L][reg_dim]
    int reg_dim;
    // ---- next layer's Q
    const float* next_src;          // null: use the tile just computed; else [B][64][E]
    const uint4* next_wq;           // [8][16] fragment-ordered W_q of the next cross-attention or null
    const float* next_bq;           // [EP]
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

__device__ __forceinline__ void next_q(const WarpMap& w, Smem& s, const StepArgs& a, int b, const float* src_tile) {
    const float* ada = a.ada + ((size_t)a.t_idx[b] * a.ada_layers + a.next_ada_layer) * ADA_ROW;
    planes_from_tile(src_tile, a.wp_pe, ada + 0, ada + EP, a.nrows, s.ah, s.al, nullptr);    // adaln_12(x + wp_pe)
    __syncthreads();
    linear_rope_q(w, s, s.ah, s.al, a.next_wq, a.next_bq, a.q_out + (size_t)b * H * ROWS * 16);
}

__global__ void __launch_bounds__(THREADS, 1) cd_step_begin_kernel(const StepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem s(smem_raw);
    const WarpMap w;
    const Map m;
    const int b = blockIdx.x;
    const float* traj_b = a.traj + (size_t)b * a.nrows * 9;
    init_common(s, traj_b, a.nrows, 9, nullptr);
    load_tile(s.t1, traj_b, a.nrows, 9, 9);
    __syncthreads();
    // trajectory encoder: Linear(9,E) -> ReLU (fp32 FMA, K = 9) -> Linear(E,E) (tensor cores)   (diffusion_head.py:43-48, 215)
    linear_to_smem<9, EP, true>(m, s.t1, a.traj_enc1, a.traj_enc1 + 9 * EP, 0, s.t2);
    __syncthreads();
    planes_from_tile(s.t2, nullptr, nullptr, nullptr, 0, s.ah, s.al, nullptr);
    __syncthreads();
    linear_tile(w, s, s.ah, s.al, a.traj_enc2, a.traj_enc2_b, s.xs);
    __syncthreads();
    if (a.lang_w)   // trajectory tokens attend to the instruction (diffusion_head.py:330-336)
        lang_attention(w, s, a.wp_pe, a.nrows, a.lang_w + LangW::WQ, a.lang_v + LangV::BQ, a.lang_w + LangW::WO,
                       a.lang_v + LangV::BO, a.lang_v + LangV::G12, a.lang_v + LangV::B12,
                       a.lang_k + (size_t)b * a.n_instr * E, a.lang_vv + (size_t)b * a.n_instr * E, a.n_instr);
    store_tile(a.x_out + (size_t)b * ROWS * E, s.xs, ROWS, E);
    if (a.next_wq) next_q(w, s, a, b, s.xs);
}

__global__ void __launch_bounds__(THREADS, 1) cd_post_kernel(const StepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem s(smem_raw);
    const WarpMap w;
    const int b = blockIdx.x;
    const uint4* lw = a.layer_w;
    const float* lv = a.layer_v;
    const float* ada = a.ada + ((size_t)a.t_idx[b] * a.ada_layers + a.ada_layer) * ADA_ROW;
    const float* traj_b = a.traj + (size_t)b * a.nrows * 9;
    prefetch_w(s, lw + AdaW::C_WO);
    init_common(s, traj_b, a.nrows, 9, a.mask ? a.mask + (size_t)b * a.nrows : nullptr);
    load_tile(s.xs, a.x_in + (size_t)b * ROWS * E, ROWS, E);
    planes_from_partials(a.att + (size_t)b * kCrossSplits * H * ROWS * 17, s.ah, s.al);
    __syncthreads();
    // ---- cross-attention epilogue: x = LN_12(x + att Wo^T + bo)          (layers.py:146-147)
    linear_tile(w, s, s.ah, s.al, lw + AdaW::C_WO, lv + AdaV::C_BO, s.t1);
    __syncthreads();
    prefetch_w(s, lw + AdaW::S_WV);
    residual_layernorm<LPR>(s.xs, s.t1, lv + AdaV::G12, lv + AdaV::B12);
    __syncthreads();
    // ---- self-attention: q = k = adaLN_1(x + pe), v = adaLN_1(x), rotary on q and k, padding mask (layers.py:165-182)
    planes_from_tile(s.xs, a.wp_pe, ada + 2 * EP, ada + 3 * EP, a.nrows, s.ah, s.al, nullptr);
    planes_from_tile(s.xs, nullptr, ada + 2 * EP, ada + 3 * EP, a.nrows, s.hh, s.hl, nullptr);
    __syncthreads();
    // V first: its input planes (hh, hl) are then free to receive Q (every warp has left the V GEMM's k loop
    // once it passes the first barrier of the K GEMM)
    linear_heads<false>(w, s, s.hh, s.hl, lw + AdaW::S_WV, lv + AdaV::S_BV, s.vh, s.vl);
    linear_heads<true>(w, s, s.ah, s.al, lw + AdaW::S_WK, lv + AdaV::S_BK, s.kh, s.kl);
    linear_heads<true>(w, s, s.ah, s.al, lw + AdaW::S_WQ, lv + AdaV::S_BQ, s.hh, s.hl);
    __syncthreads();
    prefetch_w(s, lw + AdaW::S_WO);
    mma_mha(w, s.hh, s.hl, s.kh, s.kl, s.vh, s.vl, a.nrows, a.mask ? s.mask : nullptr, s.ah, s.al);
    __syncthreads();
    linear_tile(w, s, s.ah, s.al, lw + AdaW::S_WO, lv + AdaV::S_BO, s.t2);
    __syncthreads();
    prefetch_w(s, lw + AdaW::W1, 64, 0);
    residual_layernorm<LPR>(s.xs, s.t2, lv + AdaV::G1, lv + AdaV::B1N);
    __syncthreads();
    // ---- FFN: y = adaLN_ff(x); x = LN_122(y + FFN(y))                      (layers.py:205-209)
    planes_from_tile(s.xs, nullptr, ada + 4 * EP, ada + 5 * EP, a.nrows, s.ah, s.al, s.t1);   // t1 = y (fp32)
    __syncthreads();
    ffn_tile(w, s, lw + AdaW::W1, lv + AdaV::B1, lw + AdaW::W2, lv + AdaV::B2, s.t2);
    __syncthreads();
    if (a.reg_w) prefetch_w(s, a.reg_w + MlpW::W1);
    else if (a.next_wq) prefetch_w(s, a.next_wq);
    residual_layernorm<LPR>(s.t1, s.t2, lv + AdaV::G122, lv + AdaV::B122);
    __syncthreads();
    float* x = s.t1;   // layer output
    store_tile(a.x_out + (size_t)b * ROWS * E, x, ROWS, E);

    // ---- regressor head: Linear(E,E) -> ReLU -> Linear(E,d)                (diffusion_head.py:179-198)
    if (a.reg_w) {
        planes_from_tile(x, nullptr, nullptr, nullptr, 0, s.ah, s.al, nullptr);
        __syncthreads();
        const BiasFrag breg(w, a.reg_v + MlpV::B1);
        gemm64(w, s, s.ah, s.al, a.reg_w + MlpW::W1, 16, 0, [&](int n, int r, int c, float v0, float v1) {
            v0 = fmaxf(v0 + breg.v[n][0], 0.f);
            v1 = fmaxf(v1 + breg.v[n][1], 0.f);
            uint32_t h, l;
            split_h2(v0, v1, h, l);
            *reinterpret_cast<uint32_t*>(s.hh + r * PITCH + c) = h;
            *reinterpret_cast<uint32_t*>(s.hl + r * PITCH + c) = l;
        });
        __syncthreads();
        linear_tile(w, s, s.hh, s.hl, a.reg_w + MlpW::W2, a.reg_v + MlpV::B2, s.t3);
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
        next_q(w, s, a, b, src);
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
    float* part;                   // [B][kCrossSplits][H][64][17]: unnormalised O (15) | pad | denominator l | row max m (log2 units)
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
    // this CTA's slice of the key tiles
    const int split = blockIdx.z;
    const int per = (a.ntiles + kCrossSplits - 1) / kCrossSplits;
    const int t_begin = min(split * per, a.ntiles), t_end = min(t_begin + per, a.ntiles);
    const int my_tiles = t_end - t_begin;
    const unsigned char* kv_b = a.kv + ((size_t)b * a.ntiles + t_begin) * tile_bytes;
    auto issue = [&](int t) {
        const int s = t % STAGES;
        if (t >= STAGES) mbar_wait(bar_empty + s, ((t / STAGES) - 1) & 1);
        mbar_expect_tx(bar_full + s, 4096);
        const unsigned char* src = kv_b + (size_t)t * tile_bytes;
        bulk_g2s(kvs[s], src + (size_t)h * 2048, 2048, bar_full + s);
        bulk_g2s(kvs[s] + 2048, src + (size_t)H * 2048 + (size_t)h * 2048, 2048, bar_full + s);
    };
    if (tid == 0)
        for (int t = 0; t < STAGES - 1 && t < my_tiles; ++t) issue(t);

    uint32_t qf[4];
    {
        const int row = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        ldmatrix_x4(qf, smem_u32(qs + row * 24 + 8 * (lane >> 4)));
    }
    float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float m0 = -INFINITY, m1 = -INFINITY;
    const bool tail_mask = (a.nk % kTileKeys) != 0;
    for (int t = 0; t < my_tiles; ++t) {
        const int stage = t % STAGES;
        if (tid == 0 && t + STAGES - 1 < my_tiles) issue(t + STAGES - 1);
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
        if (tail_mask && t_begin + t == a.ntiles - 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if ((t_begin + t) * kTileKeys + 8 * j + 2 * q4 + (e & 1) >= a.nk) s[j][e] = -INFINITY;
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
    // unnormalised partial, stored [17][ROWS]: planes 0..14 = O, 15 = denominator (rode in V's slot 15), 16 = row max (-inf if this
    // slice was empty); cd_post merges the kCrossSplits partials of a row
    float* out = a.part + (((size_t)b * kCrossSplits + split) * H + h) * ROWS * 17;
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int d = 8 * n + 2 * q4 + (e & 1);
            const int row = warp * 16 + g + 8 * (e >> 1);
            out[d * ROWS + row] = o[n][e];
        }
    if (q4 == 0) {
        out[16 * ROWS + warp * 16 + g] = m0;
        out[16 * ROWS + warp * 16 + g + 8] = m1;
    }
}

}  // namespace cd
}  // namespace a3d

// ===================================================================== C ABI
using namespace a3d;
using namespace a3d::cd;

extern "C" size_t cd_pack_floats(int which) {
    switch (which) {
        case 0: return LangV::SIZE;
        case 1: return AdaV::SIZE;
        case 2: return MlpV::SIZE;
        case 3: return ADA_ROW;
        case 4: return (size_t)LangW::SIZE * 4;   // fragment buffers, in 32-bit words
        case 5: return (size_t)AdaW::SIZE * 4;
        case 6: return (size_t)MlpW::SIZE * 4;
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
                           const float* vin, int n_instr, const void* w, const float* v, int nlayers, void* stream) {
    A3D_REQUIRE(tok && kin && vin && w && v, "cd_ctx_lang: null pointer");
    A3D_REQUIRE(embed == E && heads == H, "cd_ctx_lang: built for embedding_dim 120 / 8 heads (got %d / %d)", embed, heads);
    A3D_REQUIRE(batch > 0 && nctx > 0 && nctx <= tok_rows && n_instr > 0 && n_instr <= 64 && nlayers > 0,
                "cd_ctx_lang: bad sizes (nctx=%d rows=%d n_instr=%d)", nctx, tok_rows, n_instr);
    static PerDeviceOnce once_dev;
    if (bool& once = once_dev.flag(); !once) {
        if (int rc = set_smem((const void*)cd_ctx_lang_kernel, SMEM_BYTES)) return rc;
        once = true;
    }
    CtxLangArgs a{tok, tok_rows, nctx, n_instr, nlayers, batch, kin, vin, (const uint4*)w, v};
    dim3 grid((nctx + ROWS - 1) / ROWS, batch);
    cd_ctx_lang_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(a);
    return check_launch("cd_ctx_lang");
}

extern "C" int cd_step_begin(const float* traj, int batch, int length, const float* wp_pe, const int* t_idx,
                             const float* ada, int ada_layers, const float* traj_enc1, const void* traj_enc2,
                             const float* traj_enc2_b, const void* lang_w, const float* lang_v, const float* lang_k,
                             const float* lang_vv, int n_instr, float* x_out, const void* next_wq,
                             const float* next_bq, int next_ada_layer, void* q_out, void* stream) {
    A3D_REQUIRE(traj && wp_pe && t_idx && ada && traj_enc1 && traj_enc2 && traj_enc2_b && x_out, "cd_step_begin: null pointer");
    A3D_REQUIRE(batch > 0 && length > 0 && length <= ROWS, "cd_step_begin: trajectory length %d not in [1,64]", length);
    A3D_REQUIRE(!lang_w || (lang_v && lang_k && lang_vv && n_instr > 0 && n_instr <= 64), "cd_step_begin: instruction K/V missing");
    A3D_REQUIRE(!next_wq || (q_out && next_bq), "cd_step_begin: q_out / next_bq missing");
    static PerDeviceOnce once_dev;
    if (bool& once = once_dev.flag(); !once) {
        if (int rc = set_smem((const void*)cd_step_begin_kernel, SMEM_BYTES)) return rc;
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
    a.traj_enc1 = traj_enc1;
    a.traj_enc2 = (const uint4*)traj_enc2;
    a.traj_enc2_b = traj_enc2_b;
    a.lang_w = (const uint4*)lang_w;
    a.lang_v = lang_v;
    a.lang_k = lang_k;
    a.lang_vv = lang_vv;
    a.n_instr = n_instr;
    a.x_out = x_out;
    a.next_wq = (const uint4*)next_wq;
    a.next_bq = next_bq;
    a.next_ada_layer = next_ada_layer;
    a.q_out = (__half*)q_out;
    cd_step_begin_kernel<<<batch, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(a);
    return check_launch("cd_step_begin");
}

extern "C" size_t cd_cross_part_floats(int batch) { return (size_t)batch * kCrossSplits * H * ROWS * 17; }

extern "C" int cd_cross(const void* q, const void* kv, int batch, int nk, int heads, float* att, void* stream) {
    A3D_REQUIRE(q && kv && att, "cd_cross: null pointer");
    A3D_REQUIRE(heads == H && batch > 0 && nk > 0, "cd_cross: built for 8 heads (got %d)", heads);
    A3D_REQUIRE(((uintptr_t)kv & 15) == 0, "cd_cross: K/V cache must be 16-byte aligned");
    CrossArgs a{(const __half*)q, (const unsigned char*)kv, nk, (nk + kTileKeys - 1) / kTileKeys, batch, att};
    dim3 grid(H, batch, kCrossSplits);
    cd_cross_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    return check_launch("cd_cross");
}

extern "C" int cd_post(const float* traj, int batch, int length, const unsigned char* mask, const float* wp_pe,
                       const int* t_idx, const float* ada, int ada_layers, int ada_layer, const float* x_in,
                       const float* att, const void* layer_w, const float* layer_v, float* x_out, const void* reg_w,
                       const float* reg_v, float* reg_out, int reg_dim, const float* next_src, const void* next_wq,
                       const float* next_bq, int next_ada_layer, void* q_out, int do_update, int last_step,
                       float* traj_out, const float* pos_upd, const float* cond_data, const unsigned char* cond_mask,
                       const float* coef_host, const float* noise_pos, const float* noise_rot, void* stream) {
    A3D_REQUIRE(traj && wp_pe && t_idx && ada && x_in && att && layer_w && layer_v && x_out, "cd_post: null pointer");
    A3D_REQUIRE(batch > 0 && length > 0 && length <= ROWS, "cd_post: trajectory length %d not in [1,64]", length);
    A3D_REQUIRE(!reg_w || (reg_v && reg_out && reg_dim > 0 && reg_dim <= 16), "cd_post: regressor output missing");
    A3D_REQUIRE(!next_wq || (q_out && next_bq), "cd_post: q_out / next_bq missing");
    A3D_REQUIRE(!do_update || (traj_out && pos_upd && cond_data && cond_mask && coef_host && reg_w && reg_dim == 6 &&
                               (last_step || (noise_pos && noise_rot))),
                "cd_post: DDPM update needs traj_out, pos_upd, cond_*, coef, the rotation regressor and noise");
    static PerDeviceOnce once_dev;
    if (bool& once = once_dev.flag(); !once) {
        if (int rc = set_smem((const void*)cd_post_kernel, SMEM_BYTES)) return rc;
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
    a.layer_w = (const uint4*)layer_w;
    a.layer_v = layer_v;
    a.x_out = x_out;
    a.reg_w = (const uint4*)reg_w;
    a.reg_v = reg_v;
    a.reg_out = reg_out;
    a.reg_dim = reg_dim;
    a.next_src = next_src;
    a.next_wq = (const uint4*)next_wq;
    a.next_bq = next_bq;
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
    cd_post_kernel<<<batch, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(a);
    return check_launch("cd_post");
}
