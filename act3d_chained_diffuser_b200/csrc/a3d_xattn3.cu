// Fused cross-attention stack, tcgen05 / TMEM generation (same contract and packed weights as a3d_xattn2.cu).
//
// CTA = 128 query rows of one sample, 160 threads:
//   warps 0-3  "row warps": thread i owns query row i == TMEM lane i.  They run the projections / LayerNorm /
//              FFN (register-chained split-fp16 mma.sync GEMMs shared with a3d_xattn2.cu) and the softmax:
//              scores are read from tensor memory with tcgen05.ld one full row per thread, so row max /
//              exp2 / packing need no shuffles at all, and P goes back to tensor memory with tcgen05.st.
//   warp 4     one elected lane issues the K/V tile loads (1-D bulk async copies on an mbarrier ring) and all
//              tcgen05.mma instructions:  S_h = Q_h K_h^T (M=128, N=64, K=16, operands straight from shared memory
//              through SWIZZLE_32B K-major descriptors -- the K/V tile images are already in that canonical
//              layout) and O_h += P_h V_h (A = P from tensor memory, B = V MN-major).  Completion is signalled
//              with tcgen05.commit on mbarriers.
// Two passes over the keys per layer: pass 1 finds the exact row maxima (S only), pass 2 recomputes S, forms
// P = 2^(S - m) and accumulates O in tensor memory -- no running-max correction of O is ever needed, and the
// extra QK^T pass is free on a tensor pipe that the MUFU-bound softmax leaves mostly idle.
// Status (round 1): parity-green on every kernel / end-to-end test, 3.6 ms per C2 launch vs 2.8 ms for the mma.sync core,
// so it is opt-in (a3d_set_option("xattn_core", 3) / A3D_XATTN_CORE=3).  ncu (profiles/r1_xattn_ghost_v3_ncu.txt): only
// ~10 resident warps/SM, 21 % of stall samples in mbarrier spin loops, XU 58 %: the single issuing lane (several
// ~100-cycle try_waits + 5 UMMAs per 8192-score unit) and 4 row warps per CTA are the limiters, not TMEM bandwidth
// (21 %).  Next: single pass with conditional O correction, S and PV issue on separate warps, 8 row warps per CTA.
// Tensor-memory map (256 columns): O = 4 heads x 16 columns at 0..63 (slot 15 = softmax denominator),
// S/P buffer i (i = unit & 1) at 64 + 64 i (P overwrites the first 32 columns of S as packed fp16 pairs).
#include "a3d_xattn_common.cuh"

namespace a3d {

struct Xa3 {
    static constexpr int E = 60, H = 4, ROWS = 128, THREADS = 160;
    static constexpr int XP = Xa2::XP;
    static constexpr int TILE_BYTES = Xa2::TILE_BYTES, STAGES = 3;
    static constexpr size_t X_BYTES = (size_t)ROWS * XP * 4;          // parked residual stream
    static constexpr size_t Q_BYTES = (size_t)H * ROWS * 32;          // Q_h tiles [128][16] fp16, SWIZZLE_32B
    static constexpr size_t RING_BYTES = (size_t)STAGES * TILE_BYTES; // K/V ring; the O tile aliases it in the epilogue
    static constexpr size_t SMEM = X_BYTES + Q_BYTES + RING_BYTES + 256;
    static constexpr int TMEM_COLS = 256, O_COL = 0, S_COL = 64;
};

// ------------------------------------------------------------------ tcgen05 wrappers
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor: SWIZZLE_32B, 8-row groups 256 B apart (both for K-major [rows][16 halfs] tiles and
// for the MN-major V tile [keys][16 halfs]); version 1 (Blackwell); see cute/arch/mma_sm100_desc.hpp
__device__ __forceinline__ uint64_t sw32_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)6 << 61);
}
// instruction descriptor, kind::f16: D = F32 (bit 4), A = B = F16, M = 128 (bits 24-28 = 8), N >> 3 at bits 17-22
constexpr uint32_t kIdescS = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);                    // K-major A, K-major B
constexpr uint32_t kIdescPV = (1u << 4) | (1u << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);      // B (V) MN-major

#define A3D_R32(r) \
    "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), \
    "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), \
    "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), \
    "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define A3D_R32_IN(r) \
    "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), \
    "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),  \
    "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),  \
    "r"(r[31])

// 32 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : A3D_R32(r)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
        :
        : A3D_R32_IN(r), "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Xa3Bars {
    uint64_t kv_full[Xa3::STAGES], kv_empty[Xa3::STAGES];
    uint64_t s_full[2], p_full[2], sfree1[2], sfree2[2];
    uint64_t q_ready, o_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(Xa3::THREADS, 2) xattn3_kernel(const Xa2Args a) {
    using C = Xa3;
    constexpr int E = C::E, H = C::H;
    extern __shared__ __align__(1024) unsigned char smem[];
    float* xpark = reinterpret_cast<float*>(smem);                                // [128][XP] residual stream
    unsigned char* qs = smem + C::X_BYTES;                                        // [H][128][32 B] fp16, SW32
    unsigned char* kvs = smem + C::X_BYTES + C::Q_BYTES;                          // STAGES x TILE_BYTES
    float* opark = reinterpret_cast<float*>(kvs);                                 // [128][XP] attention output (epilogue)
    Xa3Bars* bars = reinterpret_cast<Xa3Bars*>(kvs + C::RING_BYTES);
    __shared__ float freq[E / 6];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, row0 = blockIdx.x * C::ROWS;
    const int g = lane >> 2, q4 = lane & 3;

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(bars->kv_full + s, 1);
            mbar_init(bars->kv_empty + s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bars->s_full + i, 1);
            mbar_init(bars->p_full + i, 128);
            mbar_init(bars->sfree1 + i, 128);
            mbar_init(bars->sfree2 + i, 1);
        }
        mbar_init(&bars->q_ready, 128);
        mbar_init(&bars->o_full, 1);
        mbar_fence_init();
    }
    if (tid < E / 6) freq[tid] = rope_freq<E>(tid);
    if (warp == 4) tmem_alloc(&bars->tmem_base, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    const int nt = a.ntiles;
    const unsigned char* kv_sample = a.kv_base + (size_t)b * nt * C::TILE_BYTES;

    if (warp == 4) {
        // =========================================================== producer: TMA loads + all tensor-core issue
        if (lane == 0) {
            uint32_t gt_load = 0;               // tiles requested so far (2 * nt per layer: pass 1 then pass 2)
            for (int layer = 0; layer < a.nlayers; ++layer) {
                const unsigned char* kv_layer = kv_sample + (size_t)layer * a.kv_layer_stride;
                const uint32_t base1 = (uint32_t)layer * 2 * nt, base2 = base1 + nt, lim = base1 + 2 * nt;
                auto load_next = [&]() {
                    if (gt_load >= lim) return;
                    const uint32_t s = gt_load % C::STAGES, use = gt_load / C::STAGES;
                    if (use >= 1) mbar_wait(bars->kv_empty + s, (use - 1) & 1);
                    const uint32_t tile = (gt_load - base1) % nt;
                    mbar_expect_tx(bars->kv_full + s, C::TILE_BYTES);
                    bulk_g2s(kvs + s * C::TILE_BYTES, kv_layer + (size_t)tile * C::TILE_BYTES, C::TILE_BYTES, bars->kv_full + s);
                    ++gt_load;
                };
                mbar_wait(&bars->q_ready, layer & 1);       // Q of this layer is in shared memory; O tile no longer read
                tc_fence_after();
                for (int i = 0; i < C::STAGES - 1; ++i) load_next();
                const uint32_t q_addr = smem_u32(qs);
                // ---------------- pass 1: S only
                for (int t = 0; t < nt; ++t) {
                    const uint32_t tau = base1 + t, s = tau % C::STAGES;
                    mbar_wait(bars->kv_full + s, (tau / C::STAGES) & 1);
                    tc_fence_after();
                    const uint32_t k_addr = smem_u32(kvs + s * C::TILE_BYTES);
                    for (int h = 0; h < H; ++h) {
                        const int u = t * H + h, i = u & 1, k = u >> 1;
                        if (k >= 1) mbar_wait(bars->sfree1 + i, (k - 1) & 1);
                        tc_fence_after();
                        umma_ss(tmem + C::S_COL + 64 * i, sw32_desc(q_addr + h * 4096), sw32_desc(k_addr + h * 2048), kIdescS, 0);
                        tc_commit(bars->s_full + i);
                    }
                    tc_commit(bars->kv_empty + s);
                    load_next();                // reuses the stage of tile tau-1, whose release was committed one iteration ago
                }
                // both S buffers must have been read by the row warps before pass 2 overwrites them
                {
                    const int last_k = (nt * H) / 2 - 1;
                    mbar_wait(bars->sfree1 + 0, last_k & 1);
                    mbar_wait(bars->sfree1 + 1, last_k & 1);
                }
                // ---------------- pass 2: S -> (row warps: P) -> O += P V; the PV of unit u-1 is issued after S of unit u
                for (int u = 0; u <= nt * H; ++u) {
                    if (u < nt * H) {
                        const int t = u / H, h = u % H, i = u & 1, k = u >> 1;
                        const uint32_t tau = base2 + t, s = tau % C::STAGES;
                        if (h == 0) mbar_wait(bars->kv_full + s, (tau / C::STAGES) & 1);
                        if (k >= 1) mbar_wait(bars->sfree2 + i, (k - 1) & 1);
                        tc_fence_after();
                        const uint32_t k_addr = smem_u32(kvs + s * C::TILE_BYTES);
                        umma_ss(tmem + C::S_COL + 64 * i, sw32_desc(q_addr + h * 4096), sw32_desc(k_addr + h * 2048), kIdescS, 0);
                        tc_commit(bars->s_full + i);
                    }
                    if (u >= 1) {
                        const int v = u - 1, t = v / H, h = v % H, j = v & 1, k = v >> 1;
                        const uint32_t s = (base2 + t) % C::STAGES;
                        mbar_wait(bars->p_full + j, k & 1);
                        tc_fence_after();
                        const uint32_t v_addr = smem_u32(kvs + s * C::TILE_BYTES) + H * 2048 + h * 2048;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_ts(tmem + C::O_COL + 16 * h, tmem + C::S_COL + 64 * j + 8 * ks, sw32_desc(v_addr + ks * 512),
                                    kIdescPV, (t > 0 || ks > 0) ? 1u : 0u);
                        tc_commit(bars->sfree2 + j);
                        if (h == H - 1) {
                            tc_commit(bars->kv_empty + s);
                            load_next();
                        }
                    }
                }
                tc_commit(&bars->o_full);
            }
        }
    } else {
        // =========================================================== row warps
        const int lrow = warp * 32 + lane;                      // this thread's row == TMEM lane
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        // ---- residual stream -> xpark (row per thread)
        {
            const int row = row0 + lrow;
            const float* xp = a.x0 + (long)b * a.x0_sb + (long)row * a.x0_sn;
            for (int c = 0; c < 64; c += 2) {
                float2 v = make_float2(0.f, 0.f);
                if (row < a.nq && c < E) v = __ldg(reinterpret_cast<const float2*>(xp + c));
                *reinterpret_cast<float2*>(xpark + lrow * C::XP + c) = v;
            }
        }
        __syncwarp();

        for (int layer = 0; layer < a.nlayers; ++layer) {
            const uint4* w = a.w + (size_t)layer * Xa2::LAYER_W;
            const float* vv = a.v + (size_t)layer * Xa2::LAYER_V;
            // ---------------------------------------------------------------- Q = rotary(x Wq^T + bq) -> smem (SW32 tiles per head)
            for (int mt = 0; mt < 2; ++mt) {
                const int r0 = warp * 32 + 16 * mt + g, r1 = r0 + 8;
                float xr[8][4], qa[8][4];
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const float2 v0 = *reinterpret_cast<const float2*>(xpark + r0 * C::XP + 8 * n + 2 * q4);
                    const float2 v1 = *reinterpret_cast<const float2*>(xpark + r1 * C::XP + 8 * n + 2 * q4);
                    xr[n][0] = v0.x;
                    xr[n][1] = v0.y;
                    xr[n][2] = v1.x;
                    xr[n][3] = v1.y;
                }
                gemm_reg(xr, w + Xa2::W_Q, lane, qa);
                float qxyz[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
                if (a.qpos) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int row = row0 + (r ? r1 : r0);
                        if (row < a.nq)
                            for (int ax = 0; ax < 3; ++ax) qxyz[r][ax] = __ldg(a.qpos + ((long)b * a.nq + row) * 3 + ax);
                    }
                }
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int c = 8 * n + 2 * q4;
                    const float b0 = __ldg(vv + Xa2::B_Q + c), b1 = __ldg(vv + Xa2::B_Q + c + 1);
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
                        const int row = r ? r1 : r0;
                        const int swz = (row >> 2) & 1;
                        auto put = [&](int hh, int d, float val) {
                            *reinterpret_cast<__half*>(qs + hh * 4096 + row * 32 + (((d >> 3) ^ swz) << 4) + (d & 7) * 2) =
                                __float2half_rn(val);
                        };
                        if (c < E) {
                            put(c / 15, c % 15, v0);
                            put((c + 1) / 15, (c + 1) % 15, v1);
                        } else {   // dims 60..63 own the pad slot (d = 15) of heads 0..3
                            put(c - E, 15, 0.f);
                            put(c + 1 - E, 15, 0.f);
                        }
                    }
                }
            }
            fence_async_smem();                 // generic-proxy writes of Q -> visible to the tensor-core (async) proxy
            tc_fence_before();
            mbar_arrive(&bars->q_ready);

            // ---------------------------------------------------------------- pass 1: exact row maxima
            float m[H];
#pragma unroll
            for (int h = 0; h < H; ++h) m[h] = -INFINITY;
            for (int t = 0; t < nt; ++t) {
                const int valid = min(kTileKeys, a.nk - t * kTileKeys);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const int u = t * H + h, i = u & 1, k = u >> 1;
                    mbar_wait(bars->s_full + i, k & 1);
                    tc_fence_after();
                    uint32_t r0[32], r1[32];
                    tmem_ld32(lane_addr + C::S_COL + 64 * i, r0);
                    tmem_ld32(lane_addr + C::S_COL + 64 * i + 32, r1);
                    tmem_wait_ld();
                    tc_fence_before();
                    mbar_arrive(bars->sfree1 + i);
                    float mx = m[h];
                    if (valid == kTileKeys) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[c]), __uint_as_float(r1[c])));
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            if (c < valid) mx = fmaxf(mx, __uint_as_float(r0[c]));
                            if (c + 32 < valid) mx = fmaxf(mx, __uint_as_float(r1[c]));
                        }
                    }
                    m[h] = mx;
                }
            }
            // ---------------------------------------------------------------- pass 2: P = 2^(S - m) -> tensor memory
            for (int t = 0; t < nt; ++t) {
                const int valid = min(kTileKeys, a.nk - t * kTileKeys);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const int u = t * H + h, i = u & 1, k = u >> 1;
                    mbar_wait(bars->s_full + i, k & 1);      // pass-2 uses follow 2*nt pass-1 uses: same parity pattern
                    tc_fence_after();
                    uint32_t r0[32], r1[32], p[32];
                    tmem_ld32(lane_addr + C::S_COL + 64 * i, r0);
                    tmem_ld32(lane_addr + C::S_COL + 64 * i + 32, r1);
                    tmem_wait_ld();
                    const float mh = m[h];
                    if (valid == kTileKeys) {      // full tile: no per-element masking code at all
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            p[c] = pack_h2(exp2_fast(__uint_as_float(r0[2 * c]) - mh), exp2_fast(__uint_as_float(r0[2 * c + 1]) - mh));
                            p[16 + c] = pack_h2(exp2_fast(__uint_as_float(r1[2 * c]) - mh), exp2_fast(__uint_as_float(r1[2 * c + 1]) - mh));
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const float e0 = (2 * c < valid) ? exp2_fast(__uint_as_float(r0[2 * c]) - mh) : 0.f;
                            const float e1 = (2 * c + 1 < valid) ? exp2_fast(__uint_as_float(r0[2 * c + 1]) - mh) : 0.f;
                            const float e2 = (2 * c + 32 < valid) ? exp2_fast(__uint_as_float(r1[2 * c]) - mh) : 0.f;
                            const float e3 = (2 * c + 33 < valid) ? exp2_fast(__uint_as_float(r1[2 * c + 1]) - mh) : 0.f;
                            p[c] = pack_h2(e0, e1);
                            p[16 + c] = pack_h2(e2, e3);
                        }
                    }
                    tmem_st32(lane_addr + C::S_COL + 64 * i, p);
                    tmem_wait_st();
                    tc_fence_before();
                    mbar_arrive(bars->p_full + i);
                }
            }
            // ---------------------------------------------------------------- O -> normalise -> opark (aliases the idle K/V ring)
            mbar_wait(&bars->o_full, layer & 1);
            tc_fence_after();
            {
                uint32_t o0[32], o1[32];
                tmem_ld32(lane_addr + C::O_COL, o0);
                tmem_ld32(lane_addr + C::O_COL + 32, o1);
                tmem_wait_ld();
                tc_fence_before();
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const uint32_t* src = (h < 2) ? (o0 + 16 * h) : (o1 + 16 * (h - 2));
                    const float inv = 1.0f / __uint_as_float(src[15]);
#pragma unroll
                    for (int d = 0; d < 16; d += 2)
                        *reinterpret_cast<float2*>(opark + lrow * C::XP + 16 * h + d) =
                            make_float2(__uint_as_float(src[d]) * inv, __uint_as_float(src[d + 1]) * inv);
                }
            }
            __syncwarp();
            // ---------------------------------------------------------------- out-proj + LN, FFN + LN (two 16-row tiles per warp)
            const bool last = (layer == a.nlayers - 1);
            for (int mt = 0; mt < 2; ++mt) {
                const int r0 = warp * 32 + 16 * mt + g, r1 = r0 + 8;
                float xr[8][4];
                {
                    float of[8][4], y[8][4];
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        const float2 v0 = *reinterpret_cast<const float2*>(opark + r0 * C::XP + 8 * n + 2 * q4);
                        const float2 v1 = *reinterpret_cast<const float2*>(opark + r1 * C::XP + 8 * n + 2 * q4);
                        of[n][0] = v0.x;
                        of[n][1] = v0.y;
                        of[n][2] = v1.x;
                        of[n][3] = v1.y;
                    }
                    gemm_reg(of, w + Xa2::W_O, lane, y);
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        const int c = 8 * n + 2 * q4;
                        const float b0 = __ldg(vv + Xa2::B_O + c), b1 = __ldg(vv + Xa2::B_O + c + 1);
                        const float2 x0v = *reinterpret_cast<const float2*>(xpark + r0 * C::XP + c);
                        const float2 x1v = *reinterpret_cast<const float2*>(xpark + r1 * C::XP + c);
                        xr[n][0] = y[n][0] + b0 + x0v.x;
                        xr[n][1] = y[n][1] + b1 + x0v.y;
                        xr[n][2] = y[n][2] + b0 + x1v.x;
                        xr[n][3] = y[n][3] + b1 + x1v.y;
                    }
                    layernorm_frag(xr, q4, vv + Xa2::G_1, vv + Xa2::BE_1);
                }
                {
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        *reinterpret_cast<float2*>(xpark + r0 * C::XP + 8 * n + 2 * q4) = make_float2(xr[n][0], xr[n][1]);
                        *reinterpret_cast<float2*>(xpark + r1 * C::XP + 8 * n + 2 * q4) = make_float2(xr[n][2], xr[n][3]);
                    }
                    float hid[8][4];
                    gemm_reg(xr, w + Xa2::W_1, lane, hid);
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        const int c = 8 * n + 2 * q4;
                        const float b0 = __ldg(vv + Xa2::B_1 + c), b1 = __ldg(vv + Xa2::B_1 + c + 1);
                        hid[n][0] = fmaxf(hid[n][0] + b0, 0.f);
                        hid[n][1] = fmaxf(hid[n][1] + b1, 0.f);
                        hid[n][2] = fmaxf(hid[n][2] + b0, 0.f);
                        hid[n][3] = fmaxf(hid[n][3] + b1, 0.f);
                    }
                    gemm_reg(hid, w + Xa2::W_2, lane, xr);
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        const int c = 8 * n + 2 * q4;
                        const float b0 = __ldg(vv + Xa2::B_2 + c), b1 = __ldg(vv + Xa2::B_2 + c + 1);
                        const float2 x0v = *reinterpret_cast<const float2*>(xpark + r0 * C::XP + c);
                        const float2 x1v = *reinterpret_cast<const float2*>(xpark + r1 * C::XP + c);
                        xr[n][0] += b0 + x0v.x;
                        xr[n][1] += b1 + x0v.y;
                        xr[n][2] += b0 + x1v.x;
                        xr[n][3] += b1 + x1v.y;
                    }
                    layernorm_frag(xr, q4, vv + Xa2::G_2, vv + Xa2::BE_2);
#pragma unroll
                    for (int n = 0; n < 8; ++n) {   // x of the next layer
                        *reinterpret_cast<float2*>(xpark + r0 * C::XP + 8 * n + 2 * q4) = make_float2(xr[n][0], xr[n][1]);
                        *reinterpret_cast<float2*>(xpark + r1 * C::XP + 8 * n + 2 * q4) = make_float2(xr[n][2], xr[n][3]);
                    }
                }
                // ---- outputs of this layer
                if (a.feat_out && (a.feat_all || last)) {
                    float* fo = a.feat_out + ((size_t)(a.feat_all ? layer : 0) * a.batch + b) * (size_t)a.feat_rows * E;
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int row = row0 + (r ? r1 : r0);
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
                            if (row0 + r0 < a.nq) lo[row0 + r0] = p0;
                            if (row0 + r1 < a.nq) lo[row0 + r1] = p1;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4) tmem_dealloc(tmem, C::TMEM_COLS);
}

}  // namespace a3d

using namespace a3d;

// launched by a3d_xattn_stack (a3d_xattn2.cu) when the "xattn_core" option selects the tcgen05 kernel
int a3d_launch_xattn3(const Xa2Args& a, dim3 grid, cudaStream_t stream) {
    static bool once = false;
    if (!once) {
        cudaError_t e = cudaFuncSetAttribute(xattn3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Xa3::SMEM);
        if (e != cudaSuccess) {
            set_error("a3d_xattn_stack(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return A3D_ECUDA;
        }
        once = true;
    }
    xattn3_kernel<<<grid, Xa3::THREADS, Xa3::SMEM, stream>>>(a);
    return check_launch("a3d_xattn_stack(tcgen05)");
}
