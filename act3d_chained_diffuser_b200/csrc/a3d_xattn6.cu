// Fused cross-attention stack, tcgen05 / TMEM generation 3 (same contract as a3d_xattn2.cu / a3d_xattn4.cu).
//
// What changed against a3d_xattn4.cu (measured there: 0.39 ms of a 2.66 ms C2 launch outside the key loop, XU pipe
// 79 % busy inside it):
//   * every linear layer of the stack (q-proj, out-proj, FFN 1/2) is a tcgen05.mma too: the activations of a row go
//     to tensor memory as split fp16 (hi, lo) A operands (TS form), the weights stream through the K/V ring as
//     16 KiB SWIZZLE_32B operand images (packing.pack_umma_weight), the accumulators come back one row per thread, so
//     bias / residual / LayerNorm / ReLU / rotary need no shuffles and no weight fragment ever comes from L2 with a
//     warp waiting on it;
//   * 8 exponential warps per 128-row tile: two threads share a row and own 32 of the 64 scores of a unit each
//     (4 warps per SM sub-partition with 2 CTAs / SM), and NP of a thread's 16 score pairs per unit are evaluated as
//     a Cody-Waite + degree-3 polynomial on packed fp32 pairs (FADD2 / FFMA2) next to the MUFU unit, which is the
//     binding unit of this kernel (tools/micro/mf_bench.cu: 560 -> 366 cycles per unit per SM at NP = 8);
//   * the issuing warp runs warp-uniform code with elect-predicated tcgen05 / bulk-copy instructions.
//
// CTA = 128 query rows of one sample, 288 threads (warps 0-7 exponentials, warps 0-3 also own the rows in the
// linear layers, warp 8 issues), 2 CTAs / SM, 256 tensor-memory columns:
//   O (4 heads x 16 columns, slot 15 = softmax denominator) at 0..63, S/P buffer i = unit % 3 at 64 + 64 i.
//   In the linear layers the idle S buffers hold A (hi: 64..95, lo: 96..127), the main accumulator (128..191)
//   and the correction accumulator (192..255).
// Ring items per layer (16 KiB each, one bulk copy): W_q | K/V tiles 0..nt-1 | W_o | W_1 | W_2.
// Softmax: single pass, stale shift folded into the QK^T product through the pad slot (Q[15] = -shift, K[15] = 1),
// unchecked fast pass + end-of-pass verdict + safe-mode replay exactly as in a3d_xattn4.cu.
#include "a3d_tcgen05.cuh"
#include "a3d_xattn_common.cuh"

namespace a3d {

struct Xa6 {
    static constexpr int E = 60, H = 4, ROWS = 128, EXP_WARPS = 8, THREADS = 32 * (EXP_WARPS + 1);
    static constexpr int TILE_BYTES = Xa2::TILE_BYTES, STAGES = 3;
    static constexpr size_t X_BYTES = (size_t)ROWS * 64 * 4;          // residual stream, one 256-byte row per query
    static constexpr size_t Q_BYTES = (size_t)H * ROWS * 32;          // Q_h tiles [128][16] fp16, SWIZZLE_32B
    static constexpr size_t RING_BYTES = (size_t)STAGES * TILE_BYTES;
    static constexpr size_t MAIL_BYTES = 2 * 2 * ROWS * 4;            // row maxima exchanged by the two threads of a row
    static constexpr size_t SMEM = X_BYTES + Q_BYTES + RING_BYTES + MAIL_BYTES + 256;
    static constexpr int TMEM_COLS = 256, O_COL = 0, S_COL = 64, NBUF = 3;
    static constexpr int A_COL = S_COL, ACC_COL = S_COL + 64, COR_COL = S_COL + 128;
};

struct Xa6Bars {
    uint64_t kv_full[Xa6::STAGES], kv_empty[Xa6::STAGES];
    uint64_t s_full[Xa6::NBUF], p_full[Xa6::NBUF], pv_done[Xa6::H];
    uint64_t q_ready, o_full, verdict, a_ready, d_full;
    uint32_t tmem_base;
    uint32_t overflow_count;
};

// layers replayed in safe mode since the last reset: a3d_debug_counter("xattn_replays")
__device__ unsigned long long g_xa6_replays = 0;

#ifdef A3D_X6_TRACE
// clock64 timeline of one CTA (study build only: A3D_NVCC_EXTRA=-DA3D_X6_TRACE; read with a3d_x6_trace_read)
__device__ unsigned long long g_x6_trace[3][4096];
__device__ __forceinline__ void x6_trace(bool on, int role, int& n, int tag, int idx) {
    if (on && n < 4096) g_x6_trace[role][n++] = ((unsigned long long)tag << 56) | ((unsigned long long)(idx & 0xffff) << 40) | (clock64() & 0xffffffffffull);
}
#define X6_TRACE(on, role, n, tag, idx) x6_trace(on, role, n, tag, idx)
#else
#define X6_TRACE(on, role, n, tag, idx)
#endif

// ---- elect-predicated issue primitives: executed by all 32 lanes of the (converged) issuing warp, one lane acts
__device__ __forceinline__ void umma_ss_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ts_e(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit_e(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_e(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e mbarrier.arrive.shared::cta.b64 _, [%0];\n\t"
        "}" ::"r"(smem_u32(bar))
        : "memory");
}
// arm the barrier with the byte count and start the bulk copy (one lane)
__device__ __forceinline__ void bulk_load_e(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
        "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t"
        "}" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// one arrival for the whole (converged) warp: every lane's preceding writes are ordered before it by the warp barrier
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
    __syncwarp();
    mbar_arrive_e(bar);
}
__device__ __forceinline__ void bar_pair(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// 2^x for two scores on the FMA pipe, packed fp32 pairs (FADD2 / FFMA2): Cody-Waite split with the 1.5 * 2^23 trick +
// degree-3 minimax polynomial on [-0.5, 0.5] (max relative error 7.5e-5, far below the fp16 rounding of P); the
// integer part goes into the exponent field.  Scores below -126 are clamped (2^-126 rounds to 0 in fp16 anyway).
__device__ __forceinline__ unsigned long long pk64(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint32_t exp2_poly_pair(uint32_t s0, uint32_t s1) {
    const float x0 = fmaxf(__uint_as_float(s0), -126.0f), x1 = fmaxf(__uint_as_float(s1), -126.0f);
    const unsigned long long X = pk64(x0, x1);
    const unsigned long long MP = pk64(12582912.0f, 12582912.0f), MN = pk64(-12582912.0f, -12582912.0f);
    const unsigned long long NEG1 = pk64(-1.0f, -1.0f);
    const unsigned long long C3 = pk64(0.055170901f, 0.055170901f), C2 = pk64(0.24260952f, 0.24260952f);
    const unsigned long long C1 = pk64(0.69326097f, 0.69326097f), C0 = pk64(0.99992818f, 0.99992818f);
    unsigned long long T, R, F, P;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(T) : "l"(X), "l"(MP));               // rint(x) in the low mantissa bits
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(R) : "l"(T), "l"(MN));               // rint(x) as a float
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(F) : "l"(R), "l"(NEG1), "l"(X)); // x - rint(x) in [-0.5, 0.5]
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(P) : "l"(C3), "l"(F), "l"(C2));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(P) : "l"(P), "l"(F), "l"(C1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(P) : "l"(P), "l"(F), "l"(C0));
    uint32_t t0, t1, p0, p1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(t0), "=r"(t1) : "l"(T));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(P));
    return pack_h2(__uint_as_float(p0 + (t0 << 23)), __uint_as_float(p1 + (t1 << 23)));
}
__device__ __forceinline__ uint32_t exp2_mufu_pair(uint32_t s0, uint32_t s1) {
    return pack_h2(exp2_fast(__uint_as_float(s0)), exp2_fast(__uint_as_float(s1)));
}

// NP: how many of a thread's 16 score pairs per unit go through the FMA-pipe polynomial (0 = all on the MUFU unit)
template <int NP>
__global__ void __launch_bounds__(Xa6::THREADS, 2) xattn6_kernel(const Xa2Args a) {
    using C = Xa6;
    constexpr int E = C::E, H = C::H;
    extern __shared__ __align__(1024) unsigned char smem[];
    float* xpark = reinterpret_cast<float*>(smem);                                // [128][64] residual stream, chunk-swizzled
    unsigned char* qs = smem + C::X_BYTES;                                        // [H][128][32 B] fp16, SW32
    unsigned char* ring = smem + C::X_BYTES + C::Q_BYTES;                         // STAGES x TILE_BYTES
    float* mail = reinterpret_cast<float*>(ring + C::RING_BYTES);                 // [2 parities][2 halves][128 rows]
    Xa6Bars* bars = reinterpret_cast<Xa6Bars*>(ring + C::RING_BYTES + C::MAIL_BYTES);
    __shared__ float freq[E / 6];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, row0 = blockIdx.x * C::ROWS;
#ifdef A3D_X6_TRACE
    const unsigned long long t_start = clock64();
#endif

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(bars->kv_full + s, 1);
            mbar_init(bars->kv_empty + s, 1);
        }
        for (int i = 0; i < C::NBUF; ++i) {
            mbar_init(bars->s_full + i, 1);
            mbar_init(bars->p_full + i, C::EXP_WARPS);      // one arrival per warp
        }
        for (int h = 0; h < C::H; ++h) mbar_init(bars->pv_done + h, 1);
        mbar_init(&bars->q_ready, C::ROWS / 32);
        mbar_init(&bars->o_full, 1);
        mbar_init(&bars->verdict, C::ROWS / 32);
        mbar_init(&bars->a_ready, C::ROWS / 32);
        mbar_init(&bars->d_full, 1);
        bars->overflow_count = 0;
        mbar_fence_init();
    }
    if (tid < E / 6) freq[tid] = rope_freq<E>(tid);
    if (warp == C::EXP_WARPS) tmem_alloc(&bars->tmem_base, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_raw = bars->tmem_base;

    const int nt = a.ntiles;
    const unsigned char* kv_sample = a.kv_base + (size_t)b * nt * C::TILE_BYTES;

    if (warp == C::EXP_WARPS) {
        // =========================================================== issuing warp (all lanes, warp-uniform control flow)
        const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_raw, 0);      // provably warp-uniform: stays in uniform registers
        uint32_t gl = 0, gu = 0;              // ring items requested / consumed
        int ld_layer = 0, ld_pos = 0;         // next item to request: pos 0 = W_q, 1..nt = K/V tile pos-1, nt+1..nt+3 = W_o, W_1, W_2
        uint32_t gg = 0;                      // linear layers issued (4 per attention layer)
        uint32_t ps = 0;                      // passes over the keys (one per layer, plus one per safe-mode replay)
        uint32_t seen_overflows = 0;
#ifdef A3D_X6_TRACE
        const bool tr_on = (blockIdx.x == 37 && blockIdx.y == 5 && lane == 0);
        int tr_n = 0;
#endif
        const uint32_t ring_addr = smem_u32(ring), q_addr = smem_u32(qs);
        // descriptors differ only in the 14-bit start-address field (bytes >> 4): offsets are plain additions
        const uint64_t qdesc = sw32_desc(q_addr), kvdesc = sw32_desc(ring_addr);

        // Ring requests are decoupled from consumption: a freed slot is re-armed as soon as its kv_empty barrier has
        // completed (checked without blocking at every unit), so that the issuing warp never sits on the commit it
        // has just made.  Invariant while the stream lasts: gl + pending == gu + STAGES.
        int pending = C::STAGES;
        auto pump = [&](bool block) {
            while (pending > 0) {
                if (ld_layer >= a.nlayers) {
                    pending = 0;
                    break;
                }
                const uint32_t s = gl % C::STAGES, use = gl / C::STAGES;
                if (use >= 1) {
                    if (block) mbar_wait(bars->kv_empty + s, (use - 1) & 1);
                    else if (!mbar_test(bars->kv_empty + s, (use - 1) & 1)) break;
                }
                const unsigned char* src;
                if (ld_pos >= 1 && ld_pos <= nt)
                    src = kv_sample + (size_t)ld_layer * a.kv_layer_stride + (size_t)(ld_pos - 1) * C::TILE_BYTES;
                else
                    src = reinterpret_cast<const unsigned char*>(a.w + (size_t)ld_layer * Xa2::LAYER_W + Xa2::W_IMG +
                                                                 (size_t)(ld_pos == 0 ? 0 : ld_pos - nt) * Xa2::MAT);
                bulk_load_e(ring + s * C::TILE_BYTES, src, C::TILE_BYTES, bars->kv_full + s);
                ++gl;
                --pending;
                if (++ld_pos == nt + 4) {
                    ld_pos = 0;
                    ++ld_layer;
                }
            }
        };
        auto want_item = [&](uint32_t it) {          // item `it` must have been requested before its kv_full is waited on
            if (gl <= it) pump(true);
        };
        // one linear layer: D (main, correction) = A (hi, lo; tensor memory) x W image (ring item gu)
        auto gemm_step = [&]() {
            mbar_wait(&bars->a_ready, gg & 1);
            const uint32_t s = gu % C::STAGES;
            want_item(gu);
            mbar_wait(bars->kv_full + s, (gu / C::STAGES) & 1);
            tc_fence_after();
            const uint32_t w_addr = ring_addr + s * C::TILE_BYTES;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t bh = sw32_desc(w_addr + ks * 2048), bl = sw32_desc(w_addr + 8192 + ks * 2048);
                umma_ts_e(tmem + C::ACC_COL, tmem + C::A_COL + 8 * ks, bh, kIdescS, ks > 0 ? 1u : 0u);
                umma_ts_e(tmem + C::COR_COL, tmem + C::A_COL + 8 * ks, bl, kIdescS, ks > 0 ? 1u : 0u);
                umma_ts_e(tmem + C::COR_COL, tmem + C::A_COL + 32 + 8 * ks, bh, kIdescS, 1u);
            }
            tc_commit_e(&bars->d_full);
            tc_commit_e(bars->kv_empty + s);
            ++gu;
            ++gg;
            ++pending;
            pump(false);
        };

        pump(true);
        for (int layer = 0; layer < a.nlayers; ++layer) {
            gemm_step();                                    // q projection
            mbar_wait(&bars->q_ready, layer & 1);           // Q tiles of this layer are in shared memory
            tc_fence_after();
            for (int attempt = 0; attempt < 2; ++attempt, ++ps) {
                // Software pipeline over units (tile t, head h): S(u) is issued two units ahead of the PV product that
                // consumes P(u-2); S(u) may overwrite the buffer of unit u-3 without a wait because tcgen05.mma
                // instructions execute in issue order.  The loop is unrolled over the heads and every index (S / P
                // buffer, ring slot, barrier parity) is carried incrementally: this warp's instruction count per
                // unit bounds the whole kernel (a division-based version of this loop cost ~300 instructions and
                // ~1000 cycles per unit, more than the exponentials of the unit).
                const uint32_t ubase = ps * nt * H;
                uint32_t bi = ubase % C::NBUF;                              // S buffer of the next S product
                uint32_t pj = bi, pk = (ubase / C::NBUF) & 1;               // P buffer / p_full parity of the next PV product
                uint32_t ts = gu % C::STAGES, tp = (gu / C::STAGES) & 1;    // ring slot / kv_full parity of tile t
                uint32_t vs = ts;                                           // ring slot of tile t - 1
                uint32_t item = gu;
                for (int t = 0; t <= nt; ++t) {                             // the extra round drains the last two PV products
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        if (pending) pump(false);
                        if (t < nt) {
                            if (h == 0) {
                                want_item(item);
                                mbar_wait(bars->kv_full + ts, tp);
                                tc_fence_after();
                            }
                            umma_ss_e(tmem + C::S_COL + 64 * bi, qdesc + h * (4096 >> 4), kvdesc + ts * (C::TILE_BYTES >> 4) + h * (2048 >> 4),
                                      kIdescS, 0);
                            tc_commit_e(bars->s_full + bi);
                            X6_TRACE(tr_on, 0, tr_n, 1, 4 * t + h);
                            bi = (bi == C::NBUF - 1) ? 0 : bi + 1;
                        }
                        // PV product of unit u - 2: head hv of tile t (h >= 2) or of tile t - 1 (h < 2)
                        constexpr int kHvOf[4] = {2, 3, 0, 1};
                        const int hv = kHvOf[h];
                        const bool has_v = (h >= 2) ? (t < nt) : (t >= 1);
                        if (has_v) {
                            const uint32_t slot = (h >= 2) ? ts : vs;
                            const uint32_t first = (h >= 2) ? (t == 0) : (t == 1);     // first tile of the pass: overwrite O_h
                            X6_TRACE(tr_on, 0, tr_n, 2, 4 * t + h - 2);
                            mbar_wait(bars->p_full + pj, pk);
                            tc_fence_after();
                            X6_TRACE(tr_on, 0, tr_n, 3, 4 * t + h - 2);
                            const uint64_t vd = kvdesc + slot * (C::TILE_BYTES >> 4) + ((H * 2048 + hv * 2048) >> 4);
                            const uint32_t pb = tmem + C::S_COL + 64 * pj;
                            // P of keys 0..31 sits in columns 0..15 of the buffer, P of keys 32..63 in columns 32..47
                            umma_ts_e(tmem + C::O_COL + 16 * hv, pb, vd, kIdescPV, first ? 0u : 1u);
                            umma_ts_e(tmem + C::O_COL + 16 * hv, pb + 8, vd + (512 >> 4), kIdescPV, 1u);
                            umma_ts_e(tmem + C::O_COL + 16 * hv, pb + 32, vd + (1024 >> 4), kIdescPV, 1u);
                            umma_ts_e(tmem + C::O_COL + 16 * hv, pb + 40, vd + (1536 >> 4), kIdescPV, 1u);
                            tc_commit_e(bars->pv_done + hv);
                            if (hv == H - 1) {
                                tc_commit_e(bars->kv_empty + slot);
                                ++gu;
                                ++pending;
                            }
                            if (pj == C::NBUF - 1) {
                                pj = 0;
                                pk ^= 1;
                            } else {
                                ++pj;
                            }
                        }
                    }
                    vs = ts;
                    ++item;
                    if (ts == C::STAGES - 1) {
                        ts = 0;
                        tp ^= 1;
                    } else {
                        ++ts;
                    }
                }
                tc_commit_e(&bars->o_full);
                if (attempt == 1) continue;
                // verdict of the row threads on the fast pass: replay this layer in safe mode if any row overflowed
                mbar_wait(&bars->verdict, layer & 1);
                const uint32_t now = *reinterpret_cast<volatile uint32_t*>(&bars->overflow_count);
                const bool redo = now != seen_overflows;
                seen_overflows = now;
                if (!redo) {
                    ++ps;
                    break;
                }
                // the items requested past the last tile (W_o, W_1 of this layer) are dropped and the tiles streamed again
                while (gu != gl) {
                    const uint32_t s = gu % C::STAGES;
                    mbar_wait(bars->kv_full + s, (gu / C::STAGES) & 1);
                    mbar_arrive_e(bars->kv_empty + s);
                    ++gu;
                }
                ld_layer = layer;
                ld_pos = 1;
                pending = C::STAGES;
                pump(true);
                if (lane == 0) atomicAdd(&g_xa6_replays, 1ull);
            }
            gemm_step();                                    // out projection
            gemm_step();                                    // FFN 1
            gemm_step();                                    // FFN 2
        }
    } else {
        // =========================================================== exponential warps; warps 0-3 also own the rows
        const uint32_t tmem = tmem_raw;
        const int wq = warp & 3, half = warp >> 2;
        const int lrow = wq * 32 + lane;                              // row inside the tile == TMEM lane
        const int row = row0 + lrow;
        const bool owner = (half == 0);
        const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
        uint32_t gg = 0;                                              // linear layers consumed (row owners)
        uint32_t ps = 0, seen_overflows = 0;
        uint32_t xch = 0;                                             // pair-exchange counter (mailbox parity)
#ifdef A3D_X6_TRACE
        const bool tr_on = (blockIdx.x == 37 && blockIdx.y == 5 && lane == 0 && wq == 0);
        int tr_n = 0, tr_n2 = 3000;
        const int tr_role = 1 + half;
#endif

        auto xchunk = [&](int c) -> float4* {                         // 16-byte chunk c (0..15) of this thread's row
            return reinterpret_cast<float4*>(xpark + lrow * 64) + (c ^ (lrow & 15));
        };
        // columns [16 j, 16 j + 16) of the A operand of the next linear layer
        auto put_a16 = [&](int j, const float (&v)[16]) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split_h2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
            tmem_st8(lane_addr + C::A_COL + 8 * j, hi);
            tmem_st8(lane_addr + C::A_COL + 32 + 8 * j, lo);
        };
        auto a_done = [&]() {
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_warp(&bars->a_ready);
        };
        auto d_wait = [&]() {
            mbar_wait(&bars->d_full, gg & 1);
            tc_fence_after();
            ++gg;
        };
        // columns [16 j, 16 j + 16) of the result of the last linear layer
        auto get_d16 = [&](int j, float (&y)[16]) {
            uint32_t ac[16], co[16];
            tmem_ld16(lane_addr + C::ACC_COL + 16 * j, ac);
            tmem_ld16(lane_addr + C::COR_COL + 16 * j, co);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = fmaf(__uint_as_float(co[i]), kLoScaleInv, __uint_as_float(ac[i]));
        };
        auto ld4 = [&](const float* p, int c, float (&o)[16]) {       // 16 consecutive floats of a warp-uniform vector
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p + c) + i);
                o[4 * i] = v.x, o[4 * i + 1] = v.y, o[4 * i + 2] = v.z, o[4 * i + 3] = v.w;
            }
        };
        // y = D + bias + residual -> parked row; LayerNorm over the E valid columns in place (eps 1e-5)
        auto residual_ln = [&](const float* bias, const float* gamma, const float* beta) {
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float y[16], bb[16];
                get_d16(j, y);
                ld4(bias, 16 * j, bb);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 x = *xchunk(4 * j + i);
                    x.x += y[4 * i] + bb[4 * i];
                    x.y += y[4 * i + 1] + bb[4 * i + 1];
                    x.z += y[4 * i + 2] + bb[4 * i + 2];
                    x.w += y[4 * i + 3] + bb[4 * i + 3];
                    if (16 * j + 4 * i < E) sum += (x.x + x.y) + (x.z + x.w);
                    *xchunk(4 * j + i) = x;
                }
            }
            const float mean = sum * (1.0f / E);
            float var = 0.f;
#pragma unroll
            for (int c = 0; c < E / 4; ++c) {
                const float4 x = *xchunk(c);
                const float d0 = x.x - mean, d1 = x.y - mean, d2 = x.z - mean, d3 = x.w - mean;
                var = fmaf(d0, d0, var);
                var = fmaf(d1, d1, var);
                var = fmaf(d2, d2, var);
                var = fmaf(d3, d3, var);
            }
            const float rstd = 1.0f / sqrtf(var * (1.0f / E) + 1e-5f);
#pragma unroll
            for (int c = 0; c < E / 4; ++c) {
                float4 x = *xchunk(c);
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), be = __ldg(reinterpret_cast<const float4*>(beta) + c);
                x.x = (x.x - mean) * rstd * g.x + be.x;
                x.y = (x.y - mean) * rstd * g.y + be.y;
                x.z = (x.z - mean) * rstd * g.z + be.z;
                x.w = (x.w - mean) * rstd * g.w + be.w;
                *xchunk(c) = x;
            }
            *xchunk(15) = make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // parked row -> A operand
        auto row_to_a = [&]() {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 x = *xchunk(4 * j + i);
                    v[4 * i] = x.x, v[4 * i + 1] = x.y, v[4 * i + 2] = x.z, v[4 * i + 3] = x.w;
                }
                put_a16(j, v);
            }
            a_done();
        };

        float qxyz[3] = {0.f, 0.f, 0.f};
        if (owner) {
            // ---- residual stream -> parked row; A operand of the first q projection
            const float* xp = a.x0 + (long)b * a.x0_sb + (long)row * a.x0_sn;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < a.nq && c < E / 4) {
                    const float2 lo = __ldg(reinterpret_cast<const float2*>(xp + 4 * c)), hi = __ldg(reinterpret_cast<const float2*>(xp + 4 * c + 2));
                    v = make_float4(lo.x, lo.y, hi.x, hi.y);
                }
                *xchunk(c) = v;
            }
            if (a.qpos && row < a.nq)
                for (int ax = 0; ax < 3; ++ax) qxyz[ax] = __ldg(a.qpos + ((long)b * a.nq + row) * 3 + ax);
            row_to_a();
        }

        const int q_swz = (lrow >> 2) & 1;
        unsigned char* q_row = qs + lrow * 32;                                   // this row in head 0's tile
        unsigned char* q_pad = q_row + ((1 ^ q_swz) << 4) + 14;                  // slot 15

        for (int layer = 0; layer < a.nlayers; ++layer) {
            const float* vv = a.v + (size_t)layer * Xa2::LAYER_V;
            X6_TRACE(tr_on, 1, tr_n2, 10, layer);
            if (owner) {
                // ------------------------------------------------------------ Q = rotary(x Wq^T + bq) -> smem (SW32 tiles per head)
                d_wait();
                X6_TRACE(tr_on, 1, tr_n2, 11, layer);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float y[16], bb[16];
                    get_d16(j, y);
                    ld4(vv + Xa2::B_Q, 16 * j, bb);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int c = 16 * j + 2 * i;
                        float v0 = y[2 * i] + bb[2 * i], v1 = y[2 * i + 1] + bb[2 * i + 1];
                        if (c < E) {
                            if (a.qpos) {
                                const int pi = c >> 1, axis = pi / (E / 6), jf = pi - axis * (E / 6);
                                const float ang = qxyz[axis] * freq[jf];
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
                            auto put = [&](int cc, float val) {
                                const int hh = cc / 15, d = cc % 15;
                                *reinterpret_cast<__half*>(q_row + hh * 4096 + (((d >> 3) ^ q_swz) << 4) + (d & 7) * 2) = __float2half_rn(val);
                            };
                            put(c, v0);
                            put(c + 1, v1);
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < H; ++h) *reinterpret_cast<__half*>(q_pad + h * 4096) = __float2half_rn(0.f);
                fence_async_smem();                 // generic-proxy writes of Q -> visible to the tensor-core (async) proxy
                tc_fence_before();
                mbar_arrive_warp(&bars->q_ready);
                X6_TRACE(tr_on, 1, tr_n2, 12, layer);
            }

            // ---------------------------------------------------------------- softmax over the keys (all 8 warps)
            constexpr float kMargin = 6.0f;
            float sh[H];
            // maximum over both halves of the row: through the mailbox, double-buffered by the exchange counter
            auto pair_max = [&](float m) -> float {
                float* slot = mail + (xch & 1) * 2 * C::ROWS;
                slot[half * C::ROWS + lrow] = m;
                bar_pair(1 + wq);
                const float o = slot[(half ^ 1) * C::ROWS + lrow];
                ++xch;
                return fmaxf(m, o);
            };
            // explicit path of one unit (first tile of a pass; every tile in safe mode): row maximum, new shift, rescale
            // of O_h when the shift moved, P with masking of the keys past nk.  r = this thread's 32 scores.
            auto slow_unit = [&](int t, int h, int valid, const uint32_t (&r)[32], uint32_t (&p)[16]) {
                float mx = -INFINITY;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (c < valid) mx = fmaxf(mx, __uint_as_float(r[c]));
                mx = pair_max(mx);
                float shift_new, delta;
                if (t == 0) {
                    shift_new = -__half2float(__float2half_rn(-(mx + kMargin)));
                    delta = shift_new;                                 // S of the first tile is unshifted
                } else {
                    shift_new = (mx >= 1.0f) ? -__half2float(__float2half_rn(-(sh[h] + mx + kMargin))) : sh[h];
                    delta = shift_new - sh[h];                         // exact: both are fp16 values
                    if (owner && __any_sync(0xffffffffu, delta != 0.f)) {
                        // O_h holds sums relative to the old shift: rescale this row once the PV product of the
                        // previous tile of this head has landed (completion number ps*nt + t of pv_done[h])
                        mbar_wait(bars->pv_done + h, (ps * nt + t - 1) & 1);
                        tc_fence_after();
                        uint32_t o[16];
                        tmem_ld16(lane_addr + C::O_COL + 16 * h, o);
                        tmem_wait_ld();
                        const float sc = exp2_fast(-delta);           // 1 for rows that keep their shift
#pragma unroll
                        for (int d = 0; d < 16; ++d) o[d] = __float_as_uint(__uint_as_float(o[d]) * sc);
                        tmem_st16(lane_addr + C::O_COL + 16 * h, o);
                    }
                }
                sh[h] = shift_new;
                if (owner) {
                    *reinterpret_cast<__half*>(q_pad + h * 4096) = __float2half_rn(-shift_new);
                    fence_async_smem();                                // visible to the S products of later tiles
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float e0 = (2 * c < valid) ? exp2_fast(__uint_as_float(r[2 * c]) - delta) : 0.f;
                    const float e1 = (2 * c + 1 < valid) ? exp2_fast(__uint_as_float(r[2 * c + 1]) - delta) : 0.f;
                    p[c] = pack_h2(e0, e1);
                }
            };
            auto explicit_tiles = [&](int t0, int t1) {
                for (int t = t0; t < t1; ++t) {
                    const int vt = min(kTileKeys, a.nk - t * kTileKeys);
                    const int valid = max(0, min(32, vt - 32 * half));
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        const uint32_t U = (ps * nt + t) * H + h, i = U % C::NBUF, k = U / C::NBUF;
                        const uint32_t sb = lane_addr + C::S_COL + 64 * i + 32 * half;
                        mbar_wait(bars->s_full + i, k & 1);
                        tc_fence_after();
                        uint32_t r[32], p[16];
                        tmem_ld32(sb, r);
                        tmem_wait_ld();
                        slow_unit(t, h, valid, r, p);
                        tmem_st16(sb, p);
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive_warp(bars->p_full + i);
                    }
                }
            };

            bool redo = false;
            for (int attempt = 0; attempt < 2; ++attempt, ++ps) {
#pragma unroll
                for (int h = 0; h < H; ++h) sh[h] = 0.f;
                if (attempt == 1) {
                    explicit_tiles(0, nt);             // safe mode: every unit tracks the exact running maximum
                } else {
                    explicit_tiles(0, 1);              // the first tile fixes the shift of every row
                    X6_TRACE(tr_on, 1, tr_n2, 13, layer);
                    // ---- remaining tiles, software-pipelined inside the warp in chunks of 16 scores: chunk B of a unit is
                    //      in flight while chunk A is exponentiated, chunk A of the NEXT unit while chunk B is; the probe of
                    //      the next unit's s_full barrier is issued before chunk A's arithmetic and consumed after it, and
                    //      the hand-over of P(u-1) (tcgen05.wait::st + arrive) sits behind chunk A as well -- nothing with a
                    //      latency is waited on right after it was started.
                    if (nt > 1) {
                        constexpr int NPA = (NP + 1) / 2, NPB = NP / 2;       // polynomial pairs of chunk A / chunk B
                        uint32_t ca[16], cb[16], p[16];
                        uint32_t U = (ps * nt + 1) * H;                         // current unit
                        uint32_t i = U % C::NBUF, k = (U / C::NBUF) & 1;
                        const uint32_t s0 = lane_addr + C::S_COL + 32 * half;
                        mbar_wait(bars->s_full + i, k);
                        tc_fence_after();
                        tmem_ld16(s0 + 64 * i, ca);
                        tmem_wait_ld();
                        int pend = -1;                 // buffer whose P hand-over is still owed (its tcgen05.st in flight)
                        const int units = (nt - 1) * H;
                        for (int u = 0; u < units; ++u) {
                            const uint32_t sb = s0 + 64 * i;
                            const uint32_t in = (i == C::NBUF - 1) ? 0 : i + 1, kn = (i == C::NBUF - 1) ? (k ^ 1) : k;
                            const bool has_next = u + 1 < units;
                            X6_TRACE(tr_on, tr_role, tr_n, 4, u + 4);
                            tmem_ld16(sb + 16, cb);
                            const bool ready = has_next ? mbar_test(bars->s_full + in, kn) : true;
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                p[c] = (c < NPA) ? exp2_poly_pair(ca[2 * c], ca[2 * c + 1]) : exp2_mufu_pair(ca[2 * c], ca[2 * c + 1]);
                            if (pend >= 0) {                   // P of the previous unit has landed by now
                                tmem_wait_st();
                                tc_fence_before();
                                mbar_arrive_warp(bars->p_full + pend);
                            }
                            tmem_wait_ld();
                            X6_TRACE(tr_on, tr_role, tr_n, 5, u + 4);
                            if (has_next) {
                                if (!ready) mbar_wait(bars->s_full + in, kn);
                                tc_fence_after();
                                tmem_ld16(s0 + 64 * in, ca);
                            }
                            X6_TRACE(tr_on, tr_role, tr_n, 6, u + 4);
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                p[8 + c] = (c < NPB) ? exp2_poly_pair(cb[2 * c], cb[2 * c + 1]) : exp2_mufu_pair(cb[2 * c], cb[2 * c + 1]);
                            tmem_st16(sb, p);
                            X6_TRACE(tr_on, tr_role, tr_n, 7, u + 4);
                            pend = (int)i;
                            if (has_next) tmem_wait_ld();
                            i = in;
                            k = kn;
                        }
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive_warp(bars->p_full + pend);
                    }
                }
                if (owner) {
                    // the pad slots go back to 0 (an unshifted first tile) for a possible replay / the next layer
#pragma unroll
                    for (int h = 0; h < H; ++h) *reinterpret_cast<__half*>(q_pad + h * 4096) = __float2half_rn(0.f);
                    fence_async_smem();
                }
                X6_TRACE(tr_on, 1, tr_n2, 14, layer);
                if (attempt == 1) continue;            // (the loop increment counts the replay pass)
                // verdict on the fast pass: a non-finite denominator anywhere in the CTA -> replay the layer in safe mode
                if (owner) {
                    mbar_wait(&bars->o_full, ps & 1);
                    tc_fence_after();
                    bool bad = false;
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        uint32_t o[16];
                        tmem_ld16(lane_addr + C::O_COL + 16 * h, o);
                        tmem_wait_ld();
                        bad = bad || !isfinite(__uint_as_float(o[15]));
                    }
                    tc_fence_before();
                    if (bad) atomicAdd(&bars->overflow_count, 1u);
                    __threadfence_block();
                    mbar_arrive_warp(&bars->verdict);
                }
                mbar_wait(&bars->verdict, layer & 1);
                const uint32_t now = *reinterpret_cast<volatile uint32_t*>(&bars->overflow_count);
                redo = now != seen_overflows;
                seen_overflows = now;
                if (!redo) {
                    ++ps;
                    break;
                }
            }
            X6_TRACE(tr_on, 1, tr_n2, 15, layer);
            if (!owner) continue;

            // ---------------------------------------------------------------- O / l -> A operand of the out projection
            if (redo) {                                 // the replay pass has its own completion of o_full
                mbar_wait(&bars->o_full, (ps - 1) & 1);
                tc_fence_after();
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                uint32_t o[16];
                float v[16];
                tmem_ld16(lane_addr + C::O_COL + 16 * h, o);
                tmem_wait_ld();
                const float inv = 1.0f / __uint_as_float(o[15]);
#pragma unroll
                for (int d = 0; d < 15; ++d) v[d] = __uint_as_float(o[d]) * inv;
                v[15] = 0.f;
                put_a16(h, v);
            }
            a_done();
            X6_TRACE(tr_on, 1, tr_n2, 16, layer);
            // ---------------------------------------------------------------- x = LN(x + O Wo^T + bo)
            d_wait();
            X6_TRACE(tr_on, 1, tr_n2, 17, layer);
            residual_ln(vv + Xa2::B_O, vv + Xa2::G_1, vv + Xa2::BE_1);
            row_to_a();
            X6_TRACE(tr_on, 1, tr_n2, 18, layer);
            // ---------------------------------------------------------------- hid = relu(x W1^T + b1)
            d_wait();
            X6_TRACE(tr_on, 1, tr_n2, 19, layer);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float y[16], bb[16];
                get_d16(j, y);
                ld4(vv + Xa2::B_1, 16 * j, bb);
#pragma unroll
                for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i] + bb[i], 0.f);
                put_a16(j, y);
            }
            a_done();
            X6_TRACE(tr_on, 1, tr_n2, 20, layer);
            // ---------------------------------------------------------------- x = LN(x + hid W2^T + b2)
            d_wait();
            X6_TRACE(tr_on, 1, tr_n2, 21, layer);
            residual_ln(vv + Xa2::B_2, vv + Xa2::G_2, vv + Xa2::BE_2);
            const bool last = (layer == a.nlayers - 1);
            if (!last) row_to_a();                      // A operand of the next layer's q projection
            X6_TRACE(tr_on, 1, tr_n2, 22, layer);
            // ---- outputs of this layer
            if (row < a.nq) {
                if (a.feat_out && (a.feat_all || last)) {
                    float* fo = a.feat_out + (((size_t)(a.feat_all ? layer : 0) * a.batch + b) * (size_t)a.feat_rows + row) * E;
#pragma unroll
                    for (int c = 0; c < E / 4; ++c) *reinterpret_cast<float4*>(fo + 4 * c) = *xchunk(c);
                }
                if (last && a.logits) {
                    for (int j = 0; j < a.nqv; ++j) {
                        const float4* qv = reinterpret_cast<const float4*>(a.qvec + ((size_t)j * a.batch + b) * E);
                        float acc = 0.f;
#pragma unroll
                        for (int c = 0; c < E / 4; ++c) {
                            const float4 x = *xchunk(c), q = __ldg(qv + c);
                            acc = fmaf(x.x, q.x, acc);
                            acc = fmaf(x.y, q.y, acc);
                            acc = fmaf(x.z, q.z, acc);
                            acc = fmaf(x.w, q.w, acc);
                        }
                        a.logits[((size_t)j * a.batch + b) * a.nq + row] = acc;
                    }
                }
            }
        }
    }
    // ---- teardown
#ifdef A3D_X6_TRACE
    if (blockIdx.x == 37 && blockIdx.y == 5 && threadIdx.x == 0) g_x6_trace[1][4000] = (23ull << 56) | (clock64() & 0xffffffffffull);
    if (blockIdx.x == 37 && blockIdx.y == 5 && threadIdx.x == 0) g_x6_trace[1][4001] = (9ull << 56) | (t_start & 0xffffffffffull);
#endif
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == C::EXP_WARPS) tmem_dealloc(tmem_raw, C::TMEM_COLS);
}

}  // namespace a3d

using namespace a3d;

int a3d_xattn6_replays(unsigned long long* value, int reset) {
    const unsigned long long zero = 0;
    if (cudaMemcpyFromSymbol(value, g_xa6_replays, sizeof(*value)) != cudaSuccess) return A3D_ECUDA;
    if (reset && cudaMemcpyToSymbol(g_xa6_replays, &zero, sizeof(zero)) != cudaSuccess) return A3D_ECUDA;
    return A3D_OK;
}

#ifdef A3D_X6_TRACE
extern "C" int a3d_x6_trace_read(unsigned long long* host) {
    return cudaMemcpyFromSymbol(host, g_x6_trace, sizeof(g_x6_trace)) == cudaSuccess ? 0 : -1;
}
#endif

template <int NP>
static int launch_np(const Xa2Args& a, dim3 grid, cudaStream_t stream) {
    static PerDeviceOnce once_dev;
    if (bool& once = once_dev.flag(); !once) {
        cudaError_t e = cudaFuncSetAttribute(xattn6_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Xa6::SMEM);
        if (e != cudaSuccess) {
            set_error("a3d_xattn_stack(tcgen05 v6): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return A3D_ECUDA;
        }
        once = true;
    }
    xattn6_kernel<NP><<<grid, Xa6::THREADS, Xa6::SMEM, stream>>>(a);
    return check_launch("a3d_xattn_stack(tcgen05 v6)");
}

// launched by a3d_xattn_stack (a3d_xattn2.cu); poly = score pairs (of 16 per thread and unit) on the FMA-pipe polynomial
int a3d_launch_xattn6(const Xa2Args& a, dim3 grid, cudaStream_t stream, int poly) {
    switch (poly) {
        case 0: return launch_np<0>(a, grid, stream);
        case 4: return launch_np<4>(a, grid, stream);
        case 6: return launch_np<6>(a, grid, stream);
        case 7: return launch_np<7>(a, grid, stream);
        case 9: return launch_np<9>(a, grid, stream);
        case 10: return launch_np<10>(a, grid, stream);
        default: return launch_np<8>(a, grid, stream);
    }
}
