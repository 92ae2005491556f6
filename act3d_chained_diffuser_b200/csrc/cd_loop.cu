// ChainedDiffuser sampling loop as ONE persistent kernel (diffusion_model.py:98-117 around diffusion_head.py:200-363):
// all denoising steps x all adaLN layers of one sample run inside a thread-block cluster of 4 CTAs that never leaves
// the SMs; the trajectory, the token tile and every intermediate stay in shared memory for the whole loop.
//
// Why this shape.  The launch-per-layer version (cd_denoiser.cu: cd_step_begin + 8 x (cd_cross, cd_post) per step,
// replayed from a CUDA graph) spends 68 % of the loop in cd_post: one CTA per sample (32 CTAs on 148 SMs) walking a
// chain of ~16 small GEMMs per layer whose weights stream through a shared-memory ring with a block barrier per k step.
// Here a sample is split BY ROWS over the 4 CTAs of a cluster (16 waypoint rows each = one m16 MMA tile):
//   * every linear layer, LayerNorm, adaLN, rotary and the cross-attention are row-local: a warp owns one 8-column
//     n tile of the [16 x 128] output, takes its weight fragments STRAIGHT from L2 with all k steps in flight (no
//     staging, no barrier inside a GEMM) and the stage costs one block barrier;
//   * the cross-attention is split BY KEYS instead: every CTA streams its own quarter of the context K/V tile images
//     (32 KiB per 64 keys, all 8 heads; cp.async.bulk on a 3-slot mbarrier ring fed by its own producer warp; an optional
//     cp.async.bulk.prefetch.L2 running n tiles ahead -- a3d_set_option("cd_prefetch_tiles", n) -- measured as a loss:
//     62.5 ms per C3 batch without it, 63.2 / 64.0 / 64.4 ms at 3 / 8 / 12 tiles, and at 12 tiles ncu shows 95 GB of
//     DRAM reads for a 54.5 GB stream: the 128 x 12 x 32 KiB of prefetched tiles evict each other before they are
//     used) against the Q of ALL 64 rows, and the unnormalised partial results
//     go to the CTA that owns the rows.  (A first version kept the rows split here too and multicast every tile to the
//     4 CTAs: 96 KiB in flight per cluster against ~2 us of HBM + multicast latency made the ring latency-bound, 54 us per
//     layer, ncu: 20 % of all stall samples on the tile barrier; the key split puts 4 x 96 KiB in flight per sample.)
//   * three exchanges per layer go through distributed shared memory, each fenced by one cluster-scope mbarrier: the
//     rotary Q rows (all-gather), the cross-attention partials (to the row owners) and the 16 rows of K / V of the
//     self-attention (all-gather into the fp16 hi / lo planes of all 4 CTAs).
// Math is the same as cd_denoiser.cu (error-compensated split-fp16 mma.sync GEMMs, exact softmax in the 64-key
// self-attention, online softmax in the cross-attention, fp32 everywhere else): the 16-row tiles leave no room for a
// 64/128-row tcgen05.mma, and at these sizes the stage latency, not tensor throughput, is the bound (the tcgen05
// round trip of a3d_xattn6.cu's linear layers -- tcgen05.st, commit, mbarrier, tcgen05.ld -- costs 3-10 k cycles per
// stage, a whole mma.sync stage here ~1 k).
#include "a3d_mma_gemm.cuh"
#include "cd_blocks.cuh"

namespace a3d {
namespace cd {

constexpr int CL = 4;                     // CTAs per cluster = row blocks of one sample
constexpr int RB = 16;                    // rows per CTA
constexpr int CW = 16;                    // compute warps
constexpr int CT = CW * 32;               // compute threads
constexpr int LT = CT + 32;               // + producer warp
constexpr int LP = 136;                   // halfs per row of an fp16 plane (272 B: conflict-free ldmatrix)
constexpr int FPT = 132;                  // floats per row of an fp32 tile
constexpr int PF_TILES = 0;               // default L2 prefetch distance of the K/V stream (0 = off; measured best, see header)
}  // namespace cd
int g_cd_prefetch_tiles = cd::PF_TILES;   // a3d_set_option("cd_prefetch_tiles", n)
namespace cd {
constexpr int KV_TILE = 2 * H * 2048;     // K image + V image of 64 keys, all heads
constexpr int ST = 3;                     // ring slots
constexpr int MAXL = 12;

#ifdef A3D_CDL_TRACE
// clock64 timeline of one layer of one CTA (study build only: A3D_NVCC_EXTRA=-DA3D_CDL_TRACE; cd_loop_trace_read)
__device__ long long g_cdl_trace[64];
#define CDL_TRACE(k) do { if (tr_on) g_cdl_trace[k] = clock64(); } while (0)
#else
#define CDL_TRACE(k)
#endif

struct LoopArgs {
    int batch, nrows, n_steps, nl, n_traj, nk, ntiles, n_instr, pf_tiles;
    float* traj;                          // [B][L][9] in: x_T (+ conditioning), out: x_0
    const float* cond;                    // [B][L][9]
    const unsigned char* cmask;           // [B][L][9]
    const unsigned char* mask;            // [B][L] key padding mask or null
    const float* wp_pe;                   // [L][E]
    const int* timesteps;                 // [n_steps] timestep of every loop iteration (device)
    const float* ada;                     // [T][nl][ADA_ROW]
    const float* coef;                    // [T][6] {c_x0, c_xt, sigma} for positions, then rotations (device)
    const float* noise_pos;               // [n_steps][B][L][3]
    const float* noise_rot;               // [n_steps][B][L][6]
    const float* enc1;                    // trajectory encoder layer 1: fp32 K-major [9][EP] + bias [EP]
    const uint4* enc2;
    const float* enc2_b;
    const uint4* lang_w;                  // LangW or null (use_instruction = 0)
    const float* lang_v;
    const float* lang_k;                  // [B][n_instr][E]
    const float* lang_vv;
    const uint4* ada_w[MAXL];
    const float* ada_v[MAXL];
    const uint4* pos_w;
    const float* pos_v;
    const uint4* rot_w;
    const float* rot_v;
    const unsigned char* kv;              // [nl] sets of [B][ntiles] tile images
    size_t kv_set_bytes;
};

struct LoopSmem {
    unsigned char* ring;                  // ST x KV_TILE
    __half *kh, *kl, *vh, *vl;            // [64][LP] self-attention / instruction K, V planes (head-padded columns)
    __half *ah, *al, *hh, *hl;            // [16][LP] GEMM input planes
    float* t1;                            // [16][FPT] GEMM output handed to the row-wise stages
    float* vecs;                          // AdaV of the current layer (biases, LayerNorm parameters): staged by a bulk copy
    float* qv;                            // [2][QV] {C_BQ, adaLN row} of the current / next layer (double-buffered)
    __half* qx;                           // [H][64][16] rotary Q of ALL rows for the next cross-attention (SWIZZLE_32B rows)
    float* part;                          // [CL sources][H][17][16] cross-attention partials of this CTA's rows (aliases the K/V planes)
    float *trj, *upd;                     // [16][9] trajectory rows; [16][9] position update (3) + rotation (6)
    float* freq;                          // [20]
    unsigned char* kmask;                 // [64]
    uint64_t *full, *empty, *kv_free, *kv_ready, *q_ready, *p_ready, *vec_full, *qv_full;
    static constexpr int QV = EP + ADA_ROW;          // floats of one {C_BQ, adaLN row} block
    static constexpr size_t BYTES = (size_t)ST * KV_TILE + (size_t)4 * 64 * LP * 2 + (size_t)4 * RB * LP * 2 + (size_t)RB * FPT * 4 +
                                    (size_t)(AdaV::SIZE + 2 * QV) * 4 + (size_t)H * 64 * 16 * 2 + 2 * RB * 9 * 4 + 32 * 4 + 64 +
                                    (2 * ST + 7) * 8 + 64;
    static_assert((size_t)CL * H * 17 * RB * 4 <= (size_t)4 * 64 * LP * 2, "partials must fit in the K/V planes");
    __device__ explicit LoopSmem(unsigned char* b) {
        ring = b;
        kh = reinterpret_cast<__half*>(b + (size_t)ST * KV_TILE);
        kl = kh + 64 * LP;
        vh = kl + 64 * LP;
        vl = vh + 64 * LP;
        ah = vl + 64 * LP;
        al = ah + RB * LP;
        hh = al + RB * LP;
        hl = hh + RB * LP;
        t1 = reinterpret_cast<float*>(hl + RB * LP);
        vecs = t1 + RB * FPT;
        qv = vecs + AdaV::SIZE;
        qx = reinterpret_cast<__half*>(qv + 2 * QV);
        part = reinterpret_cast<float*>(kh);
        trj = reinterpret_cast<float*>(qx + H * 64 * 16);
        upd = trj + RB * 9;
        freq = upd + RB * 9;
        kmask = reinterpret_cast<unsigned char*>(freq + 32);
        full = reinterpret_cast<uint64_t*>(kmask + 64);
        empty = full + ST;
        kv_free = empty + ST;
        kv_ready = kv_free + 1;
        q_ready = kv_ready + 1;
        p_ready = q_ready + 1;
        vec_full = p_ready + 1;
        qv_full = vec_full + 1;          // [2]
    }
};

// ---- cluster / distributed-shared-memory primitives
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, const uint4& v) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t raddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP_C:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_C;\n\t"
        "bra WAIT_LOOP_C;\n\t"
        "DONE_C:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t raddr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}
// warm L2 for a later bulk copy (TMA engine, no destination)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
// every thread that wrote into a peer's shared memory orders those stores at cluster scope BEFORE the block barrier after
// which one thread announces them with a release-arrive (a release by that one thread does not cover the other warps'
// in-flight remote stores: seen as a rare run-to-run difference before this fence was added)
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory"); }   // the 16 compute warps

__device__ __forceinline__ int head_slot16(int c) { return c + c / HD; }

// all weight fragments of one n tile of a [K = 16 KS][N] matrix: issued together, one L2 round trip per GEMM
template <int KS>
struct WFrag {
    uint4 b[KS];
    __device__ __forceinline__ void load(const uint4* __restrict__ wfrag, int ntiles_total, int nt, int lane) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) b[ks] = __ldg(wfrag + ((size_t)ks * ntiles_total + nt) * 32 + lane);
    }
};
// acc[4] = A[16 x 16 KS] (hi, lo planes, pitch LP) x W^T[:, n tile]: rows g, g+8; cols 8 nt + 2 q4 + {0, 1}
template <int KS>
__device__ __forceinline__ void gemm16(const __half* __restrict__ ah, const __half* __restrict__ al, const WFrag<KS>& w, int lane,
                                       float (&acc)[4]) {
    // six independent accumulator chains (main / hi*lo / lo*hi, even / odd k steps): a single chain of dependent HMMAs
    // (24 for K = 128) would cost ~30 cycles each with nothing else to issue in this warp
    float m0[4] = {0.f, 0.f, 0.f, 0.f}, m1[4] = {0.f, 0.f, 0.f, 0.f}, c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f},
          d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
    const int arow = (lane & 7) + 8 * ((lane >> 3) & 1), acol = 8 * (lane >> 4);
    const uint32_t ahb = smem_u32(ah + arow * LP + acol), alb = smem_u32(al + arow * LP + acol);
#pragma unroll
    for (int ks = 0; ks < KS; ks += 2) {
        uint32_t fh0[4], fl0[4], fh1[4], fl1[4];
        ldmatrix_x4(fh0, ahb + ks * 32);
        ldmatrix_x4(fl0, alb + ks * 32);
        ldmatrix_x4(fh1, ahb + (ks + 1) * 32);
        ldmatrix_x4(fl1, alb + (ks + 1) * 32);
        mma_16816(m0, fh0, w.b[ks].x, w.b[ks].y);
        mma_16816(c0, fh0, w.b[ks].z, w.b[ks].w);
        mma_16816(d0, fl0, w.b[ks].x, w.b[ks].y);
        mma_16816(m1, fh1, w.b[ks + 1].x, w.b[ks + 1].y);
        mma_16816(c1, fh1, w.b[ks + 1].z, w.b[ks + 1].w);
        mma_16816(d1, fl1, w.b[ks + 1].x, w.b[ks + 1].y);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] = fmaf((c0[e] + c1[e]) + (d0[e] + d1[e]), kLoScaleInv, m0[e] + m1[e]);
}

// self / instruction attention of this CTA's 16 rows: warp = head; q planes [16][LP] head-padded (carry hd^-1/2 log2 e),
// k / v planes [64][LP] head-padded; exact softmax over the <= 64 keys; output planes in natural channel order.
__device__ __forceinline__ void mha16(int h, int lane, const __half* qh, const __half* ql, const __half* kh, const __half* kl,
                                      const __half* vh, const __half* vl, int nk, const unsigned char* key_mask, __half* oh,
                                      __half* ol) {
    const int g = lane >> 2, q4 = lane & 3;
    uint32_t dead = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int key = 8 * j + 2 * q4 + e;
            if (key >= nk || (key_mask && key_mask[key])) dead |= 1u << (2 * j + e);
        }
    uint32_t fqh[4], fql[4];
    {
        const int row = (lane & 7) + 8 * ((lane >> 3) & 1);
        const int off = row * LP + 16 * h + 8 * (lane >> 4);
        ldmatrix_x4(fqh, smem_u32(qh + off));
        ldmatrix_x4(fql, smem_u32(ql + off));
    }
    float sc[8][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int key = kk * 16 + (lane & 7) + 8 * (lane >> 4);
        const int off = key * LP + 16 * h + 8 * ((lane >> 3) & 1);
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
        const int off = key * LP + 16 * h + 8 * (lane >> 4);
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
                const int row = g + 8 * (e >> 1);
                __half a, b;
                split_h(fmaf(oc[n][e], kLoScaleInv, o[n][e]) * ((e >> 1) ? i1 : i0), a, b);
                oh[row * LP + h * HD + d] = a;
                ol[row * LP + h * HD + d] = b;
            }
        }
}

__global__ void __launch_bounds__(LT, 1) cd_loop_kernel(const LoopArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    LoopSmem s(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q4 = lane & 3;
    const uint32_t rank = cluster_rank();
    const int b = blockIdx.x / CL;
    const int row0 = RB * (int)rank;                                  // first global row of this CTA
    const int nloc = max(0, min(RB, a.nrows - row0));                 // valid rows here
    const int ntiles = a.ntiles;
    // this CTA's quarter of the key tiles
    const int tiles_per = (ntiles + CL - 1) / CL;
    const int t_begin = min((int)rank * tiles_per, ntiles), my_tiles = min(t_begin + tiles_per, ntiles) - t_begin;

    // ---------------------------------------------------------------- one-time setup
    if (tid == 0) {
        for (int i = 0; i < ST; ++i) {
            mbar_init(s.full + i, 1);
            mbar_init(s.empty + i, CW);
        }
        mbar_init(s.kv_free, CL);
        mbar_init(s.kv_ready, CL);
        mbar_init(s.q_ready, CL);
        mbar_init(s.p_ready, CL);
        mbar_init(s.vec_full, 1);
        mbar_init(s.qv_full, 1);
        mbar_init(s.qv_full + 1, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < (4 * 64 + 4 * RB) * LP / 2; i += LT) reinterpret_cast<uint32_t*>(s.kh)[i] = 0u;   // all planes (pad columns stay 0)
    if (tid < E / 6) s.freq[tid] = rope_freq<E>(tid);
    if (tid < 64) s.kmask[tid] = (a.mask && tid < a.nrows) ? a.mask[(size_t)b * a.nrows + tid] : 0;
    if (tid < RB * 9) {
        const int r = tid / 9;
        s.trj[tid] = (r < nloc) ? a.traj[((size_t)b * a.nrows + row0) * 9 + tid] : 0.f;
    }
    __syncthreads();
    cluster_sync_all();                                               // barriers of every CTA are initialised and armed

    if (w == CW) {
        // ================================================================ producer warp: this CTA's K/V tile stream
        if (lane == 0 && my_tiles > 0) {
            const unsigned char* kv_b = a.kv + ((size_t)b * ntiles + t_begin) * KV_TILE;
            const uint32_t total = (uint32_t)a.n_steps * a.nl * my_tiles;
            auto src_of = [&](uint32_t j) {                       // j-th tile of the sequence (steps x layers x my tiles)
                const uint32_t l = (j / my_tiles) % a.nl, t = j % my_tiles;
                return kv_b + (size_t)l * a.kv_set_bytes + (size_t)t * KV_TILE;
            };
            const uint32_t pf = (uint32_t)a.pf_tiles;          // 0 = no L2 prefetch
            for (uint32_t j = 0; j < pf && j < total; ++j) bulk_prefetch_l2(src_of(j), KV_TILE);
            for (uint32_t j = 0; j < total; ++j) {
                if (pf && j + pf < total) bulk_prefetch_l2(src_of(j + pf), KV_TILE);
                const uint32_t slot = j % ST, use = j / ST;
                if (use >= 1) mbar_wait(s.empty + slot, (use - 1) & 1);
                mbar_expect_tx(s.full + slot, KV_TILE);
                bulk_g2s(s.ring + slot * KV_TILE, src_of(j), KV_TILE, s.full + slot);
            }
        }
    } else {
        // ================================================================ compute warps
        uint32_t kvf_remote[CL], kvr_remote[CL], kvp_remote[CL], qr_remote[CL], pr_remote[CL], qx_remote[CL];
#pragma unroll
        for (int r = 0; r < CL; ++r) {
            kvf_remote[r] = mapa(smem_u32(s.kv_free), r);
            kvr_remote[r] = mapa(smem_u32(s.kv_ready), r);
            kvp_remote[r] = mapa(smem_u32(s.kh), r);           // K/V planes == partial buffer
            qr_remote[r] = mapa(smem_u32(s.q_ready), r);
            pr_remote[r] = mapa(smem_u32(s.p_ready), r);
            qx_remote[r] = mapa(smem_u32(s.qx), r);
        }
        uint32_t gtile = 0;            // tiles consumed so far (all layers / steps): ring slot and parity
        uint32_t xphase = 0;           // layers done so far: parity of q_ready / p_ready / kv_free / kv_ready
        const float* pe_row = a.wp_pe + (size_t)(row0 + w) * E;      // this warp's row in the row-wise stages (row = warp)
        const bool row_ok = w < nloc;

        // ---- row-wise helpers: warp w owns row w; lane owns columns {2 lane, 2 lane + 1, 64 + 2 lane, 65 + 2 lane}
        auto col_of = [&](int i) { return (i < 2 ? 0 : 64) + 2 * lane + (i & 1); };
        auto ld_row4 = [&](const float* tile, float (&v)[4]) {
            const float2 x0 = *reinterpret_cast<const float2*>(tile + w * FPT + 2 * lane);
            const float2 x1 = *reinterpret_cast<const float2*>(tile + w * FPT + 64 + 2 * lane);
            v[0] = x0.x, v[1] = x0.y, v[2] = x1.x, v[3] = x1.y;
        };
        auto ld_vec4 = [&](const float* p, float (&v)[4]) {           // a [128]-padded parameter vector (shared or global)
            const float2 x0 = *reinterpret_cast<const float2*>(p + 2 * lane);
            const float2 x1 = *reinterpret_cast<const float2*>(p + 64 + 2 * lane);
            v[0] = x0.x, v[1] = x0.y, v[2] = x1.x, v[3] = x1.y;
        };
        auto ld_pe4 = [&](float (&v)[4]) {                            // waypoint embedding of this row ([E] floats, 8-byte aligned)
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = 0.f;
            if (row_ok) {
                const float2 x0 = __ldg(reinterpret_cast<const float2*>(pe_row + 2 * lane));
                v[0] = x0.x, v[1] = x0.y;
                if (lane < 28) {
                    const float2 x1 = __ldg(reinterpret_cast<const float2*>(pe_row + 64 + 2 * lane));
                    v[2] = x1.x, v[3] = x1.y;
                }
            }
        };
        // planes[row w] = split(v) in the natural channel order (columns >= E stay zero)
        auto st_planes4 = [&](__half* hi, __half* lo, const float (&v)[4]) {
            uint32_t h0, l0, h1 = 0u, l1 = 0u;
            split_h2(v[0], v[1], h0, l0);
            if (lane < 28) split_h2(v[2], v[3], h1, l1);
            *reinterpret_cast<uint32_t*>(hi + w * LP + 2 * lane) = h0;
            *reinterpret_cast<uint32_t*>(lo + w * LP + 2 * lane) = l0;
            *reinterpret_cast<uint32_t*>(hi + w * LP + 64 + 2 * lane) = h1;
            *reinterpret_cast<uint32_t*>(lo + w * LP + 64 + 2 * lane) = l1;
        };
        // v <- LayerNorm(v) over the E valid columns of the row (eps 1e-5, biased variance, two-pass)
        auto layernorm4 = [&](float (&v)[4], const float* gam, const float* bet) {
            float gg[4], bb[4];
            ld_vec4(gam, gg);
            ld_vec4(bet, bb);
            const bool hi_ok = lane < 28;
            float sum = v[0] + v[1] + (hi_ok ? v[2] + v[3] : 0.f);
            sum = warp_sum(sum);
            const float mean = sum * (1.0f / E);
            const float d0 = v[0] - mean, d1 = v[1] - mean, d2 = hi_ok ? v[2] - mean : 0.f, d3 = hi_ok ? v[3] - mean : 0.f;
            float var = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            var = warp_sum(var);
            const float rstd = 1.0f / sqrtf(var * (1.0f / E) + 1e-5f);
            v[0] = d0 * rstd * gg[0] + bb[0];
            v[1] = d1 * rstd * gg[1] + bb[1];
            v[2] = hi_ok ? d2 * rstd * gg[2] + bb[2] : 0.f;
            v[3] = hi_ok ? d3 * rstd * gg[3] + bb[3] : 0.f;
        };
        auto modulate4 = [&](float (&v)[4], const float* scale, const float* shift) {   // adaLN: x (1 + scale) + shift
            float sc[4], sh[4];
            ld_vec4(scale, sc);
            ld_vec4(shift, sh);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = (col_of(i) < E) ? v[i] * (1.0f + sc[i]) + sh[i] : 0.f;
        };
        // rotary of the accumulator pair (c, c+1) of local row r (diffusion positions are normalised coordinates)
        auto rotate = [&](int r, int c, float& v0, float& v1) {
            const int pi = c >> 1;
            if (pi < NPAIR) {
                const int axis = pi / (E / 6), j = pi - axis * (E / 6);
                float ang = __fmul_rn(s.trj[r * 9 + axis], s.freq[j]);
                ang = fmaf(-6.283185307179586f, rintf(ang * 0.15915494309189535f), ang);
                float sv, cv;
                __sincosf(ang, &sv, &cv);
                const float ev = v0, od = v1;
                v0 = ev * cv - od * sv;
                v1 = od * cv + ev * sv;
            }
        };
        // GEMM epilogues.  Warp w owns output columns 8 w + 2 q4 + {0, 1} of rows g and g + 8.
        const int ocol = 8 * w + 2 * q4;
        auto epi_tile = [&](const float (&acc)[4], const float* bias, float* tile) {          // tile = acc + bias (fp32)
            const float2 bb = *reinterpret_cast<const float2*>(bias + ocol);
            *reinterpret_cast<float2*>(tile + g * FPT + ocol) = make_float2(acc[0] + bb.x, acc[1] + bb.y);
            *reinterpret_cast<float2*>(tile + (g + 8) * FPT + ocol) = make_float2(acc[2] + bb.x, acc[3] + bb.y);
        };
        // head-padded planes of (rotary)(acc + bias): rows [prow0 + g, prow0 + g + 8] of a [.][LP] plane pair
        auto epi_heads = [&](const float (&acc)[4], const float* bias, bool rope, __half* ph, __half* pl, int prow0) {
            const float2 bb = *reinterpret_cast<const float2*>(bias + ocol);
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int r = g + 8 * hr;
                float v0 = acc[2 * hr] + bb.x, v1 = acc[2 * hr + 1] + bb.y;
                if (ocol < E) {
                    if (rope) rotate(r, ocol, v0, v1);
                    __half h0, l0, h1, l1;
                    split_h(v0, h0, l0);
                    split_h(v1, h1, l1);
                    const int s0 = (prow0 + r) * LP + head_slot16(ocol), s1 = (prow0 + r) * LP + head_slot16(ocol + 1);
                    ph[s0] = h0, pl[s0] = l0, ph[s1] = h1, pl[s1] = l1;
                } else {   // GEMM columns 120..127 own the zero pad slot of head c - 120
                    const int s0 = (prow0 + r) * LP + (ocol - E) * 16 + 15;
                    ph[s0] = pl[s0] = ph[s0 + 16] = pl[s0 + 16] = __float2half_rn(0.f);
                }
            }
        };
        // rotary Q of a cross-attention: fp16 [H][64][16], 32-byte rows with the two 16-byte halves swapped when (row>>2)&1
        // (the K tile layout: conflict-free ldmatrix); this CTA writes rows row0 .. row0+15; pad slot 15 of every head = 0
        auto qx_at = [&](int hh_, int grow, int d) { return (hh_ * 64 + grow) * 16 + ((((d >> 3) ^ ((grow >> 2) & 1))) << 3) + (d & 7); };
        auto epi_qx = [&](const float (&acc)[4], const float* bias) {
            const float2 bb = *reinterpret_cast<const float2*>(bias + ocol);
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int r = g + 8 * hr;
                float v0 = acc[2 * hr] + bb.x, v1 = acc[2 * hr + 1] + bb.y;
                if (ocol < E) {
                    rotate(r, ocol, v0, v1);
                    const int h0 = ocol / HD, h1 = (ocol + 1) / HD;
                    s.qx[qx_at(h0, row0 + r, ocol - h0 * HD)] = __float2half_rn(v0);
                    s.qx[qx_at(h1, row0 + r, ocol + 1 - h1 * HD)] = __float2half_rn(v1);
                } else {   // GEMM columns 120..127 own the pad slot of head c - 120
                    s.qx[qx_at(ocol - E, row0 + r, 15)] = __float2half_rn(0.f);
                    s.qx[qx_at(ocol + 1 - E, row0 + r, 15)] = __float2half_rn(0.f);
                }
            }
        };
        // ---- per-layer parameter vectors staged in shared memory by bulk copies (issued by one thread, a layer ahead):
        //      G = global layer index (step * nl + layer); block {C_BQ, adaLN row} of layer G in qv[G & 1], AdaV in vecs
        auto issue_qv = [&](int G) {            // tid 0 only
            const int st_ = G / a.nl, l_ = G - st_ * a.nl;
            const int tt = __ldg(a.timesteps + st_);
            float* dst = s.qv + (G & 1) * LoopSmem::QV;
            mbar_expect_tx(s.qv_full + (G & 1), LoopSmem::QV * 4);
            bulk_g2s(dst, a.ada_v[l_] + AdaV::C_BQ, EP * 4, s.qv_full + (G & 1));
            bulk_g2s(dst + EP, a.ada + ((size_t)tt * a.nl + l_) * ADA_ROW, ADA_ROW * 4, s.qv_full + (G & 1));
        };
        auto issue_vecs = [&](int G) {          // tid 0 only; every warp has left the previous layer's last use of vecs
            mbar_expect_tx(s.vec_full, AdaV::SIZE * 4);
            bulk_g2s(s.vecs, a.ada_v[G % a.nl], AdaV::SIZE * 4, s.vec_full);
        };
        const int g_total = a.n_steps * a.nl;
        // next cross-attention's Q from the residual row held in registers: A = adaLN_12[layer](x + pe); qx = rotary(A Wq^T + bq)
        auto make_q = [&](const float (&src)[4], int layer, int G) {
            WFrag<8> wq;
            wq.load(a.ada_w[layer] + AdaW::C_WQ, 16, w, lane);
            float v[4], pe[4];
            ld_pe4(pe);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = src[i] + pe[i];
            mbar_wait(s.qv_full + (G & 1), (G >> 1) & 1);
            const float* qb = s.qv + (G & 1) * LoopSmem::QV;
            modulate4(v, qb + EP, qb + 2 * EP);
            st_planes4(s.ah, s.al, v);
            csync();
            float acc[4];
            gemm16<8>(s.ah, s.al, wq, lane, acc);
            epi_qx(acc, qb);
            csync();
            // all-gather: this CTA's 16 rows of every head (512 contiguous bytes each) -> the same rows of the three peers
            // (they have all left the previous cross-attention: their K/V rows of that layer were needed to get here)
            for (int i = tid; i < H * 32; i += CT) {
                const uint32_t off = (uint32_t)(((i >> 5) * 64 + row0) * 32 + (i & 31) * 16);
                const uint4 val = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(s.qx) + off);
#pragma unroll
                for (int r = 0; r < CL; ++r)
                    if (r != (int)rank) st_cluster_v4(qx_remote[r] + off, val);
            }
            fence_cluster();
            csync();
            if (tid == 0)
                for (int r = 0; r < CL; ++r) mbar_arrive_cluster(qr_remote[r]);
        };
        // regressor head Linear(E,E) -> ReLU -> Linear(E,d) on the tile `x` -> s.upd[r][col0 + d]  (uses the kv planes as scratch)
        auto regress = [&](const float (&x)[4], const uint4* rw, const float* rv, int col0, int dim) {
            st_planes4(s.hh, s.hl, x);
            WFrag<8> w1, w2;
            w1.load(rw + MlpW::W1, 16, w, lane);
            if (w == 0) w2.load(rw + MlpW::W2, 16, 0, lane);
            csync();
            float acc[4];
            gemm16<8>(s.hh, s.hl, w1, lane, acc);
            {
                const float2 bb = __ldg(reinterpret_cast<const float2*>(rv + MlpV::B1 + ocol));
                uint32_t h0, l0, h1, l1;
                split_h2(fmaxf(acc[0] + bb.x, 0.f), fmaxf(acc[1] + bb.y, 0.f), h0, l0);
                split_h2(fmaxf(acc[2] + bb.x, 0.f), fmaxf(acc[3] + bb.y, 0.f), h1, l1);
                *reinterpret_cast<uint32_t*>(s.kh + g * LP + ocol) = h0;          // scratch planes: kh / kl rows 0..15
                *reinterpret_cast<uint32_t*>(s.kl + g * LP + ocol) = l0;
                *reinterpret_cast<uint32_t*>(s.kh + (g + 8) * LP + ocol) = h1;
                *reinterpret_cast<uint32_t*>(s.kl + (g + 8) * LP + ocol) = l1;
            }
            csync();
            if (w == 0) {
                gemm16<8>(s.kh, s.kl, w2, lane, acc);
                const float2 bb = __ldg(reinterpret_cast<const float2*>(rv + MlpV::B2 + ocol));
                if (ocol < dim) s.upd[g * 9 + col0 + ocol] = acc[0] + bb.x, s.upd[(g + 8) * 9 + col0 + ocol] = acc[2] + bb.x;
                if (ocol + 1 < dim) s.upd[g * 9 + col0 + ocol + 1] = acc[1] + bb.y, s.upd[(g + 8) * 9 + col0 + ocol + 1] = acc[3] + bb.y;
            }
            csync();
            // the scratch rows go back to zero padding-clean state is not needed: rows 0..15 are fully rewritten by every all-gather
        };

        if (tid == 0) {
            issue_qv(0);
            issue_vecs(0);
        }
        float xr[4], ys[4];            // residual row of this warp (row = warp) / output of the shared stack, in registers
#pragma unroll
        for (int i = 0; i < 4; ++i) ys[i] = 0.f;
        for (int step = 0; step < a.n_steps; ++step) {
            const int tstep = __ldg(a.timesteps + step);
            // ============================================================ step begin: trajectory encoder (+ instruction attention)
            {
                // Linear(9, E) + ReLU in fp32 (K = 9), straight into the GEMM planes
                WFrag<8> w2;
                w2.load(a.enc2, 16, w, lane);
                float v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = col_of(i);
                    float acc = 0.f;
                    if (c < E) {
                        acc = __ldg(a.enc1 + 9 * EP + c);
#pragma unroll
                        for (int k = 0; k < 9; ++k) acc = fmaf(s.trj[w * 9 + k], __ldg(a.enc1 + k * EP + c), acc);
                        acc = fmaxf(acc, 0.f);
                    }
                    v[i] = acc;
                }
                st_planes4(s.ah, s.al, v);
                csync();
                float acc[4];
                gemm16<8>(s.ah, s.al, w2, lane, acc);
                epi_tile(acc, a.enc2_b, s.t1);
                csync();
                ld_row4(s.t1, xr);
            }
            if (a.lang_w) {
                // trajectory tokens attend to the instruction tokens (diffusion_head.py:330-336): x = LN(x + Wo attn((x + pe) Wq, K, V))
                WFrag<8> wq;
                wq.load(a.lang_w + LangW::WQ, 16, w, lane);
                {
                    float v[4], pe[4];
                    ld_pe4(pe);
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = xr[i] + pe[i];
                    st_planes4(s.ah, s.al, v);
                }
                csync();                                           // every warp has read the encoder output out of t1
                // instruction K / V (fp32, per sample) -> head-padded planes, all 64 rows
                const float* lk = a.lang_k + (size_t)b * a.n_instr * E;
                const float* lv = a.lang_vv + (size_t)b * a.n_instr * E;
                for (int i = tid; i < 64 * EP; i += CT) {
                    const int r = i >> 7, slot = i & 127, hh_ = slot >> 4, d = slot & 15;
                    const bool ok = r < a.n_instr && d < HD;
                    __half x0, x1, y0, y1;
                    split_h(ok ? __ldg(lk + (size_t)r * E + hh_ * HD + d) : 0.f, x0, x1);
                    split_h(ok ? __ldg(lv + (size_t)r * E + hh_ * HD + d) : 0.f, y0, y1);
                    s.kh[r * LP + slot] = x0, s.kl[r * LP + slot] = x1, s.vh[r * LP + slot] = y0, s.vl[r * LP + slot] = y1;
                }
                csync();
                float acc[4];
                gemm16<8>(s.ah, s.al, wq, lane, acc);
                epi_heads(acc, a.lang_v + LangV::BQ, false, s.hh, s.hl, 0);
                WFrag<8> wo;
                wo.load(a.lang_w + LangW::WO, 16, w, lane);
                csync();
                if (w < H) mha16(w, lane, s.hh, s.hl, s.kh, s.kl, s.vh, s.vl, a.n_instr, nullptr, s.ah, s.al);
                csync();
                gemm16<8>(s.ah, s.al, wo, lane, acc);
                epi_tile(acc, a.lang_v + LangV::BO, s.t1);
                csync();
                float t[4];
                ld_row4(s.t1, t);
#pragma unroll
                for (int i = 0; i < 4; ++i) xr[i] += t[i];
                layernorm4(xr, a.lang_v + LangV::G12, a.lang_v + LangV::B12);
            }
            make_q(xr, 0, step * a.nl);

            for (int l = 0; l < a.nl; ++l) {
#ifdef A3D_CDL_TRACE
                const bool tr_on = (blockIdx.x == 21 && tid == 0 && step == 10 && l == 2);
#endif
                CDL_TRACE(0);
                const int G = step * a.nl + l;
                const uint4* lw = a.ada_w[l];
                const float* lv = s.vecs;                                      // AdaV of this layer (shared memory)
                const float* ada = s.qv + (G & 1) * LoopSmem::QV + EP;         // adaLN row of (timestep, layer)
                if (tid == 0 && G + 1 < g_total) issue_qv(G + 1);              // {C_BQ, adaLN row} of the next layer
                // ======================================================== cross-attention over the cached context K/V
                // Keys are split over the 4 CTAs; every CTA runs all 64 rows against its quarter of the tiles.
                // warp = (head, 32-row half): two m16 tiles, online softmax per row; the unnormalised partials
                // {O (15), denominator, row max} go to the CTA that owns the rows.
                mbar_wait_cluster(s.q_ready, xphase & 1);          // rotary Q of all 64 rows has landed (and the K/V planes of
                                                                   // every CTA are idle: they double as the partial buffers)
                CDL_TRACE(1);
                {
                    const int h = w & 7, rhalf = w >> 3;
                    uint32_t qf[2][4];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const int row = 32 * rhalf + 16 * mt + (lane & 7) + 8 * ((lane >> 3) & 1);
                        ldmatrix_x4(qf[mt], smem_u32(s.qx) + (uint32_t)((h * 64 + row) * 32 + (((lane >> 4) ^ ((row >> 2) & 1)) << 4)));
                    }
                    float o[2][2][4];
                    float mrow[2][2];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        mrow[mt][0] = mrow[mt][1] = -INFINITY;
#pragma unroll
                        for (int n = 0; n < 2; ++n)
#pragma unroll
                            for (int e = 0; e < 4; ++e) o[mt][n][e] = 0.f;
                    }
                    const bool tail_mask = (a.nk % kTileKeys) != 0;
                    for (int t = 0; t < my_tiles; ++t) {
                        const uint32_t gi = gtile + t, slot = gi % ST;
                        mbar_wait(s.full + slot, (gi / ST) & 1);
                        const uint32_t kbase = smem_u32(s.ring + slot * KV_TILE) + h * 2048, vbase = kbase + H * 2048;
                        const bool mask_tile = tail_mask && (t_begin + t == ntiles - 1);
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            float sc[8][4];
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj)
#pragma unroll
                                for (int e = 0; e < 4; ++e) sc[jj][e] = 0.f;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const int key = kk * 16 + (lane & 7) + 8 * (lane >> 4);
                                const int chunk = (lane >> 3) & 1;
                                uint32_t r[4];
                                ldmatrix_x4(r, kbase + key * 32 + ((chunk ^ ((key >> 2) & 1)) << 4));
                                mma_16816(sc[2 * kk], qf[mt], r[0], r[1]);
                                mma_16816(sc[2 * kk + 1], qf[mt], r[2], r[3]);
                            }
                            if (mask_tile) {
#pragma unroll
                                for (int jj = 0; jj < 8; ++jj)
#pragma unroll
                                    for (int e = 0; e < 4; ++e)
                                        if ((t_begin + t) * kTileKeys + 8 * jj + 2 * q4 + (e & 1) >= a.nk) sc[jj][e] = -INFINITY;
                            }
                            float mx0 = sc[0][0], mx1 = sc[0][2];
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                mx0 = fmaxf(mx0, fmaxf(sc[jj][0], sc[jj][1]));
                                mx1 = fmaxf(mx1, fmaxf(sc[jj][2], sc[jj][3]));
                            }
                            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
                            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
                            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
                            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
                            const float n0 = fmaxf(mrow[mt][0], mx0), n1 = fmaxf(mrow[mt][1], mx1);
                            const float al0 = exp2_fast(mrow[mt][0] - n0), al1 = exp2_fast(mrow[mt][1] - n1);
                            mrow[mt][0] = n0;
                            mrow[mt][1] = n1;
#pragma unroll
                            for (int n = 0; n < 2; ++n) {
                                o[mt][n][0] *= al0;
                                o[mt][n][1] *= al0;
                                o[mt][n][2] *= al1;
                                o[mt][n][3] *= al1;
                            }
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                uint32_t pa[4];
                                pa[0] = pack_h2(exp2_fast(sc[2 * kk][0] - n0), exp2_fast(sc[2 * kk][1] - n0));
                                pa[1] = pack_h2(exp2_fast(sc[2 * kk][2] - n1), exp2_fast(sc[2 * kk][3] - n1));
                                pa[2] = pack_h2(exp2_fast(sc[2 * kk + 1][0] - n0), exp2_fast(sc[2 * kk + 1][1] - n0));
                                pa[3] = pack_h2(exp2_fast(sc[2 * kk + 1][2] - n1), exp2_fast(sc[2 * kk + 1][3] - n1));
                                const int key = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
                                const int chunk = lane >> 4;
                                uint32_t r[4];
                                ldmatrix_x4_trans(r, vbase + key * 32 + ((chunk ^ ((key >> 2) & 1)) << 4));
                                mma_16816(o[mt][0], pa, r[0], r[1]);
                                mma_16816(o[mt][1], pa, r[2], r[3]);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(s.empty + slot);
                    }
                CDL_TRACE(2);
                    gtile += my_tiles;
                    // ---- partials -> the owner of the rows: m tile mt of this warp = rows of CTA 2 rhalf + mt
                    //      owner layout part[source CTA][head][17][16 rows] (floats)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const uint32_t dst = kvp_remote[2 * rhalf + mt] + (uint32_t)(((rank * H + h) * 17) * RB * 4);
#pragma unroll
                        for (int n = 0; n < 2; ++n)
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                st_cluster_f32(dst + (uint32_t)(((8 * n + 2 * q4 + (e & 1)) * RB + g + 8 * (e >> 1)) * 4), o[mt][n][e]);
                        if (q4 == 0) {
                            st_cluster_f32(dst + (uint32_t)((16 * RB + g) * 4), mrow[mt][0]);
                            st_cluster_f32(dst + (uint32_t)((16 * RB + g + 8) * 4), mrow[mt][1]);
                        }
                    }
                }
                CDL_TRACE(3);
                WFrag<8> wo;                                       // out-projection weights: in flight during the exchange
                wo.load(lw + AdaW::C_WO, 16, w, lane);
                fence_cluster();
                csync();
                if (tid == 0)
                    for (int r = 0; r < CL; ++r) mbar_arrive_cluster(pr_remote[r]);
                mbar_wait_cluster(s.p_ready, xphase & 1);          // the partials of all four key quarters are here
                CDL_TRACE(4);
                // ---- merge: warp = head, lane = (row, 8-wide half of the head dimension)
                if (w < H) {
                    const int h = w, r = lane & 15, dh = lane >> 4;
                    float m[CL], big = -INFINITY;
#pragma unroll
                    for (int sp = 0; sp < CL; ++sp) {
                        m[sp] = s.part[((sp * H + h) * 17 + 16) * RB + r];
                        big = fmaxf(big, m[sp]);
                    }
                    float acc[8];
#pragma unroll
                    for (int d = 0; d < 8; ++d) acc[d] = 0.f;
#pragma unroll
                    for (int sp = 0; sp < CL; ++sp) {
                        const float wgt = (m[sp] == -INFINITY) ? 0.f : exp2f(m[sp] - big);
#pragma unroll
                        for (int d = 0; d < 8; ++d) acc[d] = fmaf(wgt, s.part[((sp * H + h) * 17 + 8 * dh + d) * RB + r], acc[d]);
                    }
                    const float den = __shfl_sync(0xffffffffu, acc[7], r | 16);     // slot 15 = denominator (rode in V's slot 15)
                    const float inv = 1.0f / den;
#pragma unroll
                    for (int d = 0; d < 8; ++d) {
                        if (8 * dh + d < HD) {
                            __half x, y;
                            split_h(acc[d] * inv, x, y);
                            s.ah[r * LP + h * HD + 8 * dh + d] = x;
                            s.al[r * LP + h * HD + 8 * dh + d] = y;
                        }
                    }
                }
                csync();
                CDL_TRACE(5);
                // the partial buffer (= K/V planes) of this CTA is free: peers may write their K / V rows of this layer
                if (tid == 0)
                    for (int r = 0; r < CL; ++r) mbar_arrive_cluster(kvf_remote[r]);
                {
                    // ---- out projection + residual + LN_12                                       (layers.py:146-147)
                    float acc[4];
                    gemm16<8>(s.ah, s.al, wo, lane, acc);
                    mbar_wait(s.vec_full, G & 1);                  // this layer's parameter vectors have landed (issued a layer ago)
                    epi_tile(acc, lv + AdaV::C_BO, s.t1);
                }
                WFrag<8> wv, wk;
                wv.load(lw + AdaW::S_WV, 16, w, lane);
                wk.load(lw + AdaW::S_WK, 16, w, lane);
                csync();
                {
                    float v[4], t[4], pe[4];
                    ld_row4(s.t1, t);
#pragma unroll
                    for (int i = 0; i < 4; ++i) xr[i] += t[i];
                    layernorm4(xr, lv + AdaV::G12, lv + AdaV::B12);
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = xr[i];
                    // ---- self-attention inputs: q = k = adaLN_1(x + pe) -> (ah, al); v = adaLN_1(x) -> (hh, hl)   (layers.py:165-182)
                    ld_pe4(pe);
                    float qk[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) qk[i] = v[i] + pe[i];
                    modulate4(qk, ada + 2 * EP, ada + 3 * EP);
                    modulate4(v, ada + 2 * EP, ada + 3 * EP);
                    st_planes4(s.ah, s.al, qk);
                    st_planes4(s.hh, s.hl, v);
                }
                csync();
                CDL_TRACE(6);
                // every CTA has released the K/V planes of its previous use (end of the previous layer / step begin)
                mbar_wait_cluster(s.kv_free, xphase & 1);
                CDL_TRACE(7);
                {
                    float acc[4];
                    gemm16<8>(s.hh, s.hl, wv, lane, acc);
                    epi_heads(acc, lv + AdaV::S_BV, false, s.vh, s.vl, row0);
                    gemm16<8>(s.ah, s.al, wk, lane, acc);
                    epi_heads(acc, lv + AdaV::S_BK, true, s.kh, s.kl, row0);
                }
                CDL_TRACE(8);
                WFrag<8> wq;
                wq.load(lw + AdaW::S_WQ, 16, w, lane);
                csync();                                           // own rows of K / V complete; (hh, hl) free for Q
                // ---- all-gather: this CTA's 16 rows of the four planes -> the same rows of the three peers
                {
                    constexpr int V4_PER_PLANE = RB * LP * 2 / 16;               // 272 uint4 per 16-row block
                    for (int i = tid; i < 4 * V4_PER_PLANE; i += CT) {
                        const int pl = i / V4_PER_PLANE, o4 = i - pl * V4_PER_PLANE;
                        const uint32_t off = (uint32_t)((pl * 64 + row0) * LP * 2 + o4 * 16);
                        const uint4 val = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(s.kh) + off);
#pragma unroll
                        for (int r = 0; r < CL; ++r)
                            if (r != (int)rank) st_cluster_v4(kvp_remote[r] + off, val);
                    }
                }
                {
                    float acc[4];
                    gemm16<8>(s.ah, s.al, wq, lane, acc);
                    epi_heads(acc, lv + AdaV::S_BQ, true, s.hh, s.hl, 0);
                }
                WFrag<8> wso;
                wso.load(lw + AdaW::S_WO, 16, w, lane);
                fence_cluster();
                csync();                                           // remote stores ordered by every thread; Q planes complete
                if (tid == 0)
                    for (int r = 0; r < CL; ++r) mbar_arrive_cluster(kvr_remote[r]);
                CDL_TRACE(9);
                mbar_wait_cluster(s.kv_ready, xphase & 1);         // all four row blocks of K / V have landed here
                ++xphase;
                CDL_TRACE(10);
                if (w < H) mha16(w, lane, s.hh, s.hl, s.kh, s.kl, s.vh, s.vl, a.nrows, a.mask ? s.kmask : nullptr, s.ah, s.al);
                csync();
                CDL_TRACE(11);
                {
                    float acc[4];
                    gemm16<8>(s.ah, s.al, wso, lane, acc);
                    epi_tile(acc, lv + AdaV::S_BO, s.t1);
                }
                csync();
                CDL_TRACE(12);
                // ---- x = LN_1(x + sa); y = adaLN_ff(x); FFN                                      (layers.py:183-209)
                float yv[4];
                {
                    float v[4], t[4];
                    ld_row4(s.t1, t);
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = xr[i] + t[i];
                    layernorm4(v, lv + AdaV::G1, lv + AdaV::B1N);
                    modulate4(v, ada + 4 * EP, ada + 5 * EP);
#pragma unroll
                    for (int i = 0; i < 4; ++i) yv[i] = v[i];      // y stays in this warp's registers for the residual
                    st_planes4(s.ah, s.al, v);
                }
                {
                    float sum[4] = {0.f, 0.f, 0.f, 0.f};
                    WFrag<8> w1, w2;
                    w1.load(lw + AdaW::W1, 64, w, lane);
                    csync();
#pragma unroll 1
                    for (int ch = 0; ch < FFP / 128; ++ch) {
                        w2.load(lw + AdaW::W2 + (size_t)ch * 8 * 16 * FRAG, 16, w, lane);
                        float acc[4];
                        gemm16<8>(s.ah, s.al, w1, lane, acc);
                        if (ch + 1 < FFP / 128) w1.load(lw + AdaW::W1, 64, 16 * (ch + 1) + w, lane);
                        {
                            const float2 bb = *reinterpret_cast<const float2*>(lv + AdaV::B1 + 128 * ch + ocol);
                            uint32_t h0, l0, h1, l1;
                            split_h2(fmaxf(acc[0] + bb.x, 0.f), fmaxf(acc[1] + bb.y, 0.f), h0, l0);
                            split_h2(fmaxf(acc[2] + bb.x, 0.f), fmaxf(acc[3] + bb.y, 0.f), h1, l1);
                            *reinterpret_cast<uint32_t*>(s.hh + g * LP + ocol) = h0;
                            *reinterpret_cast<uint32_t*>(s.hl + g * LP + ocol) = l0;
                            *reinterpret_cast<uint32_t*>(s.hh + (g + 8) * LP + ocol) = h1;
                            *reinterpret_cast<uint32_t*>(s.hl + (g + 8) * LP + ocol) = l1;
                        }
                        csync();
                        gemm16<8>(s.hh, s.hl, w2, lane, acc);
#pragma unroll
                        for (int e = 0; e < 4; ++e) sum[e] += acc[e];
                        csync();                                   // (hh, hl) free for the next chunk
                    }
                    epi_tile(sum, lv + AdaV::B2, s.t1);
                }
                csync();
                CDL_TRACE(14);
                // ---- layer output x' = LN_122(y + FFN(y)); routing of the position / rotation heads   (diffusion_head.py:338-363)
                const int last_shared = a.n_traj - 1, p1 = a.n_traj + 1, r1 = a.n_traj + 3;
                {
                    float t[4];
                    ld_row4(s.t1, t);
#pragma unroll
                    for (int i = 0; i < 4; ++i) xr[i] = yv[i] + t[i];
                    layernorm4(xr, lv + AdaV::G122, lv + AdaV::B122);
                    if (l == last_shared) {                        // both heads start from the output of the shared stack
#pragma unroll
                        for (int i = 0; i < 4; ++i) ys[i] = xr[i];
                    }
                }
                csync();                                           // every warp is done with this layer's vectors (and with t1)
                if (tid == 0 && G + 1 < g_total) issue_vecs(G + 1);
                if (l == p1) {
                    regress(xr, a.pos_w, a.pos_v, 0, 3);
                    // the rotation head restarts from the shared-stack output
#pragma unroll
                    for (int i = 0; i < 4; ++i) xr[i] = ys[i];
                } else if (l == r1) {
                    regress(xr, a.rot_w, a.rot_v, 3, 6);
                }
                CDL_TRACE(15);
                if (l + 1 < a.nl) make_q(xr, l + 1, G + 1);
                CDL_TRACE(16);
            }
            // ============================================================ denoiser output + DDPM posterior step
            // (diffusion_head.py:271-274: position is residual on the noisy input, rotation is direct;
            //  diffusion_model.py:105-117: inpainting, prediction_type "sample", clip_sample, fixed_small variance)
            if (tid < RB * 9) {
                const int r = tid / 9, d = tid - 9 * r;
                if (r < nloc) {
                    const size_t gi = ((size_t)b * a.nrows + row0 + r) * 9 + d;
                    const float cur = s.trj[tid];
                    float out = (d < 3) ? cur + s.upd[tid] : s.upd[tid];
                    if (a.cmask[gi]) out = a.cond[gi];
                    float nxt = out;
                    if (step + 1 < a.n_steps) {
                        const float* cf = a.coef + (size_t)tstep * 6 + (d < 3 ? 0 : 3);
                        const float x0 = fminf(fmaxf(out, -1.f), 1.f);
                        nxt = __ldg(cf) * x0 + __ldg(cf + 1) * cur;
                        const size_t ni = ((size_t)step * a.batch + b) * a.nrows + row0 + r;
                        const float nz = (d < 3) ? __ldg(a.noise_pos + ni * 3 + d) : __ldg(a.noise_rot + ni * 6 + d - 3);
                        nxt += __ldg(cf + 2) * nz;
                    }
                    s.trj[tid] = nxt;
                    if (step + 1 == a.n_steps) a.traj[gi] = nxt;
                }
            }
            csync();
        }
    }
    // nobody leaves while a peer may still write into this CTA's shared memory (or multicast into it)
    __syncthreads();
    cluster_sync_all();
}

}  // namespace cd
}  // namespace a3d

using namespace a3d;
using namespace a3d::cd;

#ifdef A3D_CDL_TRACE
extern "C" int cd_loop_trace_read(long long* host) {
    return cudaMemcpyFromSymbol(host, g_cdl_trace, sizeof(g_cdl_trace)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int cd_denoise_loop(float* traj, int batch, int length, int n_steps, const float* cond, const unsigned char* cond_mask,
                               const unsigned char* mask, const float* wp_pe, const int* timesteps, const float* ada,
                               int ada_layers, int n_traj_layers, const float* coef, const float* noise_pos,
                               const float* noise_rot, const float* traj_enc1, const void* traj_enc2, const float* traj_enc2_b,
                               const void* lang_w, const float* lang_v, const float* lang_k, const float* lang_vv, int n_instr,
                               const void* const* ada_w_host, const float* const* ada_v_host, const void* pos_reg_w,
                               const float* pos_reg_v, const void* rot_reg_w, const float* rot_reg_v, const void* kv,
                               size_t kv_set_bytes, int nk, void* stream) {
    A3D_REQUIRE(traj && cond && cond_mask && wp_pe && timesteps && ada && coef && traj_enc1 && traj_enc2 && traj_enc2_b && ada_w_host &&
                    ada_v_host && pos_reg_w && pos_reg_v && rot_reg_w && rot_reg_v && kv,
                "cd_denoise_loop: null pointer");
    A3D_REQUIRE(batch > 0 && length > 0 && length <= ROWS && n_steps > 0, "cd_denoise_loop: bad sizes (B=%d L=%d steps=%d)", batch, length, n_steps);
    A3D_REQUIRE(n_steps == 1 || (noise_pos && noise_rot), "cd_denoise_loop: noise tensors missing");
    A3D_REQUIRE(ada_layers == n_traj_layers + 4 && ada_layers <= MAXL && n_traj_layers >= 1,
                "cd_denoise_loop: expects n_traj shared layers + 2 position + 2 rotation layers (got %d, %d)", ada_layers, n_traj_layers);
    A3D_REQUIRE(!lang_w || (lang_v && lang_k && lang_vv && n_instr > 0 && n_instr <= 64), "cd_denoise_loop: instruction K/V missing");
    A3D_REQUIRE(nk > 0 && ((uintptr_t)kv & 15) == 0 && (kv_set_bytes & 15) == 0, "cd_denoise_loop: bad K/V cache");
    LoopArgs a{};
    a.batch = batch, a.nrows = length, a.n_steps = n_steps, a.nl = ada_layers, a.n_traj = n_traj_layers, a.nk = nk;
    a.ntiles = (nk + kTileKeys - 1) / kTileKeys;
    a.pf_tiles = g_cd_prefetch_tiles;
    a.n_instr = n_instr;
    a.traj = traj, a.cond = cond, a.cmask = cond_mask, a.mask = mask, a.wp_pe = wp_pe, a.timesteps = timesteps, a.ada = ada, a.coef = coef;
    a.noise_pos = noise_pos, a.noise_rot = noise_rot;
    a.enc1 = traj_enc1, a.enc2 = (const uint4*)traj_enc2, a.enc2_b = traj_enc2_b;
    a.lang_w = (const uint4*)lang_w, a.lang_v = lang_v, a.lang_k = lang_k, a.lang_vv = lang_vv;
    for (int l = 0; l < ada_layers; ++l) {
        A3D_REQUIRE(ada_w_host[l] && ada_v_host[l], "cd_denoise_loop: layer %d weights missing", l);
        a.ada_w[l] = (const uint4*)ada_w_host[l];
        a.ada_v[l] = ada_v_host[l];
    }
    a.pos_w = (const uint4*)pos_reg_w, a.pos_v = pos_reg_v, a.rot_w = (const uint4*)rot_reg_w, a.rot_v = rot_reg_v;
    a.kv = (const unsigned char*)kv, a.kv_set_bytes = kv_set_bytes;
    static PerDeviceOnce once_dev;
    if (bool& once = once_dev.flag(); !once) {
        cudaError_t e = cudaFuncSetAttribute(cd_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LoopSmem::BYTES);
        if (e != cudaSuccess) {
            set_error("cd_denoise_loop: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return A3D_ECUDA;
        }
        once = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(batch * CL);
    cfg.blockDim = dim3(LT);
    cfg.dynamicSmemBytes = LoopSmem::BYTES;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, cd_loop_kernel, a);
    if (e != cudaSuccess) {
        set_error("cd_denoise_loop: launch failed: %s", cudaGetErrorString(e));
        return A3D_ECUDA;
    }
    return check_launch("cd_denoise_loop");
}
