// Geometry kernels of the Act3D path: point pyramid, local top-k context selection, token
// gather, top-ghost pick and the device-side ghost sampler.  All are HBM/latency-bound
// integer/float32 kernels: coalesced vector loads, warp-shuffle reductions, no tensor cores.
#include <stdarg.h>

#include "a3d_common.cuh"

namespace a3d {

// ------------------------------------------------------------------ error plumbing (host)
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return A3D_ECUDA;
    }
    return A3D_OK;
}

// =================================================================== point pyramid
// out[(n*h + y)*w + x][c] = 0.25 * (((p00 + p01) + p10) + p11) at the two centre pixels per axis of
// the f x f block: bilinear interpolation with lambda = 0.5 on both axes (act3d.py:379-380),
// accumulated in the order of torch's CPU kernel so that the result is bit-identical to the CPU
// reference (torch's own CUDA kernel pairs the taps differently and is 1 ulp away from it).
// One thread per output point, three channels.
__global__ void __launch_bounds__(256) pcd_pyramid_kernel(const float* __restrict__ pcd, int bn, int H, int W, int f,
                                                          float* __restrict__ out) {
    const int h = H / f, w = W / f;
    const long total = (long)bn * h * w;
    const int lo = f / 2 - 1;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const int y = (int)((i / w) % h);
        const long n = i / ((long)w * h);
        const int sy = f * y + lo, sx = f * x + lo;
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* p = pcd + ((n * 3 + c) * H + sy) * (long)W + sx;
            // weights are all 0.25; torch's CPU kernel accumulates the four taps in raster order
            const float acc = __fadd_rn(__fadd_rn(__fadd_rn(__ldg(p), __ldg(p + 1)), __ldg(p + W)), __ldg(p + W + 1));
            v[c] = __fmul_rn(0.25f, acc);
        }
        out[i * 3 + 0] = v[0];
        out[i * 3 + 1] = v[1];
        out[i * 3 + 2] = v[2];
    }
}

// =================================================================== local top-k (radix select)
// One CTA per sample.  Keys are the fp32 bit patterns of the (non-negative) distance, which
// order like the floats.  Three histogram passes (11 + 11 + 10 bits) find the exact k-th key,
// a fourth pass collects everything below it plus the lowest-index ties, then the K survivors
// are bitonic-sorted by (key, index) in shared memory.  The distance is recomputed from the
// L2-resident points in every pass (12 B/point) instead of spilling 4 B keys to HBM.
constexpr int kTopkThreads = 1024;
constexpr int kMaxK = 8192;

template <bool kTraj>
__device__ __forceinline__ uint32_t dist_key(const float* __restrict__ pts, long i, const float* c, int traj_len) {
    const float px = __ldg(pts + 3 * i), py = __ldg(pts + 3 * i + 1), pz = __ldg(pts + 3 * i + 2);
    if (!kTraj) {
        const float dx = __fsub_rn(c[0], px), dy = __fsub_rn(c[1], py), dz = __fsub_rn(c[2], pz);
        const float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        return __float_as_uint(__fsqrt_rn(s));
    } else {
        float best = INFINITY;
        for (int l = 0; l < traj_len; ++l) {
            const float dx = __fsub_rn(c[3 * l], px), dy = __fsub_rn(c[3 * l + 1], py), dz = __fsub_rn(c[3 * l + 2], pz);
            const float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            best = fminf(best, s);
        }
        return __float_as_uint(best);
    }
}

// Finds the bin where the running count crosses `need`; returns bin and the count below it.
// hist has nbins (<= 2048) entries; all threads must call; result broadcast through smem.
__device__ void find_bin(const uint32_t* hist, int nbins, uint32_t need, uint32_t* sh_scan, uint32_t* sh_out) {
    // each thread owns two consecutive bins
    const int t = threadIdx.x;
    const uint32_t a = (2 * t < nbins) ? hist[2 * t] : 0u;
    const uint32_t b = (2 * t + 1 < nbins) ? hist[2 * t + 1] : 0u;
    uint32_t v = a + b;
    // inclusive block scan over 1024 partial sums
    const int lane = t & 31, wid = t >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) sh_scan[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = sh_scan[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += n;
        }
        sh_scan[lane] = wi - w;   // exclusive warp offsets
    }
    __syncthreads();
    const uint32_t excl = sh_scan[wid] + incl - v;   // items strictly before bin 2t
    if (excl < need && need <= excl + a) {
        sh_out[0] = 2 * t;
        sh_out[1] = excl;
    } else if (excl + a < need && need <= excl + a + b) {
        sh_out[0] = 2 * t + 1;
        sh_out[1] = excl + a;
    }
    __syncthreads();
}

template <bool kTraj>
__global__ void __launch_bounds__(kTopkThreads, 1)
    topk_kernel(const float* __restrict__ center, int traj_len, const float* __restrict__ pts, int n, int k,
                int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
    extern __shared__ unsigned long long sh_sel[];   // next_pow2(k) composite (key<<32 | idx)
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t sh_scan[32];
    __shared__ uint32_t sh_res[2];
    __shared__ uint32_t sh_cnt[2];
    __shared__ float sh_c[3 * 64];

    const int b = blockIdx.x;
    const int t = threadIdx.x;
    const float* p = pts + (long)b * n * 3;
    const int clen = kTraj ? traj_len : 1;
    for (int i = t; i < 3 * clen; i += blockDim.x) sh_c[i] = center[(long)b * 3 * clen + i];

    uint32_t prefix = 0;        // key bits fixed so far
    uint32_t need = (uint32_t)k;   // how many still to take from the current candidate set
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    uint32_t fixed_mask = 0;
    for (int pass = 0; pass < 3; ++pass) {
        for (int i = t; i < 2048; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const int sh = shifts[pass];
        const uint32_t bm = (1u << widths[pass]) - 1u;
        for (int i = t; i < n; i += blockDim.x) {
            const uint32_t key = dist_key<kTraj>(p, i, sh_c, traj_len);
            if ((key & fixed_mask) == prefix) atomicAdd(&hist[(key >> sh) & bm], 1u);
        }
        __syncthreads();
        find_bin(hist, 1 << widths[pass], need, sh_scan, sh_res);
        prefix |= sh_res[0] << sh;
        fixed_mask |= bm << sh;
        need -= sh_res[1];
        __syncthreads();
    }
    const uint32_t kth = prefix;         // exact key of the k-th smallest element
    const uint32_t ties_wanted = need;   // how many elements equal to kth belong to the result

    // ---- collect: everything < kth, and ties in index order
    int kp = 1;
    while (kp < k) kp <<= 1;
    for (int i = t; i < kp; i += blockDim.x) sh_sel[i] = ~0ull;
    if (t == 0) {
        sh_cnt[0] = 0;
        sh_cnt[1] = 0;
    }
    __syncthreads();
    const uint32_t n_below = (uint32_t)k - ties_wanted;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + t;
        uint32_t key = 0xffffffffu;
        if (i < n) key = dist_key<kTraj>(p, i, sh_c, traj_len);
        if (i < n && key < kth) {
            const uint32_t slot = atomicAdd(&sh_cnt[0], 1u);
            sh_sel[slot] = ((unsigned long long)key << 32) | (uint32_t)i;
        }
        // ties: rank in index order = (ties seen in earlier chunks) + (ties at lower lanes/warps of this chunk)
        const bool tie = (i < n) && (key == kth);
        const unsigned bal = __ballot_sync(0xffffffffu, tie);
        if (__syncthreads_or(tie)) {
            // rare path (at least one tie in this chunk): ordered rank via per-warp counts
            const int lane = t & 31, wid = t >> 5;
            if (lane == 0) sh_scan[wid] = __popc(bal);
            __syncthreads();
            uint32_t before = sh_cnt[1];
            for (int w2 = 0; w2 < wid; ++w2) before += sh_scan[w2];
            const uint32_t rank = before + __popc(bal & ((1u << lane) - 1u));
            if (tie && rank < ties_wanted) sh_sel[n_below + rank] = ((unsigned long long)key << 32) | (uint32_t)i;
            __syncthreads();
            if (t == 0) {
                uint32_t tot = 0;
                for (int w2 = 0; w2 < 32; ++w2) tot += sh_scan[w2];
                sh_cnt[1] += tot;
            }
            __syncthreads();
        }
    }
    __syncthreads();

    // ---- bitonic sort of kp composites (padding = ~0 sorts last)
    for (int size = 2; size <= kp; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < (kp >> 1); i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long a = sh_sel[lo], c2 = sh_sel[hi];
                if ((a > c2) == up) {
                    sh_sel[lo] = c2;
                    sh_sel[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = t; i < k; i += blockDim.x) {
        const unsigned long long v = sh_sel[i];
        idx_out[(long)b * k + i] = (int32_t)(uint32_t)(v & 0xffffffffu);
        if (dist_out) dist_out[(long)b * k + i] = __uint_as_float((uint32_t)(v >> 32));
    }
}

// ------------------------------------------------------------------ cluster variant (n <= 65536)
// The single-CTA kernel above is latency-bound (150 us at 65536 points: four passes that recompute every distance, one
// SM per sample).  Here a thread-block cluster of 8 CTAs owns a sample: every CTA takes a contiguous eighth of the points
// and keeps its keys IN REGISTERS (<= 8 per thread, computed once from one coalesced read of the points), so the radix
// passes touch no memory but shared-memory histograms; the 8 histograms are combined in the leader CTA through
// distributed shared memory (remote atomics), the survivors are written into the leader's buffer at offsets derived from
// the per-CTA counts (CTA rank order == index order, so exact ties keep the lowest indices), and the leader sorts them.
// Same keys, same tie rule, same output as topk_kernel: bit-exact.
constexpr int kTopkCluster = 8;
constexpr int kKeysPerThread = 8;

template <bool kTraj>
__global__ void __cluster_dims__(kTopkCluster, 1, 1) __launch_bounds__(kTopkThreads, 1)
    topk_cluster_kernel(const float* __restrict__ center, int traj_len, const float* __restrict__ pts, int n, int k,
                        int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
    extern __shared__ unsigned long long sh_sel[];   // leader: next_pow2(k) composites (key << 32 | idx)
    __shared__ uint32_t hist[2048];                  // this CTA's histogram of the current pass
    __shared__ uint32_t ghist[2][2048];              // leader: cluster-wide histogram, double-buffered over the passes
    __shared__ uint32_t sh_scan[32];
    __shared__ uint32_t sh_res[2];
    __shared__ uint32_t sh_below[kTopkCluster], sh_ties[kTopkCluster];   // per-CTA counts, replicated in every CTA
    __shared__ uint32_t sh_cnt[2];
    __shared__ float sh_c[3 * 64];

    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    auto remote = [](const void* p, uint32_t r) {    // address of the same shared-memory object in CTA r of the cluster
        uint32_t a = (uint32_t)__cvta_generic_to_shared(p), o;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r));
        return o;
    };
    auto cluster_sync = []() {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    };
    const int b = blockIdx.x / kTopkCluster;
    const int t = threadIdx.x;
    const float* p = pts + (long)b * n * 3;
    const int clen = kTraj ? traj_len : 1;
    for (int i = t; i < 3 * clen; i += blockDim.x) sh_c[i] = center[(long)b * 3 * clen + i];
    for (int i = t; i < 2 * 2048; i += blockDim.x) (&ghist[0][0])[i] = 0;
    __syncthreads();

    // ---- this CTA's points: [lo, hi), thread t owns lo + j * 1024 + t
    const int per = (n + kTopkCluster - 1) / kTopkCluster;
    const int lo = min((int)rank * per, n), hi = min(lo + per, n);
    uint32_t key[kKeysPerThread];
#pragma unroll
    for (int j = 0; j < kKeysPerThread; ++j) {
        const int i = lo + j * kTopkThreads + t;
        key[j] = (i < hi) ? dist_key<kTraj>(p, i, sh_c, traj_len) : 0xffffffffu;
    }
    const uint32_t my_valid = (uint32_t)max(0, min(kKeysPerThread, (hi - lo - t + kTopkThreads - 1) / kTopkThreads));
    cluster_sync();                                   // the leader's histograms are zeroed before anyone adds to them

    uint32_t prefix = 0, need = (uint32_t)k, fixed_mask = 0;
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
        for (int i = t; i < 2048; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const int sh = shifts[pass];
        const uint32_t bm = (1u << widths[pass]) - 1u;
#pragma unroll
        for (int j = 0; j < kKeysPerThread; ++j)
            if ((uint32_t)j < my_valid && (key[j] & fixed_mask) == prefix) atomicAdd(&hist[(key[j] >> sh) & bm], 1u);
        __syncthreads();
        {   // this CTA's bins -> the leader's histogram of this pass (remote shared-memory reductions)
            const uint32_t g = remote(&ghist[pass & 1][0], 0);
            for (int i = t; i < (1 << widths[pass]); i += blockDim.x) {
                const uint32_t v = hist[i];
                if (v) asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(g + 4u * i), "r"(v) : "memory");
            }
        }
        cluster_sync();
        if (rank == 0) {
            find_bin(ghist[pass & 1], 1 << widths[pass], need, sh_scan, sh_res);
            for (int i = t; i < 2048; i += blockDim.x) ghist[pass & 1][i] = 0;   // ready for pass + 2
        }
        cluster_sync();
        {   // everybody reads the leader's verdict
            uint32_t r0, r1;
            const uint32_t a = remote(sh_res, 0);
            asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(r0) : "r"(a));
            asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(r1) : "r"(a + 4u));
            prefix |= r0 << sh;
            fixed_mask |= bm << sh;
            need -= r1;
        }
        cluster_sync();                               // sh_res may be rewritten by the next pass only after all have read it
    }
    const uint32_t kth = prefix, ties_wanted = need, n_below = (uint32_t)k - ties_wanted;

    // ---- counts of this CTA -> every CTA
    if (t == 0) {
        sh_cnt[0] = 0;
        sh_cnt[1] = 0;
    }
    __syncthreads();
    {
        uint32_t nb = 0, nt = 0;
#pragma unroll
        for (int j = 0; j < kKeysPerThread; ++j)
            if ((uint32_t)j < my_valid) {
                nb += key[j] < kth;
                nt += key[j] == kth;
            }
        nb = __reduce_add_sync(0xffffffffu, nb);
        nt = __reduce_add_sync(0xffffffffu, nt);
        if ((t & 31) == 0) {
            atomicAdd(&sh_cnt[0], nb);
            atomicAdd(&sh_cnt[1], nt);
        }
    }
    __syncthreads();
    if (t < kTopkCluster) {
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote(&sh_below[rank], t)), "r"(sh_cnt[0]) : "memory");
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote(&sh_ties[rank], t)), "r"(sh_cnt[1]) : "memory");
    }
    int kp = 1;
    while (kp < k) kp <<= 1;
    if (rank == 0)
        for (int i = t; i < kp; i += blockDim.x) sh_sel[i] = ~0ull;
    cluster_sync();
    uint32_t base_below = 0, base_tie = 0;
    for (uint32_t r = 0; r < rank; ++r) {
        base_below += sh_below[r];
        base_tie += sh_ties[r];
    }
    // ---- survivors -> the leader's buffer
    if (t == 0) {
        sh_cnt[0] = 0;
        sh_cnt[1] = 0;
    }
    __syncthreads();
    const uint32_t sel0 = remote(sh_sel, 0);
#pragma unroll
    for (int j = 0; j < kKeysPerThread; ++j) {
        const bool ok = (uint32_t)j < my_valid;
        const uint32_t i = (uint32_t)(lo + j * kTopkThreads + t);
        if (ok && key[j] < kth) {
            const uint32_t slot = base_below + atomicAdd(&sh_cnt[0], 1u);
            asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(sel0 + 8u * slot), "l"(((unsigned long long)key[j] << 32) | i) : "memory");
        }
        // ties in index order: (ties of lower ranks) + (earlier chunks of this CTA) + (lower threads of this chunk)
        const bool tie = ok && key[j] == kth;
        const unsigned bal = __ballot_sync(0xffffffffu, tie);
        if (__syncthreads_or(tie)) {
            const int lane = t & 31, wid = t >> 5;
            if (lane == 0) sh_scan[wid] = __popc(bal);
            __syncthreads();
            uint32_t before = base_tie + sh_cnt[1];
            for (int w2 = 0; w2 < wid; ++w2) before += sh_scan[w2];
            const uint32_t trank = before + __popc(bal & ((1u << lane) - 1u));
            if (tie && trank < ties_wanted)
                asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(sel0 + 8u * (n_below + trank)), "l"(((unsigned long long)key[j] << 32) | i)
                             : "memory");
            __syncthreads();
            if (t == 0) {
                uint32_t tot = 0;
                for (int w2 = 0; w2 < 32; ++w2) tot += sh_scan[w2];
                sh_cnt[1] += tot;
            }
            __syncthreads();
        }
    }
    cluster_sync();                                   // all survivors have landed in the leader
    if (rank != 0) return;

    // ---- bitonic sort of kp composites (padding = ~0 sorts last)
    for (int size = 2; size <= kp; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < (kp >> 1); i += blockDim.x) {
                const int l2 = 2 * i - (i & (stride - 1));
                const int h2 = l2 + stride;
                const bool up = ((l2 & size) == 0);
                const unsigned long long a = sh_sel[l2], c2 = sh_sel[h2];
                if ((a > c2) == up) {
                    sh_sel[l2] = c2;
                    sh_sel[h2] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = t; i < k; i += blockDim.x) {
        const unsigned long long v = sh_sel[i];
        idx_out[(long)b * k + i] = (int32_t)(uint32_t)(v & 0xffffffffu);
        if (dist_out) dist_out[(long)b * k + i] = __uint_as_float((uint32_t)(v >> 32));
    }
}

// =================================================================== token gather
// tok[b][r][c] = feat[b*ncam + cam][c][pix],  (cam, pix) = divmod(idx[b][r], hw); pos likewise.
// A block moves 32 tokens x E channels through a padded shared tile so that the scattered
// reads are issued token-major (one 32 B sector per (token, channel)) and the writes are
// fully coalesced rows of E floats.
// channels-last variant: feat[(b*ncam + cam)][pix][c] -- a token is one contiguous E-float row, so the gather is
// a plain coalesced row copy (this is the layout cuDNN's NHWC kernels leave the FPN output in).
template <int E>
__global__ void __launch_bounds__(256) gather_tokens_nhwc_kernel(const float* __restrict__ feat, const float* __restrict__ pcd,
                                                                 const int32_t* __restrict__ idx, int ncam, int hw, int k,
                                                                 float* __restrict__ tok, float* __restrict__ pos,
                                                                 int tok_rows, const float* __restrict__ bias) {
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // a warp owns rows r0 + wid + {0, 8, 16, 24}: all four index loads first, then all feature / point loads in flight
    // at once, then the stores (one dependent round trip per CTA instead of two per row)
    constexpr int NC = (E + 31) / 32;
    int s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = r0 + wid + 8 * j;
        s[j] = (r < k) ? (idx ? __ldg(idx + (long)b * k + r) : r) : -1;
    }
    float v[4][NC], p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const long row = (long)b * ncam * hw + (s[j] >= 0 ? s[j] : 0);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            v[j][i] = __ldg(feat + row * E + (c < E ? c : 0));
        }
        p[j] = __ldg(pcd + row * 3 + (lane < 3 ? lane : 0));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (s[j] < 0) continue;
        const int r = r0 + wid + 8 * j;
        float* dst = tok + ((long)b * tok_rows + r) * E;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            if (c < E) dst[c] = bias ? __fadd_rn(v[j][i], __ldg(bias + c)) : v[j][i];
        }
        if (lane < 3) pos[((long)b * tok_rows + r) * 3 + lane] = p[j];
    }
}

template <int E>
__global__ void __launch_bounds__(256) gather_tokens_kernel(const float* __restrict__ feat, const float* __restrict__ pcd,
                                                            const int32_t* __restrict__ idx, int ncam, int hw, int k,
                                                            float* __restrict__ tok, float* __restrict__ pos,
                                                            int tok_rows, const float* __restrict__ bias) {
    __shared__ float tile[32][E + 1];
    __shared__ int src[32];
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * 32;
    const int t = threadIdx.x;
    if (t < 32) {
        const int r = r0 + t;
        src[t] = (r < k) ? (idx ? idx[(long)b * k + r] : r) : -1;
    }
    __syncthreads();
    const int lane = t & 31, wid = t >> 5;
    const int s = src[lane];
    if (s >= 0) {
        const int cam = s / hw, pix = s - cam * hw;
        const float* f = feat + ((long)(b * ncam + cam) * E) * hw + pix;
        for (int c = wid; c < E; c += 8) tile[lane][c] = __ldg(f + (long)c * hw);
    }
    __syncthreads();
    for (int i = t; i < 32 * E; i += 256) {
        const int r = i / E, c = i - r * E;
        if (r0 + r < k) tok[((long)b * tok_rows + r0 + r) * E + c] = bias ? __fadd_rn(tile[r][c], __ldg(bias + c)) : tile[r][c];
    }
    if (t < 96) {
        const int r = t / 3, c = t - 3 * r;
        if (src[r] >= 0) pos[((long)b * tok_rows + r0 + r) * 3 + c] = __ldg(pcd + ((long)b * ncam * hw + src[r]) * 3 + c);
    }
}

// =================================================================== mask logits
// logits[j][b][n] = <qvec[j][b], feat[b][n]>   (act3d.py:493-494).  One thread per ghost point, both query
// vectors at once; HBM-bound: reads the (B, Ng, E) feature tensor once (16-byte loads), writes nqv floats/point.
template <int E>
__global__ void __launch_bounds__(256) mask_logits_kernel(const float* __restrict__ feat, const float* __restrict__ qvec,
                                                          int batch, int ng, int nqv, float* __restrict__ logits) {
    __shared__ float q[4][E];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < nqv * E; i += blockDim.x) q[i / E][i % E] = qvec[((size_t)(i / E) * batch + b) * E + i % E];
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= ng) return;
    const float4* row = reinterpret_cast<const float4*>(feat + ((size_t)b * ng + n) * E);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c4 = 0; c4 < E / 4; ++c4) {
        const float4 v = __ldg(row + c4);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nqv) {
                acc[j] = fmaf(v.x, q[j][4 * c4], acc[j]);
                acc[j] = fmaf(v.y, q[j][4 * c4 + 1], acc[j]);
                acc[j] = fmaf(v.z, q[j][4 * c4 + 2], acc[j]);
                acc[j] = fmaf(v.w, q[j][4 * c4 + 3], acc[j]);
            }
    }
    for (int j = 0; j < nqv; ++j) logits[((size_t)j * batch + b) * ng + n] = acc[j];
}

// =================================================================== top ghost pick
__global__ void __launch_bounds__(1024) argmax_pick_kernel(const float* __restrict__ logits, const float* __restrict__ ghost,
                                                           int ng, int32_t* __restrict__ top_idx, float* __restrict__ pos) {
    __shared__ float sv[32];
    __shared__ int si[32];
    const int b = blockIdx.x;
    const float* l = logits + (long)b * ng;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < ng; i += blockDim.x) {
        const float v = __ldg(l + i);
        if (v > best || (v == best && i < bi)) {
            best = v;
            bi = i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) {
            best = ov;
            bi = oi;
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        sv[wid] = best;
        si[wid] = bi;
    }
    __syncthreads();
    if (wid == 0) {
        best = (lane < (int)(blockDim.x >> 5)) ? sv[lane] : -INFINITY;
        bi = (lane < (int)(blockDim.x >> 5)) ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) {
                best = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            if (bi == 0x7fffffff) bi = 0;   // all-NaN row: torch.max would also return an index; pick 0
            top_idx[b] = bi;
            pos[b * 3 + 0] = ghost[((long)b * ng + bi) * 3 + 0];
            pos[b * 3 + 1] = ghost[((long)b * ng + bi) * 3 + 1];
            pos[b * 3 + 2] = ghost[((long)b * ng + bi) * 3 + 2];
        }
    }
}

// =================================================================== ghost sampler (Philox4x32-10)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
}
__device__ __forceinline__ void philox4x32(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }   // [0,1)

struct Bounds6 {
    float lo[3], hi[3];
};

__global__ void __launch_bounds__(256) sample_ghost_kernel(const float* __restrict__ anchor, float radius, Bounds6 bd,
                                                           int batch, int ng, uint64_t seed, uint64_t stream_id,
                                                           const uint64_t* __restrict__ stream_base, float* __restrict__ out) {
    if (stream_base) stream_id += *stream_base;      // device-resident call counter: fresh points on every CUDA-graph replay
    const long total = (long)batch * ng;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int b = (int)(i / ng);
        float lo[3], hi[3], c[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[a] = bd.lo[a];
            hi[a] = bd.hi[a];
            c[a] = 0.f;
        }
        if (anchor) {   // ball of `radius` around the anchor, inside the clipped box (act3d.py:418-427)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                c[a] = anchor[b * 3 + a];
                lo[a] = fminf(fmaxf(c[a] - radius, bd.lo[a]), bd.hi[a]);
                hi[a] = fminf(fmaxf(c[a] + radius, bd.lo[a]), bd.hi[a]);
            }
        }
        float p[3] = {c[0], c[1], c[2]};
        for (uint32_t attempt = 0; attempt < 256u; ++attempt) {
            uint32_t ctr[4] = {(uint32_t)i, (uint32_t)(i >> 32), attempt, (uint32_t)stream_id};
            philox4x32(ctr, (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(stream_id >> 32));
#pragma unroll
            for (int a = 0; a < 3; ++a) p[a] = lo[a] + u01(ctr[a]) * (hi[a] - lo[a]);
            if (!anchor) break;
            const float dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2];
            if (sqrtf(dx * dx + dy * dy + dz * dz) < radius) break;   // utils.py:81-82
        }
        out[i * 3 + 0] = p[0];
        out[i * 3 + 1] = p[1];
        out[i * 3 + 2] = p[2];
    }
}

__global__ void counter_add_kernel(uint64_t* ctr, uint64_t inc) { *ctr += inc; }

}  // namespace a3d

// ===================================================================== C ABI
using namespace a3d;

extern "C" const char* a3d_last_error(void) { return g_err; }
extern "C" int a3d_abi_version(void) { return 1; }

extern "C" int a3d_pcd_pyramid(const float* pcd, int bn, int height, int width, int factor, float* out, void* stream) {
    A3D_REQUIRE(pcd && out && bn > 0, "a3d_pcd_pyramid: null pointer or empty batch");
    A3D_REQUIRE(factor == 2 || factor == 4 || factor == 8, "a3d_pcd_pyramid: factor must be 2, 4 or 8 (got %d)", factor);
    A3D_REQUIRE(height % factor == 0 && width % factor == 0, "a3d_pcd_pyramid: %dx%d not divisible by %d", height, width, factor);
    const long total = (long)bn * (height / factor) * (width / factor);
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    pcd_pyramid_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pcd, bn, height, width, factor, out);
    return check_launch("a3d_pcd_pyramid");
}

static int launch_topk(bool traj, const float* center, int traj_len, const float* pts, int batch, int n, int k,
                       int32_t* idx, float* dist, void* stream) {
    A3D_REQUIRE(center && pts && idx, "a3d_topk: null pointer");
    A3D_REQUIRE(batch > 0 && n > 0 && k > 0 && k <= n, "a3d_topk: need 0 < k <= n (k=%d n=%d)", k, n);
    A3D_REQUIRE(k <= kMaxK, "a3d_topk: k=%d exceeds %d", k, kMaxK);
    A3D_REQUIRE(n <= (1 << 24), "a3d_topk: n=%d exceeds 2^24", n);
    A3D_REQUIRE(!traj || (traj_len > 0 && traj_len <= 64), "a3d_traj_topk: trajectory length %d not in [1,64]", traj_len);
    int kp = 1;
    while (kp < k) kp <<= 1;
    const size_t smem = (size_t)kp * sizeof(unsigned long long);
    if (n <= kTopkCluster * kKeysPerThread * kTopkThreads && n >= 4096) {
        // cluster of 8 CTAs per sample, keys in registers (see topk_cluster_kernel)
        if (traj) {
            cudaFuncSetAttribute(topk_cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            topk_cluster_kernel<true><<<batch * kTopkCluster, kTopkThreads, smem, (cudaStream_t)stream>>>(center, traj_len, pts, n, k, idx, dist);
        } else {
            cudaFuncSetAttribute(topk_cluster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            topk_cluster_kernel<false><<<batch * kTopkCluster, kTopkThreads, smem, (cudaStream_t)stream>>>(center, 1, pts, n, k, idx, dist);
        }
        return check_launch("a3d_topk(cluster)");
    }
    if (traj) {
        cudaFuncSetAttribute(topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        topk_kernel<true><<<batch, kTopkThreads, smem, (cudaStream_t)stream>>>(center, traj_len, pts, n, k, idx, dist);
    } else {
        cudaFuncSetAttribute(topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        topk_kernel<false><<<batch, kTopkThreads, smem, (cudaStream_t)stream>>>(center, 1, pts, n, k, idx, dist);
    }
    return check_launch("a3d_topk");
}

extern "C" int a3d_local_topk(const float* center, const float* pts, int batch, int n, int k, int32_t* idx,
                              float* dist, void* stream) {
    return launch_topk(false, center, 1, pts, batch, n, k, idx, dist, stream);
}
extern "C" int a3d_traj_topk(const float* traj, int traj_len, const float* pts, int batch, int n, int k,
                             int32_t* idx, float* dist, void* stream) {
    return launch_topk(true, traj, traj_len, pts, batch, n, k, idx, dist, stream);
}

extern "C" int a3d_gather_tokens(const float* feat, const float* pcd, const int32_t* idx, int batch, int ncam,
                                 int embed, int hw, int k, float* tok, float* pos, int tok_rows, int channels_last,
                                 const float* feat_bias, void* stream) {
    A3D_REQUIRE(feat && pcd && tok && pos, "a3d_gather_tokens: null pointer");
    A3D_REQUIRE(batch > 0 && ncam > 0 && hw > 0 && k > 0 && k <= tok_rows, "a3d_gather_tokens: bad sizes (k=%d rows=%d)", k, tok_rows);
    A3D_REQUIRE(idx || k == ncam * hw, "a3d_gather_tokens: identity gather needs k == ncam*hw");
    dim3 grid((k + 31) / 32, batch);
    if (channels_last) {
        if (embed == 60)
            gather_tokens_nhwc_kernel<60><<<grid, 256, 0, (cudaStream_t)stream>>>(feat, pcd, idx, ncam, hw, k, tok, pos, tok_rows, feat_bias);
        else if (embed == 120)
            gather_tokens_nhwc_kernel<120><<<grid, 256, 0, (cudaStream_t)stream>>>(feat, pcd, idx, ncam, hw, k, tok, pos, tok_rows, feat_bias);
        else
            A3D_REQUIRE(false, "a3d_gather_tokens: embedding_dim %d not supported (60 or 120)", embed);
        return check_launch("a3d_gather_tokens");
    }
    if (embed == 60)
        gather_tokens_kernel<60><<<grid, 256, 0, (cudaStream_t)stream>>>(feat, pcd, idx, ncam, hw, k, tok, pos, tok_rows, feat_bias);
    else if (embed == 120)
        gather_tokens_kernel<120><<<grid, 256, 0, (cudaStream_t)stream>>>(feat, pcd, idx, ncam, hw, k, tok, pos, tok_rows, feat_bias);
    else
        A3D_REQUIRE(false, "a3d_gather_tokens: embedding_dim %d not supported (60 or 120)", embed);
    return check_launch("a3d_gather_tokens");
}

extern "C" int a3d_mask_logits(const float* feat, const float* qvec, int batch, int ng, int embed, int nqv,
                               float* logits, void* stream) {
    A3D_REQUIRE(feat && qvec && logits && batch > 0 && ng > 0, "a3d_mask_logits: bad arguments");
    A3D_REQUIRE(embed == 60, "a3d_mask_logits: built for embedding_dim 60 (got %d)", embed);
    A3D_REQUIRE(nqv >= 1 && nqv <= 4, "a3d_mask_logits: 1..4 query vectors supported (got %d)", nqv);
    A3D_REQUIRE(((uintptr_t)feat & 15) == 0, "a3d_mask_logits: features must be 16-byte aligned");
    dim3 grid((ng + 255) / 256, batch);
    mask_logits_kernel<60><<<grid, 256, 0, (cudaStream_t)stream>>>(feat, qvec, batch, ng, nqv, logits);
    return check_launch("a3d_mask_logits");
}

extern "C" int a3d_argmax_pick(const float* logits, const float* ghost, int batch, int ng, int32_t* top_idx,
                               float* pos, void* stream) {
    A3D_REQUIRE(logits && ghost && top_idx && pos && batch > 0 && ng > 0, "a3d_argmax_pick: bad arguments");
    argmax_pick_kernel<<<batch, 1024, 0, (cudaStream_t)stream>>>(logits, ghost, ng, top_idx, pos);
    return check_launch("a3d_argmax_pick");
}

extern "C" int a3d_counter_add(uint64_t* counter, uint64_t inc, void* stream) {
    A3D_REQUIRE(counter, "a3d_counter_add: null pointer");
    counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, inc);
    return check_launch("a3d_counter_add");
}

static int sample_ghost_impl(const float* anchor, float radius, const float* bounds_host, int batch, int ng, uint64_t seed,
                             uint64_t stream_id, const uint64_t* stream_base, float* out, void* stream);

extern "C" int a3d_sample_ghost(const float* anchor, float radius, const float* bounds_host, int batch, int ng,
                                uint64_t seed, uint64_t stream_id, float* out, void* stream) {
    return sample_ghost_impl(anchor, radius, bounds_host, batch, ng, seed, stream_id, nullptr, out, stream);
}
extern "C" int a3d_sample_ghost_ctr(const float* anchor, float radius, const float* bounds_host, int batch, int ng,
                                    uint64_t seed, uint64_t stream_id, const uint64_t* stream_base, float* out, void* stream) {
    A3D_REQUIRE(stream_base, "a3d_sample_ghost_ctr: null counter");
    return sample_ghost_impl(anchor, radius, bounds_host, batch, ng, seed, stream_id, stream_base, out, stream);
}

static int sample_ghost_impl(const float* anchor, float radius, const float* bounds_host, int batch, int ng, uint64_t seed,
                             uint64_t stream_id, const uint64_t* stream_base, float* out, void* stream) {
    A3D_REQUIRE(bounds_host && out && batch > 0 && ng > 0, "a3d_sample_ghost: bad arguments");
    A3D_REQUIRE(!anchor || radius > 0.f, "a3d_sample_ghost: radius must be positive with an anchor");
    Bounds6 bd;
    for (int a = 0; a < 3; ++a) {
        bd.lo[a] = bounds_host[a];
        bd.hi[a] = bounds_host[3 + a];
    }
    const long total = (long)batch * ng;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    sample_ghost_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(anchor, radius, bd, batch, ng, seed, stream_id, stream_base, out);
    return check_launch("a3d_sample_ghost");
}
