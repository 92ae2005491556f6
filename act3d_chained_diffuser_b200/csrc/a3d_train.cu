// Training path of libact3d_b200: the attention core with its backward, rotary apply (forward and
// transposed), and the token gather's backward.  All fp32 on the CUDA cores: training batches are two
// orders of magnitude smaller than the inference workload (333 ghost points per level instead of
// 16384) and gradients need fp32-class accuracy, so these kernels favour exactness and
// simplicity; the tensor-core kernels (a3d_xattn2.cu, cd_denoiser.cu) stay the inference path.
//
// Attention core = softmax(q k^T [+ key padding]) v per head, head_dim 15, without ever materialising the
// (B*H, Nq, Nk) score tensor the reference builds (multihead_custom_attention.py:391-415):
//   forward  : one lane per query row, 4 warps split every 64-key tile, online softmax, log-sum-exp saved;
//   backward : recomputes p = exp(s - lse) tile by tile (flash-attention style):
//                dq kernel  (row-parallel):  ds = p (dO.v - D),  dq += ds k,    D = dO.O
//                dkv kernel (key-parallel):  dv += p dO,         dk += ds q     (atomics over query chunks)
// Dropout on the attention weights (p = 0.1 in ChainedDiffuser training, multihead_custom_attention.py:413)
// is a counter-based Bernoulli keyed on (seed, b, h, row, key): the same mask is regenerated in backward.
#include "a3d_common.cuh"

namespace a3d {
namespace {

constexpr int HD = kHeadDim;       // 15
constexpr int kRows = 32;          // query rows per CTA (one per lane)
constexpr int kSplit = 4;          // warps per CTA; each takes 16 keys of every tile
constexpr int kTile = 64;          // keys per shared-memory tile
constexpr int kKeysPerWarp = kTile / kSplit;
constexpr int kQChunk = 256;       // query rows per dkv CTA (grid.z splits the rest)

__device__ __forceinline__ uint32_t drop_bits(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

struct Tile {
    float k[kTile][16];
    float v[kTile][16];
    int live[kTile];
};

// K/V rows [t0, t0+64) of head h, sample b -> shared memory (15 floats per row, slot 15 = 0)
__device__ __forceinline__ void load_tile(Tile& t, const float* __restrict__ k, const float* __restrict__ v,
                                          const unsigned char* __restrict__ mask, int b, int h, int nk, int E, int t0) {
    for (int i = threadIdx.x; i < kTile * 16; i += kRows * kSplit) {
        const int key = i >> 4, d = i & 15, gk = t0 + key;
        float kk = 0.f, vv = 0.f;
        if (gk < nk && d < HD) {
            const long off = ((long)b * nk + gk) * E + h * HD + d;
            kk = __ldg(k + off);
            vv = __ldg(v + off);
        }
        t.k[key][d] = kk;
        t.v[key][d] = vv;
    }
    if (threadIdx.x < kTile) {
        const int gk = t0 + threadIdx.x;
        t.live[threadIdx.x] = (gk < nk) && !(mask && mask[(long)b * nk + gk]);
    }
}

__device__ __forceinline__ float dot15(const float (&a)[HD], const float* __restrict__ row) {
    const float4 r0 = *reinterpret_cast<const float4*>(row), r1 = *reinterpret_cast<const float4*>(row + 4),
                 r2 = *reinterpret_cast<const float4*>(row + 8), r3 = *reinterpret_cast<const float4*>(row + 12);
    float s = a[0] * r0.x;
    s = fmaf(a[1], r0.y, s); s = fmaf(a[2], r0.z, s); s = fmaf(a[3], r0.w, s);
    s = fmaf(a[4], r1.x, s); s = fmaf(a[5], r1.y, s); s = fmaf(a[6], r1.z, s); s = fmaf(a[7], r1.w, s);
    s = fmaf(a[8], r2.x, s); s = fmaf(a[9], r2.y, s); s = fmaf(a[10], r2.z, s); s = fmaf(a[11], r2.w, s);
    s = fmaf(a[12], r3.x, s); s = fmaf(a[13], r3.y, s); s = fmaf(a[14], r3.z, s);
    return s;
}
__device__ __forceinline__ void axpy15(float (&acc)[HD], float a, const float* __restrict__ row) {
    const float4 r0 = *reinterpret_cast<const float4*>(row), r1 = *reinterpret_cast<const float4*>(row + 4),
                 r2 = *reinterpret_cast<const float4*>(row + 8), r3 = *reinterpret_cast<const float4*>(row + 12);
    acc[0] = fmaf(a, r0.x, acc[0]); acc[1] = fmaf(a, r0.y, acc[1]); acc[2] = fmaf(a, r0.z, acc[2]);
    acc[3] = fmaf(a, r0.w, acc[3]); acc[4] = fmaf(a, r1.x, acc[4]); acc[5] = fmaf(a, r1.y, acc[5]);
    acc[6] = fmaf(a, r1.z, acc[6]); acc[7] = fmaf(a, r1.w, acc[7]); acc[8] = fmaf(a, r2.x, acc[8]);
    acc[9] = fmaf(a, r2.y, acc[9]); acc[10] = fmaf(a, r2.z, acc[10]); acc[11] = fmaf(a, r2.w, acc[11]);
    acc[12] = fmaf(a, r3.x, acc[12]); acc[13] = fmaf(a, r3.y, acc[13]); acc[14] = fmaf(a, r3.z, acc[14]);
}

// ================================================================================ forward
// grid (ceil(nq/32), B*H), 128 threads.  q/o [B][nq][E], k/v [B][nk][E] (head h = columns h*15..h*15+14),
// lse [B*H][nq].
template <bool kDrop>
__global__ void __launch_bounds__(kRows* kSplit) attn_fwd_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const unsigned char* __restrict__ mask, int H, int nq, int nk, int E, float* __restrict__ o,
    float* __restrict__ lse, uint32_t drop_thresh, float keep_scale, uint64_t seed) {
    __shared__ __align__(16) Tile tile;
    __shared__ float part[kSplit][kRows][HD + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int row = blockIdx.x * kRows + lane;
    const bool live = row < nq;

    float qr[HD], acc[HD];
    {
        const float* qp = q + ((long)b * nq + (live ? row : 0)) * E + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            qr[d] = live ? __ldg(qp + d) : 0.f;
            acc[d] = 0.f;
        }
    }
    float m = -INFINITY, l = 0.f;
    const uint64_t drop_row = ((uint64_t)bh * nq + (uint64_t)(live ? row : 0)) * (uint64_t)nk;

    for (int t0 = 0; t0 < nk; t0 += kTile) {
        __syncthreads();
        load_tile(tile, k, v, mask, b, h, nk, E, t0);
        __syncthreads();
        float s[kKeysPerWarp];
        float cmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < kKeysPerWarp; ++j) {
            const int key = warp * kKeysPerWarp + j;
            const float a = dot15(qr, tile.k[key]);
            s[j] = tile.live[key] ? a : -INFINITY;
            cmax = fmaxf(cmax, s[j]);
        }
        if (cmax > m) {
            const float r = __expf(m - cmax);      // m = -inf -> 0
            l *= r;
#pragma unroll
            for (int d = 0; d < HD; ++d) acc[d] *= r;
            m = cmax;
        }
        if (m > -INFINITY) {
#pragma unroll
            for (int j = 0; j < kKeysPerWarp; ++j) {
                const int key = warp * kKeysPerWarp + j;
                float p = __expf(s[j] - m);        // masked key: exp(-inf) = 0
                l += p;
                if (kDrop) {
                    const bool keep = drop_bits(seed, drop_row + (uint64_t)(t0 + key)) >= drop_thresh;
                    p = keep ? p * keep_scale : 0.f;
                }
                axpy15(acc, p, tile.v[key]);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) part[warp][lane][d] = acc[d];
    part[warp][lane][HD] = m;
    part[warp][lane][HD + 1] = l;
    __syncthreads();
    if (warp == 0 && live) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < kSplit; ++w) M = fmaxf(M, part[w][lane][HD]);
        float L = 0.f, out[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) out[d] = 0.f;
#pragma unroll
        for (int w = 0; w < kSplit; ++w) {
            const float mw = part[w][lane][HD];
            const float sc = (mw == -INFINITY) ? 0.f : __expf(mw - M);
            L = fmaf(part[w][lane][HD + 1], sc, L);
#pragma unroll
            for (int d = 0; d < HD; ++d) out[d] = fmaf(part[w][lane][d], sc, out[d]);
        }
        const float inv = L > 0.f ? 1.f / L : 0.f;
        float* op = o + ((long)b * nq + row) * E + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) op[d] = out[d] * inv;
        lse[(long)bh * nq + row] = L > 0.f ? M + logf(L) : 0.f;
    }
}

// ================================================================================ backward: dq (+ D)
template <bool kDrop>
__global__ void __launch_bounds__(kRows* kSplit) attn_bwd_dq_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const unsigned char* __restrict__ mask, const float* __restrict__ o, const float* __restrict__ dout,
    const float* __restrict__ lse, int H, int nq, int nk, int E, float* __restrict__ dq, float* __restrict__ dsum,
    uint32_t drop_thresh, float keep_scale, uint64_t seed) {
    __shared__ __align__(16) Tile tile;
    __shared__ float part[kSplit][kRows][HD + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int row = blockIdx.x * kRows + lane;
    const bool live = row < nq;

    float qr[HD], gr[HD], acc[HD];
    float D = 0.f, ls = 0.f;
    {
        const long base = ((long)b * nq + (live ? row : 0)) * E + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            qr[d] = live ? __ldg(q + base + d) : 0.f;
            gr[d] = live ? __ldg(dout + base + d) : 0.f;
            const float ov = live ? __ldg(o + base + d) : 0.f;
            D = fmaf(gr[d], ov, D);
            acc[d] = 0.f;
        }
        if (live) ls = lse[(long)bh * nq + row];
    }
    const uint64_t drop_row = ((uint64_t)bh * nq + (uint64_t)(live ? row : 0)) * (uint64_t)nk;

    for (int t0 = 0; t0 < nk; t0 += kTile) {
        __syncthreads();
        load_tile(tile, k, v, mask, b, h, nk, E, t0);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < kKeysPerWarp; ++j) {
            const int key = warp * kKeysPerWarp + j;
            if (!tile.live[key]) continue;                     // uniform across the warp
            const float p = __expf(dot15(qr, tile.k[key]) - ls);
            float dp = dot15(gr, tile.v[key]);
            if (kDrop) {
                const bool keep = drop_bits(seed, drop_row + (uint64_t)(t0 + key)) >= drop_thresh;
                dp = keep ? dp * keep_scale : 0.f;
            }
            axpy15(acc, p * (dp - D), tile.k[key]);
        }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) part[warp][lane][d] = acc[d];
    __syncthreads();
    if (warp == 0 && live) {
        float* dp = dq + ((long)b * nq + row) * E + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) dp[d] = (part[0][lane][d] + part[1][lane][d]) + (part[2][lane][d] + part[3][lane][d]);
        dsum[(long)bh * nq + row] = D;
    }
}

// ================================================================================ backward: dk, dv
// grid (ceil(nk/128), B*H, query chunks), 128 threads = 128 keys.  dk/dv [B][nk][E] must be zero-filled by
// the caller; every CTA adds its chunk's contribution with fp32 atomics.
struct QTile {
    float q[kRows][16];
    float g[kRows][16];
    float lse[kRows];
    float dsum[kRows];
};

template <bool kDrop>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const unsigned char* __restrict__ mask, const float* __restrict__ dout, const float* __restrict__ lse,
    const float* __restrict__ dsum, int H, int nq, int nk, int E, float* __restrict__ dk, float* __restrict__ dv,
    uint32_t drop_thresh, float keep_scale, uint64_t seed) {
    __shared__ __align__(16) QTile qt;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int key = blockIdx.x * 128 + threadIdx.x;
    const bool valid = key < nk && !(mask && mask[(long)b * nk + key]);
    const int r_begin = blockIdx.z * kQChunk;
    const int r_end = min(nq, r_begin + kQChunk);

    float kr[HD], vr[HD], gk[HD], gv[HD];
    {
        const long base = ((long)b * nk + (key < nk ? key : 0)) * E + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            kr[d] = valid ? __ldg(k + base + d) : 0.f;
            vr[d] = valid ? __ldg(v + base + d) : 0.f;
            gk[d] = 0.f;
            gv[d] = 0.f;
        }
    }
    for (int r0 = r_begin; r0 < r_end; r0 += kRows) {
        __syncthreads();
        for (int i = threadIdx.x; i < kRows * 16; i += 128) {
            const int r = i >> 4, d = i & 15, grow = r0 + r;
            float qq = 0.f, gg = 0.f;
            if (grow < r_end && d < HD) {
                const long off = ((long)b * nq + grow) * E + h * HD + d;
                qq = __ldg(q + off);
                gg = __ldg(dout + off);
            }
            qt.q[r][d] = qq;
            qt.g[r][d] = gg;
        }
        if (threadIdx.x < kRows) {
            const int grow = r0 + threadIdx.x;
            qt.lse[threadIdx.x] = grow < r_end ? lse[(long)bh * nq + grow] : 0.f;
            qt.dsum[threadIdx.x] = grow < r_end ? dsum[(long)bh * nq + grow] : 0.f;
        }
        __syncthreads();
        const int nrows = min(kRows, r_end - r0);
        if (valid) {
            for (int r = 0; r < nrows; ++r) {
                const float p = __expf(dot15(kr, qt.q[r]) - qt.lse[r]);
                float dp = dot15(vr, qt.g[r]);
                float pd = p;
                if (kDrop) {
                    const uint64_t idx = ((uint64_t)bh * nq + (uint64_t)(r0 + r)) * (uint64_t)nk + (uint64_t)key;
                    const bool keep = drop_bits(seed, idx) >= drop_thresh;
                    pd = keep ? p * keep_scale : 0.f;
                    dp = keep ? dp * keep_scale : 0.f;
                }
                axpy15(gv, pd, qt.g[r]);
                axpy15(gk, p * (dp - qt.dsum[r]), qt.q[r]);
            }
        }
    }
    if (valid) {
        const long base = ((long)b * nk + key) * E + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            atomicAdd(dk + base + d, gk[d]);
            atomicAdd(dv + base + d, gv[d]);
        }
    }
}

// ================================================================================ rotary apply
// out[2i] = x[2i] c - x[2i+1] s,  out[2i+1] = x[2i+1] c + x[2i] s   (position_encodings.py:31-34) with the
// angle of pair i taken from xyz as in RotaryPositionEncoding3D (position_encodings.py:58-97).
// sign = -1 applies the transposed (= inverse) rotation: the backward of sign = +1.
template <int E>
__global__ void __launch_bounds__(256) rope_apply_kernel(const float* __restrict__ x, const float* __restrict__ pos,
                                                         long rows, float sign, float* __restrict__ out) {
    constexpr int P = E / 2;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * P) return;
    const long row = i / P;
    const int pr = (int)(i - row * P);
    const int axis = (2 * pr) / (E / 3);
    const int j = pr - axis * (E / 6);
    const float ang = __fmul_rn(__ldg(pos + row * 3 + axis), rope_freq<E>(j));
    float s, c;
    sincosf(ang, &s, &c);
    s *= sign;
    const float2 xv = *reinterpret_cast<const float2*>(x + row * E + 2 * pr);
    float2 ov;
    ov.x = xv.x * c - xv.y * s;
    ov.y = xv.y * c + xv.x * s;
    *reinterpret_cast<float2*>(out + row * E + 2 * pr) = ov;
}

// ================================================================================ token gather backward
// dfeat[(b*ncam + cam)][c][pix] (NCHW) or [(b*ncam + cam)][pix][c] (NHWC) += dtok[b][r][c] for r < k, where
// (cam, pix) = divmod(idx[b][r] or r, hw).  top-k indices are unique per sample, so plain adds would do;
// atomics keep the kernel correct for arbitrary index lists.
__global__ void __launch_bounds__(256) gather_tokens_bwd_kernel(const float* __restrict__ dtok, const int32_t* __restrict__ idx,
                                                                int ncam, int E, int hw, int k, int tok_rows,
                                                                int channels_last, float* __restrict__ dfeat, long total) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % E);
    const long br = i / E;
    const int r = (int)(br % k);
    const int b = (int)(br / k);
    const int src = idx ? idx[(long)b * k + r] : r;
    const int cam = src / hw, pix = src - cam * hw;
    const float g = dtok[((long)b * tok_rows + r) * E + c];
    const long img = (long)b * ncam + cam;
    const long off = channels_last ? (img * hw + pix) * E + c : (img * E + c) * hw + pix;
    atomicAdd(dfeat + off, g);
}

// ================================================================================ soft cross-entropy over ghost points
// Per sample b (one CTA): label = softmax_n(-||ghost_n - gt|| / spread) (+ label smoothing),
// loss_b = -sum_n label_n log_softmax(logits)_n, dlogits_n = softmax(logits)_n - label_n
// (LossAndMetrics._compute_position_loss, main_keypose.py:387-403: F.cross_entropy with probability targets).
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
    v = is_max ? warp_max(v) : warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                       // red[] may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    float r = lane < nw ? red[lane] : (is_max ? -INFINITY : 0.f);
    r = is_max ? warp_max(r) : warp_sum(r);
    return r;
}

__global__ void __launch_bounds__(256) soft_ce_kernel(const float* __restrict__ logits, const float* __restrict__ ghost,
                                                      const float* __restrict__ gt, int ng, float inv_spread,
                                                      float smoothing, float* __restrict__ loss,
                                                      float* __restrict__ dlogits) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    const float gx = gt[b * 3], gy = gt[b * 3 + 1], gz = gt[b * 3 + 2];
    const float* lg = logits + (long)b * ng;
    const float* gp = ghost + (long)b * ng * 3;
    float amax = -INFINITY, xmax = -INFINITY;
    for (int n = threadIdx.x; n < ng; n += blockDim.x) {
        const float dx = gp[n * 3] - gx, dy = gp[n * 3 + 1] - gy, dz = gp[n * 3 + 2] - gz;
        const float a = -sqrtf(dx * dx + dy * dy + dz * dz) * inv_spread;
        amax = fmaxf(amax, a);
        xmax = fmaxf(xmax, lg[n]);
    }
    amax = block_reduce(amax, true, red);
    xmax = block_reduce(xmax, true, red);
    float sa = 0.f, sx = 0.f;
    for (int n = threadIdx.x; n < ng; n += blockDim.x) {
        const float dx = gp[n * 3] - gx, dy = gp[n * 3 + 1] - gy, dz = gp[n * 3 + 2] - gz;
        const float a = -sqrtf(dx * dx + dy * dy + dz * dz) * inv_spread;
        sa += expf(a - amax);
        sx += expf(lg[n] - xmax);
    }
    sa = block_reduce(sa, false, red);
    sx = block_reduce(sx, false, red);
    const float log_sx = logf(sx), inv_sa = 1.f / sa, inv_sx = 1.f / sx, uni = smoothing / (float)ng;
    float acc = 0.f;
    for (int n = threadIdx.x; n < ng; n += blockDim.x) {
        const float dx = gp[n * 3] - gx, dy = gp[n * 3 + 1] - gy, dz = gp[n * 3 + 2] - gz;
        const float a = -sqrtf(dx * dx + dy * dy + dz * dz) * inv_spread;
        const float t = expf(a - amax) * inv_sa * (1.f - smoothing) + uni;
        const float z = lg[n] - xmax;
        acc -= t * (z - log_sx);
        if (dlogits) dlogits[(long)b * ng + n] = expf(z) * inv_sx - t;
    }
    acc = block_reduce(acc, false, red);
    if (threadIdx.x == 0) loss[b] = acc;
}

// ================================================================================ weight / bias gradient of a linear layer
// dW[o][i] = sum_r dY[r][o] X[r][i],  db[o] = sum_r dY[r][o]   for y = x W^T + b with R rows (tokens) in the tens of
// thousands and O, I <= a few hundred: a GEMM whose output is one or a few 64 x 64 tiles and whose reduction
// dimension is huge.  cuBLAS runs it as a single-CTA SIMT sgemm (0.1 ms per call at R = 66 k: 19 % of a training
// step).  Here the rows are split over the grid in slices of 128: every CTA reduces its slice into a 64 x 64 register
// tile and writes it to a workspace, and a second small kernel sums the slices in a fixed order (deterministic, no
// atomics) -- one pass over dY and X at HBM speed plus a few MB of partials.
constexpr int kWgRows = 128;      // rows per CTA
constexpr int kWgBlk = 32;        // rows per shared-memory block

__global__ void __launch_bounds__(256) wgrad_partial_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            long rows, int O, int I, float* __restrict__ part_w,
                                                            float* __restrict__ part_b) {
    __shared__ __align__(16) float ys[kWgBlk][64 + 4];
    __shared__ __align__(16) float xs[kWgBlk][64 + 4];
    const int o0 = blockIdx.y * 64, i0 = blockIdx.z * 64;
    const long r_begin = (long)blockIdx.x * kWgRows;
    const long r_end = r_begin + kWgRows < rows ? r_begin + kWgRows : rows;
    const int to = threadIdx.x >> 4, ti = threadIdx.x & 15;       // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    for (long r0 = r_begin; r0 < r_end; r0 += kWgBlk) {
        __syncthreads();
        for (int e = threadIdx.x; e < kWgBlk * 64; e += 256) {
            const int r = e >> 6, c = e & 63;
            const long gr = r0 + r;
            ys[r][c] = (gr < r_end && o0 + c < O) ? __ldg(dy + gr * O + o0 + c) : 0.f;
            xs[r][c] = (gr < r_end && i0 + c < I) ? __ldg(x + gr * I + i0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < kWgBlk; ++r) {
            const float4 yv = *reinterpret_cast<const float4*>(&ys[r][4 * to]);
            const float4 xv = *reinterpret_cast<const float4*>(&xs[r][4 * ti]);
            const float ya[4] = {yv.x, yv.y, yv.z, yv.w}, xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(ya[a], xa[c], acc[a][c]);
                bsum[a] += ya[a];
            }
        }
    }
    float* pw = part_w + (size_t)blockIdx.x * O * I;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int o = o0 + 4 * to + a;
        if (o >= O) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = i0 + 4 * ti + c;
            if (i < I) pw[(size_t)o * I + i] = acc[a][c];
        }
        if (part_b && blockIdx.z == 0 && ti == 0) part_b[(size_t)blockIdx.x * O + o] = bsum[a];
    }
}

// out[e] = sum over slices of part[slice][e]: CTAs [0, ceil(nw/32)) reduce the weight partials, the rest the bias
// partials.  256 threads = 32 elements x 8 slice groups (group w takes slices w, w+8, ...: eight loads in flight per
// thread instead of one thread walking all ~500 slices), groups combined through shared memory in a fixed order, so
// the result is deterministic.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part_w, long nw, float* __restrict__ dw,
                                                           const float* __restrict__ part_b, long nb, float* __restrict__ db,
                                                           long slices) {
    __shared__ float red[8][33];
    const long wblocks = (nw + 31) / 32;
    const bool bias = blockIdx.x >= wblocks;
    const float* part = bias ? part_b : part_w;
    const long n = bias ? nb : nw;
    float* out = bias ? db : dw;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const long e = (bias ? (long)blockIdx.x - wblocks : (long)blockIdx.x) * 32 + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (e < n) {
        long k = grp;
        for (; k + 24 < slices; k += 32) {
            s0 += __ldg(part + (k + 0) * n + e);
            s1 += __ldg(part + (k + 8) * n + e);
            s2 += __ldg(part + (k + 16) * n + e);
            s3 += __ldg(part + (k + 24) * n + e);
        }
        for (; k < slices; k += 8) s0 += __ldg(part + k * n + e);
    }
    red[grp][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (grp == 0 && e < n) {
        float s = red[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) s += red[w][lane];
        out[e] = s;
    }
}

int drop_args(float p, uint32_t* thresh, float* scale) {
    if (p <= 0.f) {
        *thresh = 0;
        *scale = 1.f;
        return 0;
    }
    const double t = (double)p * 4294967296.0;
    *thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    *scale = 1.f / (1.f - p);
    return 1;
}

}  // namespace

// a3d_set_option("train_attn_core", 0 | 1): 0 = tensor-core kernels (a3d_train_mma.cu, default), 1 = the fp32 CUDA-core
// kernels of this file (kept as the A/B reference)
int g_train_attn_core = 0;
int a3d_launch_attn_fwd_mma(const float* q, const float* k, const float* v, const unsigned char* key_mask, int batch,
                            int heads, int nq, int nk, int embed, float* o, float* lse, int drop, uint32_t th, float sc,
                            uint64_t seed, cudaStream_t st);
int a3d_launch_attn_bwd_mma(const float* q, const float* k, const float* v, const unsigned char* key_mask,
                            const float* o, const float* dout, const float* lse, int batch, int heads, int nq, int nk,
                            int embed, float* dq, float* dk, float* dv, float* dsum, int drop, uint32_t th, float sc,
                            uint64_t seed, cudaStream_t st);
}  // namespace a3d

using namespace a3d;

extern "C" int a3d_attn_fwd(const float* q, const float* k, const float* v, const unsigned char* key_mask, int batch,
                            int heads, int nq, int nk, int embed, float* o, float* lse, float dropout_p,
                            uint64_t seed, void* stream) {
    A3D_REQUIRE(q && k && v && o && lse, "a3d_attn_fwd: null pointer");
    A3D_REQUIRE(batch > 0 && heads > 0 && nq > 0 && nk > 0, "a3d_attn_fwd: bad sizes");
    A3D_REQUIRE(embed == heads * HD, "a3d_attn_fwd: built for head_dim 15 (embed=%d heads=%d)", embed, heads);
    A3D_REQUIRE(batch * heads <= 65535, "a3d_attn_fwd: batch*heads=%d exceeds 65535", batch * heads);
    A3D_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "a3d_attn_fwd: dropout_p=%f not in [0,1)", dropout_p);
    uint32_t th;
    float sc;
    const int drop = drop_args(dropout_p, &th, &sc);
    if (g_train_attn_core == 0)
        return a3d_launch_attn_fwd_mma(q, k, v, key_mask, batch, heads, nq, nk, embed, o, lse, drop, th, sc, seed, (cudaStream_t)stream);
    dim3 grid((nq + kRows - 1) / kRows, batch * heads);
    if (drop)
        attn_fwd_kernel<true><<<grid, kRows * kSplit, 0, (cudaStream_t)stream>>>(q, k, v, key_mask, heads, nq, nk, embed, o, lse, th, sc, seed);
    else
        attn_fwd_kernel<false><<<grid, kRows * kSplit, 0, (cudaStream_t)stream>>>(q, k, v, key_mask, heads, nq, nk, embed, o, lse, th, sc, seed);
    return check_launch("a3d_attn_fwd");
}

extern "C" int a3d_attn_bwd(const float* q, const float* k, const float* v, const unsigned char* key_mask,
                            const float* o, const float* dout, const float* lse, int batch, int heads, int nq, int nk,
                            int embed, float* dq, float* dk, float* dv, float* dsum, float dropout_p, uint64_t seed,
                            void* stream) {
    A3D_REQUIRE(q && k && v && o && dout && lse && dq && dk && dv && dsum, "a3d_attn_bwd: null pointer");
    A3D_REQUIRE(batch > 0 && heads > 0 && nq > 0 && nk > 0, "a3d_attn_bwd: bad sizes");
    A3D_REQUIRE(embed == heads * HD, "a3d_attn_bwd: built for head_dim 15 (embed=%d heads=%d)", embed, heads);
    A3D_REQUIRE(batch * heads <= 65535, "a3d_attn_bwd: batch*heads=%d exceeds 65535", batch * heads);
    A3D_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "a3d_attn_bwd: dropout_p=%f not in [0,1)", dropout_p);
    uint32_t th;
    float sc;
    const int drop = drop_args(dropout_p, &th, &sc);
    cudaStream_t st = (cudaStream_t)stream;
    if (g_train_attn_core == 0)
        return a3d_launch_attn_bwd_mma(q, k, v, key_mask, o, dout, lse, batch, heads, nq, nk, embed, dq, dk, dv, dsum, drop, th, sc, seed, st);
    dim3 g1((nq + kRows - 1) / kRows, batch * heads);
    const int chunks = (nq + kQChunk - 1) / kQChunk;
    A3D_REQUIRE(chunks <= 65535, "a3d_attn_bwd: nq=%d too large", nq);
    dim3 g2((nk + 127) / 128, batch * heads, chunks);
    if (drop) {
        attn_bwd_dq_kernel<true><<<g1, kRows * kSplit, 0, st>>>(q, k, v, key_mask, o, dout, lse, heads, nq, nk, embed, dq, dsum, th, sc, seed);
        attn_bwd_dkv_kernel<true><<<g2, 128, 0, st>>>(q, k, v, key_mask, dout, lse, dsum, heads, nq, nk, embed, dk, dv, th, sc, seed);
    } else {
        attn_bwd_dq_kernel<false><<<g1, kRows * kSplit, 0, st>>>(q, k, v, key_mask, o, dout, lse, heads, nq, nk, embed, dq, dsum, th, sc, seed);
        attn_bwd_dkv_kernel<false><<<g2, 128, 0, st>>>(q, k, v, key_mask, dout, lse, dsum, heads, nq, nk, embed, dk, dv, th, sc, seed);
    }
    return check_launch("a3d_attn_bwd");
}

extern "C" int a3d_rope_apply(const float* x, const float* pos, long rows, int embed, int transpose, float* out,
                              void* stream) {
    A3D_REQUIRE(x && pos && out && rows > 0, "a3d_rope_apply: bad arguments");
    A3D_REQUIRE(((uintptr_t)x & 7) == 0 && ((uintptr_t)out & 7) == 0, "a3d_rope_apply: buffers must be 8-byte aligned");
    const float sign = transpose ? -1.f : 1.f;
    const long pairs = rows * (embed / 2);
    const unsigned blocks = (unsigned)((pairs + 255) / 256);
    if (embed == 60)
        rope_apply_kernel<60><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, pos, rows, sign, out);
    else if (embed == 120)
        rope_apply_kernel<120><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, pos, rows, sign, out);
    else
        A3D_REQUIRE(false, "a3d_rope_apply: embedding_dim %d not supported (60 or 120)", embed);
    return check_launch("a3d_rope_apply");
}

extern "C" int a3d_gather_tokens_bwd(const float* dtok, const int32_t* idx, int batch, int ncam, int embed, int hw,
                                     int k, int tok_rows, int channels_last, float* dfeat, void* stream) {
    A3D_REQUIRE(dtok && dfeat, "a3d_gather_tokens_bwd: null pointer");
    A3D_REQUIRE(batch > 0 && ncam > 0 && hw > 0 && k > 0 && k <= tok_rows && embed > 0,
                "a3d_gather_tokens_bwd: bad sizes (k=%d rows=%d)", k, tok_rows);
    A3D_REQUIRE(idx || k == ncam * hw, "a3d_gather_tokens_bwd: identity gather needs k == ncam*hw");
    const long total = (long)batch * k * embed;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    gather_tokens_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dtok, idx, ncam, embed, hw, k, tok_rows,
                                                                      channels_last, dfeat, total);
    return check_launch("a3d_gather_tokens_bwd");
}

extern "C" int a3d_soft_ce(const float* logits, const float* ghost, const float* gt, int batch, int ng, float spread,
                           float label_smoothing, float* loss, float* dlogits, void* stream) {
    A3D_REQUIRE(logits && ghost && gt && loss && batch > 0 && ng > 0, "a3d_soft_ce: bad arguments");
    A3D_REQUIRE(spread > 0.f && label_smoothing >= 0.f && label_smoothing < 1.f, "a3d_soft_ce: bad spread / smoothing");
    soft_ce_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(logits, ghost, gt, ng, 1.f / spread, label_smoothing, loss, dlogits);
    return check_launch("a3d_soft_ce");
}

extern "C" size_t a3d_linear_wgrad_workspace(long rows, int out_features, int in_features) {
    const long slices = (rows + kWgRows - 1) / kWgRows;
    return (size_t)slices * ((size_t)out_features * in_features + out_features) * sizeof(float);
}

extern "C" int a3d_linear_wgrad(const float* dy, const float* x, long rows, int out_features, int in_features, float* dw,
                                float* db, void* workspace, void* stream) {
    A3D_REQUIRE(dy && x && dw && workspace && rows > 0 && out_features > 0 && in_features > 0, "a3d_linear_wgrad: bad arguments");
    const long slices = (rows + kWgRows - 1) / kWgRows;
    A3D_REQUIRE(slices <= 2147483647L, "a3d_linear_wgrad: too many rows");
    float* part_w = (float*)workspace;
    float* part_b = part_w + (size_t)slices * out_features * in_features;
    dim3 grid((unsigned)slices, (out_features + 63) / 64, (in_features + 63) / 64);
    cudaStream_t st = (cudaStream_t)stream;
    wgrad_partial_kernel<<<grid, 256, 0, st>>>(dy, x, rows, out_features, in_features, part_w, db ? part_b : nullptr);
    const long nw = (long)out_features * in_features;
    const long nb = db ? out_features : 0;
    wgrad_reduce_kernel<<<(unsigned)((nw + 31) / 32 + (nb + 31) / 32), 256, 0, st>>>(part_w, nw, dw, part_b, nb, db, slices);
    return check_launch("a3d_linear_wgrad");
}
