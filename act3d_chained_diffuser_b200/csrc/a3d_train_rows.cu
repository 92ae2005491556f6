// Row-wise training kernels: the per-token linear layers and LayerNorms of the attention stacks
// (layers.py:300-332 RelativeCrossAttentionLayer / FeedforwardLayer, :115-218 ParallelAttentionLayer) for the tens of
// thousands of context tokens a training step pushes through them.  Under autograd these were library calls: fp32
// SIMT sgemm (TF32 is off, as in the reference) at ~20 us per [65 k x 60] x [60 x 60] product and ATen's LayerNorm
// kernels at 36 us forward / 50 us backward per call -- 4x to 6x above the time the 31 MB of traffic needs.
//
//   a3d_linear_fwd   y = x W^T + b [ReLU]   or, transposed, dx = dy W   (the data gradient of the same layer)
//       tensor cores, error-compensated fp16 pairs (a3d_mma_gemm.cuh: fp32-class accuracy).  Training weights change
//       every step, so a tiny kernel first packs W (or W^T) into the B-fragment order the GEMM consumes; the GEMM
//       kernel stages 64 rows of x as (hi, lo) planes once and walks the output columns 64 at a time.
//   a3d_layernorm_fwd   z = x + res,  y = LayerNorm(z) g + b,  saves mean / rstd      (one pass over x, res; y, z out)
//   a3d_layernorm_bwd   dz from (dy, z, mean, rstd, g); dg, db through per-CTA partials summed in a fixed order
//       (wgrad_reduce_kernel's scheme: deterministic, no atomics).
#include "a3d_mma_gemm.cuh"

namespace a3d {
namespace {

// ------------------------------------------------------------------------------------------------ weight packing
// logical weight Wl[n][k] (n: output column, k: contraction):  Wl = W (O x I) or, transposed, Wl[n][k] = W[k][n].
// out[(ks * ntiles + nt) * 32 + lane] = {b0_hi, b1_hi, b0_lo, b1_lo},  n = 8 nt + lane / 4,  k0 = 16 ks + 2 (lane % 4)
__global__ void __launch_bounds__(256) pack_frag_kernel(const float* __restrict__ w, int O, int I, int transpose, int ksteps,
                                                        int ntiles, uint4* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ksteps * ntiles * 32) return;
    const int lane = idx & 31, nt = (idx >> 5) % ntiles, ks = (idx >> 5) / ntiles;
    const int n = 8 * nt + (lane >> 2), k0 = 16 * ks + 2 * (lane & 3);
    const int N = transpose ? I : O, K = transpose ? O : I;
    auto at = [&](int nn, int kk) -> float {
        if (nn >= N || kk >= K) return 0.f;
        return transpose ? __ldg(w + (long)kk * I + nn) : __ldg(w + (long)nn * I + kk);
    };
    uint4 r;
    split_h2(at(n, k0), at(n, k0 + 1), r.x, r.z);
    split_h2(at(n, k0 + 8), at(n, k0 + 9), r.y, r.w);
    out[idx] = r;
}

// ------------------------------------------------------------------------------------------------ linear
// grid = ceil(rows / 64), 256 threads: warp = (row tile 0..3, column half 0..1); per pass 64 output columns.
template <int KSTEPS>
__global__ void __launch_bounds__(256) linear_rows_kernel(const float* __restrict__ x, long rows, int K, int N, int npad,
                                                          const uint4* __restrict__ wfrag, const float* __restrict__ bias,
                                                          int relu, float* __restrict__ y) {
    constexpr int PITCH = 16 * KSTEPS + 8;
    extern __shared__ __align__(16) unsigned char smem[];
    __half* ah = reinterpret_cast<__half*>(smem);
    __half* al = ah + 64 * PITCH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long r0 = (long)blockIdx.x * 64;
    // K is even for every layer of the models (60, 120, 480): float2 loads (row pitch K * 4 B is a multiple of 8).
    // Addresses are clamped and the value selected afterwards, so the loads of an unrolled batch issue back to back.
    constexpr int PAIRS = 8 * KSTEPS;                   // float2 per padded row
    const long last_row = rows - 1;
    const int last_pair = K / 2 - 1;
#pragma unroll 4
    for (int i = tid; i < 64 * PAIRS; i += 256) {
        const int r = i / PAIRS, cp = i - r * PAIRS;
        const long gr = r0 + r;
        const float2 t = __ldg(reinterpret_cast<const float2*>(x + (gr < rows ? gr : last_row) * K) + (cp <= last_pair ? cp : last_pair));
        const bool ok = gr < rows && cp <= last_pair;
        uint32_t hi, lo;
        split_h2(ok ? t.x : 0.f, ok ? t.y : 0.f, hi, lo);
        *reinterpret_cast<uint32_t*>(ah + r * PITCH + 2 * cp) = hi;
        *reinterpret_cast<uint32_t*>(al + r * PITCH + 2 * cp) = lo;
    }
    __syncthreads();
    const int m0 = 16 * (warp & 3), half = warp >> 2, g = lane >> 2, q = lane & 3;
    const int ntiles = npad / 8;
    for (int c0 = 0; c0 < npad; c0 += 64) {
        float acc[4][4];
        const int nt0 = c0 / 8 + 4 * half;
        mma_gemm_split<KSTEPS, 4, PITCH>(ah, al, m0, wfrag, ntiles, nt0, lane, acc);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int col = 8 * (nt0 + n) + 2 * q;
            if (col >= N) continue;                       // N is even: col + 1 < N as well
            const float b0 = bias ? __ldg(bias + col) : 0.f, b1 = bias ? __ldg(bias + col + 1) : 0.f;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const long row = r0 + m0 + g + 8 * rr;
                if (row >= rows) continue;
                float v0 = acc[n][2 * rr] + b0, v1 = acc[n][2 * rr + 1] + b1;
                if (relu) {
                    v0 = fmaxf(v0, 0.f);
                    v1 = fmaxf(v1, 0.f);
                }
                *reinterpret_cast<float2*>(y + row * N + col) = make_float2(v0, v1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// LPR lanes cooperate on one row, 4 consecutive channels per lane (float4): E = 60 -> 16 lanes (15 active),
// E = 120 -> 32 lanes (30 active).
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int E, int LPR>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            long rows, float eps, float* __restrict__ z, float* __restrict__ y,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    constexpr int RPW = 32 / LPR;                      // rows per warp
    const int lane = threadIdx.x & 31, sub = lane % LPR;
    const long row = ((long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
    const bool live = row < rows && 4 * sub < E;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        v = __ldg(reinterpret_cast<const float4*>(x + row * E) + sub);
        if (res) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(res + row * E) + sub);
            v.x += r.x, v.y += r.y, v.z += r.z, v.w += r.w;
            reinterpret_cast<float4*>(z + row * E)[sub] = v;
        }
    }
    const float mean = group_sum<LPR>((v.x + v.y) + (v.z + v.w)) * (1.0f / E);
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    const float var = group_sum<LPR>(live ? (dx * dx + dy * dy) + (dz * dz + dw * dw) : 0.f) * (1.0f / E);
    const float rstd = 1.0f / sqrtf(var + eps);
    if (live) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + sub), bt = __ldg(reinterpret_cast<const float4*>(beta) + sub);
        reinterpret_cast<float4*>(y + row * E)[sub] = make_float4(fmaf(dx * rstd, gm.x, bt.x), fmaf(dy * rstd, gm.y, bt.y),
                                                                  fmaf(dz * rstd, gm.z, bt.z), fmaf(dw * rstd, gm.w, bt.w));
        if (sub == 0) {
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
    }
}

// dz = rstd (dy g - mean_c(dy g) - xhat mean_c(dy g xhat));  per-CTA partial sums of dg = dy xhat, db = dy
template <int E, int LPR>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                            const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                            const float* __restrict__ gamma, long rows, float* __restrict__ dz,
                                                            float* __restrict__ part_g, float* __restrict__ part_b) {
    constexpr int RPW = 32 / LPR;
    __shared__ float red[8][2][E];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % LPR;
    const bool col_ok = 4 * sub < E;
    float4 gm = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok) gm = __ldg(reinterpret_cast<const float4*>(gamma) + sub);
    float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
    const long stride = (long)gridDim.x * 8 * RPW;
    const long iters = (rows + stride - 1) / stride;          // the same for every thread: the shuffles below need full warps
    for (long it = 0; it < iters; ++it) {
        const long row = ((long)blockIdx.x * 8 + warp) * RPW + lane / LPR + it * stride;
        const bool live = row < rows && col_ok;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f), v = g;
        float mean = 0.f, rstd = 0.f;
        if (live) {
            g = __ldg(reinterpret_cast<const float4*>(dy + row * E) + sub);
            v = __ldg(reinterpret_cast<const float4*>(z + row * E) + sub);
            mean = __ldg(mean_in + row);
            rstd = __ldg(rstd_in + row);
        }
        const float xh[4] = {(v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd};
        const float gy[4] = {g.x, g.y, g.z, g.w};
        const float gw[4] = {g.x * gm.x, g.y * gm.y, g.z * gm.z, g.w * gm.w};
        const float c1 = group_sum<LPR>((gw[0] + gw[1]) + (gw[2] + gw[3])) * (1.0f / E);
        const float c2 = group_sum<LPR>((gw[0] * xh[0] + gw[1] * xh[1]) + (gw[2] * xh[2] + gw[3] * xh[3])) * (1.0f / E);
        if (live) {
            reinterpret_cast<float4*>(dz + row * E)[sub] = make_float4(rstd * (gw[0] - c1 - xh[0] * c2), rstd * (gw[1] - c1 - xh[1] * c2),
                                                                       rstd * (gw[2] - c1 - xh[2] * c2), rstd * (gw[3] - c1 - xh[3] * c2));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ag[j] = fmaf(gy[j], xh[j], ag[j]);
                ab[j] += gy[j];
            }
        }
    }
    // rows of one warp that share columns, then the 8 warps, in a fixed order
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            ag[j] += __shfl_xor_sync(0xffffffffu, ag[j], o);
            ab[j] += __shfl_xor_sync(0xffffffffu, ab[j], o);
        }
    if (lane < LPR && col_ok)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[warp][0][4 * sub + j] = ag[j];
            red[warp][1][4 * sub + j] = ab[j];
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) {
        const int which = i / E, c = i - which * E;
        float s = red[0][which][c];
#pragma unroll
        for (int w = 1; w < 8; ++w) s += red[w][which][c];
        (which ? part_b : part_g)[(long)blockIdx.x * E + c] = s;
    }
}

// out[c] = sum over CTAs of part[cta][c], fixed order (one warp per 32 columns, 8 slices in flight per lane)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ part_g, const float* __restrict__ part_b, int n,
                                                     int slices, float* __restrict__ dg, float* __restrict__ db) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int blocks_g = (n + 31) / 32;
    const bool bias = (int)blockIdx.x >= blocks_g;
    const float* part = bias ? part_b : part_g;
    float* out = bias ? db : dg;
    const int e = (bias ? blockIdx.x - blocks_g : blockIdx.x) * 32 + lane;
    float s = 0.f;
    if (e < n)
        for (int k = grp; k < slices; k += 8) s += __ldg(part + (long)k * n + e);
    red[grp][lane] = s;
    __syncthreads();
    if (grp == 0 && e < n) {
        float t = red[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) t += red[w][lane];
        out[e] = t;
    }
}

constexpr int kLnBwdCtas = 296;

}  // namespace
}  // namespace a3d

using namespace a3d;

static int linear_ksteps(int k) {
    const int ks = (k + 15) / 16;
    return (ks <= 4) ? 4 : (ks <= 8) ? 8 : (ks <= 15) ? 15 : (ks <= 30) ? 30 : 0;
}

// 1 when a3d_linear_fwd supports (out_features, in_features, transpose_w)
extern "C" int a3d_linear_supported(int out_features, int in_features, int transpose_w) {
    const int K = transpose_w ? out_features : in_features, N = transpose_w ? in_features : out_features;
    return linear_ksteps(K) != 0 && (K % 2) == 0 && (N % 2) == 0 && N <= 1024;
}

extern "C" size_t a3d_linear_workspace(int out_features, int in_features, int transpose_w) {
    const int K = transpose_w ? out_features : in_features, N = transpose_w ? in_features : out_features;
    const int ks = linear_ksteps(K), npad = (N + 63) / 64 * 64;
    return (size_t)ks * (npad / 8) * 32 * sizeof(uint4);
}

extern "C" int a3d_linear_fwd(const float* x, const float* w, const float* bias, long rows, int out_features,
                              int in_features, int relu, int transpose_w, float* y, void* workspace, void* stream) {
    A3D_REQUIRE(x && w && y && workspace && rows > 0, "a3d_linear_fwd: bad arguments");
    A3D_REQUIRE(a3d_linear_supported(out_features, in_features, transpose_w),
                "a3d_linear_fwd: (out, in, transpose) = (%d, %d, %d) not supported", out_features, in_features, transpose_w);
    A3D_REQUIRE(((uintptr_t)x & 7) == 0 && ((uintptr_t)y & 7) == 0 && ((uintptr_t)workspace & 15) == 0, "a3d_linear_fwd: alignment");
    const int K = transpose_w ? out_features : in_features, N = transpose_w ? in_features : out_features;
    const int ks = linear_ksteps(K), npad = (N + 63) / 64 * 64, ntiles = npad / 8;
    cudaStream_t st = (cudaStream_t)stream;
    uint4* frag = (uint4*)workspace;
    const int total = ks * ntiles * 32;
    pack_frag_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, out_features, in_features, transpose_w, ks, ntiles, frag);
    const unsigned grid = (unsigned)((rows + 63) / 64);
    const size_t smem = (size_t)2 * 64 * (16 * ks + 8) * sizeof(__half);
#define A3D_LAUNCH_LINEAR(KS)                                                                                          \
    {                                                                                                                  \
        static PerDeviceOnce once_dev;                                                                                 \
        if (bool& once = once_dev.flag(); !once) {                                                                     \
            cudaFuncSetAttribute(linear_rows_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
            once = true;                                                                                               \
        }                                                                                                              \
        linear_rows_kernel<KS><<<grid, 256, smem, st>>>(x, rows, K, N, npad, frag, bias, relu, y);                     \
    }
    if (ks == 4) A3D_LAUNCH_LINEAR(4)
    else if (ks == 8) A3D_LAUNCH_LINEAR(8)
    else if (ks == 15) A3D_LAUNCH_LINEAR(15)
    else A3D_LAUNCH_LINEAR(30)
#undef A3D_LAUNCH_LINEAR
    return check_launch("a3d_linear_fwd");
}

extern "C" int a3d_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, long rows,
                                 int embed, float eps, float* z, float* y, float* mean, float* rstd, void* stream) {
    A3D_REQUIRE(x && gamma && beta && y && mean && rstd && rows > 0, "a3d_layernorm_fwd: bad arguments");
    A3D_REQUIRE(!res || z, "a3d_layernorm_fwd: a residual input needs the z = x + res output");
    A3D_REQUIRE(embed == 60 || embed == 120, "a3d_layernorm_fwd: embed=%d not supported (60 or 120)", embed);
    A3D_REQUIRE((((uintptr_t)x | (uintptr_t)res | (uintptr_t)z | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0,
                "a3d_layernorm_fwd: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (embed == 60)
        layernorm_fwd_kernel<60, 16><<<(unsigned)((rows + 15) / 16), 256, 0, st>>>(x, res, gamma, beta, rows, eps, z, y, mean, rstd);
    else
        layernorm_fwd_kernel<120, 32><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, res, gamma, beta, rows, eps, z, y, mean, rstd);
    return check_launch("a3d_layernorm_fwd");
}

extern "C" size_t a3d_layernorm_bwd_workspace(int embed) { return (size_t)2 * kLnBwdCtas * embed * sizeof(float); }

extern "C" int a3d_layernorm_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                                 long rows, int embed, float* dz, float* dgamma, float* dbeta, void* workspace, void* stream) {
    A3D_REQUIRE(dy && z && mean && rstd && gamma && dz && dgamma && dbeta && workspace && rows > 0, "a3d_layernorm_bwd: bad arguments");
    A3D_REQUIRE(embed == 60 || embed == 120, "a3d_layernorm_bwd: embed=%d not supported (60 or 120)", embed);
    A3D_REQUIRE((((uintptr_t)dy | (uintptr_t)z | (uintptr_t)dz | (uintptr_t)gamma) & 15) == 0, "a3d_layernorm_bwd: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    float* part_g = (float*)workspace;
    float* part_b = part_g + (size_t)kLnBwdCtas * embed;
    const long per_cta = embed == 60 ? 16 : 8;
    const int ctas = (int)((rows + per_cta - 1) / per_cta < kLnBwdCtas ? (rows + per_cta - 1) / per_cta : kLnBwdCtas);
    if (embed == 60)
        layernorm_bwd_kernel<60, 16><<<ctas, 256, 0, st>>>(dy, z, mean, rstd, gamma, rows, dz, part_g, part_b);
    else
        layernorm_bwd_kernel<120, 32><<<ctas, 256, 0, st>>>(dy, z, mean, rstd, gamma, rows, dz, part_g, part_b);
    colsum_kernel<<<2 * ((embed + 31) / 32), 256, 0, st>>>(part_g, part_b, embed, ctas, dgamma, dbeta);
    return check_launch("a3d_layernorm_bwd");
}
