// Training attention core on the tensor cores: softmax(q k^T [+ key padding]) v per head (head_dim 15 padded
// to 16) with its backward, flash-attention style (no (B*H, Nq, Nk) score tensor;
// multihead_custom_attention.py:391-415 builds three of them per layer).
//
// Why mma.sync and not tcgen05: a training batch has 333 ghost points per level, i.e. 21
// 16-row tiles per (sample, head); the kernels are bound by staging K/V tiles, not by the MMA rate, and the
// register-resident accumulators of mma.sync let the softmax, its rescaling and the gradient algebra run on the
// fragments without a TMEM round trip.  The inference kernels (a3d_xattn6.cu) are the tcgen05 path.
//
// Accuracy: every product runs as an error-compensated fp16 pair, x = hi + lo 2^-11 with hi = fp16(x),
// lo = fp16((x - hi) 2^11) (the scaling keeps the residual of small values, e.g. softmax weights of 1/4150, out of
// fp16's subnormal range): a*b ~= ah*bh + (ah*bl + al*bh) 2^-11 on m16n8k16 tensor-core MMAs, the main and the
// correction products in separate fp32 accumulators (relative error ~2^-21 per product, against 2^-11 for a plain fp16
// operand), because the golden-gradient tests compare with the fp32 reference at 1e-3 relative per parameter after 18
// such cores.  Gradients (dO) can be far below fp16's normal
// range, so they are rescaled by a power of two first: per query row in the dq kernel, per (sample, head) -- from
// the maximum the dq kernel leaves in the workspace -- in the dk/dv kernel; the scale is exact and undone at the end.
//
// One staged tile = 64 keys (or 64 queries in the dk/dv kernel), converted once per CTA into fp16 (hi, lo) planes in
// the two layouts the B operand needs: row-major [row][dim pair] (pitch 12 words) for products contracted over the
// head dimension, transposed [dim][row pair] (pitch 36 words) for products contracted over rows.  Both pitches
// make the fragment loads bank-conflict free.
#include <cuda_fp16.h>

#include "a3d_common.cuh"

namespace a3d {
namespace {

constexpr int HD = kHeadDim;   // 15
constexpr int TK = 64;         // rows per staged tile
constexpr int RP = 12;         // words per row, row-major plane
constexpr int TP = 36;         // words per row, transposed plane
constexpr int kQChunk = 256;   // query rows per dk/dv CTA (grid.z covers the rest)

struct Plane {
    uint32_t h[TK * RP];
    uint32_t l[TK * RP];
};
struct PlaneT {
    uint32_t h[16 * TP];
    uint32_t l[16 * TP];
};

__device__ __forceinline__ uint32_t drop_bits(uint64_t seed, uint64_t idx) {      // same hash as a3d_train.cu
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

// packed conversions (cvt.rn.f16x2.f32 = one F2FP on the ALU pipe); the scalar cvt.rn.f16.f32 is a quarter-rate F2F
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn((a - f.x) * 2048.f, (b - f.y) * 2048.f);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
constexpr float kLoScale = 1.0f / 2048.0f;
// c += ah * bh,  e += ah * bl + al * bh   (a * b ~= c + e 2^-11; the al * bl term is below fp32 resolution)
__device__ __forceinline__ void mma3(float (&c)[4], float (&e)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     uint32_t b0h, uint32_t b1h, uint32_t b0l, uint32_t b1l) {
    mma16816(e, al, b0h, b1h);
    mma16816(e, ah, b0l, b1l);
    mma16816(c, ah, b0h, b1h);
}
__device__ __forceinline__ void fold(float (&c)[4], const float (&e)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = fmaf(e[j], kLoScale, c[j]);
}
// The tensor core aligns and TRUNCATES the addends of its fp32 accumulation: a running sum kept in the MMA accumulator
// over the ~65 key tiles of a training context drifts by ~1e-5 relative (measured; the bias is one-sided).  So every
// staged tile is accumulated from zero inside the MMAs (4 k-steps) and added to the running sum with a rounded FADD.
__device__ __forceinline__ void add_tile(float (&acc)[4], const float (&c)[4], const float (&e)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] += fmaf(e[j], kLoScale, c[j]);
}
// B fragment words of n-tile `nt` from a row-major plane (contraction over the head dimension)
// (c is overwritten with the finished product)
__device__ __forceinline__ void mma3_rows(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const Plane& p,
                                          int nt, int g, int t) {
    const int w = (nt * 8 + g) * RP + t;
    float e[4] = {0.f, 0.f, 0.f, 0.f};
    c[0] = c[1] = c[2] = c[3] = 0.f;
    mma3(c, e, ah, al, p.h[w], p.h[w + 4], p.l[w], p.l[w + 4]);
    fold(c, e);
}
// B fragment words of dim-tile `dt`, rows [16 kk, 16 kk + 16) from a transposed plane (contraction over rows)
// (accumulates: c main, e correction; fold(c, e) once at the end)
__device__ __forceinline__ void mma3_cols(float (&c)[4], float (&e)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                          const PlaneT& p, int dt, int kk, int g, int t) {
    const int w = (dt * 8 + g) * TP + kk * 8 + t;
    mma3(c, e, ah, al, p.h[w], p.h[w + 4], p.l[w], p.l[w + 4]);
}
// accumulator fragments of two adjacent n-tiles -> A fragment (hi, lo) of the following product
__device__ __forceinline__ void acc_to_a(const float (&c0)[4], const float (&c1)[4], uint32_t (&ah)[4], uint32_t (&al)[4]) {
    split2(c0[0], c0[1], ah[0], al[0]);
    split2(c0[2], c0[3], ah[1], al[1]);
    split2(c1[0], c1[1], ah[2], al[2]);
    split2(c1[2], c1[3], ah[3], al[3]);
}

// Staging of rows [r0, r0 + 64) of head h of src[.][n][E] (already offset to the sample), times `scale`, in two steps so
// that the global loads of tile i+1 are in flight while tile i is being multiplied: fetch() -> registers (addresses
// clamped, so the eight loads of a thread issue back to back without branches), commit() -> fp16 (hi, lo) planes.
// Rows >= r_end and the pad column read as 0.
struct Fetch {
    float x[2][4];
};
__device__ __forceinline__ void fetch(Fetch& f, const float* __restrict__ src, int E, int hoff, int r0, int r_end, float scale) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int i = threadIdx.x + 128 * it;
        const int dp = i & 7, jp = i >> 3;
        const int r = r0 + 2 * jp, d = 2 * dp, d1 = min(d + 1, HD - 1);
        const float* pa = src + (long)min(r, r_end - 1) * E + hoff;
        const float* pb = src + (long)min(r + 1, r_end - 1) * E + hoff;
        const float a0 = __ldg(pa + d), a1 = __ldg(pa + d1), b0 = __ldg(pb + d), b1 = __ldg(pb + d1);
        const bool va = r < r_end, vb = r + 1 < r_end, vd = d + 1 < HD;
        f.x[it][0] = va ? a0 * scale : 0.f;
        f.x[it][1] = (va && vd) ? a1 * scale : 0.f;
        f.x[it][2] = vb ? b0 * scale : 0.f;
        f.x[it][3] = (vb && vd) ? b1 * scale : 0.f;
    }
}
template <bool kRowMajor, bool kTrans>
__device__ __forceinline__ void commit(const Fetch& f, Plane* __restrict__ pr, PlaneT* __restrict__ pt) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int i = threadIdx.x + 128 * it;
        const int dp = i & 7, jp = i >> 3, d = 2 * dp;
        if (kRowMajor) {
            split2(f.x[it][0], f.x[it][1], pr->h[(2 * jp) * RP + dp], pr->l[(2 * jp) * RP + dp]);
            split2(f.x[it][2], f.x[it][3], pr->h[(2 * jp + 1) * RP + dp], pr->l[(2 * jp + 1) * RP + dp]);
        }
        if (kTrans) {
            split2(f.x[it][0], f.x[it][2], pt->h[d * TP + jp], pt->l[d * TP + jp]);
            split2(f.x[it][1], f.x[it][3], pt->h[(d + 1) * TP + jp], pt->l[(d + 1) * TP + jp]);
        }
    }
}

// A fragment (rows rA = row0 + g, rB = rA + 8; columns 2t, 2t+1, 2t+8, 2t+9) of x[.][n][E] * (sA | sB), raw values too
__device__ __forceinline__ void load_rows(const float* __restrict__ x, int E, int hoff, int rA, int rB, int n, int t,
                                          float (&va)[4], float (&vb)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int d = 2 * t + (j & 1) + (j >> 1) * 8;
        va[j] = (rA < n && d < HD) ? __ldg(x + (long)rA * E + hoff + d) : 0.f;
        vb[j] = (rB < n && d < HD) ? __ldg(x + (long)rB * E + hoff + d) : 0.f;
    }
}
__device__ __forceinline__ void rows_to_a(const float (&va)[4], const float (&vb)[4], float sa, float sb, uint32_t (&ah)[4],
                                          uint32_t (&al)[4]) {
    split2(va[0] * sa, va[1] * sa, ah[0], al[0]);
    split2(vb[0] * sb, vb[1] * sb, ah[1], al[1]);
    split2(va[2] * sa, va[3] * sa, ah[2], al[2]);
    split2(vb[2] * sb, vb[3] * sb, ah[3], al[3]);
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
// power of two s with max * s in [1, 2); 1 for max = 0 / non-finite
__device__ __forceinline__ float pow2_scale(float mx, float* inv) {
    if (!(mx > 0.f) || !isfinite(mx)) {
        *inv = 1.f;
        return 1.f;
    }
    int e;
    frexpf(mx, &e);                      // mx = m 2^e, m in [0.5, 1)
    e = max(-100, min(100, e));
    *inv = scalbnf(1.f, e - 1);
    return scalbnf(1.f, 1 - e);
}

__device__ __forceinline__ void stage_live(int* live, const unsigned char* __restrict__ mask, int b, int nk, int t0) {
    if (threadIdx.x < TK) {
        const int gk = t0 + threadIdx.x;
        live[threadIdx.x] = (gk < nk) && !(mask && mask[(long)b * nk + gk]);
    }
}

// ================================================================================ forward
// grid (ceil(nq/64), B*H), 128 threads: warp w owns query rows [64 x + 16 w, +16) and walks every key tile.
template <bool kDrop>
__global__ void __launch_bounds__(128) attn_fwd_mma_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const unsigned char* __restrict__ mask, int H, int nq, int nk, int E, float* __restrict__ o,
    float* __restrict__ lse, uint32_t drop_thresh, float keep_scale, uint64_t seed) {
    __shared__ Plane ks;
    __shared__ PlaneT vt;
    __shared__ __align__(8) int live[TK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H, hoff = h * HD;
    const int row0 = blockIdx.x * 64 + warp * 16, rA = row0 + g, rB = rA + 8;
    const bool active = row0 < nq;
    const float* kb = k + (long)b * nk * E;
    const float* vb = v + (long)b * nk * E;

    uint32_t qh[4], ql[4];
    {
        float va[4], vb_[4];
        load_rows(q + (long)b * nq * E, E, hoff, rA, rB, nq, t, va, vb_);
        rows_to_a(va, vb_, 1.f, 1.f, qh, ql);
    }
    float acc[2][4] = {};
    float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;
    const uint64_t dropA = ((uint64_t)bh * nq + (uint64_t)min(rA, nq - 1)) * (uint64_t)nk;
    const uint64_t dropB = ((uint64_t)bh * nq + (uint64_t)min(rB, nq - 1)) * (uint64_t)nk;

    Fetch fk, fv;
    fetch(fk, kb, E, hoff, 0, nk, 1.f);
    fetch(fv, vb, E, hoff, 0, nk, 1.f);
    for (int t0 = 0; t0 < nk; t0 += TK) {
        __syncthreads();
        commit<true, false>(fk, &ks, nullptr);
        commit<false, true>(fv, nullptr, &vt);
        stage_live(live, mask, b, nk, t0);
        __syncthreads();
        if (t0 + TK < nk) {
            fetch(fk, kb, E, hoff, t0 + TK, nk, 1.f);
            fetch(fv, vb, E, hoff, t0 + TK, nk, 1.f);
        }
        if (!active) continue;
        float s[8][4];
        float tA = -INFINITY, tB = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mma3_rows(s[nt], qh, ql, ks, nt, g, t);
            const int2 lv = *reinterpret_cast<const int2*>(&live[nt * 8 + 2 * t]);
            if (!lv.x) s[nt][0] = s[nt][2] = -INFINITY;
            if (!lv.y) s[nt][1] = s[nt][3] = -INFINITY;
            tA = fmaxf(tA, fmaxf(s[nt][0], s[nt][1]));
            tB = fmaxf(tB, fmaxf(s[nt][2], s[nt][3]));
        }
        const float nA = fmaxf(mA, quad_max(tA)), nB = fmaxf(mB, quad_max(tB));
        const float cA = (nA == -INFINITY) ? 1.f : __expf(mA - nA), cB = (nB == -INFINITY) ? 1.f : __expf(mB - nB);
        mA = nA;
        mB = nB;
        lA *= cA;
        lB *= cB;
#pragma unroll
        for (int dt = 0; dt < 2; ++dt) {
            acc[dt][0] *= cA;
            acc[dt][1] *= cA;
            acc[dt][2] *= cB;
            acc[dt][3] *= cB;
        }
        const float sA = (nA == -INFINITY) ? 0.f : nA, sB = (nB == -INFINITY) ? 0.f : nB;   // s = -inf -> exp = 0 either way
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float p0 = __expf(s[nt][0] - sA), p1 = __expf(s[nt][1] - sA), p2 = __expf(s[nt][2] - sB), p3 = __expf(s[nt][3] - sB);
            lA += p0 + p1;
            lB += p2 + p3;
            if (kDrop) {
                const uint64_t key = (uint64_t)(t0 + nt * 8 + 2 * t);
                p0 = drop_bits(seed, dropA + key) >= drop_thresh ? p0 * keep_scale : 0.f;
                p1 = drop_bits(seed, dropA + key + 1) >= drop_thresh ? p1 * keep_scale : 0.f;
                p2 = drop_bits(seed, dropB + key) >= drop_thresh ? p2 * keep_scale : 0.f;
                p3 = drop_bits(seed, dropB + key + 1) >= drop_thresh ? p3 * keep_scale : 0.f;
            }
            s[nt][0] = p0;
            s[nt][1] = p1;
            s[nt][2] = p2;
            s[nt][3] = p3;
        }
        float ta[2][4] = {}, tc[2][4] = {};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t ph[4], pl[4];
            acc_to_a(s[2 * kk], s[2 * kk + 1], ph, pl);
            mma3_cols(ta[0], tc[0], ph, pl, vt, 0, kk, g, t);
            mma3_cols(ta[1], tc[1], ph, pl, vt, 1, kk, g, t);
        }
        add_tile(acc[0], ta[0], tc[0]);
        add_tile(acc[1], ta[1], tc[1]);
    }
    if (!active) return;
    lA = quad_sum(lA);
    lB = quad_sum(lB);
    const float iA = lA > 0.f ? 1.f / lA : 0.f, iB = lB > 0.f ? 1.f / lB : 0.f;
#pragma unroll
    for (int dt = 0; dt < 2; ++dt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int d = dt * 8 + 2 * t + j;
            if (d < HD) {
                if (rA < nq) o[((long)b * nq + rA) * E + hoff + d] = acc[dt][j] * iA;
                if (rB < nq) o[((long)b * nq + rB) * E + hoff + d] = acc[dt][2 + j] * iB;
            }
        }
    if (t == 0) {
        if (rA < nq) lse[(long)bh * nq + rA] = lA > 0.f ? mA + logf(lA) : 0.f;
        if (rB < nq) lse[(long)bh * nq + rB] = lB > 0.f ? mB + logf(lB) : 0.f;
    }
}

// ================================================================================ backward: dq, D = dO.O, max |dO|
// Same tiling as the forward.  p = exp(s - lse), dp = dO v^T, ds = p (dp - D), dq = ds k.
template <bool kDrop>
__global__ void __launch_bounds__(128) attn_bwd_dq_mma_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const unsigned char* __restrict__ mask, const float* __restrict__ o, const float* __restrict__ dout,
    const float* __restrict__ lse, int H, int nq, int nk, int E, float* __restrict__ dq, float* __restrict__ dsum,
    float* __restrict__ gmax, uint32_t drop_thresh, float keep_scale, uint64_t seed) {
    __shared__ Plane ks;
    __shared__ Plane vs;
    __shared__ PlaneT kt;
    __shared__ __align__(8) int live[TK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H, hoff = h * HD;
    const int row0 = blockIdx.x * 64 + warp * 16, rA = row0 + g, rB = rA + 8;
    const bool active = row0 < nq;
    const float* kb = k + (long)b * nk * E;
    const float* vb = v + (long)b * nk * E;

    uint32_t qh[4], ql[4], gh[4], gl[4];
    float DA, DB, invA, invB, lsA = 0.f, lsB = 0.f;
    {
        float va[4], vb_[4], oa[4], ob[4];
        load_rows(q + (long)b * nq * E, E, hoff, rA, rB, nq, t, va, vb_);
        rows_to_a(va, vb_, 1.f, 1.f, qh, ql);
        load_rows(dout + (long)b * nq * E, E, hoff, rA, rB, nq, t, va, vb_);
        load_rows(o + (long)b * nq * E, E, hoff, rA, rB, nq, t, oa, ob);
        float da = 0.f, db = 0.f, xa = 0.f, xb = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            da = fmaf(va[j], oa[j], da);
            db = fmaf(vb_[j], ob[j], db);
            xa = fmaxf(xa, fabsf(va[j]));
            xb = fmaxf(xb, fabsf(vb_[j]));
        }
        da = quad_sum(da);
        db = quad_sum(db);
        xa = quad_max(xa);
        xb = quad_max(xb);
        const float sa = pow2_scale(xa, &invA), sb = pow2_scale(xb, &invB);
        rows_to_a(va, vb_, sa, sb, gh, gl);
        DA = da * sa;
        DB = db * sb;
        if (rA < nq) lsA = lse[(long)bh * nq + rA];
        if (rB < nq) lsB = lse[(long)bh * nq + rB];
        if (t == 0) {
            if (rA < nq) dsum[(long)bh * nq + rA] = da;
            if (rB < nq) dsum[(long)bh * nq + rB] = db;
        }
        float wm = fmaxf(isfinite(xa) ? xa : 0.f, isfinite(xb) ? xb : 0.f);
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, off));
        if (lane == 0 && wm > 0.f) atomicMax(reinterpret_cast<unsigned int*>(gmax + bh), __float_as_uint(wm));
    }
    float acc[2][4] = {};
    const uint64_t dropA = ((uint64_t)bh * nq + (uint64_t)min(rA, nq - 1)) * (uint64_t)nk;
    const uint64_t dropB = ((uint64_t)bh * nq + (uint64_t)min(rB, nq - 1)) * (uint64_t)nk;

    Fetch fk, fv;
    fetch(fk, kb, E, hoff, 0, nk, 1.f);
    fetch(fv, vb, E, hoff, 0, nk, 1.f);
    for (int t0 = 0; t0 < nk; t0 += TK) {
        __syncthreads();
        commit<true, true>(fk, &ks, &kt);
        commit<true, false>(fv, &vs, nullptr);
        stage_live(live, mask, b, nk, t0);
        __syncthreads();
        if (t0 + TK < nk) {
            fetch(fk, kb, E, hoff, t0 + TK, nk, 1.f);
            fetch(fv, vb, E, hoff, t0 + TK, nk, 1.f);
        }
        if (!active) continue;
        float ds[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float s[4], dp[4];
            mma3_rows(s, qh, ql, ks, nt, g, t);
            mma3_rows(dp, gh, gl, vs, nt, g, t);
            const int2 lv = *reinterpret_cast<const int2*>(&live[nt * 8 + 2 * t]);
            const float p0 = lv.x ? __expf(s[0] - lsA) : 0.f, p1 = lv.y ? __expf(s[1] - lsA) : 0.f;
            const float p2 = lv.x ? __expf(s[2] - lsB) : 0.f, p3 = lv.y ? __expf(s[3] - lsB) : 0.f;
            if (kDrop) {
                const uint64_t key = (uint64_t)(t0 + nt * 8 + 2 * t);
                dp[0] = drop_bits(seed, dropA + key) >= drop_thresh ? dp[0] * keep_scale : 0.f;
                dp[1] = drop_bits(seed, dropA + key + 1) >= drop_thresh ? dp[1] * keep_scale : 0.f;
                dp[2] = drop_bits(seed, dropB + key) >= drop_thresh ? dp[2] * keep_scale : 0.f;
                dp[3] = drop_bits(seed, dropB + key + 1) >= drop_thresh ? dp[3] * keep_scale : 0.f;
            }
            ds[nt][0] = p0 * (dp[0] - DA);
            ds[nt][1] = p1 * (dp[1] - DA);
            ds[nt][2] = p2 * (dp[2] - DB);
            ds[nt][3] = p3 * (dp[3] - DB);
        }
        float ta[2][4] = {}, tc[2][4] = {};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t ah[4], al[4];
            acc_to_a(ds[2 * kk], ds[2 * kk + 1], ah, al);
            mma3_cols(ta[0], tc[0], ah, al, kt, 0, kk, g, t);
            mma3_cols(ta[1], tc[1], ah, al, kt, 1, kk, g, t);
        }
        add_tile(acc[0], ta[0], tc[0]);
        add_tile(acc[1], ta[1], tc[1]);
    }
    if (!active) return;
#pragma unroll
    for (int dt = 0; dt < 2; ++dt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int d = dt * 8 + 2 * t + j;
            if (d < HD) {
                if (rA < nq) dq[((long)b * nq + rA) * E + hoff + d] = acc[dt][j] * invA;
                if (rB < nq) dq[((long)b * nq + rB) * E + hoff + d] = acc[dt][2 + j] * invB;
            }
        }
}

// ================================================================================ backward: dk, dv
// grid (ceil(nk/64), B*H, query chunks), 128 threads: warp w owns keys [64 x + 16 w, +16) as the M dimension and
// walks the chunk's query tiles:  s^T = k q^T, p^T = exp(s^T - lse), dv += p^T dO, dp^T = v dO^T,
// ds^T = p^T (dp^T - D), dk += ds^T q.  dk / dv are zero-filled by the caller; chunks add with fp32 atomics.
template <bool kDrop>
__global__ void __launch_bounds__(128) attn_bwd_dkv_mma_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const unsigned char* __restrict__ mask, const float* __restrict__ dout, const float* __restrict__ lse,
    const float* __restrict__ dsum, const float* __restrict__ gmax, int H, int nq, int nk, int E,
    float* __restrict__ dk, float* __restrict__ dv, uint32_t drop_thresh, float keep_scale, uint64_t seed) {
    __shared__ Plane qs;
    __shared__ Plane gs;
    __shared__ PlaneT qt;
    __shared__ PlaneT gt;
    __shared__ __align__(8) float lse_s[TK];
    __shared__ __align__(8) float dsum_s[TK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H, hoff = h * HD;
    const int key0 = blockIdx.x * 64 + warp * 16, kA = key0 + g, kB = kA + 8;
    const bool validA = kA < nk && !(mask && mask[(long)b * nk + kA]);
    const bool validB = kB < nk && !(mask && mask[(long)b * nk + kB]);
    const bool active = key0 < nk;
    const int r_begin = blockIdx.z * kQChunk, r_end = min(nq, r_begin + kQChunk);
    const float* qb = q + (long)b * nq * E;
    const float* gb = dout + (long)b * nq * E;

    uint32_t kh[4], kl[4], vh[4], vl[4];
    {
        float va[4], vb_[4];
        load_rows(k + (long)b * nk * E, E, hoff, kA, kB, nk, t, va, vb_);
        rows_to_a(va, vb_, 1.f, 1.f, kh, kl);
        load_rows(v + (long)b * nk * E, E, hoff, kA, kB, nk, t, va, vb_);
        rows_to_a(va, vb_, 1.f, 1.f, vh, vl);
    }
    float inv;
    const float scale = pow2_scale(gmax[bh], &inv);
    float gk[2][4] = {}, gv[2][4] = {};

    Fetch fq, fg;
    fetch(fq, qb, E, hoff, r_begin, r_end, 1.f);
    fetch(fg, gb, E, hoff, r_begin, r_end, scale);
    for (int r0 = r_begin; r0 < r_end; r0 += TK) {
        __syncthreads();
        commit<true, true>(fq, &qs, &qt);
        commit<true, true>(fg, &gs, &gt);
        if (threadIdx.x < TK) {
            const int r = r0 + threadIdx.x;
            lse_s[threadIdx.x] = r < r_end ? lse[(long)bh * nq + r] : INFINITY;       // exp(s - inf) = 0
            dsum_s[threadIdx.x] = r < r_end ? dsum[(long)bh * nq + r] * scale : 0.f;
        }
        __syncthreads();
        if (r0 + TK < r_end) {
            fetch(fq, qb, E, hoff, r0 + TK, r_end, 1.f);
            fetch(fg, gb, E, hoff, r0 + TK, r_end, scale);
        }
        if (!active) continue;
        float tk[2][4] = {}, tv[2][4] = {}, ck[2][4] = {}, cv[2][4] = {};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float pd[2][4], ds[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int nt = 2 * kk + u;
                float s[4], dp[4];
                mma3_rows(s, kh, kl, qs, nt, g, t);
                mma3_rows(dp, vh, vl, gs, nt, g, t);
                const float2 ls = *reinterpret_cast<const float2*>(&lse_s[nt * 8 + 2 * t]);
                const float2 dd = *reinterpret_cast<const float2*>(&dsum_s[nt * 8 + 2 * t]);
                const float p0 = validA ? __expf(s[0] - ls.x) : 0.f, p1 = validA ? __expf(s[1] - ls.y) : 0.f;
                const float p2 = validB ? __expf(s[2] - ls.x) : 0.f, p3 = validB ? __expf(s[3] - ls.y) : 0.f;
                pd[u][0] = p0;
                pd[u][1] = p1;
                pd[u][2] = p2;
                pd[u][3] = p3;
                if (kDrop) {
                    const uint64_t r = (uint64_t)bh * nq + (uint64_t)(r0 + nt * 8 + 2 * t);
                    const bool k0 = drop_bits(seed, r * nk + kA) >= drop_thresh, k1 = drop_bits(seed, (r + 1) * nk + kA) >= drop_thresh;
                    const bool k2 = drop_bits(seed, r * nk + kB) >= drop_thresh, k3 = drop_bits(seed, (r + 1) * nk + kB) >= drop_thresh;
                    pd[u][0] = k0 ? p0 * keep_scale : 0.f;
                    pd[u][1] = k1 ? p1 * keep_scale : 0.f;
                    pd[u][2] = k2 ? p2 * keep_scale : 0.f;
                    pd[u][3] = k3 ? p3 * keep_scale : 0.f;
                    dp[0] = k0 ? dp[0] * keep_scale : 0.f;
                    dp[1] = k1 ? dp[1] * keep_scale : 0.f;
                    dp[2] = k2 ? dp[2] * keep_scale : 0.f;
                    dp[3] = k3 ? dp[3] * keep_scale : 0.f;
                }
                ds[u][0] = p0 * (dp[0] - dd.x);
                ds[u][1] = p1 * (dp[1] - dd.y);
                ds[u][2] = p2 * (dp[2] - dd.x);
                ds[u][3] = p3 * (dp[3] - dd.y);
            }
            uint32_t ah[4], al[4];
            acc_to_a(pd[0], pd[1], ah, al);
            mma3_cols(tv[0], cv[0], ah, al, gt, 0, kk, g, t);
            mma3_cols(tv[1], cv[1], ah, al, gt, 1, kk, g, t);
            acc_to_a(ds[0], ds[1], ah, al);
            mma3_cols(tk[0], ck[0], ah, al, qt, 0, kk, g, t);
            mma3_cols(tk[1], ck[1], ah, al, qt, 1, kk, g, t);
        }
        add_tile(gk[0], tk[0], ck[0]);
        add_tile(gk[1], tk[1], ck[1]);
        add_tile(gv[0], tv[0], cv[0]);
        add_tile(gv[1], tv[1], cv[1]);
    }
    if (!active) return;
#pragma unroll
    for (int dt = 0; dt < 2; ++dt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int d = dt * 8 + 2 * t + j;
            if (d < HD) {
                if (validA) {
                    atomicAdd(dk + ((long)b * nk + kA) * E + hoff + d, gk[dt][j] * inv);
                    atomicAdd(dv + ((long)b * nk + kA) * E + hoff + d, gv[dt][j] * inv);
                }
                if (validB) {
                    atomicAdd(dk + ((long)b * nk + kB) * E + hoff + d, gk[dt][2 + j] * inv);
                    atomicAdd(dv + ((long)b * nk + kB) * E + hoff + d, gv[dt][2 + j] * inv);
                }
            }
        }
}

// ================================================================================ a handful of query rows (nq <= 4)
// Act3D's query stack attends with ONE query token per sample (act3d.py:467-480): a 16-row MMA tile per (sample, head)
// would leave 64 CTAs walking 65 key tiles one after the other.  Here one CTA per (sample, head) spreads the KEYS over
// 256 threads (fp32 CUDA cores: 30 FMAs per key) and reduces across the block; forward and the whole backward are one
// kernel each, dk / dv rows are owned by one thread so no atomics are needed.
constexpr int kFewThreads = 256;
constexpr int kFewMax = 4;

__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float m = red[0];
#pragma unroll
    for (int w = 1; w < kFewThreads / 32; ++w) m = fmaxf(m, red[w]);
    return m;
}
// sums of 16 floats per thread across the block, fixed order; result in every thread
__device__ __forceinline__ void block_sum16(float (&v)[16], float (*red)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int j = 0; j < 16; ++j) red[threadIdx.x >> 5][j] = v[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        float s = red[0][j];
#pragma unroll
        for (int w = 1; w < kFewThreads / 32; ++w) s += red[w][j];
        v[j] = s;
    }
}
__device__ __forceinline__ float dot15g(const float (&a)[HD], const float* __restrict__ row) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s = fmaf(a[d], __ldg(row + d), s);
    return s;
}

template <bool kDrop>
__global__ void __launch_bounds__(kFewThreads) attn_fwd_few_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const unsigned char* __restrict__ mask, int H, int nq, int nk, int E, float* __restrict__ o,
    float* __restrict__ lse, uint32_t drop_thresh, float keep_scale, uint64_t seed) {
    __shared__ float red1[kFewThreads / 32];
    __shared__ float red16[kFewThreads / 32][16];
    const int bh = blockIdx.x, b = bh / H, h = bh - b * H, hoff = h * HD;
    const float* kb = k + (long)b * nk * E + hoff;
    const float* vb = v + (long)b * nk * E + hoff;
    for (int row = 0; row < nq; ++row) {
        float qr[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) qr[d] = __ldg(q + ((long)b * nq + row) * E + hoff + d);
        float m = -INFINITY;
        for (int key = threadIdx.x; key < nk; key += kFewThreads)
            if (!(mask && mask[(long)b * nk + key])) m = fmaxf(m, dot15g(qr, kb + (long)key * E));
        m = block_max(m, red1);
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        if (m > -INFINITY) {
            const uint64_t drop_row = ((uint64_t)bh * nq + (uint64_t)row) * (uint64_t)nk;
            for (int key = threadIdx.x; key < nk; key += kFewThreads) {
                if (mask && mask[(long)b * nk + key]) continue;
                float p = __expf(dot15g(qr, kb + (long)key * E) - m);
                acc[15] += p;
                if (kDrop) p = drop_bits(seed, drop_row + (uint64_t)key) >= drop_thresh ? p * keep_scale : 0.f;
#pragma unroll
                for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, __ldg(vb + (long)key * E + d), acc[d]);
            }
        }
        block_sum16(acc, red16);
        const float l = acc[15], inv = l > 0.f ? 1.f / l : 0.f;
        if (threadIdx.x < HD) o[((long)b * nq + row) * E + hoff + threadIdx.x] = acc[threadIdx.x] * inv;
        if (threadIdx.x == 0) lse[(long)bh * nq + row] = l > 0.f ? m + logf(l) : 0.f;
    }
}

// grid (B*H, key splits): every CTA takes a contiguous range of keys; dq (zero-filled by the caller) collects the
// splits' partial sums with 15 atomics per CTA and row.
constexpr int kFewSplit = 8;
template <bool kDrop>
__global__ void __launch_bounds__(kFewThreads) attn_bwd_few_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
    const unsigned char* __restrict__ mask, const float* __restrict__ o, const float* __restrict__ dout,
    const float* __restrict__ lse, int H, int nq, int nk, int E, float* __restrict__ dq, float* __restrict__ dk,
    float* __restrict__ dv, float* __restrict__ dsum, uint32_t drop_thresh, float keep_scale, uint64_t seed) {
    __shared__ float red16[kFewThreads / 32][16];
    const int bh = blockIdx.x, b = bh / H, h = bh - b * H, hoff = h * HD;
    const int per = (nk + gridDim.y - 1) / gridDim.y;
    const int k_begin = blockIdx.y * per, k_end = min(nk, k_begin + per);
    const float* kb = k + (long)b * nk * E + hoff;
    const float* vb = v + (long)b * nk * E + hoff;
    for (int row = 0; row < nq; ++row) {
        const long base = ((long)b * nq + row) * E + hoff;
        float qr[HD], gr[HD], D = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            qr[d] = __ldg(q + base + d);
            gr[d] = __ldg(dout + base + d);
            D = fmaf(gr[d], __ldg(o + base + d), D);
        }
        const float ls = lse[(long)bh * nq + row];
        const uint64_t drop_row = ((uint64_t)bh * nq + (uint64_t)row) * (uint64_t)nk;
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        for (int key = k_begin + threadIdx.x; key < k_end; key += kFewThreads) {
            if (mask && mask[(long)b * nk + key]) continue;
            const float* kr = kb + (long)key * E;
            const float* vr = vb + (long)key * E;
            float kv[HD], vv[HD];
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                kv[d] = __ldg(kr + d);
                vv[d] = __ldg(vr + d);
            }
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                s = fmaf(qr[d], kv[d], s);
                dp = fmaf(gr[d], vv[d], dp);
            }
            const float p = __expf(s - ls);
            float pd = p;
            if (kDrop) {
                const bool keep = drop_bits(seed, drop_row + (uint64_t)key) >= drop_thresh;
                pd = keep ? p * keep_scale : 0.f;
                dp = keep ? dp * keep_scale : 0.f;
            }
            const float ds = p * (dp - D);
            float* dkr = dk + ((long)b * nk + key) * E + hoff;      // (sample, head, key) belongs to this thread alone
            float* dvr = dv + ((long)b * nk + key) * E + hoff;
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                acc[d] = fmaf(ds, kv[d], acc[d]);
                if (row == 0) {                                        // zero-filled by contract: first row stores
                    dkr[d] = ds * qr[d];
                    dvr[d] = pd * gr[d];
                } else {
                    dkr[d] += ds * qr[d];
                    dvr[d] += pd * gr[d];
                }
            }
        }
        block_sum16(acc, red16);
        if (threadIdx.x < HD) atomicAdd(dq + base + threadIdx.x, acc[threadIdx.x]);
        if (threadIdx.x == 0 && blockIdx.y == 0) dsum[(long)bh * nq + row] = D;
    }
}

}  // namespace

int a3d_launch_attn_fwd_mma(const float* q, const float* k, const float* v, const unsigned char* key_mask, int batch,
                            int heads, int nq, int nk, int embed, float* o, float* lse, int drop, uint32_t th, float sc,
                            uint64_t seed, cudaStream_t st) {
    if (nq <= kFewMax) {
        if (drop)
            attn_fwd_few_kernel<true><<<batch * heads, kFewThreads, 0, st>>>(q, k, v, key_mask, heads, nq, nk, embed, o, lse, th, sc, seed);
        else
            attn_fwd_few_kernel<false><<<batch * heads, kFewThreads, 0, st>>>(q, k, v, key_mask, heads, nq, nk, embed, o, lse, th, sc, seed);
        return check_launch("a3d_attn_fwd");
    }
    dim3 grid((nq + 63) / 64, batch * heads);
    if (drop)
        attn_fwd_mma_kernel<true><<<grid, 128, 0, st>>>(q, k, v, key_mask, heads, nq, nk, embed, o, lse, th, sc, seed);
    else
        attn_fwd_mma_kernel<false><<<grid, 128, 0, st>>>(q, k, v, key_mask, heads, nq, nk, embed, o, lse, th, sc, seed);
    return check_launch("a3d_attn_fwd");
}

// dsum: batch*heads*nq floats (D per row) followed by batch*heads floats (max |dO| per (sample, head)), ZERO-FILLED
int a3d_launch_attn_bwd_mma(const float* q, const float* k, const float* v, const unsigned char* key_mask,
                            const float* o, const float* dout, const float* lse, int batch, int heads, int nq, int nk,
                            int embed, float* dq, float* dk, float* dv, float* dsum, int drop, uint32_t th, float sc,
                            uint64_t seed, cudaStream_t st) {
    if (nq <= kFewMax) {
        if (drop)
            attn_bwd_few_kernel<true><<<dim3(batch * heads, kFewSplit), kFewThreads, 0, st>>>(q, k, v, key_mask, o, dout, lse, heads, nq, nk, embed, dq, dk, dv, dsum, th, sc, seed);
        else
            attn_bwd_few_kernel<false><<<dim3(batch * heads, kFewSplit), kFewThreads, 0, st>>>(q, k, v, key_mask, o, dout, lse, heads, nq, nk, embed, dq, dk, dv, dsum, th, sc, seed);
        return check_launch("a3d_attn_bwd");
    }
    dim3 g1((nq + 63) / 64, batch * heads);
    const int chunks = (nq + kQChunk - 1) / kQChunk;
    if (chunks > 65535) {
        set_error("a3d_attn_bwd: nq=%d too large", nq);
        return A3D_EINVAL;
    }
    dim3 g2((nk + 63) / 64, batch * heads, chunks);
    float* gmax = dsum + (long)batch * heads * nq;
    if (drop) {
        attn_bwd_dq_mma_kernel<true><<<g1, 128, 0, st>>>(q, k, v, key_mask, o, dout, lse, heads, nq, nk, embed, dq, dsum, gmax, th, sc, seed);
        attn_bwd_dkv_mma_kernel<true><<<g2, 128, 0, st>>>(q, k, v, key_mask, dout, lse, dsum, gmax, heads, nq, nk, embed, dk, dv, th, sc, seed);
    } else {
        attn_bwd_dq_mma_kernel<false><<<g1, 128, 0, st>>>(q, k, v, key_mask, o, dout, lse, heads, nq, nk, embed, dq, dsum, gmax, th, sc, seed);
        attn_bwd_dkv_mma_kernel<false><<<g2, 128, 0, st>>>(q, k, v, key_mask, dout, lse, dsum, gmax, heads, nq, nk, embed, dk, dv, th, sc, seed);
    }
    return check_launch("a3d_attn_bwd");
}

}  // namespace a3d
