// Context K/V cache builder: k/v in-projection + 3-D rotary on K + head split + fp16 tile images.
// HBM-bound by design: reads 64*(E+3)*4 B, writes nsets*2*H*2 KiB per CTA (64-key tile).
#include "a3d_mma_gemm.cuh"

namespace a3d {

template <int E, int H>
struct KvCfg {
    static constexpr int EP = 16 * H;          // padded embed (64 / 128) = padded K of the projection
    static constexpr int NW = 2 * EP;          // packed output width (K | V)
    static constexpr int KSTEPS = EP / 16;     // k16 steps
    static constexpr int NTILES = NW / 8;      // n8 tiles of one set (K tiles first, then V tiles)
    static constexpr int PITCH = EP + 8;       // halfs per row of an activation plane (conflict-free ldmatrix)
    static constexpr int IMG_HALF = H * 2048;  // bytes of the K (or V) image of one tile
    static constexpr int SET_UINT4 = KSTEPS * NTILES * 32;   // fragment-ordered weight of one set
    static constexpr size_t SMEM = (size_t)2 * 64 * PITCH * 2 + 2 * 64 * (E / 2) * 4 + 2 * IMG_HALF;
};

// One CTA = one 64-key tile of one sample, 8 warps = 4 row tiles x {K half, V half}.  The 64 tokens are staged once as
// fp16 (hi, lo) planes together with their cos/sin table; for every attention layer ("set") sharing this context the
// [64 x E] x [E x 2E] projection runs on the tensor cores as the error-compensated split-fp16 GEMM of
// a3d_mma_gemm.cuh (fp32-class accuracy: the cache must agree with an fp32 projection to 1 fp16 ulp), K is rotated in
// the accumulator fragments (a rotary pair = the two columns a thread holds), rounded to fp16 and written into the
// tile image in exactly the byte layout the attention kernels consume; the image leaves with 16-byte coalesced stores.
template <int E, int H>
__global__ void __launch_bounds__(256) ctx_kv_kernel(const float* __restrict__ tok, const float* __restrict__ pos,
                                                     int tok_rows, int nk, const uint4* __restrict__ wkv,
                                                     const float* __restrict__ bkv, unsigned rope_mask, int nsets,
                                                     unsigned char* __restrict__ kv, int batch, int ntiles) {
    using C = KvCfg<E, H>;
    extern __shared__ __align__(16) unsigned char smem[];
    __half* ah = reinterpret_cast<__half*>(smem);                      // [64][PITCH] hi plane
    __half* al = ah + 64 * C::PITCH;                                   // [64][PITCH] lo plane
    float* cs = reinterpret_cast<float*>(al + 64 * C::PITCH);          // [64][E/2] cos
    float* sn = cs + 64 * (E / 2);                                     // [64][E/2] sin
    unsigned char* img = reinterpret_cast<unsigned char*>(sn + 64 * (E / 2));   // K image | V image

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, b = blockIdx.y;
    const int r0 = tile * 64;

    // ---- stage tokens as (hi, lo) fp16 planes (zero beyond nk and in the padded columns) and the rotary table
    for (int i = tid; i < 64 * (C::EP / 2); i += 256) {
        const int r = i / (C::EP / 2), c = 2 * (i - r * (C::EP / 2));
        float v0 = 0.f, v1 = 0.f;
        if (r0 + r < nk && c < E) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(tok + ((long)b * tok_rows + r0 + r) * E + c));
            v0 = t.x;
            v1 = t.y;
        }
        uint32_t hi, lo;
        split_h2(v0, v1, hi, lo);
        *reinterpret_cast<uint32_t*>(ah + r * C::PITCH + c) = hi;
        *reinterpret_cast<uint32_t*>(al + r * C::PITCH + c) = lo;
    }
    for (int i = tid; i < 64 * (E / 2); i += 256) {
        const int r = i / (E / 2), p = i - r * (E / 2);
        const int axis = p / (E / 6), j = p - axis * (E / 6);
        float c = 1.f, s = 0.f;
        if (r0 + r < nk) {
            const float ang = __fmul_rn(__ldg(pos + ((long)b * tok_rows + r0 + r) * 3 + axis), rope_freq<E>(j));
            sincosf(ang, &s, &c);
        }
        cs[i] = c;
        sn[i] = s;
    }
    __syncthreads();

    const int m0 = 16 * (warp & 3);
    const bool is_v = (warp >> 2) != 0;
    const int g = lane >> 2, q4 = lane & 3;
    constexpr int NT = 8;                                   // n tiles per GEMM call (64 output columns)

    for (int s = 0; s < nsets; ++s) {
        const bool rope = (rope_mask >> s) & 1u;
        const uint4* w_s = wkv + (size_t)s * C::SET_UINT4;
        const float* b_s = bkv + (size_t)s * C::NW + (is_v ? C::EP : 0);
        unsigned char* base = img + (is_v ? C::IMG_HALF : 0);
#pragma unroll 1
        for (int part = 0; part < C::EP / 64; ++part) {     // 64 columns of this warp's half per pass
            float acc[NT][4];
            const int nt0 = (is_v ? C::EP / 8 : 0) + part * NT;
            mma_gemm_split<C::KSTEPS, NT, C::PITCH>(ah, al, m0, w_s, C::NTILES, nt0, lane, acc);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int dim = part * 64 + 8 * n + 2 * q4;                 // even; this thread holds dims (dim, dim+1)
                const float b0 = __ldg(b_s + dim), b1 = __ldg(b_s + dim + 1);
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int row = m0 + g + 8 * rr;
                    const bool valid = (r0 + row) < nk;
                    float v0 = acc[n][2 * rr] + b0, v1 = acc[n][2 * rr + 1] + b1;
                    if (!is_v && rope && dim < E) {                          // rotary pair (2i, 2i+1), i = dim / 2
                        const float c = cs[row * (E / 2) + (dim >> 1)], sv = sn[row * (E / 2) + (dim >> 1)];
                        const float ev = v0, od = v1;
                        v0 = ev * c - od * sv;
                        v1 = od * c + ev * sv;
                    }
                    const int swz = (row >> 2) & 1;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int dd = dim + e;
                        int h, d;
                        float val;
                        if (dd < E) {
                            h = dd / 15;
                            d = dd - 15 * h;
                            val = valid ? (e ? v1 : v0) : 0.f;
                        } else {            // padded embed dims E..EP-1 own the pad slot (d = 15) of head dd-E
                            h = dd - E;
                            d = 15;
                            val = valid ? 1.f : 0.f;      // V: softmax denominator; K: carries the shift of a3d_xattn4.cu
                        }
                        const int off = h * 2048 + row * 32 + (((d >> 3) ^ swz) << 4) + (d & 7) * 2;
                        *reinterpret_cast<__half*>(base + off) = __float2half_rn(val);
                    }
                }
            }
        }
        __syncthreads();
        // ---- image out: [set][b][tile][K|V][H][64][16] fp16
        uint4* dst = reinterpret_cast<uint4*>(kv + (((size_t)s * batch + b) * ntiles + tile) * (size_t)(2 * C::IMG_HALF));
        const uint4* src = reinterpret_cast<const uint4*>(img);
        for (int i = tid; i < 2 * C::IMG_HALF / 16; i += 256) dst[i] = src[i];
        __syncthreads();
    }
}

}  // namespace a3d

using namespace a3d;

extern "C" size_t a3d_kv_bytes(int nsets, int batch, int nk, int heads) {
    const size_t ntiles = (size_t)(nk + kTileKeys - 1) / kTileKeys;
    return (size_t)nsets * batch * ntiles * 2 * heads * 2048;
}

extern "C" int a3d_ctx_kv(const float* tok, const float* pos, int batch, int tok_rows, int nk, int embed, int heads,
                          const void* wkv, const float* bkv, const int* rope_host, int nsets, void* kv,
                          void* stream) {
    A3D_REQUIRE(tok && pos && wkv && bkv && rope_host && kv, "a3d_ctx_kv: null pointer");
    A3D_REQUIRE(batch > 0 && nk > 0 && nk <= tok_rows, "a3d_ctx_kv: need 0 < nk <= tok_rows (nk=%d rows=%d)", nk, tok_rows);
    A3D_REQUIRE(nsets > 0 && nsets <= 32, "a3d_ctx_kv: nsets=%d not in [1,32]", nsets);
    unsigned mask = 0;
    for (int s = 0; s < nsets; ++s) mask |= (rope_host[s] ? 1u : 0u) << s;
    const int ntiles = (nk + kTileKeys - 1) / kTileKeys;
    dim3 grid(ntiles, batch);
    if (embed == 60 && heads == 4) {
        using C = KvCfg<60, 4>;
        cudaFuncSetAttribute(ctx_kv_kernel<60, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        ctx_kv_kernel<60, 4><<<grid, 256, C::SMEM, (cudaStream_t)stream>>>(tok, pos, tok_rows, nk, (const uint4*)wkv, bkv, mask, nsets,
                                                                         (unsigned char*)kv, batch, ntiles);
    } else if (embed == 120 && heads == 8) {
        using C = KvCfg<120, 8>;
        cudaFuncSetAttribute(ctx_kv_kernel<120, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        ctx_kv_kernel<120, 8><<<grid, 256, C::SMEM, (cudaStream_t)stream>>>(tok, pos, tok_rows, nk, (const uint4*)wkv, bkv, mask, nsets,
                                                                          (unsigned char*)kv, batch, ntiles);
    } else {
        A3D_REQUIRE(false, "a3d_ctx_kv: (embed, heads) = (%d, %d) not supported; use (60,4) or (120,8)", embed, heads);
    }
    return check_launch("a3d_ctx_kv");
}
