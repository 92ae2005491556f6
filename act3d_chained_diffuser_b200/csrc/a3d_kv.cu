// Context K/V cache builder: k/v in-projection + 3-D rotary on K + head split + fp16 tile images.
//
// One CTA = one 64-key tile of one sample.  The 64 context tokens are staged once (K-major) in
// shared memory together with their cos/sin table, then for every attention layer ("set") that
// shares this context the CTA runs the [64 x E] x [E x 2E] projection as a register-tiled fp32
// GEMM, rotates K, rounds to fp16 and assembles the tile image in shared memory in exactly the
// byte layout the attention kernels ldmatrix from; the image leaves with 16-byte coalesced
// stores.  HBM-bound: reads 64*(E+3)*4 B, writes nsets*2*H*2 KiB per CTA.
#include "a3d_linear.cuh"

namespace a3d {

template <int E, int H>
struct KvCfg {
    static constexpr int EP = 16 * H;          // padded embed (64 / 128)
    static constexpr int NW = 2 * EP;          // packed weight width (K | V)
    static constexpr int RP = 66;              // row pitch of the K-major token tile
    static constexpr int PASSES = NW / 128;    // 128 output columns per GEMM pass
    static constexpr int IMG_HALF = H * 2048;  // bytes of the K (or V) image of one tile
    static constexpr size_t SMEM = (size_t)E * RP * 4 + 2 * 64 * (E / 2) * 4 + 2 * IMG_HALF;
};

template <int E, int H>
__global__ void __launch_bounds__(256) ctx_kv_kernel(const float* __restrict__ tok, const float* __restrict__ pos,
                                                     int tok_rows, int nk, const float* __restrict__ wkv,
                                                     const float* __restrict__ bkv, unsigned rope_mask, int nsets,
                                                     unsigned char* __restrict__ kv, int batch, int ntiles) {
    using C = KvCfg<E, H>;
    extern __shared__ __align__(16) unsigned char smem[];
    float* xt = reinterpret_cast<float*>(smem);                       // [E][RP]
    float* cs = xt + E * C::RP;                                       // [64][E/2] cos
    float* sn = cs + 64 * (E / 2);                                    // [64][E/2] sin
    unsigned char* img = reinterpret_cast<unsigned char*>(sn + 64 * (E / 2));   // K image | V image

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, b = blockIdx.y;
    const int r0 = tile * 64;

    // ---- stage tokens (K-major) and the rotary table
    for (int i = tid; i < 64 * E; i += 256) {
        const int r = i / E, c = i - r * E;
        float v = 0.f;
        if (r0 + r < nk) v = __ldg(tok + ((long)b * tok_rows + r0 + r) * E + c);
        xt[c * C::RP + r] = v;
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

    const int rg = (warp & 3) * 8 + (lane & 7);
    const int cg = (warp >> 2) * 4 + (lane >> 3);
    const int row_a = 2 * rg, row_b = 2 * rg + 1;

    for (int s = 0; s < nsets; ++s) {
        const bool rope = (rope_mask >> s) & 1u;
        const float* w_s = wkv + (size_t)s * E * C::NW;
        const float* b_s = bkv + (size_t)s * C::NW;
#pragma unroll
        for (int pass = 0; pass < C::PASSES; ++pass) {
            float acc[2][16];
            gemm_2x16<E, C::RP, C::NW>(xt, w_s + pass * 128, rg, cg, acc);
            const int col0 = pass * 128 + 16 * cg;            // column in the packed [K | V] output
            const bool is_v = col0 >= C::EP;
            const int dim0 = col0 - (is_v ? C::EP : 0);       // first embed dim of this thread
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float bias = __ldg(b_s + col0 + c);
                acc[0][c] += bias;
                acc[1][c] += bias;
            }
            if (!is_v && rope) {
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int pi = (dim0 >> 1) + p;
                    if (2 * pi < E) {
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            const int row = (r == 0) ? row_a : row_b;
                            const float c = cs[row * (E / 2) + pi], sv = sn[row * (E / 2) + pi];
                            const float ev = acc[r][2 * p], od = acc[r][2 * p + 1];
                            acc[r][2 * p] = ev * c - od * sv;
                            acc[r][2 * p + 1] = od * c + ev * sv;
                        }
                    }
                }
            }
            unsigned char* base = img + (is_v ? C::IMG_HALF : 0);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = (r == 0) ? row_a : row_b;
                const bool valid = (r0 + row) < nk;
                const int swz = (row >> 2) & 1;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int dim = dim0 + c;
                    int h, d;
                    float val;
                    if (dim < E) {
                        h = dim / 15;
                        d = dim - 15 * h;
                        val = valid ? acc[r][c] : 0.f;
                    } else {            // padded embed dims E..EP-1 own the pad slot (d = 15) of head dim-E
                        h = dim - E;
                        d = 15;
                        val = valid ? 1.f : 0.f;      // V: softmax denominator; K: carries the shift of a3d_xattn4.cu
                    }
                    const int off = h * 2048 + row * 32 + (((d >> 3) ^ swz) << 4) + (d & 7) * 2;
                    *reinterpret_cast<__half*>(base + off) = __float2half_rn(val);
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
                          const float* wkv, const float* bkv, const int* rope_host, int nsets, void* kv,
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
        ctx_kv_kernel<60, 4><<<grid, 256, C::SMEM, (cudaStream_t)stream>>>(tok, pos, tok_rows, nk, wkv, bkv, mask, nsets,
                                                                         (unsigned char*)kv, batch, ntiles);
    } else if (embed == 120 && heads == 8) {
        using C = KvCfg<120, 8>;
        cudaFuncSetAttribute(ctx_kv_kernel<120, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        ctx_kv_kernel<120, 8><<<grid, 256, C::SMEM, (cudaStream_t)stream>>>(tok, pos, tok_rows, nk, wkv, bkv, mask, nsets,
                                                                          (unsigned char*)kv, batch, ntiles);
    } else {
        A3D_REQUIRE(false, "a3d_ctx_kv: (embed, heads) = (%d, %d) not supported; use (60,4) or (120,8)", embed, heads);
    }
    return check_launch("a3d_ctx_kv");
}
