// Error-compensated tensor-core GEMM for the small per-token linear layers (fp32-class accuracy).
//
// Every fp32 operand x is carried as two fp16 planes:  hi = fp16(x),  lo = fp16((x - hi) * 2^11).
// With fp32 accumulation
//        A B  ~=  Ah Bh  +  2^-11 (Ah Bl + Al Bh)                    (dropped term Al Bl ~ 2^-22)
// i.e. three m16n8k16 HMMAs per 16-wide k step, two accumulators (main, correction).  Compared with
// 3xTF32 (six k8 MMAs) this halves the tensor-core instructions; compared with single-pass
// TF32/fp16 it removes the 2^-11 operand rounding that the 100-step denoising loop cannot afford
// (measured: all linears at 10-bit mantissa -> 9.3e-4 m final position error vs the 1e-3 budget).
//
// Layouts
//   A (activations): shared memory, row-major fp16 planes  [rows][PITCH] (hi) and (lo), read with
//       ldmatrix.x4 (16 x 16 per k step); PITCH*2 bytes must be an odd multiple of 16 B mod 128.
//   B (weights): global memory, pre-split on the host and stored in FRAGMENT ORDER
//       [kstep][ntile][lane] -> uint4 {b0_hi, b1_hi, b0_lo, b1_lo}   (packing.py: pack_mma_weight)
//       so a warp fetches the operands of one (k step, n tile) with one fully coalesced 16-byte load
//       per lane; weights are tiny, shared by all CTAs and served by L2.
//   C: fp32 accumulator fragments (rows g, g+8; cols 2q, 2q+1 of each 8-wide n tile).
#pragma once
#include "a3d_common.cuh"

namespace a3d {

constexpr float kLoScale = 2048.0f;          // 2^11
constexpr float kLoScaleInv = 1.0f / 2048.0f;

__device__ __forceinline__ void split_h(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * kLoScale);
}
// two consecutive elements -> packed half2 words for the hi and lo planes (packed conversions: cvt.rn.f16x2.f32 is one
// F2FP on the ALU pipe, the scalar cvt.rn.f16.f32 a quarter-rate F2F; same round-to-nearest results)
__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(x0, x1);
    const float2 f = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn((x0 - f.x) * kLoScale, (x1 - f.y) * kLoScale);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

// acc[nt][4] (+)= A[m0 .. m0+16, 0 .. 16*KSTEPS) * W^T[:, n tiles nt0 .. nt0+NT)
//   a_hi / a_lo : shared planes, row pitch PITCH halfs;  wfrag : fragment-ordered weight, NTILES n tiles per k step
template <int KSTEPS, int NT, int PITCH>
__device__ __forceinline__ void mma_gemm_split(const __half* __restrict__ a_hi, const __half* __restrict__ a_lo, int m0,
                                               const uint4* __restrict__ wfrag, int ntiles_total, int nt0, int lane,
                                               float (&acc)[NT][4]) {
    float cor[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[n][e] = 0.f;
            cor[n][e] = 0.f;
        }
    const int arow = m0 + (lane & 7) + 8 * ((lane >> 3) & 1);
    const int acol = 8 * (lane >> 4);
    const uint32_t ah_base = smem_u32(a_hi + arow * PITCH + acol);
    const uint32_t al_base = smem_u32(a_lo + arow * PITCH + acol);
    const uint4* wp = wfrag + (size_t)nt0 * 32 + lane;

    uint4 bq[2][NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) bq[0][n] = __ldg(wp + (size_t)n * 32);
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
        const int cur = ks & 1;
        if (ks + 1 < KSTEPS) {
#pragma unroll
            for (int n = 0; n < NT; ++n) bq[cur ^ 1][n] = __ldg(wp + ((size_t)(ks + 1) * ntiles_total + n) * 32);
        }
        uint32_t ah[4], al[4];
        ldmatrix_x4(ah, ah_base + ks * 32);
        ldmatrix_x4(al, al_base + ks * 32);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const uint4 b = bq[cur][n];
            mma_16816(acc[n], ah, b.x, b.y);
            mma_16816(cor[n], ah, b.z, b.w);
            mma_16816(cor[n], al, b.x, b.y);
        }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[n][e] = fmaf(cor[n][e], kLoScaleInv, acc[n][e]);
}


// ------------------------------------------------------------------------------------------------
// CTA-wide variant: 256 threads compute a 64 x 128 output tile; warp = (m tile 0..3, n half 0..1),
// eight n tiles per warp in ONE pass over k.  The weight slab of every k step (16 n tiles x 512 B
// = 8 KiB, contiguous in the fragment layout) is streamed global -> shared with cp.async into a
// 4-stage ring, three k steps ahead, so that the L2 latency is covered by tensor-core work instead
// of being paid per k step; the head of the NEXT GEMM can be prefetched while the current epilogue
// runs (`ring_prefetch_head`).  One __syncthreads per k step.
constexpr int kRingStages = 4;
constexpr int kSlabBytes = 16 * 32 * 16;   // one k step of a 128-wide window

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct WeightRing {
    unsigned char* buf;        // kRingStages * kSlabBytes, 16-byte aligned
    const uint4* head_of;      // weight whose first (kRingStages-1) slabs are already in flight / resident, or null
};

// slab ks of a window: wfrag + (ks * ntiles_total + nt0) * 32 uint4, 512 uint4 long
__device__ __forceinline__ void ring_issue(const WeightRing& r, const uint4* wfrag, int ntiles_total, int nt0, int ks) {
    const uint4* src = wfrag + ((size_t)ks * ntiles_total + nt0) * 32;
    unsigned char* dst = r.buf + (ks % kRingStages) * kSlabBytes;
    for (int i = threadIdx.x; i < kSlabBytes / 16; i += blockDim.x) cp_async16(dst + i * 16, src + i);
}
// start fetching the first slabs of a GEMM early (call after a __syncthreads that follows the previous GEMM's k loop)
template <int KSTEPS>
__device__ __forceinline__ void ring_prefetch_head(WeightRing& r, const uint4* wfrag, int ntiles_total, int nt0) {
#pragma unroll
    for (int ks = 0; ks < kRingStages - 1; ++ks) {
        if (ks < KSTEPS) ring_issue(r, wfrag, ntiles_total, nt0, ks);
        cp_async_commit();
    }
    r.head_of = wfrag + (size_t)nt0 * 32;
}

// acc[NT][4] = A[m0 .. m0+16, 0 .. 16*KSTEPS) * W^T[:, window n tiles NT*nh .. NT*nh+NT)
// (NT = 8 with 8 warps / 256 threads, NT = 4 with 16 warps / 512 threads: warp = (m tile, n part nh))
template <int KSTEPS, int PITCH, int NT>
__device__ __forceinline__ void mma_gemm_split_cta(const __half* __restrict__ a_hi, const __half* __restrict__ a_lo, int m0,
                                                   int nh, WeightRing& r, const uint4* __restrict__ wfrag, int ntiles_total,
                                                   int nt0, int lane, float (&acc)[NT][4]) {
    float cor[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[n][e] = 0.f;
            cor[n][e] = 0.f;
        }
    if (r.head_of != wfrag + (size_t)nt0 * 32) {
        __syncthreads();   // everybody is done with the ring contents of the previous GEMM
        ring_prefetch_head<KSTEPS>(r, wfrag, ntiles_total, nt0);
    }
    r.head_of = nullptr;
    const int arow = m0 + (lane & 7) + 8 * ((lane >> 3) & 1);
    const int acol = 8 * (lane >> 4);
    const uint32_t ah_base = smem_u32(a_hi + arow * PITCH + acol);
    const uint32_t al_base = smem_u32(a_lo + arow * PITCH + acol);
#pragma unroll 1
    for (int ks = 0; ks < KSTEPS; ++ks) {
        cp_async_wait<kRingStages - 2>();
        __syncthreads();                                   // slab ks visible to all; slab ks-1 fully consumed
        if (ks + kRingStages - 1 < KSTEPS) ring_issue(r, wfrag, ntiles_total, nt0, ks + kRingStages - 1);
        cp_async_commit();
        uint32_t ah[4], al[4];
        ldmatrix_x4(ah, ah_base + ks * 32);
        ldmatrix_x4(al, al_base + ks * 32);
        const uint4* slab = reinterpret_cast<const uint4*>(r.buf + (ks % kRingStages) * kSlabBytes) + (NT * nh) * 32 + lane;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const uint4 b = slab[n * 32];
            mma_16816(acc[n], ah, b.x, b.y);
            mma_16816(cor[n], ah, b.z, b.w);
            mma_16816(cor[n], al, b.x, b.y);
        }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[n][e] = fmaf(cor[n][e], kLoScaleInv, acc[n][e]);
}

}  // namespace a3d
