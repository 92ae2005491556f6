// Memory-bound glue of the image trunk (the dense convolutions themselves stay on cuDNN, SURVEY.md
// section 8f).  Profiling the Act3D step showed ~1.4 ms of a 15 ms step in ATen elementwise kernels around
// the convolutions: input normalisation (3 launches), NHWC max-pool at 1.2 TB/s, and the FPN top-down
// path as separate bias-add / nearest-upsample / add launches over the 128 x 128 x E map.  Each of
// them is one pass over HBM here, arithmetic bit-identical to the ATen ops they replace.
//
//   a3d_trunk_normalize    (x - mean[c]) / std[c], NCHW fp32 in -> NHWC out   (transforms.Normalize,
//                          act3d.py:62 / diffusion_head.py:40 + the channels-last conversion)
//   a3d_trunk_maxpool      3 x 3, stride 2, padding 1, NHWC                     (resnet.py:44 maxpool)
//   a3d_trunk_fpn_topdown  out = (lat + bias) + nearest_upsample(top), NHWC     (torchvision
//                          FeaturePyramidNetwork.forward top-down merge)
#include "a3d_common.cuh"

namespace a3d {

struct Norm3 {
    float mean[3], stdv[3];
};

// one thread per pixel: three coalesced plane reads, one 12-byte row write -- or, with PAD4, one 16-byte write of
// (r, g, b, 0): a 4-channel NHWC input lets cuDNN run the 7x7 stem on its vectorised tensor-core kernels instead of
// the indexed 3-channel fallback (the zero channel meets zero weights: same sums)
template <bool PAD4>
__global__ void __launch_bounds__(256) trunk_normalize_kernel(const float* __restrict__ in, Norm3 nm, int hw, long total,
                                                              float* __restrict__ out) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long n = i / hw;
        const int p = (int)(i - n * hw);
        const float* src = in + n * 3 * hw + p;
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __fdiv_rn(__fsub_rn(__ldg(src + (long)c * hw), nm.mean[c]), nm.stdv[c]);
        if (PAD4) {
            reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], 0.f);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) out[i * 3 + c] = v[c];
        }
    }
}

// thread = (output pixel, 4 channels); the nine taps of neighbouring outputs overlap and are served by L1/L2
__global__ void __launch_bounds__(256) trunk_maxpool_kernel(const float4* __restrict__ in, int h, int w, int c4, int ho, int wo,
                                                            long total, float4* __restrict__ out) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % c4);
        long t = i / c4;
        const int x = (int)(t % wo);
        t /= wo;
        const int y = (int)(t % ho);
        const long n = t / ho;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int yy = 2 * y - 1 + dy;
            if (yy < 0 || yy >= h) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int xx = 2 * x - 1 + dx;
                if (xx < 0 || xx >= w) continue;
                const float4 v = __ldg(in + ((n * h + yy) * w + xx) * c4 + c);
                m.x = fmaxf(m.x, v.x);
                m.y = fmaxf(m.y, v.y);
                m.z = fmaxf(m.z, v.z);
                m.w = fmaxf(m.w, v.w);
            }
        }
        out[i] = m;
    }
}

// thread = (pixel, 4 channels).  Source index of the nearest upsample as ATen computes it:
// min(floor(dst * (in / out)), in - 1) with the scale in fp32.
__global__ void __launch_bounds__(256) trunk_fpn_topdown_kernel(const float4* __restrict__ lat, const float4* __restrict__ bias,
                                                                const float4* __restrict__ top, int h, int w, int c4, int th,
                                                                int tw, float sy, float sx, long total,
                                                                float4* __restrict__ out) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % c4);
        long t = i / c4;
        const int x = (int)(t % w);
        t /= w;
        const int y = (int)(t % h);
        const long n = t / h;
        const int ty = min((int)floorf(y * sy), th - 1), tx = min((int)floorf(x * sx), tw - 1);
        const float4 a = __ldg(lat + i);
        const float4 b = bias ? __ldg(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 u = __ldg(top + ((n * th + ty) * tw + tx) * c4 + c);
        float4 r;
        r.x = __fadd_rn(__fadd_rn(a.x, b.x), u.x);
        r.y = __fadd_rn(__fadd_rn(a.y, b.y), u.y);
        r.z = __fadd_rn(__fadd_rn(a.z, b.z), u.z);
        r.w = __fadd_rn(__fadd_rn(a.w, b.w), u.w);
        out[i] = r;
    }
}

static inline int grid_for(long total) {
    const long want = (total + 255) / 256;
    const long cap = 148L * 16;          // 16 resident 256-thread CTAs' worth per SM, grid-stride beyond
    return (int)(want < cap ? want : cap);
}

}  // namespace a3d

using namespace a3d;

extern "C" int a3d_trunk_normalize(const float* rgb, const float* mean_host, const float* std_host, int images, int hw,
                                   float* out, int out_channels, void* stream) {
    A3D_REQUIRE(rgb && mean_host && std_host && out && images > 0 && hw > 0, "a3d_trunk_normalize: bad arguments");
    A3D_REQUIRE(out_channels == 3 || out_channels == 4, "a3d_trunk_normalize: out_channels must be 3 or 4 (got %d)", out_channels);
    A3D_REQUIRE(out_channels == 3 || ((uintptr_t)out & 15) == 0, "a3d_trunk_normalize: 16-byte aligned output required");
    Norm3 nm;
    for (int c = 0; c < 3; ++c) {
        nm.mean[c] = mean_host[c];
        nm.stdv[c] = std_host[c];
    }
    const long total = (long)images * hw;
    if (out_channels == 4)
        trunk_normalize_kernel<true><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(rgb, nm, hw, total, out);
    else
        trunk_normalize_kernel<false><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(rgb, nm, hw, total, out);
    return check_launch("a3d_trunk_normalize");
}

extern "C" int a3d_trunk_maxpool(const float* in, int images, int h, int w, int channels, float* out, void* stream) {
    A3D_REQUIRE(in && out && images > 0 && h > 0 && w > 0, "a3d_trunk_maxpool: bad arguments");
    A3D_REQUIRE(channels % 4 == 0 && channels > 0, "a3d_trunk_maxpool: channels must be a multiple of 4 (got %d)", channels);
    A3D_REQUIRE((((uintptr_t)in | (uintptr_t)out) & 15) == 0, "a3d_trunk_maxpool: 16-byte aligned buffers required");
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;      // floor((h + 2 - 3) / 2) + 1
    const long total = (long)images * ho * wo * (channels / 4);
    trunk_maxpool_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(in), h, w,
                                                                            channels / 4, ho, wo, total,
                                                                            reinterpret_cast<float4*>(out));
    return check_launch("a3d_trunk_maxpool");
}

extern "C" int a3d_trunk_fpn_topdown(const float* lat, const float* bias, const float* top, int images, int h, int w,
                                     int top_h, int top_w, int channels, float* out, void* stream) {
    A3D_REQUIRE(lat && top && out && images > 0 && h > 0 && w > 0 && top_h > 0 && top_w > 0,
                "a3d_trunk_fpn_topdown: bad arguments");
    A3D_REQUIRE(channels % 4 == 0 && channels > 0, "a3d_trunk_fpn_topdown: channels must be a multiple of 4 (got %d)", channels);
    A3D_REQUIRE((((uintptr_t)lat | (uintptr_t)top | (uintptr_t)out | (uintptr_t)bias) & 15) == 0,
                "a3d_trunk_fpn_topdown: 16-byte aligned buffers required");
    const long total = (long)images * h * w * (channels / 4);
    const float sy = (float)top_h / (float)h, sx = (float)top_w / (float)w;
    trunk_fpn_topdown_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(lat), reinterpret_cast<const float4*>(bias), reinterpret_cast<const float4*>(top), h, w,
        channels / 4, top_h, top_w, sy, sx, total, reinterpret_cast<float4*>(out));
    return check_launch("a3d_trunk_fpn_topdown");
}
