"""ctypes binding of libact3d_b200.so (the C ABI declared in include/act3d_b200.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
Device pointers are taken from torch tensors (PyTorch owns all memory); the current torch
CUDA stream is passed to every call.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_long, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libact3d_b200.so")

EXPORTS = {
    # name: (restype, argtypes)
    "a3d_last_error": (c_char_p, []),
    "a3d_abi_version": (c_int, []),
    "a3d_set_option": (c_int, [c_char_p, c_int]),
    "a3d_debug_counter": (c_int, [c_char_p, c_int, ctypes.POINTER(ctypes.c_ulonglong)]),
    "a3d_pcd_pyramid": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "a3d_local_topk": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "a3d_traj_topk": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "a3d_gather_tokens": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "a3d_trunk_normalize": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), c_int, c_int,
                                    c_void_p, c_int, c_void_p]),
    "a3d_trunk_maxpool": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "a3d_trunk_fpn_topdown": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p]),
    "a3d_kv_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "a3d_ctx_kv": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                           ctypes.POINTER(c_int), c_int, c_void_p, c_void_p]),
    "a3d_xattn_layer_floats": (c_size_t, [c_int, c_int]),
    "a3d_xattn_layer_words": (c_size_t, [c_int, c_int]),
    "a3d_xattn_stack": (c_int, [c_void_p, c_long, c_long, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                                c_void_p, c_void_p]),
    "a3d_mask_logits": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "a3d_argmax_pick": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "a3d_sample_ghost": (c_int, [c_void_p, c_float, ctypes.POINTER(c_float), c_int, c_int, c_uint64, c_uint64,
                                 c_void_p, c_void_p]),
    "a3d_sample_ghost_ctr": (c_int, [c_void_p, c_float, ctypes.POINTER(c_float), c_int, c_int, c_uint64, c_uint64,
                                     c_void_p, c_void_p, c_void_p]),
    "a3d_counter_add": (c_int, [c_void_p, c_uint64, c_void_p]),
    "a3d_attn_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                             c_void_p, c_float, c_uint64, c_void_p]),
    "a3d_attn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                             c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_uint64, c_void_p]),
    "a3d_rope_apply": (c_int, [c_void_p, c_void_p, c_long, c_int, c_int, c_void_p, c_void_p]),
    "a3d_gather_tokens_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                      c_void_p]),
    "a3d_linear_wgrad_workspace": (c_size_t, [c_long, c_int, c_int]),
    "a3d_linear_wgrad": (c_int, [c_void_p, c_void_p, c_long, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "a3d_linear_supported": (c_int, [c_int, c_int, c_int]),
    "a3d_linear_workspace": (c_size_t, [c_int, c_int, c_int]),
    "a3d_linear_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "a3d_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_float, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "a3d_layernorm_bwd_workspace": (c_size_t, [c_int]),
    "a3d_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "a3d_soft_ce": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "cd_pack_floats": (c_size_t, [c_int]),
    "cd_ctx_lang": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                            c_int, c_void_p]),
    "cd_step_begin": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                              c_void_p, c_void_p]),
    "cd_cross_part_floats": (c_size_t, [c_int]),
    "cd_cross": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cd_post": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                        c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                        ctypes.POINTER(c_float), c_void_p, c_void_p, c_void_p]),
    "cd_denoise_loop": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
}

_lib = None


class A3DError(RuntimeError):
    pass


def load():
    """Load the shared library (building nothing: run __graft_entry__.build() or build.py first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise A3DError(
            f"{LIB_PATH} is missing. The CUDA extension is the product: build it with "
            "`python -m act3d_chained_diffuser_b200.build` (or __graft_entry__.build()). There is no CPU/eager fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)     # AttributeError => the .so does not match include/act3d_b200.h
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    core = os.environ.get("A3D_XATTN_CORE")          # profiling / A-B switch: 2 = mma.sync core, 3 = tcgen05 core
    if core:
        if lib.a3d_set_option(b"xattn_core", int(core)) != 0:
            raise A3DError(f"A3D_XATTN_CORE={core}: " + lib.a3d_last_error().decode())
    return lib


_launches = 0


def launch_count():
    """Number of kernel-launching C-ABI calls since the last reset (each is exactly one kernel)."""
    return _launches


def reset_launch_count():
    global _launches
    _launches = 0


def _check(code, what):
    global _launches
    _launches += 1
    if code != 0:
        msg = load().a3d_last_error()
        raise A3DError(f"{what} failed ({code}): {msg.decode() if msg else '?'}")


def set_option(name, value):
    global _launches
    _check(load().a3d_set_option(name.encode(), int(value)), "a3d_set_option")
    _launches -= 1                                    # not a kernel launch


def debug_counter(name, reset=False):
    """Diagnostics counter of the current device (synchronises it): see a3d_debug_counter in include/act3d_b200.h."""
    global _launches
    out = ctypes.c_ulonglong(0)
    _check(load().a3d_debug_counter(name.encode(), int(bool(reset)), ctypes.byref(out)), "a3d_debug_counter")
    _launches -= 1                                    # not a kernel launch
    return int(out.value)


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "libact3d_b200 works on CUDA tensors only (no CPU fallback)"
    assert t.is_contiguous(), "expected a contiguous tensor"
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t):
    assert t.dtype == torch.float32, f"expected float32, got {t.dtype}"
    return t


# ------------------------------------------------------------------------------------------------
def pcd_pyramid(pcd_flat, factor):
    """(BN, 3, H, W) -> (BN * h * w, 3) flattened; caller reshapes to (B, ncam*h*w, 3)."""
    bn, c, h, w = pcd_flat.shape
    assert c == 3
    out = torch.empty(bn * (h // factor) * (w // factor), 3, device=pcd_flat.device, dtype=torch.float32)
    _check(load().a3d_pcd_pyramid(_ptr(_f32(pcd_flat)), bn, h, w, factor, _ptr(out), _stream()), "a3d_pcd_pyramid")
    return out


def local_topk(center, pts, k, want_dist=False):
    """center (B,3), pts (B,N,3) -> idx (B,k) int32 ascending by (distance, index)."""
    b, n, _ = pts.shape
    idx = torch.empty(b, k, device=pts.device, dtype=torch.int32)
    dist = torch.empty(b, k, device=pts.device, dtype=torch.float32) if want_dist else None
    _check(load().a3d_local_topk(_ptr(_f32(center)), _ptr(_f32(pts)), b, n, k, _ptr(idx), _ptr(dist), _stream()),
           "a3d_local_topk")
    return (idx, dist) if want_dist else idx


def traj_topk(traj, pts, k, want_dist=False):
    b, n, _ = pts.shape
    idx = torch.empty(b, k, device=pts.device, dtype=torch.int32)
    dist = torch.empty(b, k, device=pts.device, dtype=torch.float32) if want_dist else None
    _check(load().a3d_traj_topk(_ptr(_f32(traj)), traj.shape[1], _ptr(_f32(pts)), b, n, k, _ptr(idx), _ptr(dist),
                                _stream()), "a3d_traj_topk")
    return (idx, dist) if want_dist else idx


def gather_tokens(feat, pcd, idx, batch, ncam, tok, pos, bias=None):
    """feat (B*ncam, E, h, w) -- NCHW-contiguous or channels-last (NHWC storage, read in place) --,
    pcd (B, ncam*h*w, 3), idx (B,K) int32 or None -> rows [0,K) of tok/pos.  bias (E,) or None is added to
    every gathered feature row (the deferred bias of the FPN output convolution)."""
    e, hw = feat.shape[1], feat.shape[2] * feat.shape[3]
    k = idx.shape[1] if idx is not None else ncam * hw
    assert feat.is_cuda and feat.dtype == torch.float32
    if feat.is_contiguous():
        nhwc = 0
    elif feat.is_contiguous(memory_format=torch.channels_last):
        nhwc = 1
    else:
        feat, nhwc = feat.contiguous(), 0
    _check(load().a3d_gather_tokens(feat.data_ptr(), _ptr(_f32(pcd)), _ptr(idx), batch, ncam, e, hw, k,
                                    _ptr(tok), _ptr(pos), tok.shape[1], nhwc, _ptr(_f32(bias)) if bias is not None else None,
                                    _stream()), "a3d_gather_tokens")
    return k


def _nhwc_like(images, channels, h, w, device):
    """fp32 tensor of logical shape (images, channels, h, w) stored channels-last."""
    return torch.empty(images, h, w, channels, device=device, dtype=torch.float32).permute(0, 3, 1, 2)


def trunk_normalize(rgb, mean, std, out_channels=3):
    """(N,3,H,W) NCHW fp32 -> (x - mean) / std as a channels-last tensor (transforms.Normalize + layout change);
    out_channels=4 appends a zero channel (16-byte pixels for the stem convolution)."""
    assert rgb.is_cuda and rgb.dim() == 4 and rgb.shape[1] == 3
    rgb = _f32(rgb)
    n, _, h, w = rgb.shape
    out = _nhwc_like(n, out_channels, h, w, rgb.device)
    m3, s3 = (ctypes.c_float * 3)(*[float(v) for v in mean]), (ctypes.c_float * 3)(*[float(v) for v in std])
    _check(load().a3d_trunk_normalize(_ptr(rgb), m3, s3, n, h * w, out.data_ptr(), int(out_channels), _stream()),
           "a3d_trunk_normalize")
    return out


def _as_nhwc(x):
    return x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous(memory_format=torch.channels_last)


def trunk_maxpool(x):
    """3x3 / stride 2 / padding 1 max-pool of a channels-last (N,C,H,W) map."""
    x = _as_nhwc(x.float())
    n, c, h, w = x.shape
    out = _nhwc_like(n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1, x.device)
    _check(load().a3d_trunk_maxpool(x.data_ptr(), n, h, w, c, out.data_ptr(), _stream()), "a3d_trunk_maxpool")
    return out


def trunk_fpn_topdown(lat, bias, top, out=None):
    """(lat + bias) + nearest-upsampled top, channels-last maps; written into `out` (default: in place on lat)."""
    lat, top = _as_nhwc(lat.float()), _as_nhwc(top.float())
    n, c, h, w = lat.shape
    out = lat if out is None else out
    _check(load().a3d_trunk_fpn_topdown(lat.data_ptr(), _ptr(_f32(bias)) if bias is not None else None, top.data_ptr(),
                                        n, h, w, top.shape[2], top.shape[3], c, out.data_ptr(), _stream()),
           "a3d_trunk_fpn_topdown")
    return out


def kv_bytes(nsets, batch, nk, heads):
    return load().a3d_kv_bytes(nsets, batch, nk, heads)


def ctx_kv(tok, pos, nk, heads, wkv, bkv, rope_flags, out=None):
    """tok (B,rows,E), pos (B,rows,3) -> uint8 buffer holding [nsets][B][ntiles][2][H][64][16] fp16."""
    b, rows, e = tok.shape
    nsets = len(rope_flags)
    nbytes = kv_bytes(nsets, b, nk, heads)
    if out is None or out.numel() < nbytes:
        out = torch.empty(nbytes, device=tok.device, dtype=torch.uint8)
    flags = (c_int * nsets)(*[int(f) for f in rope_flags])
    assert wkv.dtype == torch.int32 and wkv.numel() == nsets * 2 * (16 * heads) ** 2, "wkv: packing.pack_kv_set layout expected"
    _check(load().a3d_ctx_kv(_ptr(_f32(tok)), _ptr(_f32(pos)), b, rows, nk, e, heads, _raw(wkv), _ptr(_f32(bkv)),
                             flags, nsets, _ptr(out), _stream()), "a3d_ctx_kv")
    return out


def xattn_layer_floats(embed, ffn):
    n = load().a3d_xattn_layer_floats(embed, ffn)
    if n == 0:
        raise A3DError(f"a3d_xattn_stack is not built for embed={embed}, ffn={ffn}")
    return n


def xattn_layer_words(embed, ffn):
    return load().a3d_xattn_layer_words(embed, ffn)


def xattn_stack(x0, x0_stride_b, x0_stride_n, qpos, batch, nq, nk, embed, heads, ffn, nlayers, kv, kv_offset_bytes,
                kv_layer_stride, w, v, feat_out=None, feat_rows=0, feat_all_layers=False, qvec=None, logits=None):
    nqv = qvec.shape[0] if qvec is not None else 0
    assert w.is_cuda and w.is_contiguous()
    _check(load().a3d_xattn_stack(_ptr(_f32(x0)), x0_stride_b, x0_stride_n, _ptr(qpos), batch, nq, nk, embed, heads,
                                  ffn, nlayers, kv.data_ptr() + kv_offset_bytes, kv_layer_stride, w.data_ptr(),
                                  _ptr(_f32(v)), _ptr(feat_out), feat_rows, int(feat_all_layers), _ptr(qvec), nqv,
                                  _ptr(logits), _stream()), "a3d_xattn_stack")


def mask_logits(feat, qvec, logits):
    """feat (B, Ng, E), qvec (nqv, B, E) -> logits (nqv, B, Ng)"""
    b, ng, e = feat.shape
    _check(load().a3d_mask_logits(_ptr(_f32(feat)), _ptr(_f32(qvec)), b, ng, e, qvec.shape[0], _ptr(_f32(logits)),
                                  _stream()), "a3d_mask_logits")


def argmax_pick(logits, ghost):
    b, ng = logits.shape
    top = torch.empty(b, device=logits.device, dtype=torch.int32)
    pos = torch.empty(b, 3, device=logits.device, dtype=torch.float32)
    _check(load().a3d_argmax_pick(_ptr(_f32(logits)), _ptr(_f32(ghost)), b, ng, _ptr(top), _ptr(pos), _stream()),
           "a3d_argmax_pick")
    return top, pos


def sample_ghost(anchor, radius, bounds, batch, ng, seed, stream_id, device, counter=None):
    """counter: optional int64 device tensor (1,) added to stream_id ON THE DEVICE (CUDA-graph replays draw fresh points)."""
    out = torch.empty(batch, ng, 3, device=device, dtype=torch.float32)
    bd = (c_float * 6)(*[float(x) for x in (list(bounds[0]) + list(bounds[1]))])
    if counter is None:
        _check(load().a3d_sample_ghost(_ptr(anchor), float(radius), bd, batch, ng, int(seed) & (2**64 - 1),
                                       int(stream_id), _ptr(out), _stream()), "a3d_sample_ghost")
    else:
        _check(load().a3d_sample_ghost_ctr(_ptr(anchor), float(radius), bd, batch, ng, int(seed) & (2**64 - 1),
                                           int(stream_id), _ptr(counter), _ptr(out), _stream()), "a3d_sample_ghost_ctr")
    return out


def counter_add(counter, inc):
    _check(load().a3d_counter_add(_ptr(counter), int(inc), _stream()), "a3d_counter_add")


# ------------------------------------------------------------------------------------------------ training path
def attn_fwd(q, k, v, key_mask, heads, dropout_p=0.0, seed=0):
    """q (B,Nq,E), k/v (B,Nk,E), key_mask (B,Nk) uint8 or None -> o (B,Nq,E), lse (B*H,Nq)."""
    b, nq, e = q.shape
    nk = k.shape[1]
    o = torch.empty_like(q)
    lse = torch.empty(b * heads, nq, device=q.device, dtype=torch.float32)
    _check(load().a3d_attn_fwd(_ptr(_f32(q)), _ptr(_f32(k)), _ptr(_f32(v)), _ptr(key_mask), b, heads, nq, nk, e,
                               _ptr(o), _ptr(lse), float(dropout_p), int(seed) & (2**64 - 1), _stream()), "a3d_attn_fwd")
    return o, lse


def attn_bwd(q, k, v, key_mask, o, dout, lse, heads, dropout_p=0.0, seed=0):
    """-> dq (B,Nq,E), dk (B,Nk,E), dv (B,Nk,E)"""
    b, nq, e = q.shape
    nk = k.shape[1]
    dq = torch.zeros_like(q)
    dk, dv = torch.zeros_like(k), torch.zeros_like(v)
    dsum = torch.zeros(b * heads * (nq + 1), device=q.device, dtype=torch.float32)   # D per row, then max |dO| per (b, h)
    _check(load().a3d_attn_bwd(_ptr(_f32(q)), _ptr(_f32(k)), _ptr(_f32(v)), _ptr(key_mask), _ptr(_f32(o)),
                               _ptr(_f32(dout)), _ptr(_f32(lse)), b, heads, nq, nk, e, _ptr(dq), _ptr(dk), _ptr(dv),
                               _ptr(dsum), float(dropout_p), int(seed) & (2**64 - 1), _stream()), "a3d_attn_bwd")
    return dq, dk, dv


def rope_apply(x, pos, transpose=False):
    """x (..., E), pos (..., 3) with the same leading shape -> rotary(x) (inverse rotation if transpose)."""
    e = x.shape[-1]
    rows = x.numel() // e
    assert pos.numel() == rows * 3, "rope_apply: one xyz per row expected"
    out = torch.empty_like(x)
    _check(load().a3d_rope_apply(_ptr(_f32(x)), _ptr(_f32(pos)), rows, e, int(transpose), _ptr(out), _stream()),
           "a3d_rope_apply")
    return out


def gather_tokens_bwd(dtok, idx, batch, ncam, k, feat_shape, channels_last):
    """dtok (B, rows, E) -> gradient w.r.t. the (B*ncam, E, h, w) feature map, NCHW or channels-last storage."""
    _, e, h, w = feat_shape
    dfeat = torch.empty(feat_shape, device=dtok.device, dtype=torch.float32,
                        memory_format=torch.channels_last if channels_last else torch.contiguous_format).zero_()
    _check(load().a3d_gather_tokens_bwd(_ptr(_f32(dtok)), _ptr(idx), batch, ncam, e, h * w, k, dtok.shape[1],
                                        int(channels_last), dfeat.data_ptr(), _stream()), "a3d_gather_tokens_bwd")
    return dfeat


def linear_wgrad(dy, x, want_bias=True):
    """dy (rows, O), x (rows, I) -> dW (O, I), db (O,) or None."""
    rows, o = dy.shape
    i = x.shape[1]
    dw = torch.empty(o, i, device=dy.device, dtype=torch.float32)
    db = torch.empty(o, device=dy.device, dtype=torch.float32) if want_bias else None
    ws = torch.empty(load().a3d_linear_wgrad_workspace(rows, o, i), device=dy.device, dtype=torch.uint8)
    _check(load().a3d_linear_wgrad(_ptr(_f32(dy)), _ptr(_f32(x)), rows, o, i, _ptr(dw), _ptr(db), _ptr(ws), _stream()),
           "a3d_linear_wgrad")
    return dw, db


def linear_supported(out_features, in_features, transpose=False):
    return bool(load().a3d_linear_supported(int(out_features), int(in_features), int(transpose)))


def linear_rows(x, weight, bias=None, relu=False, transpose=False):
    """x (rows, K) fp32.  transpose=False: y = x W^T + b (W (O, I), K = I) -> (rows, O);
    transpose=True: y = x W (K = O) -> (rows, I): the data gradient of the same layer."""
    rows = x.shape[0]
    o, i = weight.shape
    n = i if transpose else o
    y = torch.empty(rows, n, device=x.device, dtype=torch.float32)
    ws = torch.empty(load().a3d_linear_workspace(o, i, int(transpose)), device=x.device, dtype=torch.uint8)
    _check(load().a3d_linear_fwd(_ptr(_f32(x)), _ptr(_f32(weight)), _ptr(bias), rows, o, i, int(relu), int(transpose),
                                 _ptr(y), _ptr(ws), _stream()), "a3d_linear_fwd")
    return y


def layernorm_fwd(x, res, gamma, beta, eps):
    """x, res (rows, E) -> y, z (= x + res, or x itself when res is None), mean, rstd (rows,)."""
    rows, e = x.shape
    y = torch.empty_like(x)
    z = torch.empty_like(x) if res is not None else x
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    _check(load().a3d_layernorm_fwd(_ptr(_f32(x)), _ptr(res), _ptr(_f32(gamma)), _ptr(_f32(beta)), rows, e, float(eps),
                                    _ptr(z) if res is not None else None, _ptr(y), _ptr(mean), _ptr(rstd), _stream()),
           "a3d_layernorm_fwd")
    return y, z, mean, rstd


def layernorm_bwd(dy, z, mean, rstd, gamma):
    """-> dz (rows, E), dgamma (E,), dbeta (E,)"""
    rows, e = dy.shape
    dz = torch.empty_like(dy)
    dg = torch.empty(e, device=dy.device, dtype=torch.float32)
    db = torch.empty_like(dg)
    ws = torch.empty(load().a3d_layernorm_bwd_workspace(e), device=dy.device, dtype=torch.uint8)
    _check(load().a3d_layernorm_bwd(_ptr(_f32(dy)), _ptr(_f32(z)), _ptr(mean), _ptr(rstd), _ptr(_f32(gamma)), rows, e,
                                    _ptr(dz), _ptr(dg), _ptr(db), _ptr(ws), _stream()), "a3d_layernorm_bwd")
    return dz, dg, db


def soft_ce(logits, ghost, gt, spread, label_smoothing=0.0, want_grad=True):
    """logits (B,Ng), ghost (B,Ng,3), gt (B,3) -> per-sample loss (B,), dlogits (B,Ng) or None."""
    b, ng = logits.shape
    loss = torch.empty(b, device=logits.device, dtype=torch.float32)
    dlog = torch.empty_like(logits) if want_grad else None
    _check(load().a3d_soft_ce(_ptr(_f32(logits)), _ptr(_f32(ghost)), _ptr(_f32(gt)), b, ng, float(spread),
                              float(label_smoothing), _ptr(loss), _ptr(dlog), _stream()), "a3d_soft_ce")
    return loss, dlog


# ------------------------------------------------------------------------------------------------ planner
def cd_pack_floats(which):
    return load().cd_pack_floats({"lang_v": 0, "ada_v": 1, "mlp_v": 2, "ada_row": 3, "lang_w": 4, "ada_w": 5,
                                  "mlp_w": 6}[which])


def _raw(t):
    """device pointer of a tensor or None (weights: any dtype)"""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return t.data_ptr()


def cd_ctx_lang(tok, nctx, kin, vin, w, v, nlayers):
    b, rows, e = tok.shape
    _check(load().cd_ctx_lang(_ptr(_f32(tok)), b, rows, nctx, e, 8, _ptr(_f32(kin)), _ptr(_f32(vin)), kin.shape[2],
                              _raw(w), _raw(v), nlayers, _stream()), "cd_ctx_lang")


def cd_step_begin(traj, wp_pe, t_idx, ada, ada_layers, enc1, enc2, enc2_b, lang_w, lang_v, lang_k, lang_vv, x_out,
                  next_wq, next_bq, next_ada_layer, q_out):
    b, length, _ = traj.shape
    n_instr = lang_k.shape[1] if lang_k is not None else 0
    _check(load().cd_step_begin(_ptr(_f32(traj)), b, length, _ptr(wp_pe), _ptr(t_idx), _ptr(ada), ada_layers,
                                _raw(enc1), _raw(enc2), _raw(enc2_b), _raw(lang_w), _raw(lang_v), _ptr(lang_k),
                                _ptr(lang_vv), n_instr, _ptr(x_out), _raw(next_wq), _raw(next_bq), next_ada_layer,
                                _ptr(q_out), _stream()), "cd_step_begin")


def cd_cross_part_floats(batch):
    return load().cd_cross_part_floats(batch)


def cd_cross(q, kv, kv_offset_bytes, batch, nk, att):
    _check(load().cd_cross(_ptr(q), kv.data_ptr() + kv_offset_bytes, batch, nk, 8, _ptr(att), _stream()), "cd_cross")


def cd_post(traj, mask, wp_pe, t_idx, ada, ada_layers, ada_layer, x_in, att, layer_w, layer_v, x_out, reg_w=None,
            reg_v=None, reg_out=None, reg_dim=0, next_src=None, next_wq=None, next_bq=None, next_ada_layer=0, q_out=None,
            update=None):
    """update = dict(last_step, traj_out, pos_upd, cond_data, cond_mask, coef (6 floats), noise_pos, noise_rot) or None"""
    b, length, _ = traj.shape
    u = update or {}
    coef = (c_float * 6)(*[float(x) for x in u.get("coef", [0.0] * 6)])
    _check(load().cd_post(_ptr(_f32(traj)), b, length, _ptr(mask), _ptr(wp_pe), _ptr(t_idx), _ptr(ada), ada_layers,
                          ada_layer, _ptr(x_in), _ptr(att), _raw(layer_w), _raw(layer_v), _ptr(x_out), _raw(reg_w),
                          _raw(reg_v), _ptr(reg_out), reg_dim, _ptr(next_src), _raw(next_wq), _raw(next_bq),
                          next_ada_layer, _ptr(q_out), int(update is not None), int(u.get("last_step", 0)),
                          _ptr(u.get("traj_out")), _ptr(u.get("pos_upd")), _ptr(u.get("cond_data")),
                          _ptr(u.get("cond_mask")), coef, _ptr(u.get("noise_pos")), _ptr(u.get("noise_rot")),
                          _stream()), "cd_post")


def cd_denoise_loop(traj, cond, cond_mask, mask, wp_pe, timesteps, ada, n_traj_layers, coef, noise_pos, noise_rot, enc1, enc2,
                    enc2_b, lang_w, lang_v, lang_k, lang_vv, ada_w, ada_v, pos_reg, rot_reg, kv, kv_set_bytes, nk):
    """The whole DDPM sampling loop in ONE persistent cluster kernel (csrc/cd_loop.cu): traj (B, L, 9) is updated in
    place from x_T to x_0.  timesteps: int32 (n_steps,) on the device; coef: fp32 (T, 6) on the device; ada_w / ada_v:
    lists of the per-layer AdaW / AdaV packs (shared layers, then the two position and the two rotation layers)."""
    b, length, _ = traj.shape
    nl = len(ada_w)
    wp = (c_void_p * nl)(*[_raw(t) for t in ada_w])
    vp = (c_void_p * nl)(*[_raw(t) for t in ada_v])
    n_instr = lang_k.shape[1] if lang_k is not None else 0
    _check(load().cd_denoise_loop(_ptr(_f32(traj)), b, length, int(timesteps.numel()), _ptr(_f32(cond)), _ptr(cond_mask), _ptr(mask),
                                  _ptr(wp_pe), _ptr(timesteps), _ptr(ada), nl, n_traj_layers, _ptr(_f32(coef)), _ptr(noise_pos),
                                  _ptr(noise_rot), _raw(enc1), _raw(enc2), _raw(enc2_b), _raw(lang_w), _raw(lang_v), _ptr(lang_k),
                                  _ptr(lang_vv), n_instr, wp, vp, _raw(pos_reg[0]), _raw(pos_reg[1]), _raw(rot_reg[0]),
                                  _raw(rot_reg[1]), kv.data_ptr(), kv_set_bytes, nk, _stream()), "cd_denoise_loop")
