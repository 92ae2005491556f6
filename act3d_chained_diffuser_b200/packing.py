"""Weight packing: nn.Parameters -> the flat fp32 buffers the kernels read.

All matrices are stored K-MAJOR ([in_features][padded out_features]) so that the register-tiled
GEMMs read 16 consecutive output columns with four 16-byte loads; vectors are padded to the same
width.  Packed buffers are cached per (parameter storage, version) so inference repacks nothing.
"""
import math

import torch

LOG2E = 1.4426950408889634


def _pad_cols(m, width):
    out = m.new_zeros(m.shape[0], width)
    out[:, : m.shape[1]] = m
    return out


def _pad_vec(v, width):
    out = v.new_zeros(width)
    out[: v.shape[0]] = v
    return out


def _kmajor(weight, width):
    """nn.Linear weight (out, in) -> (in, width) with W^T in the first `out` columns."""
    return _pad_cols(weight.detach().float().t().contiguous(), width)


def pack_xattn_layer(attn, norm, ffw, embed, heads):
    """One layer of an Act3D stack -> (w int32 words, v floats) in the Xa2 order (csrc/a3d_xattn_common.cuh):
    w = {W_q, W_o, W_1, W_2} as fragment-ordered fp16 (hi, lo) [K=64][N=64] matrices (mma.sync kernel) followed by the
    same four matrices as tcgen05 operand images (pack_umma_weight), v = {b_q, b_o, g1, be1,
    b_1, b_2, g2, be2} (64 floats each).  The q projection carries hd^-1/2 (multihead_custom_attention.py:244,325)
    and log2(e) so the kernel's softmax is a bare exp2.  W_o's input (K) index is head-padded: k = 16 h + d
    <- attention dim 15 h + d (the pad slot carries the softmax denominator in the kernel; its weight is 0)."""
    ep = 16 * heads
    e = embed
    hd = e // heads
    scale = (float(hd) ** -0.5) * LOG2E
    w_in, b_in = attn.in_proj_weight.detach().float(), attn.in_proj_bias.detach().float()
    wo = attn.out_proj.weight.detach().float()
    wo_p = wo.new_zeros(e, ep)
    for h in range(heads):
        wo_p[:, 16 * h:16 * h + hd] = wo[:, hd * h:hd * (h + 1)]
    mats = [w_in[:e] * scale, wo_p, ffw.linear1.weight, ffw.linear2.weight]
    w = torch.cat([pack_mma_weight(m, ep, ep) for m in mats] + [pack_umma_weight(m, ep, ep) for m in mats])
    vec = lambda t: _pad_vec(t.detach().float(), ep)
    v = torch.cat([_pad_vec(b_in[:e] * scale, ep), vec(attn.out_proj.bias), vec(norm.weight), vec(norm.bias),
                   vec(ffw.linear1.bias), vec(ffw.linear2.bias), vec(ffw.norm.weight), vec(ffw.norm.bias)])
    return w.contiguous(), v.contiguous()


def pack_kv_set(attn, embed, heads):
    """K/V in-projection of one attention layer -> (wkv, bkv) for a3d_ctx_kv: wkv = the [2*EP] x [EP] matrix
    (rows [0,E) = W_k, rows [EP, EP+E) = W_v: slices W[E:2E], W[2E:3E], multihead_custom_attention.py:268-303)
    as fragment-ordered split-fp16 words (pack_mma_weight), bkv [2*EP] fp32."""
    ep = 16 * heads
    e = embed
    w_in, b_in = attn.in_proj_weight.detach().float(), attn.in_proj_bias.detach().float()
    full = w_in.new_zeros(2 * ep, ep)
    full[:e, :e] = w_in[e:2 * e]
    full[ep:ep + e, :e] = w_in[2 * e:]
    bkv = torch.cat([_pad_vec(b_in[e:2 * e], ep), _pad_vec(b_in[2 * e:], ep)])
    return pack_mma_weight(full, ep, 2 * ep).contiguous(), bkv.contiguous()


class PackCache:
    """Memoise packed buffers on the identity + version of the source parameters."""

    def __init__(self):
        self._store = {}

    @staticmethod
    def _sig(params):
        return tuple((p.data_ptr(), p._version, p.device) for p in params)

    def get(self, key, params, builder):
        sig = self._sig(params)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        val = builder()
        self._store[key] = (sig, val)
        return val

    def clear(self):
        self._store.clear()


# ------------------------------------------------------------------------------------------------ planner packs
def pack_mma_weight(weight, kpad, npad):
    """nn.Linear weight (N_out, K_in) -> int32 buffer in the B-fragment order of csrc/a3d_mma_gemm.cuh:
    [kstep][ntile][lane] -> {b0_hi, b1_hi, b0_lo, b1_lo} with hi = fp16(w), lo = fp16((w - hi) * 2^11);
    b0 = (w[n][k0], w[n][k0+1]), b1 = (w[n][k0+8], w[n][k0+9]), n = 8*ntile + lane//4, k0 = 16*kstep + 2*(lane%4)."""
    w = weight.detach().float()
    full = w.new_zeros(npad, kpad)
    full[: w.shape[0], : w.shape[1]] = w
    hi = full.half()
    lo = ((full - hi.float()) * 2048.0).half()
    ks, nt = kpad // 16, npad // 8

    def frag(p):
        t = p.view(nt, 8, ks, 2, 4, 2).permute(2, 0, 1, 4, 3, 5)          # (ks, nt, g, q, hsel, pair)
        return t.reshape(ks, nt, 32, 2, 2)
    both = torch.stack([frag(hi), frag(lo)], dim=3).contiguous()          # (ks, nt, lane, plane, hsel, pair)
    return both.view(torch.int32).reshape(-1)


def pack_umma_weight(weight, kpad, npad):
    """nn.Linear weight (N_out, K_in) -> int32 buffer holding the B operand of a tcgen05.mma (csrc/a3d_xattn6.cu):
    [plane: hi | lo][k-slab = K_in // 16][n][16 halves], hi = fp16(w), lo = fp16((w - hi) * 2^11).  One slab is a
    K-major [npad rows][32 bytes] tile in the canonical SWIZZLE_32B layout -- exactly the layout of a K tile image of
    a3d_ctx_kv with `n` in the place of the key: inside a 32-byte row the two 16-byte halves are swapped when (n>>2)&1."""
    w = weight.detach().float()
    full = w.new_zeros(npad, kpad)
    full[: w.shape[0], : w.shape[1]] = w
    hi = full.half()
    lo = ((full - hi.float()) * 2048.0).half()
    swap = ((torch.arange(npad, device=w.device) >> 2) & 1).bool()

    def img(p):
        t = p.view(npad, kpad // 16, 2, 8).permute(1, 0, 2, 3).contiguous()      # (slab, n, 16-byte half, 8)
        t[:, swap] = t[:, swap].flip(2)
        return t.reshape(-1)
    return torch.cat([img(hi), img(lo)]).view(torch.int32)


def _scaled_q(attn, e, heads):
    scale = (float(e // heads) ** -0.5) * LOG2E
    w_in, b_in = attn.in_proj_weight.detach().float(), attn.in_proj_bias.detach().float()
    return w_in[:e] * scale, b_in[:e] * scale


def _vec(t, width):
    return _pad_vec(t.detach().float(), width)


def pack_lang_layer(layer, e=120, heads=8, ffp=512):
    """ParallelAttentionLayer without adaLN / self-attention (vl_attention, traj_lang_attention) ->
    (LangW int32 fragments {WQ WO W1 W2}, LangV floats {BQ BO G12 B12 B1[512] B2 G122 B122}); csrc/cd_denoiser.cu."""
    ep = 16 * heads
    wq, bq = _scaled_q(layer.cross_12, e, heads)
    w = torch.cat([pack_mma_weight(wq, ep, ep), pack_mma_weight(layer.cross_12.out_proj.weight, ep, ep),
                   pack_mma_weight(layer.ffn_12[0].weight, ep, ffp), pack_mma_weight(layer.ffn_12[3].weight, ffp, ep)])
    v = torch.cat([_pad_vec(bq, ep), _vec(layer.cross_12.out_proj.bias, ep), _vec(layer.norm_12.weight, ep),
                   _vec(layer.norm_12.bias, ep), _vec(layer.ffn_12[0].bias, ffp), _vec(layer.ffn_12[3].bias, ep),
                   _vec(layer.norm_122.weight, ep), _vec(layer.norm_122.bias, ep)])
    return w.contiguous(), v.contiguous()


def pack_ada_layer(layer, e=120, heads=8, ffp=512):
    """adaLN cross + self + FFN layer -> (AdaW {C_WQ C_WO S_WQ S_WK S_WV S_WO W1 W2},
    AdaV {C_BQ C_BO G12 B12 S_BQ S_BK S_BV S_BO G1 B1N B1[512] B2 G122 B122}).
    Both q projections carry hd^-1/2 * log2(e)."""
    ep = 16 * heads
    cq, cbq = _scaled_q(layer.cross_12, e, heads)
    sq, sbq = _scaled_q(layer.sa1, e, heads)
    w_in, b_in = layer.sa1.in_proj_weight.detach().float(), layer.sa1.in_proj_bias.detach().float()
    w = torch.cat([pack_mma_weight(cq, ep, ep), pack_mma_weight(layer.cross_12.out_proj.weight, ep, ep),
                   pack_mma_weight(sq, ep, ep), pack_mma_weight(w_in[e:2 * e], ep, ep),
                   pack_mma_weight(w_in[2 * e:], ep, ep), pack_mma_weight(layer.sa1.out_proj.weight, ep, ep),
                   pack_mma_weight(layer.ffn_12[0].weight, ep, ffp), pack_mma_weight(layer.ffn_12[3].weight, ffp, ep)])
    v = torch.cat([_pad_vec(cbq, ep), _vec(layer.cross_12.out_proj.bias, ep), _vec(layer.norm_12.weight, ep),
                   _vec(layer.norm_12.bias, ep), _pad_vec(sbq, ep), _pad_vec(b_in[e:2 * e], ep), _pad_vec(b_in[2 * e:], ep),
                   _vec(layer.sa1.out_proj.bias, ep), _vec(layer.norm_1.weight, ep), _vec(layer.norm_1.bias, ep),
                   _vec(layer.ffn_12[0].bias, ffp), _vec(layer.ffn_12[3].bias, ep),
                   _vec(layer.norm_122.weight, ep), _vec(layer.norm_122.bias, ep)])
    return w.contiguous(), v.contiguous()


def pack_mlp(seq, e=120, ep=128):
    """nn.Sequential(Linear(E, E), ReLU, [Dropout], Linear(E, out<=EP)) -> (MlpW {W1 W2}, MlpV {B1 B2})."""
    first, last = seq[0], seq[-1]
    w = torch.cat([pack_mma_weight(first.weight, ep, ep), pack_mma_weight(last.weight, ep, ep)])
    v = torch.cat([_vec(first.bias, ep), _vec(last.bias, ep)])
    return w.contiguous(), v.contiguous()


def pack_traj_encoder(seq, e=120, ep=128):
    """traj_encoder = Linear(9, E) -> ReLU -> Dropout -> Linear(E, E): first layer stays an fp32 K-major [9][EP]
    matrix + bias (K = 9 is not worth a tensor-core pass), second layer in fragment order."""
    first, last = seq[0], seq[-1]
    enc1 = torch.cat([_kmajor(first.weight, ep).reshape(-1), _vec(first.bias, ep)])
    return enc1.contiguous(), pack_mma_weight(last.weight, ep, ep).contiguous(), _vec(last.bias, ep).contiguous()
