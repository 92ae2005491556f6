"""GPU timing of the fused cross-attention kernel at the C2 shape for every attention core (one JSON per line).
The accuracy half of the study (against the fp64 oracle) lives in tests/xattn_accuracy_study.py: only tests/ may
import oracle/."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from act3d_chained_diffuser_b200 import lib  # noqa: E402
from act3d_chained_diffuser_b200.packing import pack_kv_set, pack_xattn_layer  # noqa: E402
from act3d_chained_diffuser_b200.params import XAttnStackParams  # noqa: E402
from tests.golden import synth  # noqa: E402


def _stack_sd(e, heads, layers):
    """state_dict skeleton of a RelativeCrossAttentionModule (layers.py:335-343 key names), filled by name."""
    sd = {}
    for l in range(layers):
        a, f = f"attn_layers.{l}.", f"ffw_layers.{l}."
        for k, shape in ((a + "multihead_attn.in_proj_weight", (3 * e, e)), (a + "multihead_attn.in_proj_bias", (3 * e,)),
                         (a + "multihead_attn.out_proj.weight", (e, e)), (a + "multihead_attn.out_proj.bias", (e,)),
                         (a + "norm.weight", (e,)), (a + "norm.bias", (e,)),
                         (f + "linear1.weight", (e, e)), (f + "linear1.bias", (e,)), (f + "linear2.weight", (e, e)),
                         (f + "linear2.bias", (e,)), (f + "norm.weight", (e,)), (f + "norm.bias", (e,))):
            sd[k] = torch.empty(*shape)
    return synth.fill_state_dict(sd)

E, H = 60, 4


def setup(b, nq, nk, gain=1.0):
    sd = _stack_sd(E, H, 2)
    sd["attn_layers.0.multihead_attn.in_proj_weight"][:2 * E] *= gain
    sd["attn_layers.1.multihead_attn.in_proj_weight"][:2 * E] *= gain
    stack = XAttnStackParams(E, H, 2)
    stack.load_state_dict(sd)
    x0 = synth.normal("st.x0", (1, E))
    q_xyz = synth.points_in_bounds("st.q", (b, nq))
    ctx = synth.normal("st.ctx", (b, nk, E))
    c_xyz = synth.points_in_bounds("st.c", (b, nk))
    qvec = synth.normal("st.qv", (2, b, E))
    lp = [pack_xattn_layer(stack.attn_layers[l].multihead_attn, stack.attn_layers[l].norm, stack.ffw_layers[l], E, H)
          for l in range(2)]
    w = torch.cat([x[0] for x in lp]).cuda()
    wv = torch.cat([x[1] for x in lp]).cuda()
    packs = [pack_kv_set(stack.attn_layers[l].multihead_attn, E, H) for l in range(2)]
    wkv = torch.stack([p[0] for p in packs]).cuda()
    bkv = torch.stack([p[1] for p in packs]).cuda()
    return sd, x0, q_xyz, ctx, c_xyz, qvec, w, wv, wkv, bkv


def run(b, nq, nk, tensors, iters=1):
    sd, x0, q_xyz, ctx, c_xyz, qvec, w, wv, wkv, bkv = tensors
    kv = lib.ctx_kv(ctx.cuda(), c_xyz.cuda(), nk, H, wkv, bkv, [1, 1])
    feat = torch.empty(1, b, nq, E, device="cuda")
    logits = torch.empty(2, b, nq, device="cuda")
    x0d, qd, qvd = x0.cuda(), q_xyz.cuda(), qvec.cuda()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        lib.xattn_stack(x0d, 0, 0, qd, b, nq, nk, E, H, E, 2, kv, 0, lib.kv_bytes(1, b, nk, H), w, wv,
                        feat_out=feat, feat_rows=nq, qvec=qvd, logits=logits)
    e.record()
    torch.cuda.synchronize()
    return feat.cpu(), logits.cpu(), s.elapsed_time(e) / iters


VARIANTS = [(4, 0), (6, 0), (6, 4), (6, 6), (6, 8)]          # (xattn_core, xattn_poly | xattn6_np)


def decomp(core, np_):
    """Marginal cost per key tile and fixed cost per launch: time the C2 query shape at several key counts."""
    lib.load()
    lib.set_option("xattn_core", core)
    if core == 6:
        lib.set_option("xattn6_np", np_)
    b, nq = 16, 16384
    pts = []
    for nk in (64 * 8, 64 * 16, 64 * 32, 64 * 65):
        t = setup(b, nq, nk, 1.0)
        run(b, nq, nk, t, iters=2)
        _, _, ms = run(b, nq, nk, t, iters=5)
        pts.append((nk // 64, ms))
    (t0, m0), (t1, m1) = pts[0], pts[-1]
    slope = (m1 - m0) / (t1 - t0)
    print(json.dumps({"core": core, "np": np_, "ms_by_tiles": pts, "ms_per_tile": slope, "fixed_ms": m0 - slope * t0}), flush=True)
    lib.set_option("xattn_core", 0)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "decomp":
        for core, np_ in ((4, 0), (6, 0), (6, 6)):
            decomp(core, np_)
        return
    lib.load()
    b, nq, nk = 16, 16384, 4150
    t = setup(b, nq, nk, 1.0)
    flops = 4.0 * nq * nk * E * 2 * b
    for core, poly in VARIANTS:
        lib.set_option("xattn_core", core)
        lib.set_option("xattn6_np" if core == 6 else "xattn_poly", poly)
        run(b, nq, nk, t, iters=2)
        _, _, ms = run(b, nq, nk, t, iters=5)
        print(json.dumps({"shape": [b, nq, nk], "core": core, "poly": poly, "ms_per_launch": ms, "tflops_true": flops / ms / 1e9,
                          "score_elems_per_clk_per_sm@1.965GHz": b * nq * nk * H * 2 / (ms * 1e-3) / 148 / 1.965e9}),
              flush=True)
    lib.set_option("xattn_core", 0)
    lib.set_option("xattn_poly", 0)
    lib.set_option("xattn6_np", 6)


if __name__ == "__main__":
    main()
