"""GPU study of the fused cross-attention kernel: accuracy vs an fp64 oracle and speed at the C2 shape.
Prints one JSON per line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from act3d_chained_diffuser_b200 import lib  # noqa: E402
from act3d_chained_diffuser_b200.packing import pack_kv_set, pack_xattn_layer  # noqa: E402
from act3d_chained_diffuser_b200.params import XAttnStackParams  # noqa: E402
from oracle.attention import relative_cross_attn_stack  # noqa: E402
from oracle.rope import rope3d_table  # noqa: E402
from tests.golden import synth  # noqa: E402
from tests.test_oracle_golden import _stack_sd  # noqa: E402

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


def main():
    lib.load()
    variants = [(2, 0), (4, 0), (5, 0)]
    for gain, tag in ((1.0, "unit-gain"), (4.0, "peaky x4"), (16.0, "peaky x16")):
        b, nq, nk = 2, 1024, 4097
        t = setup(b, nq, nk, gain)
        sd, x0, q_xyz, ctx, c_xyz, qvec = t[:6]
        sd64 = {k: v.double() for k, v in sd.items()}
        q_in = x0.double().unsqueeze(0).repeat(nq, b, 1)
        want64 = relative_cross_attn_stack(sd64, "", H, 2, q_in, ctx.double().transpose(0, 1),
                                           rope3d_table(q_xyz.double(), E), rope3d_table(c_xyz.double(), E))[-1].transpose(0, 1)
        lg64 = torch.einsum("jbc,bnc->jbn", qvec.double(), want64)
        rel = lambda a, r: ((a.double() - r).norm() / r.norm()).item()
        for core, poly in variants:
            lib.set_option("xattn_core", core)
            lib.set_option("xattn_poly", poly)
            feat, logits, _ = run(b, nq, nk, t)
            print(json.dumps({"case": tag, "core": core, "poly": poly, "feat_rel_l2": rel(feat[0], want64),
                              "logit_rel_l2": rel(logits, lg64),
                              "feat_maxabs_over_max": ((feat[0].double() - want64).abs().max() / want64.abs().max()).item()}),
                  flush=True)
    b, nq, nk = 16, 16384, 4150
    t = setup(b, nq, nk, 1.0)
    flops = 4.0 * nq * nk * E * 2 * b
    for core, poly in variants:
        lib.set_option("xattn_core", core)
        lib.set_option("xattn_poly", poly)
        run(b, nq, nk, t, iters=2)
        _, _, ms = run(b, nq, nk, t, iters=5)
        print(json.dumps({"shape": [b, nq, nk], "core": core, "poly": poly, "ms_per_launch": ms, "tflops_true": flops / ms / 1e9,
                          "score_elems_per_clk_per_sm@1.965GHz": b * nq * nk * H * 2 / (ms * 1e-3) / 148 / 1.965e9}),
              flush=True)
    lib.set_option("xattn_core", 0)
    lib.set_option("xattn_poly", 0)


if __name__ == "__main__":
    main()
