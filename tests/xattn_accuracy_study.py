"""GPU study (not a pytest module): accuracy of every attention core of a3d_xattn_stack against the fp64 CPU oracle at
unit / x4 / x16 logit gain (the inflated cases exercise the overflow verdict + safe-mode replay of the tcgen05 cores).
    python tests/xattn_accuracy_study.py          # one JSON per line
Lives under tests/ because it imports oracle/ (test infrastructure)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from act3d_chained_diffuser_b200 import lib  # noqa: E402
from oracle.attention import relative_cross_attn_stack  # noqa: E402
from oracle.rope import rope3d_table  # noqa: E402
from tools.xattn_study import E, H, run, setup  # noqa: E402

VARIANTS = [(2, 0), (4, 0), (6, 0), (6, 6), (6, 10)]          # (xattn_core, xattn_poly | xattn6_np)


def main():
    lib.load()
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
        for core, poly in VARIANTS:
            lib.set_option("xattn_core", core)
            lib.set_option("xattn6_np" if core == 6 else "xattn_poly", poly)
            feat, logits, _ = run(b, nq, nk, t)
            print(json.dumps({"case": tag, "core": core, "poly": poly, "feat_rel_l2": rel(feat[0], want64),
                              "logit_rel_l2": rel(logits, lg64),
                              "feat_maxabs_over_max": ((feat[0].double() - want64).abs().max() / want64.abs().max()).item()}),
                  flush=True)
    lib.set_option("xattn_core", 0)
    lib.set_option("xattn_poly", 0)
    lib.set_option("xattn6_np", 6)


if __name__ == "__main__":
    main()
