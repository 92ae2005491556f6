"""CPU: the drop-in boundary -- our modules expose the reference's constructor arguments and
state_dict keys.  The key comparison against the real reference only
runs where /root/reference exists; the frozen key listing in tests/golden/state_keys.json
(generated from the reference by this file's __main__) runs everywhere."""
import inspect
import json
import os

import pytest
import torch

from model import Act3D, DiffusionPlanner
from tests.golden import cases

KEYS = os.path.join(os.path.dirname(__file__), "golden", "state_keys.json")

VARIANTS = {
    "act3d_default": (dict(cases.ACT3D_KW, use_instruction=False), "act3d"),
    "act3d_instr_untied": (dict(cases.ACT3D_KW, use_instruction=True, weight_tying=False, gp_emb_tying=False,
                                regress_position_offset=True, rotation_parametrization="6D_from_top_ghost"), "act3d"),
    "planner_shipped": (dict(cases.PLANNER_KW), "planner"),
    "planner_untied_nogoal": (dict(cases.PLANNER_KW, weight_tying=False, use_goal=False, use_instruction=False,
                                   num_query_cross_attn_layers=4), "planner"),
}


def _shape_map(m):
    return {k: list(v.shape) for k, v in m.state_dict().items()}


def _ours(kind, kw):
    return (Act3D if kind == "act3d" else DiffusionPlanner)(**kw)


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_state_dict_keys_match_frozen_listing(name):
    kw, kind = VARIANTS[name]
    want = json.load(open(KEYS))[name]
    got = _shape_map(_ours(kind, kw))
    assert set(got) == set(want), (sorted(set(want) - set(got))[:5], sorted(set(got) - set(want))[:5])
    for k in want:
        assert got[k] == want[k], (k, got[k], want[k])


@pytest.mark.reference
@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_state_dict_loads_from_reference(name):
    from oracle.ref_import import load_reference
    ref = load_reference()
    kw, kind = VARIANTS[name]
    rm = (ref.Act3D if kind == "act3d" else ref.DiffusionPlanner)(**kw)
    ours = _ours(kind, kw)
    missing, unexpected = ours.load_state_dict(rm.state_dict(), strict=True)
    assert not missing and not unexpected
    # trainable parameter count and tying structure are the same
    assert sum(p.numel() for p in ours.parameters() if p.requires_grad) == \
        sum(p.numel() for p in rm.parameters() if p.requires_grad)


@pytest.mark.reference
def test_constructor_and_forward_signatures():
    from oracle.ref_import import load_reference
    ref = load_reference()
    for ours, theirs in ((Act3D, ref.Act3D), (DiffusionPlanner, ref.DiffusionPlanner)):
        assert list(inspect.signature(ours.__init__).parameters) == list(inspect.signature(theirs.__init__).parameters)
        for (n1, p1), (n2, p2) in zip(inspect.signature(ours.__init__).parameters.items(),
                                      inspect.signature(theirs.__init__).parameters.items()):
            assert p1.default == p2.default, (n1, p1.default, p2.default)
        assert list(inspect.signature(ours.forward).parameters) == list(inspect.signature(theirs.forward).parameters)
    assert list(inspect.signature(DiffusionPlanner.compute_trajectory).parameters) == \
        list(inspect.signature(ref.DiffusionPlanner.compute_trajectory).parameters)


def test_cpu_tensors_fail_loudly():
    m = Act3D(**dict(cases.ACT3D_KW, use_instruction=False)).eval()
    inp = cases.act3d_inputs(batch=1, ncam=1)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        m(inp["visible_rgb"], inp["visible_pcd"], inp["instruction"], inp["curr_gripper"])


if __name__ == "__main__":          # regenerate the frozen listing from the real reference
    from oracle.ref_import import load_reference
    ref = load_reference()
    out = {}
    for name, (kw, kind) in VARIANTS.items():
        out[name] = _shape_map((ref.Act3D if kind == "act3d" else ref.DiffusionPlanner)(**kw))
    json.dump(out, open(KEYS, "w"), indent=0, sort_keys=True)
    print("wrote", KEYS)
