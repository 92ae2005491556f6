"""GPU: the inference trunk (BN-folded, channels-last, cuDNN fused conv+bias+ReLU calls, restricted FPN)
computes the same features as the plain modules the reference runs (both sides are TF32 cuDNN convolutions,
so agreement is at TF32 level)."""
import pytest
import torch

from tests.golden import cases

pytestmark = pytest.mark.gpu


def test_fused_eval_trunk_matches_plain_modules():
    from model import Act3D
    torch.manual_seed(0)
    m = Act3D(**dict(cases.ACT3D_KW, use_instruction=False)).eval().cuda()
    for mod in m.backbone.modules():                  # non-trivial BN statistics
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.1)
    x = torch.rand(4, 3, 256, 256, device="cuda")
    with torch.no_grad():
        plain = m.feature_pyramid(m.backbone(m.normalize(x)))
        fused = m._eval_trunk(m.normalize, m.backbone, m.feature_pyramid, x, needed=("res3", "res1"))
    assert m._eval_trunk._fused_ok, "cuDNN fused conv entry points were not used"
    assert set(fused) == {"res1", "res3"}
    for k in fused:
        rel = ((fused[k].float() - plain[k]).norm() / plain[k].norm()).item()
        assert rel < 5e-3, (k, rel)
    # fp32-exact check of the folding itself with TF32 disabled
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            plain = m.feature_pyramid(m.backbone(m.normalize(x)))
            fused = m._eval_trunk(m.normalize, m.backbone, m.feature_pyramid, x, needed=("res3", "res1"))
        for k in fused:
            rel = ((fused[k].float() - plain[k]).norm() / plain[k].norm()).item()
            assert rel < 2e-5, (k, rel)
        # deferred output-conv bias: map + bias must reproduce the full map
        with torch.no_grad():
            maps, biases = m._eval_trunk(m.normalize, m.backbone, m.feature_pyramid, x, needed=("res3", "res1"),
                                         defer_bias=True)
        assert set(biases) == {"res1", "res3"}
        for k in maps:
            full = maps[k].float() + biases[k].view(1, -1, 1, 1)
            rel = ((full - plain[k]).norm() / plain[k].norm()).item()
            assert rel < 2e-5, (k, rel)
    finally:
        torch.backends.cudnn.allow_tf32 = True
