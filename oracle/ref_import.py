"""Import the UNMODIFIED reference from /root/reference (build container only).

Test infrastructure only.  /root/reference does not exist on the GPU box; nothing in
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.  It is used by
``tests/golden/make_golden.py`` (fixture generation) and by the container-only parts of
``tests/test_dropin_contract.py`` (constructor / forward signatures and state_dict keys against the real
reference; they skip where /root/reference is absent).

Two packages the reference imports at module scope are not installed here
(SURVEY.md App. B.1):
  * ``clip`` (model/utils/clip.py:5-6) -- stubbed with an empty ``ModifiedResNet`` base;
    only subclassed, never instantiated with backbone="resnet".
  * ``diffusers`` (diffusion_model.py:1) -- stubbed with oracle.ddpm.DDPMScheduler
    (the restatement; see the "parity unpinned" note there).
Our own repo also has a top-level package called ``model`` (the drop-in boundary), so the
reference's ``model`` / ``utils`` packages are imported under a private sys.modules
snapshot and handed back as objects; the caller's sys.modules is restored afterwards.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ACT3D_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model", "keypose_optimization"))


def _stub_modules():
    import torch.nn as nn
    from . import ddpm

    clip = types.ModuleType("clip")
    clip_model = types.ModuleType("clip.model")

    class ModifiedResNet(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
    clip_model.ModifiedResNet = ModifiedResNet
    clip.model = clip_model
    clip.load = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("clip weights are not available offline"))

    diffusers = types.ModuleType("diffusers")
    schedulers = types.ModuleType("diffusers.schedulers")
    sched_ddpm = types.ModuleType("diffusers.schedulers.scheduling_ddpm")
    sched_ddpm.DDPMScheduler = ddpm.DDPMScheduler
    schedulers.scheduling_ddpm = sched_ddpm
    diffusers.schedulers = schedulers
    return {
        "clip": clip, "clip.model": clip_model,
        "diffusers": diffusers, "diffusers.schedulers": schedulers,
        "diffusers.schedulers.scheduling_ddpm": sched_ddpm,
    }


_CACHE = {}


def load_reference():
    """Returns a namespace with the reference's ``Act3D``, ``DiffusionPlanner`` classes and the
    ``model`` package object, imported from REFERENCE_ROOT without touching our ``model`` package."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")

    shadowed = {k: v for k, v in sys.modules.items()
                if k == "model" or k.startswith("model.") or k == "utils" or k.startswith("utils.")}
    for k in shadowed:
        del sys.modules[k]
    stubs = _stub_modules()
    had = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        ref_model = importlib.import_module("model")
        ns = types.SimpleNamespace(
            model=ref_model,
            Act3D=ref_model.Act3D,
            DiffusionPlanner=ref_model.DiffusionPlanner,
            layers=importlib.import_module("model.utils.layers"),
            position_encodings=importlib.import_module("model.utils.position_encodings"),
            mha=importlib.import_module("model.utils.multihead_custom_attention"),
            utils=importlib.import_module("model.utils.utils"),
        )
    finally:
        sys.path.remove(REFERENCE_ROOT)
        ref_loaded = [k for k in sys.modules
                      if k == "model" or k.startswith("model.") or k == "utils" or k.startswith("utils.")]
        for k in ref_loaded:
            del sys.modules[k]
        sys.modules.update(shadowed)
        for k, v in had.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _CACHE["ns"] = ns
    return ns
