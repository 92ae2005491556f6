"""Frozen image backbones + FPN (stay PyTorch/cuDNN: dense convolutions).

``backbone="resnet"`` is torchvision's ResNet-50 returning its five stage outputs
(reference: model/utils/resnet.py:35-56); ``backbone="clip"`` needs the third-party ``clip``
package and its downloaded RN50 weights (reference: model/utils/clip.py:9-43), exactly like
the reference -- it raises ImportError here when that package is absent.
"""
import torch
from torch.nn.modules.utils import _pair
from torchvision import transforms
from torchvision.models.resnet import Bottleneck, ResNet

from . import lib


class ResNetStages(ResNet):
    """torchvision ResNet whose forward returns {"res1".."res5"} (strides 2,4,8,16,32)."""

    def _forward_impl(self, x):
        stem = self.relu(self.bn1(self.conv1(x)))
        c2 = self.layer1(self.maxpool(stem))
        c3 = self.layer2(c2)
        c4 = self.layer3(c3)
        c5 = self.layer4(c4)
        return {"res1": stem, "res2": c2, "res3": c3, "res4": c4, "res5": c5}


def build_backbone(kind):
    if kind == "resnet":
        net = ResNetStages(Bottleneck, [3, 4, 6, 3])
        norm = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
        return net, norm
    if kind == "clip":
        try:
            import clip  # noqa: F401
            from clip.model import ModifiedResNet
        except ImportError as exc:  # same failure mode as the reference (model/utils/clip.py:5-6)
            raise ImportError("backbone='clip' needs the OpenAI `clip` package and its RN50 weights "
                              "(not installed / not available offline); use backbone='resnet'") from exc

        class ClipStages(ModifiedResNet):
            def forward(self, x):
                x = x.type(self.conv1.weight.dtype)
                x = self.relu1(self.bn1(self.conv1(x)))
                x = self.relu2(self.bn2(self.conv2(x)))
                stem = self.relu3(self.bn3(self.conv3(x)))
                c2 = self.layer1(self.avgpool(stem))
                c3 = self.layer2(c2)
                c4 = self.layer3(c3)
                c5 = self.layer4(c4)
                return {"res1": stem, "res2": c2, "res3": c3, "res4": c4, "res5": c5}

        clip_model, clip_tf = clip.load("RN50")
        sd = clip_model.state_dict()
        layers = tuple(len({k.split(".")[2] for k in sd if k.startswith(f"visual.layer{b}")}) for b in (1, 2, 3, 4))
        net = ClipStages(layers, sd["text_projection"].shape[1], sd["visual.layer1.0.conv1.weight"].shape[0] * 32 // 64)
        net.load_state_dict(clip_model.visual.state_dict())
        return net, clip_tf.transforms[-1]
    raise ValueError(f"backbone must be 'resnet' or 'clip', got {kind!r}")


_NORM_CONSTS = {}


def normalize_images(normalize, rgb):
    """torchvision ``Normalize`` (resnet.py:20 / clip.py:34-35 in the reference) with its mean / std kept on the
    device: the stock module rebuilds both from Python lists on every call (two pageable host->device copies, which
    also forbids CUDA-graph capture of a training step).  Same arithmetic: (x - mean) / std."""
    if not hasattr(normalize, "mean"):          # a user-installed transform (e.g. nn.Identity in the test trunk)
        return normalize(rgb)
    key = (id(normalize), rgb.device, rgb.dtype)
    c = _NORM_CONSTS.get(key)
    if c is None:
        mean = torch.as_tensor(normalize.mean, dtype=rgb.dtype).view(-1, 1, 1).to(rgb.device)
        std = torch.as_tensor(normalize.std, dtype=rgb.dtype).view(-1, 1, 1).to(rgb.device)
        c = _NORM_CONSTS[key] = (mean, std)
    return (rgb - c[0]) / c[1]


class EvalTrunk:
    """Inference-time evaluation of (normalize, frozen backbone, FPN) -- still PyTorch / cuDNN library
    calls (the trunk is dense convolution work that stays on cuDNN), but organised the
    way an inference engine would:
      * BatchNorm folded into the preceding convolution (frozen backbone, eval statistics);
      * channels-last activations (NHWC tensor-core kernels, no per-layer layout conversions);
      * for torchvision ResNets, cuDNN's fused conv+bias+ReLU and conv+bias+residual+ReLU entry points
        (`torch.cudnn_convolution_relu`, `torch.cudnn_convolution_add_relu`) instead of separate bias / add /
        clamp kernels;
      * the memory-bound glue around the convolutions as single-pass kernels of this library (input
        normalisation + layout, the stem max-pool, the FPN top-down merge with the lateral bias folded in) and
        the shortcut / output-convolution biases folded or deferred instead of separate elementwise launches;
      * only the FPN levels the model reads are computed (the reference evaluates all five output
        convolutions and discards three of them).
    Only used in eval mode: while training, the reference keeps the frozen backbone's BN in train mode
    (act3d.py:72-73) and that behaviour is preserved by running the plain modules.  The folded weights are
    rebuilt whenever the backbone's parameters / buffers change."""

    def __init__(self):
        self._sig = None
        self._fused = None
        self._plan = None
        self._fused_ok = True
        self.pad_stem = True        # 4-channel (r, g, b, 0) stem input + zero-padded 7x7 weight: vectorised cuDNN kernel

    @staticmethod
    def _signature(backbone):
        return tuple((t.data_ptr(), t._version) for t in list(backbone.parameters()) + list(backbone.buffers()))

    @staticmethod
    def _fold_modules(backbone):
        import copy
        from torch.nn.utils.fusion import fuse_conv_bn_eval
        net = copy.deepcopy(backbone).eval()

        def fold(module):
            names = list(module._modules.keys())
            for a, b in zip(names, names[1:]):
                ma, mb = module._modules[a], module._modules[b]
                if isinstance(ma, torch.nn.Conv2d) and isinstance(mb, torch.nn.BatchNorm2d):
                    module._modules[a] = fuse_conv_bn_eval(ma, mb)
                    module._modules[b] = torch.nn.Identity()
            for child in module._modules.values():
                if child is not None:
                    fold(child)
        fold(net)
        return net.to(memory_format=torch.channels_last)

    @staticmethod
    def _conv_args(conv):
        w = conv.weight.detach().contiguous(memory_format=torch.channels_last)
        b = conv.bias.detach() if conv.bias is not None else torch.zeros(conv.out_channels, device=w.device)
        return w, b, list(conv.stride), list(conv.padding), list(conv.dilation), conv.groups

    def _resnet_plan(self, folded):
        """Flatten a folded torchvision ResNet into (stem, [blocks]) of cuDNN fused-call arguments."""
        stem = self._conv_args(folded.conv1)
        w4 = torch.nn.functional.pad(folded.conv1.weight.detach(), (0, 0, 0, 0, 0, 1)).contiguous(memory_format=torch.channels_last)
        plan = dict(stem=stem, stem4=(w4,) + stem[1:], pool=folded.maxpool, stages=[])
        for stage in (folded.layer1, folded.layer2, folded.layer3, folded.layer4):
            blocks = []
            for blk in stage:
                down = self._conv_args(blk.downsample[0]) if blk.downsample is not None else None
                c3 = self._conv_args(blk.conv3)
                if down is not None:     # relu(conv3 + b3 + (down + bd)) = relu(conv3 + (b3 + bd) + down_nobias)
                    c3 = (c3[0], c3[1] + down[1]) + c3[2:]
                blocks.append((self._conv_args(blk.conv1), self._conv_args(blk.conv2), c3, down))
            plan["stages"].append(blocks)
        return plan

    @staticmethod
    def _run_resnet(plan, x):
        cr, car = torch.cudnn_convolution_relu, torch.cudnn_convolution_add_relu
        stem = cr(x, *(plan["stem4"] if x.shape[1] == 4 else plan["stem"]))
        feats = {"res1": stem}
        pool = plan["pool"]
        if (_pair(pool.kernel_size), _pair(pool.stride), _pair(pool.padding), _pair(pool.dilation), pool.ceil_mode) == \
                ((3, 3), (2, 2), (1, 1), (1, 1), False) and stem.shape[1] % 4 == 0:
            y = lib.trunk_maxpool(stem)               # one NHWC pass (ATen's NHWC pool runs at ~1.2 TB/s here)
        else:
            y = pool(stem)
        for i, blocks in enumerate(plan["stages"]):
            for c1, c2, c3, down in blocks:
                # projection shortcut: its (BN-folded) bias is carried by conv3's bias (plan), so the
                # shortcut convolution runs bias-free instead of paying a separate elementwise add
                ident = y if down is None else torch.nn.functional.conv2d(y, down[0], None, down[2], down[3], down[4], down[5])
                o = cr(y, *c1)
                o = cr(o, *c2)
                y = car(o, c3[0], ident, 1.0, c3[1], c3[2], c3[3], c3[4], c3[5])
            feats[f"res{i + 2}"] = y
        return feats

    @staticmethod
    def _run_fpn(fpn, feats, needed, defer_bias):
        """torchvision FeaturePyramidNetwork.forward restricted to the requested output levels
        (identical arithmetic for those levels).  The top-down merge  (lateral conv + bias) + upsample(top)
        is one kernel (lib.trunk_fpn_topdown) fed by a bias-free lateral convolution; with `defer_bias` the
        bias of the 3x3 output convolutions is not applied to the full map but returned, and the token
        gather adds it to the rows it reads."""
        names = list(feats.keys())
        lowest = min(names.index(n) for n in needed)
        conv = torch.nn.functional.conv2d
        fusable = fpn.inner_blocks[0][0].out_channels % 4 == 0

        def layer(i, t):
            m = fpn.layer_blocks[i][0]
            if defer_bias:
                biases[names[i]] = m.bias
                return conv(t, m.weight, None, padding=1)
            return conv(t, m.weight, m.bias, padding=1)
        out, biases = {}, {}
        top = fpn.inner_blocks[len(names) - 1][0]
        last = conv(feats[names[-1]], top.weight, top.bias)
        if names[-1] in needed:
            out[names[-1]] = layer(len(names) - 1, last)
        for i in range(len(names) - 2, lowest - 1, -1):
            m = fpn.inner_blocks[i][0]
            if fusable:
                last = lib.trunk_fpn_topdown(conv(feats[names[i]], m.weight, None), m.bias, last)
            else:
                lat = conv(feats[names[i]], m.weight, m.bias)
                last = lat + torch.nn.functional.interpolate(last, size=lat.shape[-2:], mode="nearest")
            if names[i] in needed:
                out[names[i]] = layer(i, last)
        return (out, biases) if defer_bias else out

    def __call__(self, normalize, backbone, fpn, rgb, needed=("res1", "res3"), defer_bias=False):
        """Returns {level: map}; with defer_bias=True returns ({level: map without the output-conv bias},
        {level: bias}) for consumers that add the bias themselves (lib.gather_tokens)."""
        sig = self._signature(backbone)
        if self._fused is None or sig != self._sig:
            self._fused, self._sig = self._fold_modules(backbone), sig
            self._plan = self._resnet_plan(self._fused) if isinstance(backbone, ResNet) else None
        if rgb.is_cuda and isinstance(normalize, transforms.Normalize) and rgb.shape[1] == 3 and len(normalize.mean) == 3:
            pad = self.pad_stem and self._plan is not None and self._fused_ok
            x = lib.trunk_normalize(rgb.float().contiguous(), normalize.mean, normalize.std, 4 if pad else 3)
        else:
            x = normalize(rgb).contiguous(memory_format=torch.channels_last)
        if self._plan is not None and self._fused_ok and x.is_cuda:
            try:
                return self._run_fpn(fpn, self._run_resnet(self._plan, x), set(needed), defer_bias)
            except torch.cuda.OutOfMemoryError:
                raise                     # transient: must not demote the trunk to the slow path for the process lifetime
            except RuntimeError as exc:   # cuDNN fused entry points unavailable for this build / shape
                msg = str(exc)
                if any(k in msg for k in ("libact3d_b200", "a3d_", "CUDA error", "out of memory", "device-side assert")):
                    raise                 # our own kernels / device faults are never swallowed
                import warnings
                warnings.warn("fused cuDNN conv+bias+ReLU path unavailable, using the unfused trunk from now on: " + msg)
                self._fused_ok = False
        if x.shape[1] == 4:
            x = x[:, :3].contiguous(memory_format=torch.channels_last)
        out = fpn(self._fused(x))
        return (out, {}) if defer_bias else out
