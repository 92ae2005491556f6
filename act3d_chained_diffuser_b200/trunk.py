"""Frozen image backbones + FPN (stay PyTorch/cuDNN: dense convolutions).

``backbone="resnet"`` is torchvision's ResNet-50 returning its five stage outputs
(reference: model/utils/resnet.py:35-56); ``backbone="clip"`` needs the third-party ``clip``
package and its downloaded RN50 weights (reference: model/utils/clip.py:9-43), exactly like
the reference -- it raises ImportError here when that package is absent.
"""
import torch
from torchvision import transforms
from torchvision.models.resnet import Bottleneck, ResNet


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


class EvalTrunk:
    """Inference-time view of (normalize, frozen backbone, FPN): BatchNorm folded into the preceding
    convolution and channels-last activations, so that cuDNN runs NHWC tensor-core kernels without
    the per-layer NCHW<->NHWC conversions and separate BN / ReLU passes of the eager module.
    Only used in eval mode (the reference keeps the frozen backbone's BN in train mode while
    training, act3d.py:72-73 -- that behaviour is preserved by falling back to the plain modules).
    The folded copy is rebuilt when the source parameters / buffers change."""

    def __init__(self):
        self._sig = None
        self._fused = None

    @staticmethod
    def _signature(backbone):
        return tuple((t.data_ptr(), t._version) for t in list(backbone.parameters()) + list(backbone.buffers()))

    def _fuse(self, backbone):
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

    def __call__(self, normalize, backbone, fpn, rgb):
        sig = self._signature(backbone)
        if self._fused is None or sig != self._sig:
            self._fused, self._sig = self._fuse(backbone), sig
        x = normalize(rgb).contiguous(memory_format=torch.channels_last)
        return fpn(self._fused(x))
