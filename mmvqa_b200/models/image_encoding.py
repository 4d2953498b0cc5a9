"""Visual-token projector.  Mirrors models/image_encoding.py of the reference (Transfer,
ResNetTransfer, Timm_EFfNetV2, get_transfer, get_image_encoder, models_dict).

The CNN backbone is library code (torchvision / timm) and out of scope; what is rebuilt is the
``gap(act(conv1x1(feature_map)))`` stage: one GEMM over pixels with the activation and the spatial mean
fused into its epilogue, so the [B, hidden, H, W] map (616 MB at 112x112, B=16) is never written."""
import torch
import torch.nn as nn
from torchvision import models

from .. import functional as Fn
from .._lib import ACT_RELU, ACT_SERF
from .serf import SERF

try:  # timm is what the reference uses for EfficientNetV2; it is optional here
    import timm as _timm
    _timm_create = _timm.create_model
except Exception:  # pragma: no cover - depends on the environment
    _timm = None

    def _timm_create(name, features_only=True, pretrained=True, **kw):
        """timm is absent.  A randomly initialised torchvision efficientnet_v2_m (taps features[1,2,3,5,7] = 24/48/80/
        176/512 channels at strides 2/4/8/16/32) has the right SHAPES but neither the pretrained
        weights nor timm's state-dict keys, so it is only handed out when the caller opted in
        (MMVQA_ALLOW_TORCHVISION_EFFNET=1: tests, synthetic benchmarks); otherwise this raises like the reference's
        `import timm` would."""
        import os
        if os.environ.get("MMVQA_ALLOW_TORCHVISION_EFFNET") != "1":
            raise ImportError("timm is not installed: get_image_encoder('%s') cannot build the pretrained backbone. "
                              "Set MMVQA_ALLOW_TORCHVISION_EFFNET=1 to get a randomly initialised, shape-identical "
                              "torchvision stand-in (tests / synthetic benchmarks only)." % name)
        import warnings
        warnings.warn("timm missing: using a RANDOMLY INITIALISED torchvision efficientnet_v2_m stand-in "
                      "(state-dict keys differ from timm's; reference checkpoints will not load into the backbone)")
        return TorchvisionEffNetV2Features()


class TorchvisionEffNetV2Features(nn.Module):
    taps = (1, 2, 3, 5, 7)

    def __init__(self):
        super().__init__()
        self.features = models.efficientnet_v2_m(weights=None).features[:8]

    def forward(self, x):
        outs = []
        for i, m in enumerate(self.features):
            x = m(x)
            if i in self.taps:
                outs.append(x)
        return outs


# first key: num_vis, second key: image encoder name -> [constructor, channel sizes in token order]
models_dict = {5: {'resnet152': [models.resnet152, [2048, 1024, 512, 256, 64]],
                   'tf_efficientnetv2_m': [_timm_create, [24, 48, 80, 176, 512]]},
               7: {'tf_efficientnetv2_m': [_timm_create, [24, 48, 80, 160, 176, 304, 512]]}}


def get_image_encoder(args):
    m, channel_size = models_dict[args.num_vis][args.cnn_encoder]
    if 'resnet' in args.cnn_encoder:
        return m(pretrained=True), channel_size
    elif 'efficientnetv2' in args.cnn_encoder:
        return m(args.cnn_encoder, features_only=True, pretrained=True), channel_size


def get_transfer(args):
    if 'resnet' in args.cnn_encoder:
        return ResNetTransfer(args)
    elif 'efficientnetv2' in args.cnn_encoder:
        if args.num_vis == 5:
            return Timm_EFfNetV2(args)
        raise NotImplementedError("num_vis=7 (EffNetV2Transfer7Tokens) is broken in the reference and not a target")
    else:
        raise NotImplementedError


class Transfer(nn.Module):
    def __init__(self, args):
        super(Transfer, self).__init__()
        self.args = args
        self.model, self.channel_size = get_image_encoder(args)
        self.serf = SERF()
        hs = args.hidden_size
        self.conv2 = nn.Conv2d(self.channel_size[0], hs, kernel_size=(1, 1), stride=(1, 1), bias=False)
        self.gap2 = nn.AdaptiveAvgPool2d((1, 1))
        self.conv3 = nn.Conv2d(self.channel_size[1], hs, kernel_size=(1, 1), stride=(1, 1), bias=False)
        self.gap3 = nn.AdaptiveAvgPool2d((1, 1))
        self.conv4 = nn.Conv2d(self.channel_size[2], hs, kernel_size=(1, 1), stride=(1, 1), bias=False)
        self.gap4 = nn.AdaptiveAvgPool2d((1, 1))
        self.conv5 = nn.Conv2d(self.channel_size[3], hs, kernel_size=(1, 1), stride=(1, 1), bias=False)
        self.gap5 = nn.AdaptiveAvgPool2d((1, 1))
        self.conv7 = nn.Conv2d(self.channel_size[4], hs, kernel_size=(1, 1), stride=(1, 1), bias=False)
        self.gap7 = nn.AdaptiveAvgPool2d((1, 1))
        self.relu = nn.ReLU()
        self.activation = self.relu if args.use_relu else self.serf

    def _convs(self):
        return (self.conv2, self.conv3, self.conv4, self.conv5, self.conv7)

    def project_stacked(self, feats):
        """feature maps (token order) -> [num_vis, B, hidden] fp32 visual tokens (one autograd node, levels overlapped)."""
        act = ACT_RELU if self.args.use_relu else ACT_SERF
        feats = list(feats)
        return Fn.vistok_project_all(feats, [c.weight for c in self._convs()][:len(feats)], act)

    def project(self, feats):
        """feature maps (token order) -> tuple of [B, hidden] fp32 visual tokens."""
        return tuple(self.project_stacked(feats).unbind(0))


class ResNetTransfer(Transfer):
    """image_encoding.py:64-87.  The reference re-runs five prefixes of the backbone; the taps are the
    same tensors, so the backbone runs ONCE here and the five maps are tapped on the way (deep->shallow
    token order is preserved)."""

    @staticmethod
    def tap_feature_maps(backbone, img):
        """One pass over `backbone.children()[:-2]`, returning the outputs of the prefixes [:-2], [:-3], [:-4], [:-5],
        [:-7] in that (deep -> shallow) order -- the five tensors image_encoding.py:72-85 obtains by re-running each
        prefix from the image (tests/test_host_cpu.py::test_resnet_single_pass_taps_equal_the_five_prefix_runs)."""
        ch = list(backbone.children())
        n = len(ch)
        taps = {n - 2: 0, n - 3: 1, n - 4: 2, n - 5: 3, n - 7: 4}     # prefix length -> token index
        feats = [None] * 5
        x = img
        for i, m in enumerate(ch[:n - 2]):
            x = m(x)
            if (i + 1) in taps:
                feats[taps[i + 1]] = x
        return feats

    def forward(self, img):
        return self.project(self.tap_feature_maps(self.model, img))


class Timm_EFfNetV2(Transfer):
    """image_encoding.py:89-128 (incl. the Grad-CAM hooks on the deepest map)."""

    def __init__(self, args):
        super().__init__(args)
        self.grad_cam = args.grad_cam if hasattr(args, 'grad_cam') else False

    def forward(self, img):
        o = self.model(img)
        if self.grad_cam:
            o[4].register_hook(self.activations_hook)
            self.feat = o[4]
        return self.project(o)

    def activations_hook(self, grad):
        self.gradients = grad

    def get_activations_gradient(self):
        return self.gradients

    def get_activations(self):
        return self.feat
