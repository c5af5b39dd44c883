"""PyTorch feature extractor kept from the reference design (the north star leaves it in PyTorch/cuDNN):
ResNet trunk to layer3 and the multi-kernel stride-2 PatchMerging neck.  Parameter names match the
reference so its checkpoints load strictly (reference src/models/backbone.py:18-67,130-174)."""
import torch
import torch.nn as nn
import torchvision.models as tvm

_RESNETS = {18: tvm.resnet18, 34: tvm.resnet34, 50: tvm.resnet50, 101: tvm.resnet101, 152: tvm.resnet152}


class PatchMerging(nn.Module):
    """LayerNorm over channels, then parallel stride-2 convolutions with kernels `patch_size`, concatenated
    (backbone.py:28-67).  Output channels: 2*dim."""

    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm, patch_size=(2,)):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.patch_size = list(patch_size)
        self.norm = norm_layer(dim)
        self.reductions = nn.ModuleList()
        last = len(self.patch_size) - 1
        for i, ps in enumerate(self.patch_size):
            out_dim = (2 * dim) // (2 ** i if i == last else 2 ** (i + 1))
            self.reductions.append(nn.Conv2d(dim, out_dim, kernel_size=ps, stride=2, padding=(ps - 2) // 2))

    def forward(self, x):
        x = self.norm(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2).contiguous()
        return torch.cat([conv(x) for conv in self.reductions], dim=1)


class ResnetEncoder(nn.Module):
    """torchvision ResNet to `cfg.BACKBONE.LAYER`; takes NHWC images in [0,1] (backbone.py:159-174).
    `pretrained` is off by default: this build has no network access; real weights arrive via load_state_dict."""

    def __init__(self, cfg, pretrained=False):
        super().__init__()
        self.cfg = cfg
        self.last_layer = cfg.BACKBONE.LAST_LAYER
        ctor = _RESNETS[cfg.BACKBONE.NUM_LAYERS]
        net = ctor(weights="DEFAULT") if pretrained else ctor(weights=None)
        self.encoder = net           # kept as an attribute: reference checkpoints carry `backbone.encoder.*` keys
        self.layer0 = nn.Sequential(net.conv1, net.bn1, net.relu)
        self.layer1 = nn.Sequential(net.maxpool, net.layer1)
        self.layer2 = net.layer2
        self.layer3 = net.layer3
        if cfg.BACKBONE.LAYER == "layer4":
            self.layer4 = net.layer4

    def forward(self, image_nhwc):
        x = image_nhwc.permute(0, 3, 1, 2).contiguous()
        if self.cfg.NORM_INPUT:
            x = (x - 0.45) / 0.225
        x = self.layer3(self.layer2(self.layer1(self.layer0(x))))
        if self.cfg.BACKBONE.LAYER == "layer4":
            x = self.layer4(x)
        return x
