"""PyTorch feature extractor kept from the reference design (the north star leaves it in PyTorch/cuDNN):
ResNet trunk to layer3 and the multi-kernel stride-2 PatchMerging neck.  Parameter names match the
reference so its checkpoints load strictly (reference src/models/backbone.py:18-67,130-174).

Execution modes of the trunk (tuning only, the arithmetic library stays cuDNN):
  "eager"          the reference's execution: NCHW fp32 (the default; bit-compatible with the reference on the same device)
  "channels_last"  NHWC fp32 (no TF32): same arithmetic, tensor-core-friendly layout
  "tf32"           channels_last with TF32 convolutions
  "bf16"           channels_last under bf16 autocast (features returned as fp32)
`graphs=True` additionally captures the trunk into one CUDA graph per input shape and replays it (static shapes,
inference only).  Reduced-precision modes change the features by 1e-3 .. 1e-2 relative; tools/backbone_bench.py measures
the speed and the box error each mode causes through the neck and the hot path."""
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

    MODES = ("eager", "channels_last", "tf32", "bf16")

    def set_execution_mode(self, mode="eager", graphs=False):
        """See the module docstring.  Returns self."""
        if mode not in self.MODES:
            raise ValueError("execution mode %r not in %s" % (mode, self.MODES))
        self._mode, self._graphs, self._graph_cache = mode, bool(graphs), {}
        fmt = torch.contiguous_format if mode == "eager" else torch.channels_last
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data = m.weight.data.contiguous(memory_format=fmt)
        return self

    def _trunk(self, image_nhwc):
        mode = getattr(self, "_mode", "eager")
        x = image_nhwc.permute(0, 3, 1, 2)
        x = x.contiguous() if mode == "eager" else x.contiguous(memory_format=torch.channels_last)   # NHWC input: a view
        if self.cfg.NORM_INPUT:
            x = (x - 0.45) / 0.225
        x = self.layer3(self.layer2(self.layer1(self.layer0(x))))
        if self.cfg.BACKBONE.LAYER == "layer4":
            x = self.layer4(x)
        return x

    def _run(self, image_nhwc):
        mode = getattr(self, "_mode", "eager")
        if mode in ("eager", "channels_last") or not image_nhwc.is_cuda:
            return self._trunk(image_nhwc).contiguous()
        if mode == "bf16":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self._trunk(image_nhwc).float().contiguous()
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        try:
            return self._trunk(image_nhwc).contiguous()
        finally:
            torch.backends.cudnn.allow_tf32 = prev

    def forward(self, image_nhwc):
        if not getattr(self, "_graphs", False) or not image_nhwc.is_cuda or torch.is_grad_enabled():
            return self._run(image_nhwc)
        key = (tuple(image_nhwc.shape), image_nhwc.dtype, image_nhwc.device)
        entry = self._graph_cache.get(key)
        if entry is None:
            static_in = torch.empty_like(image_nhwc)
            static_in.copy_(image_nhwc)
            side = torch.cuda.Stream(device=image_nhwc.device)
            side.wait_stream(torch.cuda.current_stream(image_nhwc.device))
            with torch.cuda.stream(side):
                for _ in range(2):                            # warm-up (cuDNN plan selection) outside the capture
                    self._run(static_in)
            torch.cuda.current_stream(image_nhwc.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._run(static_in)
            entry = self._graph_cache[key] = (graph, static_in, static_out)
        graph, static_in, static_out = entry
        static_in.copy_(image_nhwc)
        graph.replay()
        return static_out.clone()
