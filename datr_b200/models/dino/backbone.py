"""ResNet backbone with frozen batch-norm, mask down-sampling and position encoding.

Mirrors the hot-path part of the reference's models/dino/backbone.py: FrozenBatchNorm2d (:36-72),
BackboneBase (:75-106; stem + layer1 frozen :79-81, nearest-neighbour mask resize :103),
Backbone (:109-129; torchvision ResNet-50/101 layout, stride on the 3x3 conv), Joiner (:132-144),
build_backbone (:147-219).  Swin / ConvNeXt backbones of the reference are outside the hot path.

The ResNet is defined here (not taken from torchvision) with torchvision's parameter names, so the
reference's checkpoints (`backbone.0.body.layerK.B.convJ.weight`, `...bnJ.{weight,bias,running_mean,
running_var}`, `...downsample.{0,1}`) load unchanged, and so that the frozen batch-norm folds into a
per-channel scale/shift applied in one pass (the reference does 4 elementwise passes, :62-72).
"""
from typing import Dict, List

import torch
import torch.nn.functional as F
from torch import nn

from datr_b200 import conv as dconv
from datr_b200 import linear as dl

from datr_b200.util.misc import NestedTensor, is_main_process
from .position_encoding import build_position_encoding


class FrozenBatchNorm2d(nn.Module):
    """y = x * w / sqrt(var + 1e-5) + (b - mean * w / sqrt(var + 1e-5)); statistics and affine are buffers."""

    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def _load_from_state_dict(self, state_dict, prefix, *rest):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, *rest)

    def scale_shift(self):
        folded = self.__dict__.get("_folded")
        if folded is not None:          # computed for all layers of the body at once (fold_frozen_bn), valid for this pass
            return folded
        scale = self.weight * (self.running_var + 1e-5).rsqrt()
        return scale, self.bias - self.running_mean * scale

    def forward(self, x):
        scale, shift = self.scale_shift()
        return torch.addcmul(shift.view(1, -1, 1, 1), x, scale.view(1, -1, 1, 1))


class fold_frozen_bn:
    """Context: (scale, shift) of every FrozenBatchNorm2d below `module` with five multi-tensor launches instead of five
    tiny kernels per layer (265 launches for ResNet-50), same operations in the same order; recomputed at every pass, so
    buffers changed by a checkpoint load or an EMA copy are always seen, also by a CUDA graph captured over the pass."""

    def __init__(self, module):
        self.bns = [m for m in module.modules() if isinstance(m, FrozenBatchNorm2d)]

    def __enter__(self):
        bns = self.bns
        if bns and bns[0].weight.is_cuda:
            inv = torch._foreach_add([b.running_var for b in bns], 1e-5)
            torch._foreach_rsqrt_(inv)
            scale = torch._foreach_mul([b.weight for b in bns], inv)
            shift = torch._foreach_sub([b.bias for b in bns], torch._foreach_mul([b.running_mean for b in bns], scale))
            for b, sc, sh in zip(bns, scale, shift):
                b.__dict__["_folded"] = (sc, sh)
        return self

    def __exit__(self, *exc):
        for b in self.bns:
            b.__dict__.pop("_folded", None)
        return False


def _pointwise_on_tensor_cores(x):
    return (dl.get_mode() == "tf32" and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
            and x.is_contiguous(memory_format=torch.channels_last))


# Backward-pass fusion inside a ResNet stage (DATR_RESNET_FUSED_BWD=0: autograd's separate passes).  Between two
# bottlenecks WITHOUT a downsample branch the gradient of a block output y = relu(bn3(conv3(.)) + x) has exactly two
# consumers, the next block's conv1 and its skip connection; autograd adds the two gradients (one pass over the map) and
# the block then applies its ReLU mask (another pass).  Here the next block's conv3 hands its skip gradient to its conv1
# (linear.GradCarrier), whose input-gradient GEMM adds it and applies (x > 0) in the epilogue
# (datr_linear_tf32_bt_masked); the producing block is told its gradient arrives masked.
import os as _os
_FUSED_BWD = _os.environ.get("DATR_RESNET_FUSED_BWD", "1") != "0"


def _conv1x1_bn(x, conv, bn, relu, residual=None, **fuse):
    """FrozenBN(conv1x1(x)) [+ residual] with optional ReLU on an NHWC tensor, as one fused GEMM over pixels."""
    assert conv.kernel_size == (1, 1) and conv.padding == (0, 0) and conv.groups == 1 and conv.bias is None
    if conv.stride != (1, 1):
        x = x[:, :, ::conv.stride[0], ::conv.stride[1]].contiguous(memory_format=torch.channels_last)
    n, cin, h, w = x.shape
    cout = conv.out_channels
    scale, shift = bn.scale_shift()
    weight = conv.weight.reshape(cout, cin) * scale[:, None]
    res = residual.permute(0, 2, 3, 1).reshape(-1, cout) if residual is not None else None
    y = dl.linear(x.permute(0, 2, 3, 1).reshape(-1, cin), weight, shift, relu=relu, residual=res, **fuse)
    return y.view(n, h, w, cout).permute(0, 3, 1, 2)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1, norm_layer=FrozenBatchNorm2d):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = norm_layer(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = norm_layer(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = norm_layer(planes * 4)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x, fuse_in=False, fuse_out=False):
        """fuse_in: x is the output of the previous bottleneck of the stage, consumed by this block only, and that block
        was called with fuse_out -- the gradient returned for x carries the previous block's ReLU mask.  Both flags are
        set by run_stage() and only on the tensor-core path."""
        if _pointwise_on_tensor_cores(x):
            # NHWC activations: the three 1x1 convolutions are GEMMs over pixels; FrozenBN folds into weight / bias
            # and ReLU / the residual add ride in the tcgen05 kernel's epilogue (datr_b200.linear); 3x3 stays cuDNN
            fuse_in = fuse_in and self.downsample is None and x.requires_grad and torch.is_grad_enabled()
            carrier = dl.GradCarrier() if fuse_in else None
            first = dict(skip_in=carrier, mask_input_grad=True) if fuse_in else {}
            last = dict(skip_out=carrier) if fuse_in else {}
            if fuse_out:
                last["grad_premasked"] = True
            y = _conv1x1_bn(x, self.conv1, self.bn1, relu=1, **first)
            if dconv.use_kernel(y, self.conv2):
                # conv2 + FrozenBN + ReLU: im2col-free implicit GEMM (4-D TMA box per tap, datr_b200.conv)
                scale, shift = self.bn2.scale_shift()
                y = dconv.conv3x3_bias_relu(y, self.conv2.weight * scale.view(-1, 1, 1, 1), shift, self.conv2.stride[0])
            else:
                y = F.relu(self.bn2(self.conv2(y)))
            if self.downsample is not None:
                x = _conv1x1_bn(x, self.downsample[0], self.downsample[1], relu=0)
            return _conv1x1_bn(y, self.conv3, self.bn3, relu=2, residual=x, **last)
        assert not (fuse_in or fuse_out), "backward-pass fusion is a property of the tensor-core path"
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        if self.downsample is not None:
            x = self.downsample(x)
        return F.relu(y + x)


class ResNet(nn.Module):
    """conv1/bn1/maxpool + layer1..layer4 of bottlenecks (ResNet-50: 3,4,6,3; ResNet-101: 3,4,23,3)."""

    def __init__(self, blocks, replace_stride_with_dilation=(False, False, False), norm_layer=FrozenBatchNorm2d):
        super().__init__()
        self.inplanes, self.dilation = 64, 1
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = norm_layer(64)
        self.layer1 = self._stage(64, blocks[0], 1, False, norm_layer)
        self.layer2 = self._stage(128, blocks[1], 2, replace_stride_with_dilation[0], norm_layer)
        self.layer3 = self._stage(256, blocks[2], 2, replace_stride_with_dilation[1], norm_layer)
        self.layer4 = self._stage(512, blocks[3], 2, replace_stride_with_dilation[2], norm_layer)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _stage(self, planes, n, stride, dilate, norm_layer):
        prev_dilation = self.dilation
        if dilate:
            self.dilation *= stride
            stride = 1
        down = None
        if stride != 1 or self.inplanes != planes * 4:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride=stride, bias=False),
                                 norm_layer(planes * 4))
        layers = [Bottleneck(self.inplanes, planes, stride, down, prev_dilation, norm_layer)]
        self.inplanes = planes * 4
        layers += [Bottleneck(self.inplanes, planes, dilation=self.dilation, norm_layer=norm_layer) for _ in range(1, n)]
        return nn.Sequential(*layers)

    def stem(self, x):
        return F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, stride=2, padding=1)


def run_stage(stage, x):
    """One ResNet stage (nn.Sequential of bottlenecks).  On the tensor-core path with gradients, consecutive blocks fuse
    the accumulation of the skip gradient and the ReLU backward between them into the next block's input-gradient GEMM
    (see _FUSED_BWD); otherwise this is stage(x)."""
    blocks = list(stage)
    trains = any(p.requires_grad for p in stage.parameters())
    if not (_FUSED_BWD and trains and torch.is_grad_enabled() and _pointwise_on_tensor_cores(x)
            and all(isinstance(b, Bottleneck) for b in blocks)):
        return stage(x)
    for j, blk in enumerate(blocks):
        # block j hands a masked gradient to block j-1 iff block j has no downsample branch (x feeds conv1 + skip only)
        fuse_in = j > 0 and blk.downsample is None
        fuse_out = j + 1 < len(blocks) and blocks[j + 1].downsample is None
        x = blk(x, fuse_in=fuse_in, fuse_out=fuse_out)
    return x


_RESNETS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}


class _Body(nn.Module):
    """Runs the ResNet and returns {"<interm index>": feature} for the requested stages (the role of
    torchvision's IntermediateLayerGetter in the reference); children keep the ResNet's names."""

    def __init__(self, net: ResNet, return_layers: Dict[str, str]):
        super().__init__()
        last = max(int(k[-1]) for k in return_layers)
        self.conv1, self.bn1 = net.conv1, net.bn1
        for i in range(1, last + 1):
            setattr(self, f"layer{i}", getattr(net, f"layer{i}"))
        self.return_layers = dict(return_layers)
        self._last = last

    def _stem(self, x):
        """conv1 -> bn1 -> ReLU -> MaxPool2d(3, 2, 1).  On NHWC CUDA tensors without gradient (the stem is frozen in every
        DATR configuration) the last three run as one kernel (datr_bn_relu_maxpool_nhwc)."""
        y = self.conv1(x)
        if (y.is_cuda and y.dtype == torch.float32 and not y.requires_grad and isinstance(self.bn1, FrozenBatchNorm2d)
                and y.is_contiguous(memory_format=torch.channels_last) and y.shape[1] % 4 == 0):
            from datr_b200 import native
            n, c, h, w = y.shape
            scale, shift = self.bn1.scale_shift()
            out = torch.empty((n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1), dtype=torch.float32, device=y.device,
                              memory_format=torch.channels_last)
            lib = native.lib()
            with torch.cuda.device(y.device):
                rc = lib.datr_bn_relu_maxpool_nhwc(y.data_ptr(), scale.contiguous().data_ptr(), shift.contiguous().data_ptr(),
                                                   n, h, w, c, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            if rc != 0:
                raise RuntimeError(f"datr_bn_relu_maxpool_nhwc failed (code {rc}): {lib.datr_decoder_ops_last_error().decode()}")
            return out
        return F.max_pool2d(F.relu(self.bn1(y)), 3, stride=2, padding=1)

    def forward(self, x):
        out = {}
        with fold_frozen_bn(self):
            x = self._stem(x)
            for i in range(1, self._last + 1):
                name = f"layer{i}"
                # stem and layer1 never train (BackboneBase): no autograd graph through them
                x = run_stage(getattr(self, name), x)
                if name in self.return_layers:
                    out[self.return_layers[name]] = x
        return out


class BackboneBase(nn.Module):
    def __init__(self, backbone: nn.Module, train_backbone: bool, num_channels, return_interm_indices: list):
        super().__init__()
        for name, p in backbone.named_parameters():
            if not train_backbone or not any(k in name for k in ("layer2", "layer3", "layer4")):
                p.requires_grad_(False)
        n = len(return_interm_indices)
        return_layers = {f"layer{5 - n + i}": str(idx) for i, idx in enumerate(return_interm_indices)}
        self.body = _Body(backbone, return_layers)
        self.num_channels = num_channels

    def forward(self, tensor_list: NestedTensor):
        feats = self.body(tensor_list.tensors)
        m = tensor_list.mask
        assert m is not None
        out: Dict[str, NestedTensor] = {}
        for name, x in feats.items():
            mask = F.interpolate(m[None].float(), size=x.shape[-2:]).to(torch.bool)[0]
            out[name] = NestedTensor(x, mask)
        return out


_TV_FILES = {"resnet50": "resnet50-0676ba61.pth", "resnet101": "resnet101-63fe2227.pth"}   # torchvision IMAGENET1K_V1


def _pretrained_state_dict(name: str):
    """ImageNet weights for the ResNet body without touching the network: a path given through
    DATR_BACKBONE_WEIGHTS, else the file torchvision's `resnet50(pretrained=True)` (reference backbone.py:118-120)
    would have downloaded into the torch hub cache.  None if neither exists."""
    import os
    cands = [os.environ.get("DATR_BACKBONE_WEIGHTS")]
    try:
        cands.append(os.path.join(torch.hub.get_dir(), "checkpoints", _TV_FILES.get(name, "")))
    except Exception:
        pass
    for path in cands:
        if path and os.path.isfile(path):
            sd = torch.load(path, map_location="cpu", weights_only=True)
            return sd.get("state_dict", sd.get("model", sd)) if isinstance(sd, dict) else sd
    return None


class Backbone(BackboneBase):
    """ResNet-50/101 with FrozenBatchNorm2d.  The reference builds `torchvision.models.resnet50(pretrained=
    is_main_process())` (:118-120), i.e. downloads ImageNet weights.  Parameter names here are torchvision's, so the
    same file is loaded when it is already on disk (DATR_BACKBONE_WEIGHTS or the torch hub cache); nothing is ever
    downloaded.  Without it the body is randomly initialised -- with a FROZEN stem / layer1 and identity FrozenBN
    statistics that is only meaningful when a checkpoint is loaded afterwards (resume / finetune / the synthetic
    benchmark), so a warning says so (DATR_BACKBONE_WEIGHTS=none silences it)."""

    def __init__(self, name: str, train_backbone: bool, dilation: bool, return_interm_indices: list,
                 batch_norm=FrozenBatchNorm2d):
        if name not in _RESNETS:
            raise NotImplementedError(f"Why you can get here with name {name}")
        assert return_interm_indices in [[0, 1, 2, 3], [1, 2, 3], [3]]
        net = ResNet(_RESNETS[name], (False, False, dilation), norm_layer=batch_norm)
        import os
        if os.environ.get("DATR_BACKBONE_WEIGHTS", "").lower() != "none":
            sd = _pretrained_state_dict(name) if is_main_process() else None
            if sd is not None:
                missing, unexpected = net.load_state_dict({k: v for k, v in sd.items() if not k.startswith("fc.")}, strict=False)
                assert not [k for k in missing if "num_batches_tracked" not in k], f"backbone weights incomplete: {missing[:5]}"
            elif is_main_process():
                import warnings
                warnings.warn("datr_b200 backbone: no ImageNet weights found (DATR_BACKBONE_WEIGHTS or torch hub cache); the "
                              "ResNet body is RANDOMLY initialised with a frozen stem/layer1 and identity FrozenBatchNorm -- "
                              "load a checkpoint (--pretrain_model_path / --resume) before training, or set "
                              "DATR_BACKBONE_WEIGHTS=<file.pth>.", stacklevel=2)
        num_channels = [256, 512, 1024, 2048][4 - len(return_interm_indices):]
        super().__init__(net, train_backbone, num_channels, return_interm_indices)


class Joiner(nn.Sequential):
    def __init__(self, backbone, position_embedding):
        super().__init__(backbone, position_embedding)

    def forward(self, tensor_list: NestedTensor):
        feats = self[0](tensor_list)
        out: List[NestedTensor] = list(feats.values())
        pos = [self[1](x).to(x.tensors.dtype) for x in out]
        return out, pos


def build_backbone(args):
    position_embedding = build_position_encoding(args)
    if not args.lr_backbone > 0:
        raise ValueError("Please set lr_backbone > 0")
    idx = args.return_interm_indices
    assert idx in [[0, 1, 2, 3], [1, 2, 3], [3]]
    if args.backbone not in _RESNETS:
        raise NotImplementedError(f"Unknown backbone {args.backbone} (the B200 hot path covers ResNet-50/101)")
    backbone = Backbone(args.backbone, True, args.dilation, idx, batch_norm=FrozenBatchNorm2d)
    model = Joiner(backbone, position_embedding)
    model.num_channels = backbone.num_channels
    return model
