"""Domain-adaptation pieces inside DINO.forward.

Mirrors the reference's models/dino/DA_utils.py: decompose_features (:5-31), GradReverse/grad_reverse
(:33-43), FCDiscriminator_img (:61-79), get_prototype_class_wise (:82-123).
"""
import torch
import torch.nn.functional as F
from torch import nn


def decompose_features(srcs, masks, poss):
    """Split every level's batch into its source (first) and target (second) half.
    Returns (src_s, mask_s, pos_s, src_all, mask_all, pos_all, src_t, mask_t, pos_t)."""
    half = srcs[0].shape[0] // 2
    first = lambda xs: [x[:half] for x in xs]
    second = lambda xs: [x[half:] for x in xs]
    return (first(srcs), first(masks), first(poss), srcs, masks, poss, second(srcs), second(masks), second(poss))


class GradReverse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.neg()


def grad_reverse(x):
    return GradReverse.apply(x)


class DA_MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i + 1 < self.num_layers:
                x = F.relu(x)
        return x


class FCDiscriminator_img(nn.Module):
    """Per-pixel domain classifier: 3x3 convs C->256->128->128->1 with LeakyReLU(0.2)."""

    def __init__(self, num_classes, ndf1=256, ndf2=128):
        super().__init__()
        self.conv1 = nn.Conv2d(num_classes, ndf1, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(ndf1, ndf2, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(ndf2, ndf2, kernel_size=3, padding=1)
        self.classifier = nn.Conv2d(ndf2, 1, kernel_size=3, padding=1)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        from datr_b200 import conv as dconv, linear as dl
        for conv in (self.conv1, self.conv2, self.conv3):
            if (dl.get_mode() == "tf32" and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
                    and dconv.use_kernel(x, conv)):
                # conv + bias + LeakyReLU(0.2) in one implicit-GEMM tcgen05 kernel; its input gradient runs on the
                # same kernel (datr_b200.conv)
                x = dconv.conv3x3_bias_act(x, conv.weight, conv.bias, 1, 2)
            else:
                x = self.leaky_relu(conv(x))
        return self.classifier(x)       # 128 -> 1 channel: outside the kernel's shapes (Cout % 4), library convolution


def get_prototype_class_wise(object_query_last_layer, outputs_class, num_classes, global_proto=None, global_amount=None):
    """Class prototypes = mean of the last-layer query features over the queries whose arg-max class is c.

    The reference materialises a [B*N, num_classes, C] masked copy (DA_utils.py:96-108); the same sums are
    one [num_classes, B*N] x [B*N, C] product here.  Returns (prototypes [K,C], class-present map [K],
    updated global prototypes [K,C] (detached running mean weighted by counts), updated counts [K],
    one-hot assignment [B*N,K])."""
    B, N, C = object_query_last_layer.shape
    feats = object_query_last_layer.reshape(B * N, C)
    label = outputs_class.sigmoid().argmax(dim=2).reshape(B * N)
    onehot = F.one_hot(label, num_classes).to(feats.dtype)
    count = onehot.sum(0)
    present = (count != 0).to(count.dtype)
    proto = (onehot.t() @ feats) / count.clamp(min=1).unsqueeze(1)
    w = count / (count + global_amount)
    w = torch.where(count == 0, torch.zeros_like(w), w).unsqueeze(1)
    global_proto = (global_proto * (1 - w) + proto * w).detach()
    return proto, present, global_proto, global_amount + count, onehot
