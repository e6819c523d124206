"""Post-raster feature head: the reference's

    feature_map = F.interpolate(feature_map.unsqueeze(0), size=gt.shape[1:], mode='bilinear', align_corners=True).squeeze(0)
    [feature_map = cnn_decoder(feature_map)]          # 1x1 conv, only with --speedup (models/networks.py:107-119)
    Ll1_feature = l1_loss(feature_map, gt)            # train.py:98-104

on two CUDA kernels of libf3dgs_b200 (csrc/feature_head.cu): a resize that, given the teacher map, directly emits the loss
and dL/d(resized map), and a gather-style resize backward that writes every element of dL/dfeature_map exactly once.
There is no CPU path (the extension raises on CPU tensors).
"""
import torch


def _C():
    from . import _C as ext  # deferred: keeps this module importable for documentation tools without the extension

    return ext


def feature_l1_loss_and_grad(feature_map: torch.Tensor, gt: torch.Tensor, weight: float = 1.0):
    """-> (loss, dL/dfeature_map) for loss = weight * mean|resize(feature_map) - gt|, no autograd graph (ViewBatch loops)."""
    C, Hg, Wg = gt.shape
    n = max(C * Hg * Wg, 1)
    sign, loss_sum = _C().feature_resize_fwd(feature_map, gt, Hg, Wg, weight / n)
    grad = _C().feature_resize_bwd(sign, feature_map.shape[1], feature_map.shape[2])
    return loss_sum[0] * (weight / n), grad


class _FeatureL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature_map, gt, weight):
        loss, grad = feature_l1_loss_and_grad(feature_map, gt, weight)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def feature_l1_loss(feature_map: torch.Tensor, gt: torch.Tensor, weight: float = 1.0) -> torch.Tensor:
    """Autograd-aware drop-in for `l1_loss(F.interpolate(feature_map[None], gt.shape[1:], 'bilinear', True)[0], gt) * weight`."""
    return _FeatureL1.apply(feature_map, gt, float(weight))


class _Resize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature_map, Hg, Wg):
        ctx.hw = (feature_map.shape[1], feature_map.shape[2])
        out, _ = _C().feature_resize_fwd(feature_map, torch.empty(0, device=feature_map.device), Hg, Wg, 0.0)
        return out

    @staticmethod
    def backward(ctx, dout):
        return _C().feature_resize_bwd(dout.contiguous(), ctx.hw[0], ctx.hw[1]), None, None


def resize_bilinear(feature_map: torch.Tensor, size) -> torch.Tensor:
    """`F.interpolate(feature_map[None], size, mode='bilinear', align_corners=True)[0]` with the gather backward; use
    it in front of the optional 1x1 decoder: `l1_loss(decoder(resize_bilinear(fm, gt.shape[1:])), gt)`."""
    return _Resize.apply(feature_map, int(size[0]), int(size[1]))
