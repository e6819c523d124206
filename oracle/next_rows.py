"""TEST INFRASTRUCTURE ONLY (like the rest of oracle/): numpy restatements of the reference's callers either side of the
rasterizer (SURVEY.md section 8 f).  Pinned against PyTorch's own CPU operators in tests/test_next_rows.py (the reference
calls exactly those operators: train.py:100-104, scene/gaussian_model.py:98-121,163-190).

  resize_bilinear_ac   F.interpolate(x[None], size, mode='bilinear', align_corners=True)[0]      (train.py:100)
  feature_l1           l1_loss(resize(x), gt) * weight and its gradient w.r.t. x                  (train.py:100-104)
  activate             sigmoid / exp / normalize / cat                                            (gaussian_model.py:98-121)
  adam_step            torch.optim.Adam single-tensor update after the activation's Jacobian      (gaussian_model.py:163-190)
"""
import numpy as np


def _src(o, r, n):
    s = np.float32(r) * o.astype(np.float32)
    i0 = s.astype(np.int64)
    i1 = i0 + (i0 < n - 1)
    l1 = (s - i0.astype(np.float32)).astype(np.float32)
    return i0, i1, (np.float32(1) - l1).astype(np.float32), l1


def resize_bilinear_ac(x, Hg, Wg):
    C, H, W = x.shape
    ry = np.float32((H - 1) / (Hg - 1)) if Hg > 1 else np.float32(0)
    rx = np.float32((W - 1) / (Wg - 1)) if Wg > 1 else np.float32(0)
    y0, y1, ly0, ly1 = _src(np.arange(Hg), ry, H)
    x0, x1, lx0, lx1 = _src(np.arange(Wg), rx, W)
    a = x[:, y0][:, :, x0] * lx0 + x[:, y0][:, :, x1] * lx1
    b = x[:, y1][:, :, x0] * lx0 + x[:, y1][:, :, x1] * lx1
    return (ly0[None, :, None] * a + ly1[None, :, None] * b).astype(np.float32)


def resize_bilinear_ac_bwd(dout, H, W):
    C, Hg, Wg = dout.shape
    ry = np.float32((H - 1) / (Hg - 1)) if Hg > 1 else np.float32(0)
    rx = np.float32((W - 1) / (Wg - 1)) if Wg > 1 else np.float32(0)
    y0, y1, ly0, ly1 = _src(np.arange(Hg), ry, H)
    x0, x1, lx0, lx1 = _src(np.arange(Wg), rx, W)
    g = np.zeros((C, H, W), np.float64)
    for (yy, wy) in ((y0, ly0), (y1, ly1)):
        for (xx, wx) in ((x0, lx0), (x1, lx1)):
            contrib = dout.astype(np.float64) * wy[None, :, None] * wx[None, None, :]
            np.add.at(g, (slice(None), yy[:, None], xx[None, :]), contrib)
    return g.astype(np.float32)


def feature_l1(x, gt, weight=1.0):
    r = resize_bilinear_ac(x, gt.shape[1], gt.shape[2])
    d = r - gt
    n = d.size
    loss = np.float32(weight) * np.abs(d).astype(np.float64).sum() / n
    return np.float32(loss), resize_bilinear_ac_bwd((np.sign(d) * np.float32(weight / n)).astype(np.float32), x.shape[1], x.shape[2])


def activate(raw_opacity, raw_scaling, raw_rotation, f_dc, f_rest):
    f32 = np.float32
    op = (f32(1) / (f32(1) + np.exp(-raw_opacity.astype(f32)))).astype(f32)
    sc = np.exp(raw_scaling.astype(f32)).astype(f32)
    nrm = np.maximum(np.sqrt((raw_rotation.astype(f32) ** 2).sum(-1, keepdims=True)), f32(1e-12))
    return op, sc, (raw_rotation / nrm).astype(f32), np.concatenate((f_dc, f_rest), axis=1).astype(f32)


def raw_gradient(kind, param, grad_act, M=16):
    """Jacobian of the activation applied to the gradient w.r.t. the activated tensor."""
    f32 = np.float32
    if kind == "sigmoid":
        o = f32(1) / (f32(1) + np.exp(-param))
        return (grad_act * o * (f32(1) - o)).astype(f32)
    if kind == "exp":
        return (grad_act * np.exp(param)).astype(f32)
    if kind == "normalize4":
        nrm = np.maximum(np.sqrt((param ** 2).sum(-1, keepdims=True)), f32(1e-12))
        q = param / nrm
        return ((grad_act - q * (q * grad_act).sum(-1, keepdims=True)) / nrm).astype(f32)
    if kind == "sh_dc":
        return grad_act[:, 0:1, :].astype(f32)
    if kind == "sh_rest":
        return grad_act[:, 1:, :].astype(f32)
    return grad_act.astype(f32)


def adam_step(param, g, m, v, lr, step, b1=0.9, b2=0.999, eps=1e-15):
    f32 = np.float32
    m = (m + f32(1 - b1) * (g - m)).astype(f32)
    v = (v * f32(b2) + f32(1 - b2) * g * g).astype(f32)
    bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
    denom = (np.sqrt(v) * f32(1.0 / np.sqrt(bc2)) + f32(eps)).astype(f32)
    return (param - f32(lr / bc1) * (m / denom)).astype(f32), m, v
