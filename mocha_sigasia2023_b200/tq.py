"""Element-wise quaternion helpers on CUDA tensors ([w,x,y,z]); thin torch plumbing used by the
set-up stage (features.py) and the NumPy-signature drop-ins (quat.py). Formulas: motion/quat.py."""
from __future__ import annotations

import torch


def cross(a, b):
    return torch.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                        a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                        a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], dim=-1)


def inv(q):
    return q * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=q.dtype, device=q.device)


def mul(x, y):
    x0, x1, x2, x3 = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    y0, y1, y2, y3 = y[..., 0], y[..., 1], y[..., 2], y[..., 3]
    return torch.stack([y0 * x0 - y1 * x1 - y2 * x2 - y3 * x3,
                        y0 * x1 + y1 * x0 - y2 * x3 + y3 * x2,
                        y0 * x2 + y1 * x3 + y2 * x0 - y3 * x1,
                        y0 * x3 - y1 * x2 + y2 * x1 + y3 * x0], dim=-1)


def mul_vec(q, x):
    u = q[..., 1:]
    t = 2.0 * cross(u.expand(x.shape[:-1] + (3,)), x) if u.shape[:-1] != x.shape[:-1] else 2.0 * cross(u, x)
    ue = u.expand(t.shape)
    return x + q[..., :1] * t + cross(ue, t)


def inv_mul(x, y):
    return mul(inv(x).expand(y.shape), y) if x.shape != y.shape else mul(inv(x), y)


def inv_mul_vec(q, x):
    return mul_vec(inv(q), x)
