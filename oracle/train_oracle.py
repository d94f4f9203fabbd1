"""CPU oracle for the callers either side of the ST-block path.  TEST INFRASTRUCTURE ONLY --
imported by tests/ only; nothing under stgcn_b200/ may import it.

numpy restatements of
  * the AdamW update the reference's default optimizer performs (main.py:147-148 builds torch.optim.AdamW; the
    algorithm is PyTorch's documented one -- third-party, pinned torch~=2.2.0 in requirements.txt:5, 2.11.0 installed),
  * the reference's own Lion optimizer (script/opt.py:34-76),
  * data_transform, the window construction (script/dataloader.py:32-48).
Pinned by tests/test_train_oracle.py: AdamW against the installed torch.optim.AdamW, Lion and data_transform against
vectors generated from the UNMODIFIED reference by tests/golden/make_train_golden.py (tests/golden/train_*.npz), and live
against the reference when /root/reference is mounted."""
from __future__ import annotations

import numpy as np


def adamw_step(p, g, m, v, t, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
    """One torch.optim.AdamW step (amsgrad=False, maximize=False) on float32 arrays; t = 1-based step number.
    Returns (p, m, v) updated copies."""
    b1, b2 = betas
    p = p.astype(np.float32) * np.float32(1 - lr * weight_decay)          # decoupled weight decay
    m = (m + (g - m) * np.float32(1 - b1)).astype(np.float32)             # exp_avg.lerp_(grad, 1 - beta1)
    v = (v * np.float32(b2) + (g * g) * np.float32(1 - b2)).astype(np.float32)
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    denom = np.sqrt(v) / np.float32(np.sqrt(bc2)) + np.float32(eps)
    p = (p - np.float32(lr / bc1) * (m / denom)).astype(np.float32)
    return p, m, v


def lion_step(p, g, m, lr=1e-3, betas=(0.9, 0.99), weight_decay=1e-2):
    """One step of the reference's Lion (opt.py:56-74): decay, sign update from the interpolated momentum, momentum
    decay with the second coefficient.  Returns (p, m)."""
    b1, b2 = betas
    p = p.astype(np.float32) * np.float32(1 - lr * weight_decay)          # opt.py:59
    update = m * np.float32(b1) + g * np.float32(1 - b1)                  # opt.py:70
    p = (p - np.float32(lr) * np.sign(update)).astype(np.float32)         # opt.py:72
    m = (m * np.float32(b2) + g * np.float32(1 - b2)).astype(np.float32)  # opt.py:75
    return p, m


def data_transform(data, n_his, n_pred):
    """x[i, 0] = data[i : i + n_his], y[i] = data[i + n_his + n_pred - 1] for i < len - n_his - n_pred
    (dataloader.py:32-48).  float32 out (the reference converts through torch.Tensor)."""
    n_vertex = data.shape[1]
    num = len(data) - n_his - n_pred
    x = np.zeros([num, 1, n_his, n_vertex], dtype=np.float32)
    y = np.zeros([num, n_vertex], dtype=np.float32)
    for i in range(num):
        x[i, 0] = data[i: i + n_his]
        y[i] = data[i + n_his + n_pred - 1]
    return x, y
