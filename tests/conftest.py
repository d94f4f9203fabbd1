"""pytest configuration: registers the ``gpu`` marker and shared golden-vector helpers.

``-m "not gpu"`` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol checks,
2-process gloo tests.  ``-m gpu`` runs on a B200: CUDA path vs oracle/golden, through the C-ABI.
"""
import glob
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def golden_case_names():
    return sorted(os.path.basename(p)[len("case_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN, "case_*.npz")))


class GoldenCase:
    """One reference-generated vector set (see tests/golden/make_golden.py)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, f"case_{name}.npz"))
        self.name = name
        self.cfg = json.loads(str(z["cfg"]))
        self.x = torch.from_numpy(z["x"])
        self.y = torch.from_numpy(z["y"])
        self.gso = torch.from_numpy(z["gso"])
        self.out = torch.from_numpy(z["out"])
        self.block0_out = torch.from_numpy(z["block0_out"])
        self.loss = float(z["loss"])
        self.dx = torch.from_numpy(z["dx"])
        self.params = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p:")}
        self.grads = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("g:")}

    def model_cfg(self):
        c = self.cfg
        return dict(blocks=c["blocks"], kt=c["Kt"], n_his=c["n_his"], act=c["act"], kind=c["kind"])


def load_gso(dataset: str, kind: str = "cheb") -> torch.Tensor:
    return torch.from_numpy(np.load(os.path.join(GOLDEN, f"gso_{dataset}_{kind}.npy")))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().reshape(-1)
    b = b.detach().double().reshape(-1)
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
