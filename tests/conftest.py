"""pytest configuration: the `gpu` marker and shared fixture helpers.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI export check (no compute calls).
`-m gpu`      : parity of the CUDA path against the oracle, through the C ABI.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name: str) -> dict:
    z = np.load(GOLDEN / f"{name}.npz")
    return {k: z[k] for k in z.files}


def rank_view(g: dict, r: int) -> dict:
    """Fixture of one rank of a multi-rank golden (keys r<r>_*), or the fixture itself."""
    pre = f"r{r}_"
    if f"{pre}counts" not in g:
        return g
    return {k[len(pre):]: v for k, v in g.items() if k.startswith(pre)}


def params_of(d: dict) -> dict:
    import refdump
    return refdump.params(d)


def rel_l2(a, b) -> float:
    nb = float(np.linalg.norm(b))
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b))) / (nb if nb > 0 else 1.0)


@pytest.fixture(scope="session")
def oracle():
    import hercules_oracle as ho
    ho.build()
    return ho
