"""pytest configuration: the `gpu` marker, import paths, and shared fixtures.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI surface (no GPU needed).
`-m gpu`       : parity tests proper -- every check calls the CUDA path through the C ABI (libftmpc.so).
Nothing here reads /root/reference (it does not exist on the GPU box).
"""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ft():
    import ftmpc_import
    return ftmpc_import.load()


@pytest.fixture(scope="session")
def oracle():
    import ftmpc_oracle
    return ftmpc_oracle


@pytest.fixture(scope="session")
def built():
    """Make sure libftmpc.so and the CPU checker exist (nvcc cross-compiles without a GPU)."""
    lib = ROOT / "fault-tolerant-mpc_b200" / "csrc" / "libftmpc.so"
    cpu = ROOT / "oracle" / "_cpu" / "libftmpc_cpu.so"
    if not lib.exists() or not cpu.exists():
        subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT, check=True)
    return lib, cpu


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(ROOT / "tests" / "golden" / "nlp_cases.npz")


@pytest.fixture(scope="session")
def accel_golden():
    """oracle KKT points for accelerating (circle) references: non-zero nominal wrench (tools/gen_golden.py --accel)"""
    import numpy as np
    return np.load(ROOT / "tests" / "golden" / "nlp_cases_accel.npz")


@pytest.fixture(scope="session")
def bench_golden():
    """oracle KKT points of 32 instances of bench.py's own workload (BASELINE configs[3]; tools/gen_golden.py --more)"""
    import numpy as np
    return np.load(ROOT / "tests" / "golden" / "nlp_cases_bench.npz")
