import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "revisit-bpr_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # the C-ABI library is a build artefact (git-ignored): a fresh checkout builds it once, in-tree,
    # exactly as __graft_entry__.build() does (nvcc cross-compiles without a GPU)
    if not (PKG / "librbpr.so").exists():
        import shutil
        import subprocess
        if shutil.which("nvcc") or Path("/usr/local/cuda/bin/nvcc").exists():
            subprocess.run(["bash", str(PKG / "csrc" / "build.sh")], check=True, capture_output=True)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
