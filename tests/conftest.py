import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
LIB = os.path.join(ROOT, "bayex_b200", "libbayex_b200.so")


def _ensure_library():
    """The shared library is a build artefact (git-ignored).  A fresh checkout builds it once here if nvcc is present - it
    cross-compiles sm_100a without a GPU; without nvcc the tests that need it are skipped (never a CPU fallback)."""
    if os.path.exists(LIB) or os.environ.get("BAYEX_B200_LIB"):
        return True
    import shutil
    import subprocess
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        subprocess.run(["bash", os.path.join(ROOT, "bayex_b200", "csrc", "build.sh")], check=False, capture_output=True)
    return os.path.exists(LIB)


HAVE_LIB = _ensure_library()


def pytest_collection_modifyitems(config, items):
    if HAVE_LIB:
        return
    skip = pytest.mark.skip(reason="libbayex_b200.so is not built and nvcc is unavailable (bayex_b200 has no CPU fallback)")
    for item in items:
        if os.path.basename(str(item.fspath)) in ("test_abi.py", "test_host_logic.py", "test_dist_gloo.py", "test_gpu_parity.py",
                                                  "test_multi_gpu.py"):
            item.add_marker(skip)


def pytest_ignore_collect(collection_path, config):
    # modules that import the package at module level cannot even be collected without the library
    if not HAVE_LIB and collection_path.name in ("test_host_logic.py", "test_gpu_parity.py"):
        return True
    return None


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "reference_shim.npz")))
