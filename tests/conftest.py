import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Make sure the in-tree libraries exist (built by `make` / __graft_entry__.build())."""
    import subprocess
    from volrend_b200 import LIB_PATH
    if not os.path.exists(LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-j8", "lib", "oracle"], cwd=ROOT)
    return True


@pytest.fixture(scope="session")
def synth_mod():
    from volrend_b200 import synth
    return synth


@pytest.fixture(scope="session")
def cfg1_tree(synth_mod):
    return synth_mod.make_config1_tree()


@pytest.fixture(scope="session")
def small_trees(synth_mod):
    """name -> SynthTree; small enough that the CPU oracle renders them in well under a second."""
    s = synth_mod
    return {
        "sh1_full4": s.make_config1_tree(),
        "sh16_d6": s.make_tree("lego", depth=6, basis_dim=16, seed=1),
        "sh9_d6": s.make_tree("lego", depth=6, basis_dim=9, seed=2),
        "sh4_d5": s.make_tree("lego", depth=5, basis_dim=4, seed=3),
        "sh25_d5": s.make_tree("drums", depth=5, basis_dim=25, seed=4),
        "rgba_d5": s.make_tree("lego", depth=5, fmt="RGBA", seed=5),
        "sg9_d5": s.make_tree("lego", depth=5, basis_dim=9, fmt="SG", seed=6),
        "asg4_d5": s.make_tree("lego", depth=5, basis_dim=4, fmt="ASG", seed=7),
        "sg7_d5": s.make_tree("lego", depth=5, basis_dim=7, fmt="SG", seed=8),   # basis not in {1,4,9,16,25}
    }
