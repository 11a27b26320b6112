import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Build every in-tree library once per session (no-op when up to date)."""
    import subprocess
    subprocess.run(["make", "-s", "host", "tools", "oracle", "cli"], cwd=ROOT, check=True,
                   stdout=subprocess.DEVNULL)
    if not (ROOT / "swarm_b200" / "libswarm_b200.so").exists():
        subprocess.run(["make", "-s", "engine"], cwd=ROOT, check=True, stdout=subprocess.DEVNULL)
    return True


def pytest_terminal_summary(terminalreporter):
    import helpers
    r = helpers.REF_CHECKS
    if r["ran"] or r["skipped"]:
        terminalreporter.write_line(f"reference-binary comparisons inside tests: {r['ran']} ran, {r['skipped']} skipped "
                                    f"({'oracle/_ref/swarm present' if helpers.have_ref() else 'oracle/_ref/swarm MISSING'})")
