"""pytest configuration: markers and import paths.

`gpu` marks tests that need a CUDA device (run on the B200 box with `-m gpu`);
everything else must pass on a CPU-only machine (`-m "not gpu"`).
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
