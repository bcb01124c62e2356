import os
import sys

# engines of one process that stand in for the ranks of a collective (k-sharded recompute) spin on each other's flags:
# every stream needs a hardware queue of its own (the default 8 would alias 7 engine streams + the default stream)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
