import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The tests need the in-tree build products (they are git-ignored).  __graft_entry__.build() makes them; if a
    # fresh checkout runs pytest first, build what is missing (nvcc cross-compiles without a GPU).
    jobs = str(max(2, min(16, os.cpu_count() or 4)))
    if not os.path.exists(os.path.join(ROOT, "oracle", "libmoe_oracle.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=False)
    if (not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmoe_ref.so"))
            and os.path.isdir("/root/reference/moe/optimal_learning/cpp")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j", jobs, "ref"], check=False)
    if not os.path.exists(os.path.join(ROOT, "cornell-moe_b200", "libcornell_moe_b200.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "cornell-moe_b200"), "-j", jobs, "all"], check=False)
