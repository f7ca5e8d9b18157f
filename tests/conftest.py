import gzip
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir(tmp_path_factory):
    """tests/golden with the gzipped indexes / databases unpacked into a session temp dir"""
    out = tmp_path_factory.mktemp("golden")
    for case in sorted(os.listdir(GOLDEN)):
        src = os.path.join(GOLDEN, case)
        if not os.path.isdir(src) or case.startswith("_"):
            continue
        dst = out / case
        dst.mkdir()
        for fn in os.listdir(src):
            if fn.endswith(".gz"):
                with gzip.open(os.path.join(src, fn), "rb") as fi, open(dst / fn[:-3], "wb") as fo:
                    shutil.copyfileobj(fi, fo)
            else:
                shutil.copy(os.path.join(src, fn), dst / fn)
    return str(out)
