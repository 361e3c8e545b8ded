import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _ensure_built():
    from oracle import oracle as orc
    if not os.path.exists(os.path.join(ROOT, "oracle", "libpml_oracle.so")):
        orc.build()
    if not os.path.exists(os.path.join(ROOT, "pepr_b200", "libpeprml.so")):
        from pepr_b200 import build
        build.build()


_ensure_built()


class Golden:
    """one reference-binary fixture: alignment + the raxmlHPC outputs recorded in tests/golden/<case>.json"""

    def __init__(self, case):
        from oracle import oracle as orc
        self.case = case
        self.meta = json.load(open(os.path.join(GOLD, case + ".json")))
        p = os.path.join(GOLD, case + ".phy")
        if os.path.exists(p):
            text = open(p).read()
        else:
            text = gzip.open(p + ".gz", "rt").read()
        toks = text.split()
        n = int(toks[0])
        self.names, self.seqs = toks[2::2][:n], toks[3::2][:n]
        self.codes = orc.encode(self.seqs)
        self.pat, self.w, self.s2p = orc.compress(self.codes)


_cache = {}


@pytest.fixture
def golden():
    def get(case):
        if case not in _cache:
            _cache[case] = Golden(case)
        return _cache[case]
    return get


@pytest.fixture(scope="session")
def gpu_ctx():
    import pepr_b200 as pb
    ctx = pb.Context(0)
    yield ctx
    ctx.close()
