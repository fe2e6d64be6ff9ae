import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: multi-GB allocations (full llama2-7B shape)")


def stories15m_path():
    """Real stories15M.bin: staged by __graft_entry__.build() into assets/ (the GPU box has no
    /root/reference), or straight from the reference mount in the build container."""
    for p in (os.path.join(ROOT, "assets", "stories15M.bin"), "/root/reference/stories15M.bin"):
        if os.path.exists(p):
            return p
    return None


@pytest.fixture(scope="session")
def stories15m():
    p = stories15m_path()
    if p is None:
        pytest.skip("stories15M.bin not staged (run __graft_entry__.build() where /root/reference exists)")
    return p


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.build_oracle()
    return oracle_lib


@pytest.fixture(scope="session")
def l2b():
    import llama2_zig_b200
    return llama2_zig_b200
