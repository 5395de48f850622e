"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/psvi_b200.h
declares, host-only size queries agree with the oracle, and compute entry points fail loudly without a GPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nat():
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ge.CSRC, "libpsvi_b200.so")):
        ge.build()
    from psvi import _native
    return _native


def test_every_header_symbol_is_exported_and_bound(nat):
    hdr = open(os.path.join(ROOT, "include", "psvi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(psvi_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(nat.exported_symbols())
    L = nat.lib()
    for s in declared:
        assert hasattr(L, s), s


def test_size_queries_match_oracle(nat):
    from oracle import psvi_oracle as po
    for dims in ([2, 2], [2, 100, 2], [2, 24, 24, 4], [256, 1024, 10]):
        m = nat.make_model(dims, 10)
        assert nat.num_theta(m) == po.p_theta(dims)
        assert nat.traj_floats(m, 7) == 7 * 8 * po.p_theta(dims)
        assert nat.gout_floats(m, 13) == 2 * po.p_theta(dims) + 13 * dims[0] + 13 + 4 * 10 + 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(nat):
    with pytest.raises(nat.NativeError):
        nat.require_cuda()
    with pytest.raises(nat.NativeError):
        nat.make_noise(torch.zeros(4))  # host tensors are rejected, never silently computed on
