"""CPU-side checks: the C-ABI library loads and exports every symbol include/awb.h declares, the
drop-in modules keep the reference's state-dict contract, and nothing computes without a GPU."""
import os
import re

import pytest
import torch

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    entry.build()


def test_header_symbols_exported():
    from awesome_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "awb.h")).read()
    declared = set(re.findall(r"\b(awb_[a-z_0-9]+)\s*\(", header))
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, f"header/binding mismatch: {declared ^ bound}"
    for name in declared:
        assert hasattr(lib, name)
    assert b"sm_100a" in lib.awb_version()


def test_state_dict_contract_matches_reference_fixture(golden):
    import awesome_b200 as A
    g = golden("icnn_c1.pt")
    m = A.ConvexNextNet(n_hidden_layers=1)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["init"].keys())
    for k in sd:
        assert sd[k].shape == g["init"][k].shape
    g2 = golden("icnn_c2.pt")
    m2 = A.ConvexNextNet(n_hidden=130, in_features=2, n_hidden_layers=2)
    assert list(m2.state_dict().keys()) == list(g2["init"].keys())
    m2.load_state_dict(g2["init"])
    # parameters stay views of one flat arena, in state_dict order
    off = m2._arena.data_ptr()
    for p in m2.parameters():
        assert p.data_ptr() == off
        off += 4 * p.numel()
    assert m2._arena.numel() == 35103
    cn = A.ConvexNet()
    assert list(cn.state_dict().keys()) == list(g2["convexnet_init"].keys())


def test_same_seed_same_init_as_reference(golden):
    import awesome_b200 as A
    g = golden("icnn_c1.pt")
    torch.manual_seed(0)
    m = A.ConvexNextNet(n_hidden_layers=1)
    for k, v in m.state_dict().items():
        assert torch.equal(v, g["init"][k]), k


def test_no_cpu_fallback():
    import awesome_b200 as A
    from awesome_b200._lib import AwbLibraryError
    m = A.ConvexNextNet()
    if not torch.cuda.is_available():
        with pytest.raises(AwbLibraryError):
            m(torch.zeros(1, 2, 4, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "awesome_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                for pat in ("import oracle", "from oracle", "oracle.", "oracle/"):
                    assert pat not in src, f"{fn} reaches into the oracle ({pat!r})"
