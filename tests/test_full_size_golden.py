"""CPU: the committed reference numbers of tests/golden/full_size_*_ref.json are reproduced by the reference
build in oracle/_ref (same script, --lib ref), and the size-independence the GPU test relies on holds for the
reference itself: energy per cell and the Gauss residual at a quarter of the golden size."""
import json
import os

import pytest

from tests import full_size as FS
from tests import helpers as H
from tests import helpers1d as H1

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden(code):
    with open(os.path.join(REPO, "tests", "golden", "full_size_%s_ref.json" % code)) as f:
        return json.load(f)


def _ref(code):
    lib = H.load_ref("em2d") if code == "em2d" else H1.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")
    return lib


@pytest.mark.parametrize("code,n,ppc", [("em2d", 256, 8), ("em1d", 1 << 14, 256)])
def test_golden_numbers_come_from_the_reference(code, n, ppc):
    g = _golden(code)
    m = FS.measure(code, _ref(code), n, ppc, g["steps"])
    for key in ("cells", "ppc", "np0", "npK"):
        assert m[key] == g[key]
    for key in ("energy_1", "energy_K", "field_energy_K", "gauss_max", "gauss_rms"):
        assert m[key] == g[key], key                      # the reference build is deterministic: same bits


@pytest.mark.parametrize("code,n,ppc", [("em2d", 128, 8), ("em1d", 1 << 12, 256)])
def test_reference_properties_do_not_depend_on_the_size(code, n, ppc):
    g = _golden(code)
    m = FS.measure(code, _ref(code), n, ppc, g["steps"])
    assert m["npK"] == m["np0"]
    assert max(m["charge_sum_rel"]) < 1e-6
    assert m["gauss_max"] < 1e-5
    for key in ("energy_1", "energy_K"):
        assert abs(m[key] / m["cells"] - g[key] / g["cells"]) < 1e-3 * abs(g[key] / g["cells"])
