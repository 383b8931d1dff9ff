"""Pins the oracle: (1) against the committed golden fixtures, which were produced by the reference's own device
programs host-compiled in oracle/_ref (tests/golden/make_golden.py), and (2) where oracle/_ref/liboppref.so is
present, directly against those programs on further scenes.  Integer results (ray counts, photon cell keys, offset
tables, hit flags) must be identical; float images within 1e-5 (the reference's `exp` resolves to the double
overload on the host, the restatement uses expf as CUDA does; splat order differs)."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from golden.make_golden import CASES, render

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_reference_golden(oracle, case):
    gold = np.load(os.path.join(GOLDEN, case[0] + ".npz"))
    got = render("oracle", *case[1:])
    assert (got["rays"] == gold["rays"]).all()
    assert np.allclose(got["output"], gold["output"], rtol=1e-5, atol=2e-5 * float(gold["output"].max()))
    for k in ("photon_keys", "offsets", "hitpoint_flags"):
        if k in gold.files:
            assert (got[k] == gold[k]).all(), k


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("scene,cfg,method,maxlen", [
    ("CornellSmall", (1 << 5) | (1 << 8) | (1 << 10), 1, 10),     # distant point light, glossy floor, large mirror sphere, BDPT
    ("CornellSmall", (1 << 2) | (1 << 6) | (1 << 9) | (1 << 0), 1, 6),   # upward area light, blue back wall, SmallVCM colours, BDPT
    ("CornellSmall", (1 << 4) | (1 << 11), 2, 10),                # strong point light, large glass sphere, PPM
])
def test_oracle_matches_host_compiled_reference(oracle, scene, cfg, method, maxlen):
    a = render("oracle", scene, cfg, method, 40, 40, 2, maxlen, 0.05 if method == 2 else None)
    b = render("ref", scene, cfg, method, 40, 40, 2, maxlen, 0.05 if method == 2 else None)
    assert (a["rays"] == b["rays"]).all()
    assert np.allclose(a["output"], b["output"], rtol=1e-5, atol=2e-5 * float(b["output"].max()))
    if method == 2:
        assert (a["photon_keys"] == b["photon_keys"]).all() and (a["offsets"] == b["offsets"]).all()
