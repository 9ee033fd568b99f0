"""RDOQ with frozen contexts: the DEFINITION, pinned to the reference -- no GPU kernel yet.
RdoQuant::QuantRdo only reads its SyntaxWriter; run against the context state a picture starts from (never
advanced) the levels of a transform unit do not depend on the units coded before it, which is what lets a
picture's units be quantised at once.  These tests hold the reference-generated golden vectors
(tests/golden/make_rdoq_golden.py) against the reference build and check that property."""
import json
import os

import numpy as np
import pytest

from oracle import bindings

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xvc_rdoq_golden.npz")


@pytest.fixture(scope="module")
def ref():
    if not bindings.have_ref():
        pytest.skip("oracle/_ref/libxvcref.so not built (needs /root/reference)")
    r = bindings.Ref()
    if not hasattr(r.L, "xref_quant_rdo_frozen"):
        pytest.skip("oracle/_ref predates xref_quant_rdo_frozen")
    return r


def _cases():
    z = np.load(GOLDEN)
    return z, json.loads(bytes(z["__cases__"]).decode())


def test_golden_file_is_consistent():
    z, cases = _cases()
    assert len(cases) >= 32
    for c in cases:
        a, b = z[c["name"] + "_in"], z[c["name"] + "_out"]
        assert a.shape == b.shape == (c["h"], c["w"]) and a.dtype == b.dtype == np.int16
    assert any(np.any(z[c["name"] + "_out"]) for c in cases)


def test_reference_reproduces_golden(ref):
    z, cases = _cases()
    for c in cases:
        lev, nz = ref.quant_rdo_frozen(c["w"], c["h"], c["bitdepth"], c["comp"], c["qp"], c["lam"], c["intra_pic"], z[c["name"] + "_in"],
                                       intra_cu=c["intra_cu"], intra_mode=1 if c["intra_cu"] else 0)
        assert np.array_equal(lev, z[c["name"] + "_out"]) and nz == c["ret"], c["name"]


def test_units_do_not_depend_on_each_other(ref):
    """Frozen contexts: the same unit gives the same levels whatever was quantised in between."""
    z, cases = _cases()
    first = cases[1]
    args = (first["w"], first["h"], first["bitdepth"], first["comp"], first["qp"], first["lam"], first["intra_pic"], z[first["name"] + "_in"])
    a, _ = ref.quant_rdo_frozen(*args)
    for c in cases[2:12]:
        ref.quant_rdo_frozen(c["w"], c["h"], c["bitdepth"], c["comp"], c["qp"], c["lam"], c["intra_pic"], z[c["name"] + "_in"], intra_cu=c["intra_cu"],
                             intra_mode=1 if c["intra_cu"] else 0)
    b, _ = ref.quant_rdo_frozen(*args)
    assert np.array_equal(a, b)


def test_rdoq_differs_from_quant_fast(ref):
    """The two quantisers are different decisions on the same coefficients (the reason RDOQ is a row of its own)."""
    z, cases = _cases()
    differ = 0
    for c in cases:
        if c["intra_cu"]:
            continue
        fast, _ = ref.quant_fast(c["w"], c["h"], c["bitdepth"], c["comp"], c["qp"], c["intra_pic"], z[c["name"] + "_in"])
        differ += int(not np.array_equal(fast, z[c["name"] + "_out"]))
    assert differ > 5
