"""Replays tests/golden/xvc_affine_golden.npz and xvc_lic_golden.npz (reference outputs of
MotionCompensation on affine / LIC CUs, see golden/make_affine_golden.py, make_lic_golden.py) against
a backend: the C oracle (CPU) or libxvc_b200.so (GPU)."""
import json
import os

import numpy as np

from oracle.bindings import Picture
from xvc_b200 import abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xvc_affine_golden.npz")


def oracle_backend(oracle):
    def run(c, r0, r1, cus, aff):
        W, H, bd = c["width"], c["height"], c["bd"]
        refs = {(0, 0): Picture(W, H, 80, r0), (1, 0): Picture(W, H, 80, r1)}
        for p in refs.values():
            oracle.pad_border(p)
        pred = Picture(W, H, 80)
        oracle.motion_compensate(refs, bd, cus, pred)
        oracle.motion_compensate_affine(refs, bd, cus, aff, pred)
        return [pred.plane(i) for i in range(3)]
    return run


def gpu_backend():
    from xvc_b200 import lib

    def run(c, r0, r1, cus, aff):
        ctx = lib.Context(c["width"], c["height"], c["bd"], num_slots=3)
        for slot, f in ((0, r0), (1, r1)):
            ctx.upload(slot, f)
            ctx.pad_border(slot)
        ctx.set_cus(cus)
        ctx.motion_compensate({(0, 0): 0, (1, 0): 1}, 2)
        ctx.motion_compensate_affine(aff, {(0, 0): 0, (1, 0): 1}, 2)
        out = ctx.download(2)
        ctx.close()
        return out
    return run


def replay(run):
    z = np.load(GOLDEN)
    cases = json.loads(bytes(z["__cases__"]).decode())
    assert len(cases) >= 3
    for c in cases:
        n = c["name"]
        r0 = [z["%s_r0_%d" % (n, i)] for i in range(3)]
        r1 = [z["%s_r1_%d" % (n, i)] for i in range(3)]
        cus = z[n + "_cus"].view(abi.cu_dtype).copy()
        aff = z[n + "_aff"].view(abi.affine_cu_dtype).copy()
        assert len(aff) == c["n_aff"]
        got = run(c, r0, r1, cus, aff)
        for i in range(3):
            assert np.array_equal(got[i], z["%s_pred_%d" % (n, i)]), (n, i)


# ---------------------------------------------------------------- LIC
LIC_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xvc_lic_golden.npz")


def oracle_lic_backend(oracle):
    def run(c, r0, r1, rec, cus, lic):
        W, H, bd = c["width"], c["height"], c["bd"]
        refs = {(0, 0): Picture(W, H, 80, r0), (1, 0): Picture(W, H, 80, r1)}
        for p in refs.values():
            oracle.pad_border(p)
        pred = Picture(W, H, 80)
        oracle.motion_compensate_lic(refs, Picture(W, H, 80, rec), bd, cus, lic, pred)
        return [pred.plane(i) for i in range(3)]
    return run


def gpu_lic_backend():
    from xvc_b200 import lib

    def run(c, r0, r1, rec, cus, lic):
        ctx = lib.Context(c["width"], c["height"], c["bd"], num_slots=4)
        for slot, f in ((0, r0), (1, r1)):
            ctx.upload(slot, f)
            ctx.pad_border(slot)
        ctx.upload(3, rec)
        ctx.set_cus(cus)
        ctx.motion_compensate_lic(lic, {(0, 0): 0, (1, 0): 1}, 3, 2)
        out = ctx.download(2)
        ctx.close()
        return out
    return run


def replay_lic(run):
    z = np.load(LIC_GOLDEN)
    cases = json.loads(bytes(z["__cases__"]).decode())
    assert len(cases) >= 3
    for c in cases:
        n = c["name"]
        g = lambda k: [z["%s_%s_%d" % (n, k, i)] for i in range(3)]   # noqa: E731
        cus = z[n + "_cus"].view(abi.cu_dtype).copy()
        lic = z[n + "_lic"].view(abi.lic_cu_dtype).copy()
        assert len(lic) == c["n_lic"]
        got = run(c, g("r0"), g("r1"), g("rec"), cus, lic)
        for i in range(3):
            assert np.array_equal(got[i], z["%s_pred_%d" % (n, i)]), (n, i)
