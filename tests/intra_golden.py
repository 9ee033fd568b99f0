"""Replays tests/golden/xvc_intra_golden.npz (reference outputs, see golden/make_intra_golden.py)
against a backend: the C oracle (CPU) or libxvc_b200.so (GPU, through the C ABI)."""
import json
import os

import numpy as np

from xvc_b200 import abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xvc_intra_golden.npz")


class OracleBackend:
    def __init__(self, oracle):
        self.o = oracle

    def refs(self, w, h, bd, nb, plane, x, y): return self.o.intra_ref_samples(w, h, bd, nb, plane, x, y)
    def predict(self, mode, w, h, bd, luma, ref, filt): return self.o.intra_predict(mode, w, h, bd, luma, ref, filt)

    def scan(self, bd, cur, rec, jobs, refs, filts):
        return np.stack([self.o.intra_satd_scan(int(j["w"]), int(j["h"]), bd, cur[0], int(j["x"]), int(j["y"]), refs[i], filts[i])
                         for i, j in enumerate(jobs)])


class GpuBackend:
    def __init__(self):
        from xvc_b200 import lib
        self.lib = lib

    def refs(self, w, h, bd, nb, plane, x, y): return self.lib.intra_ref_samples(w, h, bd, nb, plane, x, y)
    def predict(self, mode, w, h, bd, luma, ref, filt): return self.lib.intra_predict(mode, w, h, bd, luma, ref, filt)

    def scan(self, bd, cur, rec, jobs, refs, filts):
        h, w = cur[0].shape
        ctx = self.lib.Context(w, h, bd, num_slots=2)
        ctx.upload(0, cur)
        ctx.upload(1, rec)
        out = ctx.intra_satd_scan(0, 1, jobs)
        ctx.close()
        return out


def replay(backend):
    z = np.load(GOLDEN)
    cases = json.loads(bytes(z["__cases__"]).decode())
    assert len(cases) >= 2
    n_pred = 0
    for c in cases:
        bd, comp = c["bd"], c["comp"]
        cur = [z[c["name"] + "_cur%d" % i] for i in range(3)]
        rec = [z[c["name"] + "_rec%d" % i] for i in range(3)]
        jobs = z[c["name"] + "_jobs"].view(abi.intra_job_dtype)
        ref_r, filt_r, pred_r = z[c["name"] + "_ref"], z[c["name"] + "_filt"], z[c["name"] + "_pred"]
        off = 0
        for i, j in enumerate(jobs):
            w, h = int(j["w"]), int(j["h"])
            nb = (j["has_above_left"], j["has_above"], j["above_right"], j["has_left"], j["below_left"])
            ref_b, filt_b = backend.refs(w, h, bd, nb, rec[comp], int(j["x"]), int(j["y"]))
            used = np.zeros(2 * abi.INTRA_REF_STRIDE, dtype=bool)
            used[:w + h + 1] = True
            used[abi.INTRA_REF_STRIDE:abi.INTRA_REF_STRIDE + w + h] = True
            assert np.array_equal(ref_b[used], ref_r[i][used]), (c["name"], i)
            if comp == 0:
                assert np.array_equal(filt_b[used], filt_r[i][used]), (c["name"], i)
            for mode in range(abi.INTRA_NUM_MODES):
                want = pred_r[off:off + w * h].reshape(h, w)
                off += w * h
                if (i + mode) % c["pred_stride"]:
                    continue
                got = backend.predict(mode, w, h, bd, comp == 0, ref_r[i], filt_r[i] if comp == 0 else None)
                assert np.array_equal(got, want), (c["name"], i, mode)
                n_pred += 1
        if comp == 0:
            assert np.array_equal(backend.scan(bd, cur, rec, jobs, ref_r, filt_r), z[c["name"] + "_satd"]), c["name"]
    assert n_pred > 500


# ---------------------------------------------------------------- LM chroma
LM_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xvc_lm_golden.npz")


def lm_jobs(cus):
    jobs = np.zeros(len(cus), dtype=abi.intra_job_dtype)
    jobs["x"], jobs["y"], jobs["w"], jobs["h"] = cus["x"], cus["y"], cus["w"], cus["h"]
    return jobs


def lm_oracle_backend(oracle):
    def run(c, rec, cus):
        pred = [np.zeros_like(rec[1]), np.zeros_like(rec[2])]
        for cu in cus:
            x, y, w, h = int(cu["x"]), int(cu["y"]), int(cu["w"]), int(cu["h"])
            for comp in (1, 2):
                pred[comp - 1][y // 2:(y + h) // 2, x // 2:(x + w) // 2] = oracle.intra_lm_chroma(rec, comp, x, y, w, h, c["bd"])
        return pred
    return run


def lm_gpu_backend():
    from xvc_b200 import lib

    def run(c, rec, cus):
        ctx = lib.Context(c["width"], c["height"], c["bd"], num_slots=2)
        ctx.upload(0, rec)
        ctx.intra_lm_chroma(0, lm_jobs(cus), 1)
        out = ctx.download(1)
        ctx.close()
        return out[1:]
    return run


def replay_lm(run):
    z = np.load(LM_GOLDEN)
    cases = json.loads(bytes(z["__cases__"]).decode())
    assert len(cases) >= 3
    for c in cases:
        n = c["name"]
        rec = [z["%s_rec_%d" % (n, i)] for i in range(3)]
        cus = z[n + "_cus"].view(abi.cu_dtype).copy()
        assert len(cus) == c["n"]
        got = run(c, rec, cus)
        for comp in (1, 2):
            assert np.array_equal(got[comp - 1], z["%s_pred_%d" % (n, comp)]), (n, comp)
