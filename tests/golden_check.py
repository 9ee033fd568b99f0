"""Replays the golden vectors (tests/golden/xvc_hotpath_golden.npz, generated from the
unmodified reference by tests/golden/make_golden.py) against a backend: the C oracle (CPU) or
libxvc_b200.so (GPU, through the C ABI)."""
import json
import os

import numpy as np

from oracle.bindings import Picture
from xvc_b200 import abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xvc_hotpath_golden.npz")


def _picture_params(raw):
    """Stored picture parameters -> today's struct (fields appended since the vectors were made read as 0)."""
    buf = np.zeros(abi.picture_params_dtype.itemsize, dtype=np.uint8)
    buf[:len(raw)] = raw
    return buf.view(abi.picture_params_dtype).copy()


def load():
    z = np.load(GOLDEN)
    cases = json.loads(bytes(z["__cases__"]).decode())
    return z, cases


class OracleBackend:
    name = "oracle"

    def __init__(self, oracle):
        self.o = oracle

    def sad(self, a, b, w, h): return self.o.sad(int(a.dtype == np.int16), a, b, w, h)
    def ssd(self, a, b, w, h): return self.o.ssd(2 if b.dtype == np.int16 else int(a.dtype == np.int16), a, b, w, h)
    def compare(self, m, bd, a, b, w, h): return self.o.compare(m, bd, a, b, w, h)
    def interp(self, *a): return self.o.interp(*a)
    def add_avg(self, *a): return self.o.add_avg(*a)
    def fwd(self, w, h, bd, th, tv, dst, resi): return self.o.fwd_transform(w, h, bd, th, tv, dst, resi)
    def inv(self, w, h, bd, th, tv, dst, dc, c): return self.o.inv_transform(w, h, bd, th, tv, dst, dc, c)
    def quant(self, *a): return self.o.quant_fast(*a)
    def dequant(self, *a): return self.o.dequant(*a)

    def picture(self, z, c):
        o, bd, W, H, pt = self.o, c["bd"], c["width"], c["height"], c["pic_type"]
        g = lambda names: [z[n] for n in names]  # noqa: E731
        cur, r0, r1 = g(c["cur"]), g(c["r0"]), g(c["r1"])
        refs = {(0, 0): Picture(W, H, 80, r0)}
        o.pad_border(refs[(0, 0)])
        if pt == 0:
            refs[(1, 0)] = Picture(W, H, 80, r1)
            o.pad_border(refs[(1, 0)])
        orig = Picture(W, H, 0, cur)
        out = {}
        cus = z[c["me_cus"]].view(abi.cu_dtype).copy()
        jobs = z[c["me_jobs"]].view(abi.me_job_dtype).copy()
        out["me"] = o.me_search(orig, refs, bd, cus, jobs, np.sqrt(c["lam"]))
        cus2 = z[c["enc_cus"]].view(abi.cu_dtype).copy()
        prm = _picture_params(z[c["enc_prm"]])
        pred, rec = Picture(W, H, 80), Picture(W, H, 80)
        lev, me2, tu2 = o.encode_picture(orig, refs, pred, rec, bd, cus2, prm)
        out.update(enc_me=me2, enc_tu=tu2, enc_cus=cus2, enc_rec=rec.full, enc_lev=lev)
        cus3 = z[c["db_cus"]].view(abi.cu_dtype).copy()
        dbp = Picture(W, H, 80, g(c["db_in"]))
        o.deblock_picture(dbp, bd, cus3, pt, {(0, 0): 0, (1, 0): 16})
        out["db"] = dbp.planes()
        cus4 = z[c["mc_cus"]].view(abi.cu_dtype).copy()
        mcp = Picture(W, H, 80)
        o.motion_compensate(refs, bd, cus4, mcp)
        out["mc"] = mcp.planes()
        return out


class GpuBackend:
    name = "gpu"

    def __init__(self):
        from xvc_b200 import lib
        self.l = lib

    def sad(self, a, b, w, h): return self.l.sad(a, b, w, h)
    def ssd(self, a, b, w, h): return self.l.ssd(a, b, w, h)
    def compare(self, m, bd, a, b, w, h): return self.l.compare(m, bd, a, b, w, h)
    def interp(self, *a): return self.l.interp_block(*a)
    def add_avg(self, *a): return self.l.add_avg(*a)
    def fwd(self, w, h, bd, th, tv, dst, resi): return self.l.fwd_transform(w, h, bd, th, tv, dst, resi)
    def inv(self, w, h, bd, th, tv, dst, dc, c): return self.l.inv_transform(w, h, bd, th, tv, dst, dc, c)
    def quant(self, *a): return self.l.quant_fast(*a)
    def dequant(self, *a): return self.l.dequant(*a)

    def picture(self, z, c):
        bd, W, H, pt = c["bd"], c["width"], c["height"], c["pic_type"]
        g = lambda names: [z[n] for n in names]  # noqa: E731
        ctx = self.l.Context(W, H, bd, 7)
        ctx.upload(0, g(c["cur"]))
        ctx.upload(1, g(c["r0"]))
        ctx.pad_border(1)
        if pt == 0:
            ctx.upload(2, g(c["r1"]))
            ctx.pad_border(2)
        out = {}
        cus = z[c["me_cus"]].view(abi.cu_dtype).copy()
        jobs = z[c["me_jobs"]].view(abi.me_job_dtype).copy()
        jobs["ref_slot"] = 1 + jobs["list"]
        ctx.set_cus(cus)
        out["me"] = ctx.me_search(0, jobs, np.sqrt(c["lam"]))
        cus2 = z[c["enc_cus"]].view(abi.cu_dtype).copy()
        prm = _picture_params(z[c["enc_prm"]])
        prm["orig_slot"], prm["pred_slot"], prm["rec_slot"], prm["coeff_slot"] = 0, 3, 4, 5
        prm["ref_slots"][0, 0, 0], prm["ref_slots"][0, 1, 0] = 1, (2 if pt == 0 else -1)
        ctx.set_cus(cus2)
        me2, tu2 = ctx.encode_picture(prm)
        ctx.sync()
        out.update(enc_me=me2, enc_tu=tu2, enc_cus=ctx.get_cus(), enc_rec=[ctx.download_padded(4, k) for k in range(3)],
                   enc_lev=ctx.download_coeff(5))
        cus3 = z[c["db_cus"]].view(abi.cu_dtype).copy()
        ctx.upload(6, g(c["db_in"]))
        ctx.set_cus(cus3)
        ctx.deblock_picture(6, pt, {(0, 0): 0, (1, 0): 16})
        out["db"] = ctx.download(6)
        cus4 = z[c["mc_cus"]].view(abi.cu_dtype).copy()
        ctx.set_cus(cus4)
        ctx.motion_compensate({(0, 0): 1, (1, 0): 2 if pt == 0 else 1}, 3)
        out["mc"] = ctx.download(3)
        ctx.close()
        return out


def run_case(be, z, c):
    kind = c["kind"]
    if kind == "metric":
        a, b, r, r2, w, h, bd = z[c["a"]], z[c["b"]], z[c["r"]], z[c["r2"]], c["w"], c["h"], c["bd"]
        for key, exp in c["expect"].items():
            if key.startswith("ss_"):
                got = be.compare(int(key[3:]), bd, a, b, w, h)
            elif key.startswith("rs_"):
                got = be.compare(int(key[3:]), bd, r, b, w, h)
            elif key == "sad":
                got = be.sad(a, b, w, h)
            elif key == "ssd":
                got = be.ssd(a, b, w, h)
            else:
                got = be.ssd(r, r2, w, h)
            assert got == exp, (kind, key, w, h, bd)
    elif kind == "interp":
        w, h, bd, ch, ref = c["w"], c["h"], c["bd"], c["chroma"], z[c["ref"]]
        for key, name in c["outs"].items():
            fx, fy, bip = (int(v) for v in key.split("_"))
            p = np.zeros((h, w), dtype=np.int16 if bip else np.uint16)
            be.interp(ch, bip, w, h, bd, fx, fy, ref, (3, 3), p)
            assert np.array_equal(p, z[name]), (kind, ch, w, h, bd, key)   # exact (north_star tolerance: +-1 LSB)
        shift = max(2, 14 - bd) + 1
        avg = np.zeros((h, w), dtype=np.uint16)
        be.add_avg(w, h, (1 << (shift - 1)) + 2 * 8192, shift, bd, z[c["avg_a"]], z[c["avg_b"]], avg)
        assert np.array_equal(avg, z[c["avg"]])
    elif kind == "tx":
        w, h, bd, th, tv, dst = c["w"], c["h"], c["bd"], c["th"], c["tv"], c["dst"]
        assert np.array_equal(be.fwd(w, h, bd, th, tv, dst, z[c["resi"]]), z[c["coeff"]]), (kind, "fwd", w, h, th, tv)
        assert np.array_equal(be.inv(w, h, bd, th, tv, dst, 0, z[c["coeff"]]), z[c["back"]]), (kind, "inv", w, h, th, tv)
        if "full" in c:
            assert np.array_equal(be.inv(w, h, bd, th, tv, dst, 0, z[c["full"]]), z[c["full_back"]])
            assert np.array_equal(be.inv(w, h, bd, th, tv, dst, 1, z[c["dc"]]), z[c["dc_back"]])
    elif kind == "quant":
        w, h, bd, qbd = c["w"], c["h"], c["bd"], c["qp_bd"]
        l0, n0 = be.quant(w, h, bd, qbd, 0, 1, 0, z[c["coeff"]])
        l1, n1 = be.quant(w, h, bd, qbd, 1, 1, 0, z[c["coeff"]])
        assert n0 == c["nz_inter"] and np.array_equal(l0, z[c["lev_inter"]]), (kind, w, h, c["qp"])
        assert n1 == c["nz_intra"] and np.array_equal(l1, z[c["lev_intra"]]), (kind, w, h, c["qp"])
        assert np.array_equal(be.dequant(w, h, bd, qbd, z[c["lev_inter"]]), z[c["deq"]])
    elif kind == "picture":
        out = be.picture(z, c)
        exp_me = z[c["me_res"]].view(abi.me_result_dtype)
        for f in ("mv_fullpel", "mv", "cost_fullpel", "dist", "cost"):
            assert np.array_equal(out["me"][f], exp_me[f]), ("me", f)
        exp2 = z[c["enc_me"]].view(abi.me_result_dtype)
        for f in ("mv_fullpel", "mv", "cost_fullpel", "dist", "cost"):
            assert np.array_equal(out["enc_me"][f], exp2[f]), ("enc_me", f)
        assert np.array_equal(out["enc_tu"], z[c["enc_tu"]].view(abi.tu_result_dtype))
        exp_cus = z[c["enc_cus_out"]].view(abi.cu_dtype)
        for f in ("flags", "ref_idx", "mv"):
            assert np.array_equal(out["enc_cus"][f], exp_cus[f]), ("enc_cus", f)
        for k in range(3):
            assert np.array_equal(out["enc_rec"][k], z[c["enc_rec_padded"][k]]), ("enc_rec", k)
            assert np.array_equal(out["enc_lev"][k], z[c["enc_levels"][k]]), ("enc_lev", k)
            assert np.array_equal(out["db"][k], z[c["db_out"][k]]), ("deblock", k)
            assert np.array_equal(out["mc"][k], z[c["mc_out"][k]]), ("mc", k)
    else:
        raise AssertionError("unknown case kind %r" % kind)
