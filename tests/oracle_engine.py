"""CPU compute engine for exercising xvc_b200.sharding's host logic under gloo: the same
interface as sharding.GpuEngine, computed by the C oracle.  Test infrastructure only."""
import numpy as np
import torch

from oracle.bindings import Oracle, Picture


class OracleEngine:
    def __init__(self, width, height, bitdepth, cur, refs, ref_poc):
        self.o = Oracle()
        self.W, self.H, self.bd, self.ref_poc = width, height, bitdepth, ref_poc
        self.orig = Picture(width, height, 0, cur)
        self.refs = {}
        for key, planes in refs.items():
            self.refs[key] = Picture(width, height, 80, planes)
            self.o.pad_border(self.refs[key])
        self.pred, self.rec = Picture(width, height, 80), Picture(width, height, 80)
        self.cus = None

    def encode_band(self, cus_band, prm):
        p = prm.copy()
        p["deblock"], p["pad"] = 0, 0
        cus = cus_band.copy()
        self.o.encode_picture(self.orig, self.refs, self.pred, self.rec, self.bd, cus, p)
        return cus

    def set_cus(self, cus):
        self.cus = cus.copy()

    def deblock(self, pic_type, pass_mask, y0, y1):
        self.o.deblock_band(self.rec, self.bd, self.cus, pic_type, self.ref_poc, pass_mask, y0, y1)

    def get_rows(self, comp, y0, y1):
        return torch.from_numpy(np.ascontiguousarray(self.rec.plane(comp)[y0:y1]).view(np.uint8).copy())

    def put_rows(self, comp, y0, rows):
        self.rec.plane(comp)[y0:y0 + rows.shape[0]] = rows.numpy().view(np.uint16)

    def empty_rows(self, comp, n):
        return torch.empty((n, 2 * self.rec.width[comp]), dtype=torch.uint8)

    def to_comm(self, arr_u8):
        return torch.from_numpy(np.ascontiguousarray(arr_u8).copy())

    def from_comm(self, t):
        return t.numpy()

    def finish(self):
        pass


class OracleGopEngine:
    """One rank of sharding.FrameParallelGop computed by the C oracle; reconstructions travel as
    gloo broadcasts of the padded planes (the CPU stand-in for xvcb200_push_slot)."""

    def __init__(self, width, height, bitdepth, dist, inputs):
        self.o = Oracle()
        self.W, self.H, self.bd, self.dist, self.inputs = width, height, bitdepth, dist, inputs
        self.rec = {}

    def load_done(self, poc, planes):
        self.rec[poc] = Picture(self.W, self.H, 80, planes)
        self.o.pad_border(self.rec[poc])

    def encode(self, poc, pic_type, ref_pocs):
        cur, cus, prm = self.inputs(poc)
        refs = {(l, 0): self.rec[r] for l, r in enumerate(ref_pocs)}
        pred, rec = Picture(self.W, self.H, 80), Picture(self.W, self.H, 80)
        self.o.encode_picture(Picture(self.W, self.H, 0, cur), refs, pred, rec, self.bd, cus.copy(), prm)
        self.rec[poc] = rec

    def share(self, poc, owner):
        if poc not in self.rec:
            self.rec[poc] = Picture(self.W, self.H, 80)
        for c in range(3):
            t = torch.from_numpy(self.rec[poc].full[c].view(np.uint8))       # bytes: gloo has no 16-bit integer type
            self.dist.broadcast(t, src=owner)

    def fence(self):
        pass
