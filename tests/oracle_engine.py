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
