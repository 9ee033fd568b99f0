"""numpy statement of the GPU partition pre-analysis (xvc_b200/csrc/partition.cu, xvcb200_decide_partition):
test infrastructure, the checker the kernel is compared with bit for bit.  Not a restatement of reference code
-- the reference has no pre-analysis (its partition comes out of CuEncoder::CompressCu's RD recursion,
cu_encoder.cc:123-273); the rule is this repository's and is written down in include/xvc_b200.h."""
import numpy as np

from xvc_b200 import abi

R, SIDE = 8, 17
NVEC = SIDE * SIDE


def _eg_bits(v):
    u = (-v << 1) + 1 if v <= 0 else v << 1
    return 1 + 2 * (int(u).bit_length() - 1)


def decide(orig_luma, ref_luma, center16, lambda_sqrt, qp, bits_cu=8, bits_split=1):
    """orig_luma / ref_luma: tight uint16 planes.  Returns (cus, splits) like lib.Context.decide_partition."""
    h, w = orig_luma.shape
    lam = int(np.floor(65536.0 * lambda_sqrt))
    hdr_cu, hdr_split = (lam * bits_cu) >> 16, (lam * bits_split) >> 16
    cx, cy = center16[0] >> 4, center16[1] >> 4
    rate = np.array([(lam * (_eg_bits((m % SIDE - R) * 4) + _eg_bits((m // SIDE - R) * 4))) >> 16 for m in range(NVEC)], dtype=np.int64)
    # the padded plane continues the picture by replication; the kernel clamps at 80 samples, which is the same thing
    pad = 64 + R + max(abs(cx), abs(cy)) + 8
    refp = np.pad(ref_luma.astype(np.int64), pad, mode="edge")
    hb, wb = (h + 63) // 64 * 8, (w + 63) // 64 * 8
    op = np.zeros((hb * 8, wb * 8), dtype=np.int64)
    op[:h, :w] = orig_luma
    t8 = np.zeros((hb, wb, NVEC), dtype=np.int64)
    for my in range(SIDE):
        for mx in range(SIDE):
            oy, ox = pad + cy + my - R, pad + cx + mx - R
            sh = np.zeros_like(op)
            hh, ww = min(hb * 8, refp.shape[0] - oy), min(wb * 8, refp.shape[1] - ox)
            sh[:hh, :ww] = refp[oy:oy + hh, ox:ox + ww]
            d = np.abs(op - sh)
            t8[:, :, my * SIDE + mx] = np.minimum(d.reshape(hb, 8, wb, 8).sum(axis=(1, 3)), 65535)
    cus, splits = [], []

    def leaf(x, y, cw, ch, depth, m):
        mv = ((cx + m % SIDE - R) * 16, (cy + m // SIDE - R) * 16)
        cus.append((x, y, cw, ch, depth, mv))

    def best_of(tab):
        key = ((tab + rate) << 9) | np.arange(NVEC)
        k = int(key.min())
        return k >> 9, k & 511

    def node(x, y, size):
        """-> (best cost, emit function) of the node at (x, y)."""
        if x >= w or y >= h:
            return 0, None
        inside = x + size <= w and y + size <= h
        depth = {64: 0, 32: 1, 16: 2, 8: 3}[size]
        if size == 8:
            c, m = best_of(t8[y // 8, x // 8])
            return c + hdr_cu, lambda: (splits.append(0), leaf(x, y, 8, 8, depth, m))
        hs = size // 2
        kids = [node(x, y, hs), node(x + hs, y, hs), node(x, y + hs, hs), node(x + hs, y + hs, hs)]
        quad = sum(k[0] for k in kids) + hdr_split

        def emit_quad(flag):
            if flag:
                splits.append(1)
            for k in kids:
                if k[1]:
                    k[1]()
        if not inside:
            return quad, lambda: emit_quad(False)
        n8 = size // 8
        tab = t8[y // 8:y // 8 + n8, x // 8:x // 8 + n8]
        hb2 = n8 // 2
        tn, tt, tb = tab.sum(axis=(0, 1)), tab[:hb2].sum(axis=(0, 1)), tab[hb2:].sum(axis=(0, 1))
        tl, tr = tab[:, :hb2].sum(axis=(0, 1)), tab[:, hb2:].sum(axis=(0, 1))
        (cn, mn), (ct, mt), (cb, mb), (cl, ml), (cr, mr) = best_of(tn), best_of(tt), best_of(tb), best_of(tl), best_of(tr)
        none = cn + hdr_cu + hdr_split
        hor = ct + cb + 2 * hdr_cu + 2 * hdr_split
        ver = cl + cr + 2 * hdr_cu + 2 * hdr_split
        best, split = none, 0
        if hor < best:
            best, split = hor, 2
        if ver < best:
            best, split = ver, 3
        if quad < best:
            best, split = quad, 1
        if split == 0:
            return best, lambda: (splits.append(0), leaf(x, y, size, size, depth, mn))
        if split == 2:
            return best, lambda: (splits.extend([2, 0]), leaf(x, y, size, hs, depth, mt), splits.append(0), leaf(x, y + hs, size, hs, depth, mb))
        if split == 3:
            return best, lambda: (splits.extend([3, 0]), leaf(x, y, hs, size, depth, ml), splits.append(0), leaf(x + hs, y, hs, size, depth, mr))
        return best, lambda: emit_quad(True)

    for y in range(0, h, 64):
        for x in range(0, w, 64):
            node(x, y, 64)[1]()
    out = np.zeros(len(cus), dtype=abi.cu_dtype)
    for i, (x, y, cw, ch, depth, mv) in enumerate(cus):
        out[i]["x"], out[i]["y"], out[i]["w"], out[i]["h"], out[i]["depth"] = x, y, cw, ch, depth
        out[i]["mv"][0] = mv
        out[i]["mv"][1] = mv
    out["qp"] = qp
    out["ref_idx"] = -1
    return out, np.array(splits, dtype=np.uint8)
