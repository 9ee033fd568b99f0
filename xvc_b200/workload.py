"""Synthetic workloads for tests and bench.py: YUV 4:2:0 content and CU partitions.

Content generator = the one BASELINE.md section 2 records for the reference probes
(seed 1234): 8x8-block uniform noise in [0,255], 5x5 box filter, + N(0,6) per-pixel noise,
clipped to 8 bit; frame i is the (W x H) crop at offset (2i mod 64, i mod 64) -- a steady pan over 32 frames;
U = 128 + 0.25*(Y_sub - 128), V = 128 - 0.25*(Y_sub - 128).  Samples are returned at the
encoder's internal bit depth (8-bit input shifted left by bitdepth-8, as
Resampler::ConvertFrom does for the default 10-bit internal pipeline, encoder.cc:445-480).
"""
import numpy as np

from . import abi


def synth_canvas(width, height, seed=1234):
    rng = np.random.default_rng(seed)
    cw, ch = width + 64, height + 64
    blocks = rng.integers(0, 256, size=((ch + 7) // 8, (cw + 7) // 8)).astype(np.float64)
    img = np.kron(blocks, np.ones((8, 8)))[:ch, :cw]
    pad = np.pad(img, 2, mode="edge")
    box = np.zeros_like(img)
    for dy in range(5):
        for dx in range(5):
            box += pad[dy:dy + ch, dx:dx + cw]
    box /= 25.0
    box += rng.normal(0.0, 6.0, size=box.shape)
    return np.clip(np.rint(box), 0, 255).astype(np.uint16)


def synth_frame(canvas, width, height, index, bitdepth=10, frame_noise=0.0):
    """Returns (Y, U, V) uint16 planes (tight) of frame `index`.  frame_noise > 0 adds N(0, frame_noise)
    (8-bit units, seeded by the frame index) to the luma crop before chroma is derived: camera noise that
    does not follow the pan, so that no reference picture predicts a block exactly (residuals to code,
    bi-prediction averaging two noisy references pays off -- as on real video)."""
    ox, oy = (2 * index) % 64, index % 64
    y8 = canvas[oy:oy + height, ox:ox + width].astype(np.int32)
    if frame_noise > 0:
        rng = np.random.default_rng([977, int(index) & 0xffff, width, height])
        y8 = np.clip(y8 + np.rint(rng.normal(0.0, frame_noise, size=y8.shape)).astype(np.int32), 0, 255)
    sub = (y8[0::2, 0::2] + y8[0::2, 1::2] + y8[1::2, 0::2] + y8[1::2, 1::2] + 2) >> 2
    u8 = np.clip(np.rint(128 + 0.25 * (sub - 128)), 0, 255).astype(np.int32)
    v8 = np.clip(np.rint(128 - 0.25 * (sub - 128)), 0, 255).astype(np.int32)
    sh = bitdepth - 8
    return tuple(np.ascontiguousarray((p << sh).astype(np.uint16)) for p in (y8, u8, v8))


def random_frame(width, height, bitdepth, rng):
    """Uniform random samples over the full legal range (worst case for arithmetic)."""
    hi = 1 << bitdepth
    return (rng.integers(0, hi, size=(height, width), dtype=np.uint16),
            rng.integers(0, hi, size=(height // 2, width // 2), dtype=np.uint16),
            rng.integers(0, hi, size=(height // 2, width // 2), dtype=np.uint16))


def make_partition(width, height, seed=7, min_size=8, max_size=64, qp=32, uniform=None):
    """Leaf CUs covering the picture, as an abi.cu_dtype array.

    Every 64x64 CTU is split recursively (quad / horizontal / vertical binary / none) by a
    seeded RNG; CUs that cross the picture edge are always split (the reference forces
    those splits too, cu_encoder.cc:123-273).  `uniform=N` gives an N x N grid instead.
    CU dimensions are powers of two in [min_size, max_size]; width/height must be
    multiples of 8 (SegmentHeader internal size, segment_header.h:51-62).
    """
    assert width % 8 == 0 and height % 8 == 0
    rng = np.random.default_rng(seed)
    out = []

    def emit(x, y, w, h, depth):
        out.append((x, y, w, h, depth))

    def split(x, y, w, h, depth):
        if x >= width or y >= height:
            return
        over_x, over_y = x + w > width, y + h > height
        if uniform is not None:
            want = "quad" if (w > uniform or h > uniform) else "none"
            if want == "none" and (over_x or over_y):
                want = "quad"
        elif over_x or over_y or w > max_size or h > max_size:
            want = "quad" if w == h and w > 8 else ("ver" if over_x or w > h else "hor")
        else:
            opts, probs = ["none"], [0.30 + 0.5 * (w * h <= 256)]
            if w == h and w > 8 and w // 2 >= min_size:
                opts.append("quad"); probs.append(0.40)
            if h // 2 >= min_size:
                opts.append("hor"); probs.append(0.15)
            if w // 2 >= min_size:
                opts.append("ver"); probs.append(0.15)
            p = np.array(probs) / sum(probs)
            want = opts[int(rng.choice(len(opts), p=p))]
        if want == "none":
            emit(x, y, w, h, depth)
        elif want == "quad":
            hw, hh = w // 2, h // 2
            for (sx, sy) in ((x, y), (x + hw, y), (x, y + hh), (x + hw, y + hh)):
                split(sx, sy, hw, hh, depth + 1)
        elif want == "hor":
            split(x, y, w, h // 2, depth + 1)
            split(x, y + h // 2, w, h // 2, depth + 1)
        else:
            split(x, y, w // 2, h, depth + 1)
            split(x + w // 2, y, w // 2, h, depth + 1)

    for cy in range(0, height, 64):
        for cx in range(0, width, 64):
            split(cx, cy, 64, 64, 0)
    cus = np.zeros(len(out), dtype=abi.cu_dtype)
    for i, (x, y, w, h, d) in enumerate(out):
        cus[i]["x"], cus[i]["y"], cus[i]["w"], cus[i]["h"], cus[i]["depth"] = x, y, w, h, d
    cus["qp"] = qp
    cus["ref_idx"] = -1
    return cus


def make_partition_tree(width, height, seed=7, min_size=8, qp=32, binary=True):
    """A partition xvc's syntax can signal, with its CU trees: (cus, splits).  Quad splits down to
    min_size; a quad-tree leaf may be split once more horizontally or vertically (binary depth 1, no
    further split below a binary split, so no sibling split restriction ever applies --
    coding_unit.cc:105-119, cu_writer.cc:58-76).  splits = the trees of all CTUs in raster order,
    pre-order, one byte per node: 0 leaf, 1 quad, 2 horizontal, 3 vertical (SplitType, cu_types.h:37-42).
    `depth` of a CU is its QUAD depth, as CodingUnit::GetDepth.  Picture size: multiples of 64."""
    assert width % 64 == 0 and height % 64 == 0
    rng = np.random.default_rng(seed)
    out, splits = [], []

    def node(x, y, w, h, depth):
        r = rng.random()
        if w == h and w // 2 >= min_size and r < 0.55:
            splits.append(1)
            hw = w // 2
            for (sx, sy) in ((x, y), (x + hw, y), (x, y + hw), (x + hw, y + hw)):
                node(sx, sy, hw, hw, depth + 1)
            return
        if binary and w == h and w // 2 >= min_size and r < 0.8:
            hor = rng.random() < 0.5
            splits.append(2 if hor else 3)
            for k in range(2):
                splits.append(0)
                out.append((x + (0 if hor else k * w // 2), y + (k * h // 2 if hor else 0), w if hor else w // 2, h // 2 if hor else h, depth))
            return
        splits.append(0)
        out.append((x, y, w, h, depth))

    for cy in range(0, height, 64):
        for cx in range(0, width, 64):
            node(cx, cy, 64, 64, 0)
    cus = np.zeros(len(out), dtype=abi.cu_dtype)
    for i, (x, y, w, h, d) in enumerate(out):
        cus[i]["x"], cus[i]["y"], cus[i]["w"], cus[i]["h"], cus[i]["depth"] = x, y, w, h, d
    cus["qp"] = qp
    cus["ref_idx"] = -1
    return cus, np.array(splits, dtype=np.uint8)


def check_partition(cus, width, height):
    """True when the CUs tile the picture exactly once."""
    cover = np.zeros((height // 4, width // 4), dtype=np.int32)
    for c in cus:
        cover[c["y"] // 4:(c["y"] + c["h"]) // 4, c["x"] // 4:(c["x"] + c["w"]) // 4] += 1
    return bool((cover == 1).all())


def lambda_for_qp(qp, pic_type_factor=0.68):
    """PictureEncoder::CalculateLambda (picture_encoder.cc:313-354) for a non-hierarchical
    inter picture with sub-GOP length 1: 0.68 * 2^((qp-12)/3)."""
    return pic_type_factor * 2.0 ** ((qp - 12) / 3.0)


def search_range_uni(poc, ref_poc, sub_gop_length=16, rmin=96, rmax=256):
    """InterSearch::GetSearchRangeUniPred, inter_search.cc:1050-1057."""
    r = (rmax * abs(poc - ref_poc) + sub_gop_length // 2) // sub_gop_length
    return max(rmin, min(rmax, r))


def true_motion(poc, ref_poc):
    """Displacement of the synthetic content between frame `poc` and frame `ref_poc` (1/16 pel)."""
    return (16 * ((2 * poc) % 64 - (2 * ref_poc) % 64), 16 * (poc % 64 - ref_poc % 64))


def set_predictors(cus, poc, ref_pocs, seed=0, jitter=6, exact=0.5):
    """Predictor field of a picture: cus[i].mv[list] = the content's motion towards the first picture of
    the list, exact for a fraction `exact` of the CUs and off by up to `jitter`/16 pel for the rest -- what
    a neighbour-derived predictor looks like on panned content.  ref_pocs = (poc of L0[0], poc of L1[0] or None)."""
    rng = np.random.default_rng(seed)
    n = len(cus)
    for l, rp in enumerate(ref_pocs):
        if rp is None:
            continue
        mv = np.array(true_motion(poc, rp), dtype=np.int32)
        noise = rng.integers(-jitter, jitter + 1, size=(n, 2)).astype(np.int32)
        noise[rng.random(n) < exact] = 0
        cus["mv"][:, l, :] = mv[None, :] + noise
    return cus


def mv_predictors(cus, poc, lists, seed=0, jitter=6, exact=0.5):
    """Predictor per (CU, list, reference picture) -> int32 [n][columns][2]: the content's motion towards
    THAT picture (what InterSearch::GetMvpList returns on panned content: neighbour vectors scaled by POC
    distance), exact for a fraction `exact` of the CUs, off by up to `jitter`/16 pel for the rest.
    lists = (POCs of list 0, POCs of list 1); columns ordered list 0 then list 1."""
    rng = np.random.default_rng(seed)
    n = len(cus)
    pocs = list(lists[0]) + list(lists[1])
    out = np.zeros((n, len(pocs), 2), dtype=np.int32)
    noise = rng.integers(-jitter, jitter + 1, size=(n, 2)).astype(np.int32)      # a CU's neighbourhood is off by the same amount for every picture
    noise[rng.random(n) < exact] = 0
    for c, rp in enumerate(pocs):
        out[:, c, :] = np.array(true_motion(poc, rp), dtype=np.int32)[None, :] + noise
    return out


def add_objects(planes, width, height, index, bitdepth=10, n_obj=None, seed=4321):
    """Textured rectangles that move on their own (constant velocity, up to 0.75 samples per frame on top of the
    pan; about an eighth of the picture is covered) pasted into frame `index` in place: content whose motion is not one global vector, so that a partition
    decided from the content has something to find.  Object k of frame i sits at p_k + v_k * i; luma and chroma
    are both replaced.  Returns `planes`."""
    rng = np.random.default_rng(seed)
    if n_obj is None:
        n_obj = max(4, width * height // 40000)
    tex = synth_frame(synth_canvas(width, height, seed + 1), width, height, 0, bitdepth)
    for _ in range(n_obj):
        ow, oh = 2 * int(rng.integers(8, 65)), 2 * int(rng.integers(8, 65))
        px, py = float(rng.integers(0, max(1, width - ow))), float(rng.integers(0, max(1, height - oh)))
        vx, vy = float(rng.uniform(-0.75, 0.75)), float(rng.uniform(-0.75, 0.75))
        tx, ty = 2 * int(rng.integers(0, max(1, (width - ow) // 2))), 2 * int(rng.integers(0, max(1, (height - oh) // 2)))
        x = 2 * (int(round(px + vx * index)) // 2)
        y = 2 * (int(round(py + vy * index)) // 2)
        x0, y0, x1, y1 = max(0, x), max(0, y), min(width, x + ow), min(height, y + oh)
        if x1 <= x0 or y1 <= y0:
            continue
        sx, sy = tx + (x0 - x), ty + (y0 - y)
        planes[0][y0:y1, x0:x1] = tex[0][sy:sy + (y1 - y0), sx:sx + (x1 - x0)]
        for c in (1, 2):
            planes[c][y0 // 2:y1 // 2, x0 // 2:x1 // 2] = tex[c][sy // 2:sy // 2 + (y1 - y0) // 2, sx // 2:sx // 2 + (x1 - x0) // 2]
    return planes
