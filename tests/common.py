"""Shared input builders for the parity tests (seeded; no reference access at run time)."""
import numpy as np

from oracle.bindings import Picture
from xvc_b200 import abi, workload

SIZES = [4, 8, 16, 32, 64]
CSIZES = [2, 4, 8, 16, 32]


def rnd_samples(rng, h, w, bd):
    return rng.integers(0, 1 << bd, size=(h, w), dtype=np.uint16)


def rnd_resi(rng, h, w, bd):
    return rng.integers(-(1 << bd) + 1, 1 << bd, size=(h, w)).astype(np.int16)


def frames(width, height, bd, seed, content="synth"):
    """(current, ref L0, ref L1) frame triples."""
    rng = np.random.default_rng(seed)
    if content == "synth":
        canvas = workload.synth_canvas(width, height, seed)
        return [workload.synth_frame(canvas, width, height, i, bd) for i in (8, 0, 16)]
    return [workload.random_frame(width, height, bd, rng) for _ in range(3)]


def me_jobs(cus, rng, nl, ranges, spread, slots=(0, 0)):
    jobs = np.zeros(len(cus) * nl, dtype=abi.me_job_dtype)
    for i in range(len(cus)):
        for l in range(nl):
            j = jobs[i * nl + l]
            j["cu"], j["list"], j["ref_slot"] = i, l, slots[l]
            j["search_range"] = ranges[l]
            j["mvp"] = rng.integers(-spread, spread + 1, size=2)
            j["prev"] = rng.integers(-spread // 16 - 1, spread // 16 + 2, size=2)
    return jobs


def mc_cus(width, height, rng, seed, min_size=4):
    cus = workload.make_partition(width, height, seed=seed, min_size=min_size)
    for i in range(len(cus)):
        mode = i % 3
        cus[i]["ref_idx"] = [(0, -1), (-1, 0), (0, 0)][mode]
        cus[i]["mv"] = rng.integers(-3000, 3001, size=(2, 2))
        if i % 5 == 0:
            cus[i]["mv"] = (rng.integers(-20, 21, size=(2, 2)) * 16)
        if mode == 0:
            cus[i]["mv"][1] = 0
        if mode == 1:
            cus[i]["mv"][0] = 0
    return cus


def deblock_cus(width, height, rng, seed, min_size, pic_type):
    cus = workload.make_partition(width, height, seed=seed, min_size=min_size)
    n = len(cus)
    cus["qp"] = rng.integers(25, 45, size=n)
    flags = np.zeros(n, dtype=np.uint8)
    flags[rng.random(n) < 0.15] |= abi.CU_INTRA
    flags[rng.random(n) < 0.4] |= abi.CU_CBF_Y
    cus["flags"] = flags
    for i in range(n):
        if flags[i] & abi.CU_INTRA:
            cus[i]["ref_idx"] = (-1, -1)
            continue
        mode = rng.integers(0, 3) if pic_type == 0 else 0
        cus[i]["ref_idx"] = [(0, -1), (-1, 0), (0, 0)][mode]
        base = rng.integers(-2, 3, size=(2, 2)) * 16
        cus[i]["mv"] = base + rng.integers(-10, 11, size=(2, 2))
        if mode == 0:
            cus[i]["mv"][1] = 0
        if mode == 1:
            cus[i]["mv"][0] = 0
    return cus


def blocky_recon(cur, cus, rng, bd):
    """A reconstruction with real block edges: per-CU DC offsets on top of the source."""
    recp = [p.astype(np.int32) for p in cur]
    for cu in cus:
        off = int(rng.integers(-6, 7)) << (bd - 8)
        recp[0][cu["y"]:cu["y"] + cu["h"], cu["x"]:cu["x"] + cu["w"]] += off
        for c in (1, 2):
            recp[c][cu["y"] // 2:(cu["y"] + cu["h"]) // 2, cu["x"] // 2:(cu["x"] + cu["w"]) // 2] += off
    return [np.clip(p, 0, (1 << bd) - 1).astype(np.uint16) for p in recp]


def picture_params(pic_type, lam, ranges=(128, 128), pocs=(0, 16), slots=None, deblock=1, pad=1):
    prm = np.zeros(1, dtype=abi.picture_params_dtype)
    prm["pic_type"] = pic_type
    prm["search_range"][0, 0, 0], prm["search_range"][0, 1, 0] = ranges
    prm["lambda_sqrt"] = np.sqrt(lam)
    prm["chroma_offset_table"] = 1
    prm["ref_poc"][0, 0, 0], prm["ref_poc"][0, 1, 0] = pocs
    prm["num_ref"] = 1
    prm["deblock"], prm["pad"] = deblock, pad
    prm["ref_slots"] = -1
    if slots is not None:
        prm["orig_slot"], prm["pred_slot"], prm["rec_slot"], prm["coeff_slot"] = slots["orig"], slots["pred"], slots["rec"], slots["coeff"]
        prm["ref_slots"][0, 0, 0], prm["ref_slots"][0, 1, 0] = slots["ref0"], slots["ref1"]
    return prm


def affine_cus(cus, rng):
    """xvcb200_affine_cu entries for every CU that may use affine motion (CodingUnit::CanUseAffine,
    coding_unit.h:285: w > 8 and h > 8): control points = a base MV plus spreads that exercise every
    sub-block size (4 ... whole block), the translation shortcut (mv[0] == mv[1]), a zero vertical
    spread, and MVs far outside the picture (clipping of control points and of sub-block MVs)."""
    idx = [i for i in range(len(cus)) if cus[i]["w"] > 8 and cus[i]["h"] > 8 and (cus[i]["ref_idx"][0] >= 0 or cus[i]["ref_idx"][1] >= 0)]
    aff = np.zeros(len(idx), dtype=abi.affine_cu_dtype)
    for k, i in enumerate(idx):
        aff[k]["cu"] = i
        for l in range(2):
            kind = (k + l) % 7
            base = rng.integers(-600, 601, size=2) if kind != 5 else rng.integers(-40000, 40001, size=2)
            spread = [1, 2, 3, 5, 9, 40, 200][(k // 7 + l) % 7]
            d1 = rng.integers(-spread, spread + 1, size=2)
            d2 = rng.integers(-spread, spread + 1, size=2)
            if kind == 3:
                d1[:] = 0           # translation shortcut, mv[2] ignored
            if kind == 4:
                d2[:] = 0           # sub-block height = whole block
            if kind == 6:
                base = (base // 16) * 16
                d1, d2 = d1 * 16, d2 * 16       # full-pel phases
            aff[k]["mv"][l] = [base, base + d1, base + d2]
    return aff


def lic_cus(cus, width, height, indices=None):
    """xvcb200_lic_cu entries: the CU covering (x, y-4) / (x-4, y) on a 4x4 map of the picture
    (CodingUnit::GetCodingUnitAbove / Left, coding_unit.cc:227-234, 275-282, with every CU of the
    array present); checked against the reference in tests/test_oracle_vs_ref.py."""
    cmap = np.full((height // 4, width // 4), -1, dtype=np.int32)
    for i, cu in enumerate(cus):
        cmap[cu["y"] // 4:(cu["y"] + cu["h"]) // 4, cu["x"] // 4:(cu["x"] + cu["w"]) // 4] = i
    idx = [i for i in (range(len(cus)) if indices is None else indices)
           if not (cus[i]["flags"] & abi.CU_INTRA) and (cus[i]["ref_idx"][0] >= 0 or cus[i]["ref_idx"][1] >= 0)]
    lic = np.zeros(len(idx), dtype=abi.lic_cu_dtype)
    for k, i in enumerate(idx):
        x, y = int(cus[i]["x"]), int(cus[i]["y"])
        a = cmap[y // 4 - 1, x // 4] if y > 0 else -1
        l = cmap[y // 4, x // 4 - 1] if x > 0 else -1
        lic[k] = (i, cus[a]["x"] if a >= 0 else -1, cus[a]["y"] if a >= 0 else -1, cus[l]["x"] if l >= 0 else -1, cus[l]["y"] if l >= 0 else -1)
    return lic


# Hierarchical-B sub-GOP of 8 after a key picture (POC 0), coding order, (poc, pic_type, reference POCs):
# waves of independent pictures 8 | 4 | 2 6 | 1 3 5 7 (thread_encoder.cc:99-131).
GOP8 = [(8, 1, (0,)), (4, 0, (0, 8)), (2, 0, (0, 4)), (6, 0, (4, 8)), (1, 0, (0, 2)), (3, 0, (2, 4)), (5, 0, (4, 6)), (7, 0, (6, 8))]


def gop_inputs(width, height, bd, qp, seed, min_size=8):
    """frame(poc) and inputs(poc) -> (original planes, CU array, picture parameters without slots)."""
    canvas = workload.synth_canvas(width, height, seed)
    lam = workload.lambda_for_qp(qp)
    by_poc = {p[0]: p for p in GOP8}

    def frame(poc):
        return workload.synth_frame(canvas, width, height, poc, bd)

    def inputs(poc):
        _, pic_type, refs = by_poc[poc]
        cus = workload.make_partition(width, height, seed=seed + poc, min_size=min_size, qp=qp)
        ranges = tuple(min(64, workload.search_range_uni(poc, r)) for r in refs) + ((64,) if len(refs) == 1 else ())
        prm = picture_params(pic_type, lam, ranges=ranges[:2], pocs=(tuple(refs) + (0,))[:2])
        return frame(poc), cus, prm
    return frame, inputs


def oracle_refs(oracle, width, height, r0, r1=None):
    refs = {(0, 0): Picture(width, height, 80, r0)}
    oracle.pad_border(refs[(0, 0)])
    if r1 is not None:
        refs[(1, 0)] = Picture(width, height, 80, r1)
        oracle.pad_border(refs[(1, 0)])
    return refs


def intra_jobs_in_coding_order(cus, width, height, comp=0):
    """xvcb200_intra_job per CU when the CUs are coded in array order: a neighbour is available once
    its CU has been coded.  Mirrors IntraPrediction::DetermineNeighbors (intra_prediction.cc:688-707)
    with CodingUnit::GetCuSizeAboveRight / GetCuSizeBelowLeft (coding_unit.cc:304-336) on a 4x4 map;
    checked against the reference in tests/test_intra_oracle.py."""
    coded = np.zeros((height // 4 + 1, width // 4 + 1), dtype=bool)
    sh = 1 if comp else 0
    jobs = np.zeros(len(cus), dtype=abi.intra_job_dtype)

    def at(x, y):
        return 0 <= x < width and 0 <= y < height and coded[y // 4, x // 4]

    for k, c in enumerate(cus):
        x, y, w, h = int(c["x"]), int(c["y"]), int(c["w"]), int(c["h"])
        ar = bl = 0
        if y > 0:
            for i in range(h, -1, -4):
                if at(x + w - 4 + i, y - 4):
                    ar = i
                    break
        if x > 0:
            for i in range(w, -1, -4):
                if at(x - 4, y + h - 4 + i):
                    bl = i
                    break
        j = jobs[k]
        j["x"], j["y"], j["w"], j["h"] = x >> sh, y >> sh, w >> sh, h >> sh
        j["has_left"], j["has_above"], j["has_above_left"] = x > 0, y > 0, x > 0 and y > 0
        j["above_right"], j["below_left"] = ar >> sh, bl >> sh
        coded[y // 4:(y + h) // 4, x // 4:(x + w) // 4] = True
    return jobs
