#!/usr/bin/env python3
"""Generates tests/golden/xvc_hotpath_golden.npz from the UNMODIFIED reference.

Run in the dev container only (needs oracle/_ref/libxvcref.so, i.e. /root/reference):
    python tests/golden/make_golden.py
Every vector is an (input, output) pair of a reference function on the hot path, produced by
the reference's own code through oracle/ref_shim.cc (C table, i.e. simd=0; the reference's
own SimdTest proves its SIMD tables are bit-identical).  Inputs are stored next to the
outputs so the fixtures do not depend on any RNG implementation.  The GPU box has no
/root/reference: the tests only read the .npz.
"""
import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
from oracle import bindings  # noqa: E402
from xvc_b200 import abi, workload  # noqa: E402

OUT = os.path.join(HERE, "xvc_hotpath_golden.npz")


def main():
    ref = bindings.Ref()
    oracle = bindings.Oracle()          # only for the tap tables (data, checked by the filter vectors themselves)
    rng = np.random.default_rng(20261017)
    arrays, cases = {}, []

    def put(name, arr):
        arrays[name] = np.ascontiguousarray(arr)
        return name

    def case(kind, **kw):
        kw["kind"] = kind
        cases.append(kw)
        return len(cases) - 1

    # ---- metrics
    for bd in (8, 10):
        for w, h in ((4, 4), (8, 4), (4, 8), (8, 8), (16, 8), (8, 16), (16, 16), (32, 32), (64, 16), (16, 64), (64, 64), (2, 2), (2, 8), (32, 2)):
            a, b = common.rnd_samples(rng, h, w, bd), common.rnd_samples(rng, h, w, bd)
            r = common.rnd_resi(rng, h, w, bd)
            i = len(cases)
            exp = {}
            for metric in (abi.METRIC_SSD, abi.METRIC_SATD, abi.METRIC_SAD, abi.METRIC_SAD_FAST):
                if metric == abi.METRIC_SAD_FAST and h < 4:
                    continue
                exp["ss_%d" % metric] = int(ref.compare(metric, bd, a, b, w, h))
                exp["rs_%d" % metric] = int(ref.compare(metric, bd, r, b, w, h))
            exp["sad"] = int(ref.sad(0, a, b, w, h, bd))
            exp["ssd"] = int(ref.ssd(0, a, b, w, h, bd))
            exp["ssd_rr"] = int(ref.ssd(2, r, common.rnd_resi(np.random.default_rng(i), h, w, bd), w, h, bd))
            case("metric", bd=bd, w=w, h=h, a=put("c%d_a" % i, a), b=put("c%d_b" % i, b), r=put("c%d_r" % i, r),
                 r2=put("c%d_r2" % i, common.rnd_resi(np.random.default_rng(i), h, w, bd)), expect=exp)

    # ---- interpolation (all six kernels through FilterLuma/FilterChroma + bipred variants, add_avg)
    for bd in (8, 10):
        for chroma, (w, h) in itertools.product((0, 1), ((4, 4), (8, 16), (16, 8), (32, 32), (2, 4))):
            if not chroma and w < 4:
                continue
            nfrac = 32 if chroma else 16
            refblk = common.rnd_samples(rng, h + 8, w + 8, bd)
            i = len(cases)
            outs = {}
            for fx, fy in ((0, 0), (5, 0), (0, nfrac - 3), (nfrac // 2, 7), (1, nfrac - 1)):
                for bip in (0, 1):
                    p = np.zeros((h, w), dtype=np.int16 if bip else np.uint16)
                    ref.interp(chroma, bip, w, h, bd, fx, fy, refblk, (3, 3), p)
                    outs["%d_%d_%d" % (fx, fy, bip)] = put("c%d_p_%d_%d_%d" % (i, fx, fy, bip), p)
            a = rng.integers(-8192, 8191, size=(h, w)).astype(np.int16)
            b = rng.integers(-8192, 8191, size=(h, w)).astype(np.int16)
            shift = max(2, 14 - bd) + 1
            avg = np.zeros((h, w), dtype=np.uint16)
            ref.add_avg(w, h, (1 << (shift - 1)) + 2 * 8192, shift, bd, a, b, avg)
            case("interp", bd=bd, chroma=chroma, w=w, h=h, ref=put("c%d_ref" % i, refblk), outs=outs,
                 avg_a=put("c%d_avga" % i, a), avg_b=put("c%d_avgb" % i, b), avg=put("c%d_avg" % i, avg))

    # ---- transforms: DCT-2 every shape, every type pair on three shapes, DST 4x4, DC-only
    for bd in (8, 10):
        shapes = list(itertools.product((2, 4, 8, 16, 32, 64), repeat=2))
        for w, h in shapes:
            resi = common.rnd_resi(rng, h, w, bd)
            comp = 1 if (w == 2 or h == 2) else 0
            i = len(cases)
            co = ref.fwd_transform(w, h, bd, 0, 0, resi, comp=comp)
            cf = rng.integers(-32768, 32768, size=(h, w)).astype(np.int16)
            dc = np.zeros((h, w), dtype=np.int16)
            dc[0, 0] = rng.integers(-2000, 2000)
            case("tx", bd=bd, w=w, h=h, th=0, tv=0, dst=0, resi=put("c%d_resi" % i, resi), coeff=put("c%d_coeff" % i, co),
                 back=put("c%d_back" % i, ref.inv_transform(w, h, bd, 0, 0, 0, co, comp=comp)),
                 full=put("c%d_full" % i, cf), full_back=put("c%d_fullb" % i, ref.inv_transform(w, h, bd, 0, 0, 0, cf, comp=comp)),
                 dc=put("c%d_dc" % i, dc), dc_back=put("c%d_dcb" % i, ref.inv_transform(w, h, bd, 0, 0, 1, dc, comp=comp)))
        for (w, h), th, tv in itertools.product(((8, 8), (16, 4), (64, 32)), range(1, 6), range(1, 6)):
            resi = common.rnd_resi(rng, h, w, bd)
            i = len(cases)
            co = ref.fwd_transform(w, h, bd, th, tv, resi)
            case("tx", bd=bd, w=w, h=h, th=th, tv=tv, dst=0, resi=put("c%d_resi" % i, resi), coeff=put("c%d_coeff" % i, co),
                 back=put("c%d_back" % i, ref.inv_transform(w, h, bd, th, tv, 0, co)))
        resi = common.rnd_resi(rng, 4, 4, bd)
        i = len(cases)
        co = ref.fwd_transform(4, 4, bd, 0, 0, resi, comp=0, intra=1)
        case("tx", bd=bd, w=4, h=4, th=0, tv=0, dst=1, resi=put("c%d_resi" % i, resi), coeff=put("c%d_coeff" % i, co),
             back=put("c%d_back" % i, ref.inv_transform(4, 4, bd, 0, 0, 0, co, comp=0, intra=1)))

    # ---- QuantFast (+ sign hiding) and dequant
    for bd in (8, 10):
        for (w, h), qp in itertools.product(((4, 4), (8, 8), (16, 8), (4, 16), (32, 32), (64, 64), (2, 2), (8, 2)), (22, 32, 40)):
            comp = 1 if (w == 2 or h == 2) else int(rng.integers(0, 3))
            resi = rng.integers(-(1 << bd) // 2, (1 << bd) // 2, size=(h, w)).astype(np.int16)
            coeff = ref.fwd_transform(w, h, bd, 0, 0, resi, comp=1 if (w == 2 or h == 2) else 0)
            i = len(cases)
            lev0, nz0 = ref.quant_fast(w, h, bd, comp, qp, 0, coeff)
            lev1, nz1 = ref.quant_fast(w, h, bd, comp, qp, 1, coeff)
            case("quant", bd=bd, w=w, h=h, qp=qp, comp=comp, qp_bd=int(ref.qp(qp, bd)["qp_bitdepth"][comp]),
                 coeff=put("c%d_coeff" % i, coeff), lev_inter=put("c%d_l0" % i, lev0), nz_inter=int(nz0),
                 lev_intra=put("c%d_l1" % i, lev1), nz_intra=int(nz1),
                 deq=put("c%d_deq" % i, ref.dequant(w, h, bd, comp, qp, lev0)))

    # ---- one small picture through every picture-level stage
    for bd, pic_type in ((10, 0), (8, 1)):
        width, height, qp = 136, 72, 32
        cur, r0, r1 = common.frames(width, height, bd, 31 + bd)
        lam = workload.lambda_for_qp(qp)
        i = len(cases)
        # (1) ME jobs with varied predictors
        s = ref.session(width, height, bd, pic_type, qp, lam, simd=0, poc=8, sub_gop=16)
        s.set_orig(cur)
        s.add_ref(0, 0, 0, r0)
        if pic_type == 0:
            s.add_ref(1, 0, 16, r1)
        cus = workload.make_partition(width, height, seed=21, min_size=4, qp=qp)
        cus["flags"][::9] |= abi.CU_FULLPEL_MV
        s.set_cus(cus)
        nl = 2 if pic_type == 0 else 1
        jobs = common.me_jobs(cus, rng, nl, (128, 96), 300)
        me = s.me_search(jobs, lam, threads=4)
        # (2) whole picture pipeline
        s2 = ref.session(width, height, bd, pic_type, qp, lam, simd=0, poc=8, sub_gop=16)
        s2.set_orig(cur)
        s2.add_ref(0, 0, 0, r0)
        if pic_type == 0:
            s2.add_ref(1, 0, 16, r1)
        cus2 = workload.make_partition(width, height, seed=22, min_size=8, qp=qp)
        prm = common.picture_params(pic_type, lam)
        me2, tu2, cus_out = s2.encode_picture(prm, cus2, threads=4)
        # (3) deblocking alone on a blocky picture with mixed CU state (4-wide CUs -> edge chains)
        cus3 = common.deblock_cus(width, height, rng, 23, 4, pic_type)
        recp = common.blocky_recon(cur, cus3, rng, bd)
        s3 = ref.session(width, height, bd, pic_type, qp, lam, simd=0, poc=8, sub_gop=16)
        s3.add_ref(0, 0, 0, cur)
        if pic_type == 0:
            s3.add_ref(1, 0, 16, cur)
        s3.set_cus(cus3)
        s3.set_rec(recp)
        s3.deblock_picture(0, 0)
        # (4) motion compensation alone with arbitrary 1/16-pel vectors, uni + bi
        cus4 = common.mc_cus(width, height, rng, 24)
        if pic_type == 1:
            cus4["ref_idx"][:, 0], cus4["ref_idx"][:, 1] = 0, -1
            cus4["mv"][:, 1] = 0
        s4 = ref.session(width, height, bd, pic_type, qp, lam, simd=0, poc=8, sub_gop=16)
        s4.add_ref(0, 0, 0, r0)
        if pic_type == 0:
            s4.add_ref(1, 0, 16, r1)
        s4.set_cus(cus4)
        s4.motion_compensate(threads=2)
        case("picture", bd=bd, pic_type=pic_type, width=width, height=height, qp=qp, lam=lam,
             cur=[put("c%d_cur%d" % (i, c), cur[c]) for c in range(3)],
             r0=[put("c%d_r0%d" % (i, c), r0[c]) for c in range(3)],
             r1=[put("c%d_r1%d" % (i, c), r1[c]) for c in range(3)],
             me_cus=put("c%d_mecus" % i, cus.view(np.uint8)), me_jobs=put("c%d_mejobs" % i, jobs.view(np.uint8)),
             me_res=put("c%d_meres" % i, me.view(np.uint8)),
             enc_cus=put("c%d_enccus" % i, cus2.view(np.uint8)), enc_prm=put("c%d_encprm" % i, prm.view(np.uint8)),
             enc_me=put("c%d_encme" % i, me2.view(np.uint8)), enc_tu=put("c%d_enctu" % i, tu2.view(np.uint8)),
             enc_cus_out=put("c%d_enccusout" % i, cus_out.view(np.uint8)),
             enc_rec_padded=[put("c%d_encrec%d" % (i, c), s2.get_rec_padded(c)) for c in range(3)],
             enc_levels=[put("c%d_enclev%d" % (i, c), s2.get_coeff()[c]) for c in range(3)],
             db_cus=put("c%d_dbcus" % i, cus3.view(np.uint8)), db_in=[put("c%d_dbin%d" % (i, c), recp[c]) for c in range(3)],
             db_out=[put("c%d_dbout%d" % (i, c), s3.get_rec()[c]) for c in range(3)],
             mc_cus=put("c%d_mccus" % i, cus4.view(np.uint8)), mc_out=[put("c%d_mcout%d" % (i, c), s4.get_pred()[c]) for c in range(3)])

    arrays["__cases__"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print("wrote %s: %d cases, %d arrays, %.1f KiB" % (OUT, len(cases), len(arrays), os.path.getsize(OUT) / 1024.0))


if __name__ == "__main__":
    main()
