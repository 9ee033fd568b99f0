"""Bitstream conformance of the hot path ("xvcdec MD5 pass"): the decisions, levels and
reconstruction of an inter picture produced by a backend (the GPU library, or the C oracle for the CPU
check of the plumbing) are written into a real xvc bitstream by the reference's own CuWriter /
SyntaxWriter (oracle/ref_shim.cc), with the picture checksum taken over the backend's reconstruction; the
UNMODIFIED reference decoder application then decodes the stream and verifies that checksum against its
own reconstruction."""
import os
import subprocess
import tempfile

import numpy as np

import common
from oracle import bindings
from oracle.bindings import Picture
from xvc_b200 import abi, workload

XVCDEC = os.path.join(os.path.dirname(bindings.REF_SO), "xvcdec")


def gpu_partition(width, height, bd, cur, ref_rec, lam, qp):
    """(cus, splits) from xvcb200_decide_partition (the GPU pre-analysis) for one inter picture."""
    from xvc_b200 import lib
    ctx = lib.Context(width, height, bd, num_slots=2)
    ctx.upload(0, cur)
    ctx.upload(1, ref_rec)
    ctx.pad_border(1)
    out = ctx.decide_partition(0, 1, float(np.sqrt(lam)), qp)
    ctx.close()
    return out


def run(ref, backend, width=256, height=128, bd=10, qp=32, seed=3, search_range=64, n_inter=1, partition="seeded", report=None):
    """backend(cur, ref_rec, cus, prm, info) -> (cus_out, levels, rec planes incl. deblocking).
    n_inter inter pictures in a low-delay chain: picture k references picture k-1, i.e. from the second
    inter picture on the reference picture is itself a reconstruction made by the backend."""
    canvas = workload.synth_canvas(width, height, seed)
    pics = [workload.synth_frame(canvas, width, height, i, bd) for i in range(1 + n_inter)]
    conf = bindings.RefConformance(ref, width, height, bd, qp)
    if report is not None:      # the reference encoder's own stream of the same pictures (same key picture, same QP)
        own = bindings.RefConformance(ref, width, height, bd, qp)
        for p in pics:
            own.push(p)
        own.flush()
        report["reference_stream"] = own.bitstream()
        own.close()
        report["originals"] = pics
        report["inter_nal_bytes"] = []
    # The reference encoder keeps only a few pictures alive: picture k+1 is pushed after picture k has been replaced
    # (its reconstruction in the encoder's buffer is then the backend's, which is what picture k+1 references).
    conf.push(pics[0])
    recs = []
    for poc in range(1, 1 + n_inter):
        conf.push(pics[poc])
        if poc == n_inter:
            conf.flush()
        orig, ref_rec, info = conf.inter_inputs(poc, poc - 1)
        assert all(np.array_equal(a, b) for a, b in zip(orig, pics[poc]))
        if recs:
            assert all(np.array_equal(a, b) for a, b in zip(ref_rec, recs[-1]))
        # With one reference picture the reference still signals a bi-predictive picture (both lists hold
        # the same POC): the search runs on list 0 only, the deblocking decisions follow the signalled type.
        if partition == "gpu":
            cus, splits = gpu_partition(width, height, bd, orig, ref_rec, info["lam"], info["qp"])
            assert workload.check_partition(cus, width, height)
        else:
            cus, splits = workload.make_partition_tree(width, height, seed=seed + poc, min_size=8, qp=info["qp"])
        prm = common.picture_params(1, info["lam"], ranges=(search_range, search_range), pocs=(poc - 1, poc - 1),
                                    slots=dict(orig=0, ref0=1, ref1=-1, pred=2, rec=3, coeff=4), deblock=0, pad=0)
        prm["chroma_offset_table"], prm["chroma_offset_u"], prm["chroma_offset_v"] = info["chroma_table"], info["off_u"], info["off_v"]
        prm["beta_offset"], prm["tc_offset"] = info["beta_offset"], info["tc_offset"]
        info["ref_poc"] = poc - 1
        cus_out, levels, rec = backend(orig, ref_rec, cus, prm, info)
        assert np.any(cus_out["mv"]) and any(np.any(l) for l in levels)
        size = conf.write_inter(poc, cus_out, splits, levels, rec)
        assert size > 0, "the reference writer rejected the picture (%d)" % size
        recs.append(rec)
        if report is not None:
            report["inter_nal_bytes"].append(size)
    stream = conf.bitstream()
    conf.close()
    with tempfile.TemporaryDirectory() as tmp:
        bit, yuv = os.path.join(tmp, "s.xvc"), os.path.join(tmp, "out.yuv")
        open(bit, "wb").write(stream)
        res = subprocess.run([XVCDEC, "-bitstream-file", bit, "-output-file", yuv, "-output-bitdepth", str(bd)],
                             capture_output=True, text=True, timeout=600)
        log = res.stdout + res.stderr
        dec = np.fromfile(yuv, dtype=np.uint16) if os.path.exists(yuv) else None
    assert "Conformance verified" in log and res.returncode == 0, log[-2000:]
    # the decoder's output file: every inter picture equals the backend's reconstruction
    per = width * height * 3 // 2
    assert dec is not None and dec.size == (1 + n_inter) * per
    for k, rec in enumerate(recs):
        one = dec[(k + 1) * per:(k + 2) * per]
        got = [one[:width * height].reshape(height, width),
               one[width * height:width * height * 5 // 4].reshape(height // 2, width // 2),
               one[width * height * 5 // 4:].reshape(height // 2, width // 2)]
        for c in range(3):
            assert np.array_equal(got[c], rec[c]), (k, c)
    if report is not None:
        report["stream"], report["reconstructions"] = stream, recs
    return len(stream), log


def decode(stream, width, height, bd, n_pictures):
    """The UNMODIFIED xvcdec on a stream -> list of pictures (three planes each)."""
    with tempfile.TemporaryDirectory() as tmp:
        bit, yuv = os.path.join(tmp, "s.xvc"), os.path.join(tmp, "out.yuv")
        open(bit, "wb").write(stream)
        res = subprocess.run([XVCDEC, "-bitstream-file", bit, "-output-file", yuv, "-output-bitdepth", str(bd)],
                             capture_output=True, text=True, timeout=600)
        assert res.returncode == 0 and "Conformance verified" in res.stdout + res.stderr
        dec = np.fromfile(yuv, dtype=np.uint16)
    per = width * height * 3 // 2
    assert dec.size == n_pictures * per
    out = []
    for k in range(n_pictures):
        one = dec[k * per:(k + 1) * per]
        out.append([one[:width * height].reshape(height, width),
                    one[width * height:width * height * 5 // 4].reshape(height // 2, width // 2),
                    one[width * height * 5 // 4:].reshape(height // 2, width // 2)])
    return out


def oracle_backend(oracle, width, height, bd):
    def backend(cur, ref_rec, cus, prm, info):
        refs = common.oracle_refs(oracle, width, height, ref_rec)
        pred, rec = Picture(width, height, 80), Picture(width, height, 80)
        cus_o = cus.copy()
        levels, _, _ = oracle.encode_picture(Picture(width, height, 0, cur), refs, pred, rec, bd, cus_o, prm)
        if info["deblock"]:
            oracle.deblock_picture(rec, bd, cus_o, info["pic_type"], {(0, 0): info["ref_poc"], (1, 0): info["ref_poc"]}, info["beta_offset"], info["tc_offset"],
                                   info["chroma_table"], info["off_u"], info["off_v"])
        return cus_o, levels, rec.planes()
    return backend


def gpu_backend(width, height, bd):
    from xvc_b200 import lib

    def backend(cur, ref_rec, cus, prm, info):
        ctx = lib.Context(width, height, bd, num_slots=5)
        ctx.upload(0, cur)
        ctx.upload(1, ref_rec)
        ctx.pad_border(1)
        ctx.set_cus(cus)
        ctx.encode_picture(prm, want_results=False)
        if info["deblock"]:
            ctx.deblock_picture(3, info["pic_type"], {(0, 0): info["ref_poc"], (1, 0): info["ref_poc"]}, info["beta_offset"], info["tc_offset"],
                                info["chroma_table"], info["off_u"], info["off_v"])
        ctx.pad_border(3)
        ctx.sync()
        out = (ctx.get_cus(), ctx.download_coeff(4), ctx.download(3))
        ctx.close()
        return out
    return backend
