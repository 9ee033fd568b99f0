"""Rate / distortion of a sequence coded through the GPU path beside the reference encoder's own stream of the same
pictures at the same QP ("report PSNR / bitrate beside xvcenc at the same QP").

    python tests/run_rd_compare.py [W H QP N_INTER] > gpurun_out/rd.json         (needs a GPU and oracle/_ref)

Low-delay chain, one reference picture, key picture coded by the reference encoder in both streams.  GPU path:
partition from the GPU pre-analysis, uni-prediction search, QuantFast, no merge / skip / intra CUs in inter pictures,
written by the reference's CuWriter (test infrastructure) and decoded by the UNMODIFIED xvcdec ("Conformance
verified", decoder output == GPU reconstruction for every picture).  Reference: its own RDO at speed_mode 2 (all
tools, RDOQ).  The comparison says what the hot path alone costs in bits -- mode decision is outside its scope."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformance  # noqa: E402
from oracle import bindings  # noqa: E402
from xvc_b200 import lib  # noqa: E402


def psnr(a, b, bd):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(((1 << bd) - 1) ** 2 / mse)


def main():
    width, height, qp, n_inter = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (448, 256, 32, 32)))
    bd = 10
    lib.load()
    ref = bindings.Ref()
    rep = {}
    size, _ = conformance.run(ref, conformance.gpu_backend(width, height, bd), width, height, bd, qp, 21, n_inter=n_inter,
                              partition="gpu", report=rep)
    n = 1 + n_inter
    dec_ref = conformance.decode(rep["reference_stream"], width, height, bd, n)
    orig = rep["originals"]
    ours_inter = sum(rep["inter_nal_bytes"]) + 4 * n_inter
    key_and_headers = len(rep["stream"]) - ours_inter
    ref_inter = len(rep["reference_stream"]) - key_and_headers
    y_ours = [psnr(rep["reconstructions"][k][0], orig[k + 1][0], bd) for k in range(n_inter)]
    y_ref = [psnr(dec_ref[k + 1][0], orig[k + 1][0], bd) for k in range(n_inter)]
    out = {"width": width, "height": height, "qp": qp, "frames": n, "inter_pictures": n_inter,
           "xvcdec": "Conformance verified; decoder output == GPU reconstruction for every inter picture",
           "gpu_path": {"inter_bytes": ours_inter, "bytes_per_inter_picture": ours_inter / n_inter, "psnr_y_db": float(np.mean(y_ours))},
           "reference_encoder": {"inter_bytes": ref_inter, "bytes_per_inter_picture": ref_inter / n_inter, "psnr_y_db": float(np.mean(y_ref)),
                                 "settings": "xvc encoder API, low delay, one reference picture, speed_mode 2, adaptive QP off"},
           "key_picture_and_headers_bytes": key_and_headers}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
