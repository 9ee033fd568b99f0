#!/usr/bin/env python3
"""Manual multi-GPU check (not collected by pytest):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_banded_nccl.py [W H]
(BASELINE config 3: --nproc-per-node 4 ... run_banded_nccl.py 3840 2160)
One picture split into CTU-row bands across the ranks (NCCL halo exchange for deblocking) must
equal the single-GPU encode of the same picture bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from xvc_b200 import lib, sharding, workload  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, BD, QP = 1920, 1080, 10, 32
    if len(sys.argv) >= 3:
        W, H = int(sys.argv[1]), int(sys.argv[2])
    cur, r0, r1 = common.frames(W, H, BD, 77)
    cus = workload.make_partition(W, H, seed=41, min_size=4, qp=QP)
    prm = common.picture_params(0, workload.lambda_for_qp(QP), slots=dict(orig=0, ref0=1, ref1=2, pred=3, rec=4, coeff=5), pad=0)
    ctx = lib.Context(W, H, BD, 6, device=local)
    ctx.upload(0, cur)
    for s, f in ((1, r0), (2, r1)):
        ctx.upload(s, f)
        ctx.pad_border(s)
    eng = sharding.GpuEngine(ctx, dict(rec=4), BD, {(0, 0): 0, (1, 0): 16})
    enc = sharding.BandedPictureEncoder(eng, dist, rank, world, H)
    full = enc.encode(cus, prm)
    bands = [enc.gather_band_rows(c).cpu().numpy().view(np.uint16) for c in range(3)]
    # single-GPU truth on every rank (second context)
    ctx2 = lib.Context(W, H, BD, 6, device=local)
    ctx2.upload(0, cur)
    for s, f in ((1, r0), (2, r1)):
        ctx2.upload(s, f)
        ctx2.pad_border(s)
    ctx2.set_cus(cus)
    ctx2.encode_picture(prm, want_results=False)
    ctx2.sync()
    truth = ctx2.download(4)
    y0, y1 = enc.bands[rank]
    ok = all(np.array_equal(bands[c], truth[c][y0 >> (1 if c else 0):y1 >> (1 if c else 0)]) for c in range(3))
    cus_t = ctx2.get_cus()
    ok = ok and all(np.array_equal(full[f], cus_t[f]) for f in ("flags", "ref_idx", "mv"))
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("banded encode of %dx%d on %d GPUs (NCCL deblocking halo exchange) == single GPU: %s" % (W, H, world, bool(t.item())))
    dist.destroy_process_group()
    sys.exit(0 if t.item() else 1)


if __name__ == "__main__":
    main()
