#!/usr/bin/env python3
"""Frame-parallel GOP encode across GPUs (BASELINE config 5; not collected by pytest):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 tests/run_frame_parallel_gop.py [W H] [same-gpu]
A hierarchical-B sub-GOP of 8 after a key picture: the pictures of a wave (8 | 4 | 2 6 | 1 3 5 7)
are encoded on different GPUs, every finished, padded reconstruction is pushed into the same slot
of every other GPU (xvcb200_push_slot, CUDA IPC + copy engines over NVLink) and referenced there
by the next wave.  Every rank must end up with all eight reconstructions, bit for bit equal to
the serial encode of the same chain on one GPU (which rank 0 also runs).  `same-gpu`: all ranks
on cuda:0 with a gloo rendezvous (the form tests/test_gpu_configs.py launches on a one-GPU box)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from xvc_b200 import lib, sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    args = [a for a in sys.argv[1:] if a != "same-gpu"]
    same_gpu = "same-gpu" in sys.argv[1:]
    W, H = (int(args[0]), int(args[1])) if len(args) >= 2 else (1920, 1080)
    BD, QP = 10, 32
    dev = 0 if same_gpu else local
    torch.cuda.set_device(dev)
    if same_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    frame, inputs = common.gop_inputs(W, H, BD, QP, 700)
    cache = {}

    def cached_inputs(poc):          # the partition generator is host Python: not part of what is checked
        if poc not in cache:
            cache[poc] = inputs(poc)
        return cache[poc]
    pocs = [0] + [p[0] for p in common.GOP8]
    ctx = lib.Context(W, H, BD, num_slots=3 + len(pocs), device=dev)
    peers = sharding.PeerExchange(ctx, dist, rank, world)
    eng = sharding.GpuGopEngine(ctx, peers, rank, pocs, cached_inputs)
    eng.load_done(0, frame(0))
    for p in common.GOP8:
        cached_inputs(p[0])
    ctx.sync()
    dist.barrier()
    t0 = time.perf_counter()
    owners = sharding.FrameParallelGop(eng, rank, world).encode(common.GOP8, done=(0,))
    dt = time.perf_counter() - t0
    got = {poc: [ctx.download_padded(eng.slot_of[poc], c) for c in range(3)] for poc in pocs[1:]}
    # serial encode of the same chain on this rank's GPU, in a context of its own
    ctx1 = lib.Context(W, H, BD, num_slots=3 + len(pocs), device=dev)
    one = sharding.GpuGopEngine(ctx1, None, 0, pocs, cached_inputs)
    one.load_done(0, frame(0))
    for p in common.GOP8:
        one.encode(*p)
    ctx1.sync()
    ok = True
    for poc in pocs[1:]:
        for c in range(3):
            ok &= bool(np.array_equal(got[poc][c], ctx1.download_padded(one.slot_of[poc], c)))
    t = torch.tensor([1 if ok else 0])
    if not same_gpu:
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("frame-parallel sub-GOP of 8 at %dx%d on %d ranks (%s), pictures per rank %s, %.1f ms: every rank holds all reconstructions == serial encode: %s"
              % (W, H, world, "one GPU" if same_gpu else "NVLink pushes", [list(owners.values()).count(r) for r in range(world)], dt * 1e3, bool(t.item())))
    dist.destroy_process_group()
    sys.exit(0 if t.item() else 1)


if __name__ == "__main__":
    main()
