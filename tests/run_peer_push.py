#!/usr/bin/env python3
"""Frame-parallel peer exchange check (launched by tests/test_gpu_configs.py and by hand):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/run_peer_push.py [same-gpu]
Every rank uploads + pads its own picture into slot first+rank and pushes it to the peers
(xvcb200_push_slot: CUDA IPC + copy engines); afterwards every rank must hold every rank's
padded picture bit for bit.  `same-gpu`: all ranks on cuda:0 (gloo rendezvous) -- the IPC path
of a one-GPU box; otherwise one rank per GPU over NVLink (NCCL rendezvous)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from xvc_b200 import lib, sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    same_gpu = len(sys.argv) > 1 and sys.argv[1] == "same-gpu"
    dev = 0 if same_gpu else local
    torch.cuda.set_device(dev)
    if same_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    W, H, BD = 832, 480, 10
    first = 1
    ctx = lib.Context(W, H, BD, num_slots=first + world, device=dev)
    pics = [common.frames(W, H, BD, 100 + r)[0] for r in range(world)]
    ctx.upload(first + rank, pics[rank])
    ctx.pad_border(first + rank)
    ex = sharding.PeerExchange(ctx, dist, rank, world)
    for rounds in range(2):           # second round: slots rewritten after wait_own()
        ex.push(first + rank)
        ex.landed()
    ok = True
    probe = lib.Context(W, H, BD, num_slots=1, device=dev)
    for r in range(world):
        probe.upload(0, pics[r])
        probe.pad_border(0)
        for c in range(3):
            ok &= bool(np.array_equal(ctx.download_padded(first + r, c), probe.download_padded(0, c)))
    # The rendezvous on the device (xvcb200_push_slot_tagged / xvcb200_wait_slot_tag): new content, no host barrier
    # between the push and its consumption -- the consumer's stream waits for the arrival tag written behind the slot.
    for tag in (1, 2):
        pics2 = [common.frames(W, H, BD, 200 + 10 * tag + r)[0] for r in range(world)]
        ex.wait_own(first + rank)
        ctx.upload(first + rank, pics2[rank])
        ctx.pad_border(first + rank)
        ex.push_tagged(first + rank, tag)
        for r in range(world):
            if r != rank:
                ex.wait_tag(first + r, tag)
            probe.upload(0, pics2[r])
            probe.pad_border(0)
            for c in range(3):
                ok &= bool(np.array_equal(ctx.download_padded(first + r, c), probe.download_padded(0, c)))
        ex.landed()          # (before the slots are rewritten by the next round)
    t = torch.tensor([1 if ok else 0])
    if not same_gpu:
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("peer push of padded reconstructions between %d ranks (%s): %s" % (world, "one GPU" if same_gpu else "NVLink", bool(t.item())))
    dist.destroy_process_group()
    sys.exit(0 if t.item() else 1)


if __name__ == "__main__":
    main()
