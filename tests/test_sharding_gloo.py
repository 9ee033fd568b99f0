"""world_size-2 (and 3) gloo runs on CPU of the N>1 host logic: CTU-row banded encode with the
deblocking halo exchange must reproduce the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from oracle.bindings import Oracle, Picture
from xvc_b200 import sharding, workload

W, H, BD, QP = 200, 136, 10, 32


def _inputs():
    cur, r0, r1 = common.frames(W, H, BD, 301)
    cus = workload.make_partition(W, H, seed=31, min_size=4, qp=QP)
    prm = common.picture_params(0, workload.lambda_for_qp(QP), ranges=(96, 96))
    prm["pad"] = 0
    return cur, r0, r1, cus, prm


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    cur, r0, r1, cus, prm = _inputs()
    eng = OracleEngine(W, H, BD, cur, {(0, 0): r0, (1, 0): r1}, {(0, 0): 0, (1, 0): 16})
    enc = sharding.BandedPictureEncoder(eng, dist, rank, world, H)
    full = enc.encode(cus, prm)
    y0, y1 = enc.bands[rank]
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), cus=full.view(np.uint8), y0=y0, y1=y1,
             **{"p%d" % c: enc.gather_band_rows(c).numpy().view(np.uint16) for c in range(3)})
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_banded_encode_matches_single_process(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    cur, r0, r1, cus, prm = _inputs()
    o = Oracle()
    pred, rec = Picture(W, H, 80), Picture(W, H, 80)
    cus_ref = cus.copy()
    o.encode_picture(Picture(W, H, 0, cur), common.oracle_refs(o, W, H, r0, r1), pred, rec, BD, cus_ref, prm)
    planes = [np.zeros((H >> (1 if c else 0), W >> (1 if c else 0)), dtype=np.uint16) for c in range(3)]
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        y0, y1 = int(z["y0"]), int(z["y1"])
        for c in range(3):
            s = 1 if c else 0
            planes[c][y0 >> s:y1 >> s] = z["p%d" % c]
        got = z["cus"].view(cus.dtype)
        for f in ("flags", "ref_idx", "mv"):
            assert np.array_equal(got[f], cus_ref[f]), (r, f)
    for c in range(3):
        assert np.array_equal(planes[c], rec.plane(c)), c


def test_band_rows_cover_picture():
    for h in (64, 72, 136, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            bands = sharding.band_rows(h, world)
            assert bands[0][0] == 0 and bands[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            assert all(y0 % 64 == 0 or y0 == h for y0, _ in bands)


# ---------------------------------------------------------------- frame-parallel GOP (config 5)
GW, GH = 136, 72


def test_gop_waves():
    waves = sharding.gop_waves(common.GOP8, done=(0,))
    assert waves == [[8], [4], [2, 6], [1, 3, 5, 7]]
    sub16 = [(16, 1, (0,)), (8, 0, (0, 16)), (4, 0, (0, 8)), (12, 0, (8, 16))] + \
            [(p, 0, (p - 2, p + 2)) for p in (2, 6, 10, 14)] + [(p, 0, (p - 1, p + 1)) for p in range(1, 16, 2)]
    assert [len(w) for w in sharding.gop_waves(sub16, done=(0,))] == [1, 1, 2, 4, 8]
    with pytest.raises(ValueError):
        sharding.gop_waves([(4, 0, (0, 8))], done=(0,))


def _gop_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleGopEngine
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    frame, inputs = common.gop_inputs(GW, GH, BD, QP, 500)
    eng = OracleGopEngine(GW, GH, BD, dist, inputs)
    eng.load_done(0, frame(0))
    owners = sharding.FrameParallelGop(eng, rank, world).encode(common.GOP8, done=(0,))
    assert sorted(set(owners.values())) == list(range(min(world, 4)))
    np.savez(os.path.join(out_dir, "gop%d.npz" % rank), **{"p%d_%d" % (poc, c): eng.rec[poc].full[c] for poc in eng.rec for c in range(3)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_frame_parallel_gop_matches_single_process(tmp_path, world):
    """Every rank ends up with every reconstruction of the sub-GOP, equal to the serial encode."""
    mp.spawn(_gop_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle_engine import OracleGopEngine
    frame, inputs = common.gop_inputs(GW, GH, BD, QP, 500)
    eng = OracleGopEngine(GW, GH, BD, None, inputs)
    eng.load_done(0, frame(0))
    for poc, pic_type, refs in common.GOP8:
        eng.encode(poc, pic_type, refs)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "gop%d.npz" % r))
        for poc in eng.rec:
            for c in range(3):
                assert np.array_equal(z["p%d_%d" % (poc, c)], eng.rec[poc].full[c]), (r, poc, c)


# ---------------------------------------------------------------- peer exchange: agreed fallback
class _FakeCtx:
    """Stands in for lib.Context: rank `bad` cannot open its peers' arenas."""

    def __init__(self, rank, bad):
        self.rank, self.bad, self.opened = rank, bad, 0

    def ipc_export(self):
        return bytes([self.rank]) * 80

    def ipc_open_peer(self, handle):
        if self.rank == self.bad:
            raise RuntimeError("cudaIpcOpenMemHandle: invalid device context")
        self.opened += 1


def _peer_worker(rank, world, port, bad, out_dir):
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ctx = _FakeCtx(rank, bad)
    try:
        sharding.PeerExchange(ctx, dist, rank, world)
        outcome = "ok %d" % ctx.opened
    except sharding.PeerExchangeUnavailable as e:
        outcome = "unavailable: %s" % e
    open(os.path.join(out_dir, "peer%d.txt" % rank), "w").write(outcome)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("bad", [-1, 1])
def test_peer_exchange_all_ranks_agree(tmp_path, bad):
    """Either every rank opens every peer, or EVERY rank gets PeerExchangeUnavailable (naming the rank that
    failed) -- so that bench.py / an encoder falls back to the NCCL exchange on all ranks together."""
    world = 3
    mp.spawn(_peer_worker, args=(world, _free_port(), bad, str(tmp_path)), nprocs=world, join=True)
    got = [open(os.path.join(str(tmp_path), "peer%d.txt" % r)).read() for r in range(world)]
    if bad < 0:
        assert got == ["ok 2"] * world
    else:
        assert all(g.startswith("unavailable: rank %d" % bad) for g in got), got


def test_gop_arrival_tags_increase_per_slot():
    """Device-side rendezvous of the frame-parallel GOP (xvcb200_push_slot_tagged / xvcb200_wait_slot_tag): the
    consumer waits for `tag >=`, so the tags written into one slot's arrival word must grow from push to push --
    within a pass (ring slots reused by later POCs) and from the warm-up pass to the timed one."""
    from xvc_b200 import gop

    class NoCtx:
        geom = {"margin_y": [0, 0, 0], "margin_x": [0, 0, 0], "height": [8, 4, 4], "width": [8, 4, 4]}

        def plane_tensor(self, slot, comp):
            return None

    for n_sub in (2, 4, 8):
        pics = gop.hierarchical_gop(n_sub)
        eng = gop.GopEngine(NoCtx(), None, 0, pics, None, 32, 10)
        last = {}
        for pass_index in (0, 1):
            eng.pass_index = pass_index
            for poc, _, _, _ in pics:                      # coding order = push order of a slot's producers
                slot, tag = eng.slot_of(poc), eng.tag_of(poc)
                assert tag > last.get(slot, 0), (n_sub, poc, slot)
                last[slot] = tag
        # every reference picture of a picture is coded before it (the wait can be satisfied)
        order = {p[0]: k for k, p in enumerate(pics)}
        for poc, _, l0, l1 in pics:
            assert all(r == 0 or order[r] < order[poc] for r in tuple(l0) + tuple(l1))
