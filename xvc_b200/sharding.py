"""Multi-GPU decomposition of the hot path (one process per GPU, torch.distributed).

xvc itself has exactly one parallel mechanism: picture-level worker threads, a picture being
runnable once its reference pictures are finished (thread_encoder.cc:99-159).  Two
decompositions follow from the data flow of the hot path:

* frame_parallel_exchange(): every rank encodes its own picture; the only data-path exchange
  is the finished, padded reconstruction that later pictures reference -- one NCCL all-gather
  straight into reference slots (slots of a context are contiguous in memory).

* BandedPictureEncoder: ONE picture split into bands of CTU rows.  ME / MC / T-Q-recon of a
  band need nothing from other bands (reference pictures are replicated).  Deblocking needs
  one exchange: vertical edges are row-local; the horizontal edge ON a band boundary belongs to
  the lower band (its q side), reads 4 and writes 3 luma rows (2 / 1 chroma rows) of the upper
  band, and must run after the upper band finished its own edges (order dependence of edges 4
  rows apart, deblocking_filter.cc:56-77).  Protocol per boundary: upper -> lower 4 luma + 2x2
  chroma rows after the upper band's passes; lower filters, returns 3 luma + 2x1 chroma rows.
  CU metadata (cbf flags, chosen MVs) of all bands is all-gathered first because boundary
  strength looks at both sides of an edge.

The classes take an `engine` (the compute backend) so the host logic can be exercised on CPU
with gloo in tests; GpuEngine below is the product engine (libxvc_b200 through lib.Context).
"""
import numpy as np

from . import abi


def band_rows(height, world):
    """Contiguous bands of 64-row CTU rows, as even as possible: [(y0, y1), ...]."""
    ctu_rows = (height + 63) // 64
    out, start = [], 0
    for r in range(world):
        n = ctu_rows // world + (1 if r < ctu_rows % world else 0)
        out.append((min(height, 64 * start), min(height, 64 * (start + n))))
        start += n
    return out


def cus_in_band(cus, y0, y1):
    return cus[(cus["y"] >= y0) & (cus["y"] < y1)]


class GpuEngine:
    """Compute backend on one B200: a lib.Context whose stream is torch's current stream."""

    def __init__(self, ctx, slots, bitdepth, ref_poc):
        import torch
        self.torch = torch
        self.ctx, self.slots, self.bitdepth, self.ref_poc = ctx, slots, bitdepth, ref_poc
        # one non-default stream for the library's kernels, torch copies and the NCCL p2p ops
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        ctx.set_stream(self.stream.cuda_stream)
        self.device = torch.device("cuda", torch.cuda.current_device())

    def encode_band(self, cus_band, prm):
        p = prm.copy()
        p["deblock"], p["pad"] = 0, 0
        self.ctx.set_cus(cus_band)
        self.ctx.encode_picture(p, want_results=False)
        return self.ctx.get_cus()

    def set_cus(self, cus):
        self.ctx.set_cus(cus)

    def deblock(self, pic_type, pass_mask, y0, y1):
        self.ctx.deblock_band(self.slots["rec"], pic_type, self.ref_poc, pass_mask, y0, y1)

    def _rows(self, comp, y0, y1):
        g = self.ctx.geom
        t = self.ctx.plane_tensor(self.slots["rec"], comp)
        my, mx, w = int(g["margin_y"][comp]), int(g["margin_x"][comp]), int(g["width"][comp])
        return t[my + y0:my + y1, mx:mx + w]

    # rows travel as bytes: NCCL has no 16-bit integer type
    def get_rows(self, comp, y0, y1):
        return self._rows(comp, y0, y1).contiguous().view(self.torch.uint8)

    def put_rows(self, comp, y0, rows):
        self._rows(comp, y0, y0 + rows.shape[0]).copy_(rows.view(self.torch.int16))

    def empty_rows(self, comp, n):
        return self.torch.empty((n, 2 * int(self.ctx.geom["width"][comp])), dtype=self.torch.uint8, device=self.device)

    def to_comm(self, arr_u8):
        return self.torch.from_numpy(arr_u8).to(self.device)

    def from_comm(self, t):
        return t.cpu().numpy()

    def finish(self):
        self.ctx.sync()


class BandedPictureEncoder:
    """One picture across `world` ranks by CTU-row bands (see module docstring)."""

    def __init__(self, engine, dist, rank, world, height):
        self.e, self.dist, self.rank, self.world = engine, dist, rank, world
        self.bands = band_rows(height, world)

    def _gather_cus(self, mine):
        """all-gather of variable-length CU arrays -> the full CU list in picture order."""
        if self.world == 1:
            return mine
        torch = __import__("torch")
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(self.world)]
        n_local = self.e.to_comm(np.array([len(mine)], dtype=np.int64).view(np.uint8))
        all_n = [self.e.to_comm(np.zeros(8, dtype=np.uint8)) for _ in range(self.world)]
        self.dist.all_gather(all_n, n_local)
        lens = [int(self.e.from_comm(t).view(np.int64)[0]) for t in all_n]
        cap = max(lens) * abi.cu_dtype.itemsize
        buf = np.zeros(cap, dtype=np.uint8)
        buf[:mine.nbytes] = mine.view(np.uint8)
        outs = [self.e.to_comm(np.zeros(cap, dtype=np.uint8)) for _ in range(self.world)]
        self.dist.all_gather(outs, self.e.to_comm(buf))
        parts = [self.e.from_comm(t)[:n * abi.cu_dtype.itemsize].view(abi.cu_dtype) for t, n in zip(outs, lens)]
        del counts
        return np.concatenate(parts)

    def encode(self, cus, prm):
        """cus: the CU list of the WHOLE picture (every rank passes the same); returns the full
        CU list after the decisions of all bands.  The rank's band of the reconstruction (plus
        the rows its neighbours returned) is final when this returns."""
        e, r = self.e, self.rank
        y0, y1 = self.bands[r]
        pic_type = int(prm["pic_type"][0])
        mine = e.encode_band(cus_in_band(cus, y0, y1), prm) if y1 > y0 else cus[:0].copy()
        full = self._gather_cus(mine)
        if not int(prm["deblock"][0]):
            e.finish()
            return full
        e.set_cus(full)
        if y1 > y0:
            e.deblock(pic_type, 1, y0, y1)                   # vertical edges: row-local
        up = r - 1 if r > 0 and y1 > y0 and y0 > 0 else None
        down = r + 1 if r + 1 < self.world and self.bands[r + 1][1] > self.bands[r + 1][0] and y1 > y0 else None
        if up is not None:                                    # halo from the band above (its passes are done)
            for comp, n in ((0, 4), (1, 2), (2, 2)):
                t = e.empty_rows(comp, n)
                self.dist.recv(t, src=up)
                e.put_rows(comp, (y0 >> (1 if comp else 0)) - n, t)
        if y1 > y0:
            e.deblock(pic_type, 2, y0, y1)                   # horizontal edges incl. the one ON y0
        if up is not None:                                    # rows the boundary edge modified go back
            for comp, n in ((0, 3), (1, 1), (2, 1)):
                yb = y0 >> (1 if comp else 0)
                self.dist.send(e.get_rows(comp, yb - n, yb), dst=up)
        if down is not None:
            for comp, n in ((0, 4), (1, 2), (2, 2)):
                yb = y1 >> (1 if comp else 0)
                self.dist.send(e.get_rows(comp, yb - n, yb), dst=down)
            for comp, n in ((0, 3), (1, 1), (2, 1)):
                t = e.empty_rows(comp, n)
                self.dist.recv(t, src=down)
                e.put_rows(comp, (y1 >> (1 if comp else 0)) - n, t)
        e.finish()
        return full

    def gather_band_rows(self, comp):
        """Final reconstruction rows of this rank's band (for assembling the picture)."""
        y0, y1 = self.bands[self.rank]
        s = 1 if comp else 0
        return self.e.get_rows(comp, y0 >> s, y1 >> s)


def frame_parallel_exchange(ctx, dist, rank, world, first_slot, async_op=False):
    """After every rank reconstructed (and padded) its own picture into slot first_slot+rank:
    one in-place all-gather makes slots [first_slot, first_slot+world) hold all of them on every
    rank, ready to be used as reference slots.  async_op=True: the collective runs on NCCL's own
    stream behind the kernels enqueued so far and the returned work's wait() orders a later
    consumer (or the reuse of these slots) behind it -- the picture encoded next does not
    depend on it (pictures of one temporal layer, thread_encoder.cc:99-131)."""
    if world == 1:
        return None
    whole = ctx.slots_tensor(first_slot, world)
    mine = ctx.slots_tensor(first_slot + rank, 1)
    return dist.all_gather_into_tensor(whole, mine, async_op=async_op)


class PeerExchangeUnavailable(RuntimeError):
    """Raised on EVERY rank when some rank could not open a peer arena."""


class PeerExchange:
    """Frame-parallel exchange without SMs: every rank pushes its finished, padded
    reconstruction slot into the SAME slot of every other rank's arena with the copy engines
    over NVLink (xvcb200_push_slot; arenas opened through CUDA IPC).  An NCCL all-gather needs
    SMs for its copy kernels, which the persistent search kernel of the next picture occupies;
    DMA pushes overlap it completely.  All contexts must have the same geometry / slot count.

    push(slot) is ordered after the work enqueued on the context stream so far.  landed() is
    the consumer-side guarantee: own pushes finished (context stream waits for them) and a
    rendezvous of all ranks, after which slots pushed by the others may be referenced (or a
    pushed slot overwritten)."""

    def __init__(self, ctx, dist, rank, world):
        self.ctx, self.dist, self.rank, self.world = ctx, dist, rank, world
        handles = [None] * world
        dist.all_gather_object(handles, ctx.ipc_export())
        err = None
        try:
            for r in range(world):
                if r != rank:
                    ctx.ipc_open_peer(handles[r])
        except RuntimeError as e:           # e.g. no peer access between two GPUs of this box
            err = repr(e)
        errs = [None] * world
        dist.all_gather_object(errs, err)   # all ranks agree: usable everywhere or nowhere
        bad = [(r, e) for r, e in enumerate(errs) if e]
        if bad:
            raise PeerExchangeUnavailable("rank %d: %s" % bad[0])

    def push(self, slot):
        self.ctx.push_slot(slot)

    def wait_own(self, slot=-1):
        """Context stream waits for this rank's last push of `slot` (-1: all) -- before the slot
        is rewritten."""
        self.ctx.wait_pushes(slot)

    def landed(self):
        self.ctx.wait_pushes()
        self.ctx.sync()
        self.dist.barrier()

    # The same rendezvous on the device: the producer writes an arrival tag behind the pushed slot (same copy
    # stream), the consumer's context stream waits for the tag -- no host, no barrier between waves.
    def push_tagged(self, slot, tag):
        self.ctx.push_slot_tagged(slot, tag)

    def wait_tag(self, slot, tag):
        self.ctx.wait_slot_tag(slot, tag)


# ---------------------------------------------------------------------------------------------
# Frame-parallel encoding of a GOP (BASELINE config 5): the reference's own parallel model.
# ThreadEncoder (thread_encoder.cc:99-159) starts a picture as soon as all of its reference
# pictures are finished, lowest temporal layer first.  With one encoder per GPU the schedule is
# computed identically on every rank; a finished (deblocked, padded) reconstruction is pushed to
# every other GPU, where it lands in the slot the same POC has everywhere.
# ---------------------------------------------------------------------------------------------
def gop_waves(pictures, done=()):
    """pictures: [(poc, pic_type, ref_pocs), ...] in coding order; done: POCs reconstructed already.
    Returns the waves of mutually independent pictures: a picture belongs to the first wave after
    the waves of all of its reference pictures (a hierarchical-B sub-GOP of 16: 1, 1, 2, 4, 8)."""
    wave_of = {p: -1 for p in done}
    waves = []
    for poc, _, refs in pictures:
        missing = [r for r in refs if r not in wave_of]
        if missing:
            raise ValueError("picture %d references %s before it is coded" % (poc, missing))
        k = max([wave_of[r] for r in refs], default=-1) + 1
        wave_of[poc] = k
        while len(waves) <= k:
            waves.append([])
        waves[k].append(poc)
    return waves


class FrameParallelGop:
    """Runs gop_waves() across `world` ranks: picture j of a wave is encoded by rank j % world; after
    the wave every reconstruction is shared with all ranks (engine.share) and the ranks rendezvous
    (engine.fence) before a later wave may reference it.  `engine` does the work of one rank:
    sharding.GpuGopEngine (libxvc_b200 + PeerExchange) or the oracle engine of the gloo tests."""

    def __init__(self, engine, rank, world):
        self.e, self.rank, self.world = engine, rank, world

    def encode(self, pictures, done=()):
        by_poc = {p[0]: p for p in pictures}
        owners = {}
        for wave in gop_waves(pictures, done):
            for j, poc in enumerate(wave):
                owners[poc] = j % self.world
                if owners[poc] == self.rank:
                    self.e.encode(*by_poc[poc])
            for poc in wave:
                self.e.share(poc, owners[poc])
            self.e.fence()
        return owners


class GpuGopEngine:
    """One rank of FrameParallelGop on a B200.  Slots: 0 original, 1 prediction, 2 levels, then one
    reconstruction slot per POC (the same slot index on every rank, so a push lands in place).
    inputs(poc) -> (planes of the original picture, CU array, xvcb200_picture_params without slots)."""

    def __init__(self, ctx, peers, rank, pocs, inputs):
        self.ctx, self.peers, self.rank, self.inputs = ctx, peers, rank, inputs
        self.slot_of = {poc: 3 + i for i, poc in enumerate(pocs)}

    def load_done(self, poc, planes):
        self.ctx.upload(self.slot_of[poc], planes)
        self.ctx.pad_border(self.slot_of[poc])

    def encode(self, poc, pic_type, ref_pocs):
        cur, cus, prm = self.inputs(poc)
        prm = prm.copy()
        prm["orig_slot"], prm["pred_slot"], prm["coeff_slot"], prm["rec_slot"] = 0, 1, 2, self.slot_of[poc]
        prm["ref_slots"] = -1
        for l, r in enumerate(ref_pocs):
            prm["ref_slots"][0, l, 0] = self.slot_of[r]
        self.ctx.upload(0, cur)
        self.ctx.set_cus(cus)
        self.ctx.encode_picture(prm, want_results=False)

    def share(self, poc, owner):
        if owner == self.rank and self.peers is not None:
            self.peers.push(self.slot_of[poc])

    def fence(self):
        if self.peers is not None:
            self.peers.landed()
        else:
            self.ctx.sync()
