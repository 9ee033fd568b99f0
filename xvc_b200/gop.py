"""Frame-parallel encoding of a hierarchical-B sequence on one or more GPUs (BASELINE config 5).

The reference's parallel model is ThreadEncoder (thread_encoder.cc:99-159): a picture may start once all of its
reference pictures are finished, lowest temporal layer first.  sharding.gop_waves / FrameParallelGop compute that
schedule identically on every rank; this module supplies the sequence (xvc's hierarchical-B sub-GOPs of 16 after a
key picture) and GopEngine, one rank's worker: per picture the GPU partition pre-analysis
(xvcb200_decide_partition), predictors scaled by POC distance, xvcb200_encode_picture, and the push of the finished,
padded reconstruction into the slot the same POC has on every GPU.  Waves span sub-GOPs (a picture of sub-GOP k+1
whose references are done runs beside the deepest layer of sub-GOP k), so after the ramp a wave holds 16 pictures.
"""
import numpy as np

from . import abi, workload

SUB_GOP = 16


def hierarchical_gop(n_sub_gops, sub_gop=SUB_GOP):
    """Coding order of n_sub_gops sub-GOPs after key picture 0: [(poc, pic_type, (L0 POCs), (L1 POCs))].
    Anchor pictures (POC multiple of sub_gop) are uni-predicted from the previous anchor; a B picture of
    temporal distance s references POC -s and POC +s, two pictures per list, list 1 in the opposite order
    (what ReferenceListSorter gives xvc's default two reference pictures per list)."""
    pics = []
    for g in range(n_sub_gops):
        base = g * sub_gop
        pics.append((base + sub_gop, 1, (base,), ()))
        step = sub_gop // 2
        while step >= 1:
            for p in range(base + step, base + sub_gop, 2 * step):
                pics.append((p, 0, (p - step, p + step), (p + step, p - step)))
            step //= 2
    return pics


def seq_index(poc, span=31):
    """Frame of the synthetic content shown at POC `poc`: the pan runs forward for `span` frames, then backward
    (the canvas is finite), so that long sequences have no jump in the content."""
    k = poc % (2 * span)
    return k if k <= span else 2 * span - k


def as_wave_input(pics):
    """-> the (poc, pic_type, reference POCs) tuples sharding.gop_waves takes."""
    return [(poc, t, tuple(sorted(set(l0) | set(l1)))) for poc, t, l0, l1 in pics]


class GopEngine:
    """One rank of sharding.FrameParallelGop on a B200.  Slots: 0 and 3 originals (alternating: the pre-analysis of a
    rank's next picture runs ahead of the kernels of its current one), 1 prediction, 2 levels, then a ring of
    reconstruction slots indexed by POC (the same slot on every rank, so a pushed reconstruction lands in place; the
    ring is longer than the span of POCs alive at any time).  originals(poc) -> three device tensors (int16 views of
    the tight planes) resident before the timed region."""
    FIRST_RING_SLOT = 4
    ORIG_SLOTS = (0, 3)

    def __init__(self, ctx, peers, rank, pics, originals, qp, bitdepth, ring=56, max_range=128, time_events=None, owners=None):
        self.ctx, self.peers, self.rank, self.originals, self.qp = ctx, peers, rank, originals, qp
        self.by_poc = {p[0]: p for p in pics}
        # device-side rendezvous (owners: POC -> rank that codes it): a consumer's stream waits for the arrival tag
        # of every reference picture another rank produced; tags grow with the coding order and from pass to pass
        self.owners = owners
        self.order = {p[0]: k + 1 for k, p in enumerate(pics)}
        self.pass_index = 0
        self.ring, self.max_range = ring, max_range
        self.lam = workload.lambda_for_qp(qp)
        self.busy = time_events            # callable returning a (start, stop) pair of recorded-on-demand events, or None
        self.events = []
        g = ctx.geom
        self.views = [(int(g["margin_y"][c]), int(g["margin_x"][c]), int(g["height"][c]), int(g["width"][c])) for c in range(3)]
        self.orig_t = {o: [ctx.plane_tensor(o, c) for c in range(3)] for o in self.ORIG_SLOTS}

    def slot_of(self, poc):
        return self.FIRST_RING_SLOT + poc % self.ring

    def load_done(self, poc, planes):
        self.ctx.upload(self.slot_of(poc), planes)
        self.ctx.pad_border(self.slot_of(poc))

    def params(self, poc):
        _, pic_type, l0, l1 = self.by_poc[poc]
        prm = np.zeros(1, dtype=abi.picture_params_dtype)
        prm["pic_type"], prm["lambda_sqrt"], prm["chroma_offset_table"] = pic_type, np.sqrt(self.lam), 1
        prm["deblock"], prm["pad"] = 1, 1
        prm["bi_iterations"], prm["bits_mode"] = (1 if pic_type == 0 else 0), 1
        prm["orig_slot"], prm["pred_slot"], prm["coeff_slot"], prm["rec_slot"] = 0, 1, 2, self.slot_of(poc)
        prm["ref_slots"] = -1
        for l, pocs in enumerate((l0, l1)):
            prm["num_ref"][0, l] = len(pocs)
            for r, p in enumerate(pocs):
                prm["ref_slots"][0, l, r] = self.slot_of(p)
                prm["ref_poc"][0, l, r] = p
                # InterSearch::GetSearchRangeUniPred gives 256 for the anchor pictures; the search kernel stages
                # +-128 windows in shared memory, so the range is capped there (stated in the bench line)
                prm["search_range"][0, l, r] = min(self.max_range, workload.search_range_uni(poc, p))
        return prm, l0, l1

    def tag_of(self, poc):
        return self.pass_index * (len(self.order) + 1) + self.order[poc]

    def _begin(self, poc, oslot):
        """Original picture -> slot oslot (device copy), pre-analysis enqueued: the partition from the content, against
        the first list-0 picture, around the sequence's global motion."""
        _, _, l0, l1 = self.by_poc[poc]
        if self.owners is not None and self.peers is not None:
            for p in sorted(set(l0) | set(l1)):
                if p in self.owners and self.owners[p] != self.rank:      # (the key picture is resident everywhere)
                    self.peers.wait_tag(self.slot_of(p), self.tag_of(p))
        for c, t in enumerate(self.originals(poc)):
            my, mx, h, w = self.views[c]
            self.orig_t[oslot][c][my:my + h, mx:mx + w].copy_(t)
        center = workload.true_motion(seq_index(poc), seq_index(l0[0]))
        self.ctx.decide_partition_begin(oslot, self.slot_of(l0[0]), float(np.sqrt(self.lam)), self.qp, center=center)

    def _end(self, poc):
        """-> (CU array, predictors per (CU, list, reference picture): the pre-analysis vector scaled by POC distance)."""
        _, _, l0, l1 = self.by_poc[poc]
        cus, _ = self.ctx.decide_partition_end()
        mv0 = cus["mv"][:, 0, :].astype(np.int64)
        d0 = poc - l0[0]
        cols = [(mv0 * (poc - p)) // d0 for p in tuple(l0) + tuple(l1)]
        return cus, np.ascontiguousarray(np.stack(cols, axis=1).astype(np.int32))

    def encode_many(self, pocs):
        """This rank's pictures of one wave (independent of each other), pipelined: the pre-analysis kernel of picture
        k+1 is enqueued ahead of the kernels of picture k, so the host turns its result into the CU array and the
        predictors while picture k runs (the bench step's scheme)."""
        ctx = self.ctx
        if not pocs:
            return
        self._begin(pocs[0], self.ORIG_SLOTS[0])
        nxt = self._end(pocs[0])
        for i, poc in enumerate(pocs):
            oslot = self.ORIG_SLOTS[i & 1]
            prm, _, _ = self.params(poc)
            prm["orig_slot"] = oslot
            ev = self.busy() if self.busy else None
            cus, mvp = nxt
            ctx.set_cus(cus)
            ctx.set_mv_predictors(mvp)
            if ev:
                ev[0].record()
            if i + 1 < len(pocs):
                self._begin(pocs[i + 1], self.ORIG_SLOTS[(i + 1) & 1])
            ctx.encode_picture(prm, want_results=False)
            if ev:
                ev[1].record()
                self.events.append((poc, ev[0], ev[1]))
            if i + 1 < len(pocs):
                nxt = self._end(pocs[i + 1])

    def encode(self, poc, pic_type=None, ref_pocs=None):
        self.encode_many([poc])

    def share(self, poc, owner):
        if owner == self.rank and self.peers is not None:
            if self.owners is not None:
                self.peers.push_tagged(self.slot_of(poc), self.tag_of(poc))
            else:
                self.peers.push(self.slot_of(poc))

    def fence(self):
        if self.peers is not None:
            self.peers.landed()
        else:
            self.ctx.sync()
