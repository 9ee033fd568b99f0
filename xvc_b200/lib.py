"""ctypes binding of libxvc_b200.so (include/xvc_b200.h).

The Python side only marshals numpy arrays into the C ABI; every sample is produced by
the CUDA kernels.  There is no CPU fallback: loading fails loudly when the shared library
has not been built, and every call fails loudly when no CUDA device is present.
"""
import ctypes
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxvc_b200.so")
IPC_HANDLE_BYTES = 80      # XVCB200_IPC_HANDLE_BYTES

c_int, c_void_p, c_double, c_u64 = ctypes.c_int, ctypes.c_void_p, ctypes.c_double, ctypes.c_uint64
c_ssize = ctypes.c_ssize_t

# every symbol include/xvc_b200.h declares (tests/test_abi.py checks they are all exported)
EXPORTS = [
    "xvcb200_last_error", "xvcb200_last_error_string", "xvcb200_clear_error", "xvcb200_launch_count",
    "xvcb200_version", "xvcb200_abi_sizeof",
    "xvcb200_sad_sample_sample", "xvcb200_sad_short_sample", "xvcb200_ssd_sample_sample",
    "xvcb200_ssd_short_sample", "xvcb200_ssd_short_short", "xvcb200_compare_sample_sample",
    "xvcb200_compare_short_sample",
    "xvcb200_filter_h_sample_sample", "xvcb200_filter_h_sample_short", "xvcb200_filter_v_sample_sample",
    "xvcb200_filter_v_sample_short", "xvcb200_filter_v_short_sample", "xvcb200_filter_v_short_short",
    "xvcb200_add_avg", "xvcb200_filter_copy_bipred", "xvcb200_interp_block", "xvcb200_interp_block_bipred",
    "xvcb200_register_inter_prediction", "xvcb200_register_sample_metric",
    "xvcb200_fwd_transform", "xvcb200_fwd_transform_skip", "xvcb200_inv_transform", "xvcb200_inv_transform_skip",
    "xvcb200_qp_init", "xvcb200_quant_fast", "xvcb200_dequant",
    "xvcb200_ctx_create", "xvcb200_ctx_destroy", "xvcb200_ctx_set_stream", "xvcb200_stream", "xvcb200_sync",
    "xvcb200_ctx_error_string", "xvcb200_get_geometry", "xvcb200_slot_ptr", "xvcb200_slot_region",
    "xvcb200_upload_picture", "xvcb200_download_picture", "xvcb200_download_coeff", "xvcb200_upload_coeff",
    "xvcb200_download_padded", "xvcb200_pad_border", "xvcb200_set_cus", "xvcb200_get_cus", "xvcb200_set_mv_predictors", "xvcb200_set_tu_modes",
    "xvcb200_upload_picture_async", "xvcb200_download_picture_async", "xvcb200_download_coeff_async",
    "xvcb200_get_cus_async", "xvcb200_sync_copies", "xvcb200_wait_download",
    "xvcb200_me_search", "xvcb200_full_search", "xvcb200_motion_compensate", "xvcb200_motion_compensate_affine", "xvcb200_motion_compensate_lic", "xvcb200_tq_reconstruct",
    "xvcb200_dequant_reconstruct", "xvcb200_deblock_picture", "xvcb200_deblock_picture_ex", "xvcb200_deblock_picture_ext", "xvcb200_deblock_band",
    "xvcb200_encode_picture", "xvcb200_decide_partition", "xvcb200_decide_partition_begin", "xvcb200_decide_partition_end", "xvcb200_set_profiling", "xvcb200_get_stage_times",
    "xvcb200_intra_ref_samples", "xvcb200_intra_predict", "xvcb200_intra_satd_scan", "xvcb200_intra_lm_chroma",
    "xvcb200_ipc_export", "xvcb200_ipc_open_peer", "xvcb200_push_slot", "xvcb200_wait_pushes",
    "xvcb200_push_slot_tagged", "xvcb200_wait_slot_tag",
    "xvcb200_device_count",
]


class XvcB200Error(RuntimeError):
    pass


_lib = None


def device_count():
    return int(load().xvcb200_device_count())


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XvcB200Error("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(xvc_b200 has no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.xvcb200_last_error_string.restype = ctypes.c_char_p
    L.xvcb200_version.restype = ctypes.c_char_p
    L.xvcb200_ctx_error_string.restype = ctypes.c_char_p
    L.xvcb200_ctx_error_string.argtypes = [c_void_p]
    L.xvcb200_launch_count.restype = c_u64
    L.xvcb200_stream.restype = c_void_p
    L.xvcb200_stream.argtypes = [c_void_p]
    blk = [c_int, c_int, c_void_p, c_ssize, c_void_p, c_ssize]
    for name in ("xvcb200_sad_sample_sample", "xvcb200_sad_short_sample"):
        getattr(L, name).argtypes = blk
    for name in ("xvcb200_ssd_sample_sample", "xvcb200_ssd_short_sample", "xvcb200_ssd_short_short"):
        getattr(L, name).argtypes = blk
        getattr(L, name).restype = c_u64
    for name in ("xvcb200_compare_sample_sample", "xvcb200_compare_short_sample"):
        getattr(L, name).argtypes = [c_int, c_int] + blk
        getattr(L, name).restype = c_u64
    for name in ("h_sample_sample", "h_sample_short", "v_sample_sample", "v_sample_short", "v_short_sample", "v_short_short"):
        getattr(L, "xvcb200_filter_" + name).argtypes = [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_ssize, c_void_p, c_ssize]
    L.xvcb200_add_avg.argtypes = [c_int] * 5 + [c_void_p, c_ssize, c_void_p, c_ssize, c_void_p, c_ssize]
    L.xvcb200_filter_copy_bipred.argtypes = [c_int, c_int, ctypes.c_int16, c_int, c_void_p, c_ssize, c_void_p, c_ssize]
    for name in ("xvcb200_interp_block", "xvcb200_interp_block_bipred"):
        getattr(L, name).argtypes = [c_int] * 6 + [c_void_p, c_ssize, c_void_p, c_ssize]
    L.xvcb200_fwd_transform.argtypes = [c_int] * 6 + [c_void_p, c_ssize, c_void_p, c_ssize]
    L.xvcb200_inv_transform.argtypes = [c_int] * 7 + [c_void_p, c_ssize, c_void_p, c_ssize]
    for name in ("xvcb200_fwd_transform_skip", "xvcb200_inv_transform_skip"):
        getattr(L, name).argtypes = [c_int] * 3 + [c_void_p, c_ssize, c_void_p, c_ssize]
    L.xvcb200_qp_init.argtypes = [c_void_p, c_int, c_int, c_int, c_double, c_int, c_int, c_int]
    L.xvcb200_quant_fast.argtypes = [c_int] * 7 + [c_void_p, c_ssize, c_void_p, c_ssize]
    L.xvcb200_dequant.argtypes = [c_int] * 4 + [c_void_p, c_ssize, c_void_p, c_ssize]
    L.xvcb200_intra_ref_samples.argtypes = [c_int] * 8 + [c_void_p, c_ssize, c_void_p, c_void_p]
    L.xvcb200_intra_predict.argtypes = [c_int] * 5 + [c_void_p, c_void_p, c_void_p, c_ssize]
    L.xvcb200_intra_satd_scan.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]
    L.xvcb200_intra_lm_chroma.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int]
    L.xvcb200_ipc_export.argtypes = [c_void_p, c_void_p]
    L.xvcb200_ipc_open_peer.argtypes = [c_void_p, c_void_p, c_void_p]
    L.xvcb200_push_slot.argtypes = [c_void_p, c_int]
    L.xvcb200_wait_pushes.argtypes = [c_void_p, c_int]
    L.xvcb200_push_slot_tagged.argtypes = [c_void_p, c_int, ctypes.c_uint32]
    L.xvcb200_wait_slot_tag.argtypes = [c_void_p, c_int, ctypes.c_uint32]
    L.xvcb200_ctx_create.argtypes = [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int]
    L.xvcb200_ctx_destroy.argtypes = [c_void_p]
    L.xvcb200_ctx_set_stream.argtypes = [c_void_p, c_void_p]
    L.xvcb200_sync.argtypes = [c_void_p]
    L.xvcb200_get_geometry.argtypes = [c_void_p, c_void_p]
    L.xvcb200_slot_ptr.argtypes = [c_void_p, c_int, c_int, ctypes.POINTER(c_void_p)]
    L.xvcb200_slot_region.argtypes = [c_void_p, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_u64)]
    for name in ("xvcb200_upload_picture", "xvcb200_download_picture", "xvcb200_download_coeff", "xvcb200_upload_coeff",
                 "xvcb200_upload_picture_async", "xvcb200_download_picture_async", "xvcb200_download_coeff_async"):
        getattr(L, name).argtypes = [c_void_p, c_int, c_void_p, c_void_p]
    L.xvcb200_get_cus_async.argtypes = [c_void_p, c_void_p, c_int]
    L.xvcb200_sync_copies.argtypes = [c_void_p]
    L.xvcb200_wait_download.argtypes = [c_void_p, c_int]
    L.xvcb200_download_padded.argtypes = [c_void_p, c_int, c_int, c_void_p]
    L.xvcb200_pad_border.argtypes = [c_void_p, c_int]
    L.xvcb200_set_cus.argtypes = [c_void_p, c_void_p, c_int]
    L.xvcb200_get_cus.argtypes = [c_void_p, c_void_p, c_int]
    L.xvcb200_set_mv_predictors.argtypes = [c_void_p, c_void_p, c_int]
    L.xvcb200_set_tu_modes.argtypes = [c_void_p, c_void_p]
    L.xvcb200_me_search.argtypes = [c_void_p, c_int, c_void_p, c_int, c_double, c_void_p]
    L.xvcb200_full_search.argtypes = [c_void_p, c_int, c_void_p, c_int, c_double, c_void_p]
    L.xvcb200_motion_compensate.argtypes = [c_void_p, c_void_p, c_int]
    L.xvcb200_motion_compensate_affine.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int]
    L.xvcb200_motion_compensate_lic.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int]
    L.xvcb200_tq_reconstruct.argtypes = [c_void_p] + [c_int] * 9 + [c_void_p]
    L.xvcb200_dequant_reconstruct.argtypes = [c_void_p] + [c_int] * 6
    L.xvcb200_deblock_picture.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    L.xvcb200_deblock_picture_ex.argtypes = [c_void_p] + [c_int] * 7 + [c_void_p]
    L.xvcb200_deblock_picture_ext.argtypes = [c_void_p] + [c_int] * 7 + [c_void_p, c_void_p]
    L.xvcb200_deblock_band.argtypes = [c_void_p] + [c_int] * 7 + [c_void_p] + [c_int] * 3
    L.xvcb200_encode_picture.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    L.xvcb200_decide_partition.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]
    L.xvcb200_decide_partition_begin.argtypes = [c_void_p, c_void_p]
    L.xvcb200_decide_partition_end.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]
    L.xvcb200_set_profiling.argtypes = [c_void_p, c_int]
    L.xvcb200_get_stage_times.argtypes = [c_void_p, c_void_p]
    _lib = L
    return L


def _check_leaf(L):
    if L.xvcb200_last_error() != 0:
        msg = L.xvcb200_last_error_string().decode()
        L.xvcb200_clear_error()
        raise XvcB200Error(msg)


def _off(arr, row, col):
    return c_void_p(arr.ctypes.data + arr.dtype.itemsize * (row * arr.shape[1] + col))


# ---------------------------------------------------------------------- table-shaped calls
def sad(a, b, w, h):
    L = load()
    fn = L.xvcb200_sad_short_sample if a.dtype == np.int16 else L.xvcb200_sad_sample_sample
    r = fn(w, h, abi.ptr(a), a.shape[1], abi.ptr(b), b.shape[1])
    _check_leaf(L)
    return r


def ssd(a, b, w, h):
    L = load()
    if a.dtype == np.int16 and b.dtype == np.int16:
        fn = L.xvcb200_ssd_short_short
    elif a.dtype == np.int16:
        fn = L.xvcb200_ssd_short_sample
    else:
        fn = L.xvcb200_ssd_sample_sample
    r = fn(w, h, abi.ptr(a), a.shape[1], abi.ptr(b), b.shape[1])
    _check_leaf(L)
    return r


def compare(metric, bitdepth, a, b, w, h):
    L = load()
    fn = L.xvcb200_compare_short_sample if a.dtype == np.int16 else L.xvcb200_compare_sample_sample
    r = fn(metric, bitdepth, w, h, abi.ptr(a), a.shape[1], abi.ptr(b), b.shape[1])
    _check_leaf(L)
    return r


_FILTER_NAMES = ["h_sample_sample", "h_sample_short", "v_sample_sample", "v_sample_short", "v_short_sample", "v_short_short"]


def filter_block(kind, chroma, w, h, bitdepth, taps, src, src_off, dst):
    L = load()
    fn = getattr(L, "xvcb200_filter_" + _FILTER_NAMES[kind])
    fn(chroma, w, h, bitdepth, abi.ptr(np.ascontiguousarray(taps, dtype=np.int16)), _off(src, *src_off), src.shape[1],
       abi.ptr(dst), dst.shape[1])
    _check_leaf(L)


def interp_block(chroma, bipred, w, h, bitdepth, fx, fy, ref, ref_off, pred):
    L = load()
    fn = L.xvcb200_interp_block_bipred if bipred else L.xvcb200_interp_block
    fn(chroma, w, h, bitdepth, fx, fy, _off(ref, *ref_off), ref.shape[1], abi.ptr(pred), pred.shape[1])
    _check_leaf(L)


def add_avg(w, h, offset, shift, bitdepth, a, b, dst):
    L = load()
    L.xvcb200_add_avg(w, h, offset, shift, bitdepth, abi.ptr(a), a.shape[1], abi.ptr(b), b.shape[1], abi.ptr(dst), dst.shape[1])
    _check_leaf(L)


def filter_copy_bipred(w, h, offset, shift, ref, pred):
    L = load()
    L.xvcb200_filter_copy_bipred(w, h, offset, shift, abi.ptr(ref), ref.shape[1], abi.ptr(pred), pred.shape[1])
    _check_leaf(L)


def fwd_transform(w, h, bitdepth, tx_hor, tx_ver, dst4x4, resi):
    L = load()
    out = np.zeros((h, w), dtype=np.int16)
    L.xvcb200_fwd_transform(w, h, bitdepth, tx_hor, tx_ver, dst4x4, abi.ptr(resi), resi.shape[1], abi.ptr(out), w)
    _check_leaf(L)
    return out


def inv_transform(w, h, bitdepth, tx_hor, tx_ver, dst4x4, dc_only, coeff):
    L = load()
    out = np.zeros((h, w), dtype=np.int16)
    L.xvcb200_inv_transform(w, h, bitdepth, tx_hor, tx_ver, dst4x4, dc_only, abi.ptr(coeff), coeff.shape[1], abi.ptr(out), w)
    _check_leaf(L)
    return out


def transform_skip(forward, w, h, bitdepth, inp):
    L = load()
    out = np.zeros((h, w), dtype=np.int16)
    fn = L.xvcb200_fwd_transform_skip if forward else L.xvcb200_inv_transform_skip
    fn(w, h, bitdepth, abi.ptr(inp), inp.shape[1], abi.ptr(out), w)
    _check_leaf(L)
    return out


def qp_init(qp, bitdepth, lam=1.0, table=1, off_u=0, off_v=0, chroma_format=1):
    L = load()
    q = np.zeros(1, dtype=abi.qp_dtype)
    L.xvcb200_qp_init(abi.ptr(q), qp, chroma_format, bitdepth, lam, table, off_u, off_v)
    return q[0]


def quant_fast(w, h, bitdepth, qp_bd, intra_pic, sign_hiding, scan, coeff):
    L = load()
    out = np.zeros((h, w), dtype=np.int16)
    nz = L.xvcb200_quant_fast(w, h, bitdepth, qp_bd, intra_pic, sign_hiding, scan, abi.ptr(coeff), coeff.shape[1], abi.ptr(out), w)
    _check_leaf(L)
    return out, nz


def dequant(w, h, bitdepth, qp_bd, lev):
    L = load()
    out = np.zeros((h, w), dtype=np.int16)
    L.xvcb200_dequant(w, h, bitdepth, qp_bd, abi.ptr(lev), lev.shape[1], abi.ptr(out), w)
    _check_leaf(L)
    return out


def intra_ref_samples(w, h, bitdepth, nb, plane, x, y, want_filtered=True):
    """IntraPrediction::FillReferenceState for the block at (x, y) of a host plane; nb = (has_above_left,
    has_above, above_right, has_left, below_left).  Returns (ref_samples, ref_filtered) of 2 x 129 samples."""
    L = load()
    plane = np.ascontiguousarray(plane, dtype=np.uint16)
    ref = np.zeros(2 * abi.INTRA_REF_STRIDE, dtype=np.uint16)
    filt = np.zeros(2 * abi.INTRA_REF_STRIDE, dtype=np.uint16) if want_filtered else None
    L.xvcb200_intra_ref_samples(w, h, bitdepth, int(nb[0]), int(nb[1]), int(nb[2]), int(nb[3]), int(nb[4]),
                                _off(plane, y, x), plane.strides[0] // 2, abi.ptr(ref), abi.ptr(filt))
    _check_leaf(L)
    return ref, filt


def intra_predict(mode, w, h, bitdepth, is_luma, ref, filt):
    L = load()
    pred = np.zeros((h, w), dtype=np.uint16)
    L.xvcb200_intra_predict(mode, w, h, bitdepth, int(is_luma), abi.ptr(np.ascontiguousarray(ref, dtype=np.uint16)),
                            abi.ptr(None if filt is None else np.ascontiguousarray(filt, dtype=np.uint16)), abi.ptr(pred), w)
    _check_leaf(L)
    return pred


def launch_count():
    return int(load().xvcb200_launch_count())


# ---------------------------------------------------------------------- batched context
class Context:
    """xvcb200_ctx: device-resident picture slots + the per-picture hot path."""

    def __init__(self, width, height, bitdepth=10, num_slots=8, device=0):
        self.L = load()
        self.h = c_void_p()
        st = self.L.xvcb200_ctx_create(ctypes.byref(self.h), device, width, height, bitdepth, 1, num_slots)
        if st != 0:
            msg = self.L.xvcb200_last_error_string().decode()
            self.L.xvcb200_clear_error()
            self.h = None
            raise XvcB200Error("xvcb200_ctx_create failed (%d): %s" % (st, msg))
        self.width, self.height, self.bitdepth, self.num_slots = width, height, bitdepth, num_slots
        self.shapes = [(height, width), (height // 2, width // 2), (height // 2, width // 2)]
        g = np.zeros(1, dtype=abi.plane_geom_dtype)
        self.L.xvcb200_get_geometry(self.h, abi.ptr(g))
        self.geom = g[0]
        self.n_cus = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.xvcb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ok(self, st):
        if st != 0:
            msg = self.L.xvcb200_ctx_error_string(self.h).decode() or self.L.xvcb200_last_error_string().decode()
            self.L.xvcb200_clear_error()      # reported here: the thread's last-error slot is for the table-shaped entries
            raise XvcB200Error("xvc_b200 call failed (%d): %s" % (st, msg))

    def sync(self):
        self._ok(self.L.xvcb200_sync(self.h))

    def stream(self):
        return self.L.xvcb200_stream(self.h)

    def set_stream(self, cuda_stream):
        self._ok(self.L.xvcb200_ctx_set_stream(self.h, c_void_p(cuda_stream)))

    def slot_ptr(self, slot, comp):
        p = c_void_p()
        self._ok(self.L.xvcb200_slot_ptr(self.h, slot, comp, ctypes.byref(p)))
        return p.value

    def slot_region(self, slot):
        """(device address, bytes) of the whole allocation of a slot; slots are contiguous."""
        p, n = c_void_p(), c_u64()
        self._ok(self.L.xvcb200_slot_region(self.h, slot, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def ipc_export(self):
        """Handle of the slot arena (CUDA IPC handle + arena layout, XVCB200_IPC_HANDLE_BYTES) for the other
        processes of the node."""
        h = np.zeros(IPC_HANDLE_BYTES, dtype=np.uint8)
        self._ok(self.L.xvcb200_ipc_export(self.h, abi.ptr(h)))
        return h.tobytes()

    def ipc_open_peer(self, handle):
        if len(handle) != IPC_HANDLE_BYTES:
            raise XvcB200Error("ipc_open_peer: handle must be %d bytes" % IPC_HANDLE_BYTES)
        h = np.frombuffer(handle, dtype=np.uint8).copy()
        self._ok(self.L.xvcb200_ipc_open_peer(self.h, abi.ptr(h), None))

    def push_slot(self, slot):
        self._ok(self.L.xvcb200_push_slot(self.h, slot))

    def wait_pushes(self, slot=-1):
        self._ok(self.L.xvcb200_wait_pushes(self.h, slot))

    def push_slot_tagged(self, slot, tag):
        """push_slot + the arrival tag written behind it into every peer's arena (device-side rendezvous)."""
        self._ok(self.L.xvcb200_push_slot_tagged(self.h, slot, tag))

    def wait_slot_tag(self, slot, tag):
        """The context stream waits until a peer's tagged push of `slot` has arrived (no host wait)."""
        self._ok(self.L.xvcb200_wait_slot_tag(self.h, slot, tag))

    def slots_tensor(self, first, count=1):
        """torch uint8 CUDA tensor aliasing slots [first, first+count) (no copy)."""
        import torch
        base, nbytes = self.slot_region(first)

        class _Mem:
            __cuda_array_interface__ = {"shape": (count * nbytes,), "typestr": "|u1", "data": (base, False), "version": 3}
        return torch.as_tensor(_Mem(), device="cuda")

    def plane_tensor(self, slot, comp):
        """torch int16/uint16-as-int16 view [rows, pitch] of one plane incl. margins (no copy)."""
        import torch
        g = self.geom
        rows = int(g["height"][comp] + 2 * g["margin_y"][comp])
        pitch = int(g["pitch"][comp])
        p00 = self.slot_ptr(slot, comp)
        base = p00 - 2 * (int(g["margin_y"][comp]) * pitch + int(g["margin_x"][comp]))

        class _Mem:
            __cuda_array_interface__ = {"shape": (rows, pitch), "typestr": "<i2", "data": (base, False), "version": 3}
        return torch.as_tensor(_Mem(), device="cuda")

    @staticmethod
    def _strides(planes):
        return (c_ssize * 3)(*[p.strides[0] // p.itemsize for p in planes])

    def upload(self, slot, planes):
        planes = [p if p.flags["C_CONTIGUOUS"] or p.strides[1] == p.itemsize else np.ascontiguousarray(p) for p in planes]
        self._keep = planes
        self._ok(self.L.xvcb200_upload_picture(self.h, slot, abi.plane_ptr_array(planes), self._strides(planes)))

    def upload_coeff(self, slot, planes):
        planes = [np.ascontiguousarray(p, dtype=np.int16) for p in planes]
        self._keep = planes
        self._ok(self.L.xvcb200_upload_coeff(self.h, slot, abi.plane_ptr_array(planes), self._strides(planes)))

    def download(self, slot, out=None):
        planes = out if out is not None else [np.zeros(s, dtype=np.uint16) for s in self.shapes]
        self._ok(self.L.xvcb200_download_picture(self.h, slot, abi.plane_ptr_array(planes), self._strides(planes)))
        return planes

    def download_coeff(self, slot):
        planes = [np.zeros(s, dtype=np.int16) for s in self.shapes]
        self._ok(self.L.xvcb200_download_coeff(self.h, slot, abi.plane_ptr_array(planes), self._strides(planes)))
        return planes

    # transfers on the context's copy stream (see include/xvc_b200.h); the caller keeps the
    # (page-locked) host arrays alive and unchanged until sync_copies()
    def upload_async(self, slot, planes):
        self._ok(self.L.xvcb200_upload_picture_async(self.h, slot, abi.plane_ptr_array(planes), self._strides(planes)))

    def download_async(self, slot, planes):
        self._ok(self.L.xvcb200_download_picture_async(self.h, slot, abi.plane_ptr_array(planes), self._strides(planes)))

    def download_coeff_async(self, slot, planes):
        self._ok(self.L.xvcb200_download_coeff_async(self.h, slot, abi.plane_ptr_array(planes), self._strides(planes)))

    def get_cus_async(self, out):
        self._ok(self.L.xvcb200_get_cus_async(self.h, abi.ptr(out), self.n_cus))

    def wait_download(self, slot):
        self._ok(self.L.xvcb200_wait_download(self.h, slot))

    def sync_copies(self):
        self._ok(self.L.xvcb200_sync_copies(self.h))

    def download_padded(self, slot, comp):
        pad = 80 if comp == 0 else 40
        h, w = self.shapes[comp]
        out = np.zeros((h + 2 * pad, w + 2 * pad), dtype=np.uint16)
        self._ok(self.L.xvcb200_download_padded(self.h, slot, comp, abi.ptr(out)))
        return out

    def pad_border(self, slot):
        self._ok(self.L.xvcb200_pad_border(self.h, slot))

    def set_cus(self, cus):
        cus = np.ascontiguousarray(cus, dtype=abi.cu_dtype)
        self._ok(self.L.xvcb200_set_cus(self.h, abi.ptr(cus), len(cus)))
        self.n_cus = len(cus)

    def set_tu_modes(self, modes):
        """modes: abi.tu_mode_dtype [n_cus] or None to drop them; after set_cus."""
        if modes is not None:
            modes = np.ascontiguousarray(modes, dtype=abi.tu_mode_dtype)
            assert len(modes) == self.n_cus
        self._ok(self.L.xvcb200_set_tu_modes(self.h, abi.ptr(modes)))

    def set_mv_predictors(self, mvp):
        """mvp: int32 [n_cus][columns][2] (1/16 pel) or None to drop them; after set_cus."""
        if mvp is None:
            self._ok(self.L.xvcb200_set_mv_predictors(self.h, None, 0))
            return
        mvp = np.ascontiguousarray(mvp, dtype=np.int32)
        assert mvp.ndim == 3 and mvp.shape[0] == self.n_cus and mvp.shape[2] == 2
        self._ok(self.L.xvcb200_set_mv_predictors(self.h, abi.ptr(mvp), mvp.shape[1]))

    def get_cus(self):
        out = np.zeros(self.n_cus, dtype=abi.cu_dtype)
        self._ok(self.L.xvcb200_get_cus(self.h, abi.ptr(out), self.n_cus))
        return out

    def me_search(self, orig_slot, jobs, lambda_sqrt):
        jobs = np.ascontiguousarray(jobs, dtype=abi.me_job_dtype)
        res = np.zeros(len(jobs), dtype=abi.me_result_dtype)
        self._ok(self.L.xvcb200_me_search(self.h, orig_slot, abi.ptr(jobs), len(jobs), lambda_sqrt, abi.ptr(res)))
        return res

    def full_search(self, orig_slot, jobs, lambda_sqrt):
        jobs = np.ascontiguousarray(jobs, dtype=abi.fullsearch_job_dtype)
        res = np.zeros(len(jobs), dtype=abi.me_result_dtype)
        self._ok(self.L.xvcb200_full_search(self.h, orig_slot, abi.ptr(jobs), len(jobs), lambda_sqrt, abi.ptr(res)))
        return res

    @staticmethod
    def _ref_slots(ref_slots):
        arr = np.full((2, 5), -1, dtype=np.int32)
        for (l, i), s in ref_slots.items():
            arr[l, i] = s
        return arr

    def motion_compensate(self, ref_slots, pred_slot):
        arr = self._ref_slots(ref_slots)
        self._ok(self.L.xvcb200_motion_compensate(self.h, abi.ptr(arr), pred_slot))

    def motion_compensate_affine(self, aff, ref_slots, pred_slot):
        """aff: abi.affine_cu_dtype array (CU index + control-point MVs per list)."""
        aff = np.ascontiguousarray(aff, dtype=abi.affine_cu_dtype)
        arr = self._ref_slots(ref_slots)
        self._ok(self.L.xvcb200_motion_compensate_affine(self.h, abi.ptr(aff), len(aff), abi.ptr(arr), pred_slot))

    def intra_lm_chroma(self, rec_slot, jobs, pred_slot):
        """LM chroma prediction of both chroma blocks of every job (abi.intra_job_dtype, luma x/y/w/h) into pred_slot."""
        jobs = np.ascontiguousarray(jobs, dtype=abi.intra_job_dtype)
        self._ok(self.L.xvcb200_intra_lm_chroma(self.h, rec_slot, abi.ptr(jobs), len(jobs), pred_slot))

    def motion_compensate_lic(self, lic, ref_slots, rec_slot, pred_slot):
        """lic: abi.lic_cu_dtype array (CU index + position of the CU above / left, -1 = none)."""
        lic = np.ascontiguousarray(lic, dtype=abi.lic_cu_dtype)
        arr = self._ref_slots(ref_slots)
        self._ok(self.L.xvcb200_motion_compensate_lic(self.h, abi.ptr(lic), len(lic), abi.ptr(arr), rec_slot, pred_slot))

    def tq_reconstruct(self, orig_slot, pred_slot, rec_slot, coeff_slot, intra_picture=0, table=1, off_u=0, off_v=0):
        res = np.zeros(3 * self.n_cus, dtype=abi.tu_result_dtype)
        self._ok(self.L.xvcb200_tq_reconstruct(self.h, orig_slot, pred_slot, rec_slot, coeff_slot, 0, intra_picture, table,
                                               off_u, off_v, abi.ptr(res)))
        return res

    def dequant_reconstruct(self, pred_slot, rec_slot, coeff_slot, table=1, off_u=0, off_v=0):
        self._ok(self.L.xvcb200_dequant_reconstruct(self.h, pred_slot, rec_slot, coeff_slot, table, off_u, off_v))

    def deblock_picture(self, rec_slot, pic_type, ref_poc, beta_offset=0, tc_offset=0, table=1, off_u=0, off_v=0):
        poc = np.zeros((2, 5), dtype=np.int64)
        for (l, i), p in ref_poc.items():
            poc[l, i] = p
        self._ok(self.L.xvcb200_deblock_picture_ex(self.h, rec_slot, pic_type, beta_offset, tc_offset, table, off_u, off_v,
                                                   abi.ptr(poc)))

    def deblock_picture_ext(self, rec_slot, pic_type, ref_poc, affine=None, chroma_cus=None, beta_offset=0, tc_offset=0, table=1,
                            off_u=0, off_v=0):
        """Deblocking with affine CUs (abi.affine_cu_dtype) and / or the secondary CU tree of an intra picture (abi.cu_dtype)."""
        poc = np.zeros((2, 5), dtype=np.int64)
        for (l, i), p in ref_poc.items():
            poc[l, i] = p
        ext = np.zeros(4, dtype=np.uint64)        # xvcb200_deblock_ext: pointer, int32 (+ padding), pointer, int32 (+ padding)
        if affine is not None:
            affine = np.ascontiguousarray(affine, dtype=abi.affine_cu_dtype)
            ext[0], ext[1] = affine.ctypes.data, len(affine)
        if chroma_cus is not None:
            chroma_cus = np.ascontiguousarray(chroma_cus, dtype=abi.cu_dtype)
            ext[2], ext[3] = chroma_cus.ctypes.data, len(chroma_cus)
        self._ok(self.L.xvcb200_deblock_picture_ext(self.h, rec_slot, pic_type, beta_offset, tc_offset, table, off_u, off_v,
                                                    abi.ptr(poc), abi.ptr(ext)))

    def deblock_band(self, rec_slot, pic_type, ref_poc, pass_mask, y_begin, y_end, beta_offset=0, tc_offset=0, table=1,
                     off_u=0, off_v=0):
        poc = np.zeros((2, 5), dtype=np.int64)
        for (l, i), p in ref_poc.items():
            poc[l, i] = p
        self._ok(self.L.xvcb200_deblock_band(self.h, rec_slot, pic_type, beta_offset, tc_offset, table, off_u, off_v,
                                             abi.ptr(poc), pass_mask, y_begin, y_end))

    def intra_satd_scan(self, orig_slot, src_slot, jobs):
        jobs = np.ascontiguousarray(jobs, dtype=abi.intra_job_dtype)
        out = np.zeros((len(jobs), abi.INTRA_NUM_MODES), dtype=np.uint32)
        self._ok(self.L.xvcb200_intra_satd_scan(self.h, orig_slot, src_slot, abi.ptr(jobs), len(jobs), abi.ptr(out)))
        return out

    STAGES = ("me_jobs", "tz_search", "subpel_search", "motion_compensate", "tq_reconstruct", "deblock", "pad_border")

    def set_profiling(self, on=True):
        self._ok(self.L.xvcb200_set_profiling(self.h, int(on)))

    def stage_times_ms(self):
        ms = np.zeros(7, dtype=np.float32)
        self._ok(self.L.xvcb200_get_stage_times(self.h, abi.ptr(ms)))
        return dict(zip(self.STAGES, [float(v) for v in ms]))

    def decide_partition_begin(self, orig_slot, ref_slot, lambda_sqrt, qp, center=(0, 0), header_bits_cu=0, header_bits_split=0):
        """Enqueues the GPU pre-analysis (kernel + copy of its result); decide_partition_end() collects it."""
        prm = np.zeros(1, dtype=abi.partition_params_dtype)
        prm["orig_slot"], prm["ref_slot"], prm["center"], prm["lambda_sqrt"], prm["qp"] = orig_slot, ref_slot, center, lambda_sqrt, qp
        prm["header_bits_cu"], prm["header_bits_split"] = header_bits_cu, header_bits_split
        self._ok(self.L.xvcb200_decide_partition_begin(self.h, abi.ptr(prm)))

    def decide_partition_end(self):
        """-> (CU array in coding order with mv = the winning vectors, split flags)."""
        n_ctus = ((self.width + 63) // 64) * ((self.height + 63) // 64)
        if getattr(self, "_part_buf", None) is None:
            self._part_buf = (np.zeros(64 * n_ctus, dtype=abi.cu_dtype), np.zeros(128 * n_ctus, dtype=np.uint8), np.zeros(2, dtype=np.int32))
        cus, splits, n = self._part_buf
        self._ok(self.L.xvcb200_decide_partition_end(self.h, abi.ptr(cus), len(cus), ctypes.c_void_p(n.ctypes.data),
                                                     abi.ptr(splits), len(splits), ctypes.c_void_p(n.ctypes.data + 4)))
        return cus[:n[0]].copy(), splits[:n[1]].copy()

    def decide_partition(self, orig_slot, ref_slot, lambda_sqrt, qp, center=(0, 0), header_bits_cu=0, header_bits_split=0):
        """GPU pre-analysis -> (CU array in coding order with mv = the winning vectors, split flags)."""
        self.decide_partition_begin(orig_slot, ref_slot, lambda_sqrt, qp, center, header_bits_cu, header_bits_split)
        return self.decide_partition_end()

    def encode_picture(self, params, want_results=True):
        prm = params if isinstance(params, np.ndarray) else np.array([params], dtype=abi.picture_params_dtype)
        me = np.zeros(abi.num_me_columns(prm) * self.n_cus, dtype=abi.me_result_dtype) if want_results else None
        tu = np.zeros(3 * self.n_cus, dtype=abi.tu_result_dtype) if want_results else None
        self._ok(self.L.xvcb200_encode_picture(self.h, abi.ptr(prm), abi.ptr(me), abi.ptr(tu)))
        return me, tu
