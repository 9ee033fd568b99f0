"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the C oracle (libxvc_oracle.so) and, when it
has been built, for the compiled reference (oracle/_ref/libxvcref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  Nothing under xvc_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

from xvc_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libxvc_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libxvcref.so")
REF_XVCENC = os.path.join(_HERE, "_ref", "xvcenc")
REF_XVCDEC = os.path.join(_HERE, "_ref", "xvcdec")

c_int, c_void_p, c_double, c_u64, c_u32, c_i64 = (ctypes.c_int, ctypes.c_void_p, ctypes.c_double,
                                                  ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64)
c_ssize = ctypes.c_ssize_t


def build(ref=True):
    """make the oracle (and the reference when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-j8", "-C", _HERE, "ref"])


class XoPicture(ctypes.Structure):
    _fields_ = [("base", c_void_p * 3), ("stride", ctypes.c_int32 * 3), ("width", ctypes.c_int32 * 3),
                ("height", ctypes.c_int32 * 3), ("pad", ctypes.c_int32 * 3)]


class Picture:
    """Three padded uint16 planes; plane(c) is the visible area, full[c] the allocation."""

    def __init__(self, width, height, pad=80, planes=None):
        self.width = [width, width // 2, width // 2]
        self.height = [height, height // 2, height // 2]
        self.pad = [pad, pad // 2, pad // 2]
        self.full = [np.zeros((self.height[c] + 2 * self.pad[c], self.width[c] + 2 * self.pad[c]), dtype=np.uint16)
                     for c in range(3)]
        if planes is not None:
            for c in range(3):
                self.plane(c)[...] = planes[c]

    def plane(self, c):
        p = self.pad[c]
        return self.full[c][p:p + self.height[c], p:p + self.width[c]]

    def planes(self):
        return [np.ascontiguousarray(self.plane(c)) for c in range(3)]

    def c_struct(self):
        s = XoPicture()
        for c in range(3):
            stride = self.full[c].shape[1]
            s.base[c] = self.full[c].ctypes.data + 2 * (self.pad[c] * stride + self.pad[c])
            s.stride[c] = stride
            s.width[c], s.height[c], s.pad[c] = self.width[c], self.height[c], self.pad[c]
        return s

    def copy(self):
        q = Picture(self.width[0], self.height[0], self.pad[0])
        for c in range(3):
            q.full[c][...] = self.full[c]
        return q


def _refs_array(refs):
    """const xo_picture* refs[2][5] from {(list, idx): Picture}; returns (array, keepalive)."""
    arr = ((ctypes.POINTER(XoPicture) * 5) * 2)()
    keep = []
    for (l, i), pic in refs.items():
        s = pic.c_struct()
        keep.append(s)
        arr[l][i] = ctypes.pointer(s)
    return arr, keep


def _poc_array(ref_poc):
    arr = ((c_i64 * 5) * 2)()
    for (l, i), poc in ref_poc.items():
        arr[l][i] = poc
    return arr


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = self.L = ctypes.CDLL(ORACLE_SO)
        L.xo_ssd.restype = c_u64
        L.xo_satd.restype = c_u64
        L.xo_compare.restype = c_u64
        L.xo_apply_weight.restype = c_u64
        L.xo_apply_weight.argtypes = [c_u64, c_double]
        L.xo_transform_matrix.restype = ctypes.POINTER(ctypes.c_int16)
        L.xo_luma_taps.restype = ctypes.POINTER(ctypes.c_int16)
        L.xo_chroma_taps.restype = ctypes.POINTER(ctypes.c_int16)
        L.xo_exp_golomb_bits.restype = c_u32
        L.xo_qp_init.argtypes = [c_void_p, c_int, c_int, c_int, c_double, c_int, c_int, c_int]
        L.xo_me_search.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_double, c_void_p]
        L.xo_sad.argtypes = [c_int, c_int, c_int, c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_ssd.argtypes = [c_int, c_int, c_int, c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_compare.argtypes = [c_int, c_int, c_int, c_int, c_int, c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_filter.argtypes = [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_interp.argtypes = [c_int] * 7 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_add_avg.argtypes = [c_int] * 5 + [c_void_p, c_ssize, c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_filter_copy_bipred.argtypes = [c_int] * 4 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_fwd_transform.argtypes = [c_int] * 6 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_inv_transform.argtypes = [c_int] * 7 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_transform_skip.argtypes = [c_int] * 4 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_quant_fast.argtypes = [c_int] * 7 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_dequant.argtypes = [c_int] * 4 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_pad_border.argtypes = [c_void_p]
        L.xo_full_search.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_u32, c_void_p, c_void_p]
        L.xo_motion_compensate.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p]
        L.xo_intra_lm_chroma.argtypes = [c_int] * 5 + [c_void_p, c_ssize, c_void_p, c_ssize, c_void_p, c_ssize]
        L.xo_motion_compensate_lic.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]
        L.xo_motion_compensate_affine.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]
        L.xo_tq_reconstruct.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
        L.xo_dequant_reconstruct.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int]
        L.xo_deblock_picture.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
        L.xo_deblock_band.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int]
        L.xo_encode_picture.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
        L.xo_intra_ref_samples.argtypes = [c_int] * 8 + [c_void_p, c_ssize, c_void_p]
        L.xo_intra_filter_ref.argtypes = [c_int, c_int, c_void_p, c_void_p]
        L.xo_intra_predict.argtypes = [c_int] * 5 + [c_void_p, c_void_p, c_void_p, c_ssize]
        L.xo_intra_satd_scan.argtypes = [c_int, c_int, c_int, c_void_p, c_ssize, c_void_p, c_void_p, c_void_p]

    # ---- leaf
    def sad(self, kind, a, b, w, h, sa=None, sb=None):
        return self.L.xo_sad(kind, w, h, abi.ptr(a), sa or a.shape[1], abi.ptr(b), sb or b.shape[1])

    def ssd(self, kind, a, b, w, h):
        return self.L.xo_ssd(kind, w, h, abi.ptr(a), a.shape[1], abi.ptr(b), b.shape[1])

    def compare(self, metric, bitdepth, a, b, w, h):
        return self.L.xo_compare(metric, bitdepth, w, h, int(a.dtype == np.int16), abi.ptr(a), a.shape[1],
                                 abi.ptr(b), b.shape[1])

    def taps(self, chroma, frac):
        p = self.L.xo_chroma_taps(frac) if chroma else self.L.xo_luma_taps(frac)
        return np.ctypeslib.as_array(p, shape=(4 if chroma else 8,)).copy()

    def filter(self, kind, chroma, w, h, bitdepth, taps, src, src_off, dst):
        """src_off = (row, col) of the centre sample of output (0,0) inside `src`."""
        esz = src.dtype.itemsize
        sp = src.ctypes.data + esz * (src_off[0] * src.shape[1] + src_off[1])
        self.L.xo_filter(kind, chroma, w, h, bitdepth, abi.ptr(taps), c_void_p(sp), src.shape[1], abi.ptr(dst), dst.shape[1])

    def interp(self, chroma, bipred, w, h, bitdepth, fx, fy, ref, ref_off, pred):
        sp = ref.ctypes.data + 2 * (ref_off[0] * ref.shape[1] + ref_off[1])
        self.L.xo_interp(chroma, bipred, w, h, bitdepth, fx, fy, c_void_p(sp), ref.shape[1], abi.ptr(pred), pred.shape[1])

    def add_avg(self, w, h, offset, shift, bitdepth, a, b, dst):
        self.L.xo_add_avg(w, h, offset, shift, bitdepth, abi.ptr(a), a.shape[1], abi.ptr(b), b.shape[1], abi.ptr(dst), dst.shape[1])

    def filter_copy_bipred(self, w, h, offset, shift, ref, pred):
        self.L.xo_filter_copy_bipred(w, h, offset, shift, abi.ptr(ref), ref.shape[1], abi.ptr(pred), pred.shape[1])

    def matrix(self, tx_type, n):
        p = self.L.xo_transform_matrix(tx_type, n)
        return None if not p else np.ctypeslib.as_array(p, shape=(n * n,)).reshape(n, n).copy()

    def fwd_transform(self, w, h, bitdepth, tx_hor, tx_ver, dst4x4, resi):
        out = np.zeros((h, w), dtype=np.int16)
        self.L.xo_fwd_transform(w, h, bitdepth, tx_hor, tx_ver, dst4x4, abi.ptr(resi), resi.shape[1], abi.ptr(out), w)
        return out

    def inv_transform(self, w, h, bitdepth, tx_hor, tx_ver, dst4x4, dc_only, coeff):
        out = np.zeros((h, w), dtype=np.int16)
        self.L.xo_inv_transform(w, h, bitdepth, tx_hor, tx_ver, dst4x4, dc_only, abi.ptr(coeff), coeff.shape[1], abi.ptr(out), w)
        return out

    def transform_skip(self, forward, w, h, bitdepth, inp):
        out = np.zeros((h, w), dtype=np.int16)
        self.L.xo_transform_skip(forward, w, h, bitdepth, abi.ptr(inp), inp.shape[1], abi.ptr(out), w)
        return out

    def qp(self, qp, bitdepth, lam=1.0, table=1, off_u=0, off_v=0):
        q = np.zeros(1, dtype=abi.qp_dtype)
        self.L.xo_qp_init(abi.ptr(q), qp, 1, bitdepth, lam, table, off_u, off_v)
        return q[0]

    def quant_fast(self, w, h, bitdepth, qp_bd, intra_pic, sign_hiding, scan, coeff):
        out = np.zeros((h, w), dtype=np.int16)
        nz = self.L.xo_quant_fast(w, h, bitdepth, qp_bd, intra_pic, sign_hiding, scan, abi.ptr(coeff), coeff.shape[1], abi.ptr(out), w)
        return out, nz

    def dequant(self, w, h, bitdepth, qp_bd, lev):
        out = np.zeros((h, w), dtype=np.int16)
        self.L.xo_dequant(w, h, bitdepth, qp_bd, abi.ptr(lev), lev.shape[1], abi.ptr(out), w)
        return out

    # ---- picture level
    def pad_border(self, pic):
        s = pic.c_struct()
        self.L.xo_pad_border(ctypes.byref(s))

    # ---- intra prediction (xvc_oracle_intra.c)
    def intra_ref_samples(self, w, h, bitdepth, nb, plane, x, y):
        plane = np.ascontiguousarray(plane, dtype=np.uint16)
        ref = np.zeros(2 * abi.INTRA_REF_STRIDE, dtype=np.uint16)
        filt = np.zeros(2 * abi.INTRA_REF_STRIDE, dtype=np.uint16)
        blk = ctypes.c_void_p(plane.ctypes.data + (y * plane.shape[1] + x) * 2)
        self.L.xo_intra_ref_samples(w, h, bitdepth, int(nb[0]), int(nb[1]), int(nb[2]), int(nb[3]), int(nb[4]), blk,
                                    ctypes.c_ssize_t(plane.shape[1]), abi.ptr(ref))
        self.L.xo_intra_filter_ref(w, h, abi.ptr(ref), abi.ptr(filt))
        return ref, filt

    def intra_predict(self, mode, w, h, bitdepth, is_luma, ref, filt):
        pred = np.zeros((h, w), dtype=np.uint16)
        self.L.xo_intra_predict(mode, w, h, bitdepth, int(is_luma), abi.ptr(np.ascontiguousarray(ref, dtype=np.uint16)),
                                abi.ptr(None if filt is None else np.ascontiguousarray(filt, dtype=np.uint16)), abi.ptr(pred),
                                ctypes.c_ssize_t(w))
        return pred

    def intra_satd_scan(self, w, h, bitdepth, orig_plane, x, y, ref, filt):
        orig_plane = np.ascontiguousarray(orig_plane, dtype=np.uint16)
        out = np.zeros(abi.INTRA_NUM_MODES, dtype=np.uint32)
        blk = ctypes.c_void_p(orig_plane.ctypes.data + (y * orig_plane.shape[1] + x) * 2)
        self.L.xo_intra_satd_scan(w, h, bitdepth, blk, ctypes.c_ssize_t(orig_plane.shape[1]), abi.ptr(ref), abi.ptr(filt), abi.ptr(out))
        return out

    def me_search(self, orig, refs, bitdepth, cus, jobs, lambda_sqrt):
        res = np.zeros(len(jobs), dtype=abi.me_result_dtype)
        o = orig.c_struct()
        arr, keep = _refs_array(refs)
        self.L.xo_me_search(ctypes.addressof(o), ctypes.addressof(arr), bitdepth, abi.ptr(cus), abi.ptr(jobs), len(jobs),
                            lambda_sqrt, abi.ptr(res))
        return res

    def full_search(self, orig, other_pred, ref, bitdepth, cu, job, lambda_me):
        mv = np.zeros(2, dtype=np.int32)
        cost = c_u32(0)
        o, p, r = orig.c_struct(), other_pred.c_struct(), ref.c_struct()
        self.L.xo_full_search(ctypes.byref(o), ctypes.byref(p), ctypes.byref(r), bitdepth, abi.ptr(cu), abi.ptr(job),
                              c_u32(lambda_me), abi.ptr(mv), ctypes.byref(cost))
        return mv, cost.value

    def motion_compensate(self, refs, bitdepth, cus, pred):
        arr, keep = _refs_array(refs)
        p = pred.c_struct()
        self.L.xo_motion_compensate(ctypes.addressof(arr), bitdepth, abi.ptr(cus), len(cus), ctypes.byref(p))

    def motion_compensate_affine(self, refs, bitdepth, cus, aff, pred):
        arr, keep = _refs_array(refs)
        p = pred.c_struct()
        aff = np.ascontiguousarray(aff, dtype=abi.affine_cu_dtype)
        self.L.xo_motion_compensate_affine(ctypes.addressof(arr), bitdepth, abi.ptr(cus), abi.ptr(aff), len(aff), ctypes.byref(p))

    def motion_compensate_lic(self, refs, rec, bitdepth, cus, lic, pred):
        arr, keep = _refs_array(refs)
        p, r = pred.c_struct(), rec.c_struct()
        lic = np.ascontiguousarray(lic, dtype=abi.lic_cu_dtype)
        self.L.xo_motion_compensate_lic(ctypes.addressof(arr), ctypes.byref(r), bitdepth, abi.ptr(cus), abi.ptr(lic), len(lic), ctypes.byref(p))

    def intra_lm_chroma(self, rec_planes, comp, x, y, w, h, bitdepth):
        """LM chroma prediction of component comp (1 / 2) of the CU at luma (x, y, w, h) from tight reconstructed planes."""
        luma, chroma = rec_planes[0], rec_planes[comp]
        pred = np.zeros((h // 2, w // 2), dtype=np.uint16)
        lp = luma.ctypes.data + 2 * (y * luma.shape[1] + x)
        cp = chroma.ctypes.data + 2 * ((y // 2) * chroma.shape[1] + x // 2)
        self.L.xo_intra_lm_chroma(x, y, w, h, bitdepth, c_void_p(lp), luma.shape[1], c_void_p(cp), chroma.shape[1], abi.ptr(pred), w // 2)
        return pred

    def tq_reconstruct(self, orig, pred, rec, bitdepth, cus, intra_picture=0, table=1, off_u=0, off_v=0):
        levels = [np.zeros((orig.height[c], orig.width[c]), dtype=np.int16) for c in range(3)]
        res = np.zeros(3 * len(cus), dtype=abi.tu_result_dtype)
        o, p, r = orig.c_struct(), pred.c_struct(), rec.c_struct()
        lv = abi.plane_ptr_array(levels)
        self.L.xo_tq_reconstruct(ctypes.byref(o), ctypes.byref(p), ctypes.byref(r), lv, bitdepth, abi.ptr(cus), len(cus),
                                 intra_picture, table, off_u, off_v, abi.ptr(res))
        return levels, res

    def dequant_reconstruct(self, pred, rec, levels, bitdepth, cus, table=1, off_u=0, off_v=0):
        p, r = pred.c_struct(), rec.c_struct()
        lv = abi.plane_ptr_array(levels)
        self.L.xo_dequant_reconstruct(ctypes.byref(p), ctypes.byref(r), lv, bitdepth, abi.ptr(cus), len(cus), table, off_u, off_v)

    def deblock_picture(self, rec, bitdepth, cus, pic_type, ref_poc, beta_offset=0, tc_offset=0, table=1, off_u=0, off_v=0):
        r = rec.c_struct()
        poc = _poc_array(ref_poc)
        self.L.xo_deblock_picture(ctypes.byref(r), bitdepth, abi.ptr(cus), len(cus), pic_type, beta_offset, tc_offset,
                                  table, off_u, off_v, poc)

    def deblock_band(self, rec, bitdepth, cus, pic_type, ref_poc, pass_mask, y_begin, y_end, beta_offset=0, tc_offset=0,
                     table=1, off_u=0, off_v=0):
        r = rec.c_struct()
        poc = _poc_array(ref_poc)
        self.L.xo_deblock_band(ctypes.byref(r), bitdepth, abi.ptr(cus), len(cus), pic_type, beta_offset, tc_offset,
                               table, off_u, off_v, poc, pass_mask, y_begin, y_end)

    def encode_picture(self, orig, refs, pred, rec, bitdepth, cus, params):
        levels = [np.zeros((orig.height[c], orig.width[c]), dtype=np.int16) for c in range(3)]
        nl = 2 if params["pic_type"] == 0 else 1
        me = np.zeros(nl * len(cus), dtype=abi.me_result_dtype)
        tu = np.zeros(3 * len(cus), dtype=abi.tu_result_dtype)
        o, p, r = orig.c_struct(), pred.c_struct(), rec.c_struct()
        arr, keep = _refs_array(refs)
        lv = abi.plane_ptr_array(levels)
        prm = np.array([params], dtype=abi.picture_params_dtype) if not isinstance(params, np.ndarray) else params
        self.L.xo_encode_picture(ctypes.byref(o), ctypes.addressof(arr), ctypes.byref(p), ctypes.byref(r), lv, bitdepth,
                                 abi.ptr(cus), len(cus), abi.ptr(prm), abi.ptr(me), abi.ptr(tu))
        return levels, me, tu


def have_ref():
    return os.path.exists(REF_SO)


class Ref:
    """The compiled, unmodified reference behind oracle/ref_shim.cc."""

    def __init__(self):
        # where ref_shim.cc's use_simd = 2 tables find the CUDA entries (resolved with dlopen at first use)
        os.environ.setdefault("XVCB200_LIB", os.path.join(os.path.dirname(_HERE), "xvc_b200", "libxvc_b200.so"))
        L = self.L = ctypes.CDLL(REF_SO)
        L.xref_ssd.restype = c_u64
        L.xref_compare.restype = c_u64
        L.xref_transform_matrix.restype = ctypes.POINTER(ctypes.c_int16)
        L.xref_table_entries_replaced.argtypes = [c_int, c_int]
        L.xref_intra_lm_chroma.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p]
        L.xref_session_create.restype = c_void_p
        L.xref_session_create.argtypes = [c_int, c_int, c_int, c_int, c_int, c_double, c_int, c_i64, c_int, c_int, c_int, c_int]
        L.xref_intra_scan.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        L.xref_conf_create.restype = c_void_p
        L.xref_conf_create.argtypes = [c_int, c_int, c_int, c_int]
        L.xref_conf_destroy.argtypes = [c_void_p]
        L.xref_conf_push_picture.argtypes = [c_void_p, c_void_p]
        L.xref_conf_flush.argtypes = [c_void_p]
        L.xref_conf_inter_inputs.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        L.xref_conf_write_inter.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]
        L.xref_conf_bitstream.restype = ctypes.c_size_t
        L.xref_conf_bitstream.argtypes = [c_void_p, c_void_p, ctypes.c_size_t]
        L.xref_qp_info.argtypes = [c_int, c_int, c_double, c_int, c_int, c_int, c_void_p]
        L.xref_sad.argtypes = [c_int, c_int, c_int, c_int, c_int, c_void_p, c_ssize, c_void_p, c_ssize]
        L.xref_ssd.argtypes = [c_int, c_int, c_int, c_int, c_int, c_void_p, c_ssize, c_void_p, c_ssize]
        L.xref_compare.argtypes = [c_int] * 8 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xref_filter.argtypes = [c_int] * 6 + [c_void_p, c_void_p, c_ssize, c_void_p, c_ssize]
        L.xref_interp.argtypes = [c_int] * 8 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xref_add_avg.argtypes = [c_int] * 6 + [c_void_p, c_ssize, c_void_p, c_ssize, c_void_p, c_ssize]
        L.xref_filter_copy_bipred.argtypes = [c_int] * 6 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xref_fwd_transform.argtypes = [c_int] * 7 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xref_inv_transform.argtypes = [c_int] * 8 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xref_transform_skip.argtypes = [c_int] * 4 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xref_quant_fast.argtypes = [c_int] * 8 + [c_void_p, c_ssize, c_void_p, c_ssize]
        if hasattr(L, "xref_quant_rdo_frozen"):
            L.xref_quant_rdo_frozen.argtypes = [c_int] * 5 + [ctypes.c_double] + [c_int] * 3 + [c_void_p, c_ssize, c_void_p, c_ssize]
        L.xref_dequant.argtypes = [c_int] * 5 + [c_void_p, c_ssize, c_void_p, c_ssize]
        for name in ("xref_session_destroy", "xref_session_set_orig", "xref_session_set_rec", "xref_session_get_rec",
                     "xref_session_set_pred", "xref_session_get_pred", "xref_session_get_coeff"):
            getattr(L, name).argtypes = [c_void_p] + ([c_void_p] if name != "xref_session_destroy" else [])
        L.xref_session_add_ref.argtypes = [c_void_p, c_int, c_int, c_i64, c_void_p]
        L.xref_session_get_ref_padded.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p]
        L.xref_session_get_rec_padded.argtypes = [c_void_p, c_int, c_void_p]
        L.xref_session_set_cus.argtypes = [c_void_p, c_void_p, c_int]
        L.xref_session_get_cus.argtypes = [c_void_p, c_void_p, c_int]
        L.xref_deblock_picture_ext.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int]
        L.xref_session_set_tu_modes.argtypes = [c_void_p, c_void_p, c_void_p, c_int]
        L.xref_session_scan_orders.argtypes = [c_void_p, c_void_p]
        L.xref_encode_picture_mvp.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
        L.xref_search_motion_single.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        L.xref_me_search.argtypes = [c_void_p, c_void_p, c_int, c_double, c_int, c_void_p]
        L.xref_tz_search.argtypes = [c_void_p, c_void_p, c_int, c_double, c_void_p]
        L.xref_full_search.argtypes = [c_void_p, c_void_p, c_int, c_double, c_void_p]
        L.xref_motion_compensate.argtypes = [c_void_p, c_int]
        L.xref_motion_compensate_lic.argtypes = [c_void_p, c_void_p, c_int]
        L.xref_lic_neighbours.argtypes = [c_void_p, c_void_p, c_int]
        L.xref_motion_compensate_affine.argtypes = [c_void_p, c_void_p, c_int, c_int]
        L.xref_tq_reconstruct.argtypes = [c_void_p, c_int, c_void_p]
        L.xref_deblock_picture.argtypes = [c_void_p, c_int, c_int]
        L.xref_pad_border_rec.argtypes = [c_void_p]
        L.xref_encode_picture.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]

    def matrix(self, kind, n):
        p = self.L.xref_transform_matrix(kind, n)
        return None if not p else np.ctypeslib.as_array(p, shape=(n * n,)).reshape(n, n).copy()

    def sad(self, kind, a, b, w, h, bitdepth=10, simd=0, sa=None, sb=None):
        return self.L.xref_sad(kind, simd, bitdepth, w, h, abi.ptr(a), sa or a.shape[1], abi.ptr(b), sb or b.shape[1])

    def ssd(self, kind, a, b, w, h, bitdepth=10, simd=0):
        return self.L.xref_ssd(kind, simd, bitdepth, w, h, abi.ptr(a), a.shape[1], abi.ptr(b), b.shape[1])

    _METRIC = {abi.METRIC_SSD: 0, abi.METRIC_SATD: 1, abi.METRIC_SAD: 3, abi.METRIC_SAD_FAST: 4}

    def compare(self, metric, bitdepth, a, b, w, h, simd=0, comp=0, qp=32):
        return self.L.xref_compare(self._METRIC[metric], simd, bitdepth, comp, qp, w, h, int(a.dtype == np.int16),
                                   abi.ptr(a), a.shape[1], abi.ptr(b), b.shape[1])

    def filter(self, kind, chroma, w, h, bitdepth, taps, src, src_off, dst, simd=0):
        esz = src.dtype.itemsize
        sp = src.ctypes.data + esz * (src_off[0] * src.shape[1] + src_off[1])
        self.L.xref_filter(kind, chroma, simd, w, h, bitdepth, abi.ptr(taps), c_void_p(sp), src.shape[1], abi.ptr(dst), dst.shape[1])

    def interp(self, chroma, bipred, w, h, bitdepth, fx, fy, ref, ref_off, pred, simd=0):
        sp = ref.ctypes.data + 2 * (ref_off[0] * ref.shape[1] + ref_off[1])
        self.L.xref_interp(chroma, bipred, simd, w, h, bitdepth, fx, fy, c_void_p(sp), ref.shape[1], abi.ptr(pred), pred.shape[1])

    def add_avg(self, w, h, offset, shift, bitdepth, a, b, dst, simd=0):
        self.L.xref_add_avg(simd, w, h, offset, shift, bitdepth, abi.ptr(a), a.shape[1], abi.ptr(b), b.shape[1], abi.ptr(dst), dst.shape[1])

    def filter_copy_bipred(self, w, h, offset, shift, ref, pred, bitdepth=10, simd=0):
        self.L.xref_filter_copy_bipred(simd, bitdepth, w, h, offset, shift, abi.ptr(ref), ref.shape[1], abi.ptr(pred), pred.shape[1])

    def fwd_transform(self, w, h, bitdepth, tx_hor, tx_ver, resi, comp=0, intra=0):
        out = np.zeros((h, w), dtype=np.int16)
        self.L.xref_fwd_transform(w, h, bitdepth, comp, intra, tx_hor, tx_ver, abi.ptr(resi), resi.shape[1], abi.ptr(out), w)
        return out

    def inv_transform(self, w, h, bitdepth, tx_hor, tx_ver, dc_only, coeff, comp=0, intra=0):
        out = np.zeros((h, w), dtype=np.int16)
        self.L.xref_inv_transform(w, h, bitdepth, comp, intra, tx_hor, tx_ver, dc_only, abi.ptr(coeff), coeff.shape[1], abi.ptr(out), w)
        return out

    def transform_skip(self, forward, w, h, bitdepth, inp):
        out = np.zeros((h, w), dtype=np.int16)
        self.L.xref_transform_skip(forward, w, h, bitdepth, abi.ptr(inp), inp.shape[1], abi.ptr(out), w)
        return out

    def qp(self, qp, bitdepth, lam=1.0, table=1, off_u=0, off_v=0):
        q = np.zeros(1, dtype=abi.qp_dtype)
        self.L.xref_qp_info(qp, bitdepth, lam, table, off_u, off_v, abi.ptr(q))
        return q[0]

    def quant_fast(self, w, h, bitdepth, comp, qp, intra_pic, coeff, intra_cu=0, intra_mode=0):
        out = np.zeros((h, w), dtype=np.int16)
        nz = self.L.xref_quant_fast(w, h, bitdepth, comp, qp, intra_pic, intra_cu, intra_mode, abi.ptr(coeff), coeff.shape[1], abi.ptr(out), w)
        return out, nz

    def quant_rdo_frozen(self, w, h, bitdepth, comp, qp, lam, intra_pic, coeff, intra_cu=0, intra_mode=0):
        """RdoQuant::QuantRdo against the context state a picture starts from (never advanced): the definition of
        'RDOQ with frozen contexts' (SURVEY 8(f) rank 4)."""
        out = np.zeros((h, w), dtype=np.int16)
        nz = self.L.xref_quant_rdo_frozen(w, h, bitdepth, comp, qp, float(lam), intra_pic, intra_cu, intra_mode,
                                          abi.ptr(coeff), coeff.shape[1], abi.ptr(out), w)
        return out, nz

    def dequant(self, w, h, bitdepth, comp, qp, lev):
        out = np.zeros((h, w), dtype=np.int16)
        self.L.xref_dequant(w, h, bitdepth, comp, qp, abi.ptr(lev), lev.shape[1], abi.ptr(out), w)
        return out

    def session(self, *args, **kw):
        return RefSession(self, *args, **kw)


class RefConformance:
    """A two-picture xvc bitstream (key picture by the reference encoder, inter picture written by the
    reference's CuWriter from OUTSIDE decisions) for the reference decoder to verify; see
    ref_shim.cc 'bitstream conformance'."""

    def __init__(self, ref, width, height, bitdepth=10, qp=32):
        self.L = ref.L
        self.width, self.height, self.bitdepth = width, height, bitdepth
        self.shapes = [(height, width), (height // 2, width // 2), (height // 2, width // 2)]
        self.h = self.L.xref_conf_create(width, height, bitdepth, qp)
        if not self.h:
            raise RuntimeError("xref_conf_create failed")

    def close(self):
        if self.h:
            self.L.xref_conf_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def push(self, planes):
        planes = [np.ascontiguousarray(p, dtype=np.uint16) for p in planes]
        return self.L.xref_conf_push_picture(self.h, abi.plane_ptr_array(planes))

    def flush(self):
        return self.L.xref_conf_flush(self.h)

    def inter_inputs(self, poc, ref_poc):
        orig = [np.zeros(s, dtype=np.uint16) for s in self.shapes]
        rec = [np.zeros(s, dtype=np.uint16) for s in self.shapes]
        info = np.zeros(8, dtype=np.int32)
        lam = ctypes.c_double(0)
        if self.L.xref_conf_inter_inputs(self.h, poc, ref_poc, abi.plane_ptr_array(orig), abi.plane_ptr_array(rec), abi.ptr(info),
                                         ctypes.byref(lam)) != 0:
            raise RuntimeError("picture not found")
        keys = ("qp", "chroma_table", "off_u", "off_v", "deblock", "beta_offset", "tc_offset", "pic_type")
        return orig, rec, dict(zip(keys, [int(v) for v in info]), lam=lam.value)

    def write_inter(self, poc, cus, splits, levels, rec):
        cus = np.ascontiguousarray(cus, dtype=abi.cu_dtype)
        splits = np.ascontiguousarray(splits, dtype=np.uint8)
        levels = [np.ascontiguousarray(p, dtype=np.int16) for p in levels]
        rec = [np.ascontiguousarray(p, dtype=np.uint16) for p in rec]
        return self.L.xref_conf_write_inter(self.h, poc, abi.ptr(cus), len(cus), abi.ptr(splits), len(splits),
                                            abi.plane_ptr_array(levels), abi.plane_ptr_array(rec))

    def bitstream(self):
        n = self.L.xref_conf_bitstream(self.h, None, 0)
        buf = np.zeros(n, dtype=np.uint8)
        self.L.xref_conf_bitstream(self.h, abi.ptr(buf), n)
        return buf.tobytes()


class RefSession:
    def __init__(self, ref, width, height, bitdepth=10, pic_type=0, qp=32, lam=1.0, simd=1, poc=8, sub_gop=16,
                 table=1, off_u=0, off_v=0):
        self.L = ref.L
        self.width, self.height = width, height
        self.h = self.L.xref_session_create(width, height, bitdepth, pic_type, qp, lam, simd, poc, sub_gop, table, off_u, off_v)
        self.shapes = [(height, width), (height // 2, width // 2), (height // 2, width // 2)]

    def close(self):
        if self.h:
            self.L.xref_session_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _set(self, fn, planes):
        planes = [np.ascontiguousarray(p, dtype=np.uint16) for p in planes]
        fn(self.h, abi.plane_ptr_array(planes))

    def _get(self, fn, dtype=np.uint16):
        planes = [np.zeros(s, dtype=dtype) for s in self.shapes]
        fn(self.h, abi.plane_ptr_array(planes))
        return planes

    def set_orig(self, planes): self._set(self.L.xref_session_set_orig, planes)
    def set_rec(self, planes): self._set(self.L.xref_session_set_rec, planes)
    def set_pred(self, planes): self._set(self.L.xref_session_set_pred, planes)
    def get_rec(self): return self._get(self.L.xref_session_get_rec)
    def get_pred(self): return self._get(self.L.xref_session_get_pred)
    def get_coeff(self): return self._get(self.L.xref_session_get_coeff, np.int16)

    def add_ref(self, lst, idx, poc, planes):
        planes = [np.ascontiguousarray(p, dtype=np.uint16) for p in planes]
        self.L.xref_session_add_ref(self.h, lst, idx, poc, abi.plane_ptr_array(planes))

    def get_ref_padded(self, lst, idx, comp):
        pad = 80 if comp == 0 else 40
        h, w = self.shapes[comp]
        out = np.zeros((h + 2 * pad, w + 2 * pad), dtype=np.uint16)
        self.L.xref_session_get_ref_padded(self.h, lst, idx, comp, abi.ptr(out))
        return out

    def get_rec_padded(self, comp):
        pad = 80 if comp == 0 else 40
        h, w = self.shapes[comp]
        out = np.zeros((h + 2 * pad, w + 2 * pad), dtype=np.uint16)
        self.L.xref_session_get_rec_padded(self.h, comp, abi.ptr(out))
        return out

    def set_cus(self, cus):
        self.L.xref_session_set_cus(self.h, abi.ptr(cus), len(cus))

    def get_cus(self, cus):
        out = cus.copy()
        self.L.xref_session_get_cus(self.h, abi.ptr(out), len(out))
        return out

    def me_search(self, jobs, lam, threads=1):
        res = np.zeros(len(jobs), dtype=abi.me_result_dtype)
        self.L.xref_me_search(self.h, abi.ptr(jobs), len(jobs), lam, threads, abi.ptr(res))
        return res

    def tz_search(self, jobs, lam):
        mv = np.zeros((len(jobs), 2), dtype=np.int32)
        self.L.xref_tz_search(self.h, abi.ptr(jobs), len(jobs), lam, abi.ptr(mv))
        return mv

    def full_search(self, jobs, lam):
        res = np.zeros(len(jobs), dtype=abi.me_result_dtype)
        self.L.xref_full_search(self.h, abi.ptr(jobs), len(jobs), lam, abi.ptr(res))
        return res

    def motion_compensate(self, threads=1):
        self.L.xref_motion_compensate(self.h, threads)

    def motion_compensate_affine(self, aff, threads=1):
        aff = np.ascontiguousarray(aff, dtype=abi.affine_cu_dtype)
        self.L.xref_motion_compensate_affine(self.h, abi.ptr(aff), len(aff), threads)

    def motion_compensate_lic(self, lic):
        lic = np.ascontiguousarray(lic, dtype=abi.lic_cu_dtype)
        self.L.xref_motion_compensate_lic(self.h, abi.ptr(lic), len(lic))

    def lic_neighbours(self, cu_indices):
        lic = np.zeros(len(cu_indices), dtype=abi.lic_cu_dtype)
        lic["cu"] = cu_indices
        self.L.xref_lic_neighbours(self.h, abi.ptr(lic), len(lic))
        return lic

    def intra_lm_chroma(self, cus):
        """[(U block, V block)] per CU: IntraPrediction::Predict(kLmChroma) from the session's reconstruction."""
        cus = np.ascontiguousarray(cus, dtype=abi.cu_dtype)
        total = int(sum((int(c["w"]) // 2) * (int(c["h"]) // 2) for c in cus))
        pu, pv = np.zeros(total, dtype=np.uint16), np.zeros(total, dtype=np.uint16)
        self.L.xref_intra_lm_chroma(self.h, abi.ptr(cus), len(cus), abi.ptr(pu), abi.ptr(pv))
        out, off = [], 0
        for c in cus:
            w, h = int(c["w"]) // 2, int(c["h"]) // 2
            out.append((pu[off:off + w * h].reshape(h, w).copy(), pv[off:off + w * h].reshape(h, w).copy()))
            off += w * h
        return out

    def tq_reconstruct(self, n_cus, threads=1):
        res = np.zeros(3 * n_cus, dtype=abi.tu_result_dtype)
        self.L.xref_tq_reconstruct(self.h, threads, abi.ptr(res))
        return res

    def deblock_picture(self, beta_offset=0, tc_offset=0):
        self.L.xref_deblock_picture(self.h, beta_offset, tc_offset)

    def deblock_picture_ext(self, beta_offset=0, tc_offset=0, affine=None, chroma_cus=None):
        """DeblockPicture with affine CUs (abi.affine_cu_dtype) and / or the secondary CU tree (abi.cu_dtype) of an intra picture."""
        if affine is not None:
            affine = np.ascontiguousarray(affine, dtype=abi.affine_cu_dtype)
        if chroma_cus is not None:
            chroma_cus = np.ascontiguousarray(chroma_cus, dtype=abi.cu_dtype)
            assert self.L.xref_session_has_secondary_tree(self.h), "the session's picture has no secondary CU tree (not an intra picture)"
        self.L.xref_deblock_picture_ext(self.h, beta_offset, tc_offset, abi.ptr(affine), 0 if affine is None else len(affine),
                                        abi.ptr(chroma_cus), 0 if chroma_cus is None else len(chroma_cus))

    def intra_scan(self, cus, comp=0, want_pred=True):
        """CUs in coding order -> (jobs, ref_samples, ref_filtered, predictions per CU [67][h][w], satd [n][67])."""
        n = len(cus)
        sh = 1 if comp else 0
        jobs = np.zeros(n, dtype=abi.intra_job_dtype)
        ref = np.zeros((n, 2 * abi.INTRA_REF_STRIDE), dtype=np.uint16)
        filt = np.zeros((n, 2 * abi.INTRA_REF_STRIDE), dtype=np.uint16)
        sizes = [abi.INTRA_NUM_MODES * int(c["w"] >> sh) * int(c["h"] >> sh) for c in cus]
        pred = np.zeros(sum(sizes), dtype=np.uint16) if want_pred else None
        satd = np.zeros((n, abi.INTRA_NUM_MODES), dtype=np.uint32)
        self.L.xref_intra_scan(self.h, abi.ptr(np.ascontiguousarray(cus)), n, comp, abi.ptr(jobs), abi.ptr(ref), abi.ptr(filt),
                               abi.ptr(pred), abi.ptr(satd))
        preds = None
        if want_pred:
            preds, off = [], 0
            for c, sz in zip(cus, sizes):
                preds.append(pred[off:off + sz].reshape(abi.INTRA_NUM_MODES, int(c["h"] >> sh), int(c["w"] >> sh)))
                off += sz
        return jobs, ref, filt, preds, satd

    def pad_border_rec(self):
        self.L.xref_pad_border_rec(self.h)

    def set_tu_modes(self, modes, intra_luma_modes=None):
        """After set_cus: transform types / transform skip per CU and the luma intra mode (uint8 per CU) the
        reference derives the coefficient scan from."""
        n = len(modes) if modes is not None else len(intra_luma_modes)
        if modes is not None:
            modes = np.ascontiguousarray(modes, dtype=abi.tu_mode_dtype)
        if intra_luma_modes is not None:
            intra_luma_modes = np.ascontiguousarray(intra_luma_modes, dtype=np.uint8)
        self.L.xref_session_set_tu_modes(self.h, abi.ptr(modes), abi.ptr(intra_luma_modes), n)

    def scan_orders(self, n):
        out = np.zeros((n, 3), dtype=np.uint8)
        self.L.xref_session_scan_orders(self.h, abi.ptr(out))
        return out

    def search_motion_single(self, params, cu):
        """(ref_shim's restated SearchMotion, InterSearch::SearchMotion itself) for a picture holding only `cu`."""
        out = np.zeros(2, dtype=abi.cu_dtype)
        costs = np.zeros(2, dtype=np.uint64)
        one = np.array([cu], dtype=abi.cu_dtype)
        prm = np.array([params], dtype=abi.picture_params_dtype) if not isinstance(params, np.ndarray) else params
        self.L.xref_search_motion_single(self.h, abi.ptr(prm), abi.ptr(one), abi.ptr(out), abi.ptr(costs))
        return out[0], out[1]

    def encode_picture(self, params, cus, threads=1, mvp=None):
        """mvp: optional int32 [n][columns][2] predictors per (CU, list, reference picture)."""
        me = np.zeros(abi.num_me_columns(params) * len(cus), dtype=abi.me_result_dtype)
        tu = np.zeros(3 * len(cus), dtype=abi.tu_result_dtype)
        out = cus.copy()
        prm = np.array([params], dtype=abi.picture_params_dtype) if not isinstance(params, np.ndarray) else params
        if mvp is not None:
            mvp = np.ascontiguousarray(mvp, dtype=np.int32)
            assert mvp.shape == (len(cus), abi.num_me_columns(params), 2)
        self.L.xref_encode_picture_mvp(self.h, abi.ptr(prm), abi.ptr(cus), len(cus), abi.ptr(mvp), threads, abi.ptr(me), abi.ptr(tu), abi.ptr(out))
        return me, tu, out
