"""numpy / ctypes mirrors of the plain-C structs in include/xvc_b200.h.

Layouts are checked against the compiled library by tests/test_abi.py
(xvcb200_abi_sizeof) so a drift between the header and this file fails loudly.
"""
import ctypes

import numpy as np

# flags (xvcb200_cu.flags)
CU_INTRA, CU_FULLPEL_MV, CU_CBF_Y, CU_CBF_U, CU_CBF_V, CU_SKIP_ME = 1, 2, 4, 8, 16, 32
# metrics
METRIC_SSD, METRIC_SATD, METRIC_SAD, METRIC_SAD_FAST = 0, 1, 2, 3
# transform types (TransformType, cu_types.h:67-75)
TX_DEFAULT, TX_DCT2, TX_DCT5, TX_DCT8, TX_DST1, TX_DST7 = range(6)

cu_dtype = np.dtype([
    ("x", "<i2"), ("y", "<i2"), ("w", "u1"), ("h", "u1"), ("depth", "u1"), ("flags", "u1"),
    ("qp", "i1"), ("ref_idx", "i1", (2,)), ("tx_select", "u1"), ("mv", "<i4", (2, 2)),
], align=True)

affine_cu_dtype = np.dtype([("cu", "<i4"), ("mv", "<i4", (2, 3, 2))], align=True)    # xvcb200_affine_cu

lic_cu_dtype = np.dtype([("cu", "<i4"), ("above_x", "<i2"), ("above_y", "<i2"), ("left_x", "<i2"), ("left_y", "<i2")], align=True)    # xvcb200_lic_cu

me_job_dtype = np.dtype([
    ("cu", "<i4"), ("ref_slot", "<i4"), ("search_range", "<i4"), ("mvp", "<i4", (2,)),
    ("prev", "<i4", (2,)), ("list", "<i4"),
], align=True)

me_result_dtype = np.dtype([
    ("mv_fullpel", "<i4", (2,)), ("mv", "<i4", (2,)), ("cost_fullpel", "<u4"), ("dist", "<u4"),
    ("cost", "<u4"), ("num_sad", "<u4"),
], align=True)

fullsearch_job_dtype = np.dtype([
    ("cu", "<i4"), ("ref_slot", "<i4"), ("other_pred_slot", "<i4"), ("mvp", "<i4", (2,)),
    ("center", "<i4", (2,)), ("range", "<i4"),
], align=True)

tu_result_dtype = np.dtype([("ssd", "<u4"), ("num_non_zero", "<i4")], align=True)

picture_params_dtype = np.dtype([
    ("orig_slot", "<i4"), ("pred_slot", "<i4"), ("rec_slot", "<i4"), ("coeff_slot", "<i4"),
    ("ref_slots", "<i4", (2, 5)), ("ref_poc", "<i8", (2, 5)), ("num_ref", "<i4", (2,)),
    ("pic_type", "<i4"), ("search_range", "<i4", (2, 5)), ("lambda_sqrt", "<f8"),
    ("chroma_offset_table", "<i4"), ("chroma_offset_u", "<i4"), ("chroma_offset_v", "<i4"),
    ("beta_offset", "<i4"), ("tc_offset", "<i4"), ("deblock", "<i4"), ("pad", "<i4"),
    ("bi_iterations", "<i4"), ("bits_mode", "<i4"),
], align=True)

tu_mode_dtype = np.dtype([("tx_ver", "u1"), ("tx_hor", "u1"), ("tskip", "u1"), ("scan", "u1", (3,)), ("reserved", "u1", (2,))], align=True)    # xvcb200_tu_mode

partition_params_dtype = np.dtype([("orig_slot", "<i4"), ("ref_slot", "<i4"), ("center", "<i4", (2,)), ("lambda_sqrt", "<f8"), ("qp", "<i4"),
                                   ("header_bits_cu", "<i4"), ("header_bits_split", "<i4")], align=True)    # xvcb200_partition_params

plane_geom_dtype = np.dtype([
    ("width", "<i4", (3,)), ("height", "<i4", (3,)), ("pitch", "<i4", (3,)),
    ("margin_x", "<i4", (3,)), ("margin_y", "<i4", (3,)),
], align=True)

qp_dtype = np.dtype([
    ("qp_raw", "<i4", (3,)), ("qp_bitdepth", "<i4", (3,)), ("distortion_weight", "<f8", (3,)),
    ("lambda", "<f8", (3,)), ("lambda_sqrt", "<f8"),
], align=True)

# intra prediction (IntraPrediction::RefState: 2 x 129 samples; 67 modes)
INTRA_REF_STRIDE, INTRA_NUM_MODES = 129, 67
intra_job_dtype = np.dtype([
    ("x", "<i4"), ("y", "<i4"), ("w", "u1"), ("h", "u1"), ("has_above_left", "u1"), ("has_above", "u1"),
    ("has_left", "u1"), ("above_right", "u1"), ("below_left", "u1"), ("reserved", "u1"),
], align=True)

ABI_STRUCTS = {
    0: ("xvcb200_cu", cu_dtype),
    1: ("xvcb200_me_job", me_job_dtype),
    2: ("xvcb200_me_result", me_result_dtype),
    3: ("xvcb200_fullsearch_job", fullsearch_job_dtype),
    4: ("xvcb200_tu_result", tu_result_dtype),
    5: ("xvcb200_picture_params", picture_params_dtype),
    6: ("xvcb200_plane_geom", plane_geom_dtype),
    7: ("xvcb200_qp", qp_dtype),
    8: ("xvcb200_intra_job", intra_job_dtype),
    9: ("xvcb200_affine_cu", affine_cu_dtype),
    10: ("xvcb200_lic_cu", lic_cu_dtype),
    11: ("xvcb200_tu_mode", tu_mode_dtype),
    12: ("xvcb200_partition_params", partition_params_dtype),
}


def num_me_columns(params):
    """Search results per CU of xvcb200_encode_picture: one per (list, reference picture)."""
    prm = params[0] if isinstance(params, np.ndarray) and params.ndim else params
    r0 = max(int(prm["num_ref"][0]), 1)
    r1 = 0 if int(prm["pic_type"]) == 1 else max(int(prm["num_ref"][1]), 1)
    return r0 + r1


def ptr(arr):
    """void* to the first byte of a (contiguous) numpy array, or NULL for None."""
    if arr is None:
        return ctypes.c_void_p(0)
    assert arr.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(arr.ctypes.data)


def plane_ptr_array(planes):
    """uint16_t* planes[3] from three numpy arrays."""
    arr = (ctypes.c_void_p * 3)()
    for i, p in enumerate(planes):
        arr[i] = p.ctypes.data
    return arr
