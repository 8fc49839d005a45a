"""ctypes binding of libesr_b200.so (declared in include/esr_b200.h).

The product path has no fallback: if the library is missing or the device is not
sm_100, importing callers get an exception, never a silent PyTorch path.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ESR_LIB_PATH") or os.path.join(HERE, "libesr_b200.so")   # override: A/B timing of builds

MAX_KBLOCKS = 24
MAX_COUT_TILES = 8
KBLOCK_CH = 32
CEM_MAX_TAPS = 64

EPI_LRELU, EPI_RES1, EPI_RES2, EPI_ACCUM, EPI_MASK, EPI_F32_BLOCKED = 1, 2, 4, 8, 16, 32
CONV_F16, EPI_OUT_F16, EPI_RES1_HILO = 64, 128, 256


class KBlock(C.Structure):
    _fields_ = [("src", C.c_int32), ("chan", C.c_int32), ("w_off", C.c_uint32), ("dy_mask", C.c_uint8),
                ("slice_mask", C.c_uint8), ("n_dy", C.c_uint8), ("half", C.c_uint8)]


class TensorNHWC(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("channels", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("src", TensorNHWC * 2),
        ("cout_tile", C.c_int32), ("cout_tiles", C.c_int32), ("num_kblocks", C.c_int32),
        ("kblocks", KBlock * MAX_KBLOCKS),
        ("wpack", C.c_void_p), ("w_tile_bytes", C.c_uint32), ("bias", C.c_void_p),
        ("flags", C.c_uint32), ("slope", C.c_float), ("alpha", C.c_float), ("beta", C.c_float),
        ("res1", C.c_void_p), ("res1_stride", C.c_int32), ("res1_choff", C.c_int32),
        ("res2", C.c_void_p), ("res2_stride", C.c_int32), ("res2_choff", C.c_int32),
        ("out_bf16", C.c_void_p), ("out_bf16_stride", C.c_int32), ("out_bf16_choff", C.c_int32),
        ("out_bf16_lo_choff", C.c_int32), ("up", C.c_int32), ("out_bf16_scale", C.c_float),
        ("out_f32", C.c_void_p), ("out_f32_stride", C.c_int32), ("out_f32_choff", C.c_int32),
        ("out_nchw", C.c_void_p), ("cout_real", C.c_int32),
        ("mask", C.c_void_p), ("mask_stride", C.c_int32), ("mask_choff", C.c_int32),
        ("tile_choff", C.c_int16 * MAX_COUT_TILES),
        ("no_accum_tiles", C.c_uint16), ("no_bf16_tiles", C.c_uint16), ("no_res_tiles", C.c_uint16),
        ("pair", C.c_uint16), ("gamma", C.c_float),
        ("res1_hi", C.c_void_p), ("res1_hi_stride", C.c_int32), ("res1_hi_choff", C.c_int32),
        ("res1_lo", C.c_void_p), ("res1_lo_stride", C.c_int32), ("res1_lo_choff", C.c_int32),
        ("out_lo", C.c_void_p), ("out_lo_stride", C.c_int32), ("out_lo_choff", C.c_int32),
    ]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for i in range(MAX_COUT_TILES):
            self.tile_choff[i] = -1
        self.gamma, self.alpha, self.beta, self.slope = 1.0, 1.0, 1.0, 0.2
        self.up, self.out_bf16_scale, self.out_bf16_lo_choff = 1, 1.0, -1


RDB_MAX_LAYERS, RDB_MAX_KBLOCKS = 4, 8


class RdbLayer(C.Structure):
    _fields_ = [("num_kblocks", C.c_int32), ("kblocks", KBlock * RDB_MAX_KBLOCKS), ("wpack", C.c_void_p),
                ("w_tile_bytes", C.c_uint32), ("bias", C.c_void_p), ("out_choff", C.c_int32)]


class RdbGrowthDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("src", TensorNHWC * 2),
                ("num_layers", C.c_int32), ("layers", RdbLayer * RDB_MAX_LAYERS),
                ("out", C.c_void_p), ("out_stride", C.c_int32), ("mode", C.c_int32), ("slope", C.c_float),
                ("mask", C.c_void_p), ("mask_stride", C.c_int32), ("imgs_per_chunk", C.c_int32),
                ("flags", C.c_void_p), ("flags_use", C.c_int32), ("flags_zero", C.c_int32)]


class WRow(C.Structure):
    _fields_ = [("idx", C.c_int16), ("ky", C.c_int8), ("reserved", C.c_int8)]


class WSlot(C.Structure):
    _fields_ = [("idx", C.c_int16), ("ky", C.c_int8), ("term", C.c_int8)]


class XSlot(C.Structure):
    _fields_ = [("c", C.c_int8), ("dy", C.c_int8), ("term", C.c_int8), ("reserved", C.c_int8)]


class CemFilters(C.Structure):
    _fields_ = [("sf", C.c_int32), ("pre", C.c_int32), ("n_ds", C.c_int32), ("n_inv", C.c_int32),
                ("ds", C.c_float * CEM_MAX_TAPS), ("inv", C.c_float * CEM_MAX_TAPS)]


class CemFilters2d(C.Structure):
    _fields_ = [("sf", C.c_int32), ("pre", C.c_int32), ("n_ds", C.c_int32), ("n_inv", C.c_int32),
                ("ds", C.c_void_p), ("inv", C.c_void_p)]


CEM2D_MAX_SIDE = 63

class CopySeg(C.Structure):              # esr_copy_seg
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int32), ("row_elems", C.c_int32),
                ("src_pitch", C.c_int32), ("dst_pitch", C.c_int32)]


class WgradItem(C.Structure):            # esr_wgrad_item
    _fields_ = [("x", C.c_void_p), ("g", C.c_void_p), ("dw", C.c_void_p),
                ("x_stride", C.c_int32), ("x_c0", C.c_int32), ("x_f16", C.c_int32),
                ("g_stride", C.c_int32), ("g_c0", C.c_int32), ("cout", C.c_int32),
                ("n_co", C.c_int32), ("n_ci", C.c_int32), ("ci_lo", C.c_int32), ("cin_total", C.c_int32), ("ci0", C.c_int32),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("tile_begin", C.c_int32), ("tile_end", C.c_int32),
                ("db", C.c_void_p)]


class WgradTcItem(C.Structure):          # esr_wgrad_tc_item
    _fields_ = [("x_map", C.c_uint32), ("g_map", C.c_uint32), ("x_c0", C.c_int32), ("g_c0", C.c_int32), ("dw", C.c_void_p),
                ("n_ci", C.c_int32), ("n_co", C.c_int32), ("cin_total", C.c_int32), ("ci0", C.c_int32),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("tile_begin", C.c_int32), ("tile_end", C.c_int32),
                ("x_f16", C.c_int32)]


class WgradSmallItem(C.Structure):       # esr_wgrad_small_item
    _fields_ = [("g", C.c_void_p), ("g32", C.c_void_p), ("s", C.c_void_p), ("dw", C.c_void_p), ("db", C.c_void_p),
                ("g_stride", C.c_int32), ("g_c0", C.c_int32), ("cout", C.c_int32), ("n_co", C.c_int32),
                ("s_channels", C.c_int32), ("s_c0", C.c_int32), ("n_c", C.c_int32),
                ("cin_total", C.c_int32), ("ci0", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


# name -> (restype, argtypes); every symbol include/esr_b200.h declares
_i32, _i64, _vp, _f = C.c_int32, C.c_int64, C.c_void_p, C.c_float
SIGNATURES = {
    "esr_last_error": (C.c_char_p, []),
    "esr_abi_version": (C.c_int, []),
    "esr_device_check": (C.c_int, [C.c_int]),
    "esr_conv3x3_tc": (C.c_int, [C.POINTER(ConvDesc), _vp]),
    "esr_conv3x3_simt": (C.c_int, [C.POINTER(ConvDesc), _vp]),
    "esr_pack_layout": (_i64, [_i32, _i32, _i32, _i32, C.POINTER(KBlock), C.POINTER(C.c_uint32)]),
    "esr_pack_conv_weights": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _i32, _i32, _i32, _i32, C.POINTER(KBlock),
                                        C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "esr_pack_entry_bytes": (_i32, []),
    "esr_pack_entry_fill": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _i32, _i32, _i32, _i32, C.POINTER(KBlock),
                                      C.c_uint32, _vp, _vp, _vp, _vp]),
    "esr_pack_table_run": (C.c_int, [_vp, _i32, _vp]),
    "esr_copy_segments": (C.c_int, [_vp, _i32, _vp]),
    "esr_expand_rows": (C.c_int, [_vp, _i32, _i32, _i32, _i32, C.POINTER(XSlot), _i32, _vp, _vp]),
    "esr_expand_rows_bwd": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, C.POINTER(XSlot), _i32,
                                      _vp, _vp]),
    "esr_debug_set_profile_buffer": (None, [_vp]),
    "esr_debug_cem_timeout": (C.c_int, [C.POINTER(C.c_uint32)]),
    "esr_debug_cem_fused_prof": (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
    "esr_wgrad16": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "esr_wgrad16r": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "esr_wgrad_tc_map_bytes": (C.c_int32, []),
    "esr_wgrad_tc_make_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "esr_wgrad_tc": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "esr_wgrad_small": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "esr_zopt_tanh_pack": (C.c_int, [_vp, C.c_float, _i32, _i32, _i32, _vp, _vp]),
    "esr_zopt_loss_workspace_floats": (_i32, [_i32, _i32]),
    "esr_zopt_loss": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, C.c_float, C.c_float, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "esr_zopt_loss_grad": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "esr_zopt_adam": (C.c_int, [_vp, _vp, _vp, _vp, C.c_float, _i32, _i32, _i32, C.c_float, C.c_float, C.c_float, C.c_float, _vp, _vp]),
    "esr_kde_workspace_bytes": (_i64, [_i64, _i64]),
    "esr_kde_sums": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i32, C.c_double, C.c_double, C.c_double, _vp, _vp, _vp]),
    "esr_kde_grad_workspace_bytes": (_i64, [_i64, _i64, _i32]),
    "esr_kde_grad": (C.c_int, [_vp, _i64, _vp, _i64, _i32, C.c_double, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "esr_patch_select": (C.c_int, [_vp, _i64, _i32, C.c_double, _i64, _i64, _vp, _vp]),
    "esr_grad_combine": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i32,
                                   _i32, _i32, _f, _f, _vp, _i32, _i32, _i32, _vp]),
    "esr_g_input_prep": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "esr_cem_pad_input": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "esr_g_input_prep_bwd": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "esr_cem_downscale": (C.c_int, [C.POINTER(CemFilters), _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "esr_cem_inv_hth": (C.c_int, [C.POINTER(CemFilters), _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "esr_cem_upscale": (C.c_int, [C.POINTER(CemFilters), _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "esr_cem_project": (C.c_int, [C.POINTER(CemFilters), _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "esr_cem_project_fused": (C.c_int, [C.POINTER(CemFilters), _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "esr_cem_project_bwd": (C.c_int, [C.POINTER(CemFilters), _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "esr_cem2d_downscale": (C.c_int, [C.POINTER(CemFilters2d), _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "esr_cem2d_inv_hth": (C.c_int, [C.POINTER(CemFilters2d), _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "esr_cem2d_upscale": (C.c_int, [C.POINTER(CemFilters2d), _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "esr_cem2d_project": (C.c_int, [C.POINTER(CemFilters2d), _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "esr_cem2d_project_bwd": (C.c_int, [C.POINTER(CemFilters2d), _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "esr_seq_create": (_vp, []),
    "esr_seq_destroy": (None, [_vp]),
    "esr_seq_add_conv": (C.c_int, [_vp, C.POINTER(ConvDesc), _i32]),
    "esr_seq_add_rdb_growth": (C.c_int, [_vp, C.POINTER(RdbGrowthDesc)]),
    "esr_rdb_growth_flag_words": (_i64, [_i32, _i32, _i32]),
    "esr_rdb_growth_tc": (C.c_int, [C.POINTER(RdbGrowthDesc), _vp]),
    "esr_seq_run": (C.c_int, [_vp, _vp]),
    "esr_seq_num_launches": (_i32, [_vp]),
}

_lib = None


class EsrError(RuntimeError):
    pass


def lib():
    """Loads the shared library once; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EsrError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise EsrError("libesr_b200: %s (status %d)" % (lib().esr_last_error().decode(), rc))


_device_ok = {}


def require_device(device_index):
    if device_index not in _device_ok:
        check(lib().esr_device_check(device_index))
        _device_ok[device_index] = True


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class CemFilterBank2D:
    """General (non-separable) CEM filters: the 2-D taps live in device memory, uploaded once per
    device on first use (esr_cem_filters2d carries device pointers)."""

    def __init__(self, sf, pre, ds_2d, inv_2d):
        import numpy as np
        self.sf, self.pre = int(sf), int(pre)
        self.ds = np.ascontiguousarray(ds_2d, dtype=np.float32)
        self.inv = np.ascontiguousarray(inv_2d, dtype=np.float32)
        for name, k in (("ds_kernel", self.ds), ("inv_hTh", self.inv)):
            if k.ndim != 2 or k.shape[0] != k.shape[1] or k.shape[0] % 2 == 0 or k.shape[0] > CEM2D_MAX_SIDE:
                raise EsrError("%s of shape %s: the 2-D stencil path needs a square, odd-sided filter of at most %d taps a side"
                               % (name, k.shape, CEM2D_MAX_SIDE))
        self._per_device = {}

    def on(self, device_index):
        import torch
        if device_index not in self._per_device:
            dev = torch.device("cuda", device_index)
            ds, inv = torch.from_numpy(self.ds).to(dev), torch.from_numpy(self.inv).to(dev)
            f = CemFilters2d()
            f.sf, f.pre, f.n_ds, f.n_inv = self.sf, self.pre, self.ds.shape[0], self.inv.shape[0]
            f.ds, f.inv = ds.data_ptr(), inv.data_ptr()
            self._per_device[device_index] = (f, ds, inv)
        return self._per_device[device_index][0]


def cem_call(op, filters, *args):
    """Runs CEM operator `op` ('downscale', 'inv_hth', 'upscale', 'project', 'project_bwd') on the current
    device with either the separable (esr_cem_*) or the general 2-D (esr_cem2d_*) filters."""
    if isinstance(filters, CemFilterBank2D):
        import torch
        return check(getattr(lib(), "esr_cem2d_" + op)(filters.on(torch.cuda.current_device()), *args))
    return check(getattr(lib(), "esr_cem_" + op)(filters, *args))


def cem_filters_struct(sf, pre, ds_1d, inv_1d):
    f = CemFilters()
    f.sf, f.pre, f.n_ds, f.n_inv = int(sf), int(pre), len(ds_1d), len(inv_1d)
    if len(ds_1d) > CEM_MAX_TAPS or len(inv_1d) > CEM_MAX_TAPS:
        raise EsrError("CEM filter longer than %d taps" % CEM_MAX_TAPS)
    for i, v in enumerate(ds_1d):
        f.ds[i] = float(v)
    for i, v in enumerate(inv_1d):
        f.inv[i] = float(v)
    return f
