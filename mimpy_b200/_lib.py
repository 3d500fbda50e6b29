"""ctypes binding of libmimpy_b200.so (include/mimpy_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is present every entry
point raises.  PyTorch is used only to own device buffers and streams.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmimpy_b200.so")

MIMPY_E_NOCONV = -4
MIMPY_E_UNSUPPORTED = -3

_vp = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f64 = C.c_double


class MeshView(C.Structure):
    _fields_ = [("nc", _i32), ("nf", _i32), ("max_faces", _i32), ("uniform_faces", _i32),
                ("hex_nx", _i32), ("hex_ny", _i32), ("hex_nz", _i32), ("reserved0", _i32),
                ("cell_ptr", _vp), ("cell_faces", _vp), ("cell_sigma", _vp), ("face_geom", _vp),
                ("cell_k", _vp), ("cell_centroid", _vp), ("cell_volume", _vp)]


class Symbolic(C.Structure):
    _fields_ = [("flux_dof", _i32), ("nnz", _i64), ("nnz_m", _i64), ("n_blocks", _i64),
                ("face_to_flux", _vp), ("face_cell", _vp), ("indptr", _vp), ("indices", _vp),
                ("m_indptr", _vp), ("m_indices", _vp), ("m_ptr", _vp), ("pos_m", _vp),
                ("pos_dt", _vp), ("pos_d", _vp), ("diag_pos", _vp), ("role", _vp),
                ("cell_rowstart", _vp), ("face_flag", _vp), ("serial", _i64)]


class PcgInfo(C.Structure):
    _fields_ = [("rtol", _f64), ("atol", _f64), ("maxit", _i32), ("check_every", _i32),
                ("use_graph", _i32), ("iterations", _i32), ("residual", _f64), ("rhs_norm", _f64),
                ("ms_total", C.c_float), ("norm_kind", _i32), ("ref_norm", _f64), ("rz0", _f64),
                ("rz", _f64), ("sell_pat_id", _vp), ("sell_pat_table", _vp), ("sell_npat", _i32)]


class Fluid(C.Structure):
    _fields_ = [("mu_w", _f64), ("mu_o", _f64), ("rho_w", _f64), ("rho_o", _f64), ("c_w", _f64),
                ("c_o", _f64), ("s_wr", _f64), ("s_or", _f64), ("n_w", _f64), ("n_o", _f64),
                ("k_rw_o", _f64)]


_PM = C.POINTER(MeshView)
_PS = C.POINTER(Symbolic)

# name -> argtypes (restype is always int unless noted)
SIGNATURES = {
    "mimpy_b200_version": [],
    "mimpy_b200_create": [C.c_int, _vp, C.POINTER(_vp)],
    "mimpy_b200_destroy": [_vp],
    "mimpy_b200_set_stream": [_vp, _vp],
    "mimpy_b200_sync": [_vp],
    "mimpy_b200_hex_topology": [_vp, _i32, _i32, _i32, _f64, _f64, _f64, _vp, _vp, _vp, _vp],
    "mimpy_b200_geom_hex_faces": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp],
    "mimpy_b200_face_pack": [_vp, _i32, _vp, _vp, _vp, _vp],
    "mimpy_b200_geom_cells": [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mimpy_b200_mfd_local_matrices": [_vp, _PM, _i32, _i32, _vp, _vp, _vp, _vp],
    "mimpy_b200_mfd_symbolic_sizes": [_vp, _PM, _vp, _PS],
    "mimpy_b200_mfd_symbolic_fill": [_vp, _PM, _PS],
    "mimpy_b200_mfd_numeric": [_vp, _PM, _PS, _i32, _i32, _vp, _f64, _vp, _vp, _vp],
    "mimpy_b200_mfd_numeric_layers": [_vp, _PM, _PS, _i32, _i32, _vp, _f64, _vp, _vp, _i32, _i32],
    "mimpy_b200_mfd_numeric_from_blocks": [_vp, _PM, _PS, _vp, _vp, _f64, _vp, _vp],
    "mimpy_b200_mfd_coo": [_vp, _PM, _PS, _vp, _vp, _vp, _vp, _vp],
    "mimpy_b200_mfd_rhs": [_vp, _i32, _i32, _vp, _i32, _vp, _vp, _i32, _f64, _vp, _vp],
    "mimpy_b200_singlephase_terms": [_vp, _i32, _i32, _f64, _f64, _f64, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp],
    "mimpy_b200_transport_step": [_vp, _PM, _PS, _vp, _vp, _vp, _vp, _vp, _vp, _f64, _f64, _f64, _f64, _vp],
    "mimpy_b200_hybrid_setup": [_vp, _PM, _PS, _vp, _vp, _f64, _vp, _vp, _vp, _vp, _vp],
    "mimpy_b200_hybrid_setup_fused": [_vp, _PM, _PS, _i32, _i32, _vp, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i32)],
    "mimpy_b200_sell_sizes": [_vp, _i32, _vp, _vp, C.POINTER(_i64)],
    "mimpy_b200_sell_fill": [_vp, _i32, _vp, _vp, _vp, _vp],
    "mimpy_b200_pcg_sell": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(PcgInfo)],
    "mimpy_b200_spmv_sell": [_vp, _i32, _vp, _vp, _vp, _vp, _vp],
    "mimpy_b200_auxmg_create": [_vp, _PM, _PS, C.POINTER(_vp)],
    "mimpy_b200_auxmg_setup": [_vp, _vp, _PM, _PS, _vp, _f64, _vp, _vp],
    "mimpy_b200_auxmg_destroy": [_vp, _vp],
    "mimpy_b200_auxmg_cellblocks": [_vp, _vp, _PM, _PS, _vp, _vp, _vp, _i32, _vp],
    "mimpy_b200_k_anisotropy": [_vp, _i32, _vp, _f64, C.POINTER(C.c_double)],
    "mimpy_b200_pcg_aux": [_vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, C.POINTER(PcgInfo)],
    "mimpy_b200_hybrid_rhs": [_vp, _PM, _PS, _vp, _f64, _vp, _vp, _vp, _vp],
    "mimpy_b200_hybrid_recover": [_vp, _PM, _PS, _vp, _f64, _vp, _vp, _vp, _vp],
    "mimpy_b200_pcg": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(PcgInfo)],
    "mimpy_b200_spmv": [_vp, _i32, _vp, _vp, _vp, _vp, _vp],
    "mimpy_b200_twophase_mobility": [_vp, _i32, C.POINTER(Fluid), _vp, _vp, _vp, _vp, _vp],
    "mimpy_b200_twophase_upwind": [_vp, _PM, _PS, _vp, _vp],
    "mimpy_b200_twophase_newton_terms": [_vp, _i32, _i32, C.POINTER(Fluid), _f64, _vp, _vp, _vp, _vp, _vp, _vp],
    "mimpy_b200_twophase_saturation_step": [_vp, _PM, _PS, C.POINTER(Fluid), _f64, _vp, _vp, _vp,
                                            _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp],
    "mimpy_b200_twophase_saturation_step2": [_vp, _PM, _PS, C.POINTER(Fluid), _f64, _vp, _vp, _vp,
                                             _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mimpy_b200_flux_areas": [_vp, _PM, _PS, _vp],
    "mimpy_b200_sell_patterns": [_vp, _i32, _vp, _vp, _vp, _vp, C.POINTER(_i32)],
    "mimpy_b200_spmv_sell_pat": [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp],
    "mimpy_b200_set_comm": [_vp, _vp, _i32, _i32],
    "mimpy_b200_nccl_load": [C.c_char_p],
    "mimpy_b200_nccl_unique_id": [_vp],
    "mimpy_b200_nccl_init": [_vp, _vp, _i32, _i32],
    "mimpy_b200_set_halo": [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp],
    "mimpy_b200_halo_exchange_add": [_vp, _vp],
    "mimpy_b200_peer_alloc": [_vp, _i64, _vp],
    "mimpy_b200_peer_attach": [_vp, _i32, _i32, _vp],
    "mimpy_b200_peer_disable": [_vp],
    "mimpy_b200_allreduce_sum": [_vp, _vp, _i32],
}

_lib = None


class MimpyB200Error(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libmimpy_b200 error %d: %s" % (code, message))
        self.code = code


def load(required=True):
    """Loads the shared library (building nothing: see mimpy_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if required:
            raise ImportError("%s not found: run `python -m mimpy_b200.build` (there is no CPU fallback)" % LIB_PATH)
        return None
    import torch  # noqa: F401  (loads libcudart / libnccl with the right SONAMEs first)

    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, args in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            continue
        fn.argtypes = args
        fn.restype = C.c_int
    lib.mimpy_b200_last_error.argtypes = []
    lib.mimpy_b200_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc, allow=()):
    if rc != 0 and rc not in allow:
        msg = load().mimpy_b200_last_error()
        raise MimpyB200Error(rc, msg.decode("utf-8", "replace") if msg else "")
    return rc


def ptr(t):
    """Device (or pinned host) pointer of a torch tensor, None -> NULL."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("tensor passed to libmimpy_b200 must be contiguous")
    return C.c_void_p(t.data_ptr())


class Context(object):
    """One CUDA context handle + a dedicated stream on one device."""

    def __init__(self, device=None):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("mimpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device) \
            if not isinstance(device, torch.device) else device
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        h = _vp()
        check(self.lib.mimpy_b200_create(self.device.index, _vp(self.stream.cuda_stream), C.byref(h)))
        self.handle = h
        self.rank = 0
        self.world = 1

    def call(self, name, *args, **kw):
        return check(getattr(self.lib, name)(self.handle, *args), allow=kw.get("allow", ()))

    def sync(self):
        self.call("mimpy_b200_sync")

    def on_stream(self):
        """Context manager: run torch allocations/ops on this context's stream, ordered after the
        caller's current stream on entry and before it on exit."""
        return _OnStream(self)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mimpy_b200_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _OnStream(object):
    def __init__(self, ctx):
        self.ctx = ctx
        self.depth_key = "_depth"

    def __enter__(self):
        import torch

        ctx = self.ctx
        ctx._depth = getattr(ctx, "_depth", 0) + 1
        if ctx._depth == 1:
            self.outer = torch.cuda.current_stream(ctx.device)
            ctx.stream.wait_stream(self.outer)
            self.guard = torch.cuda.stream(ctx.stream)
            self.guard.__enter__()
        return ctx

    def __exit__(self, *exc):
        ctx = self.ctx
        ctx._depth -= 1
        if ctx._depth == 0:
            self.guard.__exit__(*exc)
            self.outer.wait_stream(ctx.stream)
        return False


_default_ctx = {}


def default_context(device=None):
    import torch

    idx = torch.cuda.current_device() if device is None else (device.index if hasattr(device, "index") else int(device))
    if idx not in _default_ctx:
        _default_ctx[idx] = Context(idx)
    return _default_ctx[idx]
