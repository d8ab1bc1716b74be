"""ctypes binding of libelynx_b200.so -- the C ABI declared in include/elynx_b200.h.

The library is the product; there is no Python/CPU fallback.  If the shared object is missing this
module raises ImportError (build it with `python -m elynx_b200.build`)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ELX_LIB") or os.path.join(_HERE, "libelynx_b200.so")  # ELX_LIB: A/B builds of the same ABI

ELX_OK = 0
ELX_E_INVALID = -1
ELX_E_CUDA = -2
ELX_E_UNSUPPORTED = -3
ELX_E_NOMEM = -4
ELX_E_BAD_WEIGHTS = -5
ELX_FLAG_FORCE_GENERIC = 1
ELX_FLAG_PADE = 2
ELX_FLAG_SINGLE_DRAWS = 4
ELX_FLAG_PAIR_DRAWS = 8
ELX_FLAG_EIGEN = 16
ELX_QUAD_REC_INTS = 52


class ElxError(RuntimeError):
    """Raised where the reference calls `error` (MarkovProcess.hs:35-37, RateMatrix.hs:125, ...)."""

    def __init__(self, code: int, msg: str):
        super().__init__("%s (elx status %d)" % (msg, code))
        self.code = code
        self.msg = msg


class elx_model(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("n_components", C.c_int32), ("is_mixture", C.c_int32),
                ("weights", C.POINTER(C.c_double)), ("pi", C.POINTER(C.c_double)), ("exch", C.POINTER(C.c_double))]


class elx_tree(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("parent", C.POINTER(C.c_int32)), ("brlen", C.POINTER(C.c_double)),
                ("leaf_row", C.POINTER(C.c_int32))]


class elx_opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("philox_rounds", C.c_int32), ("flags", C.c_int32), ("n_devices", C.c_int32)]


class elx_examine_result(C.Structure):
    _fields_ = [("n_sites", C.c_int64), ("n_sequences", C.c_int64), ("n_states", C.c_int32), ("reserved", C.c_int32),
                ("n_constant", C.c_int64), ("n_constant_soft", C.c_int64), ("n_no_gaps", C.c_int64), ("n_only_std", C.c_int64),
                ("n_std_chars", C.c_int64), ("n_gaps", C.c_int64), ("distribution", C.c_double * 64),
                ("keff_entropy_mean", C.c_double), ("keff_entropy_mean_no_gaps", C.c_double),
                ("keff_homoplasy_mean", C.c_double), ("keff_homoplasy_mean_no_gaps", C.c_double),
                ("hamming_mean", C.c_double), ("hamming_var", C.c_double), ("hamming_min", C.c_double), ("hamming_max", C.c_double)]


_VP, _I64, _U64, _I32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32
_CSTRS = C.POINTER(C.c_char_p)

# name -> (restype, argtypes); every symbol include/elynx_b200.h declares
SIGNATURES = {
    "elx_version": (C.c_int, []),
    "elx_last_error": (C.c_char_p, [_VP]),
    "elx_create": (C.c_int, [C.POINTER(elx_model), C.POINTER(elx_tree), C.POINTER(elx_opts), C.POINTER(_VP)]),
    "elx_destroy": (None, [_VP]),
    "elx_tree_walk_info": (C.c_int, [C.POINTER(elx_tree), C.POINTER(_I32), C.POINTER(_I32), _VP]),
    "elx_tree_pair_walk_info": (C.c_int, [C.POINTER(elx_tree), C.POINTER(_I32), C.POINTER(_I32), _VP]),
    "elx_n_devices": (_I32, [_VP]),
    "elx_shard_range": (C.c_int, [_I64, _I32, _I32, _I32, C.POINTER(_I64), C.POINTER(_I64)]),
    "elx_last_kernel": (C.c_char_p, [_VP]),
    "elx_n_leaves": (_I32, [_VP]),
    "elx_n_nodes": (_I32, [_VP]),
    "elx_n_states": (_I32, [_VP]),
    "elx_n_components": (_I32, [_VP]),
    "elx_stream": (_VP, [_VP]),
    "elx_ptables_get": (C.c_int, [_VP, _VP, C.c_int]),
    "elx_ptables_set": (C.c_int, [_VP, _VP]),
    "elx_alias_columns": (_I32, [_VP]),
    "elx_alias_get": (C.c_int, [_VP, _VP, _VP]),
    "elx_pair_groups": (_I32, [_VP]),
    "elx_pair_get": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "elx_quad_groups": (_I32, [_VP]),
    "elx_quad_get": (C.c_int, [_VP, _VP, _VP, _VP]),
    "elx_tree_quad_walk_info": (C.c_int, [C.POINTER(elx_tree), C.POINTER(_I32), C.POINTER(_I32), _VP]),
    "elx_duo_groups": (_I32, [_VP]),
    "elx_duo_get": (C.c_int, [_VP, _VP, _VP, _VP]),
    "elx_tree_duo_walk_info": (C.c_int, [C.POINTER(elx_tree), C.POINTER(_I32), C.POINTER(_I32), _VP]),
    "elx_simulate": (C.c_int, [_VP, _U64, _I64, _I64, _VP, _I64, _VP, _VP, _I64]),
    "elx_simulate_device": (C.c_int, [_VP, _U64, _I64, _I64, _VP, _I64, _VP, _VP, _I64]),
    "elx_simulate_fasta_gz": (C.c_int, [_VP, _U64, _I64, _CSTRS, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, _VP]),
    "elx_trim_device_cache": (None, []),
    "elx_result_alloc": (_VP, [_I64]),
    "elx_result_free": (None, [_VP]),
    "elx_result_pool_trim": (None, []),
    "elx_packed_bits": (_I32, [_VP]),
    "elx_packed_pitch": (_I64, [_VP, _I64]),
    "elx_simulate_packed": (C.c_int, [_VP, _U64, _I64, _I64, _VP, _I64, _VP]),
    "elx_site_components": (C.c_int, [_VP, _U64, _I64, _I64, _VP]),
    "elx_synchronize": (C.c_int, [_VP]),
    "elx_simulate_injected": (C.c_int, [_VP, _VP, _I64, _I64, _VP, _I64, _VP, _VP, _I64]),
    "elx_simulate_injected_device": (C.c_int, [_VP, _VP, _I64, _I64, _VP, _I64, _VP, _VP, _I64]),
    "elx_fasta_size": (_I64, [_VP, _CSTRS, _I64]),
    "elx_fasta_pack_device": (C.c_int, [_VP, _VP, _I64, _I64, _I64, _I64, _CSTRS, C.c_char_p, _VP]),
    "elx_simulate_fasta": (C.c_int, [_VP, _U64, _I64, _CSTRS, C.c_char_p, _VP, _I64, _VP]),
    "elx_fasta_pack_slice_device": (C.c_int, [_VP, _VP, _I64, _I64, C.c_char_p, _VP, _I64]),
    "elx_fasta_row_offset": (_I64, [_VP, _CSTRS, _I64, _I32]),
    "elx_histogram_device": (C.c_int, [_VP, _VP, _I64, _I64, _VP, _VP]),
    "elx_examine_device": (C.c_int, [_VP, _I64, _I32, _I64, _I32, C.POINTER(elx_examine_result), _VP, _VP, _VP, _VP]),
    "elx_examine": (C.c_int, [_VP, _I64, _I32, _I64, _I32, C.POINTER(elx_examine_result), _VP, _VP, _VP]),
    "elx_launch_count": (_I64, [_VP]),
    "elx_d2h_bytes": (_I64, [_VP]),
    "elx_unpack2_rows": (_I32, [_VP, _I64, _VP, _I64, _I64, _I64, C.c_char_p, C.c_int32]),
    "elx_unpack5_rows": (_I32, [_VP, _I64, _VP, _I64, _I64, _I64, C.c_char_p, C.c_int32]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libelynx_b200.so is missing (no CPU fallback exists): run `python -m elynx_b200.build`")
        _lib = C.CDLL(LIB_PATH)
        from . import _model_sigs

        for table in (SIGNATURES, _model_sigs.SIGNATURES):
            for name, (res, args) in table.items():
                fn = getattr(_lib, name)
                fn.restype = res
                fn.argtypes = args
    return _lib


def last_error(handle=None) -> str:
    s = lib().elx_last_error(handle)
    return s.decode(errors="replace") if s else ""


def check(rc: int, handle=None):
    if rc != 0:
        raise ElxError(rc, last_error(handle))
