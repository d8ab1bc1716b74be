"""ctypes signatures of include/elynx_b200_model.h (host-side model/tree helpers)."""
import ctypes as C

_VP, _I64, _I32 = C.c_void_p, C.c_int64, C.c_int32
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int)

SIGNATURES = {
    "elx_phylo_model_parse": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, _VP, C.c_int, _VP, _VP, C.c_int, C.c_int, C.POINTER(_VP)]),
    "elx_phylo_model_expand_gamma": (C.c_int, [_VP, C.c_int, C.c_double]),
    "elx_phylo_model_free": (None, [_VP]),
    "elx_phylo_model_info": (C.c_int, [_VP, _PI, _PI, _PI, _PI]),
    "elx_phylo_model_arrays": (C.c_int, [_VP, _VP, _VP, _VP]),
    "elx_phylo_model_name": (C.c_char_p, [_VP, C.c_int]),
    "elx_gamma_means": (C.c_int, [C.c_int, C.c_double, _VP]),
    "elx_parse_edm_phylobayes": (C.c_int, [C.c_char_p, _I64, _PI, _PI, C.POINTER(_PD), C.POINTER(_PD)]),
    "elx_parse_siteprofiles_phylobayes": (C.c_int, [C.c_char_p, _I64, _PI, _PI, C.POINTER(_PD), C.POINTER(_PD)]),
    "elx_free": (None, [_VP]),
    "elx_newick_parse": (C.c_int, [C.c_char_p, _I64, C.POINTER(_VP)]),
    "elx_birth_death_tree": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_uint64, C.POINTER(_VP)]),
    "elx_flat_tree_free": (None, [_VP]),
    "elx_flat_tree_size": (C.c_int, [_VP, _PI, _PI]),
    "elx_flat_tree_arrays": (C.c_int, [_VP, _VP, _VP, _VP]),
    "elx_flat_tree_name": (C.c_char_p, [_VP, C.c_int]),
    "elx_flat_tree_to_newick": (C.c_int, [_VP, C.POINTER(_VP)]),
    "elx_gzip_write": (C.c_int, [C.c_char_p, _VP, _I64, _I32, _I32, _I64]),
}
