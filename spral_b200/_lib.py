"""ctypes binding of libspral_ssids_b200.so (the C ABI in include/spral_ssids_b200.h).

The product path is the CUDA library; there is no CPU fallback.  Importing this
module fails loudly when the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C spral_b200/csrc`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspral_ssids_b200.so")


class Options(C.Structure):
    """struct spral_ssids_b200_options == cpu_factor_options (cpu_iface.hxx:24-34)."""
    _fields_ = [("print_level", C.c_int), ("action", C.c_bool), ("small", C.c_double),
                ("u", C.c_double), ("multiplier", C.c_double),
                ("small_subtree_threshold", C.c_int64), ("cpu_block_size", C.c_int),
                ("pivot_method", C.c_int), ("failed_pivot_method", C.c_int)]

    @classmethod
    def default(cls):
        # defaults of ssids_options (src/ssids/datatypes.f90:188-284)
        return cls(print_level=0, action=True, small=1e-20, u=0.01, multiplier=1.1,
                   small_subtree_threshold=4 * 10 ** 6, cpu_block_size=256,
                   pivot_method=2, failed_pivot_method=1)


class Stats(C.Structure):
    """struct spral_ssids_b200_stats == ThreadStats (ThreadStats.hxx:48-62) + cuda_error."""
    _fields_ = [("flag", C.c_int), ("num_delay", C.c_int), ("num_factor", C.c_int64),
                ("num_flops", C.c_int64), ("num_neg", C.c_int), ("num_two", C.c_int),
                ("num_zero", C.c_int), ("maxfront", C.c_int), ("maxsupernode", C.c_int),
                ("not_first_pass", C.c_int), ("not_second_pass", C.c_int),
                ("cuda_error", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Contrib(C.Structure):
    """struct spral_ssids_b200_contrib == Fortran contrib_type (contrib.f90:19-33)."""
    _fields_ = [("ready", C.c_int), ("n", C.c_int), ("val", C.c_void_p), ("ldval", C.c_int),
                ("rlist", C.c_void_p), ("ndelay", C.c_int), ("delay_perm", C.c_void_p),
                ("delay_val", C.c_void_p), ("lddelay", C.c_int), ("owner", C.c_int),
                ("posdef", C.c_bool), ("owner_ptr", C.c_void_p), ("device", C.c_int)]


class AnalysisView(C.Structure):
    _fields_ = [("n", C.c_int), ("nnodes", C.c_int), ("nparts", C.c_int),
                ("sptr", C.POINTER(C.c_int)), ("sparent", C.POINTER(C.c_int)),
                ("rptr", C.POINTER(C.c_int64)), ("rlist", C.POINTER(C.c_int)),
                ("nptr", C.POINTER(C.c_int64)), ("nlist", C.POINTER(C.c_int64)),
                ("invp", C.POINTER(C.c_int)), ("part", C.POINTER(C.c_int)),
                ("exec_loc", C.POINTER(C.c_int)), ("contrib_ptr", C.POINTER(C.c_int)),
                ("contrib_idx", C.POINTER(C.c_int)), ("contrib_dest", C.POINTER(C.c_int)),
                ("num_factor", C.c_int64), ("num_flops", C.c_int64),
                ("maxfront", C.c_int), ("maxsupernode", C.c_int), ("maxdepth", C.c_int)]


# every symbol include/spral_ssids_b200.h declares: name -> (restype, argtypes)
_vp, _i, _b, _d = C.c_void_p, C.c_int, C.c_bool, C.c_double
_ip, _lp, _dp = C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_double)
SYMBOLS = {
    "spral_ssids_b200_cuda_init": (_i, [_ip]),
    "spral_ssids_gpu_create_symbolic_subtree":
        (_vp, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, C.POINTER(Options)]),
    "spral_ssids_gpu_destroy_symbolic_subtree": (None, [_vp]),
    "spral_ssids_gpu_create_num_subtree_dbl":
        (_vp, [_b, _vp, _vp, _vp, _vp, C.POINTER(Options), C.POINTER(Stats)]),
    "spral_ssids_gpu_destroy_num_subtree_dbl": (None, [_b, _vp]),
    "spral_ssids_gpu_subtree_solve_fwd_dbl": (_i, [_b, _vp, _i, _vp, _i]),
    "spral_ssids_gpu_subtree_solve_diag_dbl": (_i, [_b, _vp, _i, _vp, _i]),
    "spral_ssids_gpu_subtree_solve_diag_bwd_dbl": (_i, [_b, _vp, _i, _vp, _i]),
    "spral_ssids_gpu_subtree_solve_bwd_dbl": (_i, [_b, _vp, _i, _vp, _i]),
    "spral_ssids_gpu_subtree_enquire_dbl": (None, [_b, _vp, _vp, _vp]),
    "spral_ssids_gpu_subtree_alter_dbl": (None, [_b, _vp, _vp]),
    "spral_ssids_gpu_subtree_get_contrib_dbl":
        (None, [_b, _vp, _ip, C.POINTER(_vp), _ip, C.POINTER(_vp), _ip, C.POINTER(_vp),
                C.POINTER(_vp), _ip]),
    "spral_ssids_gpu_subtree_get_contrib_device_dbl":
        (None, [_b, _vp, _ip, C.POINTER(_vp), _ip, C.POINTER(_vp), _ip, C.POINTER(_vp),
                C.POINTER(_vp), _ip, _ip]),
    "spral_ssids_gpu_subtree_free_contrib_dbl": (None, [_b, _vp]),
    "spral_ssids_gpu_subtree_export_contrib_ipc":
        (_i, [_vp, _vp, _ip, _ip, _lp, C.POINTER(_vp)]),
    "spral_ssids_b200_ipc_pull": (_i, [_i, _vp, C.c_int64, _vp]),
    "spral_ssids_b200_copy_to_host": (_i, [_vp, _vp, C.c_int64]),
    "spral_ssids_b200_device_alloc": (_vp, [_i, C.c_int64]),
    "spral_ssids_b200_device_free": (None, [_vp]),
    "spral_ssids_b200_contrib_fill": (None, [C.POINTER(Contrib), _b, _vp, _b]),
    "spral_ssids_gpu_symbolic_get_maps": (None, [_vp, _vp, _ip, _vp, _vp]),
    "spral_ssids_gpu_subtree_get_timings": (None, [_vp, _dp, _i]),
    "spral_ssids_b200_set_profile": (None, [_i]),
    "spral_ssids_b200_fp64_peak_tflops": (_d, [_i]),
    "spral_ssids_b200_metis_order": (_i, [_i, _vp, _vp, _vp]),
    "spral_ssids_b200_analyse":
        (_vp, [_i, _vp, _vp, _vp, _i, _i, C.c_int64, C.c_float, C.c_float, _ip]),
    "spral_ssids_b200_analysis_free": (None, [_vp]),
    "spral_ssids_b200_analysis_get": (None, [_vp, C.POINTER(AnalysisView)]),
    "spral_ssids_b200_analysis_set_partition": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "spral_ssids_b200_hungarian_scale_sym": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _ip]),
    "spral_ssids_b200_equilib_scale_sym": (_i, [_i, _vp, _vp, _vp, _vp, _i, _d, _ip]),
    "spral_ssids_b200_match_order_metis": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
    "spral_ssids_b200_auction_scale_sym": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _ip, _ip]),
}

_lib = None


def load():
    """Load the shared library and bind every declared symbol (loud on failure)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA engine is the only implementation "
            "(no CPU fallback). Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
