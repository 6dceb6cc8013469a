"""ctypes binding of libsbte_b200.so (include/sbte_b200.h). No CPU fallback: a missing library or a
missing CUDA device raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsbte_b200.so")

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/sbte_b200.h one to one
PROTOTYPES = {
    # drop-in link interface
    "initialize_coll": (None, [C.c_int, C.c_double, _dp, _dp]),
    "dealloc_coll": (None, []),
    "ComputeQ": (None, [_dp, _dp, _dp, C.POINTER(_dp)]),
    "ComputeQ_maxPreserve": (None, [_dp, _dp, _dp, C.POINTER(_dp)]),
    "fft3D": (None, [_vp, _vp, C.c_int]),
    "initialize_conservation": (None, [C.c_int, C.c_double, _dp, _vp, C.c_int]),
    "initialize_conservation_fast": (None, [C.c_int, C.c_double, _dp]),
    "conserveAllMoments": (None, [C.POINTER(_dp)]),
    "dealloc_conservation": (None, []),
    "initialize_transport": (None, [C.c_int, C.c_int, C.c_double, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double, _vp]),
    "advectOne": (None, [C.POINTER(_dp), C.POINTER(_dp), C.c_int]),
    "advectTwo": (None, [C.POINTER(_dp), C.POINTER(_dp), C.c_int]),
    "dealloc_trans": (None, []),
    # extension surface
    "sbte_last_error": (C.c_char_p, []),
    "sbte_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_double, _dp, _dp, C.c_int]),
    "sbte_destroy": (C.c_int, [_vp]),
    "sbte_device_count": (C.c_int, []),
    "sbte_enable_peer_access": (C.c_int, [_vp, _vp]),
    "sbte_sync": (C.c_int, [_vp]),
    "sbte_stream": (_vp, [_vp]),
    "sbte_launch_count": (C.c_ulonglong, [_vp]),
    "sbte_reserve": (C.c_int, [_vp, C.c_int]),
    "sbte_set_symmetrize": (C.c_int, [_vp, C.c_int]),
    "sbte_set_xy_pairing": (C.c_int, [_vp, C.c_int]),
    "sbte_xy_pairing_state": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "sbte_batch_schedule_host": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                           C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_ubyte), C.POINTER(C.c_int)]),
    "sbte_k2_profile": (C.c_int, [_vp, C.c_int]),
    "sbte_k2_profile_read": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "sbte_dev_alloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "sbte_dev_free": (C.c_int, [_vp]),
    "sbte_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "sbte_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "sbte_d2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "sbte_weights_upload_rows": (C.c_int, [_vp, C.POINTER(_dp)]),
    "sbte_weights_upload": (C.c_int, [_vp, _dp]),
    "sbte_weights_load_file": (C.c_int, [_vp, C.c_char_p]),
    "sbte_weights_bind_device": (C.c_int, [_vp, _vp]),
    "sbte_weights_fill_synthetic": (C.c_int, [_vp, C.c_ulonglong]),
    "sbte_weights_device": (_vp, [_vp]),
    "sbte_weights_generate_iso": (C.c_int, [_vp, C.c_double]),
    "sbte_weights_save_file": (C.c_int, [_vp, C.c_char_p]),
    "sbte_fft3d": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int]),
    "sbte_qhat": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int]),
    "sbte_compute_q": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int]),
    "sbte_compute_q_maxpreserve": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int]),
    "sbte_conserve": (C.c_int, [_vp, _vp, C.c_int]),
    "sbte_moment_functionals": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "sbte_moments": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "sbte_step_0d": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_int, C.c_int]),
    "sbte_compute_q_host": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int]),
    "sbte_compute_q_maxpreserve_host": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int]),
    "sbte_slab_create": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_int, C.c_int]),
    "sbte_slab_destroy": (C.c_int, [_vp]),
    "sbte_slab_set_twall_in": (C.c_int, [_vp, C.c_double]),
    "sbte_slab_f": (_vp, [_vp]),
    "sbte_slab_fconv": (_vp, [_vp]),
    "sbte_slab_upload": (C.c_int, [_vp, _dp]),
    "sbte_slab_download": (C.c_int, [_vp, _dp]),
    "sbte_slab_advect": (C.c_int, [_vp, C.c_int]),
    "sbte_slab_upwind_stage": (C.c_int, [_vp, C.c_int, C.c_int]),
    "sbte_slab_advect_finish": (C.c_int, [_vp, C.c_int]),
    "sbte_slab_halo_regions": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "sbte_slab_ipc_export": (C.c_int, [_vp, C.c_char_p]),
    "sbte_slab_ipc_import": (C.c_int, [_vp, C.c_int, C.c_char_p, C.c_int]),
    "sbte_slab_peer_attach": (C.c_int, [_vp, C.c_int, _vp]),
    "sbte_slab_halo_state": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "sbte_slab_set_halo_timeout": (C.c_int, [_vp, C.c_double]),
    "sbte_slab_set_peer_halo": (C.c_int, [_vp, C.c_int]),
    "sbte_slab_peer_detach": (C.c_int, [_vp]),
    "sbte_slab_collide": (C.c_int, [_vp, C.c_double, C.c_int]),
    "sbte_slab_step": (C.c_int, [_vp, C.c_double, C.c_int]),
    "sbte_slab_moments": (C.c_int, [_vp, _dp]),
}

_lib = None


def load():
    """Load the shared library and attach prototypes. Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libsbte_b200.so is missing (%s): build it with `python -m spectralbte_b200.build`; "
                "there is no CPU fallback" % LIB_PATH)
        # SBTE_LIB_PATH: another build of the library for A/B timing (tools/); symbols it lacks are skipped there only
        override = os.environ.get("SBTE_LIB_PATH")
        L = C.CDLL(override or LIB_PATH)  # RTLD_LOCAL: the drop-in symbol names stay private to this handle
        for name, (res, args) in PROTOTYPES.items():
            if override and not hasattr(L, name):
                continue
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class SbteError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise SbteError(load().sbte_last_error().decode())
