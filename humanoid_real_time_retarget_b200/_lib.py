"""ctypes binding of libhrt_b200.so (include/hrt_b200.h).  There is no CPU fallback: if the CUDA
library has not been built, or no GPU is visible, every product entry point raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhrt_b200.so")

_lib = None

c_float_p = C.POINTER(C.c_float)
_P = C.c_void_p

_SIGNATURES = {
    "hrt_abi_version": (C.c_int, []),
    "hrt_last_error_string": (C.c_char_p, []),
    "hrt_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "hrt_ctx_destroy": (C.c_int, [_P]),
    "hrt_ctx_sm_count": (C.c_int, [_P]),
    "hrt_set_tree": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "hrt_fk_local_quats": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P, _P, C.c_uint, _P]),
    "hrt_fk_angles": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P, C.c_int, _P, _P, C.c_uint, _P]),
    "hrt_fk_jacobian": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P]),
    "hrt_fk_vjp": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P, _P]),
    "hrt_ik_refine": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.c_int, C.c_float, C.c_float, C.c_uint, _P, _P, _P]),
    "hrt_local_from_global": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P]),
    "hrt_zero_pose_transform": (C.c_int, [_P, C.c_int, C.c_int64, _P, C.c_int, _P, _P]),
    "hrt_configure_body_quat": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "hrt_retarget_body_quat": (C.c_int, [_P, C.c_int64, _P, C.c_uint, C.c_int, C.c_float, C.c_float, _P, _P, _P, _P]),
    "hrt_retarget_body_quat_host": (C.c_int, [_P, C.c_int64, _P, C.c_uint, C.c_int, C.c_float, C.c_float, _P, _P, _P]),
    "hrt_peer_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P), C.c_char_p]),
    "hrt_peer_free": (C.c_int, [_P, _P]),
    "hrt_peer_open": (C.c_int, [_P, C.c_char_p, C.POINTER(_P)]),
    "hrt_peer_close": (C.c_int, [_P, _P]),
    "hrt_retarget_body_quat_gather": (C.c_int, [_P, C.c_int64, _P, C.c_uint, C.c_int, C.c_float, C.c_float, _P, C.c_int,
                                                C.POINTER(_P), C.c_int64, _P]),
    "hrt_peer_barrier": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), C.c_uint, _P]),
    "hrt_reassembly_layout": (C.c_int, [_P, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "hrt_retarget_body_quat_reassemble": (C.c_int, [_P, C.c_int64, _P, C.c_uint, C.c_int, C.c_float, C.c_float, _P, _P, C.c_int64, C.c_int,
                                                    C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _P, _P, C.c_uint, _P]),
    "hrt_retarget_body_quat_multicast": (C.c_int, [_P, C.c_int64, _P, C.c_uint, C.c_int, C.c_float, C.c_float, _P, _P, C.c_int64, _P]),
    "hrt_configure_pos": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int]),
    "hrt_retarget_full_body_pos": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "hrt_retarget_full_body_pos_ex": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_uint, C.c_int, C.c_float, C.c_float, _P, _P, _P, _P]),
    "hrt_retarget_full_body_pos_host": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_uint, C.c_int, C.c_float, C.c_float, _P, _P]),
    "hrt_retarget_upper_body": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P]),
    "hrt_retarget_full_body": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "hrt_stream_open": (C.c_int, [_P, C.c_uint, C.c_int, C.c_float, C.c_float]),
    "hrt_stream_frame": (C.c_int, [_P, _P, _P, _P, _P]),
    "hrt_stream_close": (C.c_int, [_P]),
    "hrt_retarget_full_body_pos_wire": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "hrt_stream_pos_open": (C.c_int, [_P, C.c_int]),
    "hrt_stream_pos_frame": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "hrt_stream_pos_frame_ex": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "hrt_stream_pos_close": (C.c_int, [_P]),
    "hrt_retarget_main_arms": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P]),
    "hrt_rescale_motion": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P, _P]),
    "hrt_rebuild_global_rotation": (C.c_int, [_P, C.c_int, C.c_int64, _P, C.c_int, _P, _P, _P, _P]),
    "hrt_motion_velocity": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_float, C.c_int, _P, _P, _P]),
    "hrt_motion_angular_velocity": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_float, C.c_int, _P, _P, _P]),
    "hrt_forward_vector": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _P, _P, _P]),
    "hrt_rot_op_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hrt_rot_op": (C.c_int, [_P, C.c_int, C.c_int64, C.POINTER(_P), C.POINTER(C.c_int64), C.c_int, C.c_float, C.POINTER(_P), _P]),
    "hrt_max_norm3": (C.c_int, [_P, C.c_int64, _P, C.POINTER(C.c_float), _P]),
    "hrt_cal_joint_quat": (C.c_int, [_P, C.c_int64, C.c_int, _P, C.c_int64, _P, _P, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class HrtError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises HrtError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HrtError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.hrt_abi_version() != 1:
        raise HrtError("libhrt_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().hrt_last_error_string().decode("utf-8", "replace")
        raise HrtError(f"hrt_b200 error {rc}: {msg}")
