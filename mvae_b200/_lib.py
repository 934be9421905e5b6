"""ctypes binding of libmvae_b200.so (C ABI declared in include/mvae_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  Loading it works on a CPU-only box (so the
symbol table can be checked there); every compute entry point needs a compute-capability-10.x device and returns
a negative mvae_status otherwise, which `check()` turns into an exception.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmvae_b200.so")

MAX_COMPONENTS = 96
ABI_VERSION = 3

# mvae_manifold
EUCLIDEAN, HYPERBOLOID, SPHERE, POINCARE, PROJ_SPHERE, UNIVERSAL = 0, 1, 2, 3, 4, 5
TYPE_OF_LETTER = {"e": EUCLIDEAN, "h": HYPERBOLOID, "s": SPHERE, "p": POINCARE, "d": PROJ_SPHERE, "u": UNIVERSAL}
# mvae_op
(OP_EXP_MAP_MU0, OP_INV_EXP_MAP_MU0, OP_EXP_MAP, OP_INV_EXP_MAP, OP_PT_MU0, OP_INV_PT_MU0, OP_DISTANCE, OP_MOBIUS_ADD,
 OP_MOBIUS_SCALAR_MUL, OP_LOGDET, OP_TO_POINCARE, OP_FROM_POINCARE) = range(12)
# mvae_epilogue
EPI_STORE, EPI_BIAS_RELU, EPI_RELU_MASK, EPI_BCE_ROWSUM, EPI_NLL_ROWSUM = range(5)
K_MAJOR, MN_MAJOR = 0, 1


class Component(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("type", "n", "d", "m_off", "l_off", "l_n", "eps_off", "z_off")]


class PmDesc(ctypes.Structure):
    _fields_ = [("C", ctypes.c_int32), ("ld_ml", ctypes.c_int32), ("ld_eps", ctypes.c_int32),
                ("ld_z", ctypes.c_int32), ("comp", Component * MAX_COMPONENTS)]


class Planes(ctypes.Structure):
    _fields_ = [("base", ctypes.c_void_p), ("plane_stride", ctypes.c_int64), ("rows", ctypes.c_int32),
                ("cols", ctypes.c_int32), ("ld", ctypes.c_int32), ("planes", ctypes.c_int32)]


class LatentPrologue(ctypes.Structure):
    _fields_ = [("draw_eps", ctypes.c_int32), ("n_zero", ctypes.c_int32), ("seed", ctypes.c_uint64),
                ("counter_dev", ctypes.c_void_p), ("zero_ptr", ctypes.c_void_p * 4), ("zero_n", ctypes.c_int64 * 4)]


class GemmArgs(ctypes.Structure):
    _fields_ = [("a", Planes), ("b", Planes), ("a_major", ctypes.c_int32), ("b_major", ctypes.c_int32),
                ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32), ("epilogue", ctypes.c_int32),
                ("split_k", ctypes.c_int32), ("bias", ctypes.c_void_p), ("out_f32", ctypes.c_void_p),
                ("ld_out", ctypes.c_int64), ("out_col", ctypes.c_void_p), ("col_split", ctypes.c_int32),
                ("out_planes", Planes), ("aux", ctypes.c_void_p), ("ld_aux", ctypes.c_int64),
                ("mask", ctypes.c_void_p), ("ld_mask", ctypes.c_int64), ("rowsum", ctypes.c_void_p),
                ("tile_n", ctypes.c_int32), ("ctas_per_sm", ctypes.c_int32), ("aux_rows", ctypes.c_int64)]


DP_MAX_RANKS, DP_HANDLE_BYTES, DP_CHANNELS, DP_SYNC_WORDS = 8, 64, 2, 528
DP_FLAG_BYTES = (DP_CHANNELS + 1) * 2 * DP_MAX_RANKS * 128


class DpComm(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("bucket", ctypes.c_void_p * DP_MAX_RANKS),
                ("flat", ctypes.c_void_p * DP_MAX_RANKS), ("flags", ctypes.c_void_p * DP_MAX_RANKS)]


class DpStepArgs(ctypes.Structure):
    _fields_ = [("n_net", ctypes.c_int64), ("begin", ctypes.c_int64), ("end", ctypes.c_int64),
                ("channel", ctypes.c_int32), ("do_tail", ctypes.c_int32), ("n_tail", ctypes.c_int32),
                ("C", ctypes.c_int32), ("exp_avg", ctypes.c_void_p), ("exp_avg_sq", ctypes.c_void_p),
                ("lr", ctypes.c_float), ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("eps", ctypes.c_float),
                ("step_dev", ctypes.c_void_p), ("radius", ctypes.c_void_p), ("radius_mask", ctypes.c_void_p),
                ("clip_mask", ctypes.c_void_p), ("radius_lr", ctypes.c_float), ("clip_max_norm", ctypes.c_float),
                ("tail_out", ctypes.c_void_p), ("sync_words", ctypes.c_void_p), ("max_ctas", ctypes.c_int32),
                ("n_targets", ctypes.c_int32), ("target_begin", ctypes.POINTER(ctypes.c_int64)),
                ("target_rows", ctypes.POINTER(ctypes.c_int32)), ("targets", ctypes.POINTER(Planes))]


_vp, _i32, _i64, _f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float

# name -> (restype, argtypes); every symbol include/mvae_b200.h declares
PROTOTYPES = {
    "mvae_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "mvae_abi_version": (ctypes.c_int, []),
    "mvae_last_cuda_error": (ctypes.c_int, []),
    "mvae_pm_desc_init": (ctypes.c_int, [ctypes.POINTER(PmDesc), _i32, ctypes.POINTER(_i32), ctypes.POINTER(_i32), _i32]),
    "mvae_pm_forward": (ctypes.c_int, [ctypes.POINTER(PmDesc), _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvae_pm_backward": (ctypes.c_int, [ctypes.POINTER(PmDesc), _i64, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp]),
    "mvae_manifold_op": (ctypes.c_int, [_i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mvae_wn_rsample": (ctypes.c_int, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvae_wn_log_prob_from_parts": (ctypes.c_int, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvae_wn_log_prob": (ctypes.c_int, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvae_split_planes": (ctypes.c_int, [_vp, _i64, _i32, _i32, ctypes.POINTER(Planes), ctypes.POINTER(Planes), _vp]),
    "mvae_gemm": (ctypes.c_int, [ctypes.POINTER(GemmArgs), _vp]),
    "mvae_skinny_rowdot": (ctypes.c_int, [_i64, _i32, _i32, _vp, _i64, ctypes.POINTER(Planes), _vp, _i64, _i64, _vp, _vp, _i64, _vp]),
    "mvae_skinny_expand": (ctypes.c_int, [_i64, _i32, _i32, _vp, _i64, _vp, _i64, _i64, _vp, _i32, _vp, _i64,
                                          ctypes.POINTER(Planes), _vp, _i64, _vp]),
    "mvae_skinny_wgrad": (ctypes.c_int, [_i64, _i32, _i32, _vp, _i64, _i32, _vp, _i64, ctypes.POINTER(Planes), _vp, _i64,
                                         _i64, _vp, _vp, _i32, _vp]),
    "mvae_latent_forward": (ctypes.c_int, [ctypes.POINTER(PmDesc), _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(Planes), _vp, _vp]),
    "mvae_latent_forward_ex": (ctypes.c_int, [ctypes.POINTER(PmDesc), _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp,
                                              _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(Planes), _vp,
                                              ctypes.POINTER(LatentPrologue), _vp]),
    "mvae_latent_backward": (ctypes.c_int, [ctypes.POINTER(PmDesc), _i64, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp,
                                            _vp, _vp, _vp, _f32, ctypes.POINTER(Planes), _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvae_recon_loss": (ctypes.c_int, [_i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mvae_elbo_reduce": (ctypes.c_int, [_i64, _i32, _vp, _vp, _f32, _vp, _vp]),
    "mvae_binarize": (ctypes.c_int, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, ctypes.c_uint64, _vp, _vp, _i64,
                                     ctypes.POINTER(Planes), _vp]),
    "mvae_iwae_latent": (ctypes.c_int, [ctypes.POINTER(PmDesc), _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvae_iwae_reduce": (ctypes.c_int, [_i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mvae_iwae_cov_norm": (ctypes.c_int, [_i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mvae_adam_step": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _i32, _f32, _vp]),
    "mvae_adam_step_dev": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _vp, _f32, _vp]),
    "mvae_clip_grad_norm": (ctypes.c_int, [_i32, _vp, _vp, _f32, _vp]),
    "mvae_sgd_step": (ctypes.c_int, [_i64, _vp, _vp, _f32, _f32, _vp]),
    "mvae_opt_step_fused": (ctypes.c_int, [_i64, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _vp,
                                           _f32, _i32, _i32, ctypes.POINTER(_i64), ctypes.POINTER(_i32),
                                           ctypes.POINTER(Planes), _vp]),
    "mvae_conv_im2col": (ctypes.c_int, [ctypes.POINTER(Planes), _i32, _i32, _i32, _i32, ctypes.POINTER(Planes), _i32, _vp]),
    "mvae_conv_col2im": (ctypes.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _i32, ctypes.POINTER(Planes),
                                        ctypes.POINTER(Planes), _vp, _i64, _vp]),
    "mvae_permute_sc": (ctypes.c_int, [_i32, _vp, _i64, _i64, _vp, _i64, _i64, _i32, _i64, _i32, _i32, _i32, _vp]),
    "mvae_colsum": (ctypes.c_int, [_vp, ctypes.POINTER(Planes), _i64, _i32, _i64, _vp, _vp]),
    "mvae_step_prologue": (ctypes.c_int, [_vp, _i64, ctypes.c_uint64, _vp, _i32, ctypes.POINTER(ctypes.c_void_p),
                                          ctypes.POINTER(_i64), _vp]),
    "mvae_ring_push": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "mvae_counter_add": (ctypes.c_int, [_vp, ctypes.c_uint64, _vp]),
    "mvae_dp_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "mvae_dp_free": (ctypes.c_int, [_vp]),
    "mvae_dp_ipc_export": (ctypes.c_int, [_vp, ctypes.c_char_p]),
    "mvae_dp_ipc_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "mvae_dp_ipc_close": (ctypes.c_int, [_vp]),
    "mvae_dp_step": (ctypes.c_int, [ctypes.POINTER(DpComm), ctypes.POINTER(DpStepArgs), _vp]),
    "mvae_dp_rendezvous": (ctypes.c_int, [ctypes.POINTER(DpComm), _vp, _vp]),
    "mvae_rt_memcpy_async": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "mvae_rt_event_record": (ctypes.c_int, [_vp, _vp]),
    "mvae_rt_stream_wait_event": (ctypes.c_int, [_vp, _vp]),
    "mvae_debug_latent": (ctypes.c_int, [_vp, _i32]),
    "mvae_debug_gemm": (ctypes.c_int, [_vp, _i32]),
    "mvae_debug_timeline": (ctypes.c_int, [_vp]),
    "mvae_debug_timeline_log": (ctypes.c_int, [ctypes.POINTER(_i32), _i32]),
    "mvae_device_info": (ctypes.c_int, [ctypes.POINTER(_i32), ctypes.POINTER(_i32), ctypes.POINTER(_i32)]),
}

_lib = None


class MvaeError(RuntimeError):
    pass


def lib():
    """Load libmvae_b200.so (fails loudly if it has not been built: `python -m mvae_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MvaeError(f"{LIB_PATH} is missing — build it with `python -m mvae_b200.build` "
                            "(there is no CPU / PyTorch fallback for this path)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.mvae_abi_version() != ABI_VERSION:
            raise MvaeError("libmvae_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        L = lib()
        msg = L.mvae_strerror(rc).decode()
        cuda = L.mvae_last_cuda_error()
        raise MvaeError(f"{what or 'mvae call'} failed: {msg} (status {rc}, last cudaError {cuda})")


def make_desc(types, dims, scalar_parametrization=False) -> PmDesc:
    C = len(types)
    d = PmDesc()
    rc = lib().mvae_pm_desc_init(ctypes.byref(d), C, (ctypes.c_int32 * C)(*types), (ctypes.c_int32 * C)(*dims),
                                 int(bool(scalar_parametrization)))
    check(rc, "mvae_pm_desc_init")
    return d
