"""ctypes binding of libguidegen_sm100.so (the C ABI declared in include/guidegen_sm100.h).

There is no fallback of any kind: if the shared library is missing, or a call returns a
non-zero status, a RuntimeError is raised.  Tensors are owned by PyTorch; only raw device
pointers, sizes and the current CUDA stream cross the boundary.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GG_LIB") or os.path.join(_HERE, "lib", "libguidegen_sm100.so")     # GG_LIB: a tuning build (build.py)

_lib = None


class GuideGenLibraryError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GuideGenLibraryError(
                f"{LIB_PATH} not found: build it with `python -m jointimagegeneration_b200.build` "
                "(the sm_100a CUDA library is the only implementation; there is no CPU/PyTorch fallback)")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


# ------------------------------------------------------------------------------------ structs
class CatArgs(C.Structure):
    _fields_ = [("x0", C.c_void_p), ("xt", C.c_void_p), ("q", C.c_void_p), ("coef", C.c_void_p),
                ("out", C.c_void_p), ("out_i64", C.c_void_p), ("labels", C.c_void_p),
                ("B", C.c_int32), ("C", C.c_int32), ("V", C.c_int64), ("clamp_min", C.c_float),
                ("mode", C.c_int32), ("seed", C.c_uint64), ("offset", C.c_uint64)]


class CatStepCLArgs(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("labels_in", C.c_void_p), ("q", C.c_void_p), ("coef", C.c_void_p),
                ("cond", C.c_void_p), ("labels_out", C.c_void_p), ("next_x", C.c_void_p), ("probs_out", C.c_void_p),
                ("B", C.c_int32), ("C", C.c_int32), ("Cpad", C.c_int32), ("n_cond", C.c_int32), ("Cin_pad", C.c_int32),
                ("V", C.c_int64), ("clamp_min", C.c_float), ("mode", C.c_int32), ("seed", C.c_uint64),
                ("offset", C.c_uint64), ("vox_base", C.c_int64)]


class DdimArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("e_t", C.c_void_p), ("noise", C.c_void_p), ("coef", C.c_void_p),
                ("x_prev", C.c_void_p), ("pred_x0", C.c_void_p), ("n", C.c_int64), ("temperature", C.c_float),
                ("e_uncond", C.c_void_p), ("guidance_scale", C.c_float)]


class DdpmArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("e_t", C.c_void_p), ("noise", C.c_void_p), ("coef", C.c_void_p),
                ("x_prev", C.c_void_p), ("x0_out", C.c_void_p), ("B", C.c_int32), ("per_sample", C.c_int64),
                ("temperature", C.c_float), ("clip_denoised", C.c_int32)]


class GnFinalizeArgs(C.Structure):
    _fields_ = [("partial1", C.c_void_p), ("C1", C.c_int32), ("nchunks1", C.c_int32),
                ("partial2", C.c_void_p), ("C2", C.c_int32), ("nchunks2", C.c_int32),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("scale_shift", C.c_void_p),
                ("N", C.c_int32), ("groups", C.c_int32), ("S", C.c_int64), ("eps", C.c_float),
                ("slab_world", C.c_int32), ("slab_rank", C.c_int32), ("slab_phase", C.c_int32),
                ("slab_tables", C.c_void_p * 8), ("slab_flag_out", C.c_void_p * 8), ("slab_flag_in", C.c_void_p * 8),
                ("slab_epoch", C.c_void_p), ("slab_done_counter", C.c_void_p)]


class PlmsArgs(C.Structure):
    _fields_ = [("e_t", C.c_void_p), ("e_uncond", C.c_void_p), ("old1", C.c_void_p), ("old2", C.c_void_p), ("old3", C.c_void_p),
                ("e_cur", C.c_void_p), ("e_prime", C.c_void_p), ("n", C.c_int64), ("order", C.c_int32), ("guidance_scale", C.c_float)]


class CatEpilogue(C.Structure):
    _fields_ = [("labels_in", C.c_void_p), ("labels_out", C.c_void_p), ("next_x", C.c_void_p), ("cond", C.c_void_p),
                ("coef", C.c_void_p), ("C", C.c_int32), ("n_cond", C.c_int32), ("Cin_pad", C.c_int32), ("mode", C.c_int32),
                ("clamp_min", C.c_float), ("seed", C.c_uint64), ("offset", C.c_uint64), ("vox_base", C.c_int64)]


class ConvSrc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("C", C.c_int32), ("centre_only", C.c_int32), ("d_shift", C.c_int32), ("reserved", C.c_int32)]


class ConvArgs(C.Structure):
    _fields_ = [("src", ConvSrc * 4), ("nsrc", C.c_int32),
                ("N", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("dims", C.c_int32),
                ("kd", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("od", C.c_int32), ("oh", C.c_int32), ("ow", C.c_int32), ("stride", C.c_int32),
                ("Do", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
                ("w_packed", C.c_void_p), ("bias", C.c_void_p), ("emb", C.c_void_p), ("emb_stride", C.c_int32),
                ("residual", C.c_void_p), ("res_stride", C.c_int32), ("y", C.c_void_p),
                ("y_sn", C.c_int64), ("y_sd", C.c_int64), ("y_sh", C.c_int64), ("y_sw", C.c_int64),
                ("y_is_f32", C.c_int32), ("Cout", C.c_int32), ("block_n", C.c_int32), ("brick", C.c_int32 * 4),
                ("gn_partial", C.c_void_p), ("gn_chunk_base", C.c_int32), ("gn_nchunks_total", C.c_int32),
                ("stats_d_min", C.c_int32), ("algo", C.c_int32), ("split_k", C.c_int32), ("workspace", C.c_void_p),
                ("src_ss", C.c_void_p * 4), ("ss_stride", C.c_int32), ("xf_silu", C.c_int32),
                ("xf_z_lo", C.c_int32), ("xf_z_hi", C.c_int32), ("cat", C.POINTER(CatEpilogue)), ("split_counters", C.c_void_p)]


class AttnArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
                ("q_bs", C.c_int64), ("k_bs", C.c_int64), ("v_bs", C.c_int64), ("o_bs", C.c_int64),
                ("q_rs", C.c_int32), ("k_rs", C.c_int32), ("v_rs", C.c_int32), ("o_rs", C.c_int32),
                ("q_hs", C.c_int32), ("k_hs", C.c_int32), ("v_hs", C.c_int32), ("o_hs", C.c_int32),
                ("B", C.c_int32), ("H", C.c_int32), ("Tq", C.c_int32), ("Tk", C.c_int32), ("d", C.c_int32),
                ("scale", C.c_float), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64)]


class PeerXchgArgs(C.Structure):
    _fields_ = [("nsend", C.c_int32), ("src", C.c_void_p * 8), ("dst", C.c_void_p * 8), ("bytes", C.c_int64 * 8),
                ("nflag_out", C.c_int32), ("flag_out", C.c_void_p * 8),
                ("nflag_in", C.c_int32), ("flag_in", C.c_void_p * 8),
                ("ncopy", C.c_int32), ("csrc", C.c_void_p * 4), ("cdst", C.c_void_p * 4), ("cbytes", C.c_int64 * 4),
                ("nzero", C.c_int32), ("zdst", C.c_void_p * 2), ("zbytes", C.c_int64 * 2),
                ("epoch", C.c_void_p), ("done_counter", C.c_void_p), ("phase", C.c_int32), ("ctas", C.c_int32)]


# the ctypes mirrors above, in the order gg_abi_sizes() reports the C structs
ABI_STRUCTS = [CatArgs, CatStepCLArgs, DdimArgs, PlmsArgs, DdpmArgs, GnFinalizeArgs, ConvSrc, ConvArgs, AttnArgs, CatEpilogue,
               PeerXchgArgs]

# every symbol include/guidegen_sm100.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "gg_version": (C.c_int, []),
    "gg_status_string": (C.c_char_p, [C.c_int]),
    "gg_device_check": (C.c_int, []),
    "gg_abi_sizes": (C.c_int, [C.POINTER(C.c_int32), C.c_int]),
    "gg_launch_count": (C.c_uint64, []),
    "gg_launch_count_reset": (None, []),
    "gg_cat_posterior_sample": (C.c_int, [C.POINTER(CatArgs), _vp]),
    "gg_cat_step_cl": (C.c_int, [C.POINTER(CatStepCLArgs), _vp]),
    "gg_ddim_update": (C.c_int, [C.POINTER(DdimArgs), _vp]),
    "gg_ddpm_update": (C.c_int, [C.POINTER(DdpmArgs), _vp]),
    "gg_plms_eps": (C.c_int, [C.POINTER(PlmsArgs), _vp]),
    "gg_labels_to_mask": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "gg_labels_gather": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _f32, _vp]),
    "gg_minmax_normalize": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _i64, _vp]),
    "gg_nchw_to_cl": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _i32, _i64, _vp]),
    "gg_cl_to_nchw": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i64, _i32, _vp]),
    "gg_gn_num_chunks": (_i32, [_i64, _i32]),
    "gg_gn_partial": (C.c_int, [_vp, _i32, _i64, _i32, _vp, _vp]),
    "gg_gn_finalize": (C.c_int, [C.POINTER(GnFinalizeArgs), _vp]),
    "gg_gn_apply": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _i32, _i64, _i32, _vp]),
    "gg_conv_pick_block_n": (_i32, [_i32]),
    "gg_conv_packed_k": (_i64, [C.POINTER(ConvArgs)]),
    "gg_conv_stats_chunks": (_i32, [C.POINTER(ConvArgs)]),
    "gg_conv_num_tiles": (_i32, [C.POINTER(ConvArgs)]),
    "gg_conv_fwd": (C.c_int, [C.POINTER(ConvArgs), _vp]),
    "gg_upsample2x": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "gg_attention_workspace_bytes": (_i64, [C.POINTER(AttnArgs)]),
    "gg_attention_fwd": (C.c_int, [C.POINTER(AttnArgs), _vp]),
    "gg_timestep_embedding": (C.c_int, [_vp, _vp, _i32, _i32, _f32, _vp]),
    "gg_small_linear": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "gg_gn_fused": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _i64, _i32, _f32, _i32, _vp]),
    "gg_gn_fused_resident": (_i32, [_i64, _i32]),
    "gg_layernorm": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    "gg_geglu": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "gg_softmax_rows": (C.c_int, [_vp, _vp, _i64, _i32, _f32, _vp]),
    "gg_transpose_bf16": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "gg_peer_alloc": (C.c_int, [_i64, C.POINTER(C.c_void_p)]),
    "gg_peer_free": (C.c_int, [_vp]),
    "gg_peer_export": (C.c_int, [_vp, C.c_char_p]),
    "gg_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "gg_peer_close": (C.c_int, [_vp]),
    "gg_peer_epoch_inc": (C.c_int, [_vp, _vp]),
    "gg_peer_exchange": (C.c_int, [C.POINTER(PeerXchgArgs), _vp]),
}


def _declare(l):
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(l, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib().gg_status_string(status).decode()
        raise GuideGenLibraryError(f"libguidegen_sm100 {what} failed: status {status} ({msg})")


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise GuideGenLibraryError("guidegen kernels need CUDA tensors (no CPU fallback exists)")


def launch_count() -> int:
    return int(lib().gg_launch_count())
