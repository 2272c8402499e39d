"""ctypes binding of libvlg.so (include/vlg.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VLG_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libvlg.so")   # override: A/B builds

VLG_F32, VLG_BF16 = 0, 1
VLG_C2I, VLG_T2I, VLG_T2V = 0, 1, 2
MODEL_TYPES = {"c2i": VLG_C2I, "t2i": VLG_T2I, "t2v": VLG_T2V}


class SampleParams(C.Structure):
    _fields_ = [("cfg_scale", C.c_float), ("guided", C.c_int), ("temperature", C.c_float), ("top_k", C.c_int),
                ("top_p", C.c_float), ("sample_logits", C.c_int), ("seed", C.c_uint64), ("offset", C.c_uint64)]


class Geometry(C.Structure):
    _fields_ = [("dim", C.c_int), ("n_layer", C.c_int), ("n_head", C.c_int), ("hidden_dim", C.c_int),
                ("vocab_size", C.c_int), ("cls_token_num", C.c_int), ("num_classes", C.c_int),
                ("caption_dim", C.c_int), ("vae_embed_dim", C.c_int), ("model_type", C.c_int), ("dtype", C.c_int),
                ("norm_eps", C.c_float)]


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("attention_norm", "wqkv", "wo", "ffn_norm", "w13", "w2")]


class Weights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("tok_embeddings", "cls_table", "cap_fc1", "cap_fc2", "adapter_in_fc1",
                                          "adapter_in_fc2", "adapter_out_fc1", "adapter_out_fc2", "norm", "output",
                                          "rope_table")] + [("rope_rows", C.c_int), ("layers", C.POINTER(LayerWeights))]


class GenerateParams(C.Structure):
    _fields_ = [("B", C.c_int), ("max_new_tokens", C.c_int), ("cfg_scale", C.c_float), ("cfg_interval", C.c_int),
                ("temperature", C.c_float), ("top_k", C.c_int), ("top_p", C.c_float), ("sample_logits", C.c_int),
                ("seed", C.c_uint64)]


class DiffGeometry(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("C", "W", "Z", "depth", "steps")]


class DiffBlock(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ln_w", "ln_b", "w1", "b1", "w2", "b2", "ada_w", "ada_b")]


class DiffWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("time_w0", "time_b0", "time_w1", "time_b1", "cond_w", "cond_b", "in_w", "in_b")] + \
               [("blocks", C.POINTER(DiffBlock))] + \
               [(n, C.c_void_p) for n in ("final_ada_w", "final_ada_b", "final_w", "final_b")]


_P, _I, _I64, _SZ, _F = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float

# name -> (restype, argtypes): exactly the symbols include/vlg.h declares
SIGNATURES = {
    "vlg_abi_version": (_I, []),
    "vlg_last_error": (C.c_char_p, []),
    "vlg_launch_count": (C.c_uint64, []),
    "vlg_vq_scratch_floats": (_SZ, [_I64, _I, _I]),
    "vlg_vq_argmin": (_I, [_P, _P, _P, _P, _P, _I64, _I, _I64, _I, _I, _P]),
    "vlg_codebook_gather": (_I, [_P, _P, _P, _I64, _I, _I64, _I, _I, _I, _P]),
    "vlg_codebook_project": (_I, [_P, _P, _P, _P, _P, _I64, _I, _I64, _I, _I, _I, _P]),
    "vlg_cfg_sample": (_I, [_P, _I, _I, _I, C.POINTER(SampleParams), _P, _P, _P, _P]),
    "vlg_top_k_top_p": (_I, [_P, _P, _I, _I, _I, _F, _P]),
    "vlg_ctx_workspace_bytes": (_SZ, [C.POINTER(Geometry), _I, _I]),
    "vlg_ctx_create": (_I, [C.POINTER(Geometry), C.POINTER(Weights), _I, _I, _P, _SZ, C.POINTER(_P)]),
    "vlg_ctx_destroy": (None, [_P]),
    "vlg_ctx_check": (_I, [_P, _P]),
    "vlg_begin": (_I, [_P, _I, _P, _P]),
    "vlg_forward_pos": (_I, [_P, _P, _P, _P, _I, _P, _P, _P]),
    "vlg_forward_prefix": (_I, [_P, _P, _P, _P, _P]),
    "vlg_mlp": (_I, [_P, _I, _P, _P, _I, _P]),
    "vlg_generate": (_I, [_P, C.POINTER(GenerateParams), _P, _P, _P, _P, _P]),
    "vlg_generate_t2v": (_I, [_P, C.POINTER(GenerateParams), _P, _P, _P, _P, _P]),
    "vlg_generate_c2i_host": (_I, [_P, C.POINTER(GenerateParams), _P, _P, _P]),
    "vlg_diff_workspace_bytes": (_SZ, [C.POINTER(DiffGeometry), _I]),
    "vlg_diff_create": (_I, [C.POINTER(DiffGeometry), C.POINTER(DiffWeights), _P, _P, _I, _P, _SZ, C.POINTER(_P)]),
    "vlg_diff_destroy": (None, [_P]),
    "vlg_diff_sample": (_I, [_P, _P, _I, _F, _F, _P, _P, C.c_uint64, C.c_uint64, _P, _P]),
    "vlg_diff_check": (_I, [_P, _P]),
    "vlg_rmsnorm": (_I, [_P, _P, _P, _I, _I, _F, _I, _P]),
    "vlg_linear": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "vlg_attn_decode": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P, _SZ, _P]),
    "vlg_attn_scratch_bytes": (_SZ, [_I, _I, _I, _I]),
    "vlg_debug_fp32_fma_peak": (_I, [_I, _I, _P, C.POINTER(C.c_longlong), _P]),
    "vlg_debug_gemm_timestamps": (_I, [_P]),
    "vlg_debug_lt_timestamps": (_I, [_P]),
    "vlg_debug_pair_gemm": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "vlg_debug_trunk_without_attention": (_I, [_P, _I, _P]),
}

_lib = None


class VlgError(RuntimeError):
    pass


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VlgError(f"{LIB_PATH} is missing: build it with `python video-llamagen_b200/build.py` "
                           "(there is no CPU or eager fallback for this path)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        if h.vlg_abi_version() != 1:
            raise VlgError("libvlg.so ABI version mismatch")
        _lib = h
    return _lib


def check(rc):
    if rc != 0:
        raise VlgError(f"libvlg error {rc}: {lib().vlg_last_error().decode()}")


def ptr(t):
    """Device/host pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what):
    if not t.is_cuda:
        raise VlgError(f"{what}: tensor is on {t.device}; the vlg path runs on a B200 only (no CPU fallback)")
    if not t.is_contiguous():
        raise VlgError(f"{what}: tensor must be contiguous")


def dtype_code(dt):
    import torch
    if dt == torch.bfloat16:
        return VLG_BF16
    if dt == torch.float32:
        return VLG_F32
    raise VlgError(f"unsupported dtype {dt}: the decode path runs in bfloat16 or float32")
