"""ctypes binding of libbpbreid_sm100.so (include/bpbreid_sm100.h).  No fallback: if the library is missing or a
call fails, a RuntimeError carrying bpb_last_error() is raised."""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbpbreid_sm100.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "bpbreid_sm100.h")

_lib = None

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_F = ctypes.c_float

_SIGS = {
    "bpb_conv_bf16": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "bpb_conv_num_mtiles": [_I, _I, _I],
    "bpb_conv_bn_bf16": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _F, _F, _P, _P, _I, _P, _P, _P, _P],
    "bpb_conv_bn_workspace_bytes": [_I],
    "bpb_conv_bn_act_bf16": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P],
    "bpb_conv_ref_bf16": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "bpb_stem_conv": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "bpb_stem_conv_tiles": [_I, _I, _I],
    "bpb_colstats_bf16": [_P, _P, _I, _I, _P],
    "bpb_colstats_tiles": [_I],
    "bpb_bn_finalize": [_P, _I, _I, _I, _L, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P],
    "bpb_bn_eval_fold": [_P, _P, _P, _P, _F, _I, _P, _P, _P],
    "bpb_bn_apply": [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "bpb_fuse_sum": [_P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    "bpb_upsample_concat": [_P, _P, _P, _P, _I, _P, _I, _I, _I, _P],
    "bpb_maxpool3x3s2": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "bpb_instnorm_stats": [_P, _P, _I, _I, _I, _I, _I, _F, _P],
    "bpb_ibn_apply": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "bpb_pixel_fold": [_P, _P, _P, _P, _I, _I, _P, _P, _P],
    "bpb_pixel_classify": [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P],
    "bpb_part_pool": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "bpb_linear_tf32": [_P, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "bpb_bn1d": [_P, _P, _P, _F, _F, _P, _P, _P, _I, _I, _I, _I, _P],
    "bpb_head_branch_stats": [_P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P],
    "bpb_head_lowres_logits": [_P, _P, _L, _I, _I, _I, _I, _P, _P],
    "bpb_head_pixel": [_P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "bpb_head_branch_pool": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "bpb_vis_unpack": [_P, _P, _I, _I, _P, _P, _P],
    "bpb_weighted_sum": [_P, _P, _I, _P, _P],
    "bpb_triplet_fwd": [_P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P],
    "bpb_triplet_fwd_rows": [_P, _P, _P, _I, _I, _I, _I, _F, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "bpb_triplet_finish": [_P, _P, _P],
    "bpb_triplet_bwd_rows": [_P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _I, _I, _I, _P, _P],
    "bpb_ce_ls_fwd": [_P, _L, _P, _P, _I, _I, _F, _I, _P, _P, _P],
    "bpb_pixel_ce_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P],
    "bpb_part_distance": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "bpb_part_distance_scratch_floats": [_I, _I, _I],
    "bpb_rank_market1501": [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "bpb_argsort_rows": [_P, _I, _I, _P, _P],
    "bpb_l2_normalize": [_P, _P, _L, _I, _F, _P],
    "bpb_resize_bilinear_f32": [_P, _P, _L, _I, _I, _I, _I, _P],
    "bpb_conv_wgrad_bf16": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "bpb_preprocess_images": [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "bpb_preprocess_masks": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _F, _F, _P, _P],
    "bpb_rerank_build": [_P, _P, _P, _I, _I, _P, _P, _P],
    "bpb_rerank_topk": [_P, _I, _I, _P, _P],
    "bpb_rerank_row_capacity": [_I],
    "bpb_rerank_kreciprocal": [_P, _P, _I, _I, _I, _P, _P, _P, _P],
    "bpb_rerank_qexpand": [_P, _I, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P],
    "bpb_rerank_csc_count": [_P, _P, _I, _I, _P, _P],
    "bpb_rerank_csc_fill": [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P],
    "bpb_rerank_jaccard_blocks": [_I, _I],
    "bpb_rerank_jaccard": [_P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _F, _F, _P, _P, _P],
    "bpb_relu_grad_sum": [_P, _I, _P, _P, _L, _I, _P],
    "bpb_bn_bwd_workspace_bytes": [_I],
    "bpb_bn_bwd_reduce": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "bpb_bn_bwd_apply": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "bpb_stem_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "bpb_adam_pack": [_P, _I, _L, _P, _P, _P, _P, _F, _F, _F, _F, _F, _I, _F, _I, _P],
    "bpb_param_segment_bytes": [],
    "bpb_memset_zero": [_P, _L, _P],
    "bpb_transpose_pad": [_P, _L, _I, _I, _P, _L, _I, _I, _P],
    "bpb_ce_ls_bwd": [_P, _L, _P, _P, _I, _I, _F, _F, _P, _L, _P],
    "bpb_triplet_bwd": [_P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _P, _P],
    "bpb_bn1d_bwd": [_P, _L, _P, _L, _P, _L, _P, _F, _I, _I, _P, _L, _I, _P, _P, _P],
    "bpb_colsum_f32": [_P, _L, _I, _I, _P, _P],
    "bpb_head_pool_bwd_prep": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "bpb_head_dmask": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "bpb_head_pixel_bwd": [_P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P, _I, _I, _F, _F, _P, _P, _P],
    "bpb_head_dwf": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "bpb_pixel_fold_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _P, _P, _P, _P, _P, _P],
    "bpb_head_branch_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
}
_PLAIN_INT = {"bpb_conv_num_mtiles", "bpb_colstats_tiles", "bpb_abi_version", "bpb_stem_conv_tiles",
              "bpb_param_segment_bytes"}


def header_symbols():
    """Every function the public header declares (used by the CPU test that checks the exports)."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bpb_[a-z0-9_]+)\s*\(", txt)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m bpbreid_b200.build` "
                           "(there is no CPU / PyTorch fallback for the hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.bpb_last_error.restype = ctypes.c_char_p
    lib.bpb_last_error.argtypes = []
    lib.bpb_abi_version.restype = _I
    lib.bpb_abi_version.argtypes = []
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = sig
        fn.restype = _L if name in ("bpb_part_distance_scratch_floats", "bpb_conv_bn_workspace_bytes",
                                    "bpb_bn_bwd_workspace_bytes") else _I
    _lib = lib
    return lib


def call(name, *args):
    """Invoke a status-returning entry point; raise with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (rc={rc}): {lib.bpb_last_error().decode()}")


def query(name, *args):
    return getattr(load(), name)(*args)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
