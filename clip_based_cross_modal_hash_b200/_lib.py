"""Build and load ``libcmh.so`` (the C-ABI of include/cmh.h) and declare its ctypes prototypes.

The library is built IN-TREE with a plain ``nvcc -shared`` for sm_100a (no torch headers: the ABI is plain
pointers and sizes).  There is no CPU fallback: if the library cannot be loaded the product raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
SO_PATH = os.path.join(_HERE, "libcmh.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libcmh.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return SO_PATH
    nvcc = os.environ.get("NVCC") or "nvcc"
    if not any(os.access(os.path.join(p, nvcc), os.X_OK) for p in os.environ.get("PATH", "").split(os.pathsep)):
        if os.path.exists("/usr/local/cuda/bin/nvcc"):
            nvcc = "/usr/local/cuda/bin/nvcc"
    tmp = "%s.tmp.%d" % (SO_PATH, os.getpid())   # per process: ranks of one torchrun launch may all find the library stale
    cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC] + sources() + ["-o", tmp]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-8000:]))
    os.replace(tmp, SO_PATH)
    if verbose:
        print(res.stderr)
    return SO_PATH


class CmhError(RuntimeError):
    pass


class Plan(ctypes.Structure):
    """Mirror of ``struct cmh_plan`` (include/cmh.h)."""

    _fields_ = [
        ("Q", ctypes.c_int64), ("N", ctypes.c_int64), ("N_geom", ctypes.c_int64), ("Qpad", ctypes.c_int64),
        ("nbits", ctypes.c_int32), ("ncls", ctypes.c_int32), ("W", ctypes.c_int32), ("LW", ctypes.c_int32),
        ("bins", ctypes.c_int32), ("nchunks", ctypes.c_int32),
        ("chunk_items", ctypes.c_int64), ("hist_elems", ctypes.c_int64), ("within_elems", ctypes.c_int64),
        ("below_elems", ctypes.c_int64), ("ap_elems", ctypes.c_int64), ("workspace_bytes", ctypes.c_int64),
    ]


class TcOperands(ctypes.Structure):
    """Mirror of ``struct cmh_tc_operands`` (include/cmh.h): int8 operand rows of the tensor-core ranking passes."""

    _fields_ = [("q_codes", ctypes.c_void_p), ("q_labels", ctypes.c_void_p), ("g_codes", ctypes.c_void_p),
                ("g_labels", ctypes.c_void_p), ("code_bytes", ctypes.c_int32), ("label_bytes", ctypes.c_int32)]


_i32, _i64, _vp, _sz = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t
_PP = ctypes.POINTER(Plan)
_OP = ctypes.POINTER(TcOperands)

# name -> argtypes (restype is int unless noted); kept in one table so tests can check every symbol of cmh.h
PROTOTYPES = {
    "cmh_abi_version": [],
    "cmh_last_error": [],
    "cmh_launch_count": [],
    "cmh_device_info": [ctypes.POINTER(_i32)] * 3,
    "cmh_code_words": [_i32],
    "cmh_label_words": [_i32],
    "cmh_pack_codes_f32": [_vp, _i64, _i32, _i64, _vp, _vp, _vp],
    "cmh_pack_labels": [_vp, _i32, _i64, _i32, _i64, _vp, _vp, _vp],
    "cmh_unpack_codes_f32": [_vp, _i64, _i32, _vp, _i64, _vp],
    "cmh_hamming_f32": [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp],
    "cmh_hamming_dense_f32": [_vp, _i64, _vp, _i64, _i32, _vp, _vp],
    "cmh_make_plan": [_i64, _i64, _i64, _i32, _i32, _i32, _PP],
    "cmh_hist": [_PP, _vp, _vp, _vp, _vp, _vp, _vp],
    "cmh_scan": [_PP, _vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "cmh_hist_totals": [_PP, _vp, _vp, _vp],
    "cmh_scan_sharded": [_PP, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "cmh_rank_map": [_PP, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp],
    "cmh_ap_reduce": [_PP, _vp, _i32, _vp, _vp],
    "cmh_map_finish": [_PP, _vp, _i32, _vp, _vp, _vp, _vp],
    "cmh_rank_topk": [_PP, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp],
    "cmh_fill_keys": [_vp, _i64, ctypes.c_uint64, _vp],
    "cmh_topk_merge": [_vp, _i32, _i64, _i64, _vp, _vp],
    "cmh_split_keys": [_vp, _i64, _vp, _vp, _vp],
    "cmh_map_k": [_PP, _vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "cmh_topk": [_PP, _vp, _vp, _i64, _i64, _vp, _sz, _vp, _vp],
    "cmh_tc_operand_bytes": [_i32],
    "cmh_tc_expand": [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp],
    "cmh_tc_hist": [_PP, _OP, _i32, _vp, _vp],
    "cmh_tc_rank_topk": [_PP, _OP, _vp, _vp, _vp, _i64, _i64, _vp, _vp],
    "cmh_tc_rank_map": [_PP, _OP, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp],
    "cmh_tc_topk_cutoff": [_PP, _vp, _i64, _i64, _vp, _vp, _vp],
    "cmh_tc_topk_sample_block": [_PP, _vp, _i64, _i32, _i64, _i64, _i32, _i32, _vp, _vp],
    "cmh_tc_topk_cutoff_sharded": [_PP, _vp, _i64, _i32, _i32, _vp, _vp, _vp],
    "cmh_tc_topk_collect": [_PP, _OP, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp],
    "cmh_tc_topk_count": [_PP, _i32, _vp, _vp, _i64, _vp, _vp, _vp],
    "cmh_tc_topk_place": [_PP, _i32, _vp, _vp, _vp, _i64, _i32, _i32, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _vp],
    "cmh_nvls_allreduce_max_s64": [_vp, _i64, _i32, _i32, _vp],
    "cmh_nvls_push_owned_s64": [_vp, _vp, _i64, _vp],
    "cmh_nvls_broadcast": [_vp, _vp, _i64, _vp],
    "cmh_label_sim_f32": [_vp, _i64, _vp, _i64, _i32, _vp, _vp],
    "cmh_cosine_sim_f32": [_vp, _i64, _vp, _i64, _i32, _vp, _vp],
    "cmh_euclid_sim_f32": [_vp, _i64, _vp, _i64, _i32, _vp, _vp],
    "cmh_gemm_bf16": [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i32, _vp, _i64, _vp, _i64, _vp],
    "cmh_gemm_force_tile": [_i32, _i32],
    "cmh_gemm_set_trace": [_vp],
    "cmh_gemm_force_units": [_i32],
    "cmh_gemm_tail_slicing": [_i32],
    "cmh_gemm_mma_lookahead": [_i32],
    "cmh_encoder_workspace_bytes": [_vp, _i64, _i32],
    "cmh_encode_image": [_vp, _vp, _i64, _vp, _sz, _vp, _vp, _vp, _vp],
    "cmh_encode_image_u8": [_vp, _vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), _i64, _vp, _sz, _vp, _vp, _vp, _vp],
    "cmh_encode_text": [_vp, _vp, _vp, _i64, _i32, _vp, _sz, _vp, _vp, _vp, _vp, _vp],
    "cmh_layernorm": [_vp, _i64, _i32, _vp, _vp, ctypes.c_float, _vp, _i32, _vp],
    "cmh_attention_bf16": [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp],
    "cmh_linear_f32": [_vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _i64, _vp],
    "cmh_head_dsph": [_vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _vp],
    "cmh_head_dcmht": [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp],
    "cmh_hyp_loss_f32": [_vp, _vp, _vp, _vp, _i64, _i32, _i32, ctypes.c_float, ctypes.c_float, _vp, _sz, _vp, _vp],
    "cmh_hyp_loss_grad_f32": [_vp, _vp, _vp, _vp, _i64, _i32, _i32, ctypes.c_float, ctypes.c_float, _vp, _sz, _vp, _vp, _vp, _vp, _vp],
    "cmh_linear_tanh_backward_f32": [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "cmh_opt_chunk_elems": [],
    "cmh_bert_adam_step": [_vp, _i32, _vp, _vp, _i32, _vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, _vp],
    "cmh_sgd_momentum_step": [_vp, _i32, _vp, _vp, _i32, ctypes.c_float, _i32, _vp],
    "cmh_head_mith_workspace_bytes": [_vp, _i64, _i32],
    "cmh_head_mith": [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _i64, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp],
}
_RESTYPES = {"cmh_last_error": ctypes.c_char_p, "cmh_launch_count": ctypes.c_ulonglong, "cmh_encoder_workspace_bytes": ctypes.c_int64,
             "cmh_head_mith_workspace_bytes": ctypes.c_int64}

ABI_VERSION = 2   # CMH_ABI_VERSION of include/cmh.h
_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    """The loaded library; builds it first if the in-tree .so is missing or older than its sources."""
    global _lib
    if _lib is None:
        path = SO_PATH
        if _stale():
            try:
                path = build()
            except Exception as e:  # no nvcc on the box: use the prebuilt file if there is one
                if not os.path.exists(SO_PATH):
                    raise CmhError("libcmh.so is missing and cannot be built: %s" % e)
                import warnings

                warnings.warn("libcmh.so is older than its sources and the rebuild failed (%s): loading the stale build" % e)
        try:
            handle = ctypes.CDLL(path)
        except OSError as e:
            raise CmhError("cannot load %s: %s" % (path, e))
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        if handle.cmh_abi_version() != ABI_VERSION:
            raise CmhError("libcmh.so ABI version %d, this package expects %d (stale build?)" % (handle.cmh_abi_version(), ABI_VERSION))
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().cmh_last_error()
        raise CmhError("libcmh error %d: %s" % (rc, msg.decode() if msg else "?"))


def make_plan(Q: int, N: int, nbits: int, ncls: int, N_geom: Optional[int] = None, target_blocks: int = 0) -> Plan:
    p = Plan()
    check(lib().cmh_make_plan(Q, N, N if N_geom is None else N_geom, nbits, ncls, target_blocks, ctypes.byref(p)))
    return p
