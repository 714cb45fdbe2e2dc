"""ctypes binding of the C-ABI CUDA library (include/tggcn_b200.h).

The header is the single source of truth: the weight table (state_dict key -> slot) and the
workspace buffer ids are parsed from it, the two structs are mirrored field by field.
There is deliberately no fallback: if the shared library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'tggcn_b200.h')
LIB_PATH = os.path.join(HERE, 'lib2ggcn_b200.so')


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'B', 'T', 'H', 'O', 'V', 'D', 'Fh', 'C_sub', 'C_aff', 'hh', 'filter', 'bn_train', 'human_seg_given',
        'object_seg_given', 'inspect', 'persistent', 'gemm_path')] + [('thr', C.c_float), ('save_for_backward', C.c_int32),
                                                                       ('cat_level_states', C.c_int32), ('mean_pool', C.c_int32),
                                                                       ('recurrent_mode', C.c_int32), ('no_fp16_split', C.c_int32),
                                                                       ('precision', C.c_int32), ('att_noscale', C.c_int32),
                                                                       ('update_strategy', C.c_int32), ('time_position', C.c_int32), ('time_periodic', C.c_int32), ('straight_through', C.c_int32), ('geo_to_human', C.c_int32), ('segment_length', C.c_int32), ('gate_layers', C.c_int32)]


class GradOutputs(C.Structure):
    _fields_ = [('d_y_hs', C.c_void_p), ('d_y_hss', C.c_void_p), ('d_y_os', C.c_void_p), ('d_y_oss', C.c_void_p),
                ('d_out_h', C.c_void_p * 4), ('d_out_o', C.c_void_p * 4)]


BWD_BUCKETS = 4          # TGGCN_BWD_BUCKETS


class BwdHooks(C.Structure):
    _fields_ = [('bucket_done', C.c_void_p * BWD_BUCKETS)]


class IO(C.Structure):
    _fields_ = [
        ('x_human', C.c_void_p), ('x_objects', C.c_void_p), ('objects_mask', C.c_void_p),
        ('human_seg', C.c_void_p), ('object_seg', C.c_void_p), ('noise', C.c_void_p),
        ('y_hs', C.c_void_p), ('y_hss', C.c_void_p), ('y_os', C.c_void_p), ('y_oss', C.c_void_p),
        ('out_h', C.c_void_p * 4), ('out_o', C.c_void_p * 4),
        ('att_frame', C.c_void_p), ('att_seg_f', C.c_void_p), ('att_seg_b', C.c_void_p),
        ('bn_running_mean', C.c_void_p), ('bn_running_var', C.c_void_p), ('bn_num_batches', C.c_void_p),
        ('dist_hh', C.c_void_p), ('dist_ho', C.c_void_p), ('dist_oo', C.c_void_p), ('steps_per_example', C.c_void_p), ('time_freq', C.c_void_p), ('status_host', C.c_void_p),
    ]


def _parse_header():
    text = open(HEADER).read()
    weights = re.findall(r'X\(\s*([A-Z0-9_]+)\s*,\s*"([^"]+)"\s*\)', text)
    body = re.search(r'enum tggcn_buf_id \{(.*?)\};', text, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    bufs = [t.split('=')[0].strip() for t in body.split(',') if t.strip()]
    assert bufs[-1] == 'TGGCN_BUF_COUNT'
    sbody = re.search(r'enum tggcn_stage_id \{(.*?)\};', text, re.S).group(1)
    sbody = re.sub(r'/\*.*?\*/', '', sbody, flags=re.S)
    stages = [t.split('=')[0].strip() for t in sbody.split(',') if t.strip()]
    assert stages[-1] == 'TGGCN_STAGE_COUNT'
    global STAGE_NAMES
    STAGE_NAMES = [n.replace('TGGCN_STAGE_', '').lower() for n in stages[:-1]]
    funcs = re.findall(r'TGGCN_API\s+[\w\s\*]+?\b(tggcn_\w+)\s*\(', text)
    return weights, bufs[:-1], funcs


STAGE_NAMES: List[str] = []
WEIGHT_TABLE, BUF_NAMES, EXPORTED = _parse_header()
WEIGHT_KEYS: List[str] = [k for _, k in WEIGHT_TABLE]
WEIGHT_INDEX: Dict[str, int] = {k: i for i, k in enumerate(WEIGHT_KEYS)}
WEIGHT_ID: Dict[str, int] = {n: i for i, (n, _) in enumerate(WEIGHT_TABLE)}
N_WEIGHTS = len(WEIGHT_TABLE)
BUF: Dict[str, int] = {n.replace('TGGCN_BUF_', ''): i for i, n in enumerate(BUF_NAMES)}

ABI_VERSION = int(re.search(r'#define TGGCN_ABI_VERSION\s+(\d+)', open(HEADER).read()).group(1))

_lib = None


class TggcnError(RuntimeError):
    pass


def lib():
    """Load (once) and return the CUDA library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TggcnError(f'{LIB_PATH} not found: build it with `python 2g-gcn_b200/build.py` '
                         '(or __graft_entry__.build()). There is no CPU / PyTorch fallback for the hot path.')
    L = C.CDLL(LIB_PATH)
    L.tggcn_abi_version.restype = C.c_int
    L.tggcn_last_error.restype = C.c_char_p
    L.tggcn_workspace_bytes.restype = C.c_size_t
    L.tggcn_workspace_bytes.argtypes = [C.POINTER(Dims)]
    L.tggcn_workspace_view.restype = C.c_int
    L.tggcn_workspace_view.argtypes = [C.POINTER(Dims), C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.tggcn_sync_status.restype = C.c_int
    L.tggcn_sync_status.argtypes = [C.POINTER(Dims), C.c_void_p, C.c_void_p]
    L.tggcn_status_decode.restype = C.c_int
    L.tggcn_status_decode.argtypes = [C.c_void_p]
    L.tggcn_forward.restype = C.c_int
    L.tggcn_forward.argtypes = [C.POINTER(Dims), C.POINTER(C.c_void_p), C.c_int, C.POINTER(IO), C.c_void_p, C.c_size_t,
                                C.c_void_p]
    L.tggcn_forward_profile.restype = C.c_int
    L.tggcn_forward_profile.argtypes = L.tggcn_forward.argtypes + [C.POINTER(C.c_float)]
    L.tggcn_launch_count.restype = C.c_ulonglong
    L.tggcn_geo_gcn_fwd.restype = C.c_int
    L.tggcn_geo_gcn_fwd.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.tggcn_linear_fwd.restype = C.c_int
    L.tggcn_linear_fwd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.tggcn_linear16_scratch_bytes.restype = C.c_size_t
    L.tggcn_linear16_scratch_bytes.argtypes = [C.c_int] * 3
    L.tggcn_linear16_fwd.restype = C.c_int
    L.tggcn_linear16_fwd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.tggcn_linear_bwd.restype = C.c_int
    L.tggcn_linear_bwd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.tggcn_bigru_fwd.restype = C.c_int
    L.tggcn_bigru_fwd.argtypes = [C.c_void_p] * 8 + [C.c_int] * 5 + [C.c_void_p]
    L.tggcn_bigru_bwd_scratch_floats.restype = C.c_size_t
    L.tggcn_bigru_bwd_scratch_floats.argtypes = [C.c_int] * 4
    L.tggcn_bigru_bwd.restype = C.c_int
    L.tggcn_bigru_bwd.argtypes = [C.c_void_p] * 12 + [C.c_int] * 5 + [C.c_void_p]
    L.tggcn_backward_workspace_bytes.restype = C.c_size_t
    L.tggcn_backward_workspace_bytes.argtypes = [C.POINTER(Dims)]
    L.tggcn_backward.restype = C.c_int
    L.tggcn_backward.argtypes = [C.POINTER(Dims), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.POINTER(IO),
                                 C.POINTER(GradOutputs), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    L.tggcn_backward_ex.restype = C.c_int
    L.tggcn_backward_ex.argtypes = L.tggcn_backward.argtypes + [C.POINTER(BwdHooks)]
    L.tggcn_backward_bucket.restype = C.c_int
    L.tggcn_backward_bucket.argtypes = [C.c_int]
    L.tggcn_upsample_argmax.restype = C.c_int
    L.tggcn_upsample_argmax.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]
    L.tggcn_f1_at_k_scratch_bytes.restype = C.c_size_t
    L.tggcn_f1_at_k_scratch_bytes.argtypes = [C.c_int] * 3
    L.tggcn_adam_step.restype = C.c_int
    L.tggcn_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_float, C.c_int, C.c_void_p]
    L.tggcn_f1_at_k.restype = C.c_int
    L.tggcn_f1_at_k.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int64,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    if L.tggcn_abi_version() != ABI_VERSION:
        raise TggcnError('lib2ggcn_b200.so ABI version mismatch; rebuild')
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        raise TggcnError(f'{what} failed (rc={rc}): {lib().tggcn_last_error().decode(errors="replace")}')


def workspace_bytes(dims: Dims) -> int:
    n = lib().tggcn_workspace_bytes(C.byref(dims))
    if n == 0:
        raise TggcnError(f'tggcn_workspace_bytes: {lib().tggcn_last_error().decode(errors="replace")}')
    return int(n)


def workspace_view(dims: Dims, name: str):
    off, nbytes = C.c_size_t(), C.c_size_t()
    check(lib().tggcn_workspace_view(C.byref(dims), BUF[name], C.byref(off), C.byref(nbytes)), 'tggcn_workspace_view')
    return int(off.value), int(nbytes.value)
