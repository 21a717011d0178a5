"""ctypes binding of libhj_b200.so (the C ABI declared in include/hj.h).

The library is built in-tree by ``make -C hephaestus-jit_b200`` (or ``__graft_entry__.build()``).
There is no fallback of any kind: if the shared library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhj_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "hj.h")

HJ_UNIQUE_ID_BYTES = 128


class HjError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"hj status {status}: {message}")
        self.status = status
        self.message = message


# status codes (include/hj.h)
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NVRTC, ERR_NCCL, ERR_NO_DEVICE, ERR_OOM = range(8)


class TypeDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("elem", ctypes.c_uint32), ("num", ctypes.c_uint32),
                ("cols", ctypes.c_uint32), ("rows", ctypes.c_uint32),
                ("first_field", ctypes.c_uint32)]


class IrVar(ctypes.Structure):
    _fields_ = [("ty", ctypes.c_uint32), ("op", ctypes.c_uint32), ("arg", ctypes.c_uint32),
                ("dep_start", ctypes.c_uint32), ("dep_end", ctypes.c_uint32),
                ("_pad", ctypes.c_uint32), ("data", ctypes.c_uint64)]


class Ir(ctypes.Structure):
    _fields_ = [("vars", ctypes.POINTER(IrVar)), ("n_vars", ctypes.c_uint32),
                ("deps", ctypes.POINTER(ctypes.c_uint32)), ("n_deps", ctypes.c_uint32),
                ("types", ctypes.POINTER(TypeDesc)), ("n_types", ctypes.c_uint32),
                ("struct_fields", ctypes.POINTER(ctypes.c_uint32)),
                ("n_struct_fields", ctypes.c_uint32), ("n_buffers", ctypes.c_uint32)]


class Pass(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("arg", ctypes.c_uint32),
                ("resources", ctypes.POINTER(ctypes.c_uint32)), ("n_resources", ctypes.c_uint32),
                ("size_buffer", ctypes.c_int32), ("ir", ctypes.POINTER(Ir)),
                ("size", ctypes.c_uint64)]


class BufferDesc(ctypes.Structure):
    _fields_ = [("size", ctypes.c_uint64), ("ty", ctypes.c_uint32), ("elem_bytes", ctypes.c_uint32)]


class PassReport(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 64), ("start_us", ctypes.c_double),
                ("duration_us", ctypes.c_double)]


class GraphReport(ctypes.Structure):
    _fields_ = [("aliasing_rate", ctypes.c_float), ("aliasing_duration_us", ctypes.c_double),
                ("backend_cpu_us", ctypes.c_double), ("n_passes", ctypes.c_uint32),
                ("passes", ctypes.POINTER(PassReport)), ("passes_capacity", ctypes.c_uint32)]


class ShardDesc(ctypes.Structure):
    _fields_ = [("placement", ctypes.c_uint32), ("deferred", ctypes.c_uint32), ("seed", ctypes.c_void_p)]


HJ_IPC_HANDLE_BYTES = 64
RES_REPLICATED, RES_SHARDED, RES_AUTO = 0, 1, 2


class Report(ctypes.Structure):
    _fields_ = [("cpu_duration_us", ctypes.c_double), ("n_passes", ctypes.c_uint32),
                ("passes", ctypes.POINTER(PassReport)), ("passes_capacity", ctypes.c_uint32)]


def declared_symbols() -> list[str]:
    """Every function name include/hj.h declares (used by the symbol-export test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hj_[a-z0-9_]+)\s*\(", text)))


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C {_HERE}` "
            "(or python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU / PyTorch fallback for this backend.")
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_vp, _sz, _i32, _u32, _u64 = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32, ctypes.c_uint32,
                              ctypes.c_uint64)
_pvp = ctypes.POINTER(ctypes.c_void_p)
_pu64, _pu32, _pi32 = ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_int32)

_SIGS = {
    "hj_last_error": (ctypes.c_char_p, []),
    "hj_abi_version": (_u32, []),
    "hj_device_count": (_i32, []),
    "hj_device_create": (_i32, [_i32, _pvp]),
    "hj_device_retain": (_i32, [_vp]),
    "hj_device_release": (_i32, [_vp]),
    "hj_device_sync": (_i32, [_vp]),
    "hj_device_stream": (_i32, [_vp, _pvp]),
    "hj_device_set_stream": (_i32, [_vp, _vp]),
    "hj_device_info": (_i32, [_vp, ctypes.POINTER(_i32), ctypes.POINTER(_i32), ctypes.POINTER(_i32),
                              ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "hj_device_pool_stats": (_i32, [_vp] + [ctypes.POINTER(_u64)] * 4),
    "hj_device_pool_trim": (_i32, [_vp]),
    "hj_device_launch_count": (_i32, [_vp, ctypes.POINTER(_u64)]),
    "hj_buffer_create": (_i32, [_vp, _sz, _pvp]),
    "hj_buffer_create_from_slice": (_i32, [_vp, _vp, _sz, _pvp]),
    "hj_buffer_create_from_host_async": (_i32, [_vp, _vp, _sz, _sz, _pvp]),
    "hj_async_chunk_schedule": (_i32, [_u64, _u64, _pu64, _pu64, _u32, ctypes.POINTER(_u32)]),
    "hj_buffer_wrap": (_i32, [_vp, _vp, _sz, _pvp]),
    "hj_buffer_retain": (_i32, [_vp]),
    "hj_buffer_release": (_i32, [_vp]),
    "hj_buffer_to_host": (_i32, [_vp, _sz, _sz, _vp]),
    "hj_buffer_upload": (_i32, [_vp, _sz, _vp, _sz]),
    "hj_buffer_fill_zero": (_i32, [_vp]),
    "hj_buffer_device_ptr": (_i32, [_vp, _pvp]),
    "hj_buffer_size": (_i32, [_vp, ctypes.POINTER(_sz)]),
    "hj_buffer_device": (_i32, [_vp, _pvp]),
    "hj_host_alloc": (_i32, [_sz, _pvp]),
    "hj_host_alloc_near": (_i32, [_vp, _sz, _pvp]),
    "hj_host_free": (_i32, [_vp]),
    "hj_reduce": (_i32, [_vp, _i32, _i32, _sz, _vp, _vp]),
    "hj_prefix_sum": (_i32, [_vp, _i32, _sz, _i32, _vp, _vp, _vp]),
    "hj_compress": (_i32, [_vp, _sz, _vp, _vp, _vp, _vp, _u32]),
    "hj_scatter_reduce": (_i32, [_vp, _i32, _i32, _sz, _vp, _vp, _u64, _vp, _sz]),
    "hj_gather": (_i32, [_vp, _sz, _sz, _vp, _vp, _vp]),
    "hj_reduce_host": (_i32, [_vp, _i32, _i32, _sz, _vp, _vp, _sz]),
    "hj_prefix_sum_host": (_i32, [_vp, _i32, _sz, _i32, _vp, _vp, _sz]),
    "hj_compress_host": (_i32, [_vp, _sz, _vp, _vp, _pu32, _u32, _sz]),
    "hj_ir_hash": (_u64, [ctypes.POINTER(Ir)]),
    "hj_ir_codegen": (_i32, [ctypes.POINTER(Ir), ctypes.POINTER(ctypes.c_void_p)]),
    "hj_free_string": (None, [_vp]),
    "hj_kernel_get": (_i32, [_vp, ctypes.POINTER(Ir), _pvp]),
    "hj_kernel_release": (_i32, [_vp]),
    "hj_ir_compile_cubin": (_i32, [ctypes.POINTER(Ir), _pvp, ctypes.POINTER(_sz)]),
    "hj_kernel_launch": (_i32, [_vp, _vp, _sz, _vp, _pvp, _u32, _u32]),
    "hj_kernel_map_host": (_i32, [_vp, _vp, _sz, _pvp, _u32, _sz]),
    "hj_device_kernel_cache_stats": (_i32, [_vp] + [ctypes.POINTER(_u64)] * 3),
    "hj_execute_graph": (_i32, [_vp, ctypes.POINTER(Pass), _u32, _pvp, ctypes.POINTER(BufferDesc),
                                _u32, ctypes.POINTER(Report)]),
    "hj_execute_graph_cached": (_i32, [_vp, _u64, ctypes.POINTER(Pass), _u32, _pvp, ctypes.POINTER(BufferDesc),
                                       _u32, ctypes.POINTER(_u32)]),
    "hj_graph_cache_drop": (_i32, [_vp, _u64]),
    "hj_graph_cache_stats": (_i32, [_vp] + [ctypes.POINTER(_u64)] * 3),
    "hj_comm_unique_id": (_i32, [_vp]),
    "hj_comm_create": (_i32, [_vp, _vp, _i32, _i32, _pvp]),
    "hj_comm_destroy": (_i32, [_vp]),
    "hj_comm_create_local": (_i32, [_vp, _i32, _i32, _pvp, _vp]),
    "hj_comm_connect": (_i32, [_vp, _vp]),
    "hj_comm_info": (_i32, [_vp, _pi32, _pi32, _pi32, _pi32]),
    "hj_comm_device": (_i32, [_vp, _pvp]),
    "hj_sharded_prefix_sum_deferred": (_i32, [_vp, _i32, _sz, _i32, _vp, _vp, _vp]),
    "hj_apply_seed": (_i32, [_vp, _i32, _sz, _vp, _vp]),
    "hj_shard_bounds": (None, [_u64, _i32, _i32, _pu64, _pu64]),
    "hj_shard_plan": (_i32, [ctypes.POINTER(Pass), _u32, ctypes.POINTER(BufferDesc), _u32, ctypes.POINTER(ShardDesc)]),
    "hj_execute_graph_sharded": (_i32, [_vp, ctypes.POINTER(Pass), _u32, _pvp, ctypes.POINTER(BufferDesc), _u32,
                                        ctypes.POINTER(ShardDesc), ctypes.POINTER(Report)]),
    "hj_execute_graph_sharded_cached": (_i32, [_vp, _u64, ctypes.POINTER(Pass), _u32, _pvp, ctypes.POINTER(BufferDesc), _u32,
                                               ctypes.POINTER(ShardDesc), ctypes.POINTER(_u32)]),
    "hj_tr_array_sharded": (_i32, [_vp, _u32, _vp, _u64, _pu64]),
    "hj_tr_from_buffer_sharded": (_i32, [_vp, _vp, _u32, _u64, _pu64]),
    "hj_tr_var_shard": (_i32, [_u64, _pi32, _pu64, _pu64, _pi32]),
    "hj_tr_materialise": (_i32, [_u64]),
    "hj_graph_launch_sharded": (_i32, [_vp, _vp, _vp, _pu64, _u32, _pu64, ctypes.POINTER(GraphReport)]),
    "hj_sharded_reduce": (_i32, [_vp, _i32, _i32, _sz, _vp, _vp]),
    "hj_sharded_prefix_sum": (_i32, [_vp, _i32, _sz, _i32, _vp, _vp]),
    "hj_sharded_compress": (_i32, [_vp, _sz, _u32, _vp, _vp, _vp, _vp]),
    "hj_sharded_scatter_reduce": (_i32, [_vp, _i32, _i32, _sz, _vp, _vp, _u64, _vp, _sz]),
    "hj_sharded_rebalance": (_i32, [_vp, _sz, _vp, _vp, _vp, _vp, ctypes.POINTER(_u64)]),
    # ---- trace / schedule / graph
    "hj_tr_type_scalar": (_u32, [_u32]),
    "hj_tr_type_vector": (_u32, [_u32, _u32]),
    "hj_tr_type_array": (_u32, [_u32, _u32]),
    "hj_tr_type_matrix": (_u32, [_u32, _u32, _u32]),
    "hj_tr_type_struct": (_u32, [ctypes.POINTER(_u32), _u32]),
    "hj_tr_type_size": (_sz, [_u32]),
    "hj_tr_type_alignment": (_sz, [_u32]),
    "hj_tr_type_offset": (_sz, [_u32, _u32]),
    "hj_tr_type_kind": (_u32, [_u32]),
    "hj_tr_var_retain": (_i32, [_u64]),
    "hj_tr_var_release": (_i32, [_u64]),
    "hj_tr_var_info": (_i32, [_u64, _pu32, _pi32, _pu64, _pi32, _pu64, _pi32]),
    "hj_tr_var_hash": (_u64, [_u64]),
    "hj_tr_is_empty": (_i32, []),
    "hj_tr_n_live": (_u64, []),
    "hj_tr_var_buffer": (_i32, [_u64, _pvp]),
    "hj_tr_index": (_i32, [_pu64]),
    "hj_tr_sized_index": (_i32, [_u64, _pu64]),
    "hj_tr_dynamic_index": (_i32, [_u64, _u64, _pu64]),
    "hj_tr_literal": (_i32, [_u32, _u64, _pu64]),
    "hj_tr_sized_literal": (_i32, [_u32, _u64, _u64, _pu64]),
    "hj_tr_array": (_i32, [_vp, _u32, _vp, _u64, _pu64]),
    "hj_tr_array_async": (_i32, [_vp, _u32, _vp, _u64, _pu64]),
    "hj_tr_from_buffer": (_i32, [_vp, _u32, _u64, _pu64]),
    "hj_tr_bop": (_i32, [_u32, _u64, _u64, _pu64]),
    "hj_tr_uop": (_i32, [_u32, _u64, _pu64]),
    "hj_tr_cast": (_i32, [_u64, _u32, _pu64]),
    "hj_tr_bitcast": (_i32, [_u64, _u32, _pu64]),
    "hj_tr_fma": (_i32, [_u64, _u64, _u64, _pu64]),
    "hj_tr_select": (_i32, [_u64, _u64, _u64, _pu64]),
    "hj_tr_extract": (_i32, [_u64, _u32, _pu64]),
    "hj_tr_extract_dyn": (_i32, [_u64, _u64, _pu64]),
    "hj_tr_composite": (_i32, [_pu64, _u32, _pu64]),
    "hj_tr_vec": (_i32, [_pu64, _u32, _pu64]),
    "hj_tr_mat": (_i32, [_pu64, _u32, _pu64]),
    "hj_tr_arr": (_i32, [_pu64, _u32, _pu64]),
    "hj_tr_gather": (_i32, [_u64, _u64, _u64, _pu64]),
    "hj_tr_scatter": (_i32, [_u64, _u64, _u64, _u64]),
    "hj_tr_scatter_reduce": (_i32, [_u64, _u64, _u64, _u64, _u32]),
    "hj_tr_scatter_atomic": (_i32, [_u64, _u64, _u64, _u64, _u32, _pu64]),
    "hj_tr_atomic_inc": (_i32, [_u64, _u64, _u64, _pu64]),
    "hj_tr_prefix_sum": (_i32, [_u64, _i32, _pu64]),
    "hj_tr_reduce": (_i32, [_u64, _u32, _pu64]),
    "hj_tr_compress": (_i32, [_u64, _pu64, _pu64]),
    "hj_tr_compress_dyn": (_i32, [_u64, _pu64]),
    "hj_tr_scope_start": (_i32, [_i32, _pu64, _u32, _pu64, _pu64]),
    "hj_tr_scope_end": (_i32, [_u64, _pu64, _u32, _pu64]),
    "hj_tr_schedule": (_i32, [_u64]),
    "hj_tr_schedule_eval": (_i32, []),
    "hj_tr_reset_schedule": (_i32, []),
    "hj_tr_var_size": (_i32, [_u64, _pu64]),
    "hj_tr_to_host": (_i32, [_u64, _u64, _u64, _vp]),
    "hj_tr_compile": (_i32, [_pvp]),
    "hj_tr_compile_fn": (_i32, [_pu64, _u32, _pu64, _u32, _pvp]),
    "hj_ir_index_zero_fill": (_i32, [ctypes.POINTER(Ir), _u32, ctypes.POINTER(_u32), ctypes.POINTER(_i32)]),
    "hj_graph_retain": (_i32, [_vp]),
    "hj_graph_release": (_i32, [_vp]),
    "hj_graph_n_passes": (_u32, [_vp]),
    "hj_graph_n_outputs": (_u32, [_vp]),
    "hj_graph_debug_string": (_i32, [_vp, _pvp]),
    "hj_graph_pass_ir": (_i32, [_vp, _u32, ctypes.POINTER(ctypes.POINTER(Ir))]),
    "hj_graph_serialize": (_i32, [_vp, _pvp, ctypes.POINTER(_sz)]),
    "hj_graph_deserialize": (_i32, [_vp, ctypes.c_char_p, _sz, _pvp]),
    "hj_graph_launch": (_i32, [_vp, _vp, _pu64, _u32, _pu64, ctypes.POINTER(GraphReport)]),
    "hj_fcache_get": (_i32, [_u64, _pvp]),
    "hj_fcache_put": (_i32, [_u64, _vp]),
    "hj_fcache_clear": (_i32, []),
    "hj_fcache_size": (_u64, []),
}

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name, None)
    if _fn is None:
        continue  # reported by the symbol-export test; using it raises AttributeError
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    msg = lib.hj_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int) -> None:
    if status != OK:
        raise HjError(status, last_error())
