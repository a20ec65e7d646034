"""ctypes binding of libmptg.so -- the C ABI declared in include/mptg/mptg.h.

The library is the product; there is no Python or CPU fallback.  Importing this module fails loudly
when the shared object is missing (run `python -m mpt_b200.build`), and creating a Context fails
loudly when there is no sm_100a device.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# MPTG_LIB: another build of the same library (kernel experiments, tools/build_variant.sh)
LIB_PATH = Path(os.environ.get("MPTG_LIB") or Path(__file__).resolve().parent / "_lib" / "libmptg.so")

MAX_PARTS = 8
MAX_SCALARS = 64
MAX_K = 128
NO_INDEX = 0xFFFFFFFF

OK, ERR_BAD_ARG, ERR_OOM, ERR_CUDA, ERR_NCCL, ERR_CAPACITY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
PART_LP, PART_SO2, PART_SO3 = 1, 2, 3
F32, F64 = 4, 8
KNN_AUTO, KNN_BRUTE, KNN_BVH = 0, 1, 2
GEOM_GRID, GEOM_SHAPES, GEOM_LINKARM, GEOM_MESH, GEOM_NAOCUP = 1, 2, 3, 4, 5


class SpacePart(C.Structure):
    _fields_ = [("kind", C.c_int32), ("p", C.c_int32), ("dim", C.c_int32), ("_pad", C.c_int32), ("weight", C.c_double)]


class SpaceDesc(C.Structure):
    _fields_ = [("n_parts", C.c_int32), ("scalar", C.c_int32), ("part", SpacePart * MAX_PARTS)]


class PrrtParams(C.Structure):
    _fields_ = [("space", C.POINTER(SpaceDesc)), ("lo", C.c_void_p), ("hi", C.c_void_p), ("range", C.c_double),
                ("goal_bias", C.c_double), ("goal_state", C.c_void_p), ("goal_radius", C.c_double), ("link_step", C.c_double),
                ("seed", C.c_uint64), ("capacity", C.c_uint32), ("max_wave", C.c_uint32)]


class PprmParams(C.Structure):
    _fields_ = [("space", C.POINTER(SpaceDesc)), ("lo", C.c_void_p), ("hi", C.c_void_p), ("goal_state", C.c_void_p), ("goal_radius", C.c_double),
                ("link_step", C.c_double), ("seed", C.c_uint64), ("capacity", C.c_uint32), ("max_wave", C.c_uint32), ("max_k", C.c_uint32)]


class MptgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmptg error {code}: {msg}")
        self.code = code


# every symbol include/mptg/mptg.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_U32 = C.c_uint32
_U32P = C.POINTER(C.c_uint32)
_U64P = C.POINTER(C.c_uint64)
_SD = C.POINTER(SpaceDesc)
SYMBOLS = {
    "mptg_abi_version": (C.c_int, []),
    "mptg_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "mptg_ctx_destroy": (C.c_int, [_P]),
    "mptg_sync": (C.c_int, [_P]),
    "mptg_last_error": (C.c_char_p, [_P]),
    "mptg_ctx_stream": (_P, [_P]),
    "mptg_ctx_launch_count": (C.c_uint64, [_P]),
    "mptg_ctx_sm_count": (C.c_int, [_P]),
    "mptg_probe_fp32_tflops": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "mptg_space_scalars": (C.c_int, [_SD]),
    "mptg_space_dimensions": (C.c_int, [_SD]),
    "mptg_distance_batch": (C.c_int, [_P, _SD, _P, _P, _U32, _P]),
    "mptg_interpolate_batch": (C.c_int, [_P, _SD, _P, _P, _P, _U32, _P]),
    "mptg_knn_create": (C.c_int, [_P, _SD, _U32, C.POINTER(_P)]),
    "mptg_knn_destroy": (C.c_int, [_P]),
    "mptg_knn_set_strategy": (C.c_int, [_P, C.c_int]),
    "mptg_knn_set_index_map": (C.c_int, [_P, _U32, _U32]),
    "mptg_knn_insert": (C.c_int, [_P, _P, _U32, _U32P]),
    "mptg_knn_insert_dev": (C.c_int, [_P, _P, _U32, _U32P]),
    "mptg_knn_size": (_U32, [_P]),
    "mptg_knn_get_states": (C.c_int, [_P, _U32, _U32, _P]),
    "mptg_knn_query": (C.c_int, [_P, _P, _U32, _U32, C.c_double, _P, _P, _P]),
    "mptg_knn_query_dev": (C.c_int, [_P, _P, _U32, _U32, C.c_double, _P, _P, _P]),
    "mptg_knn_build_index": (C.c_int, [_P]),
    "mptg_knn_last_stats": (C.c_int, [_P, _U64P]),
    "mptg_knn_merge_dev": (C.c_int, [_P, C.c_int, _U32, _U32, _U32, _P, _P, _P, _P, _P]),
    "mptg_comm_unique_id": (C.c_int, [_P]),
    "mptg_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(_P)]),
    "mptg_comm_destroy": (C.c_int, [_P]),
    "mptg_comm_rank": (C.c_int, [_P]),
    "mptg_comm_world": (C.c_int, [_P]),
    "mptg_comm_slice": (C.c_int, [_P, _U32, _U32P, _U32P]),
    "mptg_knn_insert_ids": (C.c_int, [_P, _P, _P, _U32]),
    "mptg_knn_shard_sync": (C.c_int, [_P, _P]),
    "mptg_knn_query_sharded": (C.c_int, [_P, _P, _P, _U32, _U32, C.c_double, _P, _P, _P]),
    "mptg_knn_query_sharded_dev": (C.c_int, [_P, _P, _P, _U32, _U32, C.c_double, _P, _P, _P]),
    "mptg_grid_create": (C.c_int, [_P, C.c_int, C.c_int32, C.c_int32, _P, C.POINTER(_P)]),
    "mptg_shapes_create": (C.c_int, [_P, C.c_int, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, C.POINTER(_P)]),
    "mptg_linkarm_create": (C.c_int, [_P, C.c_int, C.c_int32, _P, C.c_double, C.c_int32, _P, C.POINTER(_P)]),
    "mptg_naocup_create": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "mptg_naocup_configs": (C.c_int, [C.c_int, _P, _P, _P, _P]),
    "mptg_mesh_pair_create": (C.c_int, [_P, C.c_int, _U32, _P, _U32, _P, C.POINTER(_P)]),
    "mptg_geom_destroy": (C.c_int, [_P]),
    "mptg_geom_kind": (C.c_int, [_P]),
    "mptg_valid_batch": (C.c_int, [_P, _P, _U32, _P, _P]),
    "mptg_valid_batch_dev": (C.c_int, [_P, _P, _U32, _P, _P]),
    "mptg_link_batch": (C.c_int, [_P, _SD, _P, _P, _U32, C.c_double, _P, _P]),
    "mptg_link_batch_dev": (C.c_int, [_P, _SD, _P, _P, _U32, C.c_double, _P, _P]),
    "mptg_geom_contact_band": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "mptg_geom_last_stats": (C.c_int, [_P, _U64P]),
    "mptg_steer_batch": (C.c_int, [_P, _SD, _P, _P, _P, _U32, C.c_double, _P, _P]),
    "mptg_space_uniforms": (C.c_int, [_SD]),
    "mptg_sample_batch": (C.c_int, [_P, _SD, _P, _P, C.c_uint64, C.c_uint64, _U32, _P]),
    "mptg_sample_batch_dev": (C.c_int, [_P, _SD, _P, _P, C.c_uint64, C.c_uint64, _U32, _P]),
    "mptg_sample_transform_batch": (C.c_int, [_P, _SD, _P, _P, _P, _U32, _P]),
    "mptg_prrt_create": (C.c_int, [_P, _P, _P, C.POINTER(_P)]),
    "mptg_prrt_destroy": (C.c_int, [_P]),
    "mptg_prrt_add_start": (C.c_int, [_P, _P]),
    "mptg_prrt_wave": (C.c_int, [_P, _U32, _U32P, _U32P]),
    "mptg_prrt_size": (_U32, [_P]),
    "mptg_prrt_samples_drawn": (C.c_uint64, [_P]),
    "mptg_prrt_get_tree": (C.c_int, [_P, _U32, _U32, _P, _P]),
    "mptg_prrtstar_create": (C.c_int, [_P, _P, _P, C.c_double, C.POINTER(_P)]),
    "mptg_prrtstar_destroy": (C.c_int, [_P]),
    "mptg_prrtstar_set_rewire_radius": (C.c_int, [_P, C.c_double]),
    "mptg_prrtstar_add_start": (C.c_int, [_P, _P]),
    "mptg_prrtstar_wave": (C.c_int, [_P, _U32, _U32P, _U32P]),
    "mptg_prrtstar_size": (_U32, [_P]),
    "mptg_prrtstar_samples_drawn": (C.c_uint64, [_P]),
    "mptg_prrtstar_rewires": (C.c_uint64, [_P]),
    "mptg_prrtstar_get_tree": (C.c_int, [_P, _U32, _U32, _P, _P, _P]),
    "mptg_pprm_create": (C.c_int, [_P, _P, _P, C.POINTER(_P)]),
    "mptg_pprm_destroy": (C.c_int, [_P]),
    "mptg_pprm_set_spanner": (C.c_int, [_P, C.c_double, C.c_uint32]),
    "mptg_pprm_add_state": (C.c_int, [_P, _P, _U32, _U32P]),
    "mptg_pprm_wave": (C.c_int, [_P, _U32, _U32P, _U32P]),
    "mptg_pprm_size": (_U32, [_P]),
    "mptg_pprm_samples_drawn": (C.c_uint64, [_P]),
    "mptg_pprm_row_stride": (_U32, [_P]),
    "mptg_pprm_get_graph": (C.c_int, [_P, _U32, _U32, _P, _P, _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load libmptg.so and bind every declared symbol (raises if the library or a symbol is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m mpt_b200.build` (nvcc, sm_100a). "
            "mpt_b200 has no CPU or pure-Python fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, ctx_handle=None) -> None:
    if rc != OK:
        msg = load().mptg_last_error(ctx_handle)
        raise MptgError(rc, msg.decode() if msg else "")
