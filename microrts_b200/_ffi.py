"""ctypes binding of libmicrorts_cuda.so (include/microrts_cuda.h).

There is no CPU fallback: if the CUDA library has not been built (python -c "import __graft_entry__ as g; g.build()")
importing the binding fails loudly, and every compute call fails with MRTS_E_CUDA when no GPU is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MRTS_CUDA_LIB may point at another build of the same CUDA library (kernel tuning experiments); never a CPU path.
LIB_PATH = os.environ.get("MRTS_CUDA_LIB") or os.path.join(_HERE, "libmicrorts_cuda.so")
_lib = None


class StateHost(C.Structure):
    _fields_ = [("header", C.POINTER(C.c_int32)), ("units", C.POINTER(C.c_int32)), ("actions", C.POINTER(C.c_int32)),
                ("rng", C.POINTER(C.c_int64))]


class MctsParams(C.Structure):
    _fields_ = [("lookahead", C.c_int), ("max_depth", C.c_int), ("epsilon_l", C.c_float), ("epsilon_g", C.c_float), ("epsilon_0", C.c_float),
                ("global_strategy", C.c_int), ("force_exploration", C.c_int), ("eval_fn", C.c_int), ("algorithm", C.c_int)]


def bind(L):
    vp, i, i64, u32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32
    pvp = C.POINTER(C.c_void_p)
    sig = {
        "mrts_abi_version": (i, []), "mrts_last_error": (C.c_char_p, []),
        "mrts_utt_create": (i, [i, i, pvp]), "mrts_utt_from_json": (i, [C.c_char_p, pvp]),
        "mrts_utt_num_types": (i, [vp]), "mrts_utt_get": (i, [vp, i, i]), "mrts_utt_type_name": (C.c_char_p, [vp, i]),
        "mrts_utt_conflict_policy": (i, [vp]), "mrts_utt_max_attack_range": (i, [vp]), "mrts_utt_destroy": (None, [vp]),
        "mrts_map_load_xml": (i, [C.c_char_p, vp, pvp]), "mrts_map_from_xml": (i, [C.c_char_p, vp, pvp]),
        "mrts_map_create": (i, [i, i, vp, i, i, i, vp, vp, pvp]),
        "mrts_map_width": (i, [vp]), "mrts_map_height": (i, [vp]), "mrts_map_num_units": (i, [vp]),
        "mrts_map_get_units": (i, [vp, vp]), "mrts_map_get_terrain": (i, [vp, vp]), "mrts_map_resources": (i, [vp, i]),
        "mrts_map_destroy": (None, [vp]),
        "mrts_batch_create": (i, [vp, pvp, i, i64, i, u32, i, pvp]), "mrts_batch_destroy": (None, [vp]),
        "mrts_batch_num_games": (i64, [vp]), "mrts_batch_unit_capacity": (i, [vp]), "mrts_batch_device": (i, [vp]),
        "mrts_batch_stream": (vp, [vp]), "mrts_batch_sync": (i, [vp]),
        "mrts_batch_reset": (i, [vp, vp, i]), "mrts_batch_reset_masked": (i, [vp, vp, vp, i]),
        "mrts_batch_copy_games": (i, [vp, vp, vp, vp, i]),
        "mrts_batch_set_policy": (i, [vp, i, i, i]), "mrts_batch_set_auto_reset": (i, [vp, i]),
        "mrts_batch_set_actions": (i, [vp, i, i, vp, vp, i, i, i]),
        "mrts_batch_issue": (i, [vp, i, i, vp, vp, i, i, i, i]),
        "mrts_batch_step": (i, [vp, i, i]), "mrts_batch_set_observation_outputs": (i, [vp, i, vp, vp]), "mrts_batch_set_mask_outputs": (i, [vp, vp, vp]), "mrts_batch_set_output_stride": (i, [vp, i]), "mrts_batch_set_vec_autoreset": (i, [vp, i, i]),
        "mrts_batch_set_actions_interleaved": (i, [vp, i, vp, i, i, i, i]), "mrts_batch_vec_step": (i, [vp, vp, i, i, i]),
        "mrts_batch_set_issue_order": (i, [vp, i]), "mrts_batch_set_info_output": (i, [vp, vp]), "mrts_batch_restart_masked": (i, [vp, vp, i]), "mrts_batch_cycle_to": (i, [vp, vp, i, i]),
        "mrts_batch_rollout": (i, [vp, i, i, i, i, i, vp, vp, vp, i]),
        "mrts_batch_observe": (i, [vp, i, i, vp, i]), "mrts_batch_num_planes": (i, [vp]),
        "mrts_batch_pathfind": (i, [vp, i, vp, vp, i]), "mrts_batch_evaluate": (i, [vp, i, i, i, vp, i]),
        "mrts_batch_masks": (i, [vp, i, i, vp, i]), "mrts_batch_mask_width": (i, [vp]),
        "mrts_batch_export": (i, [vp, i64, i64, C.POINTER(StateHost)]),
        "mrts_batch_import": (i, [vp, i64, i64, C.POINTER(StateHost)]),
        "mrts_batch_results": (i, [vp, vp, i]), "mrts_batch_cycle_to_decision": (i, [vp]),
        "mrts_batch_unit_actions": (i, [vp, i, i, i, i, vp, vp, vp, vp, i]), "mrts_batch_copy_to_host": (i, [vp, vp, vp, C.c_size_t]), "mrts_batch_stats": (i, [vp, vp]),
        "mrts_batch_scatter_games": (i, [vp, vp, vp, i]),
        "mrts_batch_player_actions": (i, [vp, i64, i, vp, vp, i64, i, C.POINTER(C.c_int64)]),
        "mrts_pag_create": (i, [vp, i64, i, i, pvp]), "mrts_pag_destroy": (None, [vp]), "mrts_pag_size": (i64, [vp]), "mrts_pag_generated": (i64, [vp]),
        "mrts_pag_num_choices": (i, [vp]), "mrts_pag_next": (i, [vp, vp, i]), "mrts_pag_random": (i, [vp, C.POINTER(C.c_int64), vp, i]),
        "mrts_pag_randomize_order": (i, [vp, C.POINTER(C.c_int64)]), "mrts_java_random_seed": (i64, [i64]),
        "mrts_mcts_create": (i, [vp, i, C.POINTER(MctsParams), i, vp, pvp]), "mrts_mcts_iterate": (i, [vp, i]), "mrts_mcts_num_nodes": (i, [vp, i64]),
        "mrts_mcts_root": (i, [vp, i64, C.POINTER(C.c_int32), C.POINTER(C.c_double), vp, vp, i]), "mrts_mcts_best_actions": (i, [vp, vp, vp, i]),
        "mrts_mcts_destroy": (None, [vp]),
        "mrts_batch_launch_count": (i64, [vp]), "mrts_batch_last_kernel": (C.c_char_p, [vp]), "mrts_batch_io_bytes": (i, [vp, vp]),
        "mrts_nccl_unique_id": (i, [vp]), "mrts_nccl_comm_create": (i, [vp, i, i, i, pvp]), "mrts_nccl_comm_wrap": (i, [vp, i, pvp]),
        "mrts_nccl_comm_destroy": (None, [vp]), "mrts_batch_stats_allreduce": (i, [vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError here means the library does not export a declared symbol
        f.restype = res
        f.argtypes = args
    return L


EXPORTED_SYMBOLS = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "microrts_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib
