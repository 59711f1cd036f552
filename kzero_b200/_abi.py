"""ctypes binding of libkzb200.so (include/kzb200.h).  Fails loudly if the library is missing."""
from __future__ import annotations

import ctypes
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libkzb200.so"
_lib = None


class NetInfo(ctypes.Structure):
    _fields_ = [("input_channels", ctypes.c_int32), ("board_h", ctypes.c_int32), ("board_w", ctypes.c_int32),
                ("policy_len", ctypes.c_int32), ("channels", ctypes.c_int32), ("depth", ctypes.c_int32),
                ("max_batch", ctypes.c_int32), ("precision", ctypes.c_int32), ("device", ctypes.c_int32),
                ("conv_mode", ctypes.c_int32), ("flops_per_position", ctypes.c_double)]


_fp = ctypes.POINTER(ctypes.c_float)


class ConvWeights(ctypes.Structure):
    """kzb_conv_weights (include/kzb200.h)."""
    _fields_ = [("cin", ctypes.c_int32), ("cout", ctypes.c_int32), ("ksize", ctypes.c_int32), ("w", _fp), ("b", _fp)]


class FcWeights(ctypes.Structure):
    """kzb_fc_weights (include/kzb200.h)."""
    _fields_ = [("in_", ctypes.c_int32), ("out", ctypes.c_int32), ("w", _fp), ("b", _fp)]


class NetSpecC(ctypes.Structure):
    """kzb_net_spec (include/kzb200.h)."""
    _fields_ = [("input_channels", ctypes.c_int32), ("board_h", ctypes.c_int32), ("board_w", ctypes.c_int32), ("channels", ctypes.c_int32),
                ("depth", ctypes.c_int32), ("first", ConvWeights), ("blocks", ctypes.POINTER(ConvWeights)), ("final_scale", _fp),
                ("final_shift", _fp), ("scalar_conv", ConvWeights), ("fc1", FcWeights), ("fc2", FcWeights), ("policy_conv1", ConvWeights),
                ("policy_conv2", ConvWeights), ("has_extra", ctypes.c_int32), ("extra_conv", ConvWeights), ("extra_fc", FcWeights),
                ("policy_len", ctypes.c_int32), ("policy_src", ctypes.POINTER(ctypes.c_int32))]


class SelfplayConfig(ctypes.Structure):
    """kzb_selfplay_config (include/kzb200.h)."""
    _fields_ = [("game", ctypes.c_int32), ("visits", ctypes.c_int32), ("search_batch", ctypes.c_int32), ("gpu_batch", ctypes.c_int32),
                ("cpu_threads", ctypes.c_int32), ("gpu_threads", ctypes.c_int32), ("concurrent_games", ctypes.c_int32),
                ("max_game_length", ctypes.c_int32), ("cache_size", ctypes.c_int32), ("zero_temp_move_count", ctypes.c_int32),
                ("max_moves", ctypes.c_int32), ("max_games", ctypes.c_int32), ("part_iterations", ctypes.c_int32),
                ("full_search_prob", ctypes.c_float), ("duration_s", ctypes.c_float), ("temperature", ctypes.c_float),
                ("dirichlet_alpha", ctypes.c_float), ("dirichlet_eps", ctypes.c_float),
                ("policy_temperature_root", ctypes.c_float), ("policy_temperature_child", ctypes.c_float),
                ("exploration_weight", ctypes.c_float), ("moves_left_weight", ctypes.c_float), ("moves_left_clip", ctypes.c_float),
                ("moves_left_sharpness", ctypes.c_float), ("fpu_root", ctypes.c_float), ("fpu_root_relative", ctypes.c_int32),
                ("fpu_child", ctypes.c_float), ("fpu_child_relative", ctypes.c_int32), ("virtual_loss", ctypes.c_float),
                ("q_mode_wdl", ctypes.c_int32), ("draw_score", ctypes.c_float), ("executor_blocking_sync", ctypes.c_int32),
                ("dummy_network", ctypes.c_int32), ("output_prefix", ctypes.c_char_p), ("seed", ctypes.c_uint64)]


class SelfplayStats(ctypes.Structure):
    """kzb_selfplay_stats (include/kzb200.h)."""
    _fields_ = [("seconds", ctypes.c_double), ("real_evals", ctypes.c_uint64), ("cached_evals", ctypes.c_uint64),
                ("potential_evals", ctypes.c_uint64), ("batches", ctypes.c_uint64), ("max_batch", ctypes.c_uint64),
                ("games_finished", ctypes.c_uint64), ("moves_played", ctypes.c_uint64), ("root_visits", ctypes.c_uint64),
                ("concurrent_games", ctypes.c_uint64), ("games_written", ctypes.c_uint64), ("interrupted", ctypes.c_uint64)]


class MctsTraceOut(ctypes.Structure):
    """kzb_mcts_trace_out (include/kzb200.h)."""
    _fields_ = [("capacity", ctypes.c_int32), ("n_children", ctypes.c_int32), ("child_visits", ctypes.POINTER(ctypes.c_uint64)),
                ("child_moves", ctypes.POINTER(ctypes.c_uint32)), ("child_policy", ctypes.POINTER(ctypes.c_float)),
                ("root_values", ctypes.c_float * 5), ("root_visits", ctypes.c_uint64), ("tree_nodes", ctypes.c_uint64),
                ("evals", ctypes.c_uint64)]


# every symbol include/kzb200.h declares: name -> (restype, argtypes)
_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
SYMBOLS = {
    "kzb_device_count": (_i, []),
    "kzb_last_error": (ctypes.c_char_p, []),
    "kzb_net_create_from_onnx": (_i, [_i, _vp, _sz, _i, _i, ctypes.POINTER(_vp)]),
    "kzb_net_create": (_i, [_i, ctypes.POINTER(NetSpecC), _i, _i, ctypes.POINTER(_vp)]),
    "kzb_net_bind_mapper": (_i, [_vp, _i, _i, _i, _i, _i]),
    "kzb_net_get_info": (_i, [_vp, ctypes.POINTER(NetInfo)]),
    "kzb_onnx_inspect": (_i, [_vp, _sz, ctypes.POINTER(NetInfo)]),
    "kzb_net_destroy": (None, [_vp]),
    "kzb_eval_planes": (_i, [_vp, _vp, _i, _vp, _vp]),
    "kzb_eval_packed": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "kzb_encode_planes": (_i, [_vp, _vp, _vp, _i, _vp]),
    "kzb_net_set_symmetries": (_i, [_vp, _i, _vp, _vp]),
    "kzb_eval_packed_sym": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "kzb_stage_packed": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "kzb_time_staged": (_i, [_vp, _i, _i, _vp]),
    "kzb_profile_staged": (_i, [_vp, _i, _vp, _sz, _vp, _i, ctypes.POINTER(_i)]),
    "kzb_launches_per_eval": (_i, [_vp]),
    "kzb_selfplay_default_config": (None, [ctypes.POINTER(SelfplayConfig)]),
    "kzb_selfplay_run": (_i, [_i, _vp, _sz, _i, ctypes.POINTER(SelfplayConfig), ctypes.POINTER(SelfplayStats)]),
    "kzb_selfplay_request_stop": (None, []),
    "kzb_selfplay_clear_stop": (None, []),
    "kzb_selfplay_request_interrupt": (None, []),
    "kzb_selfplay_clear_interrupt": (None, []),
    "kzb_selfplay_session_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "kzb_selfplay_session_run": (_i, [_vp, _i, _vp, _sz, _i, ctypes.POINTER(SelfplayConfig), ctypes.POINTER(SelfplayStats)]),
    "kzb_selfplay_session_destroy": (None, [_vp]),
    "kzb_mcts_trace": (_i, [ctypes.POINTER(SelfplayConfig), ctypes.c_uint64, _i, _i, ctypes.POINTER(MctsTraceOut)]),
}


def lib_path() -> Path:
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(f"{_LIB_PATH} is missing: build it with `python -m kzero_b200.build` "
                               "(there is no CPU fallback)")
        L = ctypes.CDLL(str(_LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class KzbError(RuntimeError):
    """A non-zero return from the C ABI; the reference panics in the same situations."""


def check(rc: int) -> None:
    if rc != 0:
        raise KzbError(lib().kzb_last_error().decode(errors="replace"))
