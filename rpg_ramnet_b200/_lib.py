"""ctypes binding of libramnet_sm100a.so (include/ramnet_b200.h).

The CUDA library is the product; this module only loads it and converts error codes into
exceptions.  There is deliberately no fallback: if the shared library is missing or the device is
not sm_100 every op raises.
"""
import ctypes
import os
import threading
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libramnet_sm100a.so')

MMA_FP32, MMA_TF32 = 0, 1
EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_RES_RELU, EPI_GRU_RU, EPI_GRU_OUT, EPI_LSTM, EPI_BIAS_RELU_PRED, EPI_BIAS_RELU_ADD, EPI_BIAS_ADD = range(9)
FLAG_ROUND_TF32 = 1
FLAG_HPACK = 2
FLAG_UPCONV = 4
FLAG_S2SEG = 8
FLAG_SM_TIME = 16
FLAG_DYNAMIC = 32
LOSS_LOG_SPACE = 1
NORM_INSTANCE, NORM_RUNNING, NORM_RELU, NORM_SIGMOID, NORM_ROUND_TF32 = 1, 2, 4, 8, 16
WGRAD_FULL, WGRAD_PARTIAL_FIRST, WGRAD_PARTIAL_ADD, WGRAD_FINALIZE = range(4)
RAMNET_EUNSUPPORTED = -3


class ConvDesc(ctypes.Structure):
    """struct ramnet_conv_desc."""
    _fields_ = [(n, c_int32) for n in ('N', 'H', 'W', 'C0', 'C1', 'Cout', 'ksize', 'stride', 'epilogue',
                                       'mma_kind', 'flags', 'reserved')]


# name -> (restype, argtypes); every symbol include/ramnet_b200.h declares
SIGNATURES = {
    'ramnet_version': (c_int, []),
    'ramnet_last_error': (c_char_p, []),
    'ramnet_create': (c_int, [c_int, POINTER(c_void_p)]),
    'ramnet_destroy': (c_int, [c_void_p]),
    'ramnet_sm_count': (c_int, [c_void_p]),
    'ramnet_launch_count': (c_int64, [c_void_p]),
    'ramnet_voxel_grid': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'ramnet_voxel_grid_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'ramnet_voxel_grid_ex': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                     c_void_p, c_int, c_void_p]),
    'ramnet_voxel_votes': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    'ramnet_head_conv': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_void_p]),
    'ramnet_conv_workspace_bytes': (c_size_t, [POINTER(ConvDesc)]),
    'ramnet_conv_fwd': (c_int, [c_void_p, POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'ramnet_conv_wgrad_workspace_bytes': (c_size_t, [c_void_p, POINTER(ConvDesc)]),
    'ramnet_conv_wgrad': (c_int, [c_void_p, POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_int, c_void_p]),
    'ramnet_head_conv_wgrad': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                       c_int, c_void_p]),
    'ramnet_pack_weights_dgrad': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p]),
    'ramnet_pack_weights_hpack': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'ramnet_pack_weights_s2seg': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'ramnet_upconv_packed_floats': (c_int64, [c_int, c_int]),
    'ramnet_pack_weights_upconv': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'ramnet_plan_describe': (c_int, [POINTER(ConvDesc), c_int, c_char_p, c_size_t]),
    'ramnet_voxel_normalize': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    'ramnet_depth_to_label': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p]),
    'ramnet_depth_metrics': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float, c_void_p, c_void_p]),
    'ramnet_head_im2row': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'ramnet_pack_weights_head': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'ramnet_head_conv_tc': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p]),
    'ramnet_head_conv_wgrad_tc_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    'ramnet_head_conv_wgrad_tc': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p, c_size_t, c_int, c_void_p]),
    'ramnet_pack_weights_dgrad_s2': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'ramnet_conv_dgrad_s2': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p]),
    'ramnet_zero_insert2x': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_void_p]),
    'ramnet_colsum_scratch_bytes': (c_size_t, []),
    'ramnet_relu_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    'ramnet_gru_out_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    'ramnet_gru_ru_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                  c_int, c_void_p]),
    'ramnet_lstm_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int64, c_int, c_int, c_void_p]),
    'ramnet_norm_scratch_bytes': (c_size_t, [c_int, c_int]),
    'ramnet_norm_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int,
                                c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ramnet_norm_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ramnet_pred_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    'ramnet_upsample2x_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'ramnet_pack_weights': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'ramnet_upsample2x_add': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p]),
    'ramnet_pred_sigmoid': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                    c_int, c_void_p]),
    'ramnet_nchw_to_nhwc': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'ramnet_nhwc_to_nchw': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'ramnet_round_tf32': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'ramnet_si_loss_stats': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    'ramnet_si_loss_value': (c_int, [c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p]),
    'ramnet_si_loss_grad': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float, c_float,
                                    c_void_p, c_int, c_void_p, c_void_p]),
    'ramnet_msg_pooled_count': (c_int64, [c_int, c_int, c_int, c_int, c_int]),
    'ramnet_msg_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    'ramnet_msg_loss_stats': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    'ramnet_msg_loss_value': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'ramnet_msg_loss_grad': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                     c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ramnet_msg_sobel_preview': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'ramnet_adam_step_dev': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double,
                                     c_double, c_double, c_double, c_void_p, c_int, c_void_p]),
    'ramnet_depth_output': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_float, c_float, c_void_p, c_void_p]),
    'ramnet_tf32_pipe_rate': (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
    'ramnet_adam_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double,
                                 c_double, c_double, c_double, c_int, c_void_p]),
}

_lib = None
_lock = threading.RLock()      # re-entrant: check() -> load() may run under handle()'s lock
_handles = {}


class RamnetError(RuntimeError):
    pass


def load():
    """dlopen the library and declare every prototype (no CUDA call is made)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RamnetError(
                    f'{LIB_PATH} is missing: build it with `python -m rpg_ramnet_b200.build` '
                    '(nvcc, sm_100a). There is no CPU or PyTorch fallback for the hot path.')
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        lib = _lib if _lib is not None else load()
        raise RamnetError(f'libramnet error {rc}: {lib.ramnet_last_error().decode()}')


def handle(device_index: int) -> c_void_p:
    """One ramnet_handle per CUDA device per process.  Raises RamnetError (never blocks) when the device is missing,
    out of range or not sm_100: ramnet_create runs outside the lock and its error string is read from the already
    loaded library."""
    lib = load()
    h = _handles.get(device_index)
    if h is not None:
        return h
    new = c_void_p()
    rc = lib.ramnet_create(int(device_index), ctypes.byref(new))
    if rc != 0:
        raise RamnetError(f'libramnet error {rc}: {lib.ramnet_last_error().decode()}')
    with _lock:
        h = _handles.get(device_index)
        if h is None:
            _handles[device_index] = h = new
        else:                                   # another thread won the race
            lib.ramnet_destroy(new)
    return h


def launch_count(device_index: int) -> int:
    return int(load().ramnet_launch_count(handle(device_index)))
