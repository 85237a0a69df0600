"""ctypes binding of libbokego_b200.so (C ABI in include/bokego_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing, or the current device is not
an sm_100 GPU, every entry point raises.  PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BOKEGO_B200_SO") or os.path.join(_HERE, "libbokego_b200.so")   # the override is for measurement builds (tools/)

_lib = None
_checked_devices = set()
launch_count = 0   # kernels launched through this binding (bench.py reports it as gpu_launches)


class BokegoB200Error(RuntimeError):
    pass


def _sig(L):
    vp, i32, u32, u64, f32 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float
    L.bk_version.restype = i32
    L.bk_strerror.restype = C.c_char_p
    L.bk_strerror.argtypes = [i32]
    L.bk_device_check.restype = i32
    L.bk_feats_conv_bytes.restype = C.c_size_t
    L.bk_feats_conv_bytes.argtypes = [i32]
    L.bk_weights_blob_bytes.restype = C.c_size_t
    L.bk_weights_pack.restype = i32
    L.bk_weights_pack.argtypes = [vp] * 7
    L.bk_encode.restype = i32
    L.bk_encode.argtypes = [vp] * 10 + [i32, vp]
    L.bk_repack_f32.restype = i32
    L.bk_repack_f32.argtypes = [vp, vp, i32, vp]
    L.bk_forward.restype = i32
    L.bk_forward.argtypes = [vp] * 6 + [i32, i32, vp]
    L.bk_forward_positions.restype = i32
    L.bk_forward_positions.argtypes = [vp] * 12 + [i32, i32, vp]
    L.bk_forward_debug.restype = i32
    L.bk_forward_debug.argtypes = [vp] * 6 + [i32, i32, vp, vp, i32, vp]
    L.bk_debug_words.restype = i32
    L.bk_debug_words.argtypes = [vp]
    L.bk_playout_step.restype = i32
    L.bk_playout_step.argtypes = [vp] * 8 + [i32, u64, u32, i32, i32, vp, i32, vp]
    L.bk_playout_step_encode.restype = i32
    L.bk_playout_step_encode.argtypes = [vp] * 8 + [i32, u64, u32, i32, i32, vp, vp, i32, vp]
    L.bk_playout_run.restype = i32
    L.bk_playout_run.argtypes = [vp] * 8 + [u64, u32, i32, i32, i32, i32, i32, vp, i32, vp]
    L.bk_playout_run_debug.restype = i32
    L.bk_playout_run_debug.argtypes = [vp] * 8 + [u64, u32, i32, i32, i32, i32, i32, vp, i32, vp, vp]
    L.bk_make_moves.restype = i32
    L.bk_make_moves.argtypes = [vp] * 13 + [i32, vp]
    L.bk_tree_run.restype = i32
    L.bk_tree_run.argtypes = [vp] * 7 + [i32, i32, i32, i32, C.c_double, vp, vp, vp, i32, vp, vp, C.c_double, i32]
    L.bk_tree_finish.restype = i32
    L.bk_tree_finish.argtypes = [vp] * 5 + [i32, i32, i32, vp, vp, i32]
    L.bk_score.restype = i32
    L.bk_score.argtypes = [vp, f32, vp, vp, i32, vp]
    L.bk_pack_records.restype = i32
    L.bk_pack_records.argtypes = [vp] * 5 + [i32, i32, vp]
    L.bk_exp_draws.restype = i32
    L.bk_exp_draws.argtypes = [u64, u32, u32, u32, vp, i32, vp]
    L.bk_train_param_count.restype = C.c_size_t
    L.bk_train_workspace_bytes.restype = C.c_size_t
    L.bk_train_workspace_bytes.argtypes = [i32]
    L.bk_train_launches.restype = i32
    L.bk_train_launches.argtypes = [i32, i32, i32]
    L.bk_train_conv3_schedule.restype = i32
    L.bk_train_conv3_schedule.argtypes = [i32, i32, i32, vp]
    L.bk_train_forward.restype = i32
    L.bk_train_forward.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]
    L.bk_train_backward.restype = i32
    L.bk_train_backward.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, i32, vp, vp]
    L.bk_train_running_stats.restype = i32
    L.bk_train_running_stats.argtypes = [vp, vp, vp, i32, f32, vp]
    L.bk_adamw_step.restype = i32
    L.bk_adamw_step.argtypes = [vp, vp, vp, vp, C.c_size_t] + [C.c_double] * 5 + [i32, vp]


def lib():
    """the loaded shared library (no device needed: used by the CPU-side symbol tests)"""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise BokegoB200Error(
                f"{SO_PATH} is missing: build it with `python -m bokego_b200.build` "
                "(bokego_b200 has no CPU or PyTorch fallback)")
        L = C.CDLL(SO_PATH)
        _sig(L)
        _lib = L
    return _lib


def require_device(device):
    """raise unless `device` is a CUDA device of compute capability 10.x"""
    device = torch.device(device)
    if device.type != "cuda":
        raise BokegoB200Error(f"bokego_b200 runs on sm_100 CUDA devices only, got device '{device}'")
    if not torch.cuda.is_available():
        raise BokegoB200Error("CUDA is not available; bokego_b200 has no CPU path")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _checked_devices:
        with torch.cuda.device(idx):
            rc = lib().bk_device_check()
        if rc != 0:
            raise BokegoB200Error(lib().bk_strerror(rc).decode())
        _checked_devices.add(idx)
    return torch.device("cuda", idx)


def check(rc, what):
    if rc != 0:
        raise BokegoB200Error(f"{what}: {lib().bk_strerror(rc).decode()} ({rc})")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def count_launch(n=1):
    global launch_count
    launch_count += n
