"""ctypes binding of libfocr_sm100.so (the C-ABI declared in include/focr.h).

The library is the product: there is no Python/PyTorch fallback for any op.  Importing this module
on a machine where the .so has not been built raises immediately.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libfocr_sm100.so"


class FocrError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not _LIB_PATH.exists():
        raise FocrError(
            f"{_LIB_PATH} is missing: run `python -m fudanocr_b200.build` (needs nvcc). "
            "There is no CPU / PyTorch fallback for the focr kernels.")
    return C.CDLL(str(_LIB_PATH))


lib = _load()
lib.focr_last_error.restype = C.c_char_p
lib.focr_version.restype = C.c_int

_vp, _i, _l, _sz, _fp = C.c_void_p, C.c_int, C.c_long, C.c_size_t, C.c_void_p

_ll, _u, _f = C.c_longlong, C.c_uint, C.c_float
_pp = C.POINTER(C.c_void_p)

_SIGS = {
    "focr_sync_check": (C.c_int, [_vp]),
    "focr_conv2d_workspace_bytes": (_sz, [_i, _i, _i]),
    "focr_conv2d_fwd": (C.c_int, [_vp, _fp, _fp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "focr_conv2d_dgrad": (C.c_int, [_vp, _fp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "focr_wgrad_workspace_bytes": (_sz, []),
    "focr_conv2d_wgrad": (C.c_int, [_vp, _vp, _fp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "focr_linear_workspace_bytes": (_sz, [_i, _i]),
    "focr_linear_fwd": (C.c_int, [_vp, _fp, _fp, _vp, _vp, _l, _i, _i, _i, _vp, _sz, _vp]),
    "focr_linear_dgrad": (C.c_int, [_vp, _fp, _vp, _l, _i, _i, _vp, _sz, _vp]),
    "focr_linear_wgrad": (C.c_int, [_vp, _vp, _fp, _l, _i, _i, _vp, _sz, _vp]),
    "focr_bias_grad": (C.c_int, [_vp, _fp, _l, _i, _vp, _sz, _vp]),
    "focr_linear_wgrad_bias": (C.c_int, [_vp, _vp, _fp, _fp, _l, _i, _i, _vp, _sz, _vp]),
    "focr_bn_workspace_bytes": (_sz, []),
    "focr_bn_train_fwd": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _vp, _vp, _fp, _l, _i, _i, _vp, _sz, _vp]),
    "focr_bn_bwd": (C.c_int, [_vp, _vp, _fp, _vp, _fp, _fp, _l, _i, _i, _vp, _sz, _vp]),
    "focr_layernorm_std_fwd": (C.c_int, [_vp, _fp, _fp, _vp, _l, _f, _vp]),
    "focr_layernorm_std_bwd": (C.c_int, [_vp, _vp, _fp, _vp, _fp, _fp, _l, _f, _vp, _sz, _vp]),
    "focr_mha_drop_bits_bytes": (_sz, [_i]),
    "focr_mha_flash_fwd": (C.c_int, [_vp, _vp, _fp, _i, _f, _u, _u, _vp, _vp]),
    "focr_mha_bwd_workspace_bytes": (_sz, [_i]),
    "focr_mha_flash_bwd": (C.c_int, [_vp, _vp, _vp, _fp, _vp, _sz, _vp, _i, _f, _u, _u, _vp, _vp]),
    "focr_umma_probe": (C.c_int, [_vp, _i, C.c_ulonglong, C.c_ulonglong, _u, _i, _u, _u, _vp, _i, _vp]),
    "focr_attn_set_force_exact": (C.c_int, [_i]),
    "focr_attn_set_bwd_two_pass": (C.c_int, [_i]),
    "focr_recog_decode_prepared_bytes": (_sz, [_i, _i, _i]),
    "focr_recog_decode_prepare": (C.c_int, [_pp, _i, _i, _fp, _i, _vp, _sz, _vp]),
    "focr_recog_decode_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "focr_recog_decode": (C.c_int, [_vp, _sz, _i, _i, _i, _vp, _i, _i, _i, _vp, _fp, _vp, _sz, _vp]),
    "focr_weight_cross_entropy_workspace_bytes": (_sz, [_l]),
    "focr_weight_cross_entropy": (C.c_int, [_fp, _vp, _fp, _fp, _fp, _vp, _l, _i, _vp, _sz, _vp]),
    "focr_to_gray": (C.c_int, [_fp, _fp, _l, _i, _l, _vp]),
    "focr_to_gray_bwd": (C.c_int, [_fp, _fp, _l, _i, _l, _vp]),
    "focr_mse_loss_grad": (C.c_int, [_fp, _fp, _fp, _fp, _l, _f, _vp, _sz, _vp]),
    "focr_adam_clip_step": (C.c_int, [_vp, _i, _f, _f, _f, _f, _f, _f, _vp, _fp, _vp, _sz, _vp]),
    "focr_tbsrn_num_slots": (C.c_int, [_i]),
    "focr_tbsrn_slot_name": (C.c_char_p, [_i, _i]),
    "focr_tbsrn_workspace_bytes": (_sz, [_i, _i]),
    "focr_tbsrn_forward": (C.c_int, [_pp, _fp, _fp, _i, _i, _i, _f, _u, _vp, _sz, _vp]),
    "focr_tbsrn_forward_devseed": (C.c_int, [_pp, _fp, _fp, _i, _i, _i, _f, _vp, _vp, _sz, _vp]),
    "focr_tbsrn_backward": (C.c_int, [_pp, _pp, _fp, _fp, _i, _i, _i, _f, _u, _vp, _sz, _vp]),
    "focr_crnn_num_slots": (C.c_int, []),
    "focr_crnn_workspace_bytes": (_sz, [_i]),
    "focr_bicubic_gray_32x100": (C.c_int, [_fp, _fp, _i, _vp]),
    "focr_crnn_forward": (C.c_int, [_pp, _fp, _i, _fp, _i, _vp, _sz, _vp]),
    "focr_ctc_greedy_decode": (C.c_int, [_fp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "focr_psnr_ssim_workspace_bytes": (_sz, [_i]),
    "focr_psnr_ssim": (C.c_int, [_fp, _fp, _i, _i, _fp, _fp, _fp, _vp, _sz, _vp]),
    "focr_ctc_loss_workspace_bytes": (_sz, [_i, _i, _i]),
    "focr_ctc_loss": (C.c_int, [_fp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _i, _i, _f, _fp, _fp, _fp, _vp, _sz, _vp]),
    "focr_ctc_loss_status": (C.c_int, [_vp, _i, _i, _i, C.POINTER(C.c_int), _vp]),
    "focr_clip_contrastive_workspace_bytes": (_sz, [_i, _i]),
    "focr_clip_contrastive_loss": (C.c_int, [_fp, _fp, _fp, _vp, _i, _i, _fp, _fp, _fp, _fp, _vp, _vp, _sz, _vp]),
    "focr_recog_epoch_set": (C.c_int, [_u, _vp]),
    "focr_recog_epoch_advance": (C.c_int, [_vp]),
    "focr_mha_small_fwd": (C.c_int, [_vp, _l, _vp, _l, _vp, _l, _vp, _l, _fp, _i, _i, _i, _i, _i, _i, _f, _u, _u, _vp]),
    "focr_mha_small_bwd": (C.c_int, [_vp, _l, _vp, _l, _vp, _l, _vp, _l, _fp, _vp, _l, _vp, _l, _vp, _l, _i, _i, _i, _i, _i, _i,
                                     _f, _vp]),
    "focr_mha_small_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "focr_mha_small_bwd_ws": (C.c_int, [_vp, _l, _vp, _l, _vp, _l, _vp, _l, _fp, _vp, _l, _vp, _l, _vp, _l, _i, _i, _i, _i, _i, _i,
                                        _f, _vp, _sz, _vp]),
    "focr_layernorm_wide_fwd": (C.c_int, [_vp, _vp, _fp, _fp, _vp, _vp, _l, _i, _f, _vp]),
    "focr_layernorm_wide_workspace_bytes": (_sz, [_i]),
    "focr_layernorm_wide_bwd": (C.c_int, [_vp, _vp, _fp, _vp, _fp, _fp, _l, _i, _f, _vp, _sz, _vp]),
    "focr_text_embed_fwd": (C.c_int, [_vp, _fp, _i, _i, _i, _i, _l, _vp, _f, _u, _u, _vp, _vp]),
    "focr_text_embed_bwd": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _fp, _vp]),
    "focr_packed_ce_workspace_bytes": (_sz, [_i]),
    "focr_packed_ce": (C.c_int, [_fp, _l, _i, _i, _i, _vp, _vp, _f, _fp, _vp, _l, _vp, _sz, _vp]),
    "focr_dropout": (C.c_int, [_vp, _vp, _l, _f, _u, _u, _vp]),
    "focr_add_relu": (C.c_int, [_vp, _vp, _vp, _l, _vp]),
    "focr_relu_bwd": (C.c_int, [_vp, _vp, _vp, _l, _vp]),
    "focr_maxpool2x2_fwd": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "focr_maxpool2x2_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "focr_adadelta_step": (C.c_int, [_vp, _i, _f, _f, _f, _f, _f, _vp]),
    "focr_bn_eval_fwd": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _vp, _fp, _l, _i, _i, _vp]),
    "focr_l2norm_rows_fwd": (C.c_int, [_fp, _l, _vp, _fp, _l, _i, _vp]),
    "focr_l2norm_rows_bwd": (C.c_int, [_vp, _vp, _fp, _vp, _l, _l, _i, _vp]),
    "focr_packed_feat_mse": (C.c_int, [_vp, _i, _i, _i, _vp, _vp, _fp, _i, _f, _fp, _vp, _vp, _sz, _vp]),
    "focr_conv3x3_wgrad_tc_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "focr_conv3x3_wgrad_tc": (C.c_int, [_vp, _vp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "focr_conv3x3_gemm_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "focr_conv3x3_gemm_fwd": (C.c_int, [_vp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "focr_conv3x3_gemm_wgrad": (C.c_int, [_vp, _vp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "focr_resize_bicubic_normalize": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _fp, _vp, _vp]),
    "focr_prof_enable": (C.c_int, [_i, C.c_char_p]),
    "focr_prof_collect": (C.c_int, [C.c_char_p, _i]),
    "focr_launch_count": (_ll, []),
    "focr_tsrn_num_slots": (C.c_int, [_i]),
    "focr_tsrn_slot_name": (C.c_char_p, [_i, _i]),
    "focr_tsrn_workspace_bytes": (_sz, [_i, _i]),
    "focr_tsrn_forward": (C.c_int, [_pp, _fp, _fp, _i, _i, _i, _vp, _sz, _vp]),
    "focr_tsrn_backward": (C.c_int, [_pp, _pp, _fp, _fp, _i, _i, _i, _vp, _sz, _vp]),
    "focr_tsrn_ws_tensor": (C.c_int, [_i, _i, C.c_char_p, C.POINTER(_ll), C.POINTER(_ll), C.POINTER(_i)]),
    "focr_tbsrn_ws_tensor": (C.c_int, [_i, _i, C.c_char_p, C.POINTER(_ll), C.POINTER(_ll), C.POINTER(_i)]),
    "focr_strokenet_num_slots": (C.c_int, []),
    "focr_strokenet_slot_name": (C.c_char_p, [_i, _i]),
    "focr_strokenet_prepared_bytes": (_sz, [_i]),
    "focr_strokenet_prepare": (C.c_int, [_pp, _i, _vp, _sz, _vp]),
    "focr_focus_loss_workspace_bytes": (_sz, [_i, _i]),
    "focr_focus_loss": (C.c_int, [_vp, _sz, _i, _fp, _fp, _vp, _i, _i, _f, _f, _fp, _fp, _fp, _fp, _vp, _sz, _vp]),
    "focr_text_focus_loss": (C.c_int, [_vp, _sz, _i, _fp, _fp, _vp, _vp, _vp, _fp, _i, _i, _f, _f, _f, _fp, _fp, _fp, _fp, _fp,
                                       _vp, _sz, _vp]),
    "focr_focus_loss_ws_tensor": (C.c_int, [_i, _i, C.c_char_p, C.POINTER(_ll), C.POINTER(_ll), C.POINTER(_i)]),
}


def _bind():
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # a missing symbol is a broken build: fail loudly
        fn.restype = res
        fn.argtypes = args


_bind()


def check(rc: int, what: str = "focr call") -> None:
    if rc != 0:
        raise FocrError(f"{what} failed ({rc}): {lib.focr_last_error().decode()}")


_STATUS_CHECKS = __import__("os").environ.get("FOCR_CHECK_STATUS", "0") not in ("", "0")


def set_status_checks(on: bool) -> None:
    """Opt-in debug mode: after kernels that report bad indices / lengths through a device status word (text embedding,
    CTC loss, crop resize) read the word back (one host synchronisation per call) and raise like torch does.  Off by default:
    a training step never blocks the host.  Also enabled by FOCR_CHECK_STATUS=1 in the environment."""
    global _STATUS_CHECKS
    _STATUS_CHECKS = bool(on)


def status_checks() -> bool:
    return _STATUS_CHECKS


def ptr(t) -> int:
    """Raw device pointer of a torch tensor (0 for None)."""
    return 0 if t is None else t.data_ptr()


def cur_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


_PROF_ON = False


def prof_enable(mode: int, focus: bytes = b"") -> None:
    """event scopes on the launching stream (bench.py); while on, the trainer keeps the eager (non-graph) path"""
    global _PROF_ON
    _PROF_ON = mode != 0
    lib.focr_prof_enable(mode, focus)


def prof_enabled() -> bool:
    return _PROF_ON


def prof_collect() -> dict:
    """{scope: (launches, total_ms)} recorded since the last call (synchronises the device)."""
    buf = C.create_string_buffer(1 << 16)
    lib.focr_prof_collect(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out
