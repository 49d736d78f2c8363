"""Thin tensor-level wrappers over the C ABI (one function per kernel family).

All tensors are CUDA tensors owned by the caller; outputs are passed in (the library never allocates).
"""
import ctypes

import torch

from . import _lib
from ._lib import (MOS_OUT_BF16, MOS_OUT_F32, MOS_OUT_HEADS, MOS_SEG_ROWS, MOS_SEG_TRANSPOSED, GemmArgs, check,
                   current_stream, ptr)

BN = 160
BK = 64


def _is_bf16(t):
    return t is None or t.dtype == torch.bfloat16


def gemm(A, W, out=None, *, bias=None, bias_batch=None, rows_per_batch=0, residual=None, geglu=False,
         lora_down=None, lora_up=None, lora_seg=0, conv=None, splits=1, partial=None, stages=0,
         out_f32=False, heads=None, M=None, lda=None, ldc=None, ldr=None, bias_batch_ld=0):
    """out = epilogue(A @ W^T [+ LoRA]).

    A: bf16 [M, K] (row pitch lda) or, with conv=(B, H, Wd, C), the NHWC activation [B, H, Wd, C].
    W: bf16 [N, K] (conv: [N, 9*C]).  heads: dict(seg_ptr=[...], seg_kind=[...], seg_rows_pad=[...], heads=,
    head_dim=, dpad=, dv_pad=, tokens_per_batch=) selects the head-split epilogue (Q/K rows, V transposed).
    """
    assert A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16
    a = GemmArgs()
    a.A, a.W = ptr(A), ptr(W)
    N = W.shape[0]
    if conv is not None:
        B, H, Wd, C = conv
        a.conv, a.B, a.H, a.Wd, a.C = 1, B, H, Wd, C
        a.M, a.K = B * H * Wd, C
        a.lda = C if lda is None else lda
        assert W.shape[1] == 9 * C
    else:
        a.M = A.shape[0] if M is None else M
        a.K = W.shape[1]
        a.lda = (A.stride(0) if lda is None else lda)
    a.N = N
    a.splits, a.stages = splits, stages
    a.partial = ptr(partial)
    if bias is not None:
        assert bias.dtype == torch.float32
    a.bias = ptr(bias)
    a.bias_batch = ptr(bias_batch)
    a.rows_per_batch = rows_per_batch
    a.bias_batch_ld = bias_batch_ld
    assert _is_bf16(residual)
    a.residual = ptr(residual)
    if residual is not None:
        a.ldr = residual.stride(0) if ldr is None else ldr
    a.geglu = 1 if geglu else 0
    if lora_down is not None:
        assert lora_down.dtype == torch.bfloat16 and lora_down.shape[0] == 16 and lora_up.dtype == torch.float32
        a.lora_down, a.lora_up = ptr(lora_down), ptr(lora_up)
        a.lora_seg = lora_seg or N
    if heads is not None:
        a.out_mode = MOS_OUT_HEADS
        for i, (p_, k_, r_) in enumerate(zip(heads['seg_ptr'], heads['seg_kind'], heads['seg_rows_pad'])):
            a.seg_ptr[i] = p_.data_ptr()
            a.seg_kind[i] = k_
            a.seg_rows_pad[i] = r_
        a.heads, a.head_dim = heads['heads'], heads['head_dim']
        a.dpad, a.dv_pad = heads['dpad'], heads['dv_pad']
        a.tokens_per_batch = heads['tokens_per_batch']
    else:
        a.out_mode = MOS_OUT_F32 if out_f32 else MOS_OUT_BF16
        a.out = ptr(out)
        if out is not None:
            a.ldc = out.stride(0) if ldc is None else ldc
    check(_lib.lib().mos_gemm_bf16(ctypes.byref(a), current_stream()), 'mos_gemm_bf16')
    return out


def splitk_finalize(partial, splits, M, N, out, *, bias=None, bias_batch=None, rows_per_batch=0, residual=None,
                    ldc=None, ldr=None):
    check(_lib.lib().mos_splitk_finalize(
        ptr(partial), ctypes.c_int32(splits), ctypes.c_int64(M), ctypes.c_int64(N), ptr(bias), ptr(bias_batch),
        ctypes.c_int64(rows_per_batch), ptr(residual),
        ctypes.c_int64((residual.stride(0) if ldr is None else ldr) if residual is not None else 0), ptr(out),
        ctypes.c_int64(out.stride(0) if ldc is None else ldc), current_stream()), 'mos_splitk_finalize')
    return out
