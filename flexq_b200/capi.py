"""ctypes binding of libflexq_b200.so (the C ABI declared in include/flexq_b200.h).

PyTorch is plumbing only: tensors give device memory (``data_ptr()``) and the current CUDA
stream.  There is no CPU fallback -- if the library is missing or a call fails this module
raises.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLEXQ_B200_LIB") or os.path.join(_HERE, "libflexq_b200.so")

GROUP = 128
ROUND_CUDA = 0
ROUND_PYTHON = 1

_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/flexq_b200.h one to one
_SIGNATURES = {
    "flexq_version": (_i, []),
    "flexq_status_string": (ctypes.c_char_p, [_i]),
    "flexq_w6_packed_bytes": (_sz, [_i, _i]),
    "flexq_planes_bytes": (_sz, [_i, _i, _i]),
    "flexq_sx_ld": (_i, [_i]),
    "flexq_xscale_ref_halves": (_sz, [_i, _i]),
    "flexq_gemm_workspace_bytes": (_sz, []),
    "flexq_linear_workspace_bytes": (_sz, [_i, _i]),
    "flexq_workspace_init": (_i, [_vp, _sz, _vp]),
    "flexq_bit_packing_i32": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "flexq_bit_packing_f16": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "flexq_quant_act": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "flexq_quant_act_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "flexq_pack_w6_i32": (_i, [_vp, _vp, _i, _i, _vp]),
    "flexq_pack_w6_i8": (_i, [_vp, _vp, _i, _i, _vp]),
    "flexq_quant_pack_w6_f16": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "flexq_quant_pack_w6_f32": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "flexq_planes_to_i8": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "flexq_planes_to_w6": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "flexq_xscale_ref_to_sx": (_i, [_vp, _vp, _i, _i, _vp]),
    "flexq_w6_to_i8": (_i, [_vp, _vp, _i, _i, _vp]),
    "flexq_gemm_w6ax": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "flexq_gemm_w6ax_silu_mul": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "flexq_gemm_w6ax_groupsums": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "flexq_debug_schedule": (_i, [_i, _i, _i, _i, _i, ctypes.POINTER(ctypes.c_int), _i, ctypes.POINTER(ctypes.c_int)]),
    "flexq_debug_gemm_trace": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "flexq_linear_w6ax_f16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "flexq_gemm_ref_layout": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "flexq_rmsnorm_quant_f16": (_i, [_vp, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "flexq_silu_mul_quant_f16": (_i, [_vp, _vp, ctypes.c_longlong, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "flexq_allreduce_sum_synced_f16": (_i, [_vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), _sz, _sz, _i, _i, _vp]),
    "flexq_allreduce_oneshot_f16": (_i, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), _sz, _i, _i, _vp, _vp]),
    "flexq_set_sm_limit": (_i, [_i]),
    "flexq_set_allreduce_blocks": (_i, [_i]),
    "flexq_allreduce_sum_f16": (_i, [_vp, ctypes.POINTER(ctypes.c_void_p), _sz, _sz, _i, _i, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class FlexQError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FlexQError(
                f"{LIB_PATH} not found: build it with `python -m flexq_b200.build` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status: int, what: str):
    if status != 0:
        msg = load().flexq_status_string(status).decode()
        raise FlexQError(f"{what} failed: status {status} ({msg})")


def _ptr(t: torch.Tensor):
    assert t.is_cuda and t.is_contiguous(), "flexq_b200 needs contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_device(fn):
    """Run a wrapper with the CUDA device of its tensor arguments current (so that the launch, the stream taken by
    _stream() and the outputs allocated inside all belong to that device) and check that the operands share it.
    A model spread over several GPUs in one process (device_map='auto', as the reference's evaluation flow loads
    large models) calls into the library from whatever device happens to be current."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise FlexQError(f"{fn.__name__}: operands on different devices ({dev} and {a.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def ceil4(m: int) -> int:
    return (m + 3) // 4 * 4


# ------------------------------------------------------------------------------------------
# thin tensor-level wrappers (allocate outputs with torch, call the C ABI on the current stream)
# ------------------------------------------------------------------------------------------
_ws_pool: dict = {}


def stream_workspace(M: int | None = None, K: int | None = None, device=None) -> torch.Tensor:
    """The workspace shared by every flexq_b200 call issued on the current stream of `device` (calls on one stream run
    one after the other, so they can share the GEMM's partial-sum scratch and the quantised-activation staging area; a
    workspace must not be used by GEMMs running concurrently on different streams).  Grown on demand, zeroed once."""
    lib = load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    need = lib.flexq_gemm_workspace_bytes() if M is None else lib.flexq_linear_workspace_bytes(M, K)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_pool.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=dev)
        _ws_pool[key] = ws
    return ws


def new_workspace(M: int | None = None, K: int | None = None, device="cuda") -> torch.Tensor:
    lib = load()
    n = lib.flexq_gemm_workspace_bytes() if M is None else lib.flexq_linear_workspace_bytes(M, K)
    return torch.zeros(n, dtype=torch.uint8, device=device)


@_on_device
def bit_packing_i32(ints: torch.Tensor, bits: int) -> torch.Tensor:
    R, K = ints.shape
    out = torch.empty(R * K * bits // 32, dtype=torch.int32, device=ints.device)
    check(load().flexq_bit_packing_i32(_ptr(ints), _ptr(out), R, K, bits, _stream()), "flexq_bit_packing_i32")
    return out


@_on_device
def bit_packing_f16(x: torch.Tensor, bits: int):
    M, K = x.shape
    planes = torch.empty(M * K * bits // 32, dtype=torch.int32, device=x.device)
    xs = torch.empty(K // GROUP, 2 * ceil4(M), dtype=torch.float16, device=x.device)
    check(load().flexq_bit_packing_f16(_ptr(x), _ptr(planes), _ptr(xs), M, K, bits, _stream()), "flexq_bit_packing_f16")
    return planes, xs


@_on_device
def quant_act(x: torch.Tensor, bits: int, mode: int = ROUND_CUDA):
    """fp16 activations -> int8 containers + fp32 scales (mode selects the reference behaviour); fp32 activations
    always take the python quantiser's fp32 arithmetic (flexq_quant_act_f32)."""
    M, K = x.shape
    xq = torch.empty(M, K, dtype=torch.int8, device=x.device)
    sx = torch.empty(K // GROUP, ceil4(M), dtype=torch.float32, device=x.device)
    if x.dtype == torch.float32:
        check(load().flexq_quant_act_f32(_ptr(x), _ptr(xq), _ptr(sx), M, K, bits, _stream()), "flexq_quant_act_f32")
        return xq, sx
    assert x.dtype == torch.float16, "flexq_quant_act takes fp16 or fp32 activations"
    check(load().flexq_quant_act(_ptr(x), _ptr(xq), _ptr(sx), M, K, bits, mode, _stream()), "flexq_quant_act")
    return xq, sx


@_on_device
def rmsnorm_quant(x: torch.Tensor, gamma: torch.Tensor, eps: float, bits: int, residual: torch.Tensor | None = None,
                  want_normed: bool = False):
    """Fused (residual add +) RMSNorm + activation quantise.  `residual` is updated in place to x + residual.
    Returns (xq, sx, normed or None)."""
    M, K = x.shape
    xq = torch.empty(M, K, dtype=torch.int8, device=x.device)
    sx = torch.empty(K // GROUP, ceil4(M), dtype=torch.float32, device=x.device)
    normed = torch.empty_like(x) if want_normed else None
    check(load().flexq_rmsnorm_quant_f16(_ptr(x), _ptr(residual) if residual is not None else None, _ptr(gamma), float(eps),
                                         _ptr(normed) if normed is not None else None, _ptr(xq), _ptr(sx), M, K, bits, _stream()),
          "flexq_rmsnorm_quant_f16")
    return xq, sx, normed


@_on_device
def silu_mul_quant(gate: torch.Tensor, up: torch.Tensor, bits: int = 8, want_out: bool = False):
    """Fused SiLU(gate) * up + activation quantise.  gate/up: [M, K] fp16 views with equal row stride
    (e.g. the two halves of a fused gate_up output).  Returns (xq, sx, out or None)."""
    M, K = gate.shape
    if gate.stride(1) != 1 or up.stride(1) != 1 or gate.stride(0) != up.stride(0):
        raise FlexQError("silu_mul_quant: gate/up must be row-major with the same row stride")
    xq = torch.empty(M, K, dtype=torch.int8, device=gate.device)
    sx = torch.empty(K // GROUP, ceil4(M), dtype=torch.float32, device=gate.device)
    out = torch.empty(M, K, dtype=torch.float16, device=gate.device) if want_out else None
    gp, up_ = ctypes.c_void_p(gate.data_ptr()), ctypes.c_void_p(up.data_ptr())      # strided views: pass raw pointers
    assert gate.is_cuda and up.is_cuda and gate.dtype == torch.float16 and up.dtype == torch.float16
    check(load().flexq_silu_mul_quant_f16(gp, up_, gate.stride(0), _ptr(out) if out is not None else None,
                                          _ptr(xq), _ptr(sx), M, K, bits, _stream()), "flexq_silu_mul_quant_f16")
    return xq, sx, out


def _new_w6(N, K, device):
    return torch.empty(load().flexq_w6_packed_bytes(N, K), dtype=torch.uint8, device=device)


@_on_device
def pack_w6(w_int: torch.Tensor) -> torch.Tensor:
    N, K = w_int.shape
    w6 = _new_w6(N, K, w_int.device)
    fn = {torch.int32: "flexq_pack_w6_i32", torch.int8: "flexq_pack_w6_i8"}[w_int.dtype]
    check(getattr(load(), fn)(_ptr(w_int), _ptr(w6), N, K, _stream()), fn)
    return w6


@_on_device
def quant_pack_w6(w: torch.Tensor):
    N, K = w.shape
    w6 = _new_w6(N, K, w.device)
    ws = torch.empty(K // GROUP, N, dtype=torch.float16, device=w.device)
    fn = {torch.float16: "flexq_quant_pack_w6_f16", torch.float32: "flexq_quant_pack_w6_f32"}[w.dtype]
    check(getattr(load(), fn)(_ptr(w), _ptr(w6), _ptr(ws), N, K, _stream()), fn)
    return w6, ws


@_on_device
def planes_to_i8(planes: torch.Tensor, R: int, K: int, bits: int) -> torch.Tensor:
    out = torch.empty(R, K, dtype=torch.int8, device=planes.device)
    check(load().flexq_planes_to_i8(_ptr(planes), _ptr(out), R, K, bits, _stream()), "flexq_planes_to_i8")
    return out


@_on_device
def planes_to_w6(planes: torch.Tensor, N: int, K: int) -> torch.Tensor:
    w6 = _new_w6(N, K, planes.device)
    scratch = torch.empty(N, K, dtype=torch.int8, device=planes.device)
    check(load().flexq_planes_to_w6(_ptr(planes), _ptr(w6), _ptr(scratch), N, K, _stream()), "flexq_planes_to_w6")
    return w6


@_on_device
def xscale_ref_to_sx(xs: torch.Tensor, M: int, K: int) -> torch.Tensor:
    sx = torch.empty(K // GROUP, ceil4(M), dtype=torch.float32, device=xs.device)
    check(load().flexq_xscale_ref_to_sx(_ptr(xs), _ptr(sx), M, K, _stream()), "flexq_xscale_ref_to_sx")
    return sx


@_on_device
def w6_to_i8(w6: torch.Tensor, N: int, K: int) -> torch.Tensor:
    out = torch.empty(N, K, dtype=torch.int8, device=w6.device)
    check(load().flexq_w6_to_i8(_ptr(w6), _ptr(out), N, K, _stream()), "flexq_w6_to_i8")
    return out


@_on_device
def gemm_w6ax(xq, sx, w6, w_scale, N: int, workspace: torch.Tensor, out: torch.Tensor | None = None):
    M, K = xq.shape
    if out is None:
        out = torch.empty(M, N, dtype=torch.float16, device=xq.device)
    check(load().flexq_gemm_w6ax(_ptr(xq), _ptr(sx), _ptr(w6), _ptr(w_scale), _ptr(out), M, N, K,
                                 _ptr(workspace), workspace.numel(), _stream()), "flexq_gemm_w6ax")
    return out


@_on_device
def gemm_w6ax_silu_mul(xq, sx, w6_gate_up, w_scale_gate_up, inter: int, workspace: torch.Tensor, out: torch.Tensor | None = None):
    """gate_up GEMM with SiLU(gate) * up in the epilogue: [M, inter] fp16 from weights packed in the interleaved row
    order of ``model_pack.interleave_gate_up`` (include/flexq_b200.h: flexq_gemm_w6ax_silu_mul)."""
    M, K = xq.shape
    if out is None:
        out = torch.empty(M, inter, dtype=torch.float16, device=xq.device)
    check(load().flexq_gemm_w6ax_silu_mul(_ptr(xq), _ptr(sx), _ptr(w6_gate_up), _ptr(w_scale_gate_up), _ptr(out), M, inter, K,
                                          _ptr(workspace), workspace.numel(), _stream()), "flexq_gemm_w6ax_silu_mul")
    return out


@_on_device
def gemm_w6ax_groupsums(xq, w6, N: int) -> torch.Tensor:
    M, K = xq.shape
    S = torch.empty(M, N, K // GROUP, dtype=torch.int32, device=xq.device)
    check(load().flexq_gemm_w6ax_groupsums(_ptr(xq), _ptr(w6), _ptr(S), M, N, K, _stream()), "flexq_gemm_w6ax_groupsums")
    return S


@_on_device
def linear_w6ax(x, w6, w_scale, N: int, x_bits: int, workspace: torch.Tensor, mode: int = ROUND_CUDA,
                out: torch.Tensor | None = None):
    M, K = x.shape
    if out is None:
        out = torch.empty(M, N, dtype=torch.float16, device=x.device)
    check(load().flexq_linear_w6ax_f16(_ptr(x), _ptr(w6), _ptr(w_scale), _ptr(out), M, N, K, x_bits, mode,
                                       _ptr(workspace), workspace.numel(), _stream()), "flexq_linear_w6ax_f16")
    return out


@_on_device
def gemm_ref_layout(x_planes, x_scale, w6, w_scale, M: int, N: int, K: int, x_bits: int, workspace: torch.Tensor):
    out = torch.empty(M, N, dtype=torch.float16, device=w6.device)
    check(load().flexq_gemm_ref_layout(_ptr(x_planes), _ptr(x_scale), _ptr(w6), _ptr(w_scale), _ptr(out), M, N, K, x_bits,
                                       _ptr(workspace), workspace.numel(), _stream()), "flexq_gemm_ref_layout")
    return out


def allreduce_sum_f16(multicast_ptr: int, peer_ptrs, offset_elems: int, elems: int, rank: int, world: int):
    """In-place sum of a symmetric-memory fp16 buffer over `world` ranks (this rank's slice; see the header).
    `multicast_ptr` 0/None selects the peer-pointer path."""
    arr = (ctypes.c_void_p * 8)(*([int(p) for p in peer_ptrs] + [0] * (8 - len(peer_ptrs)))) if peer_ptrs else None
    check(load().flexq_allreduce_sum_f16(ctypes.c_void_p(int(multicast_ptr or 0)), arr, offset_elems, elems, rank, world, _stream()),
          "flexq_allreduce_sum_f16")


def allreduce_sum_synced_f16(multicast_ptr: int, peer_ptrs, flag_ptrs, offset_elems: int, elems: int, rank: int, world: int):
    """Two-shot all-reduce with the cross-rank hand-shake inside the kernel (no stream barriers)."""
    arr = (ctypes.c_void_p * 8)(*([int(p) for p in peer_ptrs] + [0] * (8 - len(peer_ptrs))))
    f = (ctypes.c_void_p * 8)(*([int(p) for p in flag_ptrs] + [0] * (8 - len(flag_ptrs))))
    check(load().flexq_allreduce_sum_synced_f16(ctypes.c_void_p(int(multicast_ptr or 0)), arr, f, offset_elems, elems, rank, world,
                                                _stream()), "flexq_allreduce_sum_synced_f16")


@_on_device
def allreduce_oneshot_f16(data_ptrs, flag_ptrs, elems: int, rank: int, world: int, out: torch.Tensor):
    """One-kernel decode all-reduce over symmetric memory (see the header)."""
    d = (ctypes.c_void_p * 8)(*([int(p) for p in data_ptrs] + [0] * (8 - len(data_ptrs))))
    f = (ctypes.c_void_p * 8)(*([int(p) for p in flag_ptrs] + [0] * (8 - len(flag_ptrs))))
    check(load().flexq_allreduce_oneshot_f16(d, f, elems, rank, world, _ptr(out), _stream()), "flexq_allreduce_oneshot_f16")
