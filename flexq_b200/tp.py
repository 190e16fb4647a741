"""Tensor-parallel W6Ax linears: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch on the B200 box, gloo in CPU tests) for the plumbing.

Follows the Megatron pairing the reference's FasterTransformer fork uses
(/root/reference/e2e/src/fastertransformer/layers/TensorParallelSiluFfnLayer.cc:41-62,83-95):

* column parallel (qkv, gate, up): W[N/tp, K], w_scale[G, N/tp]; outputs stay sharded, no
  communication.
* row parallel (o_proj, down): W[N, K/tp] split on 128-group boundaries so every rank owns
  whole groups with their own scales; the input arrives already K-sharded from the preceding
  column-parallel layer; fp16 partial outputs are summed with one all-reduce
  (ftNcclAllReduceSum, e2e/src/fastertransformer/utils/nccl_utils.cc:56-68).

INT32 group sums stay bit-exact per rank; only the fp16 summation order differs from 1 GPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import capi

GROUP = capi.GROUP


def shard_weight(w: torch.Tensor, mode: str, rank: int, world: int) -> torch.Tensor:
    """Slice an [N, K] weight for this rank."""
    N, K = w.shape
    if mode == "column":
        if N % world:
            raise ValueError(f"N={N} not divisible by tp={world}")
        n = N // world
        return w[rank * n:(rank + 1) * n].contiguous()
    if mode == "row":
        if K % GROUP or (K // GROUP) % world:
            raise ValueError(f"K={K}: K/128 groups must be divisible by tp={world}")
        k = K // world
        return w[:, rank * k:(rank + 1) * k].contiguous()
    raise ValueError(f"unknown tensor-parallel mode {mode!r}")


def shard_activation(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """K-shard of an activation for a row-parallel layer (what the preceding column-parallel
    layer would have produced on this rank)."""
    K = x.shape[-1]
    if K % GROUP or (K // GROUP) % world:
        raise ValueError(f"K={K}: K/128 groups must be divisible by tp={world}")
    k = K // world
    return x[..., rank * k:(rank + 1) * k].contiguous()


def all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class TPLinearW6Ax:
    """A packed W6Ax linear shard living on this rank's GPU.

    ``mode`` is "column", "row" or "replicated".  ``forward`` takes this rank's fp16 activations
    ([M, K] for column/replicated, [M, K/tp] for row) and returns [M, N/tp] (column) or the
    all-reduced [M, N] (row).
    """

    def __init__(self, w_full: torch.Tensor, mode: str, x_bits: int, rank: int = 0, world: int = 1,
                 act_round: int = capi.ROUND_CUDA, group=None):
        self.mode, self.x_bits, self.rank, self.world = mode, x_bits, rank, world
        self.act_round, self.group = act_round, group
        w = w_full if mode == "replicated" or world == 1 else shard_weight(w_full, mode, rank, world)
        self.N, self.K = w.shape
        self.w6, self.w_scale = capi.quant_pack_w6(w.cuda().contiguous())
        self._ws = None

    @classmethod
    def from_packed(cls, w6, w_scale, N, K, mode, x_bits, rank=0, world=1, act_round=capi.ROUND_CUDA, group=None):
        self = cls.__new__(cls)
        self.mode, self.x_bits, self.rank, self.world = mode, x_bits, rank, world
        self.act_round, self.group = act_round, group
        self.N, self.K, self.w6, self.w_scale, self._ws = N, K, w6, w_scale, None
        return self

    def workspace(self, M: int) -> torch.Tensor:
        need = capi.load().flexq_linear_workspace_bytes(M, self.K)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.zeros(need, dtype=torch.uint8, device=self.w6.device)
        return self._ws

    def forward(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        M = x.shape[0]
        y = capi.linear_w6ax(x, self.w6, self.w_scale, self.N, self.x_bits, self.workspace(M), self.act_round, out)
        if self.mode == "row" and self.world > 1:
            all_reduce_sum(y, self.group)
        return y

    __call__ = forward
