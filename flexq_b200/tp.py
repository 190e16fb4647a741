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


class PeerAllReduce:
    """fp16 sum all-reduce in a symmetric-memory buffer, done by flexq_allreduce_sum_f16 over NVLink /
    NVSwitch peer memory (multimem in-switch reduction when the allocation has a multicast mapping).

    torch.distributed._symmetric_memory is the plumbing only: it allocates the buffer, exchanges the
    handles and provides the cross-rank stream barrier; the reduction itself is our kernel.
    """

    def __init__(self, max_elems: int, device: torch.device, group=None, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.buf = symm.empty(max_elems, dtype=torch.float16, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.multicast_ptr = mc if use_multicast else 0
        # flag words for the in-kernel hand-shake (zeroed everywhere before anyone uses them)
        self.flags = symm.empty(32, dtype=torch.int32, device=device)
        self.flags.zero_()
        self.h_flags = symm.rendezvous(self.flags, self.group)
        self.flag_ptrs = [int(p) for p in self.h_flags.buffer_ptrs]
        torch.cuda.synchronize(device)
        self.h_flags.barrier(channel=0)
        torch.cuda.synchronize(device)
        self.in_kernel_sync = True

    def view(self, M: int, N: int) -> torch.Tensor:
        return self.buf[: M * N].view(M, N)

    def reduce_(self, offset_elems: int, elems: int):
        """All ranks call this on their current stream after writing buf[offset : offset+elems]."""
        if self.in_kernel_sync:                           # one kernel: arrive / reduce + publish / done
            capi.allreduce_sum_synced_f16(self.multicast_ptr, self.peer_ptrs, self.flag_ptrs, offset_elems, elems, self.rank, self.world)
            return
        self.hdl.barrier(channel=0)                       # every rank's partials are written
        capi.allreduce_sum_f16(self.multicast_ptr, self.peer_ptrs, offset_elems, elems, self.rank, self.world)
        self.hdl.barrier(channel=0)                       # every slice is published


class PeerOneShotAllReduce:
    """Decode-sized fp16 sum all-reduce in ONE kernel (flexq_allreduce_oneshot_f16): the partial lives in a
    symmetric allocation, arrival / completion are signalled through per-peer flag words, the epoch is a
    device-side counter (CUDA-graph capturable, no stream barriers).  `partial(M, N)` is where the producer
    (the row-parallel GEMM) writes; `reduce(M, N)` sums into a private tensor."""

    def __init__(self, max_elems: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.max_elems = (max_elems + 7) // 8 * 8
        self.data = symm.empty(self.max_elems, dtype=torch.float16, device=device)
        self.flags = symm.empty(32, dtype=torch.int32, device=device)
        self.flags.zero_()
        self.h_data = symm.rendezvous(self.data, self.group)
        self.h_flags = symm.rendezvous(self.flags, self.group)
        self.data_ptrs = [int(p) for p in self.h_data.buffer_ptrs]
        self.flag_ptrs = [int(p) for p in self.h_flags.buffer_ptrs]
        torch.cuda.synchronize(device)
        self.h_flags.barrier(channel=0)                    # every rank's flags are zero before anyone publishes
        torch.cuda.synchronize(device)

    def partial(self, M: int, N: int) -> torch.Tensor:
        return self.data[: M * N].view(M, N)

    def reduce(self, M: int, N: int, out: torch.Tensor | None = None) -> torch.Tensor:
        if out is None:
            out = torch.empty(M, N, dtype=torch.float16, device=self.data.device)
        capi.allreduce_oneshot_f16(self.data_ptrs, self.flag_ptrs, M * N, self.rank, self.world, out)
        return out


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
        return capi.stream_workspace(M, self.K, self.w6.device)       # one per (device, stream), shared by all layers

    def forward(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        M = x.shape[0]
        os_ = getattr(self, "_oneshot", None)
        if self.mode == "row" and self.world > 1 and os_ is not None and M * self.N <= os_.max_elems and (M * self.N) % 8 == 0:
            part = os_.partial(M, self.N)
            capi.linear_w6ax(x, self.w6, self.w_scale, self.N, self.x_bits, self.workspace(M), self.act_round, part)
            return os_.reduce(M, self.N, out)
        if self.mode == "row" and self.world > 1 and getattr(self, "_ar", None) is not None:
            y = self.forward_peer(x)
            if out is not None:
                out.copy_(y)
                return out
            return y
        y = capi.linear_w6ax(x, self.w6, self.w_scale, self.N, self.x_bits, self.workspace(M), self.act_round, out)
        if self.mode == "row" and self.world > 1:
            all_reduce_sum(y, self.group)
        return y

    def enable_oneshot_allreduce(self, max_tokens: int = 64):
        """Decode: reduce row-parallel partials of up to `max_tokens` rows with the one-kernel peer all-reduce.
        Collective: every rank must call it."""
        if self.mode == "row" and self.world > 1:
            self._oneshot = PeerOneShotAllReduce(max_tokens * self.N, self.w6.device, self.group)
        return self

    # ---- row-parallel GEMM overlapped with the peer-memory all-reduce ----
    def enable_peer_allreduce(self, max_tokens: int, chunks: int = 3, use_multicast: bool = True, sm_reserve: int = 0,
                              first_frac: float = 0.0):
        """Row-parallel shards: write the partial outputs into symmetric memory and reduce them with our
        own kernel, token-tile chunk by chunk on a second stream so the reduction of chunk i overlaps the
        GEMM of chunk i+1.  Collective: every rank must call it."""
        if self.mode != "row" or self.world == 1:
            return self
        self._ar = PeerAllReduce(max_tokens * self.N, self.w6.device, self.group, use_multicast)
        self._ar_chunks = max(1, chunks)
        self._sm_reserve = sm_reserve          # SMs the chunk GEMMs leave free for the concurrent reduction
        self._first_frac = first_frac          # 2 chunks: share of the token tiles in the first one (0 = even split)
        self._comm = torch.cuda.Stream(device=self.w6.device)
        self._ev = [torch.cuda.Event() for _ in range(self._ar_chunks)]
        return self

    def forward_peer(self, x: torch.Tensor) -> torch.Tensor:
        M = x.shape[0]
        y = self._ar.view(M, self.N)
        ws = self.workspace(M)
        tile = 192                                          # the GEMM's token tile: chunk on tile boundaries
        tiles = (M + tile - 1) // tile
        nch = min(self._ar_chunks, tiles)
        cur = torch.cuda.current_stream()
        self._comm.wait_stream(cur)                         # the buffer may still be read by earlier work
        lib = capi.load()
        if self._sm_reserve and nch > 1:
            lib.flexq_set_sm_limit(torch.cuda.get_device_properties(self.w6.device).multi_processor_count - self._sm_reserve)
            lib.flexq_set_allreduce_blocks(self._sm_reserve)
        row = 0
        for c in range(nch):
            if nch == 2 and self._first_frac > 0:
                t0 = max(1, min(tiles - 1, round(tiles * self._first_frac)))
                rows = min(M - row, (t0 if c == 0 else tiles - t0) * tile)
            else:
                rows = min(M - row, ((tiles * (c + 1)) // nch - (tiles * c) // nch) * tile)
            capi.linear_w6ax(x[row:row + rows], self.w6, self.w_scale, self.N, self.x_bits, ws, self.act_round, y[row:row + rows])
            self._ev[c].record(cur)
            with torch.cuda.stream(self._comm):
                self._comm.wait_event(self._ev[c])
                self._ar.reduce_(row * self.N, rows * self.N)
            row += rows
        if self._sm_reserve and nch > 1:
            lib.flexq_set_sm_limit(0)
            lib.flexq_set_allreduce_blocks(0)
        cur.wait_stream(self._comm)
        return y

    __call__ = forward
