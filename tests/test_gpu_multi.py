"""Multi-GPU parity of the tensor-parallel path (needs >= 2 GPUs on the box; skipped otherwise).
One process per GPU under torchrun, NCCL for the plumbing; checks run inside tools-free worker code below."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    import torch, torch.distributed as dist
    sys.path.insert(0, os.environ["FLEXQ_ROOT"])
    from flexq_b200 import capi, tp
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    capi.load()
    torch.manual_seed(7)                                   # same full weight / activations on every rank
    M, N, K = 200, 1024, 2048
    w = (0.05 * torch.randn(N, K, device=dev)).half()
    x = torch.randn(M, K, device=dev).half()
    full = tp.TPLinearW6Ax(w, "replicated", 6).forward(x)

    # column parallel: shards concatenate to the single-GPU result exactly (same integers, same fp32 order per column)
    col = tp.TPLinearW6Ax(w, "column", 6, rank, world).forward(x)
    parts = [torch.empty_like(col) for _ in range(world)]
    dist.all_gather(parts, col)
    assert torch.allclose(torch.cat(parts, 1).float(), full.float(), rtol=2e-3, atol=2e-3), "column"

    # row parallel: NCCL, peer two-shot (both paths, with and without chunk overlap), one-shot -- all equal up to fp16 rounding
    xs = tp.shard_activation(x, rank, world)
    y_nccl = tp.TPLinearW6Ax(w, "row", 6, rank, world).forward(xs)
    assert torch.allclose(y_nccl.float(), full.float(), rtol=5e-3, atol=5e-2), "row nccl"
    for mc in (False, True):
        for chunks, reserve in ((1, 0), (2, 8)):
            lin = tp.TPLinearW6Ax(w, "row", 6, rank, world).enable_peer_allreduce(M, chunks=chunks, use_multicast=mc, sm_reserve=reserve)
            for _ in range(3):
                y = lin.forward(xs)
            torch.cuda.synchronize()
            assert torch.allclose(y.float(), y_nccl.float(), rtol=5e-3, atol=5e-2), ("row peer", mc, chunks)
    lin = tp.TPLinearW6Ax(w, "row", 6, rank, world).enable_oneshot_allreduce(16)
    xd = xs[:16].contiguous()
    ref16 = tp.TPLinearW6Ax(w, "row", 6, rank, world).forward(xd)
    for _ in range(4):
        y = lin.forward(xd)
    torch.cuda.synchronize()
    assert torch.allclose(y.float(), ref16.float(), rtol=5e-3, atol=5e-2), "row oneshot"
    g = torch.cuda.CUDAGraph()                              # the one-shot reduction is graph capturable
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        lin.forward(xd)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize(); dist.barrier()
    with torch.cuda.graph(g):
        yg = lin.forward(xd)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.allclose(yg.float(), ref16.float(), rtol=5e-3, atol=5e-2), "row oneshot graph"
    dist.barrier()
    if rank == 0:
        print("TP_MULTI_OK")
    dist.destroy_process_group()
''')


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_tensor_parallel_paths_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, FLEXQ_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "TP_MULTI_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
