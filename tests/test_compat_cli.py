"""The reference's CLI harnesses and C++ entry points, rebuilt on the sm_100a library
(compat/): same argv, same stdout contract."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "compat", "bin")


def _run(name, *args, env=None):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (make -C compat)")
    e = dict(os.environ, FLEXQ_SEED="1")
    e.update(env or {})
    return subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=600, env=e)


def test_cli_usage_messages_without_gpu():
    """argument handling mirrors the reference (exit -1 + usage line); needs no GPU"""
    r = _run("test_bgemm_kernel")
    assert r.returncode != 0 and "Usage: ./test_bgemm_kernel M N K X_BITS W_BITS" in r.stdout
    r = _run("test_packing_kernel")
    assert r.returncode != 0 and "Usage: ./test_packing_kernel M K X_BITS" in r.stdout
    r = _run("test_packing_kernel", 4, 100, 6)
    assert "k must >= 128 and k % 128 == 0" in r.stdout
    r = _run("test_cublas_kernel")
    assert "Usage:" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("m,k,bits", [(1, 4096, 6), (8, 4096, 6), (8, 11008 // 128 * 128, 8), (64, 1024, 6)])
def test_packing_cli(m, k, bits):
    r = _run("test_packing_kernel", m, k, bits)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FlexQ bit packing kernel SUCCESS! consistent results!" in r.stdout
    assert re.search(r"ABQ packing [\d.]+ \(us\) exec", r.stdout) and re.search(r"FlexQ bit packing [\d.]+ \(us\) exec", r.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,k,xb", [(1, 4096, 4096, 6), (4, 1024, 4096, 8), (8, 4096, 4096, 6), (8, 4096, 11008 // 128 * 128, 8)])
def test_bgemm_cli(m, n, k, xb):
    r = _run("test_bgemm_kernel", m, n, k, xb, 6)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == ("test_w6a6_kernel" if xb == 6 else "test_w6a8_kernel")
    assert re.search(r"packing [\d.]+ \(us\) exec [\d.]+ \(us\) [\d.]+ TOPS \| [\d.]+ B-TOPS \| PASSED", r.stdout)
    assert "The best kernel config is" in r.stdout
    assert lines[-1] == "SUCCESS! consistent results!"


@pytest.mark.gpu
def test_cublas_cli():
    r = _run("test_cublas_kernel", 16, 4096, 4096)
    assert r.returncode == 0
    assert re.match(r"cuBLAS-W8A8-GEMM\. m:\s+16, n:\s+4096, k:\s+4096,\t Time: [\d.]+ ms, TFLOPS: [\d.]+", r.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("m,xb", [(1, 6), (8, 8), (16, 6)])
def test_ft_wrapper(m, xb):
    r = _run("test_ft_wrapper", m, 4096, 4096, xb)
    assert r.returncode == 0 and "FT wrapper SUCCESS" in r.stdout, r.stdout + r.stderr


def test_sweep_scripts_enumerate_the_reference_shapes(tmp_path):
    """compat/test_flexq_kernel.sh / test_cublas_kernel.sh (role of the reference's engine/test_*_kernel.sh): with the
    harness binaries replaced by stubs, they must produce one result file per (model, batch, layer) with the reference's
    file names and argv (M N K X_BITS W_BITS); needs no GPU."""
    import shutil
    work = tmp_path / "compat"
    (work / "bin").mkdir(parents=True)
    for s in ("test_flexq_kernel.sh", "test_cublas_kernel.sh"):
        shutil.copy(os.path.join(ROOT, "compat", s), work / s)
    for exe in ("test_bgemm_kernel", "test_cublas_kernel"):
        p = work / "bin" / exe
        p.write_text("#!/bin/sh\necho \"$@\"\n")
        p.chmod(0o755)
    env = dict(os.environ, BS="1 8")
    for s in ("test_flexq_kernel.sh", "test_cublas_kernel.sh"):
        subprocess.run(["bash", str(work / s)], check=True, env=env, timeout=120)
    got = sorted(os.listdir(work / "flexq_results"))
    assert len(got) == 2 * 5 * 4 and got == sorted(os.listdir(work / "cublas_results"))
    assert "llama_2_70b_8x8192x28672_w6a8.txt" in got and "llama_7b_1x12288x4096_w6a6.txt" in got and "opt_30b_1x21504x7168_w6a6.txt" in got
    assert (work / "flexq_results" / "llama_2_70b_8x8192x28672_w6a8.txt").read_text().split() == ["8", "8192", "28672", "8", "6"]
    assert (work / "cublas_results" / "llama_2_13b_1x13824x5120_w6a6.txt").read_text().split() == ["1", "13824", "5120"]
