#!/usr/bin/env python
"""Build A/B variants of libflexq_b200.so that differ only in -D flags of gemm_w6ax.cu.

    python tools/build_variants.py NAME="-DFLAG=1 -DOTHER=2" NAME2="..."

writes tools/ubench/ab/lib_NAME.so (git-ignored, travels to the GPU box with gpurun); select one with
FLEXQ_B200_LIB=<path>.  The other translation units are taken from the regular build (flexq_b200/build/*.o)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flexq_b200 import build as fb  # noqa: E402

AB = os.path.join(ROOT, "tools", "ubench", "ab")


def main():
    fb.build()
    os.makedirs(AB, exist_ok=True)
    objdir = os.path.join(fb.HERE, "build")
    others = [os.path.join(objdir, s.replace(".cu", ".o")) for s in fb.SOURCES if s != "gemm_w6ax.cu"]
    procs = []
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition("=")
        obj = os.path.join(AB, f"gemm_{name}.o")
        cmd = [fb.nvcc_path(), *flags.split(), *fb.NVCC_FLAGS, "-c", os.path.join(fb.CSRC, "gemm_w6ax.cu"), "-o", obj]
        procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise SystemExit(f"variant {name} failed:\n{out}")
        so = os.path.join(AB, f"lib_{name}.so")
        subprocess.check_call([fb.nvcc_path(), "-shared", "-o", so, obj, *others, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
        os.remove(obj)
        print(so)


if __name__ == "__main__":
    main()
