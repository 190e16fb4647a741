import os

import numpy as np
import torch


def test_torch_cpu_baseline_matches_reference_golden(golden_dir):
    """the CPU-baseline port is pinned to the reference's QuantLinear outputs"""
    from oracle.fakequant_torch import FakeQuantLinearCPU, fake_quant_sym
    g = np.load(os.path.join(golden_dir, "linear_golden.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        x, w = torch.from_numpy(g[n + "/x"]), torch.from_numpy(g[n + "/w"])
        ab = int(g[n + "/abits"])
        assert torch.equal(fake_quant_sym(w, 6), torch.from_numpy(g[n + "/wdeq"]))
        for faithful in (True, False):
            y = FakeQuantLinearCPU(w, ab, faithful=faithful)(x)
            ref = torch.from_numpy(g[n + "/y"])
            assert torch.equal(y, ref), n
