"""Runs the slab-decomposed flow step on every visible GPU (>= 2) under torchrun and checks it against the single-GPU step."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["push", "copy", "pull"])
def test_slab_step_matches_single_gpu(mode):
    """`mode` = variant of the first transpose (slab.py): peer stores from the y-pass, block-copy kernel, or x-pass pull."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (slab decomposition)")
    p = 1 << (n.bit_length() - 1)        # largest power of two
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={p}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu", "slab_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, SWRT_SLAB_MODE=mode))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "slab parity ok" in r.stdout
