"""Team mode (slab-decomposed flow step + y-band-sharded packets) under torchrun, checked against the oracle by
tests/multigpu/team_parity.py.  With fewer GPUs than ranks every rank shares cuda:0 (CUDA IPC works between processes on one
device; gloo process group, host-side team barrier), so the test also runs on the single-GPU box; with enough GPUs each rank
gets its own and the barrier is the device one (flag words over NVLink)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, nproc, port, env=None, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu", script)]
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the oracle's FFT workers are set explicitly (oracle/grid.py)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **(env or {})))


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 4])
def test_team_parity_against_the_oracle(nproc):
    import torch
    same = torch.cuda.device_count() < nproc
    r = _run("team_parity.py", nproc, 29533 + nproc, env={"SWRT_TEAM_SAME_GPU": "1" if same else "0"})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-5000:]
    assert f"team parity ok on {nproc} ranks" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [2, 3])
def test_team_parity_with_the_staged_ray_kernels_on_band_buffers(kernel):
    """The three-level tile kernel (2) and its persistent variant (3) on y-band-sharded packets: patches staged from the band +
    halo buffers (TMA where the patch is inside, cooperative fill where it wraps in x, global gathers where rows are missing)."""
    import torch
    same = torch.cuda.device_count() < 2
    r = _run("team_parity.py", 2, 29551 + kernel, env={"SWRT_TEAM_SAME_GPU": "1" if same else "0", "SWRT_TEAM_KERNEL": str(kernel)})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-5000:]
    assert "team parity ok on 2 ranks" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["push", "copy", "pull"])
def test_slab_step_variants_match_the_single_gpu_step(mode):
    """`mode` = variant of the first transpose (slab.py): peer stores from the y-pass, block-copy kernel, or x-pass pull;
    plus the NCCL all_to_all_single variant when every rank has its own GPU."""
    import torch
    same = torch.cuda.device_count() < 2
    r = _run("slab_parity.py", 2, 29541, env={"SWRT_SLAB_MODE": mode, "SWRT_TEAM_SAME_GPU": "1" if same else "0"})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-5000:]
    assert "slab parity ok" in r.stdout
