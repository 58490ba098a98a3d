"""Launches tests/multigpu_check.py on 2 GPUs when the box has them (the driver's `-m gpu` run is 1 GPU:
then this is skipped; it is run under `gpurun --gpus 2` during development, log in profiles/)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "multigpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "native_exchange=True" in r.stdout and "native_exchange=False" in r.stdout
    assert "full-rank family (pull-protocol exchange)" in r.stdout
