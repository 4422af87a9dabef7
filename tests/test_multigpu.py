"""N > 1 on real GPUs (skipped unless the box has enough of them): every sharded path — symmetric InfoNCE through the model, MIL-NCE with
one and two clips per video, the MoCo key all-gather + enqueue — must reproduce the CPU oracle evaluated on the gathered global batch
(tests/mgpu_worker.py). Logs of the 2- and 8-rank runs on B200 are committed under profiles/ (r02_mgpu_parity_*.log)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_paths_match_oracle_on_gathered_batch(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs (gpurun --gpus {world})")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(29631 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MGPU ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
