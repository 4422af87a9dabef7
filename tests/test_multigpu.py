"""N > 1 on real GPUs (skipped unless the box has >= 2): the sharded contrastive step (all-gather of embeddings, per-rank logit
blocks, reduce-scatter of remote-row gradients, W * local-share loss) must reproduce the single-process full-batch step."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_sharded_contrastive_step_matches_full_batch():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(ROOT, "tests", "mgpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU" in r.stdout and " OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
