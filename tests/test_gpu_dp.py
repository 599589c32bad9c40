"""Data-parallel correctness on hardware (tests/dp): needs >= 2 GPUs, so the single-GPU `-m gpu` run
skips it; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu -s` runs it (record: profiles/r2_dp_test.txt).
The checks live in tests/dp_worker.py (one process per GPU under torchrun, NCCL)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="data-parallel hardware test needs >= 2 GPUs")
def test_dp_gradients_and_parameters_agree_across_ranks():
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    print(r.stdout[-6000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["world"] == world
    for k, v in rep.items():
        if k.endswith("_identical"):
            assert v is True, k
