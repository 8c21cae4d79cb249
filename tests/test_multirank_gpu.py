"""2-GPU test (skipped with fewer devices): time-sliced receiver and NCCL-fed channel shards reproduce the
single-GPU results bit for bit.  Launched the way the driver launches bench.py for N>1 (torchrun)."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

import sdrpkg

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_two_gpus_reproduce_one_gpu_bitwise(tmp_path):
    S = sdrpkg.load()
    if S.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "result.json"
    env = dict(os.environ, MULTIRANK_OUT=str(out))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "multirank_gpu_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, "\n".join(ln for ln in (r.stdout + r.stderr).splitlines() if "NCCL INFO" not in ln)[-3000:]
    assert json.loads(out.read_text()) == {"ok": True, "world": 2}
