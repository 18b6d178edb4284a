"""2-GPU NCCL run of the sharded bundle adjuster: needs two devices, skipped otherwise
(run by hand with `gpurun --gpus 2 -- python -m pytest tests/test_bundle_sharded_gpu.py -m gpu`)."""
import json
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("shape", [(12, 800, 4000), (30, 3000, 14000)])
def test_sharded_compute_matches_single_gpu(shape, exchange):
    """exchange: the reduced system travels through the NVLink peer window (two ranks of one node, the default) or
    through NCCL (PTAM_B200_NO_PEER: what larger worlds and other nodes use)."""
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    if exchange == "nccl":
        env["PTAM_B200_NO_PEER"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "dist_ba_worker.py"),
           "--cams", str(shape[0]), "--points", str(shape[1]), "--meas", str(shape[2])]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("SHARDED_BA_RESULT ")]
    assert r.returncode == 0 and line, r.stdout[-3000:] + r.stderr[-3000:]
    for res in json.loads(line[0][len("SHARDED_BA_RESULT "):]):
        assert res["accepted"][0] == res["accepted"][1] and res["trials"][0] == res["trials"][1]
        assert res["outliers_equal"] and res["max_pt"] < 1e-6 and res["max_cam"] < 1e-6
