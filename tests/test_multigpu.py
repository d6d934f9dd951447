"""Multi-GPU parity on real hardware (needs >= 2 visible GPUs, otherwise skipped): one process per GPU under torchrun, NCCL,
row tiles with the peer hand-off of the in-place pass; rank 0 compares the assembled frame with the unsharded frame."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scene,fb_w,fb_h,ss", [("boxes", 60, 34, 2), ("knot:60x16", 64, 36, 4)])
def test_sharded_frame_equals_unsharded_on_real_gpus(scene, fb_w, fb_h, ss):
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    for env_extra in ({}, {"YCGE_NO_PEERS": "1"}):  # peer hand-off, then NCCL send/recv hand-off
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
               "--master-port", "29533", os.path.join(ROOT, "tools", "multigpu_check.py"), scene, str(fb_w), str(fb_h), str(ss), "3"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240, env=dict(os.environ, **env_extra))
        assert r.returncode == 0, r.stdout[-2000:]
        assert "!= unsharded" not in r.stdout, r.stdout[-2000:]
        assert r.stdout.count(": sharded x") == 3 and r.stdout.count("== unsharded") >= 3, r.stdout[-2000:]  # + the pipelined frames with the peer hand-off
