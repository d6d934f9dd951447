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


def test_multi_device_context_on_real_gpus():
    """ycge_config.n_devices over DISTINCT GPUs (peer copies over NVLink, cross-device events): equal to the one-GPU frame, bit for bit."""
    import numpy as np
    import torch
    from yetanotherconsolegameengine_b200 import api
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    for scene, fb_w, fb_h, ss, pose in (("knot:60x16", 96, 54, 4, api.BENCH_POSE), ("dragon", 480, 135, 4, api.BENCH_POSE)):
        s = api.HostScene(scene)
        one = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss)
        many = api.CudaRaytraceRenderer(s, fb_w, fb_h, ss, devices=list(range(n)))
        one.SetCamera(*pose)
        many.SetCamera(*pose)
        bufs = [np.empty((fb_h, fb_w), api.CELL_DTYPE) for _ in range(2 * n)]
        for batch in (1, 2 * n, n + 1):
            want = [one.TryFlipAndBlit() for _ in range(batch)]
            ids = [many.submit_frame(bufs[k]) for k in range(batch)]
            for k, fid in enumerate(ids):
                many.frame_wait(fid)
                assert bufs[k].tobytes() == want[k].tobytes(), f"{scene}: frame {k} of a batch of {batch} on {n} GPUs"
        many.close(); one.close(); s.close()
