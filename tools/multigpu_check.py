"""Run under torchrun on N GPUs: renders frames through the row-tile sharding (peer hand-off when possible) and checks the
assembled cells on rank 0 against the unsharded frame rendered on rank 0's GPU, bit for bit.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multigpu_check.py [scene fb_w fb_h ss frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import yetanotherconsolegameengine_b200 as pkg
from yetanotherconsolegameengine_b200 import api, sharding

scene_name = sys.argv[1] if len(sys.argv) > 1 else "dragon"
fb_w, fb_h, ss, frames = (int(x) for x in (sys.argv[2:6] if len(sys.argv) > 5 else (480, 135, 4, 3)))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scene = pkg.HostScene(scene_name)
pose = pkg.BENCH_POSE if (scene.n_meshes == 1 and scene.n_volumes == 0) else scene.default_camera()[:3]
row0, rows = sharding.tile_rows(rank, world, fb_h)
b = sharding.CudaTileBackend(scene, fb_w, fb_h, ss, row0, rows, local)
sr = sharding.ShardedRenderer(b, rank, world, fb_w, fb_h, peers=not os.environ.get("YCGE_NO_PEERS"))
sr.SetCamera(*pose)
full = None
if rank == 0:
    full = pkg.CudaRaytraceRenderer(scene, fb_w, fb_h, ss, device=local)
    full.SetCamera(*pose)
ok = True
for f in range(frames):
    got = sr.TryFlipAndBlit()
    if rank == 0:
        ref = full.TryFlipAndBlit()
        same = got.tobytes() == ref.tobytes()
        ok &= same
        print(f"frame {f + 1}: sharded x{world} (peer_handoff={sr.peer_handoff}) {'==' if same else '!='} unsharded", flush=True)
# the pipelined asynchronous path: same frames, finished out of step with their rendering
if sr.peer_handoff or world == 1:
    got = sr.render_pipelined(frames + 3, collect=True)
    torch.cuda.synchronize()
    if rank == 0:
        for f, g in enumerate(got):
            ref = full.TryFlipAndBlit()
            same = sr.assemble(g).tobytes() == ref.tobytes()
            ok &= same
            print(f"frame {frames + f + 1}: pipelined x{world} {'==' if same else '!='} unsharded", flush=True)
# the frame-parallel asynchronous path: FRONT on row tiles, BACK + FINISH of whole frames round-robin over the ranks; its own
# contexts, so it starts again at frame 1 (compared with a fresh unsharded renderer), two batches
fp = sharding.FrameParallelRenderer(scene, rank, world, fb_w, fb_h, ss, local, back_slots=2)
fp.SetCamera(*pose)
n_fp = 2 * world + 3
got = fp.render(n_fp, collect=True) + fp.render(world + 1, collect=True)
torch.cuda.synchronize()
if rank == 0:
    full2 = pkg.CudaRaytraceRenderer(scene, fb_w, fb_h, ss, device=local)
    full2.SetCamera(*pose)
    for f, g in enumerate(got):
        ref = full2.TryFlipAndBlit()
        same = fp.cells_host(g).tobytes() == ref.tobytes()
        ok &= same
        print(f"frame {f + 1}: frame-parallel x{world} {'==' if same else '!='} unsharded", flush=True)
    full2.close()
fp.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
sr.close()
dist.destroy_process_group()
sys.exit(0 if int(flag[0]) else 1)
