"""Sharding, host-side logic on CPU, world_size 2 and 3 over gloo.

Row tiles in lock step (sharding.ShardedRenderer): the tile partition and the per-frame orchestration (boundary-row hand-off
rank -> rank+1, sum all-reduce of the exposure samples, gather of the cell tiles on rank 0).
Frames in parallel (sharding.FrameParallelRenderer): FRONT tiles gathered into a back slot of the frame's root, BACK + FINISH
round-robin over the ranks, exposure state around the ring, cells to rank 0, two batches.

The backends here are fakes built on the CPU oracle (each rank renders the whole frame with the oracle and exposes only its
tile's slices), so this checks the plumbing, not the kernels; the kernels' tile / halo / front-back logic is checked on the
GPU in test_gpu_parity.py (test_row_tiles_equal_the_unsharded_frame, test_front_tiles_plus_whole_frame_back_equal_the_
unsharded_frame) and over real ranks by tools/multigpu_check.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT, TESTS
from yetanotherconsolegameengine_b200 import api, sharding


def test_tile_partition_covers_every_cell_row_once():
    for fb_h in (1, 2, 7, 90, 135):
        for world in (1, 2, 3, 4, 8):
            tiles = [sharding.tile_rows(r, world, fb_h) for r in range(world)]
            assert tiles[0][0] == 0 and tiles[-1][0] + tiles[-1][1] == fb_h
            for (a0, an), (b0, bn) in zip(tiles, tiles[1:]):
                assert a0 + an == b0
            sizes = [t[1] for t in tiles]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == fb_h


def test_balanced_tiles_equalise_the_modelled_cost():
    fb_h, world = 135, 8
    tiles = [sharding.tile_rows(r, world, fb_h) for r in range(world)]
    trace = [0.03, 0.04, 0.66, 0.82, 0.56, 0.47, 0.45, 0.37]  # measured on 8 B200s (dragon, bench pose): sky on top, mesh in the middle
    new = sharding.balanced_tiles(tiles, trace, fb_h)
    assert new[0][0] == 0 and new[-1][0] + new[-1][1] == fb_h and all(n > 0 for _, n in new)
    for (a0, an), (b0, _) in zip(new, new[1:]):
        assert a0 + an == b0
    assert new[0][1] > tiles[0][1] and new[3][1] < tiles[3][1]  # cheap sky tiles grow, the expensive mesh tile shrinks
    # modelled cost per rank is within one row's cost of the mean
    cost = np.zeros(fb_h)
    for (r0, n), t in zip(tiles, trace):
        cost[r0:r0 + n] = t / n + 0.019
    per_rank = [cost[r0:r0 + n].sum() for r0, n in new]
    assert max(per_rank) - min(per_rank) <= 2 * cost.max() + 1e-9
    # degenerate input keeps a valid partition
    flat = sharding.balanced_tiles(tiles, [0.0] * world, fb_h)
    assert sum(n for _, n in flat) == fb_h and all(n > 0 for _, n in flat)


class FakeTileBackend:
    """TileBackend made of the CPU oracle: full frame per rank, tile slices exposed as torch CPU tensors."""

    def __init__(self, scene_name, fb_w, fb_h, ss, rank, world):
        from oracle_binding import Oracle
        self.scene = api.HostScene(scene_name)
        self.o = Oracle(self.scene, fb_w, fb_h, ss)
        self.rank, self.world, self.fb_w, self.fb_h, self.ss = rank, world, fb_w, fb_h, ss
        self.row0, self.rows = sharding.tile_rows(rank, world, fb_h)
        step = max(2, 2 * ss)
        self.sh, self.sw = (fb_h * 2 * ss + step - 1) // step, (fb_w * ss + step - 1) // step
        self.logs = torch.zeros(self.sh * self.sw, dtype=torch.float32)
        self.cells = torch.zeros(self.rows * fb_w * 32, dtype=torch.uint8)
        self.frame = 0
        self.errors = []

    @staticmethod
    def token(rank, frame, n):
        return torch.full((n,), (17 * rank + 3 * frame + 1) % 251, dtype=torch.uint8)

    def set_camera(self, pos, yaw, pitch):
        self.o.set_camera(pos, yaw, pitch)

    def begin(self):
        self.frame += 1
        self.full_cells = self.o.render_frame(threads=2, fast_post=True)
        self.full_logs = self.o.debug_read(api.DBG_LOG_SAMPLES).reshape(-1).copy()
        self._halo_calls = 0
        self._recv = None

    def halo(self):
        self._halo_calls += 1
        if self._halo_calls > 1:
            return None
        n = 4 * self.fb_w * self.ss * 16
        self._recv = torch.zeros(n, dtype=torch.uint8) if self.rank > 0 else None
        self._send = torch.zeros(n, dtype=torch.uint8) if self.rank < self.world - 1 else None
        return self._recv, self._send

    def inplace(self):
        if self._recv is not None and not torch.equal(self._recv, self.token(self.rank - 1, self.frame, len(self._recv))):
            self.errors.append(f"frame {self.frame}: boundary rows from rank {self.rank - 1} did not arrive before the in-place pass")
        if self._send is not None:
            self._send.copy_(self.token(self.rank, self.frame, len(self._send)))  # valid only after the pass, like the real buffer
        # this rank's exposure samples: sample row k = cell row k (step = 2*ss)
        self.logs.zero_()
        own = torch.from_numpy(self.full_logs.reshape(self.sh, self.sw)[self.row0:self.row0 + self.rows].copy())
        self.logs.view(self.sh, self.sw)[self.row0:self.row0 + self.rows] = own

    def finish(self):
        got, exp = self.logs.numpy(), self.full_logs
        if not (np.array_equal(np.isnan(got), np.isnan(exp)) and np.array_equal(got[~np.isnan(exp)], exp[~np.isnan(exp)])):
            self.errors.append(f"frame {self.frame}: all-reduced exposure samples differ from the full-frame samples")
        tile = np.ascontiguousarray(self.full_cells[self.row0:self.row0 + self.rows])
        self.cells.copy_(torch.from_numpy(tile.view(np.uint8).reshape(-1)))


def _worker(rank, world, port, fb_w, fb_h, ss, frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path[:0] = [ROOT, TESTS]
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = FakeTileBackend("boxes", fb_w, fb_h, ss, rank, world)
        r = sharding.ShardedRenderer(b, rank, world, fb_w, fb_h)  # (the fake backend has no peer buffers: send/recv path)
        ok = True
        for f in range(frames):
            if f == 2:
                r.SetCamera((0.3, 1.2, 0.1), 0.2, -0.1)  # every rank moves its camera identically -> history reset everywhere
            out = r.TryFlipAndBlit()
            if rank == 0:
                ok &= out.tobytes() == b.full_cells.tobytes()
            else:
                ok &= out is None
        q.put((rank, ok and not b.errors, b.errors))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,fb_h", [(2, 9), (3, 10)])
def test_sharded_frame_orchestration_gloo(world, fb_h):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, 24, fb_h, 2, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, errors in results:
        assert ok, (rank, errors)


class FakeFrameBackend:
    """Backend of sharding.FrameParallelRenderer made of the CPU oracle: every rank renders every frame with its own oracle
    (identical everywhere), exposes only its tile's rows of the frame's history / guide planes as the FRONT result, and as
    the frame's root checks what the gather assembled and what the FINISH ring delivered."""
    is_cuda = False

    def __init__(self, scene_name, fb_w, fb_h, ss, rank, world, tiles, back_slots):
        from oracle_binding import Oracle
        self.scene = api.HostScene(scene_name)
        self.o = Oracle(self.scene, fb_w, fb_h, ss)
        self.rank, self.world, self.fb_w, self.fb_h, self.ss = rank, world, fb_w, fb_h, ss
        self.device = torch.device("cpu")
        self.main, self.s_comm, self.s_fin = sharding.NullStream(), sharding.NullStream(), sharding.NullStream()
        self.s_back = [sharding.NullStream() for _ in range(back_slots)]
        W, H = fb_w * ss, fb_h * 2 * ss
        self.n_px_bytes = W * H * 16
        self.py0, self.py1 = tiles[rank][0] * 2 * ss * W * 16, (tiles[rank][0] + tiles[rank][1]) * 2 * ss * W * 16
        self.slot_planes = [[torch.zeros(self.n_px_bytes, dtype=torch.uint8) for _ in range(3)] for _ in range(back_slots)]
        self.slot_cells = [torch.zeros(fb_w * fb_h * 32, dtype=torch.uint8) for _ in range(back_slots)]
        self.expo = torch.zeros(16, dtype=torch.uint8)
        self.frames = []      # per rendered frame: (planes, cells, exposure state after the frame)
        self.slot_frame = {}  # slot -> index of the frame whose BACK ran there
        self.syncs = []       # (frames rendered before the call, geometry) of every sync_scene
        self.done = 0         # frames finished so far, over all ranks (the root of frame f is the only one that counts f)
        self.errors = []

    def event(self):
        return sharding.NullEvent()

    def on(self, stream):
        return stream

    def current_stream(self):
        return sharding.NullStream()

    def synchronize(self):
        pass

    def alloc(self, nbytes):
        return torch.zeros(nbytes, dtype=torch.uint8)

    def set_camera(self, pos, yaw, pitch):
        self.o.set_camera(pos, yaw, pitch)

    def sync_scene(self, scene, geometry):
        self.syncs.append((len(self.frames), geometry))
        if geometry:
            self.o.upload_scene(scene)
        self.o.lights_update(scene.lights())

    @staticmethod
    def state_token(frame_index, ae):
        t = np.zeros(4, np.float32)
        t[0], t[1] = ae, frame_index
        return torch.from_numpy(t.view(np.uint8).copy())

    def front(self):
        cells = self.o.render_frame(threads=2, fast_post=True)
        planes = [torch.from_numpy(self.o.debug_read(k).view(np.uint8).reshape(-1).copy()) for k in (api.DBG_TAA, api.DBG_NORMAL_DEPTH, api.DBG_ALBEDO_SKY)]
        self.frames.append((planes, cells.copy(), self.state_token(len(self.frames), self.o.stats()["ae_exposure"])))

    def front_planes(self):
        out = []
        for p in self.frames[-1][0]:  # only this rank's tile rows are valid, like the tile ctx's planes
            t = torch.full_like(p, 0xEE)
            t[self.py0:self.py1] = p[self.py0:self.py1]
            out.append(t)
        return out

    def back_denoise(self, slot, stream):
        f = len(self.frames) - 1  # synchronous backend: the BACK of a frame runs right after its gather
        self.slot_frame[slot] = f
        for k in range(3):
            if not torch.equal(self.slot_planes[slot][k], self.frames[f][0][k]):
                self.errors.append(f"frame {f}: plane {k} assembled in back slot {slot} differs from the full-frame plane")

    def back_finish(self, slot, stream):
        f = self.slot_frame[slot]
        if f > 0 and not torch.equal(self.expo, self.frames[f - 1][2]):
            self.errors.append(f"frame {f}: exposure state of frame {f - 1} had not arrived before FINISH")
        self.slot_cells[slot].copy_(torch.from_numpy(self.frames[f][1].view(np.uint8).reshape(-1)))
        self.expo.copy_(self.frames[f][2])

    def close(self):
        self.o.close()
        self.scene.close()


def _fp_worker(rank, world, port, fb_w, fb_h, ss, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path[:0] = [ROOT, TESTS]
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tiles = [sharding.tile_rows(r, world, fb_h) for r in range(world)]
        b = FakeFrameBackend("entities_demo", fb_w, fb_h, ss, rank, world, tiles, back_slots=2)
        fp = sharding.FrameParallelRenderer(b, rank, world, fb_w, fb_h, ss, tiles=tiles, back_slots=2)

        def per_frame(f):  # every rank runs the same deterministic scene update, then moves the camera once
            b.scene.update(16.0)
            fp.SyncScene(b.scene, geometry=True)
            if f == 3:
                fp.SetCamera((0.3, 1.2, 0.1), 0.2, -0.1)

        got = fp.render(2 * world + 1, collect=True, set_camera=per_frame) + fp.render(world + 2, collect=True)  # two batches
        ok = b.syncs == [(f, True) for f in range(2 * world + 1)]                          # before each frame's FRONT, in order
        ok &= len({b.frames[f][1].tobytes() for f in range(2 * world + 1)}) > 1             # the scene did move between frames
        if rank == 0:
            ok &= len(got) == 3 * world + 3
            for f, g in enumerate(got):
                ok &= fp.cells_host(g).tobytes() == b.frames[f][1].tobytes()
        else:
            ok &= got == []
        # after a batch every rank holds the exposure state of the last frame
        ok &= torch.equal(b.expo, b.frames[-1][2])
        q.put((rank, bool(ok) and not b.errors, b.errors))
        fp.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,fb_h", [(2, 9), (3, 10)])
def test_frame_parallel_orchestration_gloo(world, fb_h):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_fp_worker, args=(r, world, port, 24, fb_h, 2, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, errors in results:
        assert ok, (rank, errors)
