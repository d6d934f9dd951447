"""Row-tile sharding of one frame over N processes (one per GPU), SURVEY.md 8(e).

The framebuffer is cut into contiguous tiles of cell rows; the scene is replicated.  Per frame every rank
  1. traces / TAA-blends / filters its tile plus the pixel-row halo the image passes need (recomputed, bit-identical),
  2. for the reference's in-place à-trous iteration (RaytraceRenderer.cs:718) — whose rows depend on ALL rows above —
     receives the few boundary rows from the rank above, runs its part of the wavefront, and passes its own boundary
     rows on to the rank below (point-to-point, the path's one real exchange step),
  3. contributes its log-luminance samples to a sum all-reduce (every slot is owned by exactly one rank, so the sum is
     a gather; every rank then adds them in the reference's serial order — identical exposure everywhere),
  4. converts its tile to console cells, which are gathered on rank 0.

`TileBackend` is the five-call interface of the C ABI (ycge_frame_begin / _halo / _inplace / _finish + buffers) seen
as torch tensors; `CudaTileBackend` is the product implementation on libycge.so.  The collectives are torch.distributed
(NCCL on GPUs; the orchestration is backend-agnostic and is tested on CPU with gloo and a fake tile backend).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import api


def tile_rows(rank: int, world: int, fb_h: int) -> Tuple[int, int]:
    """Cell rows [row0, row0 + rows) owned by `rank`: contiguous, aligned to cell rows, sizes differ by at most one."""
    row0 = rank * fb_h // world
    row1 = (rank + 1) * fb_h // world
    return row0, row1 - row0


def balanced_tiles(tiles: List[Tuple[int, int]], trace_ms: List[float], fb_h: int, per_row_ms: float = 0.019, min_rows: int = 1) -> List[Tuple[int, int]]:
    """Re-cut the row tiles so that every rank gets the same modelled cost.  The trace cost of a cell row is very uneven
    (sky rows end after one ray, mesh rows trace six), everything else is proportional to the number of rows
    (`per_row_ms`: à-trous passes + the wavefront's lag per row).  Density model: the measured trace time of a tile,
    spread evenly over its rows.  Tiles stay contiguous (the wavefront hands rows from rank to rank)."""
    world = len(tiles)
    cost = np.zeros(fb_h)
    for (row0, rows), t in zip(tiles, trace_ms):
        cost[row0:row0 + rows] = max(t, 0.0) / max(rows, 1) + per_row_ms
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    edges = [0]
    for g in range(1, world):
        target = cum[-1] * g / world
        e = int(np.searchsorted(cum, target))
        e = min(max(e, edges[-1] + min_rows), fb_h - (world - g) * min_rows)
        edges.append(e)
    edges.append(fb_h)
    return [(edges[g], edges[g + 1] - edges[g]) for g in range(world)]


class _DevMem:
    """A raw device range exposed through __cuda_array_interface__ so that torch can alias it (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_bytes(ptr: int, nbytes: int, device: int) -> torch.Tensor:
    return torch.as_tensor(_DevMem(ptr, nbytes), device=torch.device("cuda", device))


class CudaTileBackend:
    """One row tile on one GPU through the C ABI.  All work is enqueued on torch's current stream."""

    def __init__(self, scene: api.HostScene, fb_w: int, fb_h: int, ss: int, row0: int, rows: int, device: int):
        self.r = api.CudaRaytraceRenderer(scene, fb_w, fb_h, ss, device=device, tile_row0=row0, tile_rows=rows)
        self.device = device
        self.fb_w, self.fb_h, self.rows, self.ss = fb_w, fb_h, rows, ss
        self.r.set_stream(torch.cuda.current_stream(device).cuda_stream)
        p, n = self.r.device_ptr(api.PTR_LOG_SAMPLES)
        self.logs = device_bytes(p, n, device).view(torch.float32)
        p, n = self.r.device_ptr(api.PTR_CELLS)
        self.cells = device_bytes(p, n, device)

    def set_camera(self, pos, yaw, pitch):
        self.r.SetCamera(pos, yaw, pitch)

    def peer_export(self) -> bytes:
        return bytes(self.r.peer_export())

    def peer_attach(self, above: Optional[bytes], below: Optional[bytes], via_ipc: bool):
        """Attach to the neighbours' exported buffers: the wavefront kernels then hand the boundary rows of the in-place pass
        over themselves (peer stores over NVLink) and `halo()` reports nothing to send or receive."""
        mk = lambda b: api.Peer.from_buffer_copy(b) if b is not None else None
        self.r.peer_attach(mk(above), mk(below), via_ipc)

    def begin(self):
        self.r.frame_begin()

    def halo(self) -> Optional[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]]:
        h = self.r.frame_halo()
        if h is None:
            return None
        recv = device_bytes(h.recv_ptr, h.recv_bytes, self.device) if h.recv_bytes else None
        send = device_bytes(h.send_ptr, h.send_bytes, self.device) if h.send_bytes else None
        return recv, send

    def inplace(self):
        self.r.frame_inplace()

    def finish(self):
        self.r.frame_finish()

    # -- frame pipelining: park a finished tile, finish it later on another stream
    def stash_config(self, n_slots: int):
        self.r.stash_config(n_slots)
        self.slot_logs = []
        for k in range(n_slots):
            p, n = self.r.stash_logs_ptr(k)
            self.slot_logs.append(device_bytes(p, n, self.device).view(torch.float32))

    def stash(self, slot: int):
        self.r.frame_stash(slot)

    def finish_stashed(self, slot: int, stream: "torch.cuda.Stream"):
        self.r.frame_finish_stashed(slot, stream.cuda_stream)

    def close(self):
        self.r.close()


class ShardedRenderer:
    """IConsoleRenderer over N ranks: SetCamera + TryFlipAndBlit, the assembled frame lands on rank 0."""

    def __init__(self, backend, rank: int, world: int, fb_w: int, fb_h: int, group=None, peers: bool = True, tiles=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.b, self.rank, self.world, self.fb_w, self.fb_h = backend, rank, world, fb_w, fb_h
        self.tiles = list(tiles) if tiles is not None else [tile_rows(r, world, fb_h) for r in range(world)]
        self.max_rows = max(t[1] for t in self.tiles)
        self.cell_bytes = api.CELL_DTYPE.itemsize
        dev = backend.cells.device
        self._pad = torch.zeros(self.max_rows * fb_w * self.cell_bytes, dtype=torch.uint8, device=dev)
        self._gather = [torch.zeros_like(self._pad) for _ in range(world)] if rank == 0 else None
        self.peer_handoff = False
        # the peer hand-off needs every tile to be at least as tall as the in-place pass reaches (4 pixel rows by default);
        # every rank evaluates the same condition, so all of them take the same path
        tall_enough = min(t[1] for t in self.tiles) * 2 * getattr(backend, "ss", 1) >= 4
        if world > 1 and peers and tall_enough and hasattr(backend, "peer_export"):
            mine = backend.peer_export()
            everyone: List[Optional[bytes]] = [None] * world
            dist.all_gather_object(everyone, mine, group=group)
            backend.peer_attach(everyone[rank - 1] if rank > 0 else None, everyone[rank + 1] if rank < world - 1 else None, via_ipc=True)
            self.peer_handoff = True

    def SetCamera(self, pos, yaw, pitch):
        self.b.set_camera(pos, yaw, pitch)

    def render_device(self):
        """One frame, everything enqueued on the current stream; returns the gathered tiles (rank 0) without host sync."""
        dist, b = self.dist, self.b
        b.begin()
        while True:
            h = b.halo()
            if h is None:
                break
            recv, send = h
            if recv is not None and self.rank > 0:
                dist.recv(recv, src=self.rank - 1, group=self.group)
            b.inplace()
            if send is not None and self.rank < self.world - 1:
                dist.send(send, dst=self.rank + 1, group=self.group)
        if self.world > 1:
            dist.all_reduce(b.logs, op=dist.ReduceOp.SUM, group=self.group)
        b.finish()
        n = self.tiles[self.rank][1] * self.fb_w * self.cell_bytes
        if self.world == 1:
            return [b.cells]
        self._pad[:n].copy_(b.cells[:n])
        dist.gather(self._pad, self._gather, dst=0, group=self.group)
        return self._gather

    def render_pipelined(self, n_frames: int, collect: bool = False, set_camera=None):
        """Asynchronous path, frames pipelined over the ranks (needs the peer hand-off).  The current stream renders
        frame after frame (front end, wavefront part, stash); a side stream finishes frames in order: all-reduce of the
        stashed exposure samples, ordered exposure sum + cells, gather on rank 0.  Rank g finishes frame h only after it
        has rendered frame h + (world-1-g): the ranks run skewed by one frame each (rank 0 ahead), so that rank g's part
        of the wavefront of frame f overlaps with rank g-1's part of frame f+1, and the k-th collective of every rank
        still is the same frame at about the same time.  Returns the gathered tiles of every frame when `collect`."""
        assert self.world == 1 or self.peer_handoff, "render_pipelined needs the peer hand-off"
        dist, b = self.dist, self.b
        main = torch.cuda.current_stream()
        if not hasattr(self, "_comm"):
            self._comm = torch.cuda.Stream()
            self._slots = max(2, 2 * self.world)
            b.stash_config(self._slots)
            self._ev_stash = [torch.cuda.Event() for _ in range(self._slots)]
            self._ev_fin = [torch.cuda.Event() for _ in range(self._slots)]
        comm, S, D = self._comm, self._slots, self.world - 1 - self.rank
        n_tile = self.tiles[self.rank][1] * self.fb_w * self.cell_bytes
        out = []

        def finish(h):
            slot = h % S
            with torch.cuda.stream(comm):
                comm.wait_event(self._ev_stash[slot])
                if self.world > 1:
                    dist.all_reduce(b.slot_logs[slot], op=dist.ReduceOp.SUM, group=self.group)
                b.finish_stashed(slot, comm)
                if self.world > 1:
                    self._pad[:n_tile].copy_(b.cells[:n_tile])
                    dist.gather(self._pad, self._gather, dst=0, group=self.group)
                    if collect and self.rank == 0:
                        out.append([g.clone() for g in self._gather])
                elif collect:
                    out.append([b.cells.clone()])
                self._ev_fin[slot].record(comm)

        for f in range(n_frames):
            slot = f % S
            if f >= S:
                main.wait_event(self._ev_fin[slot])  # the slot's previous frame has been finished
            if set_camera is not None:
                set_camera(f)
            b.begin()
            while b.halo() is not None:
                b.inplace()
            b.stash(slot)
            self._ev_stash[slot].record(main)
            if f >= D:
                finish(f - D)
        for h in range(max(0, n_frames - D), n_frames):
            finish(h)
        main.wait_stream(comm)
        return out

    def close(self):
        """Unmap the neighbours' buffers on every rank BEFORE any rank frees them, then release the tile."""
        if self.peer_handoff:
            self.b.peer_attach(None, None, via_ipc=False)
            self.peer_handoff = False
        if self.world > 1:
            torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
        self.b.close()

    def assemble(self, gathered) -> np.ndarray:
        """Rank 0: the gathered tiles as one (fb_h, fb_w) cell array on the host."""
        out = np.empty((self.fb_h, self.fb_w), api.CELL_DTYPE)
        for r, (row0, rows) in enumerate(self.tiles):
            n = rows * self.fb_w * self.cell_bytes
            out[row0:row0 + rows] = gathered[r][:n].cpu().numpy().view(api.CELL_DTYPE).reshape(rows, self.fb_w)
        return out

    def TryFlipAndBlit(self) -> Optional[np.ndarray]:
        g = self.render_device()
        return self.assemble(g) if self.rank == 0 else None


class NullEvent:
    """Stream / event stand-ins for a backend whose work is synchronous (the CPU fake of tests/test_sharding.py)."""

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass


class NullStream:
    cuda_stream = 0

    def wait_event(self, ev):
        pass

    def wait_stream(self, st):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class CudaFrameBackend:
    """One rank of the frame-parallel path on one GPU: a row-tile ctx for the FRONT, a whole-frame ctx with back slots for
    BACK + FINISH, and the streams they run on, through the C ABI (ycge_frame_front, ycge_back_*)."""
    is_cuda = True

    def __init__(self, scene: api.HostScene, fb_w: int, fb_h: int, ss: int, device: int, row0: int, rows: int, back_slots: int):
        self.dev_index = device
        self.device = torch.device("cuda", device)
        self.front_r = api.CudaRaytraceRenderer(scene, fb_w, fb_h, ss, device=device, tile_row0=row0, tile_rows=rows)
        self.back_r = api.CudaRaytraceRenderer(scene, fb_w, fb_h, ss, device=device)
        self.back_r.back_config(back_slots)
        # FRONT stream at high priority: its kernels are short and every rank's front of frame f+1 waits for them, while the
        # BACK kernels of other frames (large grids) would otherwise occupy every SM slot ahead of them
        self.main = torch.cuda.Stream(device, priority=-1)
        self.front_r.set_stream(self.main.cuda_stream)
        self.s_comm = torch.cuda.Stream(device, priority=-1)   # the gathers: decoupled from the fronts by a staging ring
        self.s_back = [torch.cuda.Stream(device) for _ in range(back_slots)]
        self.s_fin = torch.cuda.Stream(device)
        W, H = fb_w * ss, fb_h * 2 * ss
        self.n_px_bytes = W * H * 16
        kinds = (api.PTR_HIST, api.PTR_GND, api.PTR_GAS)
        self.slot_planes = [[device_bytes(self.back_r.back_ptr(k, kind)[0], self.n_px_bytes, device) for kind in kinds] for k in range(back_slots)]
        self.slot_cells = [device_bytes(*self.back_r.back_ptr(k, api.PTR_CELLS), device) for k in range(back_slots)]
        p, n = self.back_r.device_ptr(api.PTR_EXPOSURE)
        self.expo = device_bytes(p, n, device)
        self._plane_cache = {}

    def event(self):
        return torch.cuda.Event()

    def on(self, stream):
        return torch.cuda.stream(stream)

    def current_stream(self):
        return torch.cuda.current_stream(self.dev_index)

    def synchronize(self):
        torch.cuda.synchronize(self.dev_index)

    def alloc(self, nbytes: int) -> torch.Tensor:
        return torch.zeros(nbytes, dtype=torch.uint8, device=self.device)

    def set_camera(self, pos, yaw, pitch):
        self.front_r.SetCamera(pos, yaw, pitch)

    def sync_scene(self, scene: api.HostScene, geometry: bool):
        """After scene.update(ms): only the FRONT context traces, so only it needs the moved lights / objects.  Both calls wait
        for the FRONT stream (the fronts already enqueued: short kernels), never for the BACK / FINISH streams."""
        if geometry:
            self.front_r.SyncGeometry(scene)
        self.front_r.SyncLights(scene)

    def front(self):
        self.front_r.frame_front()

    def front_planes(self):
        """This rank's history + guide planes of the frame just rendered (full-frame layout); the guide planes alternate between
        two sets, so the aliasing tensors are cached by pointer."""
        out = []
        for kind in (api.PTR_HIST, api.PTR_GND, api.PTR_GAS):
            p = self.front_r.device_ptr(kind)[0]
            t = self._plane_cache.get(p)
            if t is None:
                t = self._plane_cache[p] = device_bytes(p, self.n_px_bytes, self.dev_index)
            out.append(t)
        return out

    def back_denoise(self, slot: int, stream):
        self.back_r.back_denoise(slot, stream.cuda_stream)

    def back_finish(self, slot: int, stream):
        self.back_r.back_finish(slot, stream.cuda_stream)

    def close(self):
        self.front_r.close()
        self.back_r.close()


class FrameParallelRenderer:
    """Asynchronous path over N ranks, frames in parallel (ycge.h "Frame-parallel sharding").

    Row tiles cannot shorten the wavefront of the in-place à-trous pass, and running it rank after rank leaves most GPUs
    waiting.  But nothing of a frame's à-trous passes feeds the next frame: only the TAA history + guides, the exposure
    scalar and the camera memory do.  So per frame f
      FRONT   every rank traces + TAA-blends its row tile (+1 halo row) on its tile ctx            [all ranks, in step]
      GATHER  the tiles' rows of history + guides go to rank f mod N ("root") into a back slot       [NCCL send/recv]
      BACK    the root runs the à-trous passes and the exposure samples of the whole frame on its whole-frame ctx, on the
              slot's own stream, while all ranks go on with the fronts of the next frames            [N frames in flight]
      FINISH  in frame order around the ring: the root receives the exposure state of frame f-1 from rank (f-1) mod N,
              runs the ordered exposure sum + cells, sends the cells to rank 0 and the state on to rank (f+1) mod N.
    Every frame is bit-identical to the unsharded frame (tools/multigpu_check.py).  The orchestration is backend-agnostic
    (`scene` may be a ready backend object): tests/test_sharding.py runs it over gloo with a fake backend on the CPU."""

    def __init__(self, scene, rank: int, world: int, fb_w: int, fb_h: int, ss: int, device: int = 0, tiles=None, back_slots: int = 2):
        import torch.distributed as dist
        self.dist, self.rank, self.world = dist, rank, world
        self.fb_w, self.fb_h, self.ss = fb_w, fb_h, ss
        self.tiles = list(tiles) if tiles is not None else [tile_rows(r, world, fb_h) for r in range(world)]
        row0, rows = self.tiles[rank]
        self.S = back_slots
        self.b = scene if hasattr(scene, "front_planes") else CudaFrameBackend(scene, fb_w, fb_h, ss, device, row0, rows, back_slots)
        b = self.b
        self.main, self.s_comm, self.s_back, self.s_fin = b.main, b.s_comm, b.s_back, b.s_fin
        self.ev_recv = [b.event() for _ in range(back_slots)]
        self.ev_back = [b.event() for _ in range(back_slots)]
        self.ev_fin = [b.event() for _ in range(back_slots)]
        self.slot_used = [False] * back_slots
        W = fb_w * ss
        self.row_bytes = W * 16
        self.slot_planes, self.slot_cells, self.expo = b.slot_planes, b.slot_cells, b.expo
        self.cell_bytes = fb_w * fb_h * api.CELL_DTYPE.itemsize
        self.out_ring = [b.alloc(self.cell_bytes) for _ in range(4)] if rank == 0 else None
        self.frame = 0  # global frame index (0-based) of the next frame
        self._pending = []  # host-blocking backends (gloo): sends in flight, waited for at the end of a batch
        # staging ring: a copy of this rank's tile rows of (history, normal+depth, albedo+sky) per frame in flight, so that the
        # next frame's TAA may overwrite the history while the rows are still on their way to the root
        self.D = 4
        tile_bytes = self.tiles[rank][1] * 2 * ss * self.row_bytes
        self.stage = [[b.alloc(tile_bytes) for _ in range(3)] for _ in range(self.D)]
        self.ev_stage = [b.event() for _ in range(self.D)]
        self.ev_sent = [b.event() for _ in range(self.D)]
        self.stage_used = [False] * self.D
        self.pg_fin = None
        if world > 1:
            # two communicators: the gathers run in step with the fronts, the FINISH ring runs frames behind them
            self.pg_fin = dist.new_group(list(range(world)))
            t = b.alloc(4).view(torch.float32)
            dist.all_reduce(t)
            dist.all_reduce(t, group=self.pg_fin)
            b.synchronize()
            # Open every point-to-point connection NOW, pair by pair, with nothing else in flight.  NCCL sets a connection up
            # on first use with host-side rendezvous and device allocations; if that happens in the render loop while a
            # send/recv kernel of the other communicator is already spinning on one of the two GPUs, the set-up waits for
            # that kernel, the kernel for its peer, and the peer's host for the set-up: a deadlock (seen on 2 GPUs).
            for grp, pairs in ((None, [(a, c) for a in range(world) for c in range(a + 1, world)]),
                               (self.pg_fin, sorted({(min(r, (r + 1) % world), max(r, (r + 1) % world)) for r in range(world)} | {(0, r) for r in range(1, world)}))):
                for a, c in pairs:
                    if rank == a:
                        dist.send(t, dst=c, group=grp)
                        dist.recv(t, src=c, group=grp)
                    elif rank == c:
                        dist.recv(t, src=a, group=grp)
                        dist.send(t, dst=a, group=grp)
                    b.synchronize()
            dist.barrier()

    def SetCamera(self, pos, yaw, pitch):
        self.b.set_camera(pos, yaw, pitch)

    def SyncScene(self, scene, geometry: bool = False):
        """Per-frame scene changes (Scene.Update: DayNightEntity, orbiting / pulsing lights, bobbing spheres).  Every rank runs the
        same deterministic scene.update(ms) and calls this from the per-frame hook of render() before the frame's FRONT; the
        BACK and FINISH stages read image planes only."""
        self.b.sync_scene(scene, geometry)

    def _send(self, t: torch.Tensor, dst: int, group):
        """A send that never blocks the host.  NCCL: enqueued, the current stream waits for it.  Host-blocking backends (gloo): a
        copy is sent asynchronously and waited for at the end of the batch; a blocking send in the FINISH ring would keep this
        rank from serving the next gathers, which the receiver may be waiting for first."""
        if self.b.is_cuda:
            self.dist.send(t, dst=dst, group=group)
        else:
            keep = t.clone()
            self._pending.append((self.dist.isend(keep, dst=dst, group=group), keep))

    def render(self, n_frames: int, collect: bool = False, set_camera=None, host_ring=None):
        """Enqueue n_frames frames.  `set_camera(f)` is called before frame f is submitted (every rank): the per-frame hook, for
        SetCamera and for SyncScene after a scene.update(ms).  `host_ring` (rank 0):
        a list of pinned uint8 tensors; frame f's cells are copied into host_ring[f % len] as part of the frame, and the host
        waits for that copy before it reuses the entry, i.e. it runs len(host_ring) frames ahead at most (streaming end to
        end: every frame's cells land in host memory)."""
        dist, N, S, rank, b = self.dist, self.world, self.S, self.rank, self.b
        ev_host = [None] * len(host_ring) if host_ring else None
        main, s_comm, s_fin = self.main, self.s_comm, self.s_fin
        caller = b.current_stream()
        main.wait_stream(caller)
        out = []
        py = [(t[0] * 2 * self.ss, (t[0] + t[1]) * 2 * self.ss) for t in self.tiles]
        my_a, my_b = py[rank][0] * self.row_bytes, py[rank][1] * self.row_bytes
        last_root = None
        for i in range(n_frames):
            f = self.frame
            root, slot, d = f % N, (f // N) % S, f % self.D
            if set_camera is not None:
                set_camera(f)
            # ---- FRONT (all ranks), then a copy of the tile's rows into the staging ring
            with b.on(main):
                b.front()
                src = b.front_planes()
                if self.stage_used[d]:
                    main.wait_event(self.ev_sent[d])
                for k in range(3):
                    self.stage[d][k].copy_(src[k][my_a:my_b], non_blocking=True)
                self.ev_stage[d].record(main)
                self.stage_used[d] = True
            # ---- GATHER to the root's back slot (its own stream: the fronts run ahead by up to D frames)
            with b.on(s_comm):
                s_comm.wait_event(self.ev_stage[d])
                if rank == root:
                    if self.slot_used[slot]:
                        s_comm.wait_event(self.ev_fin[slot])  # the slot's previous frame has been finished
                    ops = []
                    for r in range(N):
                        lo, hi = py[r][0] * self.row_bytes, py[r][1] * self.row_bytes
                        for k in range(3):
                            if r == rank:
                                self.slot_planes[slot][k][lo:hi].copy_(self.stage[d][k], non_blocking=True)
                            else:
                                ops.append(dist.P2POp(dist.irecv, self.slot_planes[slot][k][lo:hi], r))
                    if ops:
                        for w in dist.batch_isend_irecv(ops):
                            w.wait()
                    self.ev_recv[slot].record(s_comm)
                    self.slot_used[slot] = True
                else:
                    ops = [dist.P2POp(dist.isend, self.stage[d][k], root) for k in range(3)]
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
                self.ev_sent[d].record(s_comm)
            # ---- BACK on the slot's stream
            if rank == root:
                sb = self.s_back[slot]
                sb.wait_event(self.ev_recv[slot])
                b.back_denoise(slot, sb)
                self.ev_back[slot].record(sb)
            # ---- FINISH ring (frame order), cells to rank 0
            with b.on(s_fin):
                if rank == root:
                    s_fin.wait_event(self.ev_back[slot])
                    if N > 1 and i > 0:
                        dist.recv(self.expo, src=(root - 1) % N, group=self.pg_fin)
                    b.back_finish(slot, s_fin)
                    if rank == 0:
                        self.out_ring[f % 4].copy_(self.slot_cells[slot], non_blocking=True)
                    else:
                        self._send(self.slot_cells[slot], 0, self.pg_fin)
                    if N > 1 and i + 1 < n_frames:
                        self._send(self.expo, (root + 1) % N, self.pg_fin)
                    self.ev_fin[slot].record(s_fin)
                elif rank == 0:
                    dist.recv(self.out_ring[f % 4], src=root, group=self.pg_fin)
                if collect and rank == 0:
                    out.append(self.out_ring[f % 4].clone())
                if host_ring and rank == 0:
                    h = i % len(host_ring)
                    if ev_host[h] is not None:
                        ev_host[h].synchronize()  # the host paces itself on the arrival of the frame len(host_ring) frames back
                    host_ring[h].copy_(self.out_ring[f % 4], non_blocking=True)
                    ev_host[h] = b.event()
                    ev_host[h].record(s_fin)
            last_root = root
            self.frame += 1
        # ---- end of the batch: every rank gets the exposure state, so that the next batch starts without a hand-off
        for w, _ in self._pending:
            w.wait()
        self._pending = []
        with b.on(s_fin):
            if N > 1 and last_root is not None:
                dist.broadcast(self.expo, src=last_root, group=self.pg_fin)
        for st in [main, s_comm, s_fin] + self.s_back:
            caller.wait_stream(st)
        if host_ring and rank == 0:
            for e in ev_host:
                if e is not None:
                    e.synchronize()
        return out

    def cells_host(self, t: torch.Tensor) -> np.ndarray:
        return t.cpu().numpy().view(api.CELL_DTYPE).reshape(self.fb_h, self.fb_w)

    def close(self):
        self.b.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.b.close()
