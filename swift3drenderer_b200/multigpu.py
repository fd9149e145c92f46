"""Multi-GPU host logic: one process per GPU over ``torch.distributed``.

Two partitions (BASELINE.json north_star):

* frame-parallel — ``frame_shard`` deals camera poses to ranks; no collective on the data path.
* screen bands   — ``band_edges`` splits the rows ``[0, H)`` into one contiguous band per rank; every rank
  runs the geometry stages on the whole scene and rasterises only its band (each triangle's barycentric
  walk starts at the triangle's own ymin, so a band boundary changes no pixel); ``assemble_bands``
  concatenates the bands on every rank with one all-gather (NCCL over NVLink on GPUs; gloo in the CPU test).
  ``PeerFrames`` is the fused alternative: no collective moves pixels at all — every rank's shading kernel stores
  its rows straight into the full-size frames of all ranks over NVLink peer memory (CUDA IPC), and a one-element
  all-reduce per frame orders the frames.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def band_edges(height: int, world: int) -> List[Tuple[int, int]]:
    """Rows [y0, y1) of each rank's band: as equal as possible, every band non-empty while world <= height."""
    assert world >= 1 and height >= world
    return [(height * k // world, height * (k + 1) // world) for k in range(world)]


def frame_shard(n_frames: int, rank: int, world: int) -> range:
    """Contiguous block of frame indices for `rank` (camera poses are precomputed on the host)."""
    return range(n_frames * rank // world, n_frames * (rank + 1) // world)


def assemble_bands(band: torch.Tensor, height: int, rank: int, world: int, group=None) -> torch.Tensor:
    """All-gathers per-rank bands ``(rows_k, W)`` into the full ``(H, W)`` frame on every rank.

    Bands may differ by one row, so each rank contributes a buffer padded to the tallest band and the
    padding is dropped after the collective."""
    edges = band_edges(height, world)
    assert band.shape[0] == edges[rank][1] - edges[rank][0]
    width = band.shape[1]
    tallest = max(b - a for a, b in edges)
    send = torch.zeros((tallest, width), dtype=band.dtype, device=band.device)
    send[: band.shape[0]] = band
    recv = torch.empty((world, tallest, width), dtype=band.dtype, device=band.device)
    if world == 1:
        recv[0] = send
    else:
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    return torch.cat([recv[k, : b - a] for k, (a, b) in enumerate(edges)], 0)


def equal_bands(height: int, world: int) -> bool:
    return height % world == 0


def gather_bands_inplace(frame: torch.Tensor, rank: int, world: int, group=None) -> torch.Tensor:
    """Fast path when ``height % world == 0``: every rank has rendered its band straight into rows
    ``band_edges(H, world)[rank]`` of its own full-size ``frame``; one in-place all-gather (NCCL: send buffer =
    receive buffer + rank * count, no staging copy) completes the frame on every rank."""
    height = frame.shape[0]
    assert height % world == 0 and frame.is_contiguous()
    if world > 1:
        y0, y1 = band_edges(height, world)[rank]
        dist.all_gather_into_tensor(frame.view(-1), frame[y0:y1].view(-1), group=group)
    return frame


class InterleavedAssembler:
    """Tile rows dealt round-robin to the ranks (``a % world == rank``): every rank gets an even share of the
    screen whatever the scene looks like.  Each rank renders its rows compacted (``Renderer.render_device_rows``);
    ``gather`` all-gathers the padded per-rank buffers and de-interleaves them into the frame with one
    index_select."""

    def __init__(self, height: int, width: int, rank: int, world: int, device, tile_h: int):
        from .renderer import rows_layout
        self.rank, self.world = rank, world
        tile_rows = (height + tile_h - 1) // tile_h
        self.buf_rows = ((tile_rows + world - 1) // world) * tile_h          # tallest compacted buffer, same on all ranks
        self.mine = torch.zeros((self.buf_rows, width), dtype=torch.int32, device=device)
        self.staging = torch.empty((world, self.buf_rows, width), dtype=torch.int32, device=device)
        src = torch.empty(height, dtype=torch.int64)
        for k in range(world):
            _, frame_rows, buf_rows = rows_layout(height, world, k, tile_h)
            src[torch.from_numpy(frame_rows)] = torch.from_numpy(buf_rows) + k * self.buf_rows
        self.src = src.to(device)

    def gather(self, group=None) -> torch.Tensor:
        if self.world == 1:
            self.staging[0] = self.mine
        else:
            dist.all_gather_into_tensor(self.staging.view(-1), self.mine.view(-1), group=group)
        return self.staging.view(self.world * self.buf_rows, -1).index_select(0, self.src)


def render_banded(render_band: Callable[[int, int], torch.Tensor], height: int, rank: int, world: int,
                  group=None) -> torch.Tensor:
    """`render_band(y0, y1)` must return this rank's rows as an ``(y1 - y0, W)`` int32 tensor (on the GPU it
    wraps ``Renderer.render_device(..., y0=y0, y1=y1)``); returns the assembled frame."""
    y0, y1 = band_edges(height, world)[rank]
    return assemble_bands(render_band(y0, y1), height, rank, world, group)


class PeerFrames:
    """Fused frame assembly: a ring of full-size frames on every rank, each mapped by all other ranks.

    ``render(...)`` makes this rank's renderer store its tile rows into ring slot ``k`` of EVERY rank
    (``Renderer.set_peer_frames`` + ``render_device_rows``); ``fence()`` is the per-frame ordering point (a
    one-element all-reduce).  After the fence of frame k every rank holds the complete frame in its own slot
    ``k % ring``.  Ranks must be processes of one node (CUDA IPC over NVLink / NVSwitch).

    Stream protocol: ``render`` enqueues on ``stream`` (default: torch's current stream, never the renderer's private
    one) and records an event there; ``fence`` makes torch's current stream wait for that event before the
    all-reduce, so the collective is ordered after this rank's peer stores whatever stream rendered them.
    Ring protocol: slot ``k % ring`` is overwritten by frame ``k + ring``; a consumer of frame k must have enqueued
    its reads (on a stream the next ``render`` call's stream waits for, e.g. the same stream) before ``render`` is
    called ``ring`` more times.  Because every rank passes fence k + ring - 1 only after all ranks have enqueued
    the reads that precede it in stream order, ring >= 2 keeps writers of frame k + ring behind readers of frame k.
    """

    def __init__(self, renderer, height: int, width: int, rank: int, world: int, device, ring: int = 2, group=None):
        self.r, self.rank, self.world, self.ring, self.group = renderer, rank, world, ring, group
        self.height, self.width = height, width
        self.device = device
        # the shading kernel stores 16-byte pieces: every ring slot starts on a 16-byte boundary
        self.frame_bytes = (height * width * 4 + 15) & ~15
        self.own_ptr, handle = renderer.peer_frame_alloc(self.frame_bytes * ring)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, handle, group=group)
        else:
            handles[0] = handle
        self.base = [self.own_ptr if k == rank else renderer.peer_frame_open(handles[k]) for k in range(world)]
        self.token = torch.zeros(1, dtype=torch.int32, device=device)
        self.count = 0
        self._rendered = torch.cuda.Event()
        self._pending = False
        self._fence_stream = torch.cuda.Stream(device) if torch.cuda.is_available() else None
        self._fences = {}        # frame number -> event recorded behind that frame's fence (fence_async)
        self._consumed = []      # events the next fence must wait for (consumers that have finished reading)

    def destinations(self, slot: int):
        return [b + slot * self.frame_bytes for b in self.base]

    def render(self, camera, stream: int = 0) -> int:
        """Enqueues this rank's share of the next frame; returns the ring slot it lands in."""
        slot = self.count % self.ring
        self.count += 1
        ext = torch.cuda.ExternalStream(stream, device=self.device) if stream else torch.cuda.current_stream(self.device)
        # pipelined fences (fence_async): frame k overwrites the slot of frame k - ring; every rank has finished with that
        # frame once the fence of frame k - ring + 1 has completed here (a rank contributes to a fence only behind its
        # registered consumers of the frames before it)
        guard = self._fences.pop(self.count - self.ring, None)
        if guard is not None:
            ext.wait_event(guard)
        self.r.set_peer_frames(self.destinations(slot))
        # (handle 0 would mean "the renderer's own stream" to the C ABI: torch's default stream is CUDA's legacy stream, handle 1)
        self.r.render_device_rows(camera, self.width, self.height, self.world, self.rank, 0, stream=ext.cuda_stream or 1)
        self.r.set_peer_frames([])
        self._rendered.record(ext)
        self._pending = True
        return slot

    def fence(self) -> None:
        """Orders the frames across ranks: after it (in stream order on torch's current stream) every rank's rows of
        the frames rendered so far are in place on this rank."""
        if self._pending:
            torch.cuda.current_stream(self.device).wait_event(self._rendered)
            self._pending = False
        if self.world > 1:
            dist.all_reduce(self.token, group=self.group)

    def fence_async(self) -> None:
        """The fence of the frame just rendered, on a side stream: the rendering stream goes straight on to the next
        frame's geometry while the all-reduce is in flight.  ``wait_frame`` / ``join`` order a consumer behind it."""
        fs = self._fence_stream
        fs.wait_event(self._rendered)
        self._pending = False
        for ev in self._consumed:
            fs.wait_event(ev)
        self._consumed = []
        if self.world > 1:
            with torch.cuda.stream(fs):
                dist.all_reduce(self.token, group=self.group)
        ev = torch.cuda.Event()
        ev.record(fs)
        self._fences[self.count - 1] = ev
        for k in [k for k in self._fences if k < self.count - self.ring]:
            del self._fences[k]

    def wait_frame(self, stream=None) -> None:
        """Makes `stream` (default: torch's current stream) wait for the fence of the last fenced frame."""
        if self._fences:
            (stream or torch.cuda.current_stream(self.device)).wait_event(self._fences[max(self._fences)])

    def consumed(self, stream=None) -> None:
        """Call after enqueuing a consumer's reads of assembled frames on `stream`: the next fence waits for them, so no
        rank overwrites a ring slot that is still being read."""
        ev = torch.cuda.Event()
        ev.record(stream or torch.cuda.current_stream(self.device))
        self._consumed.append(ev)

    def join(self, stream=None) -> None:
        self.wait_frame(stream)

    def drain(self) -> None:
        if self._fence_stream is not None:
            self._fence_stream.synchronize()

    def slot_ptr(self, slot: int) -> int:
        """Device pointer of this rank's copy of ring slot `slot`."""
        return self.own_ptr + slot * self.frame_bytes

    def read(self, slot: int):
        """This rank's copy of ring slot `slot` as a host array (call after the fence and a device synchronize)."""
        return self.r.read_device(self.slot_ptr(slot), (self.height, self.width))

    def close(self) -> None:
        for k, b in enumerate(self.base):
            self.r.peer_frame_release(b)
        self.base = []
