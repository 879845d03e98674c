"""Host-resident frames through the GPU remap path with copies and compute overlapped.

The remap kernel moves a 4K frame in ~50 us; PCIe moves it in ~1 ms each way.  For
frames that live in host memory (decoded video, the reference's own demo scripts:
scripts/equi2pers_torch.py:61-118 convert and upload frame by frame) the end-to-end
rate is therefore set by the bus, and the only thing software can do is keep the bus
busy in BOTH directions at once.  ``HostPipeline`` splits a pinned batch into chunks
and runs  H2D(chunk i+1) | remap(chunk i) | D2H(chunk i-1)  on three streams.

Memory is bounded: ``depth`` persistent device input buffers form a ring (chunk i + depth
is uploaded into the buffer of chunk i once chunk i's remap has finished), and the host
stops enqueueing when ``depth`` chunks are in flight (it waits on the D2H event of chunk
i - depth), so at most ``depth`` chunk outputs exist on the device at a time -- a video of
any length runs in ``depth`` x (input chunk + output chunk) bytes of device memory.

Completion contract: by default (``sync=True``) the call returns after the last D2H copy
has finished -- the returned pinned tensor is safe to read on the CPU.  With
``sync=False`` the call returns as soon as everything is enqueued; ``pipe.done`` is the
CUDA event of the last D2H copy (``pipe.done.synchronize()`` before touching ``out`` on
the host), and the caller's current stream already waits on it.

The reference has no counterpart (it uploads the whole sampling grid per call on the
default stream, equilib/equi2equi/torch.py:179-180).
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch


class HostPipeline:
    """``pipe = HostPipeline(fn, chunk=4)``; ``out = pipe(frames_pinned, rots, out=out_pinned)``.

    ``fn(img_cuda, rots_chunk) -> out_cuda`` is any transform of this package bound to
    its configuration (e.g. an ``Equi2Equi`` instance or a ``functools.partial``).
    ``frames_pinned`` is a (B, C, H, W) tensor in pinned host memory; ``out`` (optional)
    a pinned tensor for the result, allocated on first use otherwise and reused.
    """

    def __init__(self, fn: Callable, chunk: int = 4, device: Optional[torch.device] = None,
                 depth: int = 3):
        if chunk <= 0:
            raise ValueError("chunk must be positive")
        if depth < 2:
            raise ValueError("depth must be at least 2 (one chunk in flight per direction)")
        self.fn = fn
        self.chunk = chunk
        self.depth = depth
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._h2d = torch.cuda.Stream(self.device)
        self._run = torch.cuda.Stream(self.device)
        self._d2h = torch.cuda.Stream(self.device)
        self._out: Optional[torch.Tensor] = None
        self._ring: List[torch.Tensor] = []
        self.done: Optional[torch.cuda.Event] = None

    @staticmethod
    def chunks(n: int, chunk: int) -> List[range]:
        return [range(a, min(a + chunk, n)) for a in range(0, n, chunk)]

    def _ring_for(self, frames: torch.Tensor) -> List[torch.Tensor]:
        shape = (self.chunk,) + tuple(frames.shape[1:])
        if (len(self._ring) != self.depth or self._ring[0].shape != shape
                or self._ring[0].dtype != frames.dtype):
            self._ring = [torch.empty(shape, dtype=frames.dtype, device=self.device)
                          for _ in range(self.depth)]
        return self._ring

    def __call__(self, frames: torch.Tensor, rots: Sequence[dict],
                 out: Optional[torch.Tensor] = None, sync: bool = True) -> torch.Tensor:
        if frames.is_cuda:
            raise ValueError("HostPipeline takes host frames; call the transform directly")
        if frames.dim() != 4 or len(frames) != len(rots):
            raise ValueError("frames must be (B, C, H, W) with one rotation per frame")
        if not frames.is_pinned():
            frames = frames.pin_memory()
        parts = self.chunks(len(frames), self.chunk)
        ring = self._ring_for(frames)
        # the caller's stream must be done producing `frames`
        start = torch.cuda.current_stream(self.device).record_event()
        self._h2d.wait_event(start)

        ran: List[Optional[torch.cuda.Event]] = [None] * len(parts)     # remap of chunk i done
        copied: List[Optional[torch.cuda.Event]] = [None] * len(parts)  # D2H of chunk i done

        def upload(i: int):
            r = parts[i]
            buf = ring[i % self.depth][:len(r)]
            with torch.cuda.stream(self._h2d):
                if i >= self.depth:  # the buffer's previous tenant must have been consumed
                    self._h2d.wait_event(ran[i - self.depth])
                buf.copy_(frames[r.start:r.stop], non_blocking=True)
                return buf, self._h2d.record_event()

        staged = [upload(i) for i in range(min(self.depth - 1, len(parts)))]
        for i, r in enumerate(parts):
            if i >= self.depth:
                # bound the outputs alive on the device: chunk i - depth must be back on the host
                copied[i - self.depth].synchronize()
            nxt = i + self.depth - 1
            if nxt < len(parts) and nxt >= len(staged):
                staged.append(upload(nxt))
            x, ready = staged[i]
            with torch.cuda.stream(self._run):
                self._run.wait_event(ready)
                y = self.fn(x, list(rots[r.start:r.stop]))
                ran[i] = self._run.record_event()
            if out is None:
                if (self._out is None or self._out.shape[1:] != y.shape[1:]
                        or self._out.shape[0] != len(frames) or self._out.dtype != y.dtype):
                    self._out = torch.empty((len(frames),) + tuple(y.shape[1:]), dtype=y.dtype,
                                            pin_memory=True)
                out = self._out
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(ran[i])
                out[r.start:r.stop].copy_(y, non_blocking=True)
                y.record_stream(self._d2h)
                copied[i] = self._d2h.record_event()
            del y  # (the allocator hands the block back once the D2H copy has run)
        self.done = copied[-1]
        # hand completion to the caller's stream, and -- by default -- to the host
        torch.cuda.current_stream(self.device).wait_event(self.done)
        if sync:
            self.done.synchronize()
        return out


class HostViewsPipeline:
    """One host-resident equirect, many views of it, results back on the host (BASELINE
    config 5: "1 equirect x 256 rotations").  The equirect is uploaded once per call and
    sampled through a stride-0 batch (nothing is replicated); views run in chunks whose
    D2H copies overlap the next chunk's kernel.

        pipe = HostViewsPipeline(Equi2Pers(1024, 1024, 90.0), chunk=32)
        out = pipe(equi_pinned, rots, out=out_pinned)     # (len(rots), C, h, w), pinned

    Same completion contract as HostPipeline (``sync`` / ``done``); at most ``depth``
    chunk outputs are alive on the device.
    """

    def __init__(self, fn: Callable, chunk: int = 32, device: Optional[torch.device] = None,
                 depth: int = 3):
        if chunk <= 0 or depth < 2:
            raise ValueError("chunk must be positive and depth at least 2")
        self.fn, self.chunk, self.depth = fn, chunk, depth
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._run = torch.cuda.Stream(self.device)
        self._d2h = torch.cuda.Stream(self.device)
        self._dev_src: Optional[torch.Tensor] = None
        self._out: Optional[torch.Tensor] = None
        self.done: Optional[torch.cuda.Event] = None

    def __call__(self, equi: torch.Tensor, rots: Sequence[dict],
                 out: Optional[torch.Tensor] = None, sync: bool = True) -> torch.Tensor:
        if equi.is_cuda or equi.dim() != 3:
            raise ValueError("HostViewsPipeline takes one (C, H, W) host image")
        if not equi.is_pinned():
            equi = equi.pin_memory()
        if (self._dev_src is None or self._dev_src.shape != equi.shape
                or self._dev_src.dtype != equi.dtype):
            self._dev_src = torch.empty(equi.shape, dtype=equi.dtype, device=self.device)
        start = torch.cuda.current_stream(self.device).record_event()
        self._run.wait_event(start)
        copied: List[torch.cuda.Event] = []
        parts = HostPipeline.chunks(len(rots), self.chunk)
        with torch.cuda.stream(self._run):
            self._dev_src.copy_(equi, non_blocking=True)
        for i, r in enumerate(parts):
            if i >= self.depth:
                copied[i - self.depth].synchronize()
            with torch.cuda.stream(self._run):
                y = self.fn(self._dev_src[None].expand(len(r), -1, -1, -1),
                            list(rots[r.start:r.stop]))
                ran = self._run.record_event()
            if out is None:
                if (self._out is None or self._out.shape[1:] != y.shape[1:]
                        or self._out.shape[0] != len(rots) or self._out.dtype != y.dtype):
                    self._out = torch.empty((len(rots),) + tuple(y.shape[1:]), dtype=y.dtype,
                                            pin_memory=True)
                out = self._out
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(ran)
                out[r.start:r.stop].copy_(y, non_blocking=True)
                y.record_stream(self._d2h)
                copied.append(self._d2h.record_event())
            del y
        self.done = copied[-1]
        torch.cuda.current_stream(self.device).wait_event(self.done)
        if sync:
            self.done.synchronize()
        return out
