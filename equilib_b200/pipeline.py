"""Host-resident frames through the GPU remap path with copies and compute overlapped.

The remap kernel moves a 4K frame in ~70 us; PCIe moves it in ~1 ms each way.  For
frames that live in host memory (decoded video, the reference's own demo scripts:
scripts/equi2pers_torch.py:61-118 convert and upload frame by frame) the end-to-end
rate is therefore set by the bus, and the only thing software can do is keep the bus
busy in BOTH directions at once.  ``HostPipeline`` splits a pinned batch into chunks
and runs  H2D(chunk i+1) | remap(chunk i) | D2H(chunk i-1)  on three streams.

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

    def __init__(self, fn: Callable, chunk: int = 4, device: Optional[torch.device] = None):
        if chunk <= 0:
            raise ValueError("chunk must be positive")
        self.fn = fn
        self.chunk = chunk
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._h2d = torch.cuda.Stream(self.device)
        self._run = torch.cuda.Stream(self.device)
        self._d2h = torch.cuda.Stream(self.device)
        self._out: Optional[torch.Tensor] = None

    @staticmethod
    def chunks(n: int, chunk: int) -> List[range]:
        return [range(a, min(a + chunk, n)) for a in range(0, n, chunk)]

    def __call__(self, frames: torch.Tensor, rots: Sequence[dict],
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if frames.is_cuda:
            raise ValueError("HostPipeline takes host frames; call the transform directly")
        if frames.dim() != 4 or len(frames) != len(rots):
            raise ValueError("frames must be (B, C, H, W) with one rotation per frame")
        if not frames.is_pinned():
            frames = frames.pin_memory()
        parts = self.chunks(len(frames), self.chunk)
        # the caller's stream must be done producing `frames`
        start = torch.cuda.current_stream(self.device).record_event()
        self._h2d.wait_event(start)

        staged = []
        for r in parts:  # enqueue every upload; they run back to back on one stream
            with torch.cuda.stream(self._h2d):
                x = frames[r.start:r.stop].to(self.device, non_blocking=True)
                staged.append((x, self._h2d.record_event()))

        results = []
        for r, (x, ready) in zip(parts, staged):
            with torch.cuda.stream(self._run):
                self._run.wait_event(ready)
                y = self.fn(x, list(rots[r.start:r.stop]))
                x.record_stream(self._run)
                done = self._run.record_event()
            if out is None:
                if (self._out is None or self._out.shape[1:] != y.shape[1:]
                        or self._out.shape[0] != len(frames) or self._out.dtype != y.dtype):
                    self._out = torch.empty((len(frames),) + tuple(y.shape[1:]), dtype=y.dtype,
                                            pin_memory=True)
                out = self._out
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(done)
                out[r.start:r.stop].copy_(y, non_blocking=True)
                y.record_stream(self._d2h)
            results.append(y)
        # hand completion back to the caller's stream
        torch.cuda.current_stream(self.device).wait_event(self._d2h.record_event())
        return out
