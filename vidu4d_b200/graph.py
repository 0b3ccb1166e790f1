"""CUDA-graph capture of a sync-free rasterizer step, with the two re-capture paths a Stage-3 run needs.

The forward of this library never synchronises with the host (instance buffers have a CAPACITY; `num_rendered` and a status
word land in pinned memory asynchronously), which is what makes a whole warp -> rasterize -> loss -> backward step
capturable -- the reference cannot be captured: it blocks on a D2H copy in the middle of forward()
(RAST/cuda_rasterizer/rasterizer_impl.cu:282).  The price is that capacities and tensor shapes are baked into the graph:

  * densification changes the number of surfels every 100 steps (lab4d/engine/trainer.py:562-568,
    gs/scene/gaussian_model.py:431-446)                  -> the caller's `key()` changes -> re-capture;
  * the scene can outgrow the instance capacity baked in -> check_overflow() raises, the capacity hints are refreshed
                                                         -> re-capture with the larger buffers and re-run the step.
"""
from __future__ import annotations

import torch

from . import _capi, rasterizer as RZ


class GraphedStep:
    """Wraps `body()` (kernel launches only: no .item(), no host sync) into a CUDA graph.

    body   : callable; everything it allocates lives in the graph's private pool.
    key    : optional callable returning a hashable description of the shapes the body depends on (e.g. the surfel
             count); a change triggers re-capture.
    warmup : eager runs before capture (allocator pools, capacity hints, lazy initialisation).
    """

    def __init__(self, body, key=None, warmup: int = 2, device=None):
        self.body, self.key_fn, self.warmup = body, key, warmup
        self.device = device
        self.graph, self.key, self.watch = None, None, []
        self.captures = 0

    def _capture(self):
        dev = self.device
        self.graph = None                               # frees the previous graph's pool before the new capture
        RZ._pending.clear()
        RZ.set_sync_mode(True)
        self.body()                                     # learns the instance capacity of the current scene (one sync)
        RZ.set_sync_mode(False)
        for _ in range(self.warmup):
            self.body()
        RZ.check_overflow()
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(warm):
            self.body()
            RZ.check_overflow()
        torch.cuda.current_stream(dev).wait_stream(warm)
        torch.cuda.synchronize(dev)
        # the graph gets pinned status words of its OWN (not the shared pool's, which eager nosync forwards recycle)
        old_pool, old_next = RZ._host_pool, RZ._host_next
        RZ._host_pool, RZ._host_next = [], 0
        try:
            RZ.reserve_host_slots(16)
            g = torch.cuda.CUDAGraph()
            lc0 = _capi.launch_count()
            with torch.cuda.graph(g):
                self.body()
            self.launches = _capi.launch_count() - lc0  # kernels of this library one replay launches
            self.watch = list(RZ._pending)              # pinned status words the graph rewrites on every replay
            self._slots = RZ._host_pool                 # keep them alive as long as the graph
        finally:
            RZ._host_pool, RZ._host_next = old_pool, old_next
            RZ._pending.clear()
        self.graph = g
        self.key = self.key_fn() if self.key_fn else None
        self.captures += 1

    def arm(self):
        """Make check_overflow(keep=True) look at THIS graph's status words (several graphs can coexist)."""
        RZ._pending[:] = self.watch

    def __call__(self, check: bool = True):
        if self.graph is None or (self.key_fn is not None and self.key_fn() != self.key):
            self._capture()
        self.arm()
        self.graph.replay()
        if not check:
            return
        try:
            RZ.check_overflow(keep=True)                # the step's only host <-> device synchronisation
        except _capi.SurfelRasterError:
            # a frame outgrew the capacity baked into the graph: the hints were refreshed, capture again and redo the step
            self._capture()
            self.arm()
            self.graph.replay()
            RZ.check_overflow(keep=True)
