"""Replays a whole training step as ONE CUDA graph.

A step of these models is thousands of short kernels (CycleGAN ~3200, seg/depth ~13000 launches); issuing them
one by one leaves the GPU idle between launches.  ``StepGraph.run(step_fn)`` executes ``step_fn`` eagerly for a
few warm-up calls (on a side stream: autograd binds each parameter's gradient accumulator to the stream that
was current when it was created, and a node bound to the legacy default stream cannot take part in a capture),
then captures one call with ``torch.cuda.graph`` — forward, hand-written backward, losses, optimizer updates —
and from then on replays the captured graph.

Requirements on ``step_fn`` (met by the step mirrors of this package): static input tensors (copy new data into
them), no host synchronisation, no host-side decisions that change between steps (the ImagePool's random draws
are made up-front and reach the device through a pinned table; Adam reads its step count and learning rate from
device memory (``FusedAdam.sync_lr``); dropout reads its seed from a device counter that a node of the graph bumps,
``ops.dropout_seed``).
"""
import torch

from . import _lib


class StepGraph:
    WARMUP_STEPS = 3

    def __init__(self):
        self.graph = None
        self.calls = 0
        self.launches = 0          # library kernels inside one replay
        self._side = None
        self._done = None          # event recorded after the most recent replay

    def wait_previous(self):
        """Blocks the host until the previous replay has finished (call before overwriting pinned host memory
        that the graph's copy nodes read)."""
        if self._done is not None:
            self._done.synchronize()

    def run(self, step_fn, optimizers=()):
        """optimizers: FusedAdam instances stepped inside step_fn; their learning rates are copied to the device
        scalars the captured kernels read (FusedAdam.sync_lr) before every call, so scheduler updates reach replays."""
        for o in optimizers:
            if hasattr(o, 'sync_lr'):
                o.sync_lr()
        if self.graph is not None:
            self.graph.replay()
            self._done.record()
            return
        self.calls += 1
        if self.calls <= self.WARMUP_STEPS:
            if self._side is None:
                self._side = torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                step_fn()
            cur.wait_stream(self._side)
            return
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        n0 = _lib.lib().cdb_launch_count()
        # thread_local: other threads (NCCL's watchdog under data parallelism) may keep calling CUDA APIs that
        # the default global capture mode forbids while this thread captures
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            step_fn()
        self.launches = _lib.lib().cdb_launch_count() - n0
        self.graph = graph
        graph.replay()             # capture records the step, the replay performs it
        self._done = torch.cuda.Event()
        self._done.record()
