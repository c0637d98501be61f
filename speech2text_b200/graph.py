"""CUDA-graph replay of a fixed-shape training step.

The hot path launches ~350 kernels per step, most of them tens of microseconds long: issued
eagerly, the launch gaps and the host-side dispatch cost as much as several of the kernels.
``GraphedStep`` captures one eager execution of a step function (forward + backward through the
drop-in modules; the library never synchronises the host, so the whole step is capturable) and
replays it.  Shapes and tensor addresses are frozen at capture time: feed new data by copying into
the tensors the step function read during capture.
"""
from __future__ import annotations

from typing import Any, Callable, Optional

import torch


class GraphedStep:

    def __init__(self, fn: Callable[[], Any], warmup: int = 3, pool: Optional[Any] = None, priority: int = 0):
        """``priority`` < 0 captures the step on a high-priority stream: kernels that the step forks onto ordinary
        side streams (the simple-loss gradients that start in the forward pass, functional._SimpleLoss) then only
        take SMs the main chain leaves idle instead of competing with it."""
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device")
        side = torch.cuda.Stream(priority=priority)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up off the default stream, as graph capture requires
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # priority 0: torch's own capture stream (several graphs of one process then share it)
        with torch.cuda.graph(self.graph, pool=pool, **({"stream": side} if priority != 0 else {})):
            self.outputs = fn()

    def pool(self):
        return self.graph.pool()

    def __call__(self):
        self.graph.replay()
        return self.outputs
