"""CUDA-graph replay of a fixed-shape forward + backward.

A BCL step is ~20 short kernels (splat, split, contraction, slice and their backwards); at one cloud per call the
launch path, not the GPU, sets the latency.  ``GraphedStep`` captures one call of ``fn`` -- every kernel is enqueued
on the capturing stream through the C ABI, allocations come from the graph's private pool -- and replays it with a
single launch.  Inputs are the tensors ``fn`` closes over: update them in place (``copy_``) between replays.
Requirements: no host synchronisation inside ``fn`` (tile plans must exist already: ``plans.prepare``), fixed shapes.
"""
import torch


class GraphedStep:
    def __init__(self, fn, warmup=3):
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                   # warm-up off the capture: lazy initialisations, plan / image caches
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def replay(self):
        self.graph.replay()
        return self.out
