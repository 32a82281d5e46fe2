"""CUDA-graph capture of whole optimisation iterations (forward + backward + Adam): "streams and graphs instead of a
tracing compiler".  An iteration of the inversion loop launches several thousand small kernels; replaying it as one graph
removes the CPU launch path entirely.

`GraphedStep(fn, state_tensors)` warms `fn` up on a side stream, restores the listed state tensors (parameters, Adam
moments) so the warm-up iterations leave no trace, then captures one call.  `fn` must be free of host synchronisation and
must read every per-step scalar from device tensors.
"""
import torch


class GraphedStep:
    def __init__(self, fn, state_tensors=(), warmup=2):
        saved = [t.detach().clone() for t in state_tensors]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()
        with torch.no_grad():
            for t, s in zip(state_tensors, saved):
                t.copy_(s)

    def __call__(self):
        self.graph.replay()
        return self.out
