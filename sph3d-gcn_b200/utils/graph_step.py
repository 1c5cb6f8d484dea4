"""One training step as a CUDA graph.

The reference builds a static TensorFlow graph once and `sess.run`s it every step (s3dis_seg/train_s3dis.py:196-260):
no per-op host work inside a step.  Eager PyTorch pays ~10-25 us of host time per op, and an SPH3D step is 500-600
small kernels (measured: 14.2 ms wall for 8.8 ms of kernels on the ModelNet shape), so the step is host-bound.
Every entry point of libsph3d_b200 is stream-ordered, allocation-free and sync-free (include/sph3d_b200.h), hence
the whole step -- graph construction, FPS on its side stream, convolutions, loss, backward -- can be captured once
and replayed: the static-graph execution model of the reference, on CUDA graphs instead of a tracing compiler.

    step = GraphedStep(fn, parameters)   # fn(): forward + loss + backward on tensors it closes over; returns tensor(s)
    out = step()                         # replay; `out` and every parameter's .grad are rewritten in place
Inputs are static tensors: copy new data into them (tensor.copy_) before calling step().  `parameters` (a list or a
callable returning one, evaluated after the warm-up steps have created the variables) lets the step re-attach the
captured gradient buffers to `.grad` on every replay, so that code which set `.grad = None` in between (an eager
step, an optimizer's zero_grad) cannot detach them.
"""
import torch


class GraphedStep(object):
    def __init__(self, fn, parameters=None, warmup=3, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device (sph3d-gcn_b200 has no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.fn = fn
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):                  # variables, cuBLAS handles, kernel attributes: outside capture
                for _ in range(max(int(warmup), 1)):
                    fn()
            cur.wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outputs = fn()
        params = parameters() if callable(parameters) else (parameters or [])
        self.grads = [(p, p.grad) for p in params if p.grad is not None]
        self.replays = 0

    def __call__(self):
        self.graph.replay()
        for p, g in self.grads:
            p.grad = g
        self.replays += 1
        return self.outputs
