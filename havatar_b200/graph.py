"""CUDA-graph capture of a fixed-shape forward (launch-bound inner loops: the StyleUNet forward is ~250 small launches).

    g = GraphedForward(lambda style, cond: net([style], cond, noise=noise), style, cond)
    img = g(style, cond)          # copies into the static inputs, replays the graph, returns the static output(s)

The hav_* C-ABI launches go to torch's current stream, so they are captured like any torch op; weight packing and the
NoiseInjection scalars are cached by the modules during the warm-up calls, which run before capture."""
import torch


class GraphedForward:
    def __init__(self, fn, *example_inputs, warmup=3):
        self.fn = fn
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self.fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import conv

        conv.CAPTURE_GEN[0] += 1
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = self.fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
