"""CUDA-graph replay of the whole-model forward (batch-1 / small-batch latency).

A forward is ~215 dependent kernel launches; at batch 64 the GPU is busy for 9 ms and the launches hide behind it, at batch 1 the
host-side launch path (descriptor encoding + 215 cudaLaunchKernelEx) is longer than the device work.  `GraphedForward` captures
one `sunet_forward` call - the same kernels, the same programmatic-dependent-launch edges - on fixed input / output / workspace
buffers and replays it with a single launch.  The reference's counterpart is the per-image loop of demo.py:61-77.
"""
import torch


class GraphedForward:
    """``g = GraphedForward(model, batch, in_chans=3); y = g(x)`` with x (batch, in_chans, img, img) fp32 CUDA.

    ``model`` is a ``SUNet_model`` / ``SUNet``; its parameters must not change while the graph is alive (the graph holds
    the device pack that was current at capture time: rebuild it after ``load_state_dict``)."""

    def __init__(self, model, batch, in_chans=3, device=None):
        net = getattr(model, "swin_unet", model)
        p = next(net.parameters())
        self.device = torch.device(device) if device is not None else p.device
        if self.device.type != "cuda":
            raise RuntimeError("GraphedForward: the model must live on a CUDA device (there is no CPU path)")
        self.model = model
        size = net.img_size
        self.x = torch.zeros(batch, in_chans, size, size, device=self.device, dtype=torch.float32)
        self.out = torch.empty(batch, net.out_chans, size, size, device=self.device, dtype=torch.float32)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):           # warm-up: pre-pack, workspace, per-device kernel attributes - none of it may run under capture
            for _ in range(2):
                model(self.x, out=self.out)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            model(self.x, out=self.out)

    @torch.no_grad()
    def __call__(self, x, copy_out=True):
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out.clone() if copy_out else self.out
