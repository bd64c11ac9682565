"""Small-batch latency of the whole forward (the GPU analogue of BASELINE config 1: one 256x256 image): eager launches vs one
CUDA-graph replay (sunet_tf_b200.graph.GraphedForward), B = 1 / 4 / 16 / 64, device time per forward (CUDA events, median of
`reps`) and host wall time per call.  Prints one JSON document.

  python tools/bench_latency.py > profiles/r04_latency.json"""
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunet_tf_b200 import SUNet_model  # noqa: E402
from sunet_tf_b200.default_config import DEFAULT_OPT  # noqa: E402
from sunet_tf_b200.graph import GraphedForward  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SUNet_model(DEFAULT_OPT).to(dev).eval()
reps = 30
rows = []
for B in (1, 4, 16, 64):
    x = torch.rand(B, 3, 256, 256, device=dev)
    out = torch.empty(B, 1, 256, 256, device=dev)

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dev_ms, wall_ms = [], []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            wall_ms.append((time.perf_counter() - t0) * 1e3)
            dev_ms.append(e0.elapsed_time(e1))
        return statistics.median(dev_ms), statistics.median(wall_ms)

    eager = timed(lambda: model(x, out=out))
    g = GraphedForward(model, B)
    ref = model(x).clone()
    got = g(x)
    assert torch.equal(ref, got), "graph replay differs from the eager forward"
    graph = timed(lambda: g(x, copy_out=False))
    rows.append({"batch": B, "launches": model.swin_unet.launches_per_forward(B), "eager_ms": eager[0], "eager_wall_ms": eager[1],
                 "graph_ms": graph[0], "graph_wall_ms": graph[1], "images_per_s_graph": B / graph[1] * 1e3})
    del g
print(json.dumps({"what": "SUNet 256x256 forward latency, one B200, eager launches vs CUDA-graph replay (bit-identical outputs)",
                  "timing": f"median of {reps}; *_ms = CUDA events around the call, *_wall_ms = host wall clock incl. final synchronize",
                  "rows": rows}, indent=1))
