"""Experiment: one B=64 forward vs two concurrent B=32 forwards on two streams (tail-filling between independent half-batches)."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sunet_tf_b200 import SUNet_model
from sunet_tf_b200.default_config import DEFAULT_OPT
from oracle import weights as Wt

dev = torch.device("cuda:0")
sd = Wt.synth_state_dict(Wt.sunet_spec(), seed=0, style="init")
models = []
for _ in range(4):
    m = SUNet_model(DEFAULT_OPT)
    m.load_state_dict(sd)
    models.append(m.to(dev).eval())
xs = [torch.rand(64, 3, 256, 256, device=dev) for _ in range(4)]
outs = [torch.empty(64, 1, 256, 256, device=dev) for _ in range(2)]

def timeit(fn, steps=20, warm=5):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

def single(i):
    models[0](xs[i % 4], out=outs[0])

ref = models[0](xs[0]).clone()
print("single B=64: %.3f ms" % timeit(single))

for nsplit in (2, 4):
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    part = 64 // nsplit
    def multi(i):
        x = xs[i % 4]
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        for k, s in enumerate(streams):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                models[k](x[k * part:(k + 1) * part], out=outs[1][k * part:(k + 1) * part])
            e = torch.cuda.Event(); e.record(s); cur.wait_event(e)
    multi(0); torch.cuda.synchronize()
    print("split %d: max diff vs single %.3e" % (nsplit, (outs[1] - ref).abs().max().item()))
    print("%d streams x B=%d: %.3f ms" % (nsplit, part, timeit(multi)))
