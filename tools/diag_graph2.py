"""which bench-specific state breaks the capture? VAR=a: conv_prof pass + GraphedTrainStep(side-stream warm-up);
b: no conv_prof pass, side-stream warm-up; c: conv_prof pass, warm-up on the current stream"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torchseg_b200 import ops
from torchseg_b200.engine.graph import GraphedTrainStep
var = os.environ.get("VAR", "a")
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
bench.BATCH_PER_GPU = 16
model, ddp, opt, lrp = bench.build_b200(dev, 1)
host = bench.synth_batch(16, bench.H, bench.W, 100, pin=True)
batch = tuple(t.to(dev) for t in host)
for it in range(3):
    bench.train_step(model, ddp, opt, lrp, it, *batch)
if var in ("a", "c"):
    ops.conv_prof.enable()
    with bench.ClockSampler(0) as clk:
        for it in range(3):
            loss = bench.train_step(model, ddp, opt, lrp, it, *batch)
    prof = ops.conv_prof.collect(); ops.conv_prof.disable()
    print("prof", prof["launches"], float(loss.item()))
if var == "d":      # sampler only
    with bench.ClockSampler(0) as clk:
        for it in range(3):
            loss = bench.train_step(model, ddp, opt, lrp, it, *batch)
    print("sampler only", float(loss.item()))
if var == "e":      # timing events only
    ops.conv_prof.enable()
    for it in range(3):
        loss = bench.train_step(model, ddp, opt, lrp, it, *batch)
    prof = ops.conv_prof.collect(); ops.conv_prof.disable()
    print("events only", prof["launches"])
if var == "f":      # timing events, then dropped before capture
    ops.conv_prof.enable()
    for it in range(3):
        loss = bench.train_step(model, ddp, opt, lrp, it, *batch)
    prof = ops.conv_prof.collect(); ops.conv_prof.disable(); ops.conv_prof.recs = []
    import gc; gc.collect()
    print("events dropped", prof["launches"])
gs = GraphedTrainStep(model, opt, batch, warmup=2, side_stream_warmup=(var in ("a", "b")))
print("VAR", var, "graph:", gs.graph is not None, "err:", (gs.error or "")[:150].replace("\n", " "))
if gs.graph is not None:
    l = gs(*gs.static_inputs); torch.cuda.synchronize(); print("replay loss", float(l))
