"""one profiled BiSeNet-R18 training step (B=16, 1024^2) bracketed by cudaProfilerStart/Stop:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/profile_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
bench.BATCH_PER_GPU = int(os.environ.get("B", 16))
model, ddp, opt, lrp = bench.build_b200(dev, 1)
imgs, gts = bench.synth_batch(bench.BATCH_PER_GPU, bench.H, bench.W, 100)
imgs, gts = imgs.to(dev), gts.to(dev)
for it in range(int(os.environ.get("WARM", 2))):
    bench.train_step(model, ddp, opt, lrp, it, imgs, gts)
torch.cuda.synchronize()
torch.cuda.profiler.start()
bench.train_step(model, ddp, opt, lrp, 3, imgs, gts)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
