"""GraphedTrainStep flow with per-call capture-status tracing and the three capture_error_modes"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torchseg_b200 import _lib, ops
import torchseg_b200.optim as optim_mod
from torchseg_b200.engine.graph import GraphedTrainStep
rt = ctypes.CDLL("libcudart.so.12")
def cap_status():
    st = ctypes.c_int(-1)
    rc = rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return rc, st.value
log = []
orig = _lib.call
def traced(name, *a):
    s0 = cap_status()
    r = orig(name, *a)
    log.append((name, s0, cap_status()))
    return r
_lib.call = traced; ops.call = traced; optim_mod.call = traced
import torchseg_b200.seg_opr.seg_oprs as so, torchseg_b200.seg_opr.loss_opr as lo
mode = os.environ.get("MODE", "global")
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
bench.BATCH_PER_GPU = 16
model, ddp, opt, lrp = bench.build_b200(dev, 1)
host = bench.synth_batch(16, bench.H, bench.W, 100, pin=True)
batch = tuple(t.to(dev) for t in host)
for it in range(3):
    bench.train_step(model, ddp, opt, lrp, it, *batch)
torch.cuda.synchronize()
log.clear()
gs = GraphedTrainStep(model, opt, batch, warmup=int(os.environ.get("WU", 2)), capture_error_mode=mode)
print("MODE", mode, "graph:", gs.graph is not None, "err:", (gs.error or "")[:120].replace("\n", " "))
# calls made while capturing (status 1 = active, 2 = invalidated)
cap = [(i, l) for i, l in enumerate(log) if l[1][1] != 0 or l[2][1] != 0]
print("calls during capture:", len(cap))
bad = [(i, l) for i, l in cap if l[2][1] == 2 or l[2][0] != 0 or l[1][1] == 2]
if bad:
    i = bad[0][0]
    for j in range(max(0, i - 4), min(len(log), i + 2)):
        print("   ", j, log[j])
