"""find the call that invalidates the whole-step CUDA graph capture at bench sizes"""
import ctypes, os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torchseg_b200 import _lib, ops

rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
print("cudart:", rt)

def cap_status():
    st = ctypes.c_int(-1)
    rc = rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return rc, st.value

log = []
orig_call = _lib.call
def traced(name, *a):
    r = orig_call(name, *a)
    if rt is not None:
        s = cap_status()
        log.append((name, s))
    return r
_lib.call = traced
ops.call = traced
import torchseg_b200.optim as optim_mod, torchseg_b200.seg_opr.seg_oprs as so
optim_mod.call = traced

dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
bench.BATCH_PER_GPU = int(os.environ.get("B", 16))
if os.environ.get("HW"):
    bench.H = bench.W = int(os.environ["HW"])
model, ddp, opt, lrp = bench.build_b200(dev, 1)
batch = bench.synth_batch(bench.BATCH_PER_GPU, bench.H, bench.W, 100)
batch = [t.to(dev) for t in batch]
for it in range(3):
    bench.train_step(model, ddp, opt, lrp, it, *batch)
torch.cuda.synchronize()
from torchseg_b200.engine.graph import GraphedTrainStep
log.clear()
mode = os.environ.get("MODE", "global")
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, capture_error_mode=mode):
        opt.zero_grad(); loss = model(*batch); loss.backward(); opt.step()
    print("capture OK in mode", mode, "calls", len(log))
    g.replay(); torch.cuda.synchronize(); print("replay OK loss", float(loss))
except Exception as e:
    print("capture FAILED in mode", mode, type(e).__name__, str(e)[:200])
    bad = [i for i, (n, s) in enumerate(log) if s[1] != 1 or s[0] != 0]
    print("calls logged", len(log), "first non-active status at", bad[:1])
    if bad:
        i = bad[0]
        for j in range(max(0, i - 3), min(len(log), i + 2)):
            print("   ", j, log[j])
