"""Which condition invalidates the whole-step capture? Each variant runs in its own process (a failed capture poisons
the process): python tools/diag_graph.py            (driver)   /   python tools/diag_graph.py VARIANT   (one variant)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    # name: (batch, size, eager steps before, capture_error_mode, side stream for everything[, extras])
    "eager3_b16": (16, 1024, 3, "global", True),
    "eager3_b16_pinned": (16, 1024, 3, "global", True, "pinned"),
    "eager3_b16_loss_alive": (16, 1024, 3, "global", True, "loss"),
    "eager3_b16_nosync": (16, 1024, 3, "global", True, "nosync"),
    "eager3_b16_all": (16, 1024, 3, "global", True, "pinned,loss,nosync"),
}


def one(name):
    import torch
    import bench
    from torchseg_b200.engine.graph import GraphedTrainStep
    B, size, eager, mode, side = VARIANTS[name][:5]
    extras = VARIANTS[name][5].split(",") if len(VARIANTS[name]) > 5 else []
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    if side:
        torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    bench.BATCH_PER_GPU = B
    bench.H = bench.W = size
    model, ddp, opt, lrp = bench.build_b200(dev, 1)
    host = bench.synth_batch(B, size, size, 100, pin="pinned" in extras)
    batch = tuple(t.to(dev) for t in host)
    for it in range(eager):
        loss = bench.train_step(model, ddp, opt, lrp, it, *batch)
        if "loss" not in extras:
            del loss
    if "nosync" not in extras:
        torch.cuda.synchronize()
    g = GraphedTrainStep(model, opt, batch, warmup=2, capture_error_mode=mode)
    if g.graph is None:
        print("RESULT %s FAIL %s" % (name, (g.error or "").replace("\n", " | ")[-700:]))
        return
    for _ in range(3):
        loss = g(*batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        loss = g(*batch)
    e1.record()
    torch.cuda.synchronize()
    print("RESULT %s OK launches=%d ms/step=%.3f loss=%.4f" % (name, g.launches_per_step, e0.elapsed_time(e1) / 10, loss.item()))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1])
    else:
        for name in VARIANTS:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True, timeout=400)
            lines = [l for l in (r.stdout + r.stderr).splitlines() if l.startswith("RESULT")]
            print(lines[-1] if lines else "RESULT %s CRASH rc=%d %s" % (name, r.returncode, (r.stderr or "")[-300:].replace("\n", " | ")))
            sys.stdout.flush()
