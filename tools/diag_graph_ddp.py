"""2-GPU smoke of the captured multi-GPU step on a SMALL problem (fast, bounded): eager steps → capture → replays.
torchrun --nproc-per-node 2 tools/diag_graph_ddp.py [overlap]"""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(70, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", init_method="env://", device_id=dev)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
B, size = int(os.environ.get("B", 4)), int(os.environ.get("SIZE", 256))
bench.BATCH_PER_GPU = B; bench.H = bench.W = size
model, ddp, opt, lrp = bench.build_b200(dev, world)
batch = tuple(t.to(dev) for t in bench.synth_batch(B, size, size, 100 + rank))
def say(msg):
    print("[rank %d] %s" % (rank, msg)); sys.stdout.flush()
for it in range(3):
    bench.train_step(model, ddp, opt, lrp, it, *batch)
torch.cuda.synchronize(); say("eager steps ok")
from torchseg_b200.engine.graph import GraphedTrainStep
g = GraphedTrainStep(model, opt, batch, warmup=2, ddp=ddp)
if "overlap" in sys.argv:
    ddp.overlap = True
torch.cuda.synchronize(); say("capture %s" % ("ok, %d launches" % g.launches_per_step if g.graph is not None else "FAILED: " + (g.error or "")[-400:]))
if g.graph is not None:
    for k in range(5):
        loss = g(*batch)
        torch.cuda.synchronize(); say("replay %d loss %.4f" % (k, loss.item()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g(*batch)
    e1.record(); torch.cuda.synchronize()
    say("graph %.3f ms/step" % (e0.elapsed_time(e1) / 10))
    flat = opt.flat_param.clone(); other = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(other, flat)
    say("ranks identical after replays: %s" % all(torch.equal(o, other[0]) for o in other))
g.release()
dist.barrier(); dist.destroy_process_group(); say("done")
