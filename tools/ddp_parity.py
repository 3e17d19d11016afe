"""2-GPU check of the DDP / SyncBN replacements: after one step on different per-rank batches, both ranks hold
identical parameters, and the SyncBN global-batch statistics make the 2x8-image step equal (to bf16 noise) to a
single-process step on the concatenated 16-image batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torchseg_b200
from torchseg_b200 import optim, ops
from torchseg_b200.apex.parallel import DistributedDataParallel, SyncBatchNorm
from torchseg_b200.networks import BiSeNet
from torchseg_b200.seg_opr.loss_opr import ProbOhemCrossEntropy2d
from torchseg_b200.utils.init_func import init_weight, group_weight

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", init_method="env://", device_id=dev)
N, HW = 8, 128
def build(norm, n_imgs):
    torch.manual_seed(0)
    crit = ProbOhemCrossEntropy2d(255, thresh=0.7, min_kept=n_imgs * HW * HW // 16)
    m = BiSeNet(19, True, crit, None, norm)
    init_weight(m.business_layer, torch.nn.init.kaiming_normal_, norm, 1e-5, 0.1, mode='fan_in', nonlinearity='relu')
    m.to(dev); torchseg_b200.prepare_model(m); m.train()
    opt = optim.SGD(group_weight([], m, norm, 1e-2), lr=1e-2, momentum=0.9, weight_decay=5e-4)
    return m, opt
g = torch.Generator().manual_seed(5)
X = torch.randn(world * N, 3, HW, HW, generator=g); Y = torch.randint(0, 19, (world * N, HW, HW), generator=g); Y[:, :12] = 255
m, opt = build(SyncBatchNorm, N)
ddp = DistributedDataParallel(m)
opt.zero_grad()
loss = ddp(X[rank * N:(rank + 1) * N].to(dev), Y[rank * N:(rank + 1) * N].to(dev))
loss.backward(); ddp.finish_reduce(); opt.step(); torch.cuda.synchronize()
flat = opt.flat_param.clone()
other = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(other, flat)
same = all(torch.equal(o, other[0]) for o in other)
gl = loss.detach().clone(); dist.all_reduce(gl); gl /= world
if rank == 0:
    ops.set_sync_group(None, 1)
    ops.grad_ready_hook = None
    m1, o1 = build(torch.nn.BatchNorm2d, world * N)
    o1.zero_grad()
    l1 = m1(X.to(dev), Y.to(dev)); l1.backward(); o1.step(); torch.cuda.synchronize()
    d = (o1.flat_param - flat).norm() / (o1.flat_param - 0).norm()
    # update direction agreement
    print("ranks identical:", same, "| mean loss 2x8:", float(gl), "single 16:", float(l1), "| rel param diff after step %.3e" % float(d))
    assert same
    assert abs(float(gl) - float(l1)) < 2e-2 * abs(float(l1))
    print("OK")
dist.barrier(); dist.destroy_process_group()
