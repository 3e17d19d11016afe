import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchseg_b200
from torchseg_b200.networks import PSPNet
from oracle import torch_ref
cuda = torch.device("cuda:0")
torch.manual_seed(2)
N, HW = int(os.environ.get("DN", 8)), int(os.environ.get("DH", 96))
m = PSPNet(150, torch.nn.CrossEntropyLoss(ignore_index=-1))
for mod in m.modules():
    if isinstance(mod, torch.nn.Dropout2d): mod.p = 0.0
def mk():
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k: v.requires_grad_(True)
    return sd
g = torch.Generator().manual_seed(3)
x = torch.randn(N, 3, HW, HW, generator=g); y = torch.randint(-1, 150, (N, HW, HW), generator=g)
sd32 = mk(); l32, _ = torch_ref.pspnet_loss(x, y, sd32); l32.backward()
torch_ref.set_bf16_emulation(True); sde = mk(); le, _ = torch_ref.pspnet_loss(x, y, sde); le.backward(); torch_ref.set_bf16_emulation(False)
m.to(cuda); torchseg_b200.prepare_model(m); m.train()
loss = m(x.to(cuda), y.to(cuda)); loss.backward(); torch.cuda.synchronize()
print("loss dev %.5f fp32 %.5f emu %.5f" % (loss.item(), float(l32), float(le)))
def cos(a, b):
    a, b = a.float().cpu().reshape(-1), b.reshape(-1)
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))
names = ["psp_layer.conv6.2.weight", "psp_layer.conv6.0.conv.weight", "psp_layer.ppm.0.psp/cbr.conv.weight", "psp_layer.ppm.3.psp/cbr.conv.weight",
         "aux_layer.2.weight", "aux_layer.0.conv.weight", "backbone.layer4.2.conv3.weight", "backbone.layer4.2.conv1.weight", "backbone.layer4.1.conv2.weight",
         "backbone.layer4.0.conv3.weight", "backbone.layer4.0.conv1.weight", "backbone.layer3.22.conv3.weight", "backbone.layer3.10.conv2.weight",
         "backbone.layer3.0.conv1.weight", "backbone.layer2.0.conv1.weight", "backbone.layer1.0.conv1.weight", "backbone.conv1.6.weight", "backbone.conv1.0.weight"]
P = dict(m.named_parameters())
for n in names:
    print("%-45s cos(dev,fp32) %.3f  cos(dev,emu) %.3f  cos(emu,fp32) %.3f  |g| dev %.3e fp32 %.3e" % (
        n, cos(P[n].grad, sd32[n].grad), cos(P[n].grad, sde[n].grad), cos(sde[n].grad, sd32[n].grad), float(P[n].grad.norm()), float(sd32[n].grad.norm())))
