"""micro-benchmark of single conv launches (CUDA events, L2 flushed between iterations)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchseg_b200 import ops, _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

def case(N, C, H, W, K, R, st, pad, dil, tag=""):
    x = ops.nhwc_empty(N, C, H, W); x.normal_()
    wb = torch.randn(K, R, R, C, device=dev).to(torch.bfloat16)
    wt = torch.randn(C, R, R, K, device=dev).to(torch.bfloat16)
    P, Q = ops.conv_out_size(H, R, st, pad, dil), ops.conv_out_size(W, R, st, pad, dil)
    y = ops.nhwc_empty(N, K, P, Q); gy = ops.nhwc_empty(N, K, P, Q); gy.normal_()
    stats = torch.zeros(2, K, device=dev)
    dw = torch.zeros(K, R, R, C, device=dev)
    dx = ops.nhwc_empty(N, C, H, W)
    fl = 2.0 * N * P * Q * K * C * R * R
    t1 = timeit(lambda: ops.conv_fprop(x, wb, K, R, st, pad, dil, out=y, stats=stats))
    t2 = timeit(lambda: ops.conv_fprop(x, wb, K, R, st, pad, dil, out=y))
    t3 = timeit(lambda: ops.conv_dgrad(gy, wt, (N, C, H, W), K, R, st, pad, dil, out=dx))
    t4 = timeit(lambda: ops.conv_wgrad(x, gy, K, R, st, pad, dil, dw))
    print("%-28s fprop+stats %7.1f us (%5.0f TF/s) | fprop %7.1f (%5.0f) | dgrad %7.1f (%5.0f) | wgrad %7.1f (%5.0f)" % (
        tag or str((N, C, H, W, K, R, st)), t1, fl / t1 / 1e6, t2, fl / t2 / 1e6, t3, fl / t3 / 1e6, t4, fl / t4 / 1e6))

if __name__ == "__main__":
    for k, v in [a.split("=") for a in sys.argv[1:] if "=" in a]:
        _lib.call("tsb_debug_set", int(k), int(v))
    case(16, 64, 256, 256, 64, 3, 1, 1, 1, "layer1 64->64 @256")
    case(16, 64, 512, 512, 64, 3, 2, 1, 1, "spatial 64->64 s2 @512")
    case(16, 128, 128, 128, 128, 3, 1, 1, 1, "layer2 128->128 @128")
    case(16, 256, 64, 64, 256, 3, 1, 1, 1, "layer3 256->256 @64")
    case(16, 512, 32, 32, 512, 3, 1, 1, 1, "layer4 512->512 @32")
    case(16, 128, 128, 128, 256, 3, 1, 1, 1, "head1 128->256 @128")
    case(16, 256, 128, 128, 256, 1, 1, 0, 1, "ffm 1x1 256->256 @128")
    case(16, 64, 256, 256, 128, 3, 2, 1, 1, "layer2.0 64->128 s2")
