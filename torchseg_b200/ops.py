"""torch.autograd.Function wrappers around the libtsb C ABI (include/tsb.h).

PyTorch drives only the graph, the memory and the streams; every op below runs a hand-written sm_100a
kernel. Activations are bf16 tensors of LOGICAL shape [N,C,H,W] stored NHWC (`nhwc_empty`), possibly a
channel slice of a wider buffer (channel stride `cs`). There is no CPU / ATen fallback: without a GPU or
without libtsb.so these functions raise.
"""
import torch

from . import _lib
from ._lib import BF16, F32, ConvShape, call, ptr, stream

import ctypes
import weakref

_BF = torch.bfloat16


# --------------------------------------------------------------------------------------------------
# tensor helpers
# --------------------------------------------------------------------------------------------------
def nhwc_empty(N, C, H, W, dtype=_BF, device="cuda", cs=None):
    """logical [N,C,H,W] view of a fresh NHWC buffer (channel stride cs >= C)"""
    cs = C if cs is None else cs
    buf = torch.empty((N, H, W, cs), dtype=dtype, device=device)
    return buf[..., :C].permute(0, 3, 1, 2)


def nhwc_zeros(N, C, H, W, dtype=_BF, device="cuda", cs=None):
    cs = C if cs is None else cs
    buf = torch.zeros((N, H, W, cs), dtype=dtype, device=device)
    return buf[..., :C].permute(0, 3, 1, 2)


def cs_of(t):
    """channel stride (elements between consecutive pixels) of an NHWC-backed logical NCHW tensor"""
    assert t.dim() == 4, "expected a 4-D tensor"
    N, C, H, W = t.shape
    s = t.stride()
    if C > 1 and s[1] != 1:
        raise ValueError("tensor is not NHWC-backed (channel stride %d); use to_nhwc()" % s[1])
    if W > 1:
        cs = s[3]
    elif H > 1:
        cs = s[2]
    elif N > 1:
        cs = s[0]
    else:
        cs = C
    if (W > 1 and H > 1 and s[2] != W * cs) or (H * W > 1 and N > 1 and s[0] != H * W * cs) or cs < C:
        raise ValueError("tensor is not a dense NHWC buffer: shape %s strides %s" % (tuple(t.shape), s))
    return cs


def to_nhwc(t, dtype=_BF):
    """any [N,C,H,W] tensor → NHWC-backed logical NCHW tensor of `dtype` (ATen copy; boundary only)"""
    N, C, H, W = t.shape
    out = nhwc_empty(N, C, H, W, dtype=dtype, device=t.device)
    out.copy_(t)
    return out


def _krsc_ptr(w):
    """pointer to a dense [K,R,S,C] fp32 buffer for conv weight w ([K,C,R,S] logical)"""
    wp = w.permute(0, 2, 3, 1)
    if not wp.is_contiguous():
        raise ValueError("conv weight must be channels_last (KRSC); call torchseg_b200.prepare_model(model)")
    return wp


def conv_out_size(H, k, stride, pad, dil):
    return (H + 2 * pad - dil * (k - 1) - 1) // stride + 1


def make_shape(N, H, W, C, K, R, stride, pad, dil):
    P = conv_out_size(H, R, stride, pad, dil)
    Q = conv_out_size(W, R, stride, pad, dil)
    return ConvShape(N, H, W, C, K, R, R, stride, pad, dil, P, Q)


# --------------------------------------------------------------------------------------------------
# bf16 weight packs, cached per optimiser step
# --------------------------------------------------------------------------------------------------
class _PackCache:
    """bf16 operand packs of conv weights that are not served by the flat mirror (optim.FlatPack): weights before the
    first fused optimiser step, torch.optim.* flows, padded / viewed weights, stems. At most ONE entry per
    (weight object, operand kind); an entry is overwritten when the weight's version changes and dropped when the
    weight is gone (DFN's per-forward padded temporaries), so the cache cannot grow with the step count."""

    def __init__(self):
        self.step = 0
        self.cache = {}
        self._sweep_at = 256

    def invalidate(self):
        self.step += 1
        self.cache.clear()
        zero_arena.reset()

    def _sweep(self):
        dead = [k for k, v in self.cache.items() if v[2]() is None]
        for k in dead:
            del self.cache[k]
        self._sweep_at = max(256, 2 * len(self.cache))

    def get(self, w, want_t, pad_k=None, stem=False):
        fp = getattr(w, "_tsb_pack", None)
        if fp is not None and not stem and pad_k is None:
            hit = fp[0].lookup(w, fp[1], want_t)   # bf16 mirror of the flat parameter buffer (optim.FlatPack)
            if hit is not None:
                return hit
        key = (id(w), pad_k, stem)
        ver = (w.data_ptr(), w._version)
        hit = self.cache.get(key)
        if hit is not None and hit[2]() is w and hit[3] == ver and (hit[1] is not None or not want_t or stem):
            return hit[0], hit[1]
        K, C, R, S = w.shape
        if stem:
            wp = torch.empty((K, 4, 4, 16), dtype=_BF, device=w.device)
            call("tsb_pack_stem_weight", ptr(_krsc_ptr(w)), K, R, ptr(wp), stream())
            out = (wp, None, weakref.ref(w), ver)
        else:
            src = _krsc_ptr(w)
            if pad_k is not None and pad_k != K:  # classifier: pad K (19) up to 64 rows of zeros
                padded = torch.zeros((pad_k, R, S, C), dtype=torch.float32, device=w.device)
                padded[:K].copy_(src)
                src, K = padded, pad_k
            wb = torch.empty((K, R, S, C), dtype=_BF, device=w.device)
            wt = torch.empty((C, R, S, K), dtype=_BF, device=w.device) if want_t else None
            call("tsb_pack_weight", ptr(src), K, R, S, C, ptr(wb), ptr(wt), stream())
            out = (wb, wt, weakref.ref(w), ver)
        self.cache[key] = out
        if len(self.cache) > self._sweep_at:
            self._sweep()
        return out[0], out[1]


pack_cache = _PackCache()


def _install_optimizer_hook():
    """torch.optim.* flows (INTEGRATION.md keeps them supported): every optimiser step changes the weights, so the
    per-tensor packs are dropped and the zero arena is recycled, exactly as the fused SGD does itself"""
    try:
        from torch.optim.optimizer import register_optimizer_step_post_hook
    except ImportError:  # pragma: no cover
        return

    def _post(opt, args, kwargs):
        pack_cache.invalidate()

    register_optimizer_step_post_hook(_post)


class _ZeroArena(object):
    """Small fp32 scratch buffers that must start at zero (BN statistics, reduction accumulators): carved out of ONE
    buffer that is cleared with a single memset per step (`reset`, called from zero_grad / optimizer.step) instead of
    one fill kernel per buffer. Falls back to torch.zeros when exhausted or never reset."""

    def __init__(self, nfloats=1 << 20):
        self.n = nfloats
        self.buf = None
        self.off = 0

    def reset(self):
        if self.buf is not None:
            self.buf.zero_()
        self.off = 0

    def take(self, shape, device):
        n = 1
        for d in shape:
            n *= int(d)
        n4 = (n + 3) // 4 * 4
        if self.buf is None or self.buf.device != device:
            self.buf = torch.zeros(self.n, dtype=torch.float32, device=device)
            self.off = 0
        if self.off + n4 > self.n:
            return torch.zeros(shape, dtype=torch.float32, device=device)
        out = self.buf[self.off:self.off + n].view(shape)
        self.off += n4
        return out


zero_arena = _ZeroArena()
_install_optimizer_hook()

# set by the DDP shim: called with a parameter whose gradient has just been written DIRECTLY into its flat .grad view
grad_ready_hook = None


def _direct_grad(p):
    """True when this parameter's .grad is a persistent flat-buffer view the kernels may accumulate into directly"""
    return getattr(p, "_tsb_direct", False) and p.grad is not None


def _notify(*params):
    if grad_ready_hook is not None:
        for p in params:
            grad_ready_hook(p)

# process group for SyncBN statistics (set by apex_shim.parallel.SyncBatchNorm / DistributedDataParallel)
_sync = {"group": None, "world": 1, "p2p": None, "p2p_tried": False}


def set_sync_group(group, world):
    _sync["group"], _sync["world"] = group, world
    if world <= 1:
        _sync["p2p"], _sync["p2p_tried"] = None, False


class _PeerExchange(object):
    """SyncBN statistics exchange over NVLink peer memory (csrc/p2p.cu, tsb_p2p_allreduce_sum): one symmetric buffer per
    rank (torch symmetric memory supplies the allocation and the peer-mapped addresses — plumbing only), ONE small
    kernel per exchange instead of an NCCL all-reduce. All ranks must issue the same exchanges in the same order."""
    NSLOTS = 64
    SLOT_FLOATS = 4096      # 2 x 2048 channels (ResNet-101 layer4)

    def __init__(self, group, world, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        pg = group if group is not None else dist.group.WORLD
        self.world = world
        self.rank = dist.get_rank(pg)
        nbytes = int(_lib.lib().tsb_p2p_buffer_bytes(world, self.NSLOTS, self.SLOT_FLOATS))
        assert nbytes > 0
        self.buf = symm.empty(nbytes // 4, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, pg.group_name)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(ptrs) == world and all(ptrs)
        self.peers = (ctypes.c_ulonglong * world)(*ptrs)
        # the exchange counter lives on the device (pre-incremented by every exchange kernel): identical on all ranks
        # because all ranks issue the same exchanges, and replayable inside a captured CUDA graph
        self.seq_dev = torch.zeros(1, dtype=torch.int32, device=device)

    def allreduce(self, buf, acc_hi=None, acc_lo=None):
        call("tsb_p2p_allreduce_sum", ptr(buf), buf.numel(), ctypes.cast(self.peers, ctypes.c_void_p), self.rank, self.world,
             0, ptr(self.seq_dev), self.NSLOTS, self.SLOT_FLOATS, ptr(acc_hi), ptr(acc_lo), stream())


def _peer_exchange(device):
    """the NVLink exchange object, or None (single process, CPU / gloo, TSB_SYNCBN_P2P=0, no symmetric memory).
    Collective: every rank of the SyncBN group reaches this at the same point (its first exchange), and the ranks AGREE on
    the outcome — a rank that could not set the exchange up while its peers spin in the kernel would deadlock the job."""
    if _sync["world"] <= 1 or device.type != "cuda":
        return None
    if not _sync["p2p_tried"]:
        _sync["p2p_tried"] = True
        import os
        import sys
        import torch.distributed as dist
        x, why = None, "disabled by TSB_SYNCBN_P2P=0"
        if os.environ.get("TSB_SYNCBN_P2P", "1") != "0":
            try:
                x = _PeerExchange(_sync["group"], _sync["world"], device)
            except Exception as e:  # noqa: BLE001 — no NVLink P2P / symmetric memory on this box: keep NCCL
                x, why = None, "%s: %s" % (type(e).__name__, str(e).split("\n")[0][:200])
        ok = torch.tensor([1.0 if x is not None else 0.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=_sync["group"])
        if float(ok.item()) < 1.0:
            if x is not None:
                why = "another rank could not set it up"
            x = None
            if os.environ.get("TSB_SYNCBN_P2P", "1") != "0":
                sys.stderr.write("torchseg_b200: NVLink SyncBN exchange unavailable (%s); using NCCL all-reduce\n" % why)
        else:
            torch.cuda.synchronize(device)
            dist.barrier(group=_sync["group"])   # every rank's buffer is zeroed before anybody's first exchange writes into it
        _sync["p2p"] = x
    return _sync["p2p"]


def _allreduce_stats(buf, acc_hi=None, acc_lo=None):
    """sum the packed per-channel statistics over the SyncBN group, in place. acc_hi / acc_lo (backward): the LOCAL
    halves are first accumulated into the parameter gradients (dgamma += buf[1], dbeta += buf[0]). Returns True when
    that accumulation was done here."""
    if _sync["world"] <= 1:
        return False
    x = _peer_exchange(buf.device)
    n = buf.numel()
    if x is not None and n % 4 == 0 and n <= x.SLOT_FLOATS and buf.is_contiguous():
        x.allreduce(buf, acc_hi, acc_lo)
        return acc_hi is not None
    import torch.distributed as dist
    if acc_hi is not None:
        acc_hi.add_(buf[1])   # local sums must be taken before the all-reduce
        acc_lo.add_(buf[0])
    dist.all_reduce(buf, group=_sync["group"])
    return acc_hi is not None


# --------------------------------------------------------------------------------------------------
# optional CUDA-event bracketing of the conv launches (bench.py roofline: algorithmic FLOPs / kernel time)
# --------------------------------------------------------------------------------------------------
class _ConvProf(object):
    def __init__(self):
        self.on = False
        self.recs = []

    def enable(self):
        self.on, self.recs = True, []

    def disable(self):
        self.on = False

    def begin(self):
        if not self.on:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, e0, flops):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.recs.append((e0, e1, flops))

    def collect(self):
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b, _ in self.recs)
        return dict(ms=ms, flops=float(sum(f for _, _, f in self.recs)), launches=len(self.recs))


conv_prof = _ConvProf()


def _conv_flops(shp):
    return 2.0 * shp.N * shp.P * shp.Q * shp.K * shp.C * shp.R * shp.S


# --------------------------------------------------------------------------------------------------
# low-level conv calls
# --------------------------------------------------------------------------------------------------
def conv_fprop(x, wb, K, R, stride, pad, dil, bias=None, out_dtype=_BF, out=None, ocs=None, stats=None):
    N, C, H, W = x.shape
    shp = make_shape(N, H, W, C, K, R, stride, pad, dil)
    if out is None:
        ocs_alloc = ocs if ocs is not None else (K + 7) // 8 * 8
        out = nhwc_empty(N, K, shp.P, shp.Q, dtype=out_dtype, device=x.device, cs=ocs_alloc)
    s1 = s2 = None
    if stats is not None:
        s1, s2 = stats[0], stats[1]
    ev = conv_prof.begin()
    call("tsb_conv2d_fprop", ctypes.byref(shp), ptr(x), cs_of(x), ptr(wb), ptr(bias), ptr(out), _lib.dt(out), cs_of(out),
         ptr(s1), ptr(s2), stream())
    conv_prof.end(ev, _conv_flops(shp))
    return out


def conv_fprop_fused(x, wb, K, R, stride, pad, dil, bias, res=None, relu=False):
    """inference conv with the BatchNorm folded into (wb, bias): y = relu?(conv(x) + bias (+ res)) in ONE kernel
    (tsb_conv2d_fprop_fused) — no raw tensor, no separate normalisation pass"""
    N, C, H, W = x.shape
    shp = make_shape(N, H, W, C, K, R, stride, pad, dil)
    out = nhwc_empty(N, K, shp.P, shp.Q, device=x.device, cs=(K + 7) // 8 * 8)
    if res is not None and (res.dtype != _BF or res.stride(1) != 1 or cs_of(res) != cs_of(out)):
        res = to_nhwc(res) if cs_of(out) == K else None
        assert res is not None, "residual layout must match the output"
    ev = conv_prof.begin()
    call("tsb_conv2d_fprop_fused", ctypes.byref(shp), ptr(x), cs_of(x), ptr(wb), ptr(bias), ptr(res),
         cs_of(res) if res is not None else 0, int(relu), ptr(out), BF16, cs_of(out), stream())
    conv_prof.end(ev, _conv_flops(shp))
    return out


def conv_stem_fprop_fused(xs2d, wp, K, bias, relu=True):
    """the 7x7/2 (or 3x3/2) C_in = 3 stem on the space-to-depth image with folded BN + ReLU in the epilogue"""
    N, _, H2, WP = xs2d.shape
    H, W = H2 * 2, (WP - 4) * 2
    out = nhwc_empty(N, K, H2, WP - 4, device=xs2d.device)
    ev = conv_prof.begin()
    call("tsb_conv_stem_fprop_fused", ptr(xs2d), N, H, W, ptr(wp), K, ptr(bias), int(relu), ptr(out), K, stream())
    conv_prof.end(ev, 2.0 * N * H2 * (WP - 4) * K * 147)
    return out


class _FoldCache(object):
    """eval-mode BatchNorm folded into the conv operands: w' = w · γ/√(σ²+ε) (bf16 KRSC pack), b' = β − μ·γ/√(σ²+ε).
    One entry per (conv weight, bn) pair, rebuilt when any of the five tensors changes (their versions, plus
    pack_cache.step because the fused SGD / a graph replay updates parameters without bumping versions)."""

    def __init__(self):
        self.cache = {}

    def get(self, w, bn, stem):
        key = (id(w), id(bn))
        ver = (w.data_ptr(), w._version, bn.weight._version, bn.bias._version, bn.running_mean._version,
               bn.running_var._version, pack_cache.step, bool(stem))
        hit = self.cache.get(key)
        if hit is not None and hit[0] == ver and hit[3]() is w:
            return hit[1], hit[2]
        with torch.no_grad():
            scale = (bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps))
            bias = (bn.bias.float() - bn.running_mean.float() * scale).contiguous()
            wk = (_krsc_ptr(w).float() * scale[:, None, None, None]).contiguous()
            K, R, S, C = wk.shape
            if stem:
                wb = torch.empty((K, 4, 4, 16), dtype=_BF, device=w.device)
                call("tsb_pack_stem_weight", ptr(wk), K, R, ptr(wb), stream())
            else:
                wb = torch.empty((K, R, S, C), dtype=_BF, device=w.device)
                call("tsb_pack_weight", ptr(wk), K, R, S, C, ptr(wb), None, stream())
        self.cache[key] = (ver, wb, bias, weakref.ref(w))
        if len(self.cache) > 1024:
            for k in [k for k, v in self.cache.items() if v[3]() is None]:
                del self.cache[k]
        return wb, bias


fold_cache = _FoldCache()


def conv_dgrad(dy, wt, xshape, K, R, stride, pad, dil, out=None, accumulate=False):
    N, C, H, W = xshape
    shp = make_shape(N, H, W, C, K, R, stride, pad, dil)
    if out is None:
        out = nhwc_empty(N, C, H, W, device=dy.device)
    ev = conv_prof.begin()
    try:
        call("tsb_conv2d_dgrad", ctypes.byref(shp), ptr(dy), cs_of(dy), ptr(wt), ptr(out), cs_of(out), int(accumulate), stream())
    except RuntimeError as e:
        raise RuntimeError("%s [dgrad shape N=%d H=%d W=%d C=%d K=%d R=%d stride=%d pad=%d dil=%d]" % (
            e, N, H, W, C, K, R, stride, pad, dil)) from None
    conv_prof.end(ev, _conv_flops(shp))
    return out


def conv_wgrad(x, dy, K, R, stride, pad, dil, dw_krsc):
    N, C, H, W = x.shape
    shp = make_shape(N, H, W, C, K, R, stride, pad, dil)
    ev = conv_prof.begin()
    call("tsb_conv2d_wgrad", ctypes.byref(shp), ptr(x), cs_of(x), ptr(dy), cs_of(dy), ptr(dw_krsc), stream())
    conv_prof.end(ev, _conv_flops(shp))


def _new_wgrad(w):
    """zeroed gradient with the parameter's (KRSC) strides: returns (logical view, dense KRSC tensor)"""
    K, C, R, S = w.shape
    g = torch.zeros((K, R, S, C), dtype=torch.float32, device=w.device)
    return g.permute(0, 3, 1, 2), g


# --------------------------------------------------------------------------------------------------
# Conv + (Sync)BatchNorm(train) + residual + ReLU  — ConvBnRelu / BasicBlock tail
# --------------------------------------------------------------------------------------------------
class GradShare(object):
    """One gradient buffer for a tensor consumed `n` times inside a block (x → conv1, x → downsample, x → residual).

    Each consumer's backward calls contribute(g): the first stores its gradient and returns None to autograd; later ones
    have already ACCUMULATED into the stored buffer (conv_dgrad(out=stash, accumulate=True)) and the last returns it. The
    sum therefore reaches autograd as a single tensor and the element-wise bf16 adds (one read-read-write pass over
    the activation per fork) disappear. Consumers outside the block are untouched: autograd adds their gradient to the
    returned tensor as usual."""

    def __init__(self, n):
        self.n = n
        self.left = n
        self.stash = None

    def contribute(self, g, accumulated=False):
        """g: this consumer's input gradient; accumulated=True means g IS the stash (the consumer's dgrad epilogue added
        into it in place). Returns what the consumer hands to autograd: None until the last contributor."""
        self.left -= 1
        if self.stash is None:
            self.stash = g
        elif not accumulated:
            self.stash = self.stash + g   # a contributor that cannot accumulate in place arrived late (not the case for
            #                               the block orders used here: the shortcut's backward always runs first)
        if self.left > 0:
            return None
        out, self.stash, self.left = self.stash, None, self.n
        return out


class ConvBNActFn(torch.autograd.Function):
    """y = act(BN_train(conv(x, w)) + residual).

    Replaces ConvBnRelu.forward (/root/reference/furnace/seg_opr/seg_oprs.py:39-46) and the conv→bn→(+res)→relu
    groups of BasicBlock.forward (/root/reference/furnace/base_model/resnet.py:33-53). The conv epilogue emits
    the per-channel Σ/Σ² (BN statistics), tsb_bn_finalize turns them into scale/shift (and updates the
    running stats), tsb_bn_apply normalises (+residual, +ReLU) in one pass.
    """

    @staticmethod
    def forward(ctx, x, w, gamma, beta, residual, running_mean, running_var, stride, pad, dil, relu, eps, momentum,
                training, stem, in_share=None, res_share=None):
        K = w.shape[0]
        R = w.shape[2]
        dev = x.device
        if stem:
            N, _, H2, WP = x.shape  # xs2d [N, 16, H/2, W/2+4]
            H, W = H2 * 2, (WP - 4) * 2
            P, Q = H2, WP - 4
        else:
            N, C, H, W = x.shape
            P, Q = conv_out_size(H, R, stride, pad, dil), conv_out_size(W, R, stride, pad, dil)
        count = float(N * P * Q) * _sync["world"]
        stats = zero_arena.take((2, K), dev) if training else None
        raw = nhwc_empty(N, K, P, Q, device=dev)
        if stem:
            wp, _ = pack_cache.get(w, False, stem=True)
            ev = conv_prof.begin()
            call("tsb_conv_stem_fprop", ptr(x), N, H, W, ptr(wp), K, ptr(raw), K, ptr(stats[0]) if training else None,
                 ptr(stats[1]) if training else None, stream())
            conv_prof.end(ev, 2.0 * N * P * Q * K * 147)
        else:
            wb, _ = pack_cache.get(w, False)
            conv_fprop(x, wb, K, R, stride, pad, dil, out=raw, stats=stats)
        aux = torch.empty((4, K), dtype=torch.float32, device=dev)  # mean, invstd, scale, shift
        if training:
            _allreduce_stats(stats)
            call("tsb_bn_finalize", ptr(stats[0]), ptr(stats[1]), count, K, ptr(gamma), ptr(beta), eps, momentum,
                 ptr(aux[0]), ptr(aux[1]), ptr(aux[2]), ptr(aux[3]), ptr(running_mean), ptr(running_var), stream())
        else:
            # eval: scale/shift from the running statistics (sum := mean*1, sumsq := var + mean^2, count 1)
            s = torch.stack([running_mean, running_var + running_mean * running_mean]).contiguous()
            call("tsb_bn_finalize", ptr(s[0]), ptr(s[1]), 1.0, K, ptr(gamma), ptr(beta), eps, 0.0, ptr(aux[0]),
                 ptr(aux[1]), ptr(aux[2]), ptr(aux[3]), None, None, stream())
        y = nhwc_empty(N, K, P, Q, device=dev)
        call("tsb_bn_apply", ptr(raw), K, ptr(aux[2]), ptr(aux[3]), ptr(residual), cs_of(residual) if residual is not None else 0,
             int(relu), ptr(y), K, N * P * Q, K, stream())
        ctx.save_for_backward(x, w, gamma, raw, y, aux)
        ctx.beta_ref = beta
        ctx.cfg = (stride, pad, dil, relu, stem, residual is not None, count, training)
        # gradient sharing at a fork (GradShare): in_share — this conv's input x has other consumers in the same block;
        # res_share — the residual IS that x. The contributors chain their input gradients through one buffer (dgrad
        # epilogue accumulate) instead of handing autograd several tensors to add.
        ctx.in_share = in_share
        ctx.res_share = res_share if residual is not None else None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, gamma, raw, y, aux = ctx.saved_tensors
        stride, pad, dil, relu, stem, has_res, count, training = ctx.cfg
        assert training, "backward through eval-mode BN is not supported"
        K = w.shape[0]
        R = w.shape[2]
        N, _, P, Q = raw.shape
        dev = dy.device
        npix = N * P * Q
        if dy.dtype != _BF or dy.stride(1) != 1:
            dy = to_nhwc(dy)
        red = zero_arena.take((2, K), dev)
        # ReLU mask: recomputed from raw*scale+shift when there is no residual (saves reading y in both passes)
        y_mask = ptr(y) if (relu and has_res) else None
        call("tsb_bn_bwd_reduce", ptr(dy), cs_of(dy), y_mask, K, ptr(raw), K, ptr(aux[0]), ptr(aux[1]), int(relu), npix, K,
             ptr(red[0]), ptr(red[1]), ptr(aux[2]), ptr(aux[3]), stream())
        # dgamma = Σ dz·x̂, dbeta = Σ dz are LOCAL sums (the DDP all-reduce averages parameter grads);
        # the dx formula needs the GLOBAL sums under SyncBN.
        beta = ctx.beta_ref
        direct = _direct_grad(w) and _direct_grad(gamma) and _direct_grad(beta)
        fold = direct and _sync["world"] == 1   # single GPU: the accumulation is folded into tsb_bn_bwd_apply
        if direct:
            dgamma = dbeta = None
            # multi-GPU: the exchange kernel adds the LOCAL sums to gamma.grad / beta.grad before it reduces
            _allreduce_stats(red, None if fold else gamma.grad, None if fold else beta.grad)
        else:
            dgamma = red[1].clone()
            dbeta = red[0].clone()
            _allreduce_stats(red)
        draw = nhwc_empty(N, K, P, Q, device=dev)
        dres = nhwc_empty(N, K, P, Q, device=dev) if has_res else None
        call("tsb_bn_bwd_apply", ptr(dy), cs_of(dy), y_mask, K, ptr(raw), K, ptr(aux[0]), ptr(aux[1]), ptr(gamma), ptr(red[0]),
             ptr(red[1]), count, int(relu), ptr(draw), K, ptr(dres), K if has_res else 0, npix, K,
             ptr(gamma.grad) if fold else None, ptr(beta.grad) if fold else None, ptr(aux[2]), ptr(aux[3]), stream())
        if direct:
            dw_view, dw = None, w.grad.permute(0, 2, 3, 1)   # KRSC view of the flat gradient buffer
        else:
            dw_view, dw = _new_wgrad(w)
        dx = None
        if stem:
            H, W = P * 2, Q * 2
            dwp = zero_arena.take((K, 4, 64), dev)
            ev = conv_prof.begin()
            call("tsb_conv_stem_wgrad", ptr(x), N, H, W, ptr(draw), K, K, ptr(dwp), stream())
            conv_prof.end(ev, 2.0 * N * P * Q * K * 147)
            call("tsb_unpack_stem_wgrad", ptr(dwp), K, R, ptr(dw), stream())
        else:
            conv_wgrad(x, draw, K, R, stride, pad, dil, dw)
            if ctx.needs_input_grad[0]:
                _, wt = pack_cache.get(w, True)
                share = ctx.in_share
                if share is not None and share.stash is not None:
                    dx = conv_dgrad(draw, wt, x.shape, K, R, stride, pad, dil, out=share.stash, accumulate=True)
                    dx = share.contribute(dx, accumulated=True)
                else:
                    dx = conv_dgrad(draw, wt, x.shape, K, R, stride, pad, dil)
                    if share is not None:
                        dx = share.contribute(dx)
        if has_res and ctx.res_share is not None and ctx.needs_input_grad[4]:
            dres = ctx.res_share.contribute(dres)
        if direct:
            _notify(w, gamma, beta)
        return dx, dw_view, dgamma, dbeta, dres, None, None, None, None, None, None, None, None, None, None, None, None


class StemPairFn(torch.autograd.Function):
    """The TWO 7x7/2 stems of BiSeNet (spatial_path.conv_7x7, bisenet network.py:118, and context_path.conv1,
    resnet.py:126) read the same image: one tcgen05 launch computes both (K = 64 + 64 output channels, N = 128 MMA
    tiles), one BN-apply pass normalises both, one wgrad launch produces both weight gradients. Returns the two
    64-channel halves as channel-slice views of one NHWC buffer."""

    @staticmethod
    def forward(ctx, xs, w1, g1, b1, rm1, rv1, w2, g2, b2, rm2, rv2, eps1, mom1, eps2, mom2):
        N, _, H2, WP = xs.shape
        H, W, P, Q = H2 * 2, (WP - 4) * 2, H2, WP - 4
        dev = xs.device
        K1, K2 = w1.shape[0], w2.shape[0]
        K = K1 + K2
        wp = torch.empty((K, 4, 4, 16), dtype=_BF, device=dev)
        call("tsb_pack_stem_weight", ptr(_krsc_ptr(w1)), K1, w1.shape[2], ptr(wp), stream())
        call("tsb_pack_stem_weight", ptr(_krsc_ptr(w2)), K2, w2.shape[2], ptr(wp[K1:]), stream())
        stats = zero_arena.take((2, K), dev)
        raw = nhwc_empty(N, K, P, Q, device=dev)
        # two N=64 launches into the channel halves of one buffer: this layer is bound by the epilogue's stores and the
        # 64-column epilogue (one chunk per warp, register running sums) is the faster one; the wgrad below is fused
        for off, Kh in ((0, K1), (K1, K2)):
            ev = conv_prof.begin()
            call("tsb_conv_stem_fprop", ptr(xs), N, H, W, ptr(wp[off:]), Kh, ptr(raw[:, off:]), K, ptr(stats[0][off:]),
                 ptr(stats[1][off:]), stream())
            conv_prof.end(ev, 2.0 * N * P * Q * Kh * 147)
        _allreduce_stats(stats)
        count = float(N * P * Q) * _sync["world"]
        aux = torch.empty((4, K), dtype=torch.float32, device=dev)
        for (off, Kh, g, b, rm, rv, eps, mom) in ((0, K1, g1, b1, rm1, rv1, eps1, mom1), (K1, K2, g2, b2, rm2, rv2, eps2, mom2)):
            call("tsb_bn_finalize", ptr(stats[0][off:]), ptr(stats[1][off:]), count, Kh, ptr(g), ptr(b), eps, mom,
                 ptr(aux[0][off:]), ptr(aux[1][off:]), ptr(aux[2][off:]), ptr(aux[3][off:]), ptr(rm), ptr(rv), stream())
        y = nhwc_empty(N, K, P, Q, device=dev)
        call("tsb_bn_apply", ptr(raw), K, ptr(aux[2]), ptr(aux[3]), None, 0, 1, ptr(y), K, N * P * Q, K, stream())
        ctx.save_for_backward(xs, w1, g1, w2, g2, raw, aux)
        ctx.refs = (b1, b2)
        ctx.count = count
        return y[:, :K1], y[:, K1:]

    @staticmethod
    def backward(ctx, dy1, dy2):
        xs, w1, g1, w2, g2, raw, aux = ctx.saved_tensors
        b1, b2 = ctx.refs
        N, K, P, Q = raw.shape
        K1 = w1.shape[0]
        dev = raw.device
        npix = N * P * Q
        draw = nhwc_empty(N, K, P, Q, device=dev)
        grads = []
        for (off, w, g, b, dy) in ((0, w1, g1, b1, dy1), (K1, w2, g2, b2, dy2)):
            Kh = w.shape[0]
            if dy.dtype != _BF or dy.stride(1) != 1:
                dy = to_nhwc(dy)
            red = zero_arena.take((2, Kh), dev)
            xr = raw[:, off:off + Kh]
            call("tsb_bn_bwd_reduce", ptr(dy), cs_of(dy), None, 0, ptr(xr), K, ptr(aux[0][off:]), ptr(aux[1][off:]), 1, npix, Kh,
                 ptr(red[0]), ptr(red[1]), ptr(aux[2][off:]), ptr(aux[3][off:]), stream())
            direct = _direct_grad(w) and _direct_grad(g) and _direct_grad(b)
            fold = direct and _sync["world"] == 1
            dg, db = (None, None) if direct else (red[1].clone(), red[0].clone())
            _allreduce_stats(red, g.grad if (direct and not fold) else None, b.grad if (direct and not fold) else None)
            call("tsb_bn_bwd_apply", ptr(dy), cs_of(dy), None, 0, ptr(xr), K, ptr(aux[0][off:]), ptr(aux[1][off:]), ptr(g),
                 ptr(red[0]), ptr(red[1]), ctx.count, 1, ptr(draw[:, off:off + Kh]), K, None, 0, npix, Kh,
                 ptr(g.grad) if fold else None, ptr(b.grad) if fold else None, ptr(aux[2][off:]), ptr(aux[3][off:]), stream())
            grads.append((direct, dg, db))
        H, W = P * 2, Q * 2
        dwp = zero_arena.take((K, 4, 64), dev)
        ev = conv_prof.begin()
        call("tsb_conv_stem_wgrad", ptr(xs), N, H, W, ptr(draw), K, K, ptr(dwp), stream())
        conv_prof.end(ev, 2.0 * N * P * Q * K * 147)
        outs = []
        for (off, w, g, b), (direct, dg, db) in zip(((0, w1, g1, b1), (K1, w2, g2, b2)), grads):
            Kh = w.shape[0]
            if direct:
                dwv, dw = None, w.grad.permute(0, 2, 3, 1)
            else:
                dwv, dw = _new_wgrad(w)
            call("tsb_unpack_stem_wgrad", ptr(dwp[off:]), Kh, w.shape[2], ptr(dw), stream())
            if direct:
                _notify(w, g, b)
            outs.append((dwv, dg, db))
        (dw1, dg1, db1), (dw2, dg2, db2) = outs
        return None, dw1, dg1, db1, None, None, dw2, dg2, db2, None, None, None, None, None, None


class ConvFn(torch.autograd.Function):
    """plain convolution (+bias), bf16 or fp32 output with channel stride `ocs` — the 1x1 classifier heads
    (bisenet network.py:156-161) and the BN-free SE convs of FeatureFusion (seg_oprs.py:224-229)."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, pad, dil, out_f32, ocs):
        K, C, R, _ = w.shape
        wb, _ = pack_cache.get(w, False)
        y = conv_fprop(x, wb, K, R, stride, pad, dil, bias=bias, out_dtype=torch.float32 if out_f32 else _BF, ocs=ocs)
        ctx.save_for_backward(x, w)
        ctx.bias_ref = bias
        ctx.cfg = (stride, pad, dil, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, pad, dil, has_bias = ctx.cfg
        K, C, R, _ = w.shape
        N, _, P, Q = dy.shape
        dev = dy.device
        # dgrad needs the GEMM-K (= out channels) padded to a multiple of 64 in a bf16 NHWC buffer
        Kp = (K + 63) // 64 * 64
        if dy.dtype == _BF and dy.stride(1) == 1 and cs_of(dy) >= Kp and K == Kp:
            dyb = dy
        else:
            if dy.stride(1) == 1 and dy.dtype in (_BF, torch.float32):
                dyb = nhwc_empty(N, Kp, P, Q, device=dev)     # tsb_cast_pad writes all Kp channels (zero pads)
                call("tsb_cast_pad", ptr(dy), _lib.dt(dy), cs_of(dy), K, ptr(dyb), Kp, Kp, N * P * Q, stream())
            else:
                dyb = nhwc_zeros(N, Kp, P, Q, device=dev)
                dyb[:, :K].copy_(dy)
        direct = _direct_grad(w) and (not has_bias or _direct_grad(ctx.bias_ref))
        if direct:
            dw_view, dw = None, w.grad.permute(0, 2, 3, 1)
        else:
            dw_view, dw = _new_wgrad(w)
        conv_wgrad(x, dyb[:, :K], K, R, stride, pad, dil, dw)
        db = None
        if has_bias:
            db = ctx.bias_ref.grad if direct else torch.zeros((K,), dtype=torch.float32, device=dev)
            call("tsb_bias_grad", ptr(dyb), Kp, N * P * Q, K, ptr(db), stream())
            if direct:
                db = None
        dx = None
        if ctx.needs_input_grad[0]:
            _, wt = pack_cache.get(w, True, pad_k=Kp)
            dx = conv_dgrad(dyb, wt, x.shape, Kp, R, stride, pad, dil)
        if direct:
            _notify(*([w, ctx.bias_ref] if has_bias else [w]))
        return dx, dw_view, db, None, None, None, None, None


# --------------------------------------------------------------------------------------------------
# pooling / resize / attention scale
# --------------------------------------------------------------------------------------------------
class MaxPool3x3S2Fn(torch.autograd.Function):
    """nn.MaxPool2d(3, 2, 1) — /root/reference/furnace/base_model/resnet.py:132"""

    @staticmethod
    def forward(ctx, x):
        N, C, H, W = x.shape
        P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = nhwc_empty(N, C, P, Q, device=x.device)
        idx = torch.empty((N, P, Q, C), dtype=torch.uint8, device=x.device)
        call("tsb_maxpool3x3s2_fwd", ptr(x), cs_of(x), ptr(y), C, ptr(idx), N, C, H, W, stream())
        ctx.save_for_backward(idx)
        ctx.shape = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        N, C, H, W = ctx.shape
        if dy.dtype != _BF or dy.stride(1) != 1:
            dy = to_nhwc(dy)
        dx = nhwc_empty(N, C, H, W, device=dy.device)
        call("tsb_maxpool3x3s2_bwd", ptr(idx), ptr(dy), cs_of(dy), ptr(dx), C, N, C, H, W, stream())
        return dx


class AdaptiveAvgPoolFn(torch.autograd.Function):
    """nn.AdaptiveAvgPool2d(S) — seg_oprs.py:200,223; bisenet network.py:35; pspnet network.py:83.
    Output bf16 [N,C,S,S] (NHWC)."""

    @staticmethod
    def forward(ctx, x, S):
        N, C, H, W = x.shape
        o32 = torch.empty((N, S, S, C), dtype=torch.float32, device=x.device)
        call("tsb_adaptive_avgpool_fwd", ptr(x), cs_of(x), N, C, H, W, S, ptr(o32), stream())
        out = nhwc_empty(N, C, S, S, device=x.device)
        call("tsb_cast_scale", ptr(o32), F32, C, ptr(out), BF16, C, N * S * S, C, None, stream())
        ctx.shape = (N, C, H, W, S)
        return out

    @staticmethod
    def backward(ctx, dy):
        N, C, H, W, S = ctx.shape
        if dy.stride(1) != 1:
            dy = to_nhwc(dy, dy.dtype)
        d32 = torch.empty((N, S, S, C), dtype=torch.float32, device=dy.device)
        call("tsb_cast_scale", ptr(dy), _lib.dt(dy), cs_of(dy), ptr(d32), F32, C, N * S * S, C, None, stream())
        dx = nhwc_empty(N, C, H, W, device=dy.device)
        call("tsb_adaptive_avgpool_bwd", ptr(d32), N, C, H, W, S, ptr(dx), C, 0, stream())
        return dx, None


_scalars = {}


def _const_scalar(v, device):
    """cached 1-element fp32 device tensor (scale factors read by kernels from device memory)"""
    key = (v, str(device))
    t = _scalars.get(key)
    if t is None:
        t = torch.tensor([v], dtype=torch.float32, device=device)
        _scalars[key] = t
    return t


class BilinearFn(torch.autograd.Function):
    """F.interpolate(mode='bilinear', align_corners=True) — bisenet network.py:82-84,93-94"""

    @staticmethod
    def forward(ctx, x, Ho, Wo):
        N, C, Hi, Wi = x.shape
        y = nhwc_empty(N, C, Ho, Wo, device=x.device)
        call("tsb_bilinear_fwd", ptr(x), _lib.dt(x), cs_of(x), ptr(y), BF16, C, N, C, Hi, Wi, Ho, Wo, stream())
        ctx.shape = (N, C, Hi, Wi, Ho, Wo)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, Hi, Wi, Ho, Wo = ctx.shape
        if dy.dtype != _BF or dy.stride(1) != 1:
            dy = to_nhwc(dy)
        dx = nhwc_empty(N, C, Hi, Wi, device=dy.device)
        if Hi == 1 and Wi == 1:
            # broadcast forward (1x1 → HxW, bisenet network.py:82-84) ⇒ the backward is a plain spatial sum: reuse the
            # pooling reduction (mean) and scale by H*W instead of a one-block gather
            m32 = torch.empty((N, 1, 1, C), dtype=torch.float32, device=dy.device)
            call("tsb_adaptive_avgpool_fwd", ptr(dy), cs_of(dy), N, C, Ho, Wo, 1, ptr(m32), stream())
            call("tsb_cast_scale", ptr(m32), F32, C, ptr(dx), BF16, C, N, C, ptr(_const_scalar(float(Ho * Wo), dy.device)), stream())
        else:
            call("tsb_bilinear_bwd", ptr(dy), BF16, cs_of(dy), ptr(dx), BF16, C, N, C, Hi, Wi, Ho, Wo, 0, stream())
        return dx, None, None


class ChanScaleFn(torch.autograd.Function):
    """y = x * (base + sigmoid(a)) + add.  ARM: base 0 (+ the caller's `fm += last_fm`, bisenet network.py:91-92);
    FFM: base 1 (fm + fm*se, seg_oprs.py:237). `a` is the bf16 [N,C,1,1] pre-sigmoid attention logit."""

    @staticmethod
    def forward(ctx, x, a, add, base):
        N, C, H, W = x.shape
        a32 = torch.empty((N, C), dtype=torch.float32, device=x.device)
        call("tsb_cast_scale", ptr(a), _lib.dt(a), cs_of(a), ptr(a32), F32, C, N, C, None, stream())
        y = nhwc_empty(N, C, H, W, device=x.device)
        call("tsb_chan_scale_fwd", ptr(x), cs_of(x), ptr(a32), float(base), ptr(add), cs_of(add) if add is not None else 0,
             ptr(y), C, N, H * W, C, stream())
        ctx.save_for_backward(x, a32)
        ctx.cfg = (base, add is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, a32 = ctx.saved_tensors
        base, has_add = ctx.cfg
        N, C, H, W = x.shape
        if dy.dtype != _BF or dy.stride(1) != 1:
            dy = to_nhwc(dy)
        dx = nhwc_empty(N, C, H, W, device=x.device)
        da32 = torch.zeros((N, C), dtype=torch.float32, device=x.device)
        call("tsb_chan_scale_bwd", ptr(dy), cs_of(dy), ptr(x), cs_of(x), ptr(a32), float(base), ptr(dx), C, ptr(da32), N,
             H * W, C, stream())
        da = nhwc_empty(N, C, 1, 1, device=x.device)
        call("tsb_cast_scale", ptr(da32), F32, C, ptr(da), BF16, C, N, C, None, stream())
        return dx, da, (dy if has_add else None), None


class ConcatFn(torch.autograd.Function):
    """torch.cat([x1, x2, ...], dim=1) (FeatureFusion seg_oprs.py:234, PyramidPooling pspnet network.py:106) as strided
    copies into ONE NHWC buffer; backward hands out channel-slice views (no copy)."""

    @staticmethod
    def forward(ctx, *xs):
        N, _, H, W = xs[0].shape
        chans = [int(x.shape[1]) for x in xs]
        Ct = sum(chans)
        out = nhwc_empty(N, Ct, H, W, device=xs[0].device)
        npix = N * H * W
        off = 0
        for x, c in zip(xs, chans):
            call("tsb_cast_scale", ptr(x), BF16, cs_of(x), ptr(out[:, off:]), BF16, Ct, npix, c, None, stream())
            off += c
        ctx.chans = chans
        return out

    @staticmethod
    def backward(ctx, dy):
        if dy.dtype != _BF or dy.stride(1) != 1:
            dy = to_nhwc(dy)
        outs, off = [], 0
        for c in ctx.chans:
            outs.append(dy[:, off:off + c])
            off += c
        return tuple(outs)


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        N, C, H, W = a.shape
        y = nhwc_empty(N, C, H, W, device=a.device)
        call("tsb_add", ptr(a), cs_of(a), ptr(b), cs_of(b), ptr(y), C, N * H * W, C, stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


class PSABmmFn(torch.autograd.Function):
    """out = bmm(x.view(b,c,-1), softmax(att.view(b,L,-1), dim=1)) with L = h*w attention channels per pixel
    (PointwiseSpatialAttention.forward, psanet network.py:122-126,131-138).

    NHWC form per image: out[j, ch] = Σ_i S[j, i] · x[i, ch] with S = row-softmax of att[j, :]. That is a 1x1
    convolution whose "image" is S (L pixels x L channels, zero-padded to a multiple of 64 channels) and whose weight
    operand is the image's own feature map transposed ([c, L]); dS is the matching dgrad (weight operand = x itself),
    dx the matching wgrad. All three run on the tcgen05 conv kernels, one launch per image."""

    @staticmethod
    def forward(ctx, x, att):
        b, c, h, w = x.shape
        L = att.shape[1]
        assert L == h * w and att.shape[0] == b and tuple(att.shape[2:]) == (h, w), "PSA needs h*w attention channels"
        assert c % 64 == 0 and L % 8 == 0
        dev = x.device
        if cs_of(x) != c:
            x = to_nhwc(x)
        Lp = (L + 63) // 64 * 64
        S = nhwc_empty(b, Lp, h, w, device=dev)
        call("tsb_softmax_rows_fwd", ptr(att), _lib.dt(att), cs_of(att), ptr(S), Lp, b * h * w, L, Lp, stream())
        out = nhwc_empty(b, c, h, w, device=dev)
        xt = torch.empty((c, Lp), dtype=_BF, device=dev)
        for n in range(b):
            call("tsb_transpose_pad", ptr(x[n:n + 1]), BF16, c, ptr(xt), Lp, L, c, Lp, stream())
            conv_fprop(S[n:n + 1], xt, c, 1, 1, 0, 1, out=out[n:n + 1])
        ctx.save_for_backward(x, S)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, S = ctx.saved_tensors
        b, c, h, w = x.shape
        Lp = S.shape[1]
        L = h * w
        dev = x.device
        if dout.dtype != _BF or dout.stride(1) != 1:
            dout = to_nhwc(dout)
        dS = nhwc_empty(b, L, h, w, device=dev, cs=Lp)
        dx = nhwc_empty(b, c, h, w, device=dev)
        dxt = torch.empty((c, Lp), dtype=torch.float32, device=dev)
        for n in range(b):
            conv_dgrad(dout[n:n + 1], x[n:n + 1], (1, L, h, w), c, 1, 1, 0, 1, out=dS[n:n + 1])
            dxt.zero_()
            conv_wgrad(S[n:n + 1], dout[n:n + 1], c, 1, 1, 0, 1, dxt)
            call("tsb_transpose_pad", ptr(dxt), F32, Lp, ptr(dx[n:n + 1]), c, c, L, c, stream())
        dA = nhwc_empty(b, L, h, w, device=dev, cs=Lp)
        call("tsb_softmax_rows_bwd", ptr(S), Lp, ptr(dS), Lp, ptr(dA), Lp, b * h * w, L, Lp, stream())
        return dx, dA


# --------------------------------------------------------------------------------------------------
# image packing (network input boundary)
# --------------------------------------------------------------------------------------------------
def pack_image_s2d(img):
    """NCHW fp32 image → [N,16,H/2,W/2+4] bf16 (NHWC) operand of the 7x7/2 stems"""
    N, C, H, W = img.shape
    assert C == 3 and img.dtype == torch.float32 and img.is_contiguous()
    if (H | W) & 1:
        # odd sizes (713, 473): one extra zero row / column at the bottom / right is exactly the stem's own zero padding
        # (the stride-2 output size floor((H + 2p - k) / 2) + 1 is the same for H and H + 1 when H is odd)
        img = torch.nn.functional.pad(img, (0, W & 1, 0, H & 1)).contiguous()
        H, W = H + (H & 1), W + (W & 1)
    out = nhwc_empty(N, 16, H // 2, W // 2 + 4, device=img.device)
    call("tsb_pack_image_s2d", ptr(img), N, H, W, ptr(out), stream())
    return out


# --------------------------------------------------------------------------------------------------
# OHEM cross-entropy
# --------------------------------------------------------------------------------------------------
def _ohem_forward(pt_call, n, labels, ignore_label, thresh, min_kept, class_weight, dev):
    state = torch.empty((_lib.OHEM_STATE_WORDS,), dtype=torch.int32, device=dev)
    p = torch.empty((n,), dtype=torch.float32, device=dev)
    nll = torch.empty((n,), dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    call("tsb_ohem_begin", ptr(state), stream())
    pt_call(p, nll, state)
    call("tsb_ohem_select", ptr(p), n, int(min_kept), float(thresh), ptr(state), stream())
    call("tsb_ohem_loss", ptr(p), ptr(nll), ptr(labels), n, int(ignore_label), ptr(class_weight), ptr(state), ptr(loss),
         stream())
    return loss, p, state


class OhemCEFn(torch.autograd.Function):
    """ProbOhemCrossEntropy2d.forward on MATERIALISED logits [N,C,H,W] (any strides, fp32/bf16) —
    /root/reference/furnace/seg_opr/loss_opr.py:68-98."""

    @staticmethod
    def forward(ctx, pred, target, ignore_label, thresh, min_kept, class_weight):
        N, C, H, W = pred.shape
        target = target.contiguous()
        sn, sc, sy, sx = pred.stride()
        dtp = _lib.dt(pred)

        def pt(p, nll, state):
            call("tsb_ohem_ptarget", ptr(pred), dtp, sn, sc, sy, sx, ptr(target), N, C, H, W, int(ignore_label),
                 float(thresh), ptr(p), ptr(nll), ptr(state), stream())

        loss, p, state = _ohem_forward(pt, N * H * W, target, ignore_label, thresh, min_kept, class_weight, pred.device)
        ctx.save_for_backward(pred, target, p, state, class_weight if class_weight is not None else torch.empty(0))
        ctx.cfg = (ignore_label, class_weight is not None)
        return loss

    @staticmethod
    def backward(ctx, g):
        pred, target, p, state, cw = ctx.saved_tensors
        ignore_label, has_cw = ctx.cfg
        N, C, H, W = pred.shape
        sn, sc, sy, sx = pred.stride()
        d = torch.empty_strided(pred.shape, pred.stride(), dtype=pred.dtype, device=pred.device)
        g = g.to(torch.float32).contiguous()
        call("tsb_ohem_grad", ptr(pred), _lib.dt(pred), sn, sc, sy, sx, ptr(target), ptr(p), N, C, H, W, int(ignore_label),
             ptr(cw) if has_cw else None, ptr(state), ptr(g), ptr(d), stream())
        return d, None, None, None, None, None


class OhemUpCEFn(torch.autograd.Function):
    """Fused head tail: bilinear x`scale` upsample (align_corners) of the LOW-res fp32 NHWC logits +
    ProbOhemCrossEntropy2d; the full-resolution logits are never written (bisenet network.py:104-108,163-166)."""

    @staticmethod
    def forward(ctx, lo, target, H, W, num_classes, ignore_label, thresh, min_kept, class_weight):
        N, _, h, w = lo.shape
        assert lo.dtype == torch.float32
        target = target.contiguous()
        cs = cs_of(lo)

        def pt(p, nll, state):
            call("tsb_ohem_ptarget_up", ptr(lo), cs, h, w, ptr(target), N, num_classes, H, W, int(ignore_label),
                 float(thresh), ptr(p), ptr(nll), ptr(state), stream())

        loss, p, state = _ohem_forward(pt, N * H * W, target, ignore_label, thresh, min_kept, class_weight, lo.device)
        ctx.save_for_backward(lo, target, p, state, class_weight if class_weight is not None else torch.empty(0))
        ctx.cfg = (H, W, num_classes, ignore_label, class_weight is not None)
        return loss

    @staticmethod
    def backward(ctx, g):
        lo, target, p, state, cw = ctx.saved_tensors
        H, W, C, ignore_label, has_cw = ctx.cfg
        N, Cl, h, w = lo.shape
        cs = cs_of(lo)
        d = nhwc_zeros(N, Cl, h, w, dtype=torch.float32, device=lo.device, cs=cs)
        g = g.to(torch.float32).contiguous()
        call("tsb_ohem_grad_up", ptr(lo), cs, h, w, ptr(target), ptr(p), N, C, H, W, int(ignore_label),
             ptr(cw) if has_cw else None, ptr(state), ptr(g), ptr(d), stream())
        return d, None, None, None, None, None, None, None, None


class CEUpFn(torch.autograd.Function):
    """Fused head tail for MANY classes (ADE's 150): bilinear up-sampling (align_corners) of the LOW-res fp32 NHWC logits to
    H x W + CrossEntropyLoss(mean, ignore_index); the full-resolution [N,C,H,W] logits are never written
    (pspnet network.py:46-57, psanet network.py:46-55). 1 <= C <= 152."""

    @staticmethod
    def forward(ctx, lo, target, H, W, num_classes, ignore_label):
        N, _, h, w = lo.shape
        assert lo.dtype == torch.float32
        target = target.contiguous()
        cs = cs_of(lo)
        dev = lo.device
        state = torch.empty((8,), dtype=torch.int32, device=dev)
        lse = torch.empty((N * H * W,), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        call("tsb_ce_up_fwd", ptr(lo), cs, h, w, ptr(target), N, int(num_classes), int(H), int(W), int(ignore_label), ptr(lse),
             ptr(state), ptr(loss), stream())
        ctx.save_for_backward(lo, target, lse, state)
        ctx.cfg = (int(H), int(W), int(num_classes), int(ignore_label))
        return loss

    @staticmethod
    def backward(ctx, g):
        lo, target, lse, state = ctx.saved_tensors
        H, W, C, ignore_label = ctx.cfg
        N, Cl, h, w = lo.shape
        cs = cs_of(lo)
        d = nhwc_zeros(N, Cl, h, w, dtype=torch.float32, device=lo.device, cs=cs)
        g = g.to(torch.float32).contiguous()
        call("tsb_ce_up_bwd", ptr(lo), cs, h, w, ptr(target), ptr(lse), N, C, H, W, ignore_label, ptr(state), ptr(g), ptr(d),
             stream())
        return d, None, None, None, None, None


FUSED_CE_MAX_CLASSES = 152   # tsb_ce_up_* register budget (38 class quads)


def ohem_state_dict(state):
    """decode the OHEM state words (host sync) — for tests / logging"""
    s = state.cpu()
    f = s.view(torch.float32)
    return dict(num_valid=int(s[_lib.OHEM_ST_NUM_VALID]), count_le=int(s[_lib.OHEM_ST_COUNT_LE]),
                active=bool(s[_lib.OHEM_ST_ACTIVE]), T=float(f[_lib.OHEM_ST_THRESH]), kept=int(s[_lib.OHEM_ST_KEPT]),
                loss=float(f[_lib.OHEM_ST_LOSS]))
