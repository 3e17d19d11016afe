"""CPU tests: the oracle against the golden vectors generated from the LIVE reference (tools/make_golden.py),
against the live reference itself when /root/reference is present, and the two oracle restatements (C and
torch) against each other."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from golden_cases import ohem_case, bisenet_case, fcn_case, fcn_r101_case, pspnet_case, dfn_case, psanet_case, OHEM_REGIMES

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_outputs.json")))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("regime", OHEM_REGIMES)
def test_c_oracle_ohem_vs_golden(regime):
    from oracle import c_oracle
    logits, labels, min_kept = ohem_case(regime)
    g = GOLD["ohem"][regime]
    out = c_oracle.ohem(logits.numpy(), labels.numpy(), 255, 0.7, min_kept, want_grad=True)
    if g["loss"] == "nan":
        assert np.isnan(out["loss"])
        return
    assert abs(out["loss"] - g["loss"]) < 1e-5 * abs(g["loss"])
    assert int(out["kept"].sum()) == g["kept_count"]
    assert _sha(out["kept"].astype(np.uint8)) == g["kept_sha256"], "kept index set differs from the reference"
    assert abs(float(np.abs(out["dlogits"]).sum()) - g["grad_abs_sum"]) < 1e-4 * g["grad_abs_sum"]


@pytest.mark.parametrize("regime", OHEM_REGIMES)
def test_torch_oracle_ohem_vs_golden_and_c(regime):
    from oracle import c_oracle, torch_ref
    logits, labels, min_kept = ohem_case(regime)
    g = GOLD["ohem"][regime]
    loss, valid, thr = torch_ref.ohem_ce(logits, labels, 255, 0.7, min_kept, return_aux=True)
    out = c_oracle.ohem(logits.numpy(), labels.numpy(), 255, 0.7, min_kept)
    if g["loss"] == "nan":
        assert torch.isnan(loss)
        return
    assert abs(float(loss) - g["loss"]) < 1e-6 * abs(g["loss"])
    assert np.array_equal(valid.numpy(), out["kept"])


def test_ohem_weighted_and_focal_vs_golden():
    from oracle import c_oracle, torch_ref
    from torchseg_b200.seg_opr.loss_opr import ProbOhemCrossEntropy2d
    logits, labels, min_kept = ohem_case("B")
    w = np.array(ProbOhemCrossEntropy2d.CITYSCAPES_WEIGHT, np.float32)
    out = c_oracle.ohem(logits.numpy(), labels.numpy(), 255, 0.7, min_kept, class_weight=w)
    assert abs(out["loss"] - GOLD["ohem"]["B_weighted"]["loss"]) < 1e-5 * abs(GOLD["ohem"]["B_weighted"]["loss"])
    g = torch.Generator().manual_seed(10)
    pred = torch.randn(2, 1, 32, 40, generator=g)
    tgt = torch.randint(0, 2, (2, 32, 40), generator=g)
    tgt[:, :3] = 255
    assert abs(float(torch_ref.sigmoid_focal(pred, tgt, 255)) - GOLD["focal"]["loss"]) < 1e-6


def _ref_sd(build, seed):
    torch.manual_seed(seed)
    m = build()
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def test_torch_oracle_bisenet_vs_golden():
    """same seed → same initial weights (our module mirrors the reference's construction order) → same loss"""
    from oracle import torch_ref
    from torchseg_b200.networks import BiSeNet
    x, y, min_kept, seed = bisenet_case()
    m, sd = _ref_sd(lambda: BiSeNet(19, True, None, None, torch.nn.BatchNorm2d), seed)
    g = GOLD["bisenet_r18"]
    assert sum(p.numel() for p in m.parameters()) == g["n_params"] and len(sd) == g["n_state"]
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    loss, _ = torch_ref.bisenet_r18_loss(x, y, sd, min_kept)
    assert abs(float(loss) - g["loss"]) < 1e-5 * g["loss"]
    loss.backward()
    for n, ref in g["grad_norms"].items():
        assert abs(float(sd[n].grad.norm()) - ref) < 1e-4 * ref, n


def test_torch_oracle_fcn_r18_vs_golden():
    """BASELINE configs[0] (plumbing case): FCN-32s R18, 2 x 256 x 256, CPU"""
    from oracle import torch_ref
    from torchseg_b200.networks import FCN
    x, y, seed = fcn_case()
    torch.manual_seed(seed)
    m = FCN(19, backbone="R18")
    sd = {}
    for k, v in m.state_dict().items():   # reference key names: backbone.*, head.cbr.*, head.conv1x1.*
        sd[k] = v.detach().clone()
    loss = torch_ref.fcn_r18_loss(x, y, sd)
    assert abs(float(loss) - GOLD["fcn_r18"]["loss"]) < 1e-5 * GOLD["fcn_r18"]["loss"]


def test_torch_oracle_fcn_r101_vs_golden():
    """the shipped FCN-32s R101_v1c: our module reproduces the reference construction (seed → same weights, same keys)
    and the oracle restatement reproduces the live reference's loss / gradient norms"""
    from oracle import torch_ref
    from torchseg_b200.networks import FCN
    x, y, seed = fcn_r101_case()
    torch.manual_seed(seed)
    m = FCN(21, torch.nn.CrossEntropyLoss(ignore_index=255), backbone="R101")
    g = GOLD["fcn_r101"]
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert sum(p.numel() for p in m.parameters()) == g["n_params"] and len(sd) == g["n_state"]
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    loss, _ = torch_ref.fcn_r101_loss(x, y, sd)
    assert abs(float(loss) - g["loss"]) < 1e-5 * g["loss"]
    loss.backward()
    for n, ref in g["grad_norms"].items():
        assert abs(float(sd[n].grad.norm()) - ref) < 1e-4 * ref, n


def test_torch_oracle_pspnet_vs_golden():
    """PSPNet-R101_v1c dilated-8: our module reproduces the reference construction (same seed → same weights, same
    state_dict keys) and the oracle restatement reproduces the live reference's loss / gradient norms"""
    from oracle import torch_ref
    from torchseg_b200.networks import PSPNet
    x, y, seed = pspnet_case()
    torch.manual_seed(seed)
    m = PSPNet(150, torch.nn.CrossEntropyLoss(ignore_index=-1))
    g = GOLD["pspnet_r101"]
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert sum(p.numel() for p in m.parameters()) == g["n_params"] and len(sd) == g["n_state"]
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    loss, _ = torch_ref.pspnet_loss(x, y, sd)
    assert abs(float(loss) - g["loss"]) < 1e-5 * g["loss"]
    loss.backward()
    for n, ref in g["grad_norms"].items():
        assert abs(float(sd[n].grad.norm()) - ref) < 1e-4 * ref, n


def test_torch_oracle_dfn_vs_golden():
    """DFN-R101_v1c: module construction (seed → same weights / keys / 90.22 M params) and the oracle restatement of
    RefineResidual / ChannelAttention / DFNHead / DFN.forward against the live reference's loss and gradient norms"""
    from oracle import torch_ref
    from torchseg_b200.networks import DFN
    from torchseg_b200.seg_opr.loss_opr import SigmoidFocalLoss
    x, y, e, seed = dfn_case()
    torch.manual_seed(seed)
    m = DFN(19, torch.nn.CrossEntropyLoss(ignore_index=255), SigmoidFocalLoss(255), 0.1)
    g = GOLD["dfn_r101"]
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert sum(p.numel() for p in m.parameters()) == g["n_params"] and len(sd) == g["n_state"]
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    loss, _ = torch_ref.dfn_loss(x, y, e, sd)
    assert abs(float(loss) - g["loss"]) < 1e-5 * g["loss"]
    loss.backward()
    for n, ref in g["grad_norms"].items():
        assert abs(float(sd[n].grad.norm()) - ref) < 1e-4 * ref, n


def test_torch_oracle_psanet_vs_golden():
    """PSANet-R101_v1c: construction parity (79.58 M params, 684 state entries) and the oracle restatement
    of PointwiseSpatialAttention (softmax over the 3600 attention channels + bmm) against the live reference"""
    from oracle import torch_ref
    from torchseg_b200.networks import PSANet
    x, y, seed = psanet_case()
    torch.manual_seed(seed)
    m = PSANet(150, torch.nn.CrossEntropyLoss(ignore_index=-1))
    g = GOLD["psanet_r101"]
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert sum(p.numel() for p in m.parameters()) == g["n_params"] and len(sd) == g["n_state"]
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    loss, _ = torch_ref.psanet_loss(x, y, sd)
    assert abs(float(loss) - g["loss"]) < 1e-5 * g["loss"]
    loss.backward()
    for n, ref in g["grad_norms"].items():
        assert abs(float(sd[n].grad.norm()) - ref) < 1e-4 * ref, n


def test_exp_det_properties():
    from oracle import c_oracle
    assert c_oracle.exp_det(0.0) == 1.0
    assert c_oracle.exp_det(-200.0) == 0.0
    xs = np.linspace(-80, 0, 2001).astype(np.float32)
    e = np.array([c_oracle.exp_det(float(v)) for v in xs])
    r = np.exp(xs.astype(np.float64))
    assert np.max(np.abs(e - r) / r) < 4e-7
    assert np.all(np.diff(e) >= 0)  # monotone


def test_bilinear_oracle_matches_aten():
    from oracle import c_oracle
    g = torch.Generator().manual_seed(4)
    lo = torch.randn(2, 6, 9, 32, generator=g)
    ref = torch.nn.functional.interpolate(lo[..., :19].permute(0, 3, 1, 2), scale_factor=8, mode="bilinear", align_corners=True)
    out = c_oracle.bilinear_nhwc_to_nchw(lo.numpy(), 19, 48, 72)
    assert np.max(np.abs(out - ref.numpy())) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# live reference (authoring container only)
# ---------------------------------------------------------------------------------------------------------------
def _live():
    from oracle import reference_live as rl
    if not rl.available():
        pytest.skip("/root/reference not present (GPU box): pinned by tests/golden instead")
    return rl


def test_oracle_vs_live_reference_ohem():
    rl = _live()
    from oracle import c_oracle
    lo = rl.load_loss_opr()
    for regime in ["A", "B", "C", "D"]:
        for seed in (1, 2, 3):
            logits, labels, min_kept = ohem_case(regime, seed=seed, H=32, W=40)
            crit = lo.ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=min_kept, use_weight=False)
            x = logits.clone().requires_grad_(True)
            loss = crit(x, labels)
            loss.backward()
            kept_ref = (x.grad.abs().sum(dim=1) > 0).reshape(-1).numpy()
            out = c_oracle.ohem(logits.numpy(), labels.numpy(), 255, 0.7, min_kept, want_grad=True)
            assert np.array_equal(out["kept"], kept_ref), (regime, seed)
            assert abs(out["loss"] - float(loss)) < 1e-5 * abs(float(loss))
            assert np.max(np.abs(out["dlogits"] - x.grad.numpy())) < 1e-6


def test_oracle_vs_live_reference_bisenet_exact():
    rl = _live()
    from oracle import torch_ref
    net = rl.load_network('bisenet/cityscapes.bisenet.R18')
    lo = rl.load_loss_opr()
    x, y, min_kept, seed = bisenet_case(N=2, HW=64, seed=5)
    torch.manual_seed(seed)
    crit = lo.ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=min_kept, use_weight=False)
    m = net.BiSeNet(19, True, crit, None, torch.nn.BatchNorm2d)
    m.train()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    loss_ref = m(x, y)
    loss_ref.backward()
    loss, _ = torch_ref.bisenet_r18_loss(x, y, sd, min_kept)
    loss.backward()
    assert float(loss) == float(loss_ref)
    for n, p in m.named_parameters():
        assert torch.equal(p.grad, sd[n].grad), n


def test_module_key_names_match_reference():
    """checkpoint compatibility: our BiSeNet exposes exactly the reference's state_dict keys and shapes (App. D)"""
    rl = _live()
    from torchseg_b200.networks import BiSeNet
    net = rl.load_network('bisenet/cityscapes.bisenet.R18')
    ref = net.BiSeNet(19, True, None, None, torch.nn.BatchNorm2d).state_dict()
    ours = BiSeNet(19, True, None, None, torch.nn.BatchNorm2d).state_dict()
    assert list(ref.keys()) == list(ours.keys())
    for k in ref:
        assert ref[k].shape == ours[k].shape, k


def test_metric_functions_vs_live_reference():
    """seg_opr.metric (confusion matrix / IoU / ADE intersection-union) against the reference's metric.py on seeded
    predictions, incl. ignore labels (255 → outside [0, n_cl), -1 for ADE) and classes that never occur"""
    import importlib.util
    from torchseg_b200.seg_opr import metric as mine
    rl = _live()
    spec = importlib.util.spec_from_file_location("ref_metric", os.path.join(rl.REF, "furnace", "seg_opr", "metric.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(5)
    n_cl = 19
    gt = rng.integers(0, 17, size=(3, 64, 80))          # classes 17, 18 never occur → NaN IoU, skipped by nanmean
    gt[:, :5] = 255
    pred = np.where(rng.random(gt.shape) < 0.7, np.clip(gt, 0, 16), rng.integers(0, 17, size=gt.shape))
    h1, l1, c1 = mine.hist_info(n_cl, pred, gt)
    h2, l2, c2 = ref.hist_info(n_cl, pred, gt)
    assert np.array_equal(h1, h2) and l1 == l2 and c1 == c2
    for a, b in zip(mine.compute_score(h1, c1, l1), ref.compute_score(h2, c2, l2)):
        assert np.allclose(a, b, equal_nan=True)
    ht, lt, ct = mine.hist_info_torch(n_cl, torch.from_numpy(pred), torch.from_numpy(gt))
    assert np.array_equal(ht.numpy(), h2) and int(lt) == l2 and int(ct) == c2
    # ADE protocol
    lab = rng.integers(-1, 150, size=(48, 64))
    pr = np.where(rng.random(lab.shape) < 0.6, np.clip(lab, 0, 149), rng.integers(0, 150, size=lab.shape))
    i1, u1 = mine.intersectionAndUnion(pr, lab, 150)
    i2, u2 = ref.intersectionAndUnion(pr, lab, 150)
    assert np.array_equal(i1, i2) and np.array_equal(u1, u2)
    for a, b in zip(mine.meanIoU(np.stack([i1, i1], 1), np.stack([u1, u1], 1)), ref.meanIoU(np.stack([i2, i2], 1), np.stack([u2, u2], 1))):
        assert np.allclose(a, b, equal_nan=True)
    for a, b in zip(mine.pixelAccuracy(pr, lab), ref.pixelAccuracy(pr, lab)):
        assert a == b
    assert mine.mean_pixel_accuracy([3, 4], [5, 6]) == ref.mean_pixel_accuracy([3, 4], [5, 6])
    assert mine.accuracy(pr, lab)[0] == ref.accuracy(pr, lab)[0]


def test_evaluator_sliding_and_whole_eval_vs_live_reference(monkeypatch):
    """engine.evaluator (batched sliding-window evaluation) against the reference Evaluator run on the CPU with a small
    deterministic network: same window grid, same summed (not averaged) overlap handling, same flip / multi-scale
    accumulation, same cv2 resizes → identical arg-max maps and matching scores"""
    rl = _live()
    ref_mod = rl.load_evaluator()
    from torchseg_b200.engine.evaluator import Evaluator
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self, raising=False)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.c1 = torch.nn.Conv2d(3, 8, 3, padding=1)
            self.c2 = torch.nn.Conv2d(8, 5, 3, padding=2, dilation=2)

        def forward(self, x):
            return torch.log_softmax(self.c2(torch.relu(self.c1(x))) * 3.0, dim=1)

    class RefWrap(object):       # the reference calls .eval() / .to(get_device()) (-1 on the CPU) on its val_func
        def __init__(self, m):
            self.m = m

        def eval(self):
            self.m.eval()

        def to(self, *a):
            return self

        def __call__(self, x):
            return self.m(x)

    net = Net().double().float()
    mean, std = np.array([0.485, 0.456, 0.406]), np.array([0.229, 0.224, 0.225])
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(150, 210, 3), dtype=np.uint8)
    kw = dict(dataset=None, class_num=5, image_mean=mean, image_std=std, network=None, multi_scales=[0.75, 1.0, 1.5],
              is_flip=True, devices=[0])
    mine = Evaluator(crop_batch=3, **kw)
    mine.val_func = net

    class _DS(object):
        def get_length(self):
            return 0
    kw_ref = dict(kw, dataset=_DS())
    ref = ref_mod.Evaluator(**kw_ref)
    ref.val_func = RefWrap(net)
    # sliding evaluation: crop 96 on a 150x210 image → 2x3 .. 3x5 windows per scale, incl. the single-window scale path
    p_ref = ref.sliding_eval(img, 96, 2 / 3.0, None)
    p_mine = mine.sliding_eval(img, 96, 2 / 3.0, None)
    assert p_ref.shape == p_mine.shape == (150, 210)
    assert (p_ref == p_mine).mean() > 0.999
    s_ref = ref.scale_process(img, (150, 210), 96, 2 / 3.0, None)
    s_mine = mine.scale_process(img, (150, 210), 96, 2 / 3.0, None)
    assert np.abs(s_ref - s_mine).max() < 1e-4 * np.abs(s_ref).max()
    assert mine.window_grid(150, 210, 96, 2 / 3.0) == [(0, 0), (0, 64), (0, 114), (54, 0), (54, 64), (54, 114)]
    # small image → single padded window
    small = img[:80, :70]
    assert np.array_equal(ref.sliding_eval(small, 96, 2 / 3.0, None), mine.sliding_eval(small, 96, 2 / 3.0, None))
    # whole-image evaluation with padding to a fixed input size and resize to an output size
    w_ref = ref.whole_eval(img, (75, 105), (160, 224), None)
    w_mine = mine.whole_eval(img, (75, 105), (160, 224), None)
    assert (w_ref == w_mine).mean() > 0.999
    # grey-scale input path
    g = img[:, :, :1]
    assert (ref.whole_eval(g, None, None, None) == mine.whole_eval(g, None, None, None)).mean() > 0.999


def test_oracle_dfn_blocks_vs_live_reference_modules():
    """RefineResidual / BNRefine / ChannelAttention restatements against the live reference modules (bit-exact, fp32)"""
    from oracle import torch_ref as tr
    rl = _live()
    so = rl.load_seg_oprs()
    torch.manual_seed(3)
    x = torch.randn(2, 16, 9, 11)
    for relu in (True, False):
        m = so.BNRefine(16, 16, 3, has_relu=relu).train()
        sd = {"m." + k: v.detach().clone() for k, v in m.state_dict().items()}
        assert torch.equal(tr.bn_refine(x, sd, "m", relu, 1e-5, 0.1, True), m(x))
        m = so.RefineResidual(16, 24, 3, has_relu=relu).train()
        sd = {"m." + k: v.detach().clone() for k, v in m.state_dict().items()}
        assert torch.equal(tr.refine_residual(x, sd, "m", relu, 1e-5, 0.1, True), m(x))
    m = so.ChannelAttention(32, 16, 1)
    sd = {"m." + k: v.detach().clone() for k, v in m.state_dict().items()}
    x2 = torch.randn(2, 16, 9, 11)
    assert torch.allclose(tr.channel_attention(x, x2, sd, "m"), m(x, x2), atol=1e-6)
