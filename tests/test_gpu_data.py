"""Device input pipeline (tsb_train_preprocess behind torchseg_b200.utils.gpu_pipeline.TrainPreGPU) against the oracle
restatement of the reference's cv2 pipeline and against the fixtures generated from the LIVE reference TrainPre —
bit-exact (float32 bit patterns and labels)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest
import torch

from golden_cases import pipeline_case, pipeline_case_dfn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "data_pipeline.json")))


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.cpu().numpy()).tobytes()).hexdigest()


def test_pipeline_matches_live_reference_golden(cuda):
    from torchseg_b200.utils.gpu_pipeline import TrainPreGPU
    bgr, gt, crop, scales, mean, std = pipeline_case()
    pre = TrainPreGPU(mean, std, crop, scales, cuda, bgr_input=True)
    img_d, gt_d = torch.from_numpy(bgr).to(cuda), torch.from_numpy(gt).to(cuda)
    seeds = sorted(GOLD["cases"], key=int)
    params = []
    for seed in seeds:                       # one batch with all the seeds: per-sample parameters differ inside a launch
        random.seed(int(seed))
        params.append(pre.draw(bgr.shape[:2]))
    data, label = pre([img_d] * len(seeds), [gt_d] * len(seeds), params)
    assert tuple(data.shape) == (len(seeds), 3) + tuple(crop) and data.dtype == torch.float32 and label.dtype == torch.int64
    for i, seed in enumerate(seeds):
        ent = GOLD["cases"][seed]
        assert _sha(data[i]) == ent["data_sha256"], seed
        assert _sha(label[i]) == ent["label_sha256"], seed


@pytest.mark.parametrize("seed", [0, 1, 3, 5])
def test_pipeline_cityscapes_frame_vs_oracle(cuda, seed):
    """a full 1024 x 2048 frame → 1024 x 1024 crop (BASELINE configs[1] input), every scale of config.py:86"""
    from oracle import data_ref
    from torchseg_b200.utils.gpu_pipeline import TrainPreGPU
    _, _, _, scales, mean, std = pipeline_case()
    rng = np.random.default_rng(100 + seed)
    bgr = rng.integers(0, 256, (1024, 2048, 3), dtype=np.uint8)
    gt = rng.integers(0, 19, (1024, 2048), dtype=np.uint8)
    gt[:100] = 255
    crop = (1024, 1024)
    pre = TrainPreGPU(mean, std, crop, scales, cuda, bgr_input=True)
    random.seed(seed)
    prm = pre.draw(bgr.shape[:2])
    random.seed(seed)
    ref_prm = data_ref.draw_params(bgr.shape[:2], crop, scales)
    data, label = pre([torch.from_numpy(bgr).to(cuda)], [torch.from_numpy(gt).to(cuda)], [prm])
    ref_d, ref_l = data_ref.train_pre(bgr[:, :, ::-1], gt, ref_prm, crop, mean, std)
    assert np.array_equal(data[0].cpu().numpy().view(np.uint32), ref_d.view(np.uint32)), ref_prm
    assert np.array_equal(label[0].cpu().numpy(), ref_l)


def test_pipeline_feeds_the_network(cuda):
    """the batch goes straight into BiSeNet.forward (NCHW fp32 'data', int64 'label' — train.py:119-124)"""
    from torchseg_b200.utils.gpu_pipeline import TrainPreGPU
    bgr, gt, crop, scales, mean, std = pipeline_case(H=160, W=256, crop=(128, 128))
    pre = TrainPreGPU(mean, std, crop, scales, cuda)
    random.seed(2)
    data, label = pre([torch.from_numpy(bgr).to(cuda)] * 2, [torch.from_numpy(gt).to(cuda)] * 2)
    assert torch.isfinite(data).all() and int(label.max()) <= 255 and int(label.min()) >= 0
    assert abs(float(data.mean())) < 1.0 and 0.5 < float(data.std()) < 2.0


def test_dfn_pipeline_edge_labels_match_live_reference_golden(cuda):
    """DFN TrainPre on the device: data / label / aux_label (Canny apertureSize 7 + 7x7 dilate border map) bit-exact vs
    the fixtures generated from the LIVE reference (dfn dataloader.py:11-44)"""
    from torchseg_b200.utils.gpu_pipeline import TrainPreGPU
    bgr, gt, crop, scales, mean, std = pipeline_case_dfn()
    pre = TrainPreGPU(mean, std, crop, scales, cuda, bgr_input=True, edge_labels=True)
    img_d, gt_d = torch.from_numpy(bgr).to(cuda), torch.from_numpy(gt).to(cuda)
    seeds = sorted(GOLD["cases_dfn"], key=int)
    params = []
    for seed in seeds:
        random.seed(int(seed))
        params.append(pre.draw(bgr.shape[:2]))
    data, label, aux = pre([img_d] * len(seeds), [gt_d] * len(seeds), params)
    assert aux.dtype == torch.int64 and tuple(aux.shape) == (len(seeds),) + tuple(crop)
    for i, seed in enumerate(seeds):
        ent = GOLD["cases_dfn"][seed]
        assert _sha(data[i]) == ent["data_sha256"] and _sha(label[i]) == ent["label_sha256"], seed
        assert int((aux[i] == 1).sum()) == ent["aux_ones"], (seed, int((aux[i] == 1).sum()), ent["aux_ones"])
        assert _sha(aux[i]) == ent["aux_sha256"], seed


@pytest.mark.parametrize("seed", [0, 2])
def test_dfn_pipeline_cityscapes_frame_vs_oracle(cuda, seed):
    """full 1024 x 2048 frame → 800 x 800 crop (dfn config.py:65-66), border labels vs the oracle"""
    from oracle import data_ref
    from torchseg_b200.utils.gpu_pipeline import TrainPreGPU
    _, _, _, scales, mean, std = pipeline_case_dfn()
    rng = np.random.default_rng(200 + seed)
    bgr = rng.integers(0, 256, (1024, 2048, 3), dtype=np.uint8)
    gt = np.zeros((1024, 2048), np.uint8)
    for _ in range(200):
        y0, x0 = rng.integers(0, 1024), rng.integers(0, 2048)
        gt[y0:y0 + rng.integers(8, 300), x0:x0 + rng.integers(8, 500)] = rng.integers(0, 19)
    gt[:60] = 255
    crop = (800, 800)
    pre = TrainPreGPU(mean, std, crop, scales, cuda, edge_labels=True)
    random.seed(seed)
    prm = pre.draw(bgr.shape[:2])
    random.seed(seed)
    ref_prm = data_ref.draw_params(bgr.shape[:2], crop, scales)
    data, label, aux = pre([torch.from_numpy(bgr).to(cuda)], [torch.from_numpy(gt).to(cuda)], [prm])
    ref_d, ref_l, ref_a = data_ref.train_pre_dfn(bgr[:, :, ::-1], gt, ref_prm, crop, mean, std)
    assert np.array_equal(label[0].cpu().numpy(), ref_l)
    assert np.array_equal(aux[0].cpu().numpy(), ref_a), (ref_prm, int((aux[0].cpu().numpy() != ref_a).sum()))
    assert np.array_equal(data[0].cpu().numpy().view(np.uint32), ref_d.view(np.uint32))


def test_pipeline_speed_config_label_downsampling(cuda):
    """cityscapes.bisenet.R18.speed: labels at 1/8 resolution (dataloader.py:28-30) vs the live-reference fixtures"""
    from torchseg_b200.utils.gpu_pipeline import TrainPreGPU
    bgr, gt, crop, scales, mean, std = pipeline_case()
    pre = TrainPreGPU(mean, std, crop, scales, cuda, gt_down_sampling=8)
    for seed, ent in GOLD["cases_speed"].items():
        random.seed(int(seed))
        data, label = pre([torch.from_numpy(bgr).to(cuda)], [torch.from_numpy(gt).to(cuda)])
        assert list(label.shape[1:]) == ent["label_shape"] and _sha(label[0]) == ent["label_sha256"], seed
        assert tuple(data.shape[2:]) == tuple(crop)
