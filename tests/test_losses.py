"""Training losses against the reference: CharbonnierLoss closed form and ContrastLoss vs the golden
fixture produced by the reference's own My_CR.ContrastLoss (oracle/make_golden_cr.py; seeded random-init VGG19, fp32 CPU)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "contrast_loss.npz")


def test_charbonnier_matches_reference_formula():
    from lewin_b200.losses import CharbonnierLoss
    g = torch.Generator().manual_seed(0)
    x, y = torch.rand(2, 3, 16, 16, generator=g), torch.rand(2, 3, 16, 16, generator=g)
    want = np.mean(np.sqrt((x.numpy().astype(np.float64) - y.numpy()) ** 2 + 1e-6))      # losses.py:49-51, eps = 1e-3
    assert abs(float(CharbonnierLoss()(x, y)) - want) < 1e-6


def test_contrast_loss_matches_reference_golden():
    from lewin_b200.losses import ContrastLoss
    z = np.load(GOLDEN)
    for name, ab in (("full", False), ("ablation", True)):
        torch.manual_seed(int(z["seed"]))
        crit = ContrastLoss(ablation=ab, pretrained=False)
        a = torch.from_numpy(z["a"]).requires_grad_(True)
        loss, ap, an = crit(a, torch.from_numpy(z["p"]), torch.from_numpy(z["n"]))
        loss.backward()
        assert abs(float(loss) - float(z[name + "_loss"])) < 1e-5 * max(1.0, abs(float(z[name + "_loss"])))
        assert abs(float(ap) - float(z[name + "_ap"])) < 1e-6
        assert abs(float(an) - float(z[name + "_an"])) < 1e-6
        ref = z[name + "_da"]
        # d_ap / d_an is a ratio of nearly equal distances at random init: its gradient is a difference of two close terms,
        # which amplifies the fp32 summation-order differences of the channels-last convolutions (measured 3e-3 of max)
        got = a.grad.numpy()
        assert np.abs(got - ref).max() < 5e-2 * np.abs(ref).max()          # + isolated sign(fa - fp) flips of the L1 terms
        cos = float((got.ravel() @ ref.ravel()) / (np.linalg.norm(got) * np.linalg.norm(ref)))
        assert cos > 0.9999
        assert all(not p.requires_grad for p in crit.parameters())                         # frozen VGG (My_CR.py:78-80)


def test_vgg_slices_keep_reference_state_dict_keys():
    from lewin_b200.losses import Vgg19
    keys = list(Vgg19(pretrained=False).state_dict().keys())
    conv_idx = [0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28]                            # convs of features[0:30]
    cut = {0: 1, 2: 2, 5: 2, 7: 3, 10: 3, 12: 4, 14: 4, 16: 4, 19: 4, 21: 5, 23: 5, 25: 5, 28: 5}
    want = [f"slice{cut[i]}.{i}.{w}" for i in conv_idx for w in ("weight", "bias")]
    assert keys == want
