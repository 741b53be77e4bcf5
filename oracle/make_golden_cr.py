"""Golden fixture for the contrastive loss: runs the REFERENCE's My_CR.ContrastLoss (imported from /root/reference, CPU,
fp32) on a small seeded input and stores loss / all_ap / all_an and d loss / d a.

The reference constructor downloads ImageNet weights and calls .cuda(); neither exists in this container, so
`torchvision.models.vgg19` is wrapped to build the seeded random-init network (torch.manual_seed(SEED) right before
construction, the same call lewin_b200.losses.ContrastLoss(pretrained=False) makes) and `.cuda()` is a no-op.  Nothing of
the reference's arithmetic is touched.  Usage (this container only):  python oracle/make_golden_cr.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF_DIR", "/root/reference")
SEED = 11


def main():
    sys.path.insert(0, os.path.join(REF, "Uformer_ProbSparse"))
    from torchvision import models
    real = models.vgg19
    models.vgg19 = lambda pretrained=False, **kw: real(weights=None)
    torch.nn.Module.cuda = lambda self, device=None: self
    import My_CR
    out = {}
    for name, ab in (("full", False), ("ablation", True)):
        torch.manual_seed(SEED)
        crit = My_CR.ContrastLoss(ablation=ab)
        g = torch.Generator().manual_seed(SEED + 1)
        a = torch.rand(2, 3, 48, 48, generator=g).requires_grad_(True)
        p = torch.rand(2, 3, 48, 48, generator=g)
        n = torch.rand(2, 3, 48, 48, generator=g)
        loss, ap, an = crit(a, p, n)
        loss.backward()
        out[name + "_loss"] = np.float64(loss.item())
        out[name + "_ap"] = np.float64(float(ap))
        out[name + "_an"] = np.float64(float(an))
        out[name + "_da"] = a.grad.numpy().copy()
        if not ab:
            out["a"], out["p"], out["n"] = a.detach().numpy().copy(), p.numpy().copy(), n.numpy().copy()
    out["seed"] = np.int64(SEED)
    path = os.path.join(ROOT, "tests", "golden", "contrast_loss.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
