"""Checkpoints WRITTEN BY THE REFERENCE'S OWN CODE, committed as fixtures.

Builds the reference's unmodified ``methods/DKT.py::DKT(backbone.Conv4, ...)`` and ``methods/DKT_regression.py::DKT(
backbone.Conv3())`` from /root/reference (real ``backbone.py``; GPyTorch replaced by oracle/gpytorch_standin, whose
module tree / parameter names / shapes follow GPyTorch 1.0.1), gives every parameter and buffer a seeded non-trivial
value, and saves them exactly the way the reference does:

  * classification: ``torch.save({'epoch': epoch, 'state': model.state_dict()}, outfile)``        (train.py:57-65)
  * regression:     ``model.save_checkpoint(path)``                                                  (DKT_regression.py:99-104)

Next to each checkpoint the reference model's own outputs with those weights are recorded (``get_logits`` / ``correct``
on a seeded episode; regression ``test_loop``-style prediction), so that tests/test_compat.py (key set, strict load,
tensor equality; CPU) and tests/test_dkt_gpu.py::test_reference_checkpoint_* (same outputs from the CUDA path after
``load_state_dict``; GPU) can check that a reference-trained model is usable as is.

Run:  python tests/golden/make_golden_state.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def perturb(module, seed):
    g = torch.Generator().manual_seed(seed)
    seen = set()
    with torch.no_grad():
        for name, p in list(module.named_parameters()) + list(module.named_buffers()):
            if id(p) in seen or not p.is_floating_point():
                continue
            seen.add(id(p))
            if name.endswith("running_var"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif name.endswith("raw_noise") and not p.requires_grad:
                continue                                          # the classifier's fixed noise 0.1 (DKT.py:346-347)
            elif "raw_variance" in name and not p.requires_grad:
                continue                                          # cossim / bncossim: variance := 1, frozen
            elif "raw_lengthscale" in name:
                p.copy_(30.0 + 5.0 * torch.rand(p.shape, generator=g))
            elif "raw_mixture" in name:
                p.copy_(-4.0 + 0.3 * torch.randn(p.shape, generator=g))
            else:
                p.add_(0.1 * torch.randn(p.shape, generator=g))


def episode(seed, n_way, per_class, image):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_way, per_class, 3, image, image, generator=g)
    return x + 0.5 * torch.randn(n_way, 1, 3, 1, 1, generator=g)


def main():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "gpytorch_standin"))
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    import backbone as ref_backbone
    import methods.DKT as ref_dkt
    import methods.DKT_regression as ref_reg
    assert ref_dkt.__file__.startswith(REF) and ref_reg.__file__.startswith(REF) and ref_backbone.__file__.startswith(REF)
    torch.set_num_threads(1)

    for kernel in ("bncossim", "rbf"):
        ref_dkt.kernel_type = kernel
        torch.manual_seed(1)
        model = ref_dkt.DKT(ref_backbone.Conv4, n_way=5, n_support=1)
        perturb(model, seed=7)
        path = os.path.join(HERE, "ref_checkpoint_cls_%s.tar" % kernel)
        torch.save({'epoch': 3, 'state': model.state_dict()}, path)               # train.py:57-65
        model.eval()
        x = episode(31, 5, 1 + 4, 84)
        out = {"logits": model.get_logits(x).detach().numpy(), "correct": np.array(model.correct(x), dtype=np.float64),
               "keys": np.array(sorted(model.state_dict().keys()))}
        np.savez_compressed(os.path.join(HERE, "ref_checkpoint_cls_%s.npz" % kernel), **out)
        print(kernel, len(out["keys"]), "keys", os.path.getsize(path), "bytes", "correct", out["correct"])

    for kernel in ("rbf", "spectral"):
        ref_reg.kernel_type = kernel
        torch.manual_seed(1)
        model = ref_reg.DKT(ref_backbone.Conv3())
        perturb(model, seed=9)
        path = os.path.join(HERE, "ref_checkpoint_reg_%s.tar" % kernel)
        model.save_checkpoint(path)                                                # DKT_regression.py:99-104
        g = torch.Generator().manual_seed(41)
        x_all = torch.rand(19, 3, 100, 100, generator=g)
        y_all = torch.rand(19, generator=g) * 2 - 1
        ind = [12, 2, 1, 16, 4]
        # DKT_regression.py:83-95 with fixed draws
        z_support = model.feature_extractor(x_all[ind]).detach()
        model.model.set_train_data(inputs=z_support, targets=y_all[ind], strict=False)
        model.model.eval(); model.feature_extractor.eval(); model.likelihood.eval()
        with torch.no_grad():
            pred = model.likelihood(model.model(model.feature_extractor(x_all).detach()))
            lower, upper = pred.confidence_region()
        out = {"mean": pred.mean.numpy(), "lower": lower.numpy(), "upper": upper.numpy(), "support_ind": np.array(ind),
               "gp_keys": np.array(sorted(model.model.state_dict().keys())),
               "likelihood_keys": np.array(sorted(model.likelihood.state_dict().keys())),
               "net_keys": np.array(sorted(model.feature_extractor.state_dict().keys()))}
        np.savez_compressed(os.path.join(HERE, "ref_checkpoint_reg_%s.npz" % kernel), **out)
        print("regression", kernel, os.path.getsize(path), "bytes", "mean", out["mean"][:4])


if __name__ == "__main__":
    main()
