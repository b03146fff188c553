"""Generate golden vectors from the UNMODIFIED reference ``backbone.py`` (imported from
/root/reference, which exists only in the authoring container -- fixtures are committed).

For every architecture on the hot path: draw parameters with oracle.backbone.init_params(seed),
load them into the reference's own module (``load_state_dict``), run the reference forward in
train mode (batch statistics + running-stat update) and eval mode on seeded inputs, and store
ONLY the outputs (+ a parameter checksum so RNG drift is detected).  tests/test_oracle.py
regenerates parameters/inputs from the same seeds and requires oracle.backbone to match.

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

CASES = [  # arch, n images, image size, bn_out
    ("Conv4", 6, 84, True),
    ("Conv4", 5, 84, False),
    ("Conv6", 4, 84, False),
    ("ResNet10", 3, 224, False),
    ("ResNet18", 3, 224, True),
    ("ResNet50", 2, 224, False),
    ("Conv3", 5, 100, False),
]


def case_inputs(arch, n, size, seed=7):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 3, size, size, generator=g)


def param_checksum(p):
    s = 0.0
    for k in sorted(p):
        if p[k].is_floating_point():
            s += float(p[k].double().abs().sum())
    return s


def main():
    sys.path.insert(0, REF)
    import backbone as ref_backbone          # the reference's own file
    sys.path.pop(0)
    from oracle import backbone as obb

    for arch, n, size, bn_out in CASES:
        torch.manual_seed(0)
        p = obb.init_params(arch, seed=3, bn_out=bn_out)
        m = getattr(ref_backbone, arch)()
        if bn_out:   # what DKT.__init__ does (methods/DKT.py:45-48)
            m.trunk.add_module("bn_out", torch.nn.BatchNorm1d(int(np.prod(m.final_feat_dim))))
        sd = m.state_dict()
        new = {}
        for k in sd:
            kk = k
            if kk not in p and ".trunk." in kk[6:]:
                # ConvBlock registers C/BN twice: trunk.i.C.* and trunk.i.trunk.{0,1}.* (backbone.py:115-127)
                head, tail = kk.split(".trunk.", 1)
                idx, rest = tail.split(".", 1)
                kk = head + (".C." if idx == "0" else ".BN.") + rest
            assert kk in p, (arch, k)
            new[k] = p[kk].clone()
        m.load_state_dict(new)
        x = case_inputs(arch, n, size)
        m.train()
        out_train = m(x).detach()
        sd_after = {k: v.clone() for k, v in m.state_dict().items() if "running" in k}
        m.eval()
        with torch.no_grad():
            out_eval = m(x)
        name = "backbone_%s%s.npz" % (arch, "_bnout" if bn_out else "")
        # keep a couple of running-stat vectors (first BN and last BN) to pin the update rule
        rk = sorted(k for k in sd_after if ".trunk." not in k[6:])
        keep = {("stat_" + k): sd_after[k].numpy() for k in (rk[:2] + rk[-2:])}
        np.savez_compressed(os.path.join(HERE, name), out_train=out_train.numpy(), out_eval=out_eval.numpy(),
                            checksum=np.float64(param_checksum(p)), n=n, size=size, **keep)
        print(name, tuple(out_train.shape), "checksum %.6f" % param_checksum(p))


if __name__ == "__main__":
    main()
