"""Host-side drop-in surface (no GPU): the repo-root shims in front of a reference checkout, the checkpoint file
helpers (io_utils.py:66-86) and loading reference-written checkpoints."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _model(kernel="bncossim", n_way=3):
    from deep_kernel_transfer_b200 import backbone
    from deep_kernel_transfer_b200.methods.DKT import DKT
    torch.manual_seed(0)
    return DKT(lambda: backbone.ConvNet(4, image_size=32), n_way, 1, kernel=kernel)


def _as_gpytorch(state, new_style):
    """Re-shape a state_dict of this package the way GPyTorch writes it: Interval constraint bounds as buffers, and for
    the >= 1.9 layout a scalar ``raw_constant`` and a [1] ``raw_outputscale``."""
    out = {}
    for k, v in state.items():
        if k.endswith("mean_module.constant") and new_style:
            out[k[:-len("constant")] + "raw_constant"] = v.reshape(())
            continue
        if k.endswith("raw_outputscale") and new_style:
            v = v.reshape(1)
        out[k] = v
        leaf = k.rsplit(".", 1)[-1]
        if leaf.startswith("raw_") and leaf != "raw_constant":
            out[k + "_constraint.lower_bound"] = torch.tensor(0.0)
            out[k + "_constraint.upper_bound"] = torch.tensor(float("inf"))
    return out


@pytest.mark.parametrize("kernel", ["bncossim", "rbf"])
@pytest.mark.parametrize("new_style", [False, True])
def test_load_reference_style_checkpoint(tmp_path, kernel, new_style):
    src = _model(kernel)
    with torch.no_grad():
        for p in src.parameters():
            p.add_(torch.randn_like(p) * 0.1)
        for b in src.buffers():
            if b.is_floating_point():
                b.add_(torch.rand_like(b))
    f = tmp_path / "best_model.tar"
    torch.save({"epoch": 7, "state": _as_gpytorch(src.state_dict(), new_style)}, str(f))       # train.py:57-65
    dst = _model(kernel)
    tmp = torch.load(str(f))
    dst.load_state_dict(tmp["state"])                                                             # test.py:125-126
    a, b = src.state_dict(), dst.state_dict()
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # aliases still share storage after loading (backbone.py:115-127; DKT.py:41)
    assert dst.feature.trunk[0].C.weight is dst.feature.trunk[0].trunk[0].weight
    assert dst.feature_extractor is dst.feature
    with pytest.raises(RuntimeError):
        dst.load_state_dict({**tmp["state"], "model.models.0.bogus": torch.zeros(1)})


GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("kernel", ["bncossim", "rbf"])
def test_checkpoint_written_by_the_reference_loads(kernel):
    """tests/golden/ref_checkpoint_cls_*.tar was written by the reference's own DKT(backbone.Conv4) via train.py's
    ``torch.save({'epoch', 'state': model.state_dict()})`` (tests/golden/make_golden_state.py).  The drop-in has the SAME
    key set (incl. the mll.mlls / model.likelihood aliases GPyTorch registers), strict-loads it, holds the same tensors."""
    import numpy as np
    from deep_kernel_transfer_b200 import backbone
    from deep_kernel_transfer_b200.methods.DKT import DKT
    ck = torch.load(os.path.join(GOLD, "ref_checkpoint_cls_%s.tar" % kernel))
    ref_keys = set(np.load(os.path.join(GOLD, "ref_checkpoint_cls_%s.npz" % kernel))["keys"].tolist())
    assert ck["epoch"] == 3 and set(ck["state"].keys()) == ref_keys
    dst = DKT(backbone.Conv4, 5, 1, kernel=kernel)
    assert set(dst.state_dict().keys()) == ref_keys, set(dst.state_dict().keys()) ^ ref_keys
    dst.load_state_dict(ck["state"])                                   # strict
    own = dst.state_dict()
    for k, v in ck["state"].items():
        assert own[k].shape == v.shape and torch.equal(own[k], v), k
    # the frozen pieces stay frozen, the aliases stay aliases
    assert not dst.model.models[0].likelihood.noise_covar.raw_noise.requires_grad
    assert dst.mll.mlls[2].model is dst.model.models[2] and dst.model.likelihood.likelihoods[1] is dst.likelihood.likelihoods[1]


@pytest.mark.parametrize("kernel", ["rbf", "spectral"])
def test_regression_checkpoint_written_by_the_reference_loads(kernel):
    """Same for DKT_regression.save_checkpoint (DKT_regression.py:99-104), written by the reference's own class."""
    import numpy as np
    from deep_kernel_transfer_b200 import backbone
    from deep_kernel_transfer_b200.methods.DKT_regression import DKT as DKTR
    path = os.path.join(GOLD, "ref_checkpoint_reg_%s.tar" % kernel)
    meta = np.load(os.path.join(GOLD, "ref_checkpoint_reg_%s.npz" % kernel))
    ck = torch.load(path)
    dst = DKTR(backbone.Conv3(), kernel=kernel)
    assert set(dst.model.state_dict().keys()) == set(meta["gp_keys"].tolist()) == set(ck["gp"].keys())
    assert set(dst.likelihood.state_dict().keys()) == set(meta["likelihood_keys"].tolist())
    assert set(dst.feature_extractor.state_dict().keys()) == set(meta["net_keys"].tolist())
    dst.load_checkpoint(path)
    for part, mod in (("gp", dst.model), ("likelihood", dst.likelihood), ("net", dst.feature_extractor)):
        own = mod.state_dict()
        for k, v in ck[part].items():
            assert torch.equal(own[k].reshape(v.shape), v), (part, k)


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (authoring container only)")
def test_reference_strict_loads_our_checkpoint(tmp_path):
    """The other direction: a checkpoint written by the drop-in strict-loads into the reference's own DKT / regression
    DKT (unmodified methods/*.py + backbone.py on the GPyTorch stand-in)."""
    code = (
        "import sys, types, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "torch.Tensor.cuda = lambda self, *a, **k: self; torch.nn.Module.cuda = lambda self, *a, **k: self\n"
        "sys.modules.setdefault('h5py', types.ModuleType('h5py'))\n"
        "import backbone, methods.DKT as D, methods.DKT_regression as R\n"
        "assert D.__file__.startswith(%r)\n"
        "D.kernel_type = 'bncossim'\n"
        "m = D.DKT(backbone.Conv4, n_way=5, n_support=1)\n"
        "ck = torch.load(%r)\n"
        "m.load_state_dict(ck['state'])\n"
        "R.kernel_type = 'spectral'\n"
        "r = R.DKT(backbone.Conv3())\n"
        "r.load_checkpoint(%r)\n" % (os.path.join(ROOT, "oracle", "gpytorch_standin"), REF, REF,
                                     str(tmp_path / "ours_cls.tar"), str(tmp_path / "ours_reg.tar")))
    from deep_kernel_transfer_b200 import backbone
    from deep_kernel_transfer_b200.methods.DKT import DKT
    from deep_kernel_transfer_b200.methods.DKT_regression import DKT as DKTR
    torch.manual_seed(0)
    torch.save({"epoch": 1, "state": DKT(backbone.Conv4, 5, 1, kernel="bncossim").state_dict()}, str(tmp_path / "ours_cls.tar"))
    DKTR(backbone.Conv3(), kernel="spectral").save_checkpoint(str(tmp_path / "ours_reg.tar"))
    r = subprocess.run([sys.executable, "-c", code], cwd="/tmp", capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]


def test_checkpoint_file_helpers(tmp_path):
    sys.path.insert(0, ROOT)
    import io_utils
    d = str(tmp_path)
    assert io_utils.get_resume_file(d) is None and io_utils.get_best_file(d) is None
    for n in ("3.tar", "12.tar", "best_model.tar"):
        open(os.path.join(d, n), "w").close()
    assert io_utils.get_resume_file(d) == os.path.join(d, "12.tar")
    assert io_utils.get_best_file(d) == os.path.join(d, "best_model.tar")
    assert io_utils.get_assigned_file(d, 3) == os.path.join(d, "3.tar")
    assert set(io_utils.model_dict) >= {"Conv4", "Conv6", "ResNet10", "ResNet18", "ResNet34", "ResNet50", "ResNet101"}


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (authoring container only)")
def test_reference_train_driver_imports_through_the_shims():
    """``train.py`` of the reference, unmodified, with this repo in front of it on sys.path: its imports resolve, the DKT
    method, the backbones and the episodic data manager are this package's, everything else is the reference's own."""
    code = ("import train, io_utils, backbone, methods.protonet as pn\n"
            "assert train.DKT.__module__ == 'deep_kernel_transfer_b200.methods.DKT', train.DKT.__module__\n"
            "assert train.SetDataManager.__module__ == 'deep_kernel_transfer_b200.episode_feed'\n"
            "assert train.SimpleDataManager.__module__ == '_reference_data_datamgr'\n"
            "assert io_utils.model_dict['Conv4'] is backbone.Conv4 and backbone.__file__.startswith(%r)\n"
            "assert pn.__file__.startswith(%r) and callable(io_utils.parse_args)\n"
            "m = train.DKT(io_utils.model_dict['Conv4'], n_way=5, n_support=1)\n"
            "assert m.feature.final_feat_dim == 1600\n" % (ROOT, REF))
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + REF)
    r = subprocess.run([sys.executable, "-c", code], cwd="/tmp", env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
