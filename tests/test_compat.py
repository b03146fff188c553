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
