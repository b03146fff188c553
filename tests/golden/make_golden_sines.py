"""Golden vectors for BASELINE configs[0] (sines), produced by RUNNING THE REFERENCE'S OWN, UNMODIFIED
``sines/train_DKT.py::main()`` (imported from /root/reference; it exists only in the authoring container).

The script needs gpytorch, matplotlib and seaborn, none of which is installable offline: gpytorch is
oracle/gpytorch_standin (independent scipy / LAPACK arithmetic, see its header), matplotlib / seaborn are inert stub
modules (the plots at the end of ``main()`` are drawn into mocks).  ``main()`` is executed as written -- 50 000 training
iterations (sines/train_DKT.py:171-180), 500 test tasks (199-229), 10 plots -- and observed from outside:

  * ``Task_Distribution`` / ``Sine_Task.sample_data`` are wrapped to record the data of the first training tasks and of
    the first test tasks;
  * ``torch.optim.Adam.step`` is wrapped to snapshot the network + GP state before the first steps;
  * ``ExactMarginalLogLikelihood.forward`` is wrapped to record the first losses, ``likelihood(gp(z))`` in eval mode to
    record the predictive mean / confidence region of the first test tasks, and ``print`` captures the final
    "Average MSE" line.

Committed: tests/golden/sines_reference.npz (inputs, initial / per-step parameters, losses, trained state, test-task
predictions, the reference's own average test MSE).  tests/test_oracle.py replays oracle/episode.py::OracleSines and
tests/test_dkt_gpu.py the CUDA ``SinesDKT`` against it.

Run:  python tests/golden/make_golden_sines.py      (about two minutes)
"""
import builtins
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
N_TRAIN_REC, N_TEST_REC = 5, 6


def main():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "gpytorch_standin"))
    plt = mock.MagicMock()
    plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    sys.modules["seaborn"] = mock.MagicMock()
    sys.path.insert(0, os.path.join(REF, "sines"))
    import gpytorch
    import train_DKT as ref
    assert ref.__file__.startswith(REF)

    rec = {"train_x": [], "train_y": [], "loss": [], "state": [], "test": [], "lines": []}
    state = {"net": None, "gp": None, "lik": None, "steps": 0, "phase": "train", "n_sample": 0}

    # -- observe the data the script draws
    sample_orig = ref.Sine_Task.sample_data

    def sample_data(self, size=1, noise=0.0, sort=False):
        x, y = sample_orig(self, size, noise, sort)
        if state["phase"] == "train" and len(rec["train_x"]) < N_TRAIN_REC:
            rec["train_x"].append(x.numpy().copy())
            rec["train_y"].append(y.numpy().copy())
        if state["phase"] == "test":
            state["last_all"] = (x.numpy().copy(), y.numpy().copy())
        return x, y
    ref.Sine_Task.sample_data = sample_data

    # -- observe the modules it builds
    feat_init = ref.Feature.__init__

    def feature_init(self):
        feat_init(self)
        state["net"] = self
    ref.Feature.__init__ = feature_init
    gp_init = ref.ExactGPModel.__init__

    def gpmodel_init(self, train_x, train_y, likelihood):
        gp_init(self, train_x, train_y, likelihood)
        state["gp"], state["lik"] = self, likelihood
    ref.ExactGPModel.__init__ = gpmodel_init

    def snapshot():
        out = {"net." + k: v.detach().numpy().copy() for k, v in state["net"].state_dict().items()}
        out.update({"gp." + k: v.detach().numpy().copy() for k, v in state["gp"].state_dict().items()})
        return out

    # -- losses of the first steps; parameters before each of them
    mll_forward = gpytorch.mlls.ExactMarginalLogLikelihood.forward

    def forward(self, output, target):
        v = mll_forward(self, output, target)
        if len(rec["loss"]) < N_TRAIN_REC:
            rec["state"].append(snapshot())
            rec["loss"].append(float(-v.detach()))
        return v
    gpytorch.mlls.ExactMarginalLogLikelihood.forward = forward

    # -- test phase: likelihood(gp(z_query)) in eval mode
    lik_forward = gpytorch.likelihoods.GaussianLikelihood.forward

    def lik_fwd(self, dist):
        out = lik_forward(self, dist)
        if not state["gp"].training and len(rec["test"]) < N_TEST_REC and out.mean.shape[0] == 195:
            lo, hi = out.confidence_region()
            gp = state["gp"]
            rec["test"].append({"x_all": state["last_all"][0], "y_all": state["last_all"][1],
                                "z_support": gp.train_inputs[0].detach().numpy().copy(),
                                "y_support": gp.train_targets.detach().numpy().copy(),
                                "mean": out.mean.detach().numpy().copy(), "lower": lo.detach().numpy().copy(),
                                "upper": hi.detach().numpy().copy()})
        return out
    gpytorch.likelihoods.GaussianLikelihood.forward = lik_fwd

    real_print = builtins.print

    def tee(*a, **k):
        s = " ".join(str(x) for x in a)
        if s.startswith("Test, please wait"):
            state["phase"] = "test"
            rec["trained"] = snapshot()
        if "Average MSE" in s or s.startswith("[0]") or s.startswith("[49900]"):
            rec["lines"].append(s)
            real_print(s)
    ref.print = tee                                     # module-level name lookup: builtins stay untouched elsewhere

    np.random.seed(0)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    cwd = os.getcwd()
    os.chdir("/tmp")                                    # plt.savefig is a mock, but keep any stray file out of the repo
    try:
        ref.main()                                      # the reference's script, unmodified
    finally:
        os.chdir(cwd)

    out = {"train_x": np.stack(rec["train_x"]), "train_y": np.stack(rec["train_y"]), "loss": np.array(rec["loss"])}
    for i, s in enumerate(rec["state"]):
        for k, v in s.items():
            out["step%d.%s" % (i, k)] = v
    for k, v in rec["trained"].items():
        out["trained." + k] = v
    for i, t in enumerate(rec["test"]):
        for k, v in t.items():
            out["test%d.%s" % (i, k)] = v
    avg = [l for l in rec["lines"] if "Average MSE" in l][0]
    out["average_mse"] = np.float64(avg.split("Average MSE:")[1].split("+-")[0])
    out["average_mse_std"] = np.float64(avg.split("+-")[1])
    np.savez_compressed(os.path.join(HERE, "sines_reference.npz"), **out)
    real_print("saved", len(out), "arrays; losses", out["loss"], "average MSE", out["average_mse"])


if __name__ == "__main__":
    main()
