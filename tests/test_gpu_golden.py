"""CUDA path against the committed golden vectors (tests/golden/*.npz, produced by the
float64 oracle) at the named configurations of SURVEY.md 8(d).  Tolerance 1e-4 relative
(north_star); per-tensor max|a-b| / max|b|."""
import os

import numpy as np
import pytest
import torch

from oracle import cbfssm_oracle as O      # only for PARAM_NAMES / shapes (checker side)
from tests.helpers import NAMED_CASES, named_case, rel_inf
from tests.test_gpu_parity import run_engine

pytestmark = pytest.mark.gpu
TOL = 1e-4
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("flags", [12, 1], ids=["register_or_tensor", "cooperative"])
@pytest.mark.parametrize("name", list(NAMED_CASES))
def test_named_configuration_matches_golden(name, flags):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    cfg, params, u, y, eps_b, z_b, eps_f, cond = named_case(name)
    eng, out, yd = run_engine(cfg, params, u, y, eps_b, z_b, eps_f, cond, flags)
    for k in ("loss", "loglik", "kl_x", "entropy", "kl_z_f", "kl_z_b"):
        ref, got = float(gold[k]), float(out[k])
        assert abs(got - ref) <= TOL * max(abs(ref), 1e-3), (k, got, ref)
    xf, yt = eng.export_states(yd)
    pm, pv = eng.moments(xf, cfg.dim_y, eng.var_y)
    im, iv = eng.moments(xf, cfg.dim_x, None)
    torch.cuda.synchronize()
    assert rel_inf(xf.cpu().numpy()[:, ::15, ::10, :], gold["x_final_sample"]) < TOL
    assert rel_inf(yt.cpu().numpy()[:, ::15, ::10, :], gold["y_tilde_sample"]) < TOL
    assert rel_inf(pm.cpu().numpy(), gold["pred_mean"]) < TOL
    assert rel_inf(pv.cpu().numpy(), gold["pred_var"]) < TOL
    assert rel_inf(im.cpu().numpy(), gold["internal_mean"]) < TOL
    assert rel_inf(iv.cpu().numpy(), gold["internal_var"]) < TOL
    grads = eng.get_grads()
    worst = {k: rel_inf(grads[k], gold["grad." + k]) for k in O.PARAM_NAMES}
    bad = {k: v for k, v in worst.items() if not v < TOL}
    assert not bad, bad
