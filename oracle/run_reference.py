"""Run the reference's UNMODIFIED model source on one minibatch.  TEST INFRASTRUCTURE ONLY.

``/root/reference/cbfssm/model/{tf_transform,gp_tf,base_model,cbfssm,cbfssmhalf}.py`` are
imported as they are and executed under ``oracle/tf_shim`` (an eager float64 stand-in for the
TensorFlow-1.8 symbols they use; TensorFlow itself cannot be installed here).  This is how the
restated oracle (``oracle/cbfssm_oracle.py``) is pinned: ``tests/golden/make_golden.py`` calls
this module to write ``tests/golden/ref_*.npz``; ``tests/test_oracle.py`` asserts the oracle
reproduces those files and the GPU golden test reads them.  The reference checkout exists only
in the build container, so nothing that runs on the GPU box imports this module.

The normal draws are handed out by (loop body, run, t) read from the reference's own frames
(``tf_shim.random_normal``), i.e. *which* resample draws are used is decided by the reference's
``tf.cond`` predicates (cbfssm.py:123-136), not by a schedule restated here.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("CBFSSM_REFERENCE_ROOT", "/root/reference")
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_shim")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "cbfssm", "model", "cbfssm.py"))


def _import_reference():
    """Import tf (the shim) and the reference model modules straight from the checkout."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    mod = sys.modules.get("tensorflow")
    if mod is not None and not getattr(mod, "__version__", "").endswith("-shim"):
        raise RuntimeError("a real tensorflow is already imported; the shim must not shadow it")
    for p in (SHIM_DIR, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    tf = importlib.import_module("tensorflow")
    # the model modules are imported one by one: ``cbfssm/model/__init__.py`` would also pull in
    # PRSSM and the Voliro model, which are off the path
    pkg = importlib.import_module("cbfssm")
    if "cbfssm.model" not in sys.modules:
        import types
        m = types.ModuleType("cbfssm.model")
        m.__path__ = [os.path.join(REFERENCE_ROOT, "cbfssm", "model")]
        sys.modules["cbfssm.model"] = m
        pkg.model = m
    ref_cbfssm = importlib.import_module("cbfssm.model.cbfssm")
    ref_half = importlib.import_module("cbfssm.model.cbfssmhalf")
    for m in (ref_cbfssm, ref_half):
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), m.__file__
    return tf, ref_cbfssm, ref_half


def reference_config(cfg, learning_rate=0.01, batch_size=None, **extra):
    """The dict of run/template.py:19-40 from an ``OracleConfig``-like object."""
    ds = type("DS", (), {"dim_u": cfg.dim_u, "dim_y": cfg.dim_y})
    out = {
        "ds": ds, "batch_size": batch_size or 32, "shuffle": 10000,
        "samples": cfg.samples, "dim_x": cfg.dim_x, "ind_pnt_num": cfg.ind_pnt_num,
        "learning_rate": learning_rate, "loss_factors": np.asarray(cfg.loss_factors, dtype=np.float64),
        "k_factor": float(cfg.k_factor), "recog_len": cfg.recog_len,
        "zeta_pos": cfg.zeta_pos, "zeta_mean": cfg.zeta_mean, "zeta_var": cfg.zeta_var,
        "var_x": np.asarray(cfg.var_x, dtype=np.float64), "var_y": np.asarray(cfg.var_y, dtype=np.float64),
        "gp_var": cfg.gp_var, "gp_len": cfg.gp_len,
    }
    out.update(extra)
    return out


def _np(t):
    return t.detach().numpy().copy()


def run_cbfssm(cfg, param_values, u, y, eps_b, z_b, eps_f, condition=True, learning_rate=0.01):
    """Execute ``CBFSSM(config)`` (cbfssm/model/cbfssm.py) on the minibatch (u, y).

    param_values: the 12 raw tensors in the reference's creation order (gp_f: zeta_pos,
    zeta_mean, zeta_var_unc, kern variance_unc, kern lengthscales_unc; gp_b likewise;
    var_x_unc; var_y_unc).  Draw shapes as in SURVEY Appendix B: eps_b, z_b [2,T,B,S];
    eps_f [T-1,B,S].  Returns every tensor the graph exposes plus d loss / d variables.
    """
    tf, ref_cbfssm, _ = _import_reference()
    B, T, _ = np.asarray(u).shape
    S = cfg.samples
    eps_b, z_b, eps_f = (np.asarray(a, dtype=np.float64) for a in (eps_b, z_b, eps_f))

    def provider(body, run, t, in_branch):
        if body == "_backward_body":
            return (z_b if in_branch else eps_b)[run, t].reshape(B, S, 1)
        assert body == "_forward_body" and not in_branch
        return eps_f[t].reshape(B, S, 1)

    tf.configure(u, y, condition, variable_values=list(param_values), draw_provider=provider)
    model = ref_cbfssm.CBFSSM(reference_config(cfg, learning_rate, batch_size=B))
    st = tf.shim
    assert len(st.variables) == 12, len(st.variables)
    out = {
        "loss": _np(model.loss), "kl_x": _np(model.kl_x), "entropy": _np(model.entropy),
        "kl_z_f": _np(model.gp_f.prior_kl()), "kl_z_b": _np(model.gp_b.prior_kl()),
        "x_final": _np(model.x_final), "y_tilde": _np(model.y_tilde),
        "pred_mean": _np(model.pred_mean), "pred_var": _np(model.pred_var),
        "internal_mean": _np(model.internal_mean), "internal_var": _np(model.internal_var),
        "mse": _np(model.mse), "sde": _np(model.sde),
        "grads": [_np(g) for g in st.gradients],
        "adam": [_np(v) for v in st.adam],
        "var_dict": {k: _np(v) for k, v in model.var_dict.items()},
        "draw_log": list(st.draw_log),
    }
    lf, Sf = np.asarray(cfg.loss_factors, dtype=np.float64), float(S)
    # loglik is a local of _build_loss (cbfssm.py:251); recover it from the exposed terms
    out["loglik"] = (-out["loss"] + lf[0] / Sf * out["kl_x"] - lf[1] / Sf * out["entropy"]
                     + out["kl_z_f"] + out["kl_z_b"]) * Sf / lf[0]
    return out


def run_cbfssmhalf(cfg, param_values, u, y, eps_f, condition=True, recog_model="rnn", learning_rate=0.01):
    """Execute ``CBFSSMHALF(config)`` (cbfssm/model/cbfssmhalf.py).  param_values in creation
    order: the 5 gp_f tensors, var_x_unc, var_y_unc [dy], then -- for recog_model 'rnn' -- GRU
    gates kernel, gates bias, candidate kernel, candidate bias, dense kernel, dense bias."""
    tf, _, ref_half = _import_reference()
    B, T, _ = np.asarray(u).shape
    S = cfg.samples
    eps_f = np.asarray(eps_f, dtype=np.float64)

    def provider(body, run, t, in_branch):
        assert body == "_forward_body" and not in_branch
        return eps_f[t].reshape(B, S, 1)

    tf.configure(u, y, condition, variable_values=list(param_values), draw_provider=provider)
    config = reference_config(cfg, learning_rate, batch_size=B, recog_model=recog_model)
    config["var_y"] = np.asarray(cfg.var_y, dtype=np.float64)[:cfg.dim_y]
    model = ref_half.CBFSSMHALF(config)
    st = tf.shim
    return {
        "loss": _np(model.loss), "kl_x": _np(model.kl_x), "kl_z_f": _np(model.gp_f.prior_kl()),
        "x_final": _np(model.x_final), "pred_mean": _np(model.pred_mean), "pred_var": _np(model.pred_var),
        "internal_mean": _np(model.internal_mean), "internal_var": _np(model.internal_var),
        "grads": [_np(g) for g in st.gradients], "adam": [_np(v) for v in st.adam],
    }
