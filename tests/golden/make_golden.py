#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN SOURCE.

    python tests/golden/make_golden.py [case ...]     # writes tests/golden/ref_<case>.npz

Runs only in the build container (it needs the reference checkout at /root/reference):
``oracle/run_reference.py`` imports the unmodified ``cbfssm/model/{gp_tf,cbfssm,cbfssmhalf}.py``
and executes them under ``oracle/tf_shim`` (eager float64 stand-in for TensorFlow 1.8, which
cannot be installed here).  Inputs are regenerated from seeds by ``tests.helpers``; each file
stores the reference's loss terms, every parameter gradient, the Adam-updated parameters,
the predictive moments and a strided sample of the states.  ``tests/test_oracle.py`` asserts
that the restated oracle reproduces these files (which is what pins it) and the GPU tests
compare the CUDA path with them.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cbfssm_oracle as O          # noqa: E402  (names / parameter order only)
from oracle import cbfssmhalf_oracle as H      # noqa: E402
from oracle import run_reference as RR         # noqa: E402
from tests.helpers import HALF_REF_CASES, NAMED_CASES, STRONG_REF_CASES, half_ref_case, named_case, strong_ref_case   # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def _save(name, arrays, t0):
    path = os.path.join(OUT, "ref_" + name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: loss {float(arrays['loss']):.6f}  reference run {time.time() - t0:.1f}s  "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB", flush=True)


def full_case(name, cfg, params, u, y, eps_b, z_b, eps_f, cond):
    t0 = time.time()
    ref = RR.run_cbfssm(cfg, [params[k].numpy() for k in O.PARAM_NAMES], u, y, eps_b, z_b, eps_f, cond)
    arrays = {f"grad.{k}": g.reshape(params[k].shape) for k, g in zip(O.PARAM_NAMES, ref["grads"])}
    arrays.update({f"adam.{k}": g.reshape(params[k].shape) for k, g in zip(O.PARAM_NAMES, ref["adam"])})
    for k in ("loss", "loglik", "kl_x", "entropy", "kl_z_f", "kl_z_b", "mse"):
        arrays[k] = np.asarray(float(ref[k]))
    for k in ("pred_mean", "pred_var", "internal_mean", "internal_var"):
        arrays[k] = ref[k]
    arrays["x_final_sample"] = ref["x_final"][:, ::15, ::10, :]
    arrays["y_tilde_sample"] = ref["y_tilde"][:, ::15, ::10, :]
    arrays["sde_sample"] = ref["sde"][:, ::15]
    # which (run, t) the reference's own tf.cond predicates resampled at
    arrays["resampled_at"] = np.asarray([(r, t) for body, r, t, br in ref["draw_log"] if br], dtype=np.int64).reshape(-1, 2)
    _save(name, arrays, t0)


def half_case(name):
    cfg, params, w, u, y, eps_f, cond, recog = half_ref_case(name)
    t0 = time.time()
    names = list(H.HALF_PARAM_NAMES) + (list(w) if recog == "rnn" else [])
    vals = [params[k].numpy() for k in H.HALF_PARAM_NAMES] + ([w[k] for k in w] if recog == "rnn" else [])
    ref = RR.run_cbfssmhalf(cfg, vals, u, y, eps_f, cond, recog)
    arrays = {f"grad.{k}": g.reshape(np.shape(v)) for k, g, v in zip(names, ref["grads"], vals)}
    arrays.update({f"adam.{k}": g.reshape(np.shape(v)) for k, g, v in zip(names, ref["adam"], vals)})
    for k in ("loss", "kl_x", "kl_z_f"):
        arrays[k] = np.asarray(float(ref[k]))
    for k in ("pred_mean", "pred_var", "internal_mean", "internal_var", "x_final"):
        arrays[k] = ref[k]
    _save(name, arrays, t0)


def main(argv):
    want = set(argv)
    for name in NAMED_CASES:
        if not want or name in want:
            full_case(name, *named_case(name))
    for name in STRONG_REF_CASES:
        if not want or name in want:
            full_case(name, *strong_ref_case(name))
    for name in HALF_REF_CASES:
        if not want or name in want:
            half_case(name)


if __name__ == "__main__":
    main(sys.argv[1:])
