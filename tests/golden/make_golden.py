#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the float64 oracle.

    python tests/golden/make_golden.py            # writes tests/golden/<case>.npz

PARITY UNPINNED: the reference ships no golden vectors and TensorFlow 1.8 cannot run
here, so these files pin the *oracle* (oracle/cbfssm_oracle.py), not TensorFlow output.
Inputs are regenerated from seeds by tests.helpers.named_case(); each file stores the
oracle's loss terms, all 12 parameter gradients, the predictive moments and a strided
sample of x_final.  The GPU parity tests compare the CUDA path with these files at the
named configurations without re-running the oracle on the GPU box.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cbfssm_oracle as O          # noqa: E402
from tests.helpers import NAMED_CASES, named_case   # noqa: E402


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name in NAMED_CASES:
        cfg, params, u, y, eps_b, z_b, eps_f, cond = named_case(name)
        t0 = time.time()
        res, gd = O.loss_and_grads(cfg, params, u, y, eps_b, z_b, eps_f, cond)
        dt = time.time() - t0
        arrays = {f"grad.{k}": v.numpy() for k, v in gd.items()}
        for k in ("loss", "loglik", "kl_x", "entropy", "kl_z_f", "kl_z_b"):
            arrays[k] = np.asarray(float(getattr(res, k).detach()))
        arrays["pred_mean"] = res.pred_mean.detach().numpy()
        arrays["pred_var"] = res.pred_var.detach().numpy()
        arrays["internal_mean"] = res.internal_mean.detach().numpy()
        arrays["internal_var"] = res.internal_var.detach().numpy()
        xf = res.x_final.detach().numpy()
        arrays["x_final_sample"] = xf[:, ::15, ::10, :]
        arrays["y_tilde_sample"] = res.y_tilde.detach().numpy()[:, ::15, ::10, :]
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: loss {float(res.loss):.6f}  oracle {dt:.1f}s  -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
