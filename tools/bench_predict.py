#!/usr/bin/env python
"""Latency of the prediction path the reference's Outputs uses (cbfssm/outputs/outputs.py:121-141: the whole test
experiment as one sequence, B = 1, S = config['samples'], condition = False) through the public API, with and
without the dead backward-message work.  Prints one JSON line per case (kept under profiles/)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from cbf_ssm_b200.model import CBFSSM


def main():
    dev = torch.device("cuda", 0)
    for name, M, T in (("predict_b1_t5000_m100", 100, 5000), ("predict_b1_t5000_m20", 20, 5000)):
        class DS:
            dim_u, dim_y = 2, 2
        cfg = {"ds": DS, "batch_size": 1, "shuffle": 1, "dim_x": 4, "ind_pnt_num": M, "samples": 50, "learning_rate": 0.01,
               "loss_factors": np.asarray([10.0, 0.0]), "k_factor": 1.0, "recog_len": 50, "zeta_pos": 2.0, "zeta_mean": 0.01,
               "zeta_var": 1e-4, "var_x": np.full(4, 0.01), "var_y": np.full(4, 1.0), "gp_var": 0.01, "gp_len": 1.0}
        model = CBFSSM(cfg, device=dev, seed=1)
        g = np.random.default_rng(0)
        u = torch.from_numpy(g.standard_normal((1, T, 2)).astype(np.float32)).pin_memory()
        y = torch.from_numpy(g.standard_normal((1, T, 2)).astype(np.float32)).pin_memory()
        res = {}
        for label, names in (("prediction_handles_only", ["pred_mean", "pred_var"]),
                             ("with_loss_handle_full_message", ["pred_mean", "pred_var", "loss"])):
            for _ in range(3):
                model.evaluate_batch(u, y, names, False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 10
            for _ in range(n):
                model.evaluate_batch(u, y, names, False)      # returns host arrays: includes D2H
            dt = (time.perf_counter() - t0) / n
            res[label] = {"ms_per_call": dt * 1e3, "particle_steps_per_s": 50 * T / dt}
        print(json.dumps({"case": name, "B": 1, "S": 50, "T": T, "M": M, "recog_len": 50, **res}), flush=True)


if __name__ == "__main__":
    main()
