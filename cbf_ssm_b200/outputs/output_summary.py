"""Mean / spread of the test RMSE over repeated runs (cbfssm/outputs/output_summary.py:8-31): the file
``summary.txt`` the reference's multi-run scripts write (run/run_sarcos.py:53-79, run/run_smallscale.py:66-92),
same text layout, plus a copy of the launching script."""
from __future__ import annotations

import os
import shutil
import sys

import numpy as np


class OutputSummary:

    def __init__(self, out_dir, copy_main=True):
        self.out_dir = out_dir
        self.rmse_all = []
        os.makedirs(out_dir, exist_ok=True)
        main = os.path.abspath(sys.argv[0]) if sys.argv and sys.argv[0] else ""
        if copy_main and os.path.isfile(main):                   # output_summary.py:15
            shutil.copyfile(main, os.path.join(out_dir, "main.py"))

    def add_outputs(self, outputs):
        """Record ``outputs.get_last_rmse()`` of one finished run (None if that run wrote no test RMSE)."""
        self.rmse_all.append(outputs.get_last_rmse())

    def write_summary(self):
        runs = list(self.rmse_all)
        if not runs or runs[0] is None:                          # output_summary.py:21,30-31
            print("RMSE summary skipped")
            return None
        vals = np.asarray(runs, dtype=np.float64)
        lines = ["RMSE", "====", "", "Runs:"] + ["  %f" % v for v in vals]
        lines += ["Mean: %f" % np.mean(vals), "Std:  %f" % np.std(vals)]      # population std, like np.std
        path = os.path.join(self.out_dir, "summary.txt")
        with open(path, "w") as f:
            f.write("\n".join(lines) + "\n")
        return path
