"""SUMMARY.txt of a directory of bench.py JSON lines (tools/bench_scale8.sh):  python tools/scale8_summary.py DIR"""
import glob
import json
import os
import sys

d = sys.argv[1]
print("# 8-GPU box (B200 x8, NCCL over NVLink), tools/bench_scale8.sh; one bench.py JSON line per file in this directory")
print("%-34s %2s %-7s %9s %9s %10s %8s %18s %8s %13s" % ("run", "N", "scaling", "global B", "part/GPU", "p-steps/s", "ms/step",
                                                          "rollout kernels ms", "rest ms", "e2e p-steps/s"))
for f in sorted(glob.glob(os.path.join(d, "*.json"))):
    try:
        j = json.load(open(f))
    except ValueError:
        continue
    c, sb = j["config"], j.get("step_breakdown", {})
    print("%-34s %2d %-7s %9d %9d %10.4g %8.2f %18.2f %8.2f %13.4g" % (
        os.path.basename(f)[:-5], j["n_gpus"], j["scaling"], c.get("global_batch", 0), c.get("particles_per_gpu", 0), j["value"],
        j["ms_per_step"], sb.get("rollout_and_accumulation_kernels_ms", float("nan")), sb.get("everything_else_ms", float("nan")),
        j["e2e"]["value"]))
