#!/usr/bin/env python
"""ELBO fwd+bwd particle-steps/s of the CBF-SSM hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU restatement of the TF path

Workload (config.workload): BASELINE.json configs[1], "RoboMove-shaped CBF-SSM, state
dim 4, M=20 inducing points": dx/du/dy = 4/2/2, M = 20, S = 50 particles, T = 300,
recog_len 50, k_factor 1, loss_factors (20, 0) (run/run_robomove.py:18-44 with M=20),
synthetic AR(1) sequences and random-init parameters.  A step is one ELBO forward +
backward over one minibatch: both GP prologues, the two backward-message runs, the
forward rollout, the loss, all 12 parameter gradients (one all-reduce when N>1) and the
TF-style Adam update.  One particle-step = one (b, s, t) cell, i.e. B*S*T per step.
Per-GPU batch is fixed (weak scaling); inputs are larger than L2 (see config.l2).

value : device-timed (CUDA events), inputs and draws resident in HBM.
e2e   : the same step through the public API ``CBFSSM.evaluate_batch`` with pinned HOST
        minibatch buffers: H2D of u,y, in-library Philox draws, D2H of the loss, every step.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "elbo_fwd_bwd_particle_steps_per_sec"
UNIT = "particle-steps/s"
WORKLOADS = {
    # BASELINE.json configs[1]: the headline workload (default).  Batches are sized to whole waves of the
    # kernels' resident CTAs (148 SMs x 3 CTAs x 128 particles for the register path, 148 x 2 x 128 for the
    # tensor path): 4544 x 50 particles = 4.0 waves, 1515 x 50 = 2.0 waves.
    "robomove_m20": dict(dx=4, du=2, dy=2, M=20, S=50, T=300, R=50, kap=1.0, lf=(20.0, 0.0), batch=4544,
                         name="RoboMove-shaped CBF-SSM dx4/du2/dy2 M20 S50 T300 R50 (BASELINE.json configs[1])"),
    # run/template.py defaults (north_star target shape M=100, D=4); secondary, not the driver's line
    "template_m100": dict(dx=4, du=2, dy=2, M=100, S=50, T=100, R=50, kap=1.0, lf=(10.0, 0.0), batch=1515,
                          name="run/template.py CBF-SSM dx4/du2/dy2 M100 S50 T100 R50"),
    "sarcos_m100": dict(dx=14, du=7, dy=7, M=100, S=20, T=250, R=16, kap=50.0, lf=(6.0, 0.0), batch=1894,
                        name="Sarcos-shaped CBF-SSM dx14/du7/dy7 M100 S20 T250 R16 (BASELINE.json configs[2])"),
}
# BASELINE.json configs[4] (scaling sweep) corner D=8, M=100, T=500: 2.0 waves of two-tile CTAs; its reverse pass
# needs ~200 GB of operand tiles and therefore runs in time windows
WORKLOADS["sweep_d8_m100"] = dict(dx=8, du=1, dy=4, M=100, S=1024, T=500, R=16, kap=1.0, lf=(10.0, 0.0), batch=74,
                                  name="sweep corner CBF-SSM dx8/du1/dy4 M100 S1024 T500 R16 (BASELINE.json configs[4])")
# BASELINE.json configs[3]: Voliro-shaped multi-experiment windows, 256 sequences x 1024 particles per GPU
WORKLOADS["voliro_m20"] = dict(dx=13, du=6, dy=7, M=20, S=1024, T=64, R=16, kap=1.0, lf=(20.0, 0.0), batch=256,
                               name="Voliro-shaped CBF-SSM dx13/du6/dy7 M20 S1024 T64 R16 (BASELINE.json configs[3])")
# BASELINE.json configs[0]: SpringNonLinear with the run/template.py defaults at the reference's own batch size
WORKLOADS["spring_template_b32"] = dict(dx=4, du=1, dy=1, M=100, S=50, T=100, R=50, kap=1.0, lf=(10.0, 0.0), batch=32,
                                        name="SpringNonLinear run/template.py defaults dx4/du1/dy1 M100 S50 T100 R50, B=32 "
                                             "(BASELINE.json configs[0])")
WORK = dict(WORKLOADS["robomove_m20"])
CFG_INIT = dict(zeta_pos=2.0, zeta_mean=0.01, zeta_var=1e-4, gp_var=0.01, gp_len=1.0)


def flop_model(M, Din, Dout):
    """algorithmic FLOPs of one sparse-GP evaluation, forward / reverse."""
    f_fwd = 2 * M * M + M * (3 * Din + 4 * Dout + 5)
    f_bwd = 4 * M * M + M * (6 * Din + 8 * Dout + 10)
    return f_fwd, f_bwd


def ar1(rng, shape, rho=0.95):
    e = rng.standard_normal(shape).astype(np.float32)
    out = np.empty(shape, dtype=np.float32)
    out[:, 0] = e[:, 0]
    c = np.float32(np.sqrt(1 - rho * rho))
    for t in range(1, shape[1]):
        out[:, t] = rho * out[:, t - 1] + c * e[:, t]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except ValueError:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


# --------------------------------------------------------------------------------------
# CPU arm: the oracle (float64 PyTorch-CPU restatement of the TF graph), all host threads
# --------------------------------------------------------------------------------------
def cpu_oracle_rate(steps, warmup, B=32):
    import torch
    from oracle import cbfssm_oracle as O
    w = WORK
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.OracleConfig(dim_x=w["dx"], dim_u=w["du"], dim_y=w["dy"], ind_pnt_num=w["M"], samples=w["S"],
                         recog_len=w["R"], k_factor=w["kap"], loss_factors=w["lf"], **CFG_INIT)
    params = O.init_params(cfg, 1)
    rng = np.random.default_rng(1)
    u, y = ar1(rng, (B, w["T"], w["du"])), ar1(rng, (B, w["T"], w["dy"]))
    eb, zb, ef = O.draw_noise(B, w["S"], w["T"], 1)
    for _ in range(warmup):
        O.loss_and_grads(cfg, params, u, y, eb, zb, ef, True)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.loss_and_grads(cfg, params, u, y, eb, zb, ef, True)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    psteps = B * w["S"] * w["T"]
    return psteps / dt, dt, cores, f"one minibatch B={B} S={w['S']} T={w['T']} M={w['M']} per step (reference batch size), float64"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, dt, cores, sample = cpu_oracle_rate(args.steps, min(args.warmup, 1))
    w = WORK
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"],
                       "batch_per_step": 32, **{k: w[k] for k in ("M", "S", "T", "R")}},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "CPU restatement of the TF-1.8 path (oracle/), not TensorFlow: TF 1.8 cannot be installed here"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from cbf_ssm_b200 import _lib
    from cbf_ssm_b200.model import CBFSSM

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    lib = _lib.load()

    w = WORK
    B, S, T = args.batch, w["S"], w["T"]

    class DS:
        dim_u, dim_y = w["du"], w["dy"]

    config = {"ds": DS, "batch_size": B, "shuffle": 1, "dim_x": w["dx"], "ind_pnt_num": w["M"], "samples": S,
              "learning_rate": 0.01, "loss_factors": np.asarray(w["lf"]), "k_factor": w["kap"], "recog_len": w["R"],
              "var_x": np.asarray([0.1 ** 2] * w["dx"]), "var_y": np.asarray([1.0] * w["dx"]), **CFG_INIT}
    # each rank owns B whole sequences (weak scaling); gradients/terms are all-reduced
    model = CBFSSM(config, device=dev, group=None, seed=1)
    eng = model.engine
    eng.flags = args.flags
    eng.group = group
    if world > 1:
        dist.broadcast(eng.theta, src=0)

    rng = np.random.default_rng(100 + rank)
    u_host = torch.from_numpy(ar1(rng, (B, T, w["du"]))).pin_memory()
    y_host = torch.from_numpy(ar1(rng, (B, T, w["dy"]))).pin_memory()
    u_dev, y_dev = u_host.to(dev), y_host.to(dev)
    N = B * S
    eb = torch.empty(2, T, N, dtype=torch.float32, device=dev)
    zb = torch.empty(2, T, N, dtype=torch.float32, device=dev)
    ef = torch.empty(T - 1, N, dtype=torch.float32, device=dev)
    for i, t in enumerate((eb, zb, ef)):
        eng.fill_normal(t, 1234 + rank, i)
    psteps_local = N * T
    psteps_global = psteps_local * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        eng.forward(u_dev, y_dev, eb, zb, ef, True)
        eng.backward()
        eng.adam_step(config["learning_rate"])

    def e2e_step():
        return model.evaluate_batch(u_host, y_host, ["train", "loss"], True)[1]

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident timing ----------------
    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    eng.launches = 0
    lib.cbf_timing_enable(1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        device_step()
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launches
    clocks = sampler.stop() if rank == 0 else None
    kms = (ctypes.c_double * 8)()
    kcnt = (ctypes.c_int64 * 8)()
    _lib.check(lib.cbf_timing_read(kms, kcnt))
    lib.cbf_timing_enable(0)
    ms_per_step = ms_total / args.steps
    value = psteps_global / (ms_per_step * 1e-3)

    # ---------------- end-to-end through the public API, host buffers ----------------
    for _ in range(3):
        e2e_step()
    barrier()
    ev0.record()
    loss = None
    for _ in range(args.steps):
        loss = e2e_step()
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    e2e_value = psteps_global / (e2e_ms * 1e-3)
    h2d = u_host.numel() * 4 + y_host.numel() * 4
    d2h = 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    peaks, how = measured_peaks()
    Din, dx, dh, M = w["dx"] + w["du"], w["dx"], w["dx"] - w["dy"], w["M"]
    ff, fb = flop_model(M, Din, dx)
    bf, bb = flop_model(M, Din, dh)
    # algorithmic FLOPs per particle-step attributed to each kernel (1 gp_f + 2 gp_b evaluations)
    kflops = [2 * bf, ff, fb, 2 * bb]
    knames = ["bm_forward", "fw_forward", "fw_reverse", "bm_reverse"]
    # kernel time per step (a kernel may be launched several times per step: chain batches, time windows)
    kavg = [kms[i] / args.steps for i in range(4)]
    dom = int(np.argmax([kms[i] for i in range(4)]))
    simt_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12     # TFLOP/s
    achieved = kflops[dom] * psteps_local / (kavg[dom] * 1e-3) / 1e12
    all_flops = sum(kflops) * psteps_local
    step_frac = all_flops / (ms_per_step * 1e-3) / 1e12 / simt_peak
    bytes_pstep = 2 * (4 * dx + 8 * dh + 12)
    traffic = None            # DRAM bytes per launch of the dominant kernel, from the committed ncu capture
    try:
        tr = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")))
        ent = tr.get(args.workload, {})
        if ent.get("batch") == B:
            traffic = ent.get(knames[dom])
    except (OSError, ValueError):
        pass
    roofline = {"bound": "fp32_simt", "kernel": knames[dom], "achieved": achieved, "peak": simt_peak,
                "unit": "TFLOP/s", "frac": achieved / simt_peak, "traffic": traffic,
                "peak_source": f"148 SM x 128 lanes x 2 x sm_max_mhz ({how} MEASURED_PEAKS.json); the path is FP32-SIMT "
                               "compute-bound (SURVEY 8d), not HBM- or tensor-bound",
                "flops_per_particle_step": {k: v for k, v in zip(knames, kflops)},
                "kernel_ms_avg": {**{k: v for k, v in zip(knames, kavg)},
                                  **({"outer_f": kms[4] / args.steps, "outer_b": kms[5] / args.steps} if kcnt[4] else {})},
                "kernel_share_of_step": {k: kavg[i] / ms_per_step for i, k in enumerate(knames)},
                "whole_step_frac": step_frac,
                **({"note": "tensor path: the M^2 contractions run on tcgen05 (fp16/bf16 splits, 3 MMA passes), so "
                            "algorithmic FLOPs against the FP32-SIMT peak can exceed 1; the kernels are bound by the "
                            "O(M*D) SIMT work at 8 warps/SM (profiles/r01i_m100_*)"} if kcnt[4] else {}),
                "hbm": {"algorithmic_bytes_per_particle_step": bytes_pstep,
                        "achieved_gbs": bytes_pstep * psteps_local / (ms_per_step * 1e-3) / 1e9,
                        "peak_gbs": peaks.get("hbm_gbs")}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, dt, cores, sample = cpu_oracle_rate(3, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + "; 1 warm-up + 3 timed",
               "ms_per_step": dt * 1e3}

    resident_bytes = sum(t.numel() * 4 for t in (eb, zb, ef, u_dev, y_dev))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"],
                       "batch_per_gpu": B, "global_batch": B * world, "particles_per_gpu": N, "seq_len": T,
                       "M": M, "S": S, "R": w["R"], "parallelism": f"dp{world} over sequences, 1 all-reduce/step",
                       "l2": f"inputs larger than L2: {resident_bytes / 2**20:.0f} MiB of draws+data and "
                             f"{(T * (dx + 3 * dh) * N * 4) / 2**20:.0f} MiB of states streamed per step"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "loss": float(loss)},
            "gpu_launches": launches, "roofline": roofline}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="sequences per GPU per step (0 = workload default)")
    ap.add_argument("--workload", default="robomove_m20", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--M", type=int, default=0, help="override the workload's number of inducing points (kernel-path studies)")
    ap.add_argument("--S", type=int, default=0, help="override the particles per sequence")
    ap.add_argument("--T", type=int, default=0, help="override the sequence length")
    ap.add_argument("--R", type=int, default=0, help="override recog_len")
    ap.add_argument("--flags", type=int, default=0, help="cbf_shape.flags (1 cooperative kernels, 2 no tensor cores)")
    args = ap.parse_args()
    WORK.clear()
    WORK.update(WORKLOADS[args.workload])
    for key in ("M", "S", "T", "R"):
        v = getattr(args, key)
        if v > 0:
            WORK[key] = v
            WORK["name"] += " [%s overridden to %d]" % (key, v)
    if args.batch <= 0:
        args.batch = WORK["batch"]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
