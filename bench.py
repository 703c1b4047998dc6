#!/usr/bin/env python
"""ELBO fwd+bwd particle-steps/s of the CBF-SSM hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework
    python bench.py --impl reference [--steps K] [--warmup W]      # the CPU arm (float64 oracle) on the same config

Headline workload (config.workload): the shape BASELINE.json's target is quoted on -- run/template.py's
CBF-SSM at M = 100 inducing points, state dim 4 (dx/du/dy = 4/2/2), S = 50 particles, T = 100, recog_len 50,
loss_factors (10, 0) -- on synthetic AR(1) sequences with random-init parameters.  The same JSON line carries,
under ``extra.robomove_m20``, BASELINE.json configs[1] (RoboMove-shaped, M = 20, T = 300) measured the same way.

A step is one ELBO forward + backward over one minibatch: the step's normal draws (in-library Philox), both
float64 GP prologues, the two backward-message runs, the forward rollout, the loss, all 12 parameter gradients
(one all-reduce when N > 1), the prologue adjoints and the TF-style Adam update.  One particle-step = one
(b, s, t) cell.  The model is built with the real ``group=`` path: every rank holds the global minibatch's
(tiny) u, y and rolls out its contiguous share of the particles n = b*S + s.

value : device-timed (CUDA events), u / y resident in HBM.
e2e   : the same step through the public API ``CBFSSM.evaluate_batch`` with pinned HOST minibatch buffers:
        H2D of u, y and D2H of the loss every step.
--scaling weak (default): the global batch grows with N (fixed work per GPU); strong: the global batch is fixed.
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
    # run/template.py defaults = the shape BASELINE.json's north-star target is quoted on (M=100, D=4): the headline.
    # Batches are sized to whole waves of the kernels' resident CTAs (148 SMs x 2 CTAs x 128 particles for the
    # tensor path, 148 x 3 x 128 for the register path): 1515 x 50 particles = 2.0 waves, 4544 x 50 = 4.0 waves.
    "template_m100": dict(dx=4, du=2, dy=2, M=100, S=50, T=100, R=50, kap=1.0, lf=(10.0, 0.0), batch=1515, cpu_batch=128,
                          name="run/template.py CBF-SSM dx4/du2/dy2 M100 S50 T100 R50 (north-star target shape M=100, D=4)"),
    # BASELINE.json configs[1]
    "robomove_m20": dict(dx=4, du=2, dy=2, M=20, S=50, T=300, R=50, kap=1.0, lf=(20.0, 0.0), batch=4544, cpu_batch=256,
                         name="RoboMove-shaped CBF-SSM dx4/du2/dy2 M20 S50 T300 R50 (BASELINE.json configs[1])"),
    # BASELINE.json configs[2]
    "sarcos_m100": dict(dx=14, du=7, dy=7, M=100, S=20, T=250, R=16, kap=50.0, lf=(6.0, 0.0), batch=1894, cpu_batch=64,
                        name="Sarcos-shaped CBF-SSM dx14/du7/dy7 M100 S20 T250 R16 (BASELINE.json configs[2])"),
    # BASELINE.json configs[4] (scaling sweep) corner D=8, M=100, T=500: 2.0 waves of two-tile CTAs
    "sweep_d8_m100": dict(dx=8, du=1, dy=4, M=100, S=1024, T=500, R=16, kap=1.0, lf=(10.0, 0.0), batch=74, cpu_batch=1,
                          name="sweep corner CBF-SSM dx8/du1/dy4 M100 S1024 T500 R16 (BASELINE.json configs[4])"),
    # BASELINE.json configs[4]: the 2^20-particle corner at M=100, D=4 (strong-scaling subject: --scaling strong)
    "sweep_1m_m100": dict(dx=4, du=2, dy=2, M=100, S=1024, T=500, R=16, kap=1.0, lf=(10.0, 0.0), batch=1024, cpu_batch=1,
                          name="sweep corner CBF-SSM dx4/du2/dy2 M100 S1024 B1024 (2^20 particles) T500 R16 (BASELINE.json configs[4])"),
    # BASELINE.json configs[4] corner M=500, D=8: the float64 batched path (P = 2 MB no longer fits an SM)
    "sweep_d8_m500": dict(dx=8, du=1, dy=4, M=500, S=1024, T=500, R=16, kap=1.0, lf=(10.0, 0.0), batch=16, cpu_batch=1,
                          name="sweep corner CBF-SSM dx8/du1/dy4 M500 S1024 T500 R16 (BASELINE.json configs[4]), float64 batched path"),
    # BASELINE.json configs[3]: Voliro-shaped multi-experiment windows, 256 sequences x 1024 particles per GPU
    "voliro_m20": dict(dx=13, du=6, dy=7, M=20, S=1024, T=64, R=16, kap=1.0, lf=(20.0, 0.0), batch=256, cpu_batch=2,
                       name="Voliro-shaped CBF-SSM dx13/du6/dy7 M20 S1024 T64 R16 (BASELINE.json configs[3])"),
    # BASELINE.json configs[0]: SpringNonLinear with the run/template.py defaults at the reference's own batch size
    "spring_template_b32": dict(dx=4, du=1, dy=1, M=100, S=50, T=100, R=50, kap=1.0, lf=(10.0, 0.0), batch=32, cpu_batch=32,
                                name="SpringNonLinear run/template.py defaults dx4/du1/dy1 M100 S50 T100 R50, B=32 "
                                     "(BASELINE.json configs[0])"),
}
HEADLINE, EXTRA = "template_m100", "robomove_m20"
CFG_INIT = dict(zeta_pos=2.0, zeta_mean=0.01, zeta_var=1e-4, gp_var=0.01, gp_len=1.0)
TENSOR_PASSES = 3          # fp16 / bf16 two-term splits: hi*hi + lo*hi + hi*lo per product (kernels_tc.cuh)


def flop_model(M, Din, Dout):
    """algorithmic FLOPs of one sparse-GP evaluation, forward / reverse, split into the
    M^2 contractions (tensor-core candidates) and the O(M*D) remainder."""
    return dict(fwd_m2=2 * M * M, fwd_md=M * (3 * Din + 4 * Dout + 5),
                bwd_m2=4 * M * M, bwd_md=M * (6 * Din + 8 * Dout + 10))


def live_message_steps(T, R):
    """Backward-message evaluations that feed something, of 2T (SURVEY 8a note 5): chains are cut at each
    resample (cbfssm.py:123-136); steps after a chain's last written step are dead and skipped."""
    live = 0
    for run in (0, 1):
        off = 1 if run == 0 else R + 1
        starts = [T - 1] + [t for t in range(T - 2, -1, -1) if (t + off) % (2 * R) == 0]
        for i, t_hi in enumerate(starts):
            t_next = starts[i + 1] if i + 1 < len(starts) else -1
            written = [t for t in range(t_hi, t_next, -1) if ((t % (2 * R)) < R) == (run == 0)]
            if written:
                live += t_hi - min(written) + 1
    return live


def ar1(rng, shape, rho=0.95):
    e = rng.standard_normal(shape).astype(np.float32)
    out = np.empty(shape, dtype=np.float32)
    out[:, 0] = e[:, 0]
    c = np.float32(np.sqrt(1 - rho * rho))
    for t in range(1, shape[1]):
        out[:, t] = rho * out[:, t - 1] + c * e[:, t]
    return out


def config_dict(w, batch_per_gpu, world, scaling):
    """The workload description both arms print (identical by construction)."""
    gb = batch_per_gpu * world if scaling == "weak" else batch_per_gpu
    return {"workload": w["name"], "dx": w["dx"], "du": w["du"], "dy": w["dy"], "M": w["M"], "S": w["S"],
            "seq_len": w["T"], "R": w["R"], "k_factor": w["kap"], "loss_factors": list(w["lf"]),
            "global_batch": gb, "particles_per_gpu": -(-gb * w["S"] // world),
            "particle_steps_per_step": gb * w["S"] * w["T"],
            "parallelism": f"dp{world}: particles n=b*S+s split contiguously over ranks, 1 all-reduce/step",
            "l2": "working set (states, messages, adjoints, operand tiles, draws re-generated every step) far larger "
                  "than the 126 MB L2: see l2_working_set_mib"}


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
            return json.load(f), "MEASURED_PEAKS.json"
    except OSError:
        # B200_PROFILING.md fallback figures
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1590.0, "sm_max_mhz": 1965.0}, \
            "B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)"


# --------------------------------------------------------------------------------------
# CPU arm: the oracle (float64 PyTorch-CPU restatement of the TF graph, pinned against the reference's own
# source -- tests/golden/ref_*.npz)
# --------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, w, B):
        import torch
        from oracle import cbfssm_oracle as O
        self.torch, self.O, self.w, self.B = torch, O, w, B
        self.cfg = O.OracleConfig(dim_x=w["dx"], dim_u=w["du"], dim_y=w["dy"], ind_pnt_num=w["M"], samples=w["S"],
                                  recog_len=w["R"], k_factor=w["kap"], loss_factors=w["lf"], **CFG_INIT)
        self.params = O.init_params(self.cfg, 1)
        rng = np.random.default_rng(1)
        self.u, self.y = ar1(rng, (B, w["T"], w["du"])), ar1(rng, (B, w["T"], w["dy"]))
        self.draws = O.draw_noise(B, w["S"], w["T"], 1)
        self.psteps = B * w["S"] * w["T"]

    def step(self):
        t0 = time.perf_counter()
        self.O.loss_and_grads(self.cfg, self.params, self.u, self.y, *self.draws, True)
        return time.perf_counter() - t0

    def run(self, steps, warmup, threads):
        self.torch.set_num_threads(threads)
        for _ in range(warmup):
            self.step()
        return [self.step() for _ in range(steps)]

    def sample(self):
        w = self.w
        return (f"each step = one ELBO fwd+bwd over B={self.B} sequences of the workload (S={w['S']}, T={w['T']}, "
                f"M={w['M']}; {self.psteps} particle-steps), float64 PyTorch-CPU oracle")


def cpu_baseline(w, steps, warmup, with_five=True):
    """All host cores, median of ``steps`` after ``warmup``; plus the reference's own thread setting
    (intra-op 5, cbfssm/training/trainer.py:22-26).  The sample batch is past the point where the rate
    saturates (printed sweep in profiles/)."""
    cores = os.cpu_count() or 1
    arm = CpuArm(w, w["cpu_batch"])
    ts = arm.run(steps, warmup, cores)
    med = float(np.median(ts))
    out = {"value": arm.psteps / med, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": med * 1e3,
           "sample": arm.sample() + f"; {warmup} warm-up + {steps} timed, median",
           "note": "CPU restatement of the TF-1.8 path (oracle/), not TensorFlow: TF 1.8 cannot be installed here"}
    if with_five:
        t5 = arm.run(3, 1, min(5, cores))
        out["five_threads"] = {"value": arm.psteps / float(np.median(t5)), "threads": min(5, cores),
                               "why": "intra_op_parallelism_threads = 5 (cbfssm/training/trainer.py:24); 1 warm-up + 3 timed"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    world = args.gpus
    cores = os.cpu_count() or 1
    arm = CpuArm(w, w["cpu_batch"])
    ts = arm.run(args.steps, args.warmup, cores)
    dt = float(np.mean(ts))
    rate = arm.psteps / dt
    t5 = arm.run(2, 0, min(5, cores))
    sweep = {}
    for b in sorted({max(1, w["cpu_batch"] // 4), w["cpu_batch"]}):
        if b == w["cpu_batch"]:
            sweep[str(b)] = arm.psteps / float(np.median(ts))
        else:
            a2 = CpuArm(w, b)
            sweep[str(b)] = a2.psteps / float(np.median(a2.run(2, 1, cores)))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(w, args.batch, world, args.scaling),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": arm.sample(),
                             "five_threads": {"value": arm.psteps / float(np.median(t5)), "threads": min(5, cores),
                                              "why": "intra_op_parallelism_threads = 5 (cbfssm/training/trainer.py:24)"},
                             "rate_vs_sample_batch": sweep,
                             "note": "CPU restatement of the TF-1.8 path (oracle/), not TensorFlow: TF 1.8 cannot be "
                                     "installed here; the rate is per particle-step and saturates in the sample batch"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def measure_workload(name, args, steps, ctx):
    """Time one workload on this rank set; returns (per-rank partial dict).  ctx: torch, dist, world, rank, dev, lib."""
    import torch
    from cbf_ssm_b200 import _lib
    from cbf_ssm_b200.model import CBFSSM
    dist, world, rank, dev, lib, group = ctx["dist"], ctx["world"], ctx["rank"], ctx["dev"], ctx["lib"], ctx["group"]
    w = dict(WORKLOADS[name])
    for key in ("M", "S", "T", "R"):
        v = getattr(args, key)
        if v > 0 and name == args.workload:
            w[key] = v
            w["name"] += " [%s overridden to %d]" % (key, v)
    bpg = args.batch if (args.batch > 0 and name == args.workload) else w["batch"]
    Bg = bpg * world if args.scaling == "weak" else bpg          # global minibatch
    S, T = w["S"], w["T"]

    class DS:
        dim_u, dim_y = w["du"], w["dy"]

    config = {"ds": DS, "batch_size": Bg, "shuffle": 1, "dim_x": w["dx"], "ind_pnt_num": w["M"], "samples": S,
              "learning_rate": 0.01, "loss_factors": np.asarray(w["lf"]), "k_factor": w["kap"], "recog_len": w["R"],
              "var_x": np.asarray([0.1 ** 2] * w["dx"]), "var_y": np.asarray([1.0] * w["dx"]), **CFG_INIT}
    model = CBFSSM(config, device=dev, group=group, seed=1)       # the product's own sharding / all-reduce path
    eng = model.engine
    eng.flags = args.flags

    rng = np.random.default_rng(100)                              # same global minibatch on every rank
    u_host = torch.from_numpy(ar1(rng, (Bg, T, w["du"]))).pin_memory()
    y_host = torch.from_numpy(ar1(rng, (Bg, T, w["dy"]))).pin_memory()
    u_dev, y_dev = u_host.to(dev), y_host.to(dev)
    n0, nl = model._shard(Bg)
    psteps_global = Bg * S * T
    lr = config["learning_rate"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        eb, zb, ef = model._draws(Bg, T, n0, nl)                  # in-library Philox, this rank's particles
        eng.forward(u_dev, y_dev, eb, zb, ef, True, n_offset=n0, n_local=nl)
        eng.backward()                                            # all-reduce inside when world > 1
        eng.adam_step(lr)

    def e2e_step():
        return model.evaluate_batch(u_host, y_host, ["train", "loss"], True)[1]

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    warm = max(args.warmup, 3)
    for _ in range(warm):
        device_step()
    barrier()
    eng.launches = 0
    lib.cbf_timing_enable(1)
    sampler = ClockSampler(ctx["local"])
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
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
    ms_per_step = ms_total / steps

    for _ in range(3):
        e2e_step()
    barrier()
    ev0.record()
    loss = None
    for _ in range(steps):
        loss = e2e_step()
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1)) / steps
    out = dict(w=w, bpg=bpg, Bg=Bg, nl=nl, ms_per_step=ms_per_step, value=psteps_global / (ms_per_step * 1e-3),
               e2e_ms=e2e_ms, e2e_value=psteps_global / (e2e_ms * 1e-3), loss=float(loss), launches=launches,
               clocks=clocks, kms=[kms[i] / steps for i in range(8)], kcnt=[int(kcnt[i]) for i in range(8)],
               h2d=u_host.numel() * 4 + y_host.numel() * 4, steps=steps, warm=warm,
               tensor_path=bool(kcnt[4] or kcnt[5]), ws_bytes=int(eng._ws.numel()) if eng._ws is not None else 0,
               f64_path=bool(eng.kernel_path == 3 or (int(args.flags) & 128)))
    if out["f64_path"]:      # measured FP64 GEMM roof of this GPU (cuBLAS DGEMM through torch, best of 5)
        a64 = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        best = 0.0
        for _ in range(6):
            ev0.record()
            torch.matmul(a64, a64)
            ev1.record()
            torch.cuda.synchronize()
            best = max(best, 2 * 4096 ** 3 / (ev0.elapsed_time(ev1) * 1e-3) / 1e12)
        out["dgemm_peak"] = best
        del a64
    del model, eng
    torch.cuda.empty_cache()
    return out


def roofline_of(m, peaks, peaks_src, fp32_peak, traffic_table, workload_key):
    """Roofline of the dominant kernel (by measured time) of one measured workload.

    Register path (M <= 32): every FLOP is FP32 SIMT -> algorithmic FLOPs of the kernel / its CUDA-event time
    against the FP32 FMA rate measured on this GPU a moment ago (cbf_measure_fp32_peak).
    Tensor path: the kernel's M^2 contractions run on tcgen05 as 3-pass two-term splits (fp32-equivalent
    products), the O(M*D) work on the SIMT pipes, so the line carries both fractions: M^2 FLOPs against the measured
    sustained bf16 rate / 3 passes, and O(M*D) FLOPs against the measured FP32 rate.  Only LIVE backward-message
    evaluations are counted (SURVEY 8a note 5)."""
    w, nl = m["w"], m["nl"]
    Din, dx, dh, M, T, R = w["dx"] + w["du"], w["dx"], w["dx"] - w["dy"], w["M"], w["T"], w["R"]
    ff, fb = flop_model(M, Din, dx), flop_model(M, Din, dh)
    live = live_message_steps(T, R)
    ev_f, ev_b = nl * (T - 1), nl * live                          # GP evaluations per launch group, this rank
    kern = {   # name: (evaluations, M^2 flops / evaluation, O(MD) flops / evaluation, timed kinds)
        "bm_forward": (ev_b, fb["fwd_m2"], fb["fwd_md"], (0,)),
        "fw_forward": (ev_f, ff["fwd_m2"], ff["fwd_md"], (1,)),
        "fw_reverse": (ev_f, ff["bwd_m2"], ff["bwd_md"], (2, 4)),   # on the tensor path the accumulation GEMM
        "bm_reverse": (ev_b, fb["bwd_m2"], fb["bwd_md"], (3, 5)),   # (kinds 4 / 5) is part of the reverse pass
    }
    times = {k: sum(m["kms"][i] for i in v[3]) for k, v in kern.items()}
    dom = max(times, key=times.get)
    ev, m2, md, _ = kern[dom]
    t_s = times[dom] * 1e-3
    tens_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))) / TENSOR_PASSES
    bytes_pstep = 2 * (4 * dx + 8 * dh + 12)
    traffic = None
    ent = traffic_table.get(workload_key, {})
    if ent.get("batch") == m["bpg"]:
        traffic = ent.get(dom)
    share = {k: times[k] / m["ms_per_step"] for k in kern}
    all_m2 = sum(v[0] * v[1] for v in kern.values())
    all_md = sum(v[0] * v[2] for v in kern.values())
    common = {"kernel": dom, "kernel_ms_per_step": times, "kernel_share_of_step": share,
              "live_message_evaluations_of_2T": [live, 2 * T], "traffic": traffic,
              "hbm": {"algorithmic_bytes_per_particle_step": bytes_pstep,
                      "algorithmic_bytes_per_launch": bytes_pstep * nl * T // 4,
                      "dram_bytes_over_algorithmic": (traffic / (bytes_pstep * nl * T / 4.0)) if traffic else None,
                      "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peaks_src}}
    if m.get("f64_path"):
        # float64 batched path: kinds 1 / 2 time the whole forward / backward pass; the M^2 work is cuBLAS DGEMM
        t_f, t_b = m["kms"][1] * 1e-3, m["kms"][2] * 1e-3
        f_fwd = ev_f * ff["fwd_m2"] + ev_b * fb["fwd_m2"]
        f_bwd = ev_f * (ff["fwd_m2"] + ff["bwd_m2"]) + ev_b * (fb["fwd_m2"] + fb["bwd_m2"])   # the reverse pass recomputes a = P k
        a_t = f_bwd / t_b / 1e12
        return {"bound": "tensor", "kernel": "f64 backward pass (DGEMM + elementwise kernels)", "achieved": a_t,
                "peak": m["dgemm_peak"], "unit": "TFLOP/s", "frac": a_t / m["dgemm_peak"], "traffic": None,
                "peak_source": "FP64 GEMM rate measured on this GPU in this run (cuBLAS DGEMM 4096^3 through torch, best of 6)",
                "forward_pass": {"achieved": f_fwd / t_f / 1e12, "ms": m["kms"][1]}, "backward_pass_ms": m["kms"][2],
                "note": "M^2 FLOPs incl. the recomputation of a = P k in the reverse pass (it is a DGEMM here); the [n, M] "
                        "matrices make the elementwise kernels HBM-bound",
                "live_message_evaluations_of_2T": [live, 2 * T],
                "hbm": {"peak_gbs": peaks.get("hbm_gbs"), "peak_source": peaks_src}}
    if m["tensor_path"]:
        a_t = ev * m2 / t_s / 1e12
        a_s = ev * md / t_s / 1e12
        return {"bound": "tensor", "achieved": a_t, "peak": tens_peak, "unit": "TFLOP/s", "frac": a_t / tens_peak,
                "peak_source": f"bf16_tflops_sustained ({peaks_src}) / {TENSOR_PASSES} MMA passes per fp32-equivalent product "
                               "(two-term fp16/bf16 splits)",
                "simt": {"achieved": a_s, "peak": fp32_peak, "unit": "TFLOP/s", "frac": a_s / fp32_peak,
                         "peak_source": "packed FP32 FMA rate measured on this GPU in this run (cbf_measure_fp32_peak)"},
                "whole_step": {"tensor_frac": all_m2 / (m["ms_per_step"] * 1e-3) / 1e12 / tens_peak,
                               "simt_frac": all_md / (m["ms_per_step"] * 1e-3) / 1e12 / fp32_peak},
                "limiter": "shared-memory bandwidth and issue slots of the O(M*D) SIMT work (DESIGN.md 5.3), not the "
                           "tensor pipe or HBM", **common}
    a = ev * (m2 + md) / t_s / 1e12
    return {"bound": "fp32_simt", "achieved": a, "peak": fp32_peak, "unit": "TFLOP/s", "frac": a / fp32_peak,
            "peak_source": "packed FP32 FMA rate measured on this GPU in this run (cbf_measure_fp32_peak); the register "
                           "path is FP32-SIMT compute-bound (SURVEY 8d), not HBM- or tensor-bound",
            "whole_step": {"simt_frac": (all_m2 + all_md) / (m["ms_per_step"] * 1e-3) / 1e12 / fp32_peak}, **common}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from cbf_ssm_b200 import _lib

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
    ctx = dict(dist=dist, world=world, rank=rank, local=local, dev=dev, lib=lib, group=group)

    # FP32 roof of this GPU, measured now (rank 0's figure is reported)
    scratch = torch.empty(8 * 256 * 160, dtype=torch.float32, device=dev)
    tf = ctypes.c_double(0.0)
    _lib.check(lib.cbf_measure_fp32_peak(_lib.ptr(scratch), 20000, ctypes.byref(tf),
                                         ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    fp32_peak = float(tf.value)

    main = measure_workload(args.workload, args, args.steps, ctx)
    extra = None
    if args.workload == HEADLINE and not args.no_extra:
        extra = measure_workload(EXTRA, args, max(3, min(args.steps, 10)), ctx)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peaks_src = measured_peaks()
    try:
        traffic_table = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        traffic_table = {}

    def describe(m, key):
        w = m["w"]
        dx, dh = w["dx"], w["dx"] - w["dy"]
        return {"value": m["value"], "unit": UNIT, "ms_per_step": m["ms_per_step"], "steps": m["steps"], "warmup": m["warm"],
                "config": config_dict(w, m["bpg"], world, args.scaling),
                "l2_working_set_mib": {"workspace": m["ws_bytes"] / 2**20, "draws": 3 * (2 * w["T"]) * m["nl"] * 4 / 2**20},
                "e2e": {"value": m["e2e_value"], "unit": UNIT, "ms_per_step": m["e2e_ms"], "h2d_bytes_per_step": m["h2d"],
                        "d2h_bytes_per_step": 8, "loss": m["loss"]},
                "gpu_launches": m["launches"], "clocks": m["clocks"],
                "roofline": roofline_of(m, peaks, peaks_src, fp32_peak, traffic_table, key)}

    d = describe(main, args.workload)
    kern_ms = sum(main["kms"][i] for i in range(6))
    breakdown = {"rollout_and_accumulation_kernels_ms": kern_ms, "everything_else_ms": main["ms_per_step"] - kern_ms,
                 "note": "everything else = normal draws, both float64 GP prologues and their adjoints, reductions, the "
                         "all-reduce of the flat gradient (N > 1) and Adam; with few particles per GPU the rollout kernels do "
                         "not shrink further because a time step is a serial chain (DESIGN.md 5.7, 7)"}
    line = {"metric": METRIC, "value": d["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": main["warm"], "ms_per_step": d["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64" if main.get("f64_path") else "f32", "data": "synthetic",
            "config": d["config"], "clocks": d["clocks"],
            "e2e": d["e2e"], "gpu_launches": d["gpu_launches"], "roofline": d["roofline"],
            "l2_working_set_mib": d["l2_working_set_mib"], "step_breakdown": breakdown,
            "fp32_fma_peak_measured_tflops": fp32_peak}
    if extra is not None:
        line["extra"] = {EXTRA: describe(extra, EXTRA)}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(main["w"], 10, 3)
        if extra is not None:
            line["extra"][EXTRA]["cpu_baseline"] = cpu_baseline(extra["w"], 3, 1, with_five=False)
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
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: global batch = batch x N; strong: global batch = batch, split over the N ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra.robomove_m20 measurement")
    ap.add_argument("--M", type=int, default=0, help="override the workload's number of inducing points (kernel-path studies)")
    ap.add_argument("--S", type=int, default=0, help="override the particles per sequence")
    ap.add_argument("--T", type=int, default=0, help="override the sequence length")
    ap.add_argument("--R", type=int, default=0, help="override recog_len")
    ap.add_argument("--flags", type=int, default=0, help="cbf_shape.flags (1 cooperative kernels, 2 no tensor cores)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.batch <= 0:
            args.batch = WORKLOADS[args.workload]["batch"]
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
