#!/usr/bin/env python
"""bench.py -- batched KalmanODE solves/s on the headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path over one batch: FitzHugh-Nagumo (n_state 6, n_meas 2),
n_eval 400, 65,536 theta draws per GPU, solve_mv (posterior mean + variance for every
system and time point).  For N > 1 launch under torchrun (one rank per GPU); the batch is
replicated per rank (weak scaling), no collective on the data path.

Prints ONE JSON line (rank 0).  `value` times the device-pointer entry point with inputs
resident in HBM; `e2e` times the host-buffer entry point (pinned host buffers, H2D of the
inputs and D2H of every output inside the timed region); `roofline` is the dominant
(backward) kernel against the measured HBM bandwidth with SURVEY section 8(d)'s algorithmic bytes;
`cpu_baseline` is the C oracle (a port of the reference's arithmetic) on the host cores.
`--impl reference` times that CPU port alone with all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched KalmanODE solves/sec (FitzHugh-Nagumo, 64k theta)"
UNIT = "solves/s"
BATCH = 65536
N_EVAL, N_STATE, N_MEAS = 400, 6, 2
# SURVEY section 8(d), general mode, per system per step: the reference's four dense history
# arrays written once by the filter and read once by the smoother, plus the solve_mv output.
HIST_BYTES_STEP = 8 * (2 * N_STATE + 2 * N_STATE * N_STATE)          # 672
OUT_BYTES_STEP = 8 * (N_STATE + N_STATE * N_STATE)                   # 336
ALG_BYTES_SOLVE = N_EVAL * (2 * HIST_BYTES_STEP + OUT_BYTES_STEP)    # 672,000
ALG_BYTES_BACKWARD = N_EVAL * (HIST_BYTES_STEP + OUT_BYTES_STEP)     # the smoother kernel's share
ALG_BYTES_FORWARD = N_EVAL * HIST_BYTES_STEP


def workload_config(extra=None):
    c = {"workload": "fitzhugh-nagumo solve_mv, n_state 6, n_meas 2, n_eval 400, "
                     "65536 theta draws per GPU (BASELINE configs[1])",
         "batch_per_gpu": BATCH, "n_eval": N_EVAL, "mode": "solve_mv", "prior": "ibm q=3, sigma 0.1",
         "l2": "no flush: each step streams >3.7 GB of filter history (written, then read back) and "
               "8.8 GB of output per GPU, >100x the 126 MB L2"}
    if extra:
        c.update(extra)
    return c


REF_FLOPS_PER_SOLVE = 3621 * 400          # SURVEY 8(d), C2, general mode


def measured_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/traffic.json, written by tools/ncu_traffic.py from `ncu --set full`), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


def write_pattern_peak():
    """GB/s a no-compute kernel reaches writing the variance OUTPUT PATTERN of this workload (65,536
    interleaved streams of 288-byte runs, tools/microbench/write_pattern.cu), measured live; None if
    the binary is missing.  The output layout [b][t][n][n] is the reference's; this is what the memory
    system sustains for it, against 7.4 TB/s for a contiguous fill."""
    exe = os.path.join(ROOT, "tools", "microbench", "write_pattern")
    if not os.path.exists(exe):
        return None
    try:
        import subprocess
        out = subprocess.run([exe, "--quick"], capture_output=True, text=True, timeout=60).stdout
        return float(json.loads(out.strip().splitlines()[0])["GBps"])
    except Exception:
        return None


def fp64_peak():
    """Sustained FP64 FMA rate of this GPU in TFLOP/s (tools/microbench/fp64_peak.cu, measured live; the
    best of its occupancy settings), or None."""
    exe = os.path.join(ROOT, "tools", "microbench", "fp64_peak")
    if not os.path.exists(exe):
        return None
    try:
        import subprocess
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        vals = [json.loads(l).get("fp64_TFLOPs") for l in out.strip().splitlines()]
        return max(v for v in vals if v)
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:            # NVML missing: report that, do not fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(2.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_baseline(seconds=10.0, chunk=1024, nthreads=0):
    """The C oracle (port of rodeo/eigen/KalmanTV.h + KalmanTVODE.h), OpenMP over the batch,
    on a bounded sample of the SAME workload: chunks of `chunk` solves until `seconds` elapse."""
    import numpy as np
    import oracle
    from tests import workloads as wl
    oracle.build()
    flags = oracle.select_timing_build()
    w, x0, theta = wl.fitz_batch(chunk * 64, seed=0)
    p = wl.make_oracle_problem(w)
    if nthreads <= 0:            # torchrun exports OMP_NUM_THREADS=1: ask for every usable core
        nthreads = host_cores()
    cores = nthreads
    oracle.solve_batch(p, x0[:max(cores, 8)], theta[:max(cores, 8)], None, oracle.MODE_MV, nthreads)
    done, t0 = 0, time.perf_counter()
    while True:
        lo = done % (chunk * 64)
        oracle.solve_batch(p, x0[lo:lo + chunk], theta[lo:lo + chunk], None, oracle.MODE_MV, nthreads)
        done += chunk
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return {"value": done / el, "unit": UNIT, "cores": int(cores), "kind": "port",
            "sample": f"{done} of the workload's solves (chunks of {chunk}), {el:.1f} s, "
                      f"C port of the reference arithmetic, {flags}, OpenMP over systems"}, done, el


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  Its compiled backends
    cannot be built in this image (Eigen and `kalmantv` are absent), so this is the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    chunk = 2048
    import numpy as np
    import oracle
    from tests import workloads as wl
    oracle.build()
    flags = oracle.select_timing_build()
    w, x0, theta = wl.fitz_batch(chunk, seed=0)
    p = wl.make_oracle_problem(w)
    cores = host_cores()         # explicit: torchrun exports OMP_NUM_THREADS=1
    for _ in range(max(1, min(args.warmup, 2))):
        oracle.solve_batch(p, x0, theta, None, oracle.MODE_MV, cores)
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.solve_batch(p, x0, theta, None, oracle.MODE_MV, cores)
    el = time.perf_counter() - t0
    v = steps * chunk / el
    sample = (f"each step = {chunk} of the workload's 65536 solves; {steps} steps timed "
              f"(capped at 20 so the run ends within minutes); C port of the reference arithmetic, {flags}")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config({"reference_sample": sample}),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": int(cores), "kind": "port",
                             "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def host_memory_ok(nbytes):
    try:
        avail = None
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
        lim = None
        for pth in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
            if os.path.exists(pth):
                s = open(pth).read().strip()
                if s.isdigit():
                    lim = int(s)
        cap = min(x for x in (avail, lim) if x is not None)
        return nbytes * 2.5 < cap
    except Exception:
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--dense", action="store_true",
                    help="force the general dense kernel family (default: block-diagonal is auto-selected)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    from tests import workloads as wl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    w, x0, theta = wl.fitz_batch(B, seed=rank)
    k = wl.make_kode(w, force_dense=args.dense)
    family = "/".join(k.kernel_family)
    k.reserve(B)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()

    # ---- device-resident throughput ("value") ---------------------------------------------
    for _ in range(args.warmup):
        mu, var = k.solve_mv(x0d, theta=thd)
    barrier()
    k.set_timing(True)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = k.launch_count
    fwd_ms = bwd_ms = 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        mu, var = k.solve_mv(x0d, theta=thd)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = k.launch_count - launches0
    # per-kernel device times of the last timed step (events on the launching stream)
    f, b = k.last_kernel_ms()
    fwd_ms, bwd_ms = f, b
    k.set_timing(False)
    status_bad = int(k.status.abs().sum())
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer entry point ------------------------------------
    e2e = None
    if not args.no_e2e:
        del mu, var
        torch.cuda.empty_cache()
        # several ranks share one host: keep the pinned footprint of the e2e leg modest there
        # (throughput is PCIe-bound per GPU and does not depend on the batch size)
        Be = B if world == 1 else min(B, 16384)
        out_bytes = Be * (N_EVAL + 1) * (N_STATE + N_STATE * N_STATE) * 8
        while Be > 4096 and not host_memory_ok(out_bytes * max(1, min(world, 8))):
            Be //= 2
            out_bytes = Be * (N_EVAL + 1) * (N_STATE + N_STATE * N_STATE) * 8
        pin = lambda *s: torch.empty(s, dtype=torch.float64, pin_memory=True)
        x0p, thp = pin(Be, N_STATE), pin(Be, 3)
        x0p.copy_(torch.from_numpy(x0[:Be]))
        thp.copy_(torch.from_numpy(theta[:Be]))
        mup, varp = pin(Be, N_EVAL + 1, N_STATE), pin(Be, N_EVAL + 1, N_STATE, N_STATE)
        x0n, thn, mun, varn = x0p.numpy(), thp.numpy(), mup.numpy(), varp.numpy()
        k.solve_mv_into(x0n, thn, mun, varn)           # warm-up (allocates staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            k.solve_mv_into(x0n, thn, mun, varn)       # synchronous: returns with results on the host
        barrier()
        el = time.perf_counter() - t0
        te = torch.tensor([el], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        el = float(te.item())
        e2e = {"value": world * Be * args.e2e_steps / el, "unit": UNIT,
               "h2d_bytes_per_step": Be * (N_STATE + 3) * 8,
               "d2h_bytes_per_step": out_bytes + Be * 4,
               "batch_per_gpu": Be, "steps": args.e2e_steps, "ms_per_step": 1e3 * el / args.e2e_steps,
               "host_buffers": "pinned", "checksum": float(mun[:, -1, 0].sum())}
        del mup, varp

    # ---- the general dense kernel family on the same workload (extra line, fewer steps) -------
    dense = None
    if not args.dense and world == 1:
        try:
            del mu, var
        except NameError:
            pass
        torch.cuda.empty_cache()
        kd = wl.make_kode(w, force_dense=True)
        kd.reserve(B)
        for _ in range(3):
            kd.solve_mv(x0d, theta=thd)
        torch.cuda.synchronize()
        kd.set_timing(True)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nd = max(3, args.steps // 3)
        d0.record()
        for _ in range(nd):
            kd.solve_mv(x0d, theta=thd)
        d1.record()
        torch.cuda.synchronize()
        dms = d0.elapsed_time(d1) / nd
        df, db = kd.last_kernel_ms()
        dense = {"family": "/".join(kd.kernel_family), "value": B / (dms * 1e-3), "unit": UNIT,
                 "ms_per_step": dms, "forward_kernel_ms": df, "backward_kernel_ms": db, "steps": nd,
                 "backward_achieved_GBps": ALG_BYTES_BACKWARD * B / (db * 1e-3) / 1e9}
        del kd
        torch.cuda.empty_cache()

    # ---- shared-covariance mode on the same workload: a DIFFERENT amount of work per system
    # (covariances once per prior, mean recursions per system; var returned as one shared array),
    # reported on its own line with its own byte count -- never against the general-mode roofline
    shared = None
    if world == 1 and not args.dense:
        try:
            del mu, var
        except NameError:
            pass
        torch.cuda.empty_cache()
        ks = wl.make_kode(w, shared_cov=True)
        for _ in range(3):
            ks.solve_mv(x0d, theta=thd)
        torch.cuda.synchronize()
        ks.set_timing(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ns = args.steps
        s0.record()
        for _ in range(ns):
            ks.solve_mv(x0d, theta=thd)
        s1.record()
        torch.cuda.synchronize()
        sms = s0.elapsed_time(s1) / ns
        sf, sb = ks.last_kernel_ms()
        sc_bytes = N_EVAL * 8 * 3 * N_STATE * B          # mu_filt written + read, mu_smooth written
        shared = {"mode": "shared covariance (SURVEY F5): tables cached per prior, mean recursions per system, "
                          "var = one (N+1,n,n) array for the batch",
                  "value": B / (sms * 1e-3), "unit": UNIT, "ms_per_step": sms, "forward_kernel_ms": sf,
                  "backward_kernel_ms": sb, "steps": ns,
                  "algorithmic_bytes_per_step": sc_bytes, "achieved_GBps": sc_bytes / (sms * 1e-3) / 1e9,
                  "frac_of_hbm_peak": sc_bytes / (sms * 1e-3) / 1e9 / peaks()[0]}
        if not args.no_e2e:
            Bs = B
            pin = lambda *s_: torch.empty(s_, dtype=torch.float64, pin_memory=True)
            x0p, thp = pin(Bs, N_STATE), pin(Bs, 3)
            x0p.copy_(torch.from_numpy(x0))
            thp.copy_(torch.from_numpy(theta))
            mup, varp = pin(Bs, N_EVAL + 1, N_STATE), pin(N_EVAL + 1, N_STATE, N_STATE)
            ks.solve_mv_into(x0p.numpy(), thp.numpy(), mup.numpy(), varp.numpy())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                ks.solve_mv_into(x0p.numpy(), thp.numpy(), mup.numpy(), varp.numpy())
            el = time.perf_counter() - t0
            shared["e2e"] = {"value": Bs * args.e2e_steps / el, "unit": UNIT,
                             "h2d_bytes_per_step": Bs * (N_STATE + 3) * 8,
                             "d2h_bytes_per_step": Bs * (N_EVAL + 1) * N_STATE * 8
                                                   + (N_EVAL + 1) * N_STATE * N_STATE * 8 + Bs * 4,
                             "ms_per_step": 1e3 * el / args.e2e_steps}
            del mup, varp
        del ks
        torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _, _ = cpu_baseline(args.cpu_seconds)

    if rank == 0:
        peak, peak_src = peaks()
        key = "backward_mv_fitz6_" + ("dense" if args.dense else "blockdiag")
        traffic = measured_traffic(key)
        wpat = write_pattern_peak()
        f64 = fp64_peak()
        bwd_gbs = ALG_BYTES_BACKWARD * B / (bwd_ms * 1e-3) / 1e9
        fwd_gbs = ALG_BYTES_FORWARD * B / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else None
        step_gbs = ALG_BYTES_SOLVE * B * args.steps / (ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config({"parallelism": f"batch replicated x{world} (weak), no collective"}),
            "roofline": {"bound": "hbm", "kernel": f"kalman_backward_staged (smoother, solve_mv, {family})",
                         "achieved": bwd_gbs, "peak": peak, "unit": "GB/s", "frac": bwd_gbs / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "dram_frac": (traffic / (bwd_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "write_pattern_peak": wpat,
                         "frac_of_write_pattern_peak": (traffic / (bwd_ms * 1e-3) / 1e9 / wpat) if traffic and wpat else None,
                         "note": "achieved = SURVEY 8(d) algorithmic bytes (the reference's four dense "
                                 "history arrays + solve_mv output) / kernel time; the kernel moves fewer "
                                 "bytes (traffic) because predicted moments are recomputed and covariances "
                                 "are stored packed, so frac can exceed 1; dram_frac = traffic / time / peak; "
                                 "write_pattern_peak = GB/s of a no-compute kernel writing this workload's "
                                 "288-byte-run output pattern (tools/microbench/write_pattern.cu, measured in "
                                 "this run): 82 % of the kernel's traffic is that pattern",
                         "fp64": {"peak_TFLOPs": f64, "source": "tools/microbench/fp64_peak.cu, measured in this run "
                                                                "(nominal 37.2)",
                                  "reference_flops_per_solve": REF_FLOPS_PER_SOLVE,
                                  "reference_equivalent_TFLOPs": value / world * REF_FLOPS_PER_SOLVE / 1e12,
                                  "note": "SURVEY 8(d): 3621 flop/step x 400 steps as the reference's dense "
                                          "KalmanTV executes them; the block-diagonal kernels skip the structural "
                                          "zeros, so this can exceed the peak -- the second bound, never the binding one"},
                         "algorithmic_bytes_per_launch": ALG_BYTES_BACKWARD * B,
                         "kernel_ms": bwd_ms, "forward_kernel_ms": fwd_ms,
                         "forward_achieved": fwd_gbs,
                         "whole_step_achieved": step_gbs, "whole_step_frac": step_gbs / peak},
            "kernel_family": family, "dense_path": dense, "shared_covariance_path": shared,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "status_nonzero": status_bad,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
