#!/usr/bin/env python
"""bench.py -- batched KalmanODE solves/s on the headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path over one batch: FitzHugh-Nagumo (n_state 6, n_meas 2),
n_eval 400, 65,536 theta draws per GPU, solve_mv (posterior mean + variance for every
system and time point).  For N > 1 launch under torchrun (one rank per GPU); the batch is
replicated per rank (weak scaling), no collective on the data path.

Prints ONE JSON line (rank 0):
  value        device-pointer entry point, inputs resident in HBM, CUDA events, max over ranks
  e2e          host-buffer entry point (pinned host buffers; H2D of the inputs and D2H of every output inside the
               timed region), the SAME 65,536-system batch per GPU at every N, with a per-rank H2D / D2H probe
  roofline     the dominant (smoother) kernel: `frac` = DRAM bytes it moves / its average time over the timed
               steps / measured HBM copy bandwidth -- the bytes are computed from the data layout (history records
               read + outputs written) and checked against the committed ncu figure; `algorithmic_frac` keeps
               SURVEY section 8(d)'s reference-layout bytes beside it.  The forward kernel has the same block.
  c5_sweep     BASELINE configs[4]: the 1,048,576-theta log-likelihood sweep, STRONG-scaled (sharded over the N ranks)
               with its one NCCL all_gather (8 bytes per system) inside the CUDA-event region
  cpu_baseline the C oracle (a port of the reference's arithmetic) on the host cores (N = 1 only)
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
C5_TOTAL = 1 << 20
N_EVAL, N_STATE, N_MEAS = 400, 6, 2
# SURVEY section 8(d), general mode, per system per step: the reference's four dense history
# arrays written once by the filter and read once by the smoother, plus the solve_mv output.
HIST_BYTES_STEP = 8 * (2 * N_STATE + 2 * N_STATE * N_STATE)          # 672
OUT_BYTES_STEP = 8 * (N_STATE + N_STATE * N_STATE)                   # 336
ALG_BYTES_SOLVE = N_EVAL * (2 * HIST_BYTES_STEP + OUT_BYTES_STEP)    # 672,000
ALG_BYTES_BACKWARD = N_EVAL * (HIST_BYTES_STEP + OUT_BYTES_STEP)     # the smoother kernel's share
ALG_BYTES_FORWARD = N_EVAL * HIST_BYTES_STEP
REF_FLOPS_PER_SOLVE = 3621 * 400          # SURVEY 8(d), C2, general mode


def workload_config(extra=None):
    c = {"workload": "fitzhugh-nagumo solve_mv, n_state 6, n_meas 2, n_eval 400, "
                     "65536 theta draws per GPU (BASELINE configs[1])",
         "batch_per_gpu": BATCH, "n_eval": N_EVAL, "mode": "solve_mv", "prior": "ibm q=3, sigma 0.1",
         "l2": "no flush: each step streams >1.8 GB of filter history (written, then read back) and "
               "8.8 GB of output per GPU, >80x the 126 MB L2"}
    if extra:
        c.update(extra)
    return c


def layout_traffic(kode, B):
    """DRAM bytes one launch of each kernel moves, FROM THE DATA LAYOUT: the forward
    kernel writes, and the smoother reads, R records of `hist_doubles` doubles per system (R = checkpointed time
    indices); the smoother writes the solve_mv output.  Inputs (x0, theta) are < 0.1 %."""
    hist = kode.history_bytes_per_system * B
    out = (N_EVAL + 1) * OUT_BYTES_STEP * B
    return {"forward": hist, "backward": hist + out, "history": hist, "output": out}


def ncu_traffic(kernel_key):
    """DRAM bytes per launch from the committed ncu capture of HEAD's kernels (profiles/traffic.json, written by
    tools/ncu_traffic.py from `ncu --set full`), or None: cross-check of layout_traffic(), not its source."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_key)
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


def oracle_problem(w):
    import oracle
    return oracle.Problem(w["ode"], w["W"], w["tmin"], w["tmax"], w["N"], w["Q"], w["c"], w["R"])


def cpu_baseline(seconds=10.0, chunk=1024, nthreads=0):
    """The C oracle (port of rodeo/eigen/KalmanTV.h + KalmanTVODE.h), OpenMP over the batch,
    on a bounded sample of the SAME workload: chunks of `chunk` solves until `seconds` elapse."""
    import oracle
    from rodeo_legacy_b200 import workloads as wl
    oracle.build()
    flags = oracle.select_timing_build()
    w, x0, theta = wl.fitz_batch(chunk * 64, seed=0)
    p = oracle_problem(w)
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
    import oracle
    from rodeo_legacy_b200 import workloads as wl
    oracle.build()
    flags = oracle.select_timing_build()
    w, x0, theta = wl.fitz_batch(chunk, seed=0)
    p = oracle_problem(w)
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
        return nbytes * 1.25 < cap
    except Exception:
        return False


def bind_near_gpu(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index` BEFORE any pinned host memory is
    allocated (first-touch places the staging buffers on that NUMA node).  Returns a short description."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        nv.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return f"{len(cpus)} cpus [{cpus[0]}..{cpus[-1]}] (nvmlDeviceSetCpuAffinity)"
    except Exception as e:
        return f"unbound ({type(e).__name__})"


_HOST_KEEP = []


def host_buffer(torch, shape, mode="pinned"):
    """float64 host tensor for the e2e leg: "pinned" = cudaHostAlloc through torch; "thp" = anonymous mmap with
    MADV_HUGEPAGE, touched, then cudaHostRegister'ed (2 MB pages: 512x fewer IOMMU / TLB entries for the DMA engines
    than 4 KB pages when several GPUs copy into tens of GB of host memory at once)."""
    import numpy as np
    if mode == "pinned":
        return torch.empty(shape, dtype=torch.float64, pin_memory=True)
    import mmap
    n = int(np.prod(shape)) * 8
    huge = 2 << 20
    m = mmap.mmap(-1, n + huge)
    try:
        m.madvise(mmap.MADV_HUGEPAGE)
    except Exception:
        pass
    raw = np.frombuffer(m, dtype=np.uint8)
    off = (-raw.ctypes.data) % huge
    arr = raw[off:off + n].view(np.float64).reshape(shape)
    arr[...] = 0.0                                   # fault the pages in (as huge pages where the kernel allows)
    t = torch.from_numpy(arr)
    rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), n, 0)
    if int(rc) != 0:
        raise RuntimeError(f"cudaHostRegister failed: {rc}")
    _HOST_KEEP.append((m, t))
    return t


def copy_probe(torch, nbytes=1 << 30, reps=3):
    """H2D / D2H GB/s of this rank between pinned host memory and its GPU (best of `reps`), all ranks at once."""
    host = torch.empty(nbytes // 8, dtype=torch.float64, pin_memory=True)
    dev = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    out = {}
    for name, (dst, src) in {"h2d": (dev, host), "d2h": (host, dev)}.items():
        best = 0.0
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        out[name] = best
    del host, dev
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-host", default="pinned", choices=["pinned", "thp"],
                    help="host buffers of the e2e leg: cudaHostAlloc, or huge-page mmap + cudaHostRegister")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the dense / shared-covariance / C5 side measurements")
    ap.add_argument("--checkpoint", type=int, default=0, help="checkpoint interval of the filter history (0 = auto)")
    ap.add_argument("--no-structure", action="store_true",
                    help="general block-diagonal kernels (no upper-triangular-Q / unit-w specialisation)")
    ap.add_argument("--dense", action="store_true",
                    help="force the general dense kernel family (default: block-diagonal is auto-selected)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    # stdout carries exactly ONE line (the JSON): anything libraries print there (NCCL's version banner, ...)
    # goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    from rodeo_legacy_b200 import workloads as wl
    from rodeo_legacy_b200.dist import gather_batch, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    binding = bind_near_gpu(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(v):
        return -max_over_ranks(-v)

    B = args.batch
    w, x0, theta = wl.fitz_batch(B, seed=rank)
    k = wl.make_kode(w, force_dense=args.dense, use_structure=not args.no_structure, checkpoint=args.checkpoint)
    family = "/".join(k.kernel_family) + ("/structured" if k.structured else "")
    k.reserve(B)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()

    # ---- device-resident throughput ("value") ---------------------------------------------
    for _ in range(args.warmup):
        mu, var = k.solve_mv(x0d, theta=thd)
    barrier()
    k.set_timing(True)           # per-kernel events inside the library, on the launching stream, EVERY timed step
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = k.launch_count
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
    f_tot, b_tot, n_timed = k.kernel_ms_total()
    fwd_ms, bwd_ms = f_tot / n_timed, b_tot / n_timed      # averages over the timed steps
    k.set_timing(False)
    status_bad = int(k.status.abs().sum())
    ms = max_over_ranks(ms)
    value = world * B * args.steps / (ms * 1e-3)
    traffic = layout_traffic(k, B)
    del mu, var
    torch.cuda.empty_cache()

    # ---- end to end through the host-buffer entry point: the same batch per GPU at every N ------------
    e2e = None
    if not args.no_e2e:
        Be = B
        out_bytes = Be * (N_EVAL + 1) * OUT_BYTES_STEP
        fits = host_memory_ok(out_bytes * world + (1 << 30) * world)
        fits = min_over_ranks(1.0 if fits else 0.0) > 0.5
        if not fits:       # never silently: the reduced batch is named in the line
            while Be > 4096 and not host_memory_ok(out_bytes * world):
                Be //= 2
                out_bytes = Be * (N_EVAL + 1) * OUT_BYTES_STEP
            Be = int(min_over_ranks(float(Be)))
            out_bytes = Be * (N_EVAL + 1) * OUT_BYTES_STEP
        probe = copy_probe(torch)
        pin = lambda *s: torch.empty(s, dtype=torch.float64, pin_memory=True)
        x0p, thp = pin(Be, N_STATE), pin(Be, 3)
        x0p.copy_(torch.from_numpy(x0[:Be]))
        thp.copy_(torch.from_numpy(theta[:Be]))
        mup = host_buffer(torch, (Be, N_EVAL + 1, N_STATE), args.e2e_host)
        varp = host_buffer(torch, (Be, N_EVAL + 1, N_STATE, N_STATE), args.e2e_host)
        x0n, thn, mun, varn = x0p.numpy(), thp.numpy(), mup.numpy(), varp.numpy()
        k.solve_mv_into(x0n, thn, mun, varn)           # warm-up (allocates staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            k.solve_mv_into(x0n, thn, mun, varn)       # synchronous: returns with results on the host
        barrier()
        el = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * Be * args.e2e_steps / el, "unit": UNIT,
               "h2d_bytes_per_step": Be * (N_STATE + 3) * 8,
               "d2h_bytes_per_step": out_bytes + Be * 4,
               "batch_per_gpu": Be, "steps": args.e2e_steps, "ms_per_step": 1e3 * el / args.e2e_steps,
               "host_buffers": "pinned (cudaHostAlloc)" if args.e2e_host == "pinned" else
                               "2 MB transparent huge pages, cudaHostRegister", "cpu_binding": binding,
               "copy_probe_GBps": {"h2d_min_over_ranks": min_over_ranks(probe["h2d"]),
                                   "d2h_min_over_ranks": min_over_ranks(probe["d2h"]),
                                   "d2h_sum_over_ranks": None, "note": "1 GiB pinned <-> device, all ranks at once"},
               "d2h_GBps_achieved_per_gpu": (out_bytes + Be * 4) * args.e2e_steps / el / 1e9,
               "limiter": "PCIe D2H of the full (B, N+1, n, n) variance output: the step moves 8.83 GB per GPU to the host",
               "checksum": float(mun[:, -1, 0].sum())}
        t = torch.tensor([probe["d2h"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e2e["copy_probe_GBps"]["d2h_sum_over_ranks"] = float(t.item())
        del mup, varp, x0p, thp

    # ---- BASELINE configs[4]: 1,048,576-theta log-likelihood sweep, sharded over the ranks, one NCCL gather ----
    c5 = None
    if not args.no_extras and not args.dense:
        torch.cuda.empty_cache()
        lo, hi = shard_range(C5_TOTAL, rank, world)
        th5 = torch.from_numpy(wl.fitz_sweep_theta(C5_TOTAL, lo, hi, w)).cuda()
        x05 = torch.from_numpy(np.tile(w["x0"], (hi - lo, 1))).cuda()
        Y, ind, sind = wl.fitz_observations(k, w)
        reps = max(3, min(args.steps, 10))
        for _ in range(2):
            ll = gather_batch(k.loglik(x05, th5, Y, ind, sind, 0.2), C5_TOTAL)
        barrier()
        c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        tot = gat = 0.0
        for _ in range(reps):
            barrier()
            c0.record()
            ll_local = k.loglik(x05, th5, Y, ind, sind, 0.2)
            c1.record()
            ll = gather_batch(ll_local, C5_TOTAL)          # all_gather_into_tensor, 8 bytes per system
            c2.record()
            torch.cuda.synchronize()
            tot += max_over_ranks(c0.elapsed_time(c2))
            gat += max_over_ranks(c1.elapsed_time(c2))
        c5_ms, gather_ms = tot / reps, gat / reps
        c5 = {"workload": "fitzhugh-nagumo log-likelihood sweep, 1,048,576 theta, thinned to t = 1..40, gamma 0.2 "
                          "(BASELINE configs[4]; examples/inference.py:85-94)",
              "value": C5_TOTAL / (c5_ms * 1e-3), "unit": UNIT, "scaling": "strong", "n_gpus": world,
              "ms_per_sweep": c5_ms, "gather_ms": gather_ms, "systems_per_gpu": hi - lo, "reps": reps,
              "parallelism": f"sharded x{world} (contiguous slices), one all_gather_into_tensor of 8 B/system "
                             "inside the timed region" if world > 1 else "one GPU, no collective",
              "gathered": int(ll.numel()), "argmax_theta_index": int(ll.argmax()),
              "loglik_checksum": float(ll.sum())}
        del th5, x05, ll

    # ---- side measurements on one GPU: other kernel families, other entry points ------------------------
    dense = shared = e2e_other = None
    if world == 1 and not args.no_extras and not args.dense:
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
        df, db, dn = kd.kernel_ms_total()
        dtr = layout_traffic(kd, B)
        dense = {"family": "/".join(kd.kernel_family), "value": B / (dms * 1e-3), "unit": UNIT,
                 "ms_per_step": dms, "forward_kernel_ms": df / dn, "backward_kernel_ms": db / dn, "steps": nd,
                 "backward_traffic": dtr["backward"],
                 "backward_frac": dtr["backward"] / (db / dn * 1e-3) / 1e9 / peaks()[0]}
        del kd
        torch.cuda.empty_cache()

        # shared-covariance mode: a DIFFERENT amount of work per system (covariances once per prior, mean
        # recursions per system; var returned as one shared array) -- its own line, its own byte count
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
        sf, sb, sn = ks.kernel_ms_total()
        sc_bytes = N_EVAL * 8 * 3 * N_STATE * B          # mu_filt written + read, mu_smooth written
        shared = {"mode": "shared covariance (SURVEY F5): tables cached per prior, mean recursions per system, "
                          "var = one (N+1,n,n) array for the batch",
                  "value": B / (sms * 1e-3), "unit": UNIT, "ms_per_step": sms, "forward_kernel_ms": sf / sn,
                  "backward_kernel_ms": sb / sn, "steps": ns,
                  "algorithmic_bytes_per_step": sc_bytes, "achieved_GBps": sc_bytes / (sms * 1e-3) / 1e9,
                  "frac_of_hbm_peak": sc_bytes / (sms * 1e-3) / 1e9 / peaks()[0]}
        if not args.no_e2e:
            pin = lambda *s_: torch.empty(s_, dtype=torch.float64, pin_memory=True)
            x0p, thp = pin(B, N_STATE), pin(B, 3)
            x0p.copy_(torch.from_numpy(x0))
            thp.copy_(torch.from_numpy(theta))
            mup, varp = pin(B, N_EVAL + 1, N_STATE), pin(N_EVAL + 1, N_STATE, N_STATE)
            ks.solve_mv_into(x0p.numpy(), thp.numpy(), mup.numpy(), varp.numpy())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                ks.solve_mv_into(x0p.numpy(), thp.numpy(), mup.numpy(), varp.numpy())
            el = time.perf_counter() - t0
            shared["e2e"] = {"value": B * args.e2e_steps / el, "unit": UNIT,
                             "h2d_bytes_per_step": B * (N_STATE + 3) * 8,
                             "d2h_bytes_per_step": B * (N_EVAL + 1) * N_STATE * 8
                                                   + (N_EVAL + 1) * N_STATE * N_STATE * 8 + B * 4,
                             "ms_per_step": 1e3 * el / args.e2e_steps}
            del mup, varp
        del ks
        torch.cuda.empty_cache()

        # the other two entry points end to end (host buffers in, host buffers out), general mode
        if not args.no_e2e:
            z = np.asfortranarray(np.random.default_rng(2).standard_normal((N_STATE, N_EVAL)))
            kz = wl.make_kode(w, z_state=z)
            xs = kz.solve_sim(x0, theta=theta)                 # NumPy in -> NumPy out (pageable result)
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                xs = kz.solve_sim(x0, theta=theta)
            el_sim = (time.perf_counter() - t0) / args.e2e_steps
            xp = torch.empty((B, N_EVAL + 1, N_STATE), dtype=torch.float64, pin_memory=True)   # allocated once by the caller
            kz.solve_sim_into(x0, theta, xp.numpy())
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                kz.solve_sim_into(x0, theta, xp.numpy())
            el_simp = (time.perf_counter() - t0) / args.e2e_steps
            sim_same = bool(np.array_equal(xp.numpy(), xs))
            del xp
            Y, ind, sind = wl.fitz_observations(kz, w)
            lln = kz.loglik(x0, theta, Y, ind, sind, 0.2)      # NumPy in -> NumPy out
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                lln = kz.loglik(x0, theta, Y, ind, sind, 0.2)
            el_ll = (time.perf_counter() - t0) / args.e2e_steps
            e2e_other = {
                "solve_sim": {"value": B / el_sim, "unit": UNIT, "ms_per_step": 1e3 * el_sim,
                              "h2d_bytes_per_step": B * (N_STATE + 3) * 8 + z.nbytes,
                              "d2h_bytes_per_step": B * (N_EVAL + 1) * N_STATE * 8 + B * 4,
                              "host_buffers": "pageable NumPy arrays (the call a reference user makes)"},
                "solve_sim_into": {"value": B / el_simp, "unit": UNIT, "ms_per_step": 1e3 * el_simp,
                                   "h2d_bytes_per_step": B * (N_STATE + 3) * 8 + z.nbytes,
                                   "d2h_bytes_per_step": B * (N_EVAL + 1) * N_STATE * 8 + B * 4,
                                   "host_buffers": "result into a pinned buffer the caller allocated once",
                                   "same_draws_as_solve_sim": sim_same},
                "loglik": {"value": B / el_ll, "unit": UNIT, "ms_per_step": 1e3 * el_ll,
                           "h2d_bytes_per_step": B * (N_STATE + 3) * 8, "d2h_bytes_per_step": B * 12,
                           "host_buffers": "pageable NumPy arrays", "checksum": float(np.sum(lln))}}
            del kz, xs
            # opt-in compact variance layouts (rodeo_cuda_set_var_layout): the same solve_mv, fewer output bytes
            for lay, vw in (("blocks", 2 * 9), ("diag", 6)):
                kl = wl.make_kode(w, var_layout=lay)
                pin = lambda *s_: torch.empty(s_, dtype=torch.float64, pin_memory=True)
                x0p, thp = pin(B, N_STATE), pin(B, 3)
                x0p.copy_(torch.from_numpy(x0))
                thp.copy_(torch.from_numpy(theta))
                mup = pin(B, N_EVAL + 1, N_STATE)
                varp = pin(B, N_EVAL + 1, 2, 3, 3) if lay == "blocks" else pin(B, N_EVAL + 1, N_STATE)
                kl.solve_mv_into(x0p.numpy(), thp.numpy(), mup.numpy(), varp.numpy())
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    kl.solve_mv_into(x0p.numpy(), thp.numpy(), mup.numpy(), varp.numpy())
                el_l = (time.perf_counter() - t0) / args.e2e_steps
                for _ in range(3):
                    kl.solve_mv(x0d, theta=thd)
                torch.cuda.synchronize()
                l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                l0.record()
                for _ in range(10):
                    kl.solve_mv(x0d, theta=thd)
                l1.record()
                torch.cuda.synchronize()
                e2e_other["solve_mv_var_" + lay] = {
                    "value": B / el_l, "unit": UNIT, "ms_per_step": 1e3 * el_l,
                    "h2d_bytes_per_step": B * (N_STATE + 3) * 8,
                    "d2h_bytes_per_step": B * (N_EVAL + 1) * (N_STATE + vw) * 8 + B * 4,
                    "host_buffers": "pinned", "device_resident_value": B * 10 / (l0.elapsed_time(l1) * 1e-3),
                    "note": "opt-in output layout, never the headline: var = " +
                            ("the two (3, 3) diagonal blocks" if lay == "blocks" else "the marginal variances") +
                            " of the reference's (6, 6) matrix, same numbers"}
                del kl, mup, varp

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _, _ = cpu_baseline(args.cpu_seconds)

    if rank == 0:
        peak, peak_src = peaks()
        key = "backward_mv_fitz6_" + ("dense" if args.dense else "blockdiag")
        tr_ncu = ncu_traffic(key)
        tr = traffic["backward"]
        if args.checkpoint not in (0, 2) or args.no_structure:
            tr_ncu = None      # the committed capture is of the default build (checkpoint 2)
        if tr_ncu:       # the layout figure must describe what the hardware counted for this kernel
            assert abs(tr - tr_ncu) <= 0.03 * tr_ncu, (tr, tr_ncu)
        f64 = fp64_peak() if not args.no_extras else None
        bwd_s, fwd_s = bwd_ms * 1e-3, fwd_ms * 1e-3
        step_s = ms * 1e-3 / args.steps
        tpb = family.startswith("blockdiag/registers")
        kname = "kalman_backward_tpb" if tpb else "kalman_backward_staged"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config({"parallelism": f"batch replicated x{world} (weak), no collective on the "
                                                      "solve_mv path; the sharded sweep with its NCCL gather is c5_sweep"}),
            "roofline": {"bound": "hbm", "kernel": f"{kname} (smoother, solve_mv, {family})",
                         "achieved": tr / bwd_s / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": tr / bwd_s / 1e9 / peak, "traffic": tr, "peak_source": peak_src,
                         "traffic_source": "data layout: history records read (%d B) + solve_mv output written (%d B) per "
                                           "launch; cross-checked against the committed ncu capture" %
                                           (traffic["history"], traffic["output"]),
                         "traffic_ncu": tr_ncu, "kernel_ms": bwd_ms,
                         "kernel_ms_source": f"CUDA events on the launching stream, mean of {n_timed} timed steps",
                         "algorithmic_bytes_per_launch": ALG_BYTES_BACKWARD * B,
                         "algorithmic_frac": ALG_BYTES_BACKWARD * B / bwd_s / 1e9 / peak,
                         "algorithmic_note": "SURVEY 8(d) bytes (the reference's four dense history arrays + the output); "
                                             "the kernel moves fewer because predicted moments are recomputed, covariances "
                                             "are stored packed per block and every second record is re-run (checkpoint 2)",
                         "forward": {"kernel": "kalman_forward_wpb" if tpb else "kalman_forward",
                                     "kernel_ms": fwd_ms, "traffic": traffic["forward"],
                                     "achieved": traffic["forward"] / fwd_s / 1e9,
                                     "frac": traffic["forward"] / fwd_s / 1e9 / peak,
                                     "algorithmic_frac": ALG_BYTES_FORWARD * B / fwd_s / 1e9 / peak},
                         "whole_step": {"ms": ms / args.steps, "traffic": traffic["forward"] + tr,
                                        "frac": (traffic["forward"] + tr) / step_s / 1e9 / peak,
                                        "algorithmic_frac": ALG_BYTES_SOLVE * B / step_s / 1e9 / peak},
                         "fp64": {"peak_TFLOPs": f64, "source": "tools/microbench/fp64_peak.cu, measured in this run "
                                                                "(nominal 37.2)",
                                  "reference_flops_per_solve": REF_FLOPS_PER_SOLVE,
                                  "reference_equivalent_TFLOPs": value / world * REF_FLOPS_PER_SOLVE / 1e12,
                                  "note": "SURVEY 8(d): 3621 flop/step x 400 steps as the reference's dense "
                                          "KalmanTV executes them; the block-diagonal kernels skip the structural "
                                          "zeros, so this can exceed the peak -- the second bound, never the binding one"}},
            "kernel_family": family, "c5_sweep": c5, "dense_path": dense, "shared_covariance_path": shared,
            "e2e_other_entry_points": e2e_other,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "status_nonzero": status_bad,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
