#!/usr/bin/env python
"""bench.py -- Kalman logp+grad state-steps/s (batched series) on B200, next to the CPU oracle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl b200|reference]

One "step" = one batched evaluation of sum_t ll_t AND its gradient w.r.t. all nine state-space matrices
(workload C3: filter + smoother) over the whole synthetic batch of the workload, through the C ABI
(libbkf.so).  Default workload = BASELINE.json configs[1]: SARIMA(1,1,1)x(1,0,1,12), UnivariateFilter,
16384 draws x T=500 per GPU.  Prints ONE JSON line (rank 0).

  value   state-steps/s with inputs already resident in HBM (device pointers), CUDA-event timed, max over ranks,
          includes the final NCCL all-gather of the packed [logp | grads] buffer when N > 1 (weak scaling:
          every rank owns a full batch of draws).
  e2e     the same metric through the same C-ABI call with HOST (pinned) buffers: H2D of all inputs and D2H of
          logp + gradients inside the timed region.
  roofline / cpu_baseline: see DESIGN.md "Measurement".
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kalman_logp_grad_state_steps_per_sec"
UNIT = "state-steps/s"
NAMES = ("a0", "P0", "c", "d", "T", "Z", "R", "H", "Q")
STATIC_NDIM = {"a0": 1, "P0": 2, "c": 1, "d": 1, "T": 2, "Z": 2, "R": 2, "H": 2, "Q": 2}

WORKLOADS = {
    # name: (synth fn name, filter kind, default B per GPU, default T, mode)
    "C1": ("local_level", "standard", 1, 100, "logp_grad"),
    "C2": ("sarima", "univariate", 16384, 500, "logp_grad"),
    "C3": ("structural", "standard", 8192, 240, "filter_smooth"),
    "C4": ("varmax", "cholesky", 512, 1000, "logp_grad"),
    "C5": ("ets", "standard", 1_000_000, 100, "logp_grad"),
    # not a BASELINE config: C4's model under the reference's DEFAULT filter (StandardFilter), which has no specialised
    # kernel for p > 1 and runs on the generic CTA-per-series kernels
    "C4s": ("varmax", "standard", 512, 200, "logp_grad"),
}
DESCR = {
    "C1": "local-level structural (m=2,p=1,r=2) StandardFilter logp+grad",
    "C2": "SARIMA(1,1,1)x(1,0,1,12) (m=15,p=1,r=1) UnivariateFilter logp+grad",
    "C3": "structural trend+seasonal(12)+cycle (m=15,p=1,r=5) StandardFilter + KalmanSmoother",
    "C4": "VARMAX(2,0) k_endog=16 (m=32,p=16,r=16) SquareRootFilter logp+grad",
    "C5": "ETS(A,Ad,N) (m=3,p=1,r=1) StandardFilter logp+grad",
    "C4s": "VARMAX(2,0) k_endog=16 (m=32,p=16,r=16) StandardFilter logp+grad (generic kernels; not a BASELINE config)",
}


# ---- algorithmic cost model (dense, mul and add each count 1) -----------------------
def fwd_flops(kind, m, p, r):
    if kind == "standard":
        return 8 * m**3 + m**2 * (6 * p + 2 * r + 10) + m * (6 * p**2 + 2 * r**2 + 4 * p + 1) + p**3 + 3 * p**2 + 3 * p
    if kind == "univariate":
        return 4 * m**3 + m**2 * (5 * p + 2 * r + 8) + m * (7 * p + 2 * r**2 + 1) + 6 * p
    return (4 / 3 * (p + m) ** 3 + 2 * m**3 + 2 * m**2 * (r + 2 * m / 3) + 2 * p * m**2 + 2 * m**2
            + 2 * m * r**2 + (p**3 + r**3) / 3)


def smoother_flops(m, r, svd_pinv=False):
    """SURVEY 8d counts 36 m^3 per smoother step, 22 m^3 of which is the reference's SVD-based pinv.  The engine
    inverts the SPD P_hat directly, so the roofline is charged the conservative figure in which the pinv is a
    factor-and-solve (m^3/3 + 2 m^3); the SURVEY figure is reported beside it (svd_pinv=True)."""
    pinv = 22 * m**3 if svd_pinv else (m**3 / 3 + 2 * m**3)
    return 14 * m**3 + pinv + 2 * m**2 * r + 2 * m * r**2 + 14 * m**2


def step_model(kind, mode, m, p, r):
    """Algorithmic flops and HBM bytes per state-step, split per kernel class."""
    f = fwd_flops(kind, m, p, r)
    spill = (m + m * m) * 8
    if mode == "logp_grad":
        return {"forward": (f, 8 * p + spill), "backward": (2 * f, 8 * p + spill)}
    return {"forward": (f, 8 * p + spill), "smoother": (smoother_flops(m, r), 2 * spill)}


# ---- helpers -------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled DURING the timed region: NVML in a thread every 5 ms (a timed
    region is only ~100 ms long), `nvidia-smi -lms` as the fall-back when NVML is not importable."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []       # nvidia-smi rows
        self.samples = []    # (sm_mhz, reasons bitmask) from NVML
        self.proc = None
        self.nvml = None
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices; NVML indexes physical ones
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                why = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, why))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1)
            nv = self.nvml
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [x[0] for x in self.samples]
            reasons = sorted(k for k, bit in names.items() if any(x[1] & bit for x in self.samples))
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def ncu_traffic(name, kernel, B, T):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture of this
    exact workload size (profiles/ncu_traffic.json), else None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            ent = json.load(f).get(name)
    except (OSError, ValueError):
        return None
    if not ent or ent.get("B") != B or ent.get("T") != T or kernel not in ent:
        return None
    return {"bytes_per_launch": ent[kernel], "source": ent["source"]}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(name, B, T, seed_offset=0):
    from pymc_experimental_b200 import synth

    fn, kind, _, _, mode = WORKLOADS[name]
    cfg = getattr(synth, fn)(B, T) if seed_offset == 0 else getattr(synth, fn)(B, T, seed=synth.SEED_BASE + 100 + seed_offset)
    return cfg, kind, mode


# ---- CPU arm (oracle port; also the `--impl reference` arm) ----------------------------------------------
def cpu_engine(kind, mode):
    """Which oracle times the CPU arm: the compiled C/OpenMP port where it exists, else the torch / NumPy one."""
    if mode == "logp_grad" and kind in ("standard", "univariate"):
        return "c"
    return "torch" if mode == "logp_grad" else "numpy"


def cpu_run(name, kind, mode, cfg, sample_B):
    """One pass of the oracle over the first sample_B series of the workload; returns seconds."""
    import torch

    from oracle import kalman_np as onp
    from oracle import kalman_torch as oth

    sub = {k: cfg[k][:sample_B] for k in NAMES}
    y = cfg["y"] if cfg["y"].ndim == 2 else cfg["y"][:sample_B]
    if cpu_engine(kind, mode) == "c":
        from oracle import kalman_c

        kalman_c.lib()
        t0 = time.perf_counter()
        kalman_c.logp_and_grad(kind, y, *[sub[k] for k in NAMES])
        return time.perf_counter() - t0
    t0 = time.perf_counter()
    if mode == "logp_grad":
        oth.logp_and_grad(kind, y, *[sub[k] for k in NAMES])
    else:
        with torch.no_grad():
            for b in range(sample_B):
                outs = onp.kalman_filter(kind, y if y.ndim == 2 else y[b], *[sub[k][b] for k in NAMES])
                onp.kalman_smoother(sub["T"][b], sub["R"][b], sub["Q"][b], outs[0], outs[3])
    return time.perf_counter() - t0


def cpu_sample_size(name, mode):
    return {"C1": 1, "C2": 2048, "C3": 16, "C4": 16, "C5": 200000, "C4s": 16}[name] if mode == "logp_grad" else 16


def cpu_descr(kind, mode):
    eng = cpu_engine(kind, mode)
    if eng == "c":
        from oracle import kalman_c

        return kalman_c.num_threads(), "oracle/kalman_c.c (C port of the reference formulas + adjoint, gcc -O3 -fopenmp)"
    if eng == "torch":
        import torch

        return torch.get_num_threads(), "oracle/kalman_torch.py (batched torch.float64 autograd)"
    return 1, "oracle/kalman_np.py filter+smoother (NumPy/LAPACK, Python loop)"


def cpu_baseline(name, kind, mode, cfg, T, steps=1):
    import torch

    sB = min(cpu_sample_size(name, mode), cfg["a0"].shape[0])
    cpu_run(name, kind, mode, cfg, min(sB, 8))  # warm-up (library load, thread pool)
    best = min(cpu_run(name, kind, mode, cfg, sB) for _ in range(steps))
    cores, what = cpu_descr(kind, mode)
    return {
        "value": sB * T / best, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{sB} series x T={T} of the same workload through {what}",
        "seconds": best,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 in every worker; the CPU arm runs alone on rank 0 and may use the whole host
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or "OMP_NUM_THREADS" not in os.environ:
        os.environ["OMP_NUM_THREADS"] = os.environ.get("BKF_CPU_THREADS", str(os.cpu_count() or 1))
    name = args.workload or "C2"
    _, kind, dB, dT, mode = WORKLOADS[name]
    T = args.T or dT
    sB = cpu_sample_size(name, mode)
    cfg, kind, mode = make_workload(name, sB, T)
    for _ in range(args.warmup):
        cpu_run(name, kind, mode, cfg, sB)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_run(name, kind, mode, cfg, sB)
    dt = time.perf_counter() - t0
    cores, what = cpu_descr(kind, mode)
    value = sB * T * args.steps / dt
    sample = f"each step = {sB} series x T={T} of workload {name} through {what}; a restatement of the " \
             f"reference's PyTensor filter (PyTensor itself is not installable here)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}: {DESCR[name]}", "series_per_step": sB, "T": T},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---- GPU arm ------------------------------------------------------------------------------------------------
def run_b200(args):
    """Headline line = the workload asked for (default C2).  With no --workload the line also carries, under "also",
    one record each for C3, C4 and C5 at their BASELINE sizes (single GPU only), measured the same way."""
    import torch
    import torch.distributed as dist

    from pymc_experimental_b200.engine import KalmanEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    eng = KalmanEngine(local_rank)
    ctx = {"world": world, "rank": rank, "local_rank": local_rank, "dev": dev, "eng": eng}

    name = args.workload or "C2"
    line = measure_workload(args, ctx, name, args.batch, args.T, args.steps, with_cpu=world == 1 and not args.no_cpu)
    if rank == 0 and args.workload is None and world == 1 and not args.no_also:
        line["also"] = {}
    if args.workload is None and world == 1 and not args.no_also:
        for extra in ("C3", "C4", "C5", "C4s"):
            rec = measure_workload(args, ctx, extra, 0, 0, 2 if extra == "C4s" else max(2, min(args.steps, 3)), with_cpu=False)
            if rank == 0:
                line["also"][extra] = {k: rec[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "dtype", "config",
                                                           "clocks", "gpu_launches", "e2e", "roofline") if k in rec}
                for k in ("e2e_sample", "e2e_single_call"):
                    if k in rec:
                        line["also"][extra][k] = rec[k]
        # C4's BASELINE total (4 096 draws x T = 1 000) fits one GPU (35 GB of record): the same kernels with every SM holding
        # several series -- the per-GPU share above (512 draws) leaves one series per scheduler and is latency-bound
        rec = measure_workload(args, ctx, "C4", 4096, 0, 2, with_cpu=False)
        if rank == 0:
            line["also"]["C4"]["baseline_total_on_one_gpu"] = {
                k: rec[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "clocks", "e2e", "roofline")
                if k in rec}
    if world > 1 and not args.no_also and not args.batch:
        # strong scaling beside the (default) weak figure: the workload's BASELINE total split over the ranks
        _, _, dB, _, _ = WORKLOADS[name]
        total = {"C4": 4096}.get(name, dB)  # C4's BASELINE total is 4 096 draws (512 per GPU on 8 GPUs)
        rec = measure_workload(args, ctx, name, max(total // world, 1), args.T, args.steps, with_cpu=False)
        if rank == 0:
            line["strong_scaling"] = {"total_series": total, "series_per_gpu": max(total // world, 1), "value": rec["value"],
                                      "unit": rec["unit"], "ms_per_step": rec["ms_per_step"], "e2e": rec["e2e"],
                                      "note": "same workload, the BASELINE total split over the ranks (fixed total work)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def measure_workload(args, ctx, name, batch, T_arg, steps, with_cpu):
    import torch
    import torch.distributed as dist

    from pymc_experimental_b200.engine import F32, F64

    world, rank, local_rank, dev, eng = ctx["world"], ctx["rank"], ctx["local_rank"], ctx["dev"], ctx["eng"]
    _, kind, dB, dT, mode = WORKLOADS[name]
    B = batch or dB
    T = T_arg or dT
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    npdtype = np.float64 if args.dtype == "f64" else np.float32
    cfg, kind, mode = make_workload(name, B, T, seed_offset=rank)
    m, r = cfg["R"].shape[-2:]
    p = cfg["Z"].shape[-2]
    jitter = 1e-8 if args.dtype == "f64" else 1e-6

    # device-resident inputs
    dargs = [torch.as_tensor(cfg["y"], dtype=dtype, device=dev)] + [torch.as_tensor(cfg[k], dtype=dtype, device=dev) for k in NAMES]
    layout, total = eng.packed_layout(B, m, p, r)
    packed = torch.empty(total, dtype=dtype, device=dev)
    gathered = torch.empty(total * world, dtype=dtype, device=dev) if world > 1 else None
    # N > 1: the batch goes through the kernels in two halves, each with its own packed [logp | grads] buffer; the
    # all-gather of the first half (NCCL's own stream, async_op) runs under the kernels of the second half, so only half of
    # the one collective of the path is exposed.  (Series are independent: halves give bit-identical results.)
    halves = []
    # (only where a part still fills the GPU: a latency-bound workload -- C4: 512 series, one per scheduler -- takes as long
    #  for half the batch as for all of it, so it goes through in one piece; its gather is 8 MB)
    parts = int(os.environ.get("BKF_BENCH_PARTS", "2"))
    if world > 1 and mode == "logp_grad" and parts > 1 and B // parts >= 4096:
        bounds = [B * i // parts for i in range(parts + 1)]
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            _, tot_h = eng.packed_layout(hi - lo, m, p, r)
            sl = [dargs[0][lo:hi] if cfg["y"].ndim == 3 else dargs[0]]
            sl += [a[lo:hi] if a.ndim == STATIC_NDIM[k] + 1 else a for k, a in zip(NAMES, dargs[1:])]
            halves.append((sl, torch.empty(tot_h, dtype=dtype, device=dev), torch.empty(tot_h * world, dtype=dtype, device=dev)))

    def device_step():
        if mode == "logp_grad":
            if halves:  # the only inter-GPU traffic: one all-gather of [logp | grads] over NVLink, in two overlapped parts
                works = []
                for sl, pk, gt in halves:
                    eng.logp_grad(kind, *sl, packed_out=pk, cov_jitter=jitter)
                    works.append(dist.all_gather_into_tensor(gt, pk, async_op=True))
                for w in works:
                    w.wait()
                return None
            out = eng.logp_grad(kind, *dargs, packed_out=packed, cov_jitter=jitter)
            if world > 1:
                dist.all_gather_into_tensor(gathered, packed)
            return out
        return eng.filter_smooth(kind, *dargs, cov_jitter=jitter)

    # host (pinned) buffers for the end-to-end arm
    def pinned(a):
        t = torch.empty(a.shape, dtype=dtype, pin_memory=True)
        t.copy_(torch.as_tensor(a, dtype=dtype))
        return t.numpy()

    hargs = [pinned(cfg["y"])] + [pinned(cfg[k]) for k in NAMES]
    h2d = sum(a.nbytes for a in hargs)

    def host_step():
        if mode == "logp_grad":
            logp, grads, _ = eng.logp_grad(kind, *hargs, dtype=npdtype, cov_jitter=jitter)
            return logp.nbytes + sum(g.nbytes for g in grads.values())
        ll, sa, sP, _ = eng.filter_smooth(kind, *hargs, dtype=npdtype, cov_jitter=jitter)
        return ll.nbytes + sa.nbytes + sP.nbytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()

    # ---- timed region: device-resident ----
    fp64_peak = eng.measure_fma_peak(F64 if args.dtype == "f64" else F32)
    eng.enable_timing(True)
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        device_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = eng.launch_count - launches0
    ktimes = eng.kernel_times()
    eng.enable_timing(False)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    units_per_step = B * T * world
    value = units_per_step * steps / (ms * 1e-3)

    # ---- timed region: end to end through host buffers ----
    for _ in range(max(args.warmup, 3)):  # the same W >= 3 warm-up steps as the device-resident leg (first calls grow the
        host_step()                        # page-locked pool and fault in its pages)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(steps):
        d2h = host_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = units_per_step * steps / float(t_e.item())

    # ---- the same with the copies hidden: slices of the batch on two handles / streams / host threads ----
    e2e_overlapped = None
    if mode == "logp_grad" and args.dtype == "f64" and B >= 8192 and not os.environ.get("BKF_BENCH_NO_OVERLAP"):
        chunks = int(os.environ.get("BKF_BENCH_CHUNKS", "4"))

        def overlapped_step():
            logp, grads, _ = eng.logp_grad_overlapped(kind, *hargs, chunks=chunks, dtype=npdtype, cov_jitter=jitter)
            return logp.nbytes + sum(g.nbytes for g in grads.values())

        for _ in range(max(args.warmup, 3)):
            overlapped_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            o_d2h = overlapped_step()
        torch.cuda.synchronize()
        t_o = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_o, op=dist.ReduceOp.MAX)
        e2e_overlapped = {"value": units_per_step * steps / float(t_o.item()), "unit": UNIT, "chunks": chunks,
                          "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(o_d2h),
                          "note": "KalmanEngine.logp_grad_overlapped: the same dense call, host buffers in and out, the batch in "
                                  "slices on two handles / streams / host threads so that H2D and D2H run under the kernels"}

    # ---- the same through the compact parameterisation (row f-1): only the draw-dependent entries cross PCIe ----
    e2e_compact = None
    if name == "C2" and mode == "logp_grad":
        from pymc_experimental_b200 import synth

        base, positions, u = synth.sarima_compact(cfg)
        hu, hy = pinned(u), hargs[0]
        hbase = {k: np.ascontiguousarray(v, dtype=npdtype) for k, v in base.items()}

        def compact_step():
            lp, gu, _ = eng.logp_grad_compact(kind, hy, hbase, positions, hu, dtype=npdtype, cov_jitter=jitter)
            return lp.nbytes + gu.nbytes

        compact_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            c_d2h = compact_step()
        torch.cuda.synchronize()
        t_c = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_c, op=dist.ReduceOp.MAX)
        e2e_compact = {"value": units_per_step * steps / float(t_c.item()), "unit": UNIT,
                       "h2d_bytes_per_step": int(hu.nbytes + hy.nbytes + sum(v.nbytes for v in hbase.values())),
                       "d2h_bytes_per_step": int(c_d2h),
                       "note": "bkf_filter_logp_grad_compact: 7 draw-dependent matrix entries per draw in, logp + 7 "
                               "gradient entries out (SURVEY 8 f-1); extra figure, `e2e` above is the dense call"}

    # ---- filter + smoother + one draw per (draw, step), smoothed moments kept on the device (row f-3) ----
    e2e_sample = None
    if mode == "filter_smooth" and m <= 32:
        draw = [0]

        def sample_step():  # variates drawn in the kernel (Philox4x32-10, counter advanced per step): nothing but y and the
            draw[0] += 1    # parameters goes in, only the draws and loglike_obs come out
            x, ll, _ = eng.sample_conditional_posterior(kind, *hargs, dtype=npdtype, cov_jitter=jitter, seed=20261020,
                                                        offset=draw[0] * B * T * m)
            return x.nbytes + ll.nbytes

        sample_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            s_d2h = sample_step()
        torch.cuda.synchronize()
        t_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
        e2e_sample = {"value": units_per_step * steps / float(t_s.item()), "unit": UNIT,
                      "h2d_bytes_per_step": int(sum(a.nbytes for a in hargs)),
                      "d2h_bytes_per_step": int(s_d2h),
                      "note": "bkf_filter_smooth_sample: filter + smoother + one conditional-posterior draw per (draw, "
                              "step) through host buffers; the smoothed covariances stay on the device and the standard "
                              "normals are drawn in the kernel (SURVEY 8 f-3); extra figure, `e2e` above returns the "
                              "smoothed moments themselves"}

    # ---- roofline of the dominant kernel ----
    model = step_model(kind, mode, m, p, r)
    hbm_peak, peak_src = load_peaks()
    dom = max((k for k in model if ktimes[k][1] > 0), key=lambda k: ktimes[k][0])
    dom_ms = ktimes[dom][0] / ktimes[dom][1]
    flops, bytes_ = model[dom]
    if args.dtype == "f32":  # step_model counts float64 bytes
        model = {k: (f, b_ // 2) for k, (f, b_) in model.items()}
        bytes_ //= 2
    Bl = B / (len(halves) or 1)  # series per kernel launch (N > 1 runs the batch as two launches per kernel)
    ach_tf = flops * Bl * T / (dom_ms * 1e-3) / 1e12
    ach_gbs = bytes_ * Bl * T / (dom_ms * 1e-3) / 1e9
    t_fp = flops / (fp64_peak * 1e12)
    t_hbm = bytes_ / (hbm_peak * 1e9)
    if t_fp >= t_hbm:
        roof = {"bound": "fp64" if args.dtype == "f64" else "fp32", "achieved": ach_tf, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": ach_tf / fp64_peak, "traffic": None}
    else:
        roof = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "traffic": None}
    tr = ncu_traffic(name, dom, B, T) if args.dtype == "f64" else None  # the captures are of the float64 kernels
    if tr is not None:
        roof["traffic"] = tr["bytes_per_launch"]
        roof["traffic_source"] = tr["source"]
    roof.update({
        "kernel": dom, "kernel_ms": dom_ms, "kernel_share_of_step": ktimes[dom][0] / (ms if ms > 0 else 1.0),
        "algorithmic_flops_per_state_step": flops, "algorithmic_bytes_per_state_step": bytes_,
        "achieved_tflops": ach_tf, "fma_peak_tflops": fp64_peak, "fma_peak_source": "bkf_measure_fma_peak (live)",
        "achieved_hbm_gbs": ach_gbs, "hbm_peak_gbs": hbm_peak, "hbm_peak_source": peak_src,
        "all_kernels_ms": {k: (v[0] / v[1] if v[1] else None) for k, v in ktimes.items()},
    })
    # every kernel of the step against the same peak: the whole step (device-timed, collectives included) and each kernel
    step_flops = sum(f for f, _ in model.values()) * B * T
    roof["whole_step"] = {"achieved_tflops": step_flops / (ms / steps * 1e-3) / 1e12,
                          "frac_of_fma_peak": step_flops / (ms / steps * 1e-3) / 1e12 / fp64_peak}
    roof["per_kernel_frac_of_fma_peak"] = {
        k: (model[k][0] * Bl * T / (ktimes[k][0] / ktimes[k][1] * 1e-3) / 1e12 / fp64_peak) for k in model if ktimes[k][1] > 0}
    if dom == "smoother":
        roof["flops_note"] = ("smoother charged 14 m^3 + (7/3) m^3 (pinv as factor-and-solve) + O(m^2); SURVEY 8d's "
                              "figure with the reference's SVD pinv (22 m^3) is in achieved_tflops_survey_formula")
        roof["achieved_tflops_survey_formula"] = smoother_flops(m, r, True) * Bl * T / (dom_ms * 1e-3) / 1e12

    line = None
    if rank == 0:
        cpu = cpu_baseline(name, kind, mode, cfg, T) if with_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{name}: {DESCR[name]}", "series_per_gpu": B, "T": T, "m": int(m), "p": int(p),
                       "r": int(r), "mode": mode, "kernel_path": eng.last_path,
                       "l2": "working set (trajectory workspace + parameters) exceeds the 126 MB L2; no explicit flush",
                       "parallelism": f"independent series sharded over {world} GPU(s); one final NCCL all-gather of [logp | "
                                      "grads]" + (f", issued in {len(halves)} parts: each under the next part's kernels" if halves else "")},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "roofline": roof,
        }
        if e2e_compact is not None:
            line["e2e_compact"] = e2e_compact
        if e2e_overlapped is not None and world > 1:
            # several ranks share the host links and cores: the two-thread overlapped call measured slower than the single
            # call at N = 2 (686 against 776 M state-steps/s), so the single call stays the end-to-end figure there
            line["e2e_overlapped"] = e2e_overlapped
        elif e2e_overlapped is not None:
            # the headline end-to-end figure is the public call that hides the copies; the plain single call rides along
            line["e2e_single_call"] = dict(line["e2e"], note="KalmanEngine.logp_grad, one call: H2D, kernels, D2H back to back")
            line["e2e"] = {k: e2e_overlapped[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")}
            line["e2e"]["api"] = f"KalmanEngine.logp_grad_overlapped(chunks={e2e_overlapped['chunks']})"
            line["e2e"]["note"] = e2e_overlapped["note"]
        if e2e_sample is not None:
            line["e2e_sample"] = e2e_sample
        if cpu is not None:
            line["cpu_baseline"] = cpu
    del dargs, packed, gathered, hargs
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="one workload only (default: C2 as the headline plus C3, C4, C5 under \"also\")")
    ap.add_argument("--no-also", action="store_true", help="headline workload only")
    ap.add_argument("--batch", type=int, default=0, help="series per GPU (default: the workload's)")
    ap.add_argument("--T", type=int, default=0, help="time steps (default: the workload's)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
