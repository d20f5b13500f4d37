#!/usr/bin/env python
"""Benchmark of the hot path: forward + adjoint SENSE NUFFT (complex64).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl b200|reference]

One *step* = one forward SENSE NUFFT (image -> k-space) followed by one adjoint SENSE
NUFFT (k-space -> coil-combined image) of the workload, through the public modules
(``KbNufft`` / ``KbNufftAdjoint``).  The metric is non-uniform points per second
counted in coil-points: ``2 * B * C * M`` per step and per GPU.  With N > 1 every
rank runs its own slices (batch sharding, no data-path collective): weak scaling.

Rank 0 prints ONE JSON line (see the keys below).  ``--impl reference`` times the
CPU oracle port of the reference's algorithm on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

if "--impl" in sys.argv and "reference" in sys.argv:
    # The CPU arm gets every host thread.  torchrun exports OMP_NUM_THREADS=1 to its workers, which makes libgomp
    # rebuild its thread team for every parallel region of the oracle (measured: 1.4x slower) -- undo that before
    # any OpenMP runtime is loaded.
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "NU points/sec (fwd+adj SENSE NUFFT, c64)"
UNIT = "coil-points/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(wl, B):
    """compulsory HBM traffic, every tensor touched once, each FFT one
    read + one write of the oversampled grid (complex64 = 8 B, omega float32)."""
    N = int(np.prod(wl.im_size))
    K = int(np.prod(wl.grid_size))
    C, M, d = wl.n_coils, wl.n_points, len(wl.im_size)
    fwd = 8 * (B * N + C * N + N + 4 * B * C * K + B * C * M) + 4 * d * M
    adj = 8 * (B * C * M + 3 * B * C * K + B * C * N + C * N + B * N + N) + 4 * d * M
    interp = 8 * (B * C * K + B * C * M) + 4 * d * M  # each direction, interpolation kernel alone
    return fwd, adj, interp


class ClockSampler:
    """SM clock / throttle-reason sampling DURING the timed region: an NVML polling thread
    (5 ms period; the main thread sits in cudaDeviceSynchronize with the GIL released), with
    nvidia-smi as a fallback when pynvml is unavailable."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._stop = False

    def _poll(self, nvml, handle):
        bits = {
            "hw_slowdown": getattr(nvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nvml, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        get_reasons = getattr(nvml, "nvmlDeviceGetCurrentClocksEventReasons",
                              getattr(nvml, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
        while not self._stop:
            try:
                self.samples.append(float(nvml.nvmlDeviceGetClockInfo(handle, nvml.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    mask = int(get_reasons(handle))
                    for name, bit in bits.items():
                        if mask & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading

            import pynvml as nvml

            nvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[self.gpu_index]) if visible and visible.split(",")[0].isdigit() \
                else self.gpu_index
            handle = nvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(nvml.nvmlDeviceGetMaxClockInfo(handle, nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, args=(nvml, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=statistics.median(self.samples), sm_max_mhz=self.sm_max,
                           reasons=sorted(self.reasons), samples=len(self.samples), source="nvml")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm),
                       source="nvidia-smi")
        return out


def make_config(wl, B, world, adjoint_mode, fft):
    """The `config` object of the JSON line: identical keys (and, for one workload, values) in both arms."""
    return {"workload": wl.name, "description": wl.description, "batch_per_gpu": B, "coils": wl.n_coils,
            "points": wl.n_points, "im_size": list(wl.im_size), "grid_size": list(wl.grid_size), "numpoints": 6,
            "adjoint_mode": adjoint_mode,
            "parallelism": f"batch-sharded x{world} (no collective)",
            "l2": "512 MiB buffer written between timed steps (L2 flush)",
            "plan": "trajectory plan cached across steps (built in warm-up)",
            "fft": fft}


LAUNCH_GRAPH = ("library-owned CUDA-graph replay (torchkbnufft_b200.set_graph_mode(True)): each operator call of the "
                "timed steps is one cudaGraphLaunch of its captured kernels")
LAUNCH_EAGER = "eager: every kernel launched by the host path of the module call"
FFT_OWN = "own pruned passes (compile-time plans, libb200nufft.so)"
FFT_CUFFT = "cuFFT via torch.fft + own pad/crop kernels"


def import_stock_reference():
    """The UNMODIFIED reference package installed by oracle/install_ref.sh under oracle/_ref (git-ignored; it
    travels to the GPU box with the snapshot).  Returns (module, None) or (None, reason)."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "torchkbnufft")):
        return None, "oracle/_ref/torchkbnufft missing (run `sh oracle/install_ref.sh` where /root/reference exists)"
    import importlib
    import warnings

    sys.path.insert(0, ref_dir)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod = importlib.import_module("torchkbnufft")
        return mod, None
    except Exception as exc:  # pragma: no cover
        return None, f"import failed: {exc!r}"
    finally:
        sys.path.remove(ref_dir)


def stock_reference_pair(wl, B, device, budget_s=25.0, warmup=1, max_steps=20):
    """Forward + adjoint SENSE NUFFT of the workload through the STOCK torchkbnufft modules (table mode, the call
    pattern of the reference's profile_torchkbnufft.py:98-139) on `device`: wall clock on the CPU, CUDA events with
    the L2 flushed between steps on a GPU.  Returns a dict for the JSON line."""
    import warnings

    import torch

    from torchkbnufft_b200 import workloads

    ref, why = import_stock_reference()
    if ref is None:
        return {"unavailable": why}
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0, n_batch=B)
    x, s, om = (torch.from_numpy(a).to(device) for a in (image, smaps, omega))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nu = ref.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(device)
        na = ref.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(device)
    cuda = torch.device(device).type == "cuda"
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=device) if cuda else None

    def pair():
        with torch.no_grad():
            k = nu(x, om, smaps=s)
            return na(k, om, smaps=s)

    times = []
    t_begin = time.perf_counter()
    for it in range(warmup + max_steps):
        if cuda:
            flush.fill_(it & 0xFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pair()
            e1.record()
            torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) * 1e-3
        else:
            t0 = time.perf_counter()
            pair()
            dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 3:
            break
    units = 2 * B * wl.n_coils * wl.n_points
    med = statistics.median(times)
    return {"value": units / med, "unit": UNIT, "ms_per_step": 1e3 * med, "steps": len(times), "warmup": warmup,
            "device": str(device), "threads": torch.get_num_threads() if not cuda else None,
            "package": f"torchkbnufft {getattr(ref, '__version__', '?')} (unmodified, oracle/_ref)",
            "note": "stock KbNufft / KbNufftAdjoint modules, table interpolation, complex64, same seeded inputs; "
                    "median over the timed steps"}


def oracle_pair_seconds(wl, B, steps, warmup, threads):
    """Time the CPU oracle (port of the reference's algorithm) on the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch

    import kbnufft_oracle as orc
    from torchkbnufft_b200 import workloads
    from torchkbnufft_b200._nufft import utils

    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0, n_batch=B)
    pre = utils.init_fn(im_size=wl.im_size, dtype=torch.complex64)
    scaling = utils.compute_scaling_coefs(pre.im_size.tolist(), pre.grid_size.tolist(), pre.numpoints.tolist(),
                                          pre.alpha.tolist(), pre.order.tolist()).to(torch.complex64).numpy()
    tables = [t.numpy() for t in pre.tables]
    J, L, ns = pre.numpoints.tolist(), pre.table_oversamp.tolist(), pre.n_shift.numpy()
    args = (omega, tables, ns, J, L, scaling, wl.im_size, wl.grid_size)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        k = orc.nufft_forward(image, *args, smaps=smaps, nthreads=threads)
        orc.nufft_adjoint(k, *args, smaps=smaps, nthreads=threads)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args, wl, rank):
    """--impl reference: the oracle port on all host cores (rank 0 only)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = wl.n_batch
    frac = 1.0
    sample = wl
    # bound the sample so the whole run ends within minutes (3-D / batched configs)
    est_units = 2 * B * wl.n_coils * wl.n_points * int(np.prod([6] * len(wl.im_size)))
    if est_units > 4e9:
        frac = 4e9 / est_units
        sample = wl.scaled(frac)
    times = oracle_pair_seconds(sample, B, args.steps, max(1, args.warmup), threads)
    units = 2 * B * sample.n_coils * sample.n_points
    total = sum(times)
    value = units * len(times) / total
    config = make_config(wl, B if args.gpus <= 1 or wl.n_batch == 1 else max(1, wl.n_batch // args.gpus),
                         max(1, args.gpus), "atomic", FFT_OWN)
    sample_txt = f"{sample.n_spokes}/{wl.n_spokes} spokes" if frac < 1 else "full workload"
    # the UNMODIFIED reference package beside the port (bounded to a few steps: it is an order of magnitude slower)
    stock = None
    try:
        import torch

        torch.set_num_threads(threads)
        stock = stock_reference_pair(sample, B, "cpu", budget_s=40.0, warmup=1, max_steps=5)
        if "value" in stock:
            stock.update(cores=threads, kind="reference", sample=sample_txt)
    except Exception as exc:  # pragma: no cover
        stock = {"unavailable": repr(exc)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"full {wl.name} fwd+adj pair x{len(times)}" if frac == 1.0 else
                         f"{sample.n_spokes} of {wl.n_spokes} spokes, fwd+adj pair x{len(times)}"},
        "cpu_baseline_reference": stock,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--adjoint-mode", default=None, choices=["atomic", "sorted"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip timing the stock reference on the GPU")
    ap.add_argument("--no-partitions", dest="partitions", action="store_false",
                    help="skip the cfg5 strong-scaling and coil-sharded sections (default workload only)")
    ap.add_argument("--breakdown", action="store_true", help="print per-stage device times to stderr")
    ap.add_argument("--launch", choices=["graph", "eager"], default="graph",
                    help="measured steps go through the library's CUDA-graph replay (tkbn.set_graph_mode) or eager launches")
    ap.add_argument("--opt", action="append", default=[], help="engine option id=value (A/B experiments)")
    ap.add_argument("--fft", choices=["auto", "cufft", "own"], default="auto",
                    help="FFT passes: auto (default; own compile-time planned passes when every grid length has a "
                         "plan, else cuFFT), cufft (always cuFFT + element-wise kernels), own (own passes for any "
                         "supported length)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    from torchkbnufft_b200 import workloads

    wl = workloads.WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank)
        return

    import torch
    import torch.distributed as dist

    import torchkbnufft_b200 as tkbn
    from torchkbnufft_b200 import _lib
    from torchkbnufft_b200._nufft import fft as eng_fft
    from torchkbnufft_b200._nufft import interp as eng_interp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU path); "
                         "use --impl reference for the CPU oracle timing")
    _lib.load()  # fail loudly if the native library is missing
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.adjoint_mode:
        tkbn.set_adjoint_mode(args.adjoint_mode)
    eng_fft.use_fused_fft = {"auto": "auto", "cufft": False, "own": True}[args.fft]
    for kv in args.opt:
        k, v = kv.split("=")
        _lib.check(_lib.load().b2n_set_option(int(k), int(v)), "b2n_set_option")

    B = wl.n_batch if world == 1 else max(1, wl.n_batch // world) if wl.n_batch > 1 else 1
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=rank, n_batch=B)
    omega = wl.trajectory(np.float32)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    x, s, y, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, kdata, omega))
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        k = nu(x, om, smaps=s)
        return k, na(k, om, smaps=s)

    # Launch mode of the measured steps: the library's own CUDA-graph replay (tkbn.set_graph_mode, public API: from the
    # fourth call with the same argument buffers an operator call is one cudaGraphLaunch) or plain eager launches.
    # Setup (plan build, twiddles, scratch, graph capture) happens in untimed calls BEFORE the W warm-up steps.
    from torchkbnufft_b200._nufft import graphs as eng_graphs
    use_graphs = args.launch == "graph"
    tkbn.set_graph_mode(use_graphs)
    for _ in range(3):
        step()
    torch.cuda.synchronize()  # the plan's device-side counts have been read back: graphs are captured in the next calls
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps, device time, L2 flushed between steps -----------------
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    def own_launches():  # kernels of the library: launched eagerly (counted at every launch) + replayed from graphs
        return int(_lib.load().b2n_launch_count()) + eng_graphs.replayed_kernel_count()

    launches_before = own_launches()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        starts[i].record()
        step()
        ends[i].record()
    gpu_launches = own_launches() - launches_before
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    step_ms = [a.elapsed_time(b) for a, b in zip(starts, ends)]
    total_ms = sum(step_ms)
    # the same K steps in the other launch mode (reported beside value as "other_launch_mode")
    tkbn.set_graph_mode(not use_graphs)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        starts[i].record()
        step()
        ends[i].record()
    torch.cuda.synchronize()
    eager_ms = sum(a.elapsed_time(b) for a, b in zip(starts, ends))
    if world > 1:
        t = torch.tensor([eager_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        eager_ms = float(t.item())
    tkbn.set_graph_mode(False)
    # The same K steps once more, back to back, now with CUDA events around the two interpolation launches (the
    # dominant kernels): their durations feed "roofline".  The four extra event records per step cost 11-15 us of
    # stream time (profiles/r01_h_pdl_ab.log), so this pass is kept out of "value" and reported beside it.
    eng_interp.kernel_timer = eng_interp.KernelTimer()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        starts[i].record()
        step()
        ends[i].record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    timer, eng_interp.kernel_timer = eng_interp.kernel_timer, None
    bracketed_ms = sum(a.elapsed_time(b) for a, b in zip(starts, ends)) / args.steps
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    units_per_step = 2 * B * wl.n_coils * wl.n_points
    value = world * units_per_step * args.steps / (total_ms * 1e-3)

    # ---- per-stage device times (same launches, separate pass) --------------------------------
    geo_args = (nu.tables, nu.n_shift, nu.numpoints, nu.table_oversamp)
    grid_size = tuple(wl.grid_size)
    ndim = len(grid_size)

    fused = eng_fft.fused_fft_available(torch.complex64, grid_size)

    def stage_times(reps):
        if fused:
            names = ["fft_fwd_fused", "interp_fwd", "interp_adj", "fft_adj_fused"]
        else:
            names = ["apod_pad", "fft", "interp_fwd", "interp_adj", "ifft", "crop_coilsum"]
        acc = {n: [] for n in names}
        for r in range(reps):
            flush.fill_(r & 0xFF)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            ev[0].record()
            if fused:
                g = eng_fft.fused_fft_forward(x, grid_size, s, nu.scaling_coef, 1.0)
                ev[1].record()
                k = eng_interp.table_interp(g, om, *geo_args)
                ev[2].record()
                g2 = eng_interp.table_interp_adjoint(k, om, *geo_args, None, nu.grid_size)
                ev[3].record()
                eng_fft.fused_fft_adjoint(g2, wl.im_size, s, nu.scaling_coef, 1.0)
                ev[4].record()
            else:
                g = eng_fft.apod_pad(x, grid_size, s, nu.scaling_coef, 1.0)
                ev[1].record()
                g = eng_fft.fft_grid(g, ndim, inverse=False)
                ev[2].record()
                k = eng_interp.table_interp(g, om, *geo_args)
                ev[3].record()
                g2 = eng_interp.table_interp_adjoint(k, om, *geo_args, None, nu.grid_size)
                ev[4].record()
                g2 = eng_fft.fft_grid(g2, ndim, inverse=True)
                ev[5].record()
                eng_fft.crop_apod_coilsum(g2, wl.im_size, s, nu.scaling_coef, 1.0)
                ev[6].record()
            torch.cuda.synchronize()
            for j, n in enumerate(names):
                acc[n].append(ev[j].elapsed_time(ev[j + 1]))
        return {n: statistics.median(v) for n, v in acc.items()}

    stages = stage_times(max(5, min(20, args.steps)))
    fwd_b, adj_b, interp_b = algorithmic_bytes(wl, B)
    peak, peak_src = load_peaks()
    live = {n: timer.mean_ms(n) for n in ("interp_fwd", "interp_adj")}  # measured inside the timed steps
    dom = "interp_adj" if live["interp_adj"] >= live["interp_fwd"] else "interp_fwd"
    achieved = interp_b / (live[dom] * 1e-3) / 1e9
    pair_achieved = (fwd_b + adj_b) / (total_ms / args.steps * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if wl.name == "cfg2" and B == 1 and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom)  # DRAM bytes of one launch from the committed ncu --set full capture

    # ---- end to end through the public API with HOST buffers ----------------------------------
    # Every step uploads that step's image, sensitivity maps, trajectory and measured k-space from pinned host
    # memory, runs forward + adjoint through the module API (the trajectory plan is rebuilt because the
    # trajectory tensor was rewritten) and reads both results back to pinned host memory.  "value" is the
    # steady-state throughput of a 3-deep software pipeline (upload stream / compute stream / download stream,
    # the way a reconstruction service feeds slices); "serial_value" is the same work on one stream with
    # nothing overlapped.  Both time all copies of all timed steps.
    e2e = None
    try:
        hx, hs, hy, hom = (torch.from_numpy(a).pin_memory() for a in (image, smaps, kdata, omega))
        NBUF = 3
        dbuf = [[torch.empty_like(t, device=dev) for t in (hx, hs, hy, hom)] for _ in range(NBUF)]
        hk = [torch.empty((B, wl.n_coils, wl.n_points), dtype=torch.complex64).pin_memory() for _ in range(NBUF)]
        hi = [torch.empty((B, 1) + tuple(wl.im_size), dtype=torch.complex64).pin_memory() for _ in range(NBUF)]
        compute = torch.cuda.current_stream(dev)
        up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ready = [torch.cuda.Event() for _ in range(NBUF)]
        freed = [torch.cuda.Event() for _ in range(NBUF)]
        drained = [torch.cuda.Event() for _ in range(NBUF)]

        def upload(b, stream):
            dx, ds, dy, dom_t = dbuf[b]
            with torch.cuda.stream(stream):
                dx.copy_(hx, non_blocking=True)
                ds.copy_(hs, non_blocking=True)
                dy.copy_(hy, non_blocking=True)
                dom_t.copy_(hom, non_blocking=True)  # new trajectory contents -> plan is rebuilt

        def run(b):
            dx, ds, dy, dom_t = dbuf[b]
            return nu(dx, dom_t, smaps=ds), na(dy, dom_t, smaps=ds)

        def download(b, k, im, stream):
            with torch.cuda.stream(stream):
                hk[b].copy_(k, non_blocking=True)
                hi[b].copy_(im, non_blocking=True)
            k.record_stream(stream)
            im.record_stream(stream)

        def serial_step():
            upload(0, compute)
            k, im = run(0)
            download(0, k, im, compute)

        def pipelined(n):
            for i in range(n):
                b = i % NBUF
                if i >= NBUF:
                    up.wait_event(freed[b])        # compute of step i-NBUF has consumed these inputs
                    down.wait_event(drained[b])    # and its results have left hk[b] / hi[b] (same stream: ordered)
                upload(b, up)
                ready[b].record(up)
                compute.wait_event(ready[b])
                k, im = run(b)
                freed[b].record(compute)
                down.wait_event(freed[b])
                download(b, k, im, down)
                drained[b].record(down)
            compute.wait_stream(up)
            compute.wait_stream(down)

        def timed(fn):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        # enough steps for the fill and drain of the 3-deep pipeline (one upload, one compute + download that overlap
        # nothing) to be a few per cent of the region, not 10 % as with 20 steps
        n_e2e = max(60, min(200, args.steps))
        for _ in range(3):
            serial_step()
        # median of five repetitions each (the host side of these copies is shared with whatever else runs on the
        # node; the best repetition is reported beside it)
        serial_all = [timed(lambda: [serial_step() for _ in range(n_e2e)]) for _ in range(5)]
        pipelined(NBUF)
        pipe_all = [timed(lambda: pipelined(n_e2e)) for _ in range(5)]
        serial_ms, pipe_ms = statistics.median(serial_all), statistics.median(pipe_all)
        h2d = sum(t.numel() * t.element_size() for t in (hx, hs, hom, hy))
        d2h = sum(t.numel() * t.element_size() for t in (hk[0], hi[0]))

        # second figure -- the reconstruction-service case: sensitivity maps and trajectory stay resident on the
        # device (their plan is cached), every step uploads only that step's image and k-space and downloads both
        # results
        def upload_data(b, stream):
            dx, _ds, dy, _dom = dbuf[b]
            with torch.cuda.stream(stream):
                dx.copy_(hx, non_blocking=True)
                dy.copy_(hy, non_blocking=True)

        def run_resident(b):
            dx, _ds, dy, _dom = dbuf[b]
            return nu(dx, om, smaps=s), na(dy, om, smaps=s)

        def pipelined_resident(n):
            for i in range(n):
                b = i % NBUF
                if i >= NBUF:
                    up.wait_event(freed[b])
                    down.wait_event(drained[b])
                upload_data(b, up)
                ready[b].record(up)
                compute.wait_event(ready[b])
                k, im = run_resident(b)
                freed[b].record(compute)
                down.wait_event(freed[b])
                download(b, k, im, down)
                drained[b].record(down)
            compute.wait_stream(up)
            compute.wait_stream(down)

        pipelined_resident(NBUF)
        res_ms = statistics.median([timed(lambda: pipelined_resident(n_e2e)) for _ in range(5)])
        h2d_res = sum(t.numel() * t.element_size() for t in (hx, hy))
        e2e = {"value": world * units_per_step * n_e2e / (pipe_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": pipe_ms / n_e2e,
               "best_ms_per_step": min(pipe_all) / n_e2e,
               "serial_value": world * units_per_step * n_e2e / (serial_ms * 1e-3),
               "serial_ms_per_step": serial_ms / n_e2e, "steps": n_e2e, "repetitions": 5,
               "resident_operator": {"value": world * units_per_step * n_e2e / (res_ms * 1e-3), "unit": UNIT,
                                     "ms_per_step": res_ms / n_e2e, "h2d_bytes_per_step": h2d_res,
                                     "d2h_bytes_per_step": d2h,
                                     "note": "sensitivity maps and trajectory resident on the device (plan cached); "
                                             "image and k-space uploaded, both results downloaded every step"},
               "note": "pinned host buffers; image, smaps, trajectory and k-space uploaded and both results "
                       "downloaded every step, trajectory plan rebuilt every step; value = 3-deep pipeline over "
                       "upload/compute/download streams, serial_value = one stream, no overlap; MEDIAN of 5 "
                       "repetitions of `steps` steps each"}
    except Exception as exc:  # pragma: no cover
        e2e = {"error": repr(exc)}

    # ---- the same step captured once in a CUDA graph and replayed (no per-launch host work, no launch gaps) -----
    graph_info = None
    try:
        step()  # the e2e section above rebuilt many plans: make sure this trajectory's plan is cached again, so
        step()  # that the captured graph holds the six kernels of the step and not a plan build
        torch.cuda.synchronize()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            step()
        torch.cuda.synchronize()
        n_graph = max(10, min(100, args.steps))
        g0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_graph)]
        g1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_graph)]
        for i in range(n_graph):
            flush.fill_(i & 0xFF)
            g0[i].record()
            graph.replay()
            g1[i].record()
        torch.cuda.synchronize()
        g_ms = sum(a.elapsed_time(b) for a, b in zip(g0, g1)) / n_graph
        if world > 1:
            t = torch.tensor([g_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            g_ms = float(t.item())
        graph_info = {"ms_per_step": g_ms, "value": world * units_per_step / (g_ms * 1e-3), "steps": n_graph,
                      "note": "same forward+adjoint step captured with torch.cuda.graph and replayed, L2 flushed between "
                              "replays; informational -- `value` above is the eager module API"}
        del graph
    except Exception as exc:  # pragma: no cover
        graph_info = {"error": repr(exc)}

    # ---- the north-star partitions, device-timed with the max over ranks (BASELINE.json config 5 / coil split) -----
    def timed_steps(fn, n_steps, n_warm=7):
        """Mean device ms per call of fn over n_steps, L2 flushed between steps, barrier + synchronize on both
        sides, MAX over ranks."""
        for i in range(n_warm):
            fn()
            if i == 2:
                torch.cuda.synchronize()  # plan counts read back before the calls that capture graphs (graph mode)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps)]
        for i in range(n_steps):
            flush.fill_(i & 0xFF)
            e0[i].record()
            fn()
            e1[i].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1)) / n_steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    partitions = {}
    if args.partitions and wl.name == "cfg2":
        from torchkbnufft_b200 import parallel

        n_part = max(5, min(20, args.steps))
        tkbn.set_graph_mode(use_graphs)  # as in the headline loop (timed_steps warms up before it times)
        # (a) BASELINE config 5, STRONG scaling: 64 slices x 16 coils, 64 / N slices per rank, shared trajectory,
        #     no collective on the data path (every slice is independent)
        try:
            w5 = workloads.WORKLOADS["cfg5"]
            lo, hi = parallel.shard_bounds(w5.n_batch, rank, world)
            im5, sm5, _kd5, om5 = workloads.make_inputs(w5, seed=100 + rank, n_batch=hi - lo)
            nu5 = tkbn.KbNufft(im_size=w5.im_size, dtype=torch.complex64).to(dev)
            na5 = tkbn.KbNufftAdjoint(im_size=w5.im_size, dtype=torch.complex64).to(dev)
            x5, s5, o5 = (torch.from_numpy(a).to(dev) for a in (im5, sm5, om5))

            def step5():
                k = nu5(x5, o5, smaps=s5)
                return na5(k, o5, smaps=s5)

            ms5 = timed_steps(step5, n_part)
            units5 = 2 * w5.n_batch * w5.n_coils * w5.n_points  # whole job: all 64 slices
            f5, a5, _ = algorithmic_bytes(w5, hi - lo)
            partitions["cfg5_strong"] = {
                "value": units5 / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5, "scaling": "strong",
                "slices_total": w5.n_batch, "slices_per_gpu": hi - lo, "coils": w5.n_coils, "steps": n_part,
                "collective": None, "pair_frac_per_gpu": (f5 + a5) / (ms5 * 1e-3) / 1e9 / peak,
                "note": "BASELINE config 5: forward + adjoint SENSE NUFFT of all 64 slices, batch-sharded; total "
                        "work fixed as N grows; device time, max over ranks"}
            del x5, s5, o5, nu5, na5
            tkbn.clear_caches()
            torch.cuda.empty_cache()
        except Exception as exc:  # pragma: no cover
            partitions["cfg5_strong"] = {"error": repr(exc)}
        # (b) coil sharding of cfg2 (B = 1 < N): C / N coils per rank, replicated image / trajectory; the forward
        #     needs no exchange, the adjoint ends with ONE NCCL sum all-reduce of the coil-combined image
        #     (reference coupling point: modules/kbnufft.py:404-405), inside the timed region
        try:
            lo, hi = parallel.shard_bounds(wl.n_coils, rank, world)
            im0, sm0, _kd0, _om0 = workloads.make_inputs(wl, seed=0, n_batch=1)  # the SAME problem on every rank
            x0, s0 = torch.from_numpy(im0).to(dev), torch.from_numpy(sm0).to(dev)
            s_loc = s0[:, lo:hi].contiguous()

            def step_sharded():
                k_loc = parallel.coil_sharded_forward(nu, x0, om, s_loc)
                return parallel.coil_sharded_adjoint(na, k_loc, om, s_loc)

            got = step_sharded()
            want = na(nu(x0, om, smaps=s0), om, smaps=s0)
            err = float(torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want))
            ms_sh = timed_steps(step_sharded, n_part)
            buf = torch.zeros_like(want)
            ms_ar = timed_steps(lambda: parallel.all_reduce_complex_(buf), n_part) if world > 1 else 0.0
            units2 = 2 * wl.n_coils * wl.n_points  # the whole 16-coil problem, whatever N is
            # the same partition with the engine's own all-reduce kernel over NVLink peer memory (csrc/b2n_peer.cu)
            peer_info = None
            if world > 1:
                try:
                    peer = parallel.PeerAllReduce(max_values=want.numel(), dtype=torch.complex64)

                    def step_peer():
                        k_loc = parallel.coil_sharded_forward(nu, x0, om, s_loc)
                        return parallel.coil_sharded_adjoint(na, k_loc, om, s_loc, reducer=peer)

                    got_p = step_peer().clone()
                    err_p = float(torch.linalg.vector_norm(got_p - want) / torch.linalg.vector_norm(want))
                    same = [torch.empty_like(got_p) for _ in range(world)]
                    dist.all_gather(same, got_p)
                    ms_p = timed_steps(step_peer, n_part)
                    parallel.fuse_allreduce = False  # A/B: the all-reduce as one more kernel behind the last pass
                    try:
                        ms_p_sep = timed_steps(step_peer, n_part)
                    finally:
                        parallel.fuse_allreduce = True
                    ms_par = timed_steps(lambda: peer(buf), n_part)
                    peer_info = {
                        "value": units2 / (ms_p * 1e-3), "unit": UNIT, "ms_per_step": ms_p,
                        "collective": "b2n_fft_adjoint_fused_allreduce: the adjoint's last inverse FFT pass pushes "
                                      "every finished image row into the peers' CUDA-IPC windows over NVLink and adds "
                                      "the arrivals in rank order (compute and collective in one kernel, replayed with "
                                      "the adjoint's graph); inside the timed region",
                        "ms_per_step_allreduce_as_separate_kernel": ms_p_sep,
                        "allreduce_ms_alone": ms_par, "rel_l2_vs_unsharded": err_p,
                        "bit_identical_on_all_ranks": bool(all(torch.equal(g, same[0]) for g in same))}
                    peer.close()
                except Exception as exc:  # pragma: no cover
                    peer_info = {"error": repr(exc)}
            nccl_info = {
                "value": units2 / (ms_sh * 1e-3), "unit": UNIT, "ms_per_step": ms_sh,
                "collective": "NCCL all-reduce(sum) of the coil-combined image, inside the timed region"
                              if world > 1 else None,
                "allreduce_ms_alone": ms_ar, "allreduce_share": (ms_ar / ms_sh) if ms_sh > 0 else None,
                "rel_l2_vs_unsharded": err}
            best = peer_info if (peer_info and "error" not in peer_info) else nccl_info
            partitions["coil_sharded"] = {
                "value": best["value"], "unit": UNIT, "ms_per_step": best["ms_per_step"], "scaling": "strong",
                "coils_total": wl.n_coils, "coils_per_gpu": hi - lo, "steps": n_part,
                "collective": best["collective"], "allreduce_bytes": int(want.numel() * 8),
                "allreduce_ms_alone": best["allreduce_ms_alone"],
                "allreduce_share": (best["allreduce_ms_alone"] / best["ms_per_step"]) if best["ms_per_step"] > 0 else None,
                "rel_l2_vs_unsharded": best["rel_l2_vs_unsharded"],
                "peer_memory_allreduce": peer_info, "nccl_allreduce": nccl_info,
                "note": "cfg2 forward + adjoint with the coils split over the ranks; strong scaling of ONE slice; "
                        "device time, max over ranks; value = the engine's peer-memory all-reduce kernel when the ranks "
                        "could map each other's windows, else the NCCL path (both reported)"}
            del x0, s0, s_loc, buf
        except Exception as exc:  # pragma: no cover
            partitions["coil_sharded"] = {"error": repr(exc)}
        tkbn.set_graph_mode(False)

    # ---- the stock reference package on the same GPU ("the existing Blackwell path") -----------------
    reference_cuda = None
    if rank == 0 and not args.no_reference_cuda:
        try:
            reference_cuda = stock_reference_pair(wl, B, dev, budget_s=20.0, warmup=5, max_steps=20)
            tkbn.clear_caches()
            torch.cuda.empty_cache()
        except Exception as exc:  # pragma: no cover
            reference_cuda = {"unavailable": repr(exc)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = wl
        est = 2 * B * wl.n_coils * wl.n_points * 6 ** ndim
        if est > 2e9:
            sample = wl.scaled(2e9 / est)
        times = oracle_pair_seconds(sample, B, 3, 1, threads)
        cpu_units = 2 * B * sample.n_coils * sample.n_points
        cpu_baseline = {"value": cpu_units / statistics.mean(times), "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": (f"full {wl.name} fwd+adj pair x3" if sample is wl else
                                   f"{sample.n_spokes} of {wl.n_spokes} spokes, fwd+adj pair x3")}

    if rank == 0:
        if args.breakdown:
            print("stage ms:", {k: round(v, 4) for k, v in stages.items()}, file=sys.stderr)
            print("step ms: min %.4f median %.4f max %.4f" % (min(step_ms), statistics.median(step_ms), max(step_ms)),
                  file=sys.stderr)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c64", "data": "synthetic",
            "config": make_config(wl, B, world, tkbn.get_adjoint_mode(), FFT_OWN if fused else FFT_CUFFT),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": live[dom], "kernels_ms_in_step": live,
                         "kernel_timing": "CUDA events around the two interpolation launches in a second pass of the "
                                          "same K steps, run back to back with the timed one (the brackets cost "
                                          "stream time, so they stay out of value)",
                         "ms_per_step_with_kernel_events": bracketed_ms,
                         "algorithmic_bytes": interp_b,
                         "pair_algorithmic_bytes": fwd_b + adj_b, "pair_achieved": pair_achieved,
                         "pair_frac": pair_achieved / peak},
            "stages_ms": {k: round(v, 5) for k, v in stages.items()},
            "launch": LAUNCH_GRAPH if use_graphs else LAUNCH_EAGER,
            "other_launch_mode": {"launch": LAUNCH_EAGER if use_graphs else LAUNCH_GRAPH,
                                  "ms_per_step": eager_ms / args.steps,
                                  "value": world * units_per_step * args.steps / (eager_ms * 1e-3), "unit": UNIT,
                                  "note": "the same K steps through the same module calls in the other launch mode "
                                          "(bench.py --launch)"},
            "cuda_graph": graph_info,
            "partitions": partitions,
            "reference_cuda": reference_cuda,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            # kernels of libb200nufft.so launched inside the timed region: counted by the library at every eager
            # launch (b2n_launch_count) and, for replayed graphs, as the launches recorded at capture x replays
            # (7 per step with the own FFT passes: rows, columns, gather / sample pre-pass, spread, columns,
            # rows + coil sum; cudaMemsetAsync, the L2-flush fill and cuFFT are not counted)
            "gpu_launches": gpu_launches,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
