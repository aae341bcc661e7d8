#!/usr/bin/env python
"""Benchmark of the hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Default workload = BASELINE.json config 3: 10-word discrete HMM (N=4, M=256) Baum-Welch over
100 000 synthetic sequences per word at T=200 (200 M frames per EM iteration), per GPU (weak
scaling: every rank holds its own shard of that size; the pi/A/B accumulators are combined
with one NCCL all-reduce per iteration).  One "step" = one full EM iteration (E-step forward +
backward/accumulate, reduce, all-reduce, M-step) over the whole batch.

Printed JSON (rank 0, one line): metric/value = Baum-Welch frames/s/iteration with the
codewords resident in HBM; `e2e` = the same through the public packed-array API
(engine.bw_fit -> hmmb_bw_*) from pinned HOST buffers, H2D/D2H inside the timed region;
`roofline` for the dominant kernel from CUDA events recorded around every launch on the
launching stream; `cpu_baseline` = the CPU oracle (numpy port of the reference) on a bounded
sample using all host cores; secondary blocks `vq_encode`, `lbg`, `score` for the other
BASELINE configs.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

WORKLOADS = {
    # name: (W, S per word per GPU, T, N, M)
    "bw_c3": dict(W=10, S=100_000, T=200, N=4, M=256, desc="config 3: 10 words x 100k seq x T=200, N=4, M=256"),
    "bw_c4": dict(W=1000, S=500, T=200, N=16, M=1024, desc="config 4: 1000 words x 500 seq x T=200, N=16, M=1024"),
    "bw_n8": dict(W=1000, S=500, T=200, N=8, M=1024, desc="1000 words x 500 seq x T=200, N=8, M=1024 (left-to-right kernels at 8 states)"),
    "bw_c1": dict(W=10, S=20, T=100, N=4, M=256, desc="config 1: 10 words x 20 seq x T~100 (latency-bound parity config)"),
}


def measured_traffic(workload: str, kernel: str, frames: int):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum per frame of the captured launch, scaled to this
    launch's frame count); None if no capture is recorded for the workload."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload][kernel]
        return rec["dram_bytes_per_frame"] * frames, rec.get("source")
    except Exception:
        return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (B200_PROFILING.md):
    NVML every ~2 ms from a thread (the timed region of the default workload is tens of
    milliseconds, shorter than one nvidia-smi call); nvidia-smi is the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []          # (sm_mhz, power_w, reasons bitmask)
        self.sm_max = None
        self.source = None
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
        except Exception:
            self._nvml = None
            self.source = "nvidia-smi"

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        try:
            pw = n.nvmlDeviceGetPowerUsage(self._h) / 1000.0
        except Exception:
            pw = None
        try:
            rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
        except Exception:
            try:
                rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            except Exception:
                rs = 0
        self.samples.append((sm, pw, rs))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 7:
            bits = 0
            for i, b in enumerate((0x8, 0x40, 0x20, 0x4)):  # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
                if parts[3 + i].lower().startswith("active"):
                    bits |= b
            self.sm_max = float(parts[1])
            self.samples.append((float(parts[0]), float(parts[2]), bits))

    def _run(self):
        try:  # first NVML query of a thread is slow: take (and drop) it before the timed region starts
            if self._nvml:
                self._sample_nvml()
                self.samples.clear()
        except Exception:
            pass
        self._ready.set()
        while not self._stop.is_set():
            try:
                if self._nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.002 if self._nvml else 0.2)

    def __enter__(self):
        self.samples = []
        self._stop.clear()
        self._ready = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        self._ready.wait(5.0)
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "source": self.source}
        sm = sorted(s[0] for s in self.samples)
        pw = [s[1] for s in self.samples if s[1] is not None]
        bits = 0
        for s in self.samples:
            bits |= s[2]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = [n for b, n in names.items() if bits & b]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "power_w_max": max(pw) if pw else None,
                "reasons": reasons, "samples": len(self.samples), "source": self.source}


class SmiLoopSampler:
    """Clocks for multi-rank runs: one `nvidia-smi -lms` child process on rank 0 (the recipe's own way).  A
    Python sampler thread cannot be used there: the all-reduce hook is a Python callback, so every EM iteration
    needs the GIL, and an NVML-polling thread made rank 0 the straggler of every all-reduce (measured on 8 GPUs:
    4.0-4.5 ms per iteration with the thread, 3.53 ms without)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.samples = []
        self.sm_max = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={ClockSampler.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7 and parts[0].replace(".", "").isdigit():
                bits = 0
                for i, b in enumerate((0x8, 0x40, 0x20, 0x4)):
                    if parts[3 + i].lower().startswith("active"):
                        bits |= b
                self.sm_max = float(parts[1])
                pw = float(parts[2]) if parts[2].replace(".", "").isdigit() else None
                self.samples.append((float(parts[0]), pw, bits))

    def summary(self):
        c = ClockSampler.__new__(ClockSampler)
        c.samples, c.sm_max, c.source = self.samples, self.sm_max, "nvidia-smi -lms 20 (rank 0, from the warm-up on)"
        return ClockSampler.summary(c)


# ----------------------------------------------------------------------------- CPU arm
def _cpu_train_word(args):
    from oracle import hmm_oracle as O
    seqs, N, M, iters, init = args
    t = time.perf_counter()
    O.hmm_training(seqs, N=N, M=M, epsilon=-1.0, max_iterations=iters, init=init)
    return time.perf_counter() - t


def cpu_baum_welch(cfg, seq_per_word: int, iters: int, procs: int):
    """The CPU oracle (numpy restatement of the reference's log-space Baum-Welch) on a bounded
    sample of the same workload, one process per word over `procs` host cores."""
    import multiprocessing as mp
    from hmm_training_b200 import engine, synthetic
    W, T, N, M = min(cfg["W"], max(procs, 1) * 2), cfg["T"], cfg["N"], cfg["M"]
    obs, offsets, wos = synthetic.fixed_length_codewords(12345, W, seq_per_word, T, N, M)
    init = engine.default_init(N, M)
    jobs = []
    for w in range(W):
        rows = obs.reshape(W * seq_per_word, T)[w * seq_per_word:(w + 1) * seq_per_word]
        jobs.append(([r.astype(np.int64) for r in rows], N, M, iters, init))
    t0 = time.perf_counter()
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            pool.map(_cpu_train_word, jobs)
    else:
        for j in jobs:
            _cpu_train_word(j)
    dt = time.perf_counter() - t0
    frames = W * seq_per_word * T
    return frames * iters / dt, dt, f"{W} words x {seq_per_word} seq x T={T} (N={N}, M={M}), {iters} EM iterations"


def run_reference(args, cfg, real_stdout):
    """--impl reference: the reference's algorithm on the host cores.  The reference itself is
    pure Python and only exists in the build container, so this arm times oracle/ (the
    validated numpy port, kind="port") with every host core, on the same workload shape."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    spw = int(os.environ.get("HMMB_CPU_SEQ_PER_WORD", "2000" if cfg["N"] <= 4 else "150"))
    for _ in range(max(args.warmup, 0) and 1):
        cpu_baum_welch(cfg, max(spw // 4, 1), 1, cores)
    vals = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, dt, sample = cpu_baum_welch(cfg, spw, 1, cores)
        vals.append(v)
    value = float(np.mean(vals))
    wall = time.perf_counter() - t_all
    line = {
        "impl": "reference", "metric": "baum_welch_frames_per_s_per_iter", "value": value, "unit": "frames/s/iter",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["desc"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "frames/s/iter", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s/iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(real_stdout, line)


# ----------------------------------------------------------------------------- GPU arm
def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    line to stdout on this image), so file descriptor 1 is pointed at stderr for the whole run and
    the JSON line goes to the saved original descriptor."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_fd: int, line: dict) -> None:
    sys.stdout.flush()
    os.write(real_fd, (json.dumps(line) + "\n").encode())


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bw_c3", choices=list(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the workload's sequences (debug)")
    ap.add_argument("--no-extras", action="store_true", help="skip VQ / LBG / scoring / CPU-baseline blocks")
    args = ap.parse_args()
    cfg = dict(WORKLOADS[args.workload])
    cfg["S"] = max(1, int(round(cfg["S"] * args.scale)))
    if args.impl == "reference":
        run_reference(args, cfg, real_stdout)
        return
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # torchrun exports OMP_NUM_THREADS=1; give every rank its share of the host cores for the library's
        # host-side blocking (hmmb_init reads HMMB_HOST_THREADS)
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        os.environ.setdefault("HMMB_HOST_THREADS", str(max(1, min(8, (os.cpu_count() or 1) // max(local_world, 1)))))
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from hmm_training_b200 import _lib, engine, synthetic
    from hmm_training_b200 import dist as hdist
    bound_cpus = hdist.bind_to_gpu_cpus(local_rank) if world > 1 else 0  # NUMA-local pinned buffers and host threads
    lib = _lib.load()
    _lib.init(local_rank)
    hdist.bind_torch_stream()
    hbm_peak, peak_src = measured_peaks()

    W, S, T, N, M = cfg["W"], cfg["S"], cfg["T"], cfg["N"], cfg["M"]
    obs, offsets, wos = synthetic.fixed_length_codewords(1000 + rank, W, S, T, N, M)
    frames_rank = int(offsets[-1])
    pi0, A0, B0 = engine.default_init(N, M)
    pi0, A0, B0 = np.tile(pi0, (W, 1)), np.tile(A0, (W, 1, 1)), np.tile(B0, (W, 1, 1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = lib.hmmb_launch_count()
    bw = engine.BaumWelch(obs, offsets, wos, W, N, M)
    bw.set_params(pi0, A0, B0)
    # default 1 = off: measured on config 4 at 2 GPUs, 4 / 8 groups cost more (7.77 / 7.80 ms per iteration) than the
    # plain all-reduce after the E-step (7.26 ms): per-group launches serialise the tails of the backward pass
    overlap_groups = int(os.environ.get("HMMB_OVERLAP_GROUPS", "1"))
    overlap = world > 1 and bw.kernel_family() == "left_to_right" and overlap_groups > 1
    if world > 1:
        # left-to-right path (config 4: 133 MB of accumulators): all-reduce per word group on a side stream,
        # overlapped with the next group's backward pass
        bw.set_dist(rank, world, hdist.make_allreduce(overlap=overlap))
        if overlap:
            bw.set_overlap(overlap_groups)
    cap = args.warmup + args.steps + 8
    smi = SmiLoopSampler(local_rank).start() if (world > 1 and rank == 0) else None
    if smi:
        time.sleep(1.0)  # nvidia-smi needs about a second to start printing
    bw.iterate(args.warmup, -1.0, cap, sync_each=False)
    _lib.check(lib.hmmb_set_profiling(1))
    _lib.check(lib.hmmb_phase_reset())
    barrier()
    l_before = lib.hmmb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # single GPU: NVML thread every ~2 ms; several ranks: an nvidia-smi child of rank 0 (see SmiLoopSampler)
    with (ClockSampler(local_rank) if world == 1 else contextlib.nullcontext()) as clk:
        ev0.record()
        bw.iterate(args.steps, -1.0, cap, sync_each=False)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    if smi:
        time.sleep(0.05)
        smi.stop()
        clk = smi
    gpu_launches = int(lib.hmmb_launch_count() - l_before)
    phases = {}
    for name in ("bw_forward", "bw_exact", "bw_backward", "bw_reduce", "bw_mstep"):
        pms, n = _lib.phase_ms(name)
        if n:
            phases[name] = {"ms_per_launch": pms / n, "launches": n}
    _lib.check(lib.hmmb_set_profiling(0))
    exact_passes, bwd_handover = bw.diagnostics()
    bw_family = bw.kernel_family()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    frames_total = frames_rank * world
    value = frames_total / (ms_per_step * 1e-3)
    bw.close()

    # ---- roofline of the dominant kernel (algorithmic bytes: DESIGN.md §Kernels)
    sym_b = 1 if M <= 256 else 2
    alg = {"bw_forward": frames_rank * (sym_b + 8 * N), "bw_backward": frames_rank * (sym_b + 8 * N)}
    dom = max((k for k in alg if k in phases), key=lambda k: phases[k]["ms_per_launch"], default=None)
    roofline = None
    if dom:
        ach = alg[dom] / (phases[dom]["ms_per_launch"] * 1e-3) / 1e9
        estep_ms = sum(phases[k]["ms_per_launch"] for k in ("bw_forward", "bw_backward") if k in phases)
        family = bw_family
        kname = {"n4_left_to_right": "k_bw_%s4<true>", "n4_dense": "k_bw_%s4<false>", "left_to_right": "k_bw_%sL",
                 "generic": "k_bw_%sG"}.get(family, "k_bw_%s") % ("bwd" if dom == "bw_backward" else "fwd")
        traffic, traffic_src = measured_traffic(args.workload, dom, frames_rank)
        roofline = {"bound": "hbm", "kernel": kname, "achieved": ach,
                    "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
                    "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg[dom],
                    "estep": {"algorithmic_bytes_per_frame": 2 * sym_b + 16 * N,
                              "achieved": frames_rank * (2 * sym_b + 16 * N) / (estep_ms * 1e-3) / 1e9,
                              "frac": frames_rank * (2 * sym_b + 16 * N) / (estep_ms * 1e-3) / 1e9 / hbm_peak},
                    "phases": phases}

    # ---- e2e: public packed-array API from pinned host buffers, one EM iteration per call
    e2e = None
    try:
        obs_p = torch.empty(obs.shape, dtype=torch.uint8 if obs.dtype == np.uint8 else torch.int16, pin_memory=True)
        obs_h = obs_p.numpy().view(obs.dtype)
        obs_h[:] = obs
        off_h = torch.from_numpy(offsets).pin_memory().numpy()
        wos_h = torch.from_numpy(wos).pin_memory().numpy()
        ar = hdist.make_allreduce() if world > 1 else None
        k_e2e = max(2, min(args.steps, 5))
        engine.bw_fit(obs_h, off_h, wos_h, W, N, M, pi0, A0, B0, -1.0, 1, allreduce=ar, rank=rank, world=world)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            out = engine.bw_fit(obs_h, off_h, wos_h, W, N, M, pi0, A0, B0, -1.0, 1, allreduce=ar, rank=rank, world=world)
        barrier()
        dt = (time.perf_counter() - t0) / k_e2e
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = obs_h.nbytes + off_h.nbytes + wos_h.nbytes + pi0.nbytes + A0.nbytes + B0.nbytes
        d2h = sum(x.nbytes for x in out)
        e2e = {"value": frames_total / dt, "unit": "frames/s/iter", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3, "steps": k_e2e,
               "api": "engine.bw_fit (hmmb_bw_create/set_params/iterate/get_params) with pinned host buffers, 1 EM iteration per call"}
    except Exception as exc:  # pragma: no cover
        e2e = {"error": repr(exc)}

    extras = {}
    if not args.no_extras and rank == 0:
        extras = run_extras(torch, lib, _lib, engine, synthetic, hbm_peak, world == 1)

    cpu_baseline = None
    if world == 1 and not args.no_extras:
        cores = os.cpu_count() or 1
        # bounded sample of the same workload: ~10 s of CPU work on all host cores
        spw = int(os.environ.get("HMMB_CPU_SEQ_PER_WORD", "6000" if N <= 4 else "500"))
        v, dt, sample = cpu_baum_welch(cfg, spw, 1, cores)
        cpu_baseline = {"value": v, "unit": "frames/s/iter", "cores": cores, "kind": "port", "sample": sample,
                        "seconds": dt}

    if rank == 0:
        line = {
            "metric": "baum_welch_frames_per_s_per_iter", "value": value, "unit": "frames/s/iter", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["desc"], "frames_per_gpu_per_iter": frames_rank, "W": W, "seq_per_word_per_gpu": S,
                       "T": T, "N": N, "M": M, "parallelism": f"sequences sharded over {world} GPU(s), 1 allreduce/iter" +
                       (f" in {overlap_groups} word groups overlapped with the backward pass" if overlap else ""),
                       "l2": "inputs larger than L2 (codewords + alpha spill >> 126 MB)",
                       "cpus_bound_per_rank": bound_cpus},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": gpu_launches,
            "clocks": clk.summary(), "precision_guard": {"exact_sequence_passes": exact_passes,
                                                         "backward_handovers": bwd_handover},
        }
        line.update(extras)
        _emit(real_stdout, line)
    # orderly teardown: free device memory and the library's events before NCCL goes away
    torch.cuda.synchronize()
    lib.hmmb_set_stream(None)
    lib.hmmb_shutdown()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_extras(torch, lib, _lib, engine, synthetic, hbm_peak, with_cpu):
    """Secondary blocks for BASELINE configs 2 (VQ / LBG, 1M frames) and 5 (recognition)."""
    out = {}
    F, K = 1_000_000, 256
    X = synthetic.mfcc_mixture(0, F, K)
    C = synthetic.random_codebook(1, K)
    dX = torch.from_numpy(X).cuda()
    dC = torch.from_numpy(C).cuda()
    dI = torch.empty(F, dtype=torch.int32, device="cuda")
    call = lambda: _lib.check(lib.hmmb_vq_encode_dev(dX.data_ptr(), F, dC.data_ptr(), K, dI.data_ptr(), None))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    ev0.record()
    for _ in range(reps):
        call()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    Xp = torch.from_numpy(X).pin_memory().numpy()
    engine.vq_encode(Xp, C)
    t0 = time.perf_counter()
    for _ in range(3):
        engine.vq_encode(Xp, C)
    dt = (time.perf_counter() - t0) / 3
    flops = F * K * 36.0
    out["vq_encode"] = {"metric": "vq_encode_frames_per_s", "value": F / (ms * 1e-3), "unit": "frames/s",
                        "frames": F, "K": K, "ms": ms,
                        "roofline": {"bound": "fp64", "achieved_tflops": flops / (ms * 1e-3) / 1e12,
                                     "hbm_achieved_gbs": F * 108 / (ms * 1e-3) / 1e9,
                                     "hbm_frac": F * 108 / (ms * 1e-3) / 1e9 / hbm_peak,
                                     "note": "12 DADD + 12 DFMA per frame-centroid pair; FP64-pipe bound (AI ~ 85 flop/B)"},
                        "e2e": {"value": F / dt, "unit": "frames/s", "h2d_bytes_per_step": int(X.nbytes + C.nbytes),
                                "d2h_bytes_per_step": 4 * F}}
    t0 = time.perf_counter()
    Cb, gens, assign, iters, gd = engine.lbg_fit(None, K, 100, 1e-3, x_dev_ptr=dX.data_ptr(), F=F)
    dt = time.perf_counter() - t0
    passes = int(iters.sum())
    out["lbg"] = {"metric": "lbg_codebook_build_s", "value": dt, "unit": "s", "frames": F, "K": K,
                  "lloyd_passes": passes, "iters_per_generation": [int(i) for i in iters],
                  "frame_passes_per_s": F * passes / dt}
    if with_cpu:
        from oracle import vq_oracle
        n = 20000
        vq_oracle.encode(X[:1000], C)
        t0 = time.perf_counter()
        vq_oracle.encode(X[:n], C)
        dtc = time.perf_counter() - t0
        out["vq_encode"]["cpu_baseline"] = {"value": n / dtc, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                            "sample": f"{n} frames x K={K} (C oracle, OpenMP)"}
    del dX, dI
    # recognition (BASELINE config 5): U utterances x 10 models, pinned host buffers, warm
    U, Wm = 1_000_000, 10
    rng = np.random.default_rng(5)
    obs, offsets, _ = synthetic.fixed_length_codewords(77, Wm, U // Wm, 100, 4, 256)
    pi, A, B = engine.default_init(4, 256)
    Bm = rng.dirichlet(np.ones(256) * 0.3, size=(Wm, 4))
    pim, Am = np.tile(pi, (Wm, 1)), np.tile(A, (Wm, 1, 1))
    obs_p = torch.empty(obs.shape, dtype=torch.uint8, pin_memory=True).numpy()
    obs_p[:] = obs
    ll_p = torch.empty((U, Wm), dtype=torch.float64, pin_memory=True).numpy()
    engine.score(obs_p, offsets, 4, 256, pim, Am, Bm, out_ll=ll_p)
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        ll, arg = engine.score(obs_p, offsets, 4, 256, pim, Am, Bm, out_ll=ll_p)
    dt = (time.perf_counter() - t0) / reps
    # kernel time from a separate pass with the library's per-launch CUDA events on (they cost ~0.4 ms per call)
    _lib.check(lib.hmmb_set_profiling(1))
    _lib.check(lib.hmmb_phase_reset())
    for _ in range(reps):
        engine.score(obs_p, offsets, 4, 256, pim, Am, Bm, out_ll=ll_p)
    kms, kn = _lib.phase_ms("score")
    _lib.check(lib.hmmb_set_profiling(0))
    t0 = time.perf_counter()
    for _ in range(reps):
        engine.score(obs_p, offsets, 4, 256, pim, Am, Bm, want_ll=False)
    dt_arg = (time.perf_counter() - t0) / reps
    kernel_ms = kms / reps  # all k_score4 launches of one call (the pipelined scorer launches it once per stage)
    fm = U * 100 * Wm  # frame x model pairs
    out["score"] = {"metric": "recognition_utterances_per_s", "value": U / dt,
                    "unit": "utterances/s (host API end to end, [U,W] log-likelihoods + argmax back on the host)",
                    "utterances": U, "models": Wm, "T": 100, "seconds": dt,
                    "argmax_only": {"value": U / dt_arg, "seconds": dt_arg},
                    "kernel": {"name": "k_score4", "ms": kernel_ms, "frame_models_per_s": fm / (kernel_ms * 1e-3),
                               "fp64_tflops": fm * 45.0 / (kernel_ms * 1e-3) / 1e12,
                               "note": "45 flop per frame x model (bidiagonal N=4); issue-bound, no HBM stream to speak of"},
                    "h2d_bytes_per_step": int(obs_p.nbytes + offsets.nbytes), "d2h_bytes_per_step": int(ll_p.nbytes + 4 * U)}
    if with_cpu:
        from oracle import hmm_oracle as O
        nu = 400
        seqs = [obs[u * 100:(u + 1) * 100].astype(np.int64) for u in range(nu)]
        models = [(Am[w], Bm[w], pim[w]) for w in range(Wm)]
        O.score_batch(seqs[:20], models)
        t0 = time.perf_counter()
        O.score_batch(seqs, models)
        dtc = time.perf_counter() - t0
        out["score"]["cpu_baseline"] = {"value": nu / dtc, "unit": "utterances/s", "cores": 1, "kind": "port",
                                        "sample": f"{nu} utterances x {Wm} models, T=100 (numpy oracle, 1 core)"}
    # MFCC front-end (parity unpinned): 100k frames of 320 samples
    Ym = np.random.default_rng(9).normal(size=(100_000, 320)) * 1000.0
    Ymp = torch.from_numpy(Ym).pin_memory().numpy()
    engine.mfcc_frames(Ymp)  # warm: device blocks of this size enter the library's caching allocator
    _lib.check(lib.hmmb_set_profiling(1))
    _lib.check(lib.hmmb_phase_reset())
    t0 = time.perf_counter()
    engine.mfcc_frames(Ymp)
    dtm = time.perf_counter() - t0
    mms, _ = _lib.phase_ms("mfcc")
    _lib.check(lib.hmmb_set_profiling(0))
    out["mfcc"] = {"metric": "mfcc_frames_per_s", "value": len(Ym) / dtm, "unit": "frames/s (host API end to end)",
                   "frames": len(Ym), "frame_len": 320, "kernel_ms": mms, "kernel_frames_per_s": len(Ym) / (mms * 1e-3),
                   "h2d_bytes_per_step": int(Ym.nbytes), "d2h_bytes_per_step": len(Ym) * 13 * 8,
                   "note": "parity with librosa unpinned (oracle/mfcc_oracle.py)"}
    # config 5 stress variant: 1000 left-to-right models with 16 states and 1024 codewords (k_scoreL)
    Us, Ws, Ns, Ms, Ts = 100_000, 1000, 16, 1024, 100
    obs, offsets, _ = synthetic.fixed_length_codewords(78, 10, Us // 10, Ts, Ns, Ms)
    pi, A, B = engine.default_init(Ns, Ms)
    Bs = np.empty((Ws, Ns, Ms))
    base = rng.dirichlet(np.ones(Ms) * 0.3, size=(8, Ns))
    for w in range(Ws):
        Bs[w] = np.roll(base[w % 8], 7 * w, axis=1)
    pis, As = np.tile(pi, (Ws, 1)), np.tile(A, (Ws, 1, 1))
    obs_p = torch.empty(obs.shape, dtype=torch.int16, pin_memory=True).numpy().view(np.uint16)
    obs_p[:] = obs
    engine.score(obs_p[:Ts * 1000], offsets[:1001], Ns, Ms, pis, As, Bs, want_ll=False)
    _lib.check(lib.hmmb_set_profiling(1))
    _lib.check(lib.hmmb_phase_reset())
    t0 = time.perf_counter()
    engine.score(obs_p, offsets, Ns, Ms, pis, As, Bs, want_ll=False)
    dt = time.perf_counter() - t0
    kms, kn = _lib.phase_ms("score")
    _lib.check(lib.hmmb_set_profiling(0))
    fm = Us * Ts * Ws
    out["score_n16"] = {"metric": "recognition_utterances_per_s", "value": Us / dt, "unit": "utterances/s (host API, argmax only)",
                        "utterances": Us, "models": Ws, "N": Ns, "M": Ms, "T": Ts, "seconds": dt,
                        "kernel": {"name": "k_scoreL<16>", "ms": kms / max(kn, 1),
                                   "frame_models_per_s": fm / (kms / max(kn, 1) * 1e-3)}}
    return out


if __name__ == "__main__":
    main()
