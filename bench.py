#!/usr/bin/env python
"""bench.py -- throughput of the 802.11 channel-estimation hot path on B200 (BASELINE.json's metric).

    python bench.py --gpus 1 --steps 20 --warmup 5                # our arm (sm_100a kernels through the C-ABI)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # the reference's own CPU code on the host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank per GPU, frame-sharded, no collective

Headline workload (config.workload): BASELINE.json configs[2] -- shared-filter PS_MMSE over 1 Mi synthetic frames of the
inputs.h shape per GPU, FP32 I/O: one step = one pass  H[n][53] = (rx/tx)[n][53] W^T  over every local frame (LS divide
fused into the GEMM kernel).  `value` = frames all ranks processed / max-over-ranks device time (inputs resident in HBM);
`e2e` = the same call with HOST buffers (pinned), H2D and D2H inside the timed region.  Inputs (1.27 GB per pass) are
10x larger than L2, so no flush is needed between iterations.  `extras` carries the secondary kernels (LT_LS, fused
PS_Linear/Cubic/Sinc, equalizer, per-frame solves), each with its own roofline fraction.
"""
import argparse
import ctypes
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

NSC, NBLK = 53, 15
METRIC = "MMSE channel estimates/sec (53-subcarrier frames)"
UNIT = "frames/s"
OW2 = 9.6172e-08
AMP = 8.875
# DRAM bytes of ONE mmse_shared_tc launch over 1 Mi frames from the ncu --set full capture (0.8896 GB read + 0.3875 GB written, profiles/r02_ncu_kernels.txt)
NCU_DRAM_BYTES_PER_MI_FRAMES = 1.2772e9


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        self.join(timeout=1)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def time_steps(fn, steps, warmup, torch, dist=None):
    """W untimed + K timed steps; device time from CUDA events per step; returns (total_ms, per_step_ms list)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return ev[0].elapsed_time(ev[steps]), per


class numa_local:
    """Allocate pinned host buffers on the NUMA node next to the GPU: the thread is bound to the GPU's CPU set (NVML) while
    cudaHostAlloc places and pins the pages, then the previous affinity is restored (the CPU baseline keeps every core).
    With 8 ranks copying at once, buffers on the far socket put every byte on the inter-socket link."""

    def __init__(self, index):
        self.index, self.prev, self.state = index, None, "unbound"

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * i + b for i, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
            prev = os.sched_getaffinity(0)
            cpus &= prev
            if cpus and cpus != prev:
                os.sched_setaffinity(0, cpus)
                self.prev, self.state = prev, "bound to %d of %d cpus" % (len(cpus), len(prev))
            elif cpus:
                self.state = "single node"
        except Exception as e:          # no NVML / no permission: allocate wherever the process runs
            self.state = "unbound (%s)" % type(e).__name__
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            try:
                os.sched_setaffinity(0, self.prev)
            except Exception:
                pass
        return False


def pinned(wifi, shape, dtype):
    lib = wifi._lib.load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    if lib.wifi_host_alloc(ctypes.byref(p), n) != 0:
        raise MemoryError("wifi_host_alloc(%d)" % n)
    buf = (ctypes.c_char * n).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def cpu_reference_rate(sample_frames, seconds_target=12.0):
    """The reference's own routines (oracle/_ref: multiply utils.c:16-31 + the LS divide), frame-parallel over all host
    threads, on a bounded sample of the same workload.  Falls back to the oracle port when _ref is not built."""
    import synth
    from oracle.pyoracle import Oracle, Reference
    o = Oracle()
    cores = os.cpu_count() or 1
    R = synth.channel_covariance()
    d = np.full(NSC, OW2 / AMP ** 2); d[26] = OW2 / 1e-8
    W = o.mmse_filter(R, d)
    fr = synth.make_frames(4096, seed=5)
    tx, rx = fr["tx_symb"][:, 0, :].copy(), fr["rx_symb"][:, 0, :].copy()
    if Reference.available():
        ref = Reference()
        cores = ref.set_threads(cores)           # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
        run = lambda t, r: ref.mmse_shared_omp(W, t, r)
        kind = "reference"
    else:
        run = lambda t, r: o.mmse_apply(W, r / t)
        kind, cores = "port", 1
    t0 = time.perf_counter(); run(tx, rx); dt = time.perf_counter() - t0
    n = int(min(sample_frames, max(4096, 4096 * seconds_target / max(dt, 1e-6))))
    reps = -(-n // 4096)
    txb, rxb = np.tile(tx, (reps, 1))[:n], np.tile(rx, (reps, 1))[:n]
    t0 = time.perf_counter(); run(txb, rxb); dt = time.perf_counter() - t0
    seq = None
    if kind == "reference":                       # the same routine on ONE core (the reference's sequential build), ~2 s
        ref.set_threads(1)
        m = int(max(4096, min(n, 2.0 * (n / dt) / max(cores, 1))))
        t1 = time.perf_counter(); run(txb[:m], rxb[:m]); d1 = time.perf_counter() - t1
        ref.set_threads(cores)
        seq = {"value": m / d1, "unit": UNIT, "cores": 1, "sample": "%d frames, %.1f s" % (m, d1)}
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind, "sequential": seq,
            "not_run": "the reference's MPI build (no mpirun/mpi.h on this box; main_mpi.c:1015-1080 numbers are in BASELINE.md) and its "
                       "intra-frame OpenMP PS_MMSE (returns NaN after 278 s per frame, SURVEY 6.2)",
            "sample": "%d frames of the shared-filter MMSE workload (LS divide + 53x53 filter), %.1f s, %s" %
                      (n, dt, "oracle/_ref ref_mmse_shared_omp (reference multiply(), OpenMP over frames)" if kind == "reference"
                       else "oracle port, single thread")}, (run, txb, rxb)


def cpu_estimator_baselines(seconds_each=1.5):
    """SURVEY 8(d) CPU baselines for the four LS / interpolation estimators, all from the reference's own code compiled in
    place (oracle/_ref): (i) the sequential functions on ONE core at -O2 and at -O0 (what the reference's compile.c builds),
    (ii) the reference's OpenMP build as it is (main_openmp.c: a 53-thread intra-frame team per call) on a small N',
    (iii) the frame-parallel harness (ref_estimate_omp: `omp parallel for` over frames around the sequential functions) on all
    host threads.  Bounded: ~1.5 s per timing."""
    import synth
    from oracle.pyoracle import Reference, ReferenceOpenMP
    if not Reference.available():
        return {"unavailable": "oracle/_ref not built on this box"}
    ref = Reference()
    ref0 = Reference("O0") if Reference.available("O0") else None
    romp = ReferenceOpenMP() if ReferenceOpenMP.available() else None
    cores = ref.set_threads(os.cpu_count() or 1)
    fr = synth.make_frames(4096, seed=6)
    tx, rx = fr["tx_symb"][:, 0, :].copy(), fr["rx_symb"][:, 0, :].copy()
    out = {"cores": cores, "unit": UNIT,
           "not_run": "the MPI build (no mpirun / mpi.h on this box; its published numbers, main_mpi.c:1015-1080, are in BASELINE.md)"}

    def timed(fn, a, b, per_frame_guess):
        n = int(max(64, min(1 << 20, seconds_each / per_frame_guess)))
        reps = -(-n // len(a))
        A, B = np.tile(a, (reps, 1))[:n], np.tile(b, (reps, 1))[:n]
        t0 = time.perf_counter(); fn(A, B); dt = time.perf_counter() - t0
        return {"value": n / dt, "sample": "%d frames, %.2f s" % (n, dt)}

    for w in ("lt_ls", "ps_linear", "ps_cubic", "ps_sinc"):
        a, b = (fr["tx_pre"], fr["rx_pre"]) if w == "lt_ls" else (tx, rx)
        t0 = time.perf_counter(); ref.estimate(w, a, b); per = (time.perf_counter() - t0) / len(a)      # ~ -O2, one core
        e = {"sequential_O2_1core": timed(lambda x, y: ref.estimate(w, x, y), a, b, per)}
        if ref0 is not None:
            e["sequential_O0_1core"] = timed(lambda x, y: ref0.estimate(w, x, y), a, b, 2 * per)
        e["frame_parallel_omp_all_cores"] = dict(timed(lambda x, y: ref.estimate(w, x, y, omp=True), a, b, per / max(cores, 1) * 1.5), cores=cores)
        if romp is not None:
            t0 = time.perf_counter(); romp.estimate(w, a[:16], b[:16]); per_o = (time.perf_counter() - t0) / 16
            e["reference_openmp_build_as_is"] = dict(timed(lambda x, y: romp.estimate(w, x, y), a, b, per_o),
                                                     note="main_openmp.c: num_threads(53) team per call, frames one after the other")
        out[w] = e
    # configs[3]: the per-frame 53 x 53 solve.  The reference's own PS_MMSE returns NaN after 278 s per frame, so this
    # row is the oracle's long-double LU restatement of WiFi_channel_estimation_PS_MMSE.m:16-33 on ONE core -- a port, stated as such
    try:
        from oracle.pyoracle import Oracle
        o = Oracle()
        frp = synth.make_frames(256, seed=7, sigma2="perframe")
        txp, rxp, R = frp["tx_symb"][:, 0, :].copy(), frp["rx_symb"][:, 0, :].copy(), synth.channel_covariance()
        t0 = time.perf_counter(); o.mmse_perframe(R, txp[:16], rxp[:16], frp["sigma2"][:16]); per = (time.perf_counter() - t0) / 16
        n = int(max(16, min(256, seconds_each / per)))
        t0 = time.perf_counter(); o.mmse_perframe(R, txp[:n], rxp[:n], frp["sigma2"][:n]); dt = time.perf_counter() - t0
        out["mmse_perframe"] = {"oracle_port_long_double_1core": {"value": n / dt, "sample": "%d frames, %.2f s" % (n, dt), "kind": "port", "cores": 1,
                                                                  "note": "53 x 53 complex LU + back-substitution per frame in long double "
                                                                          "(oracle/wifi_oracle.c); the reference's PS_MMSE: 278 s per frame, NaN"}}
    except Exception as exc:                                          # the checker library is test infrastructure: never fail the bench on it
        out["mmse_perframe"] = {"unavailable": str(exc)[:120]}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, (run, txb, rxb) = cpu_reference_rate(1 << 22, seconds_target=max(1.0, min(12.0, 90.0 / (args.steps + args.warmup))))
    n = len(txb)
    for _ in range(args.warmup):
        run(txb[: n // 4], rxb[: n // 4])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(txb, rxb)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    cb["value"] = val
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f80",
        "data": "synthetic", "config": {"workload": "configs[2]: shared-filter PS_MMSE, bounded sample of %d frames per step on the host cores" % n},
        "cpu_baseline": cb, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


class Extras:
    """Secondary kernels: every rank times the same list on its own shard, the per-entry device time is the MAX over ranks
    (one all-reduce of the whole list after the last entry) and rank 0 turns it into whole-job frames/s and roofline fractions
    (algorithmic bytes / flops per frame from SURVEY 8(d))."""

    def __init__(self, torch, dist, world, hbm, steps, warmup):
        self.torch, self.dist, self.world, self.hbm, self.steps, self.warmup = torch, dist, world, hbm, steps, warmup
        self.entries = []

    def rate(self, name, fn, n, bytes_per_frame=None, flops_per_frame=None, peak_tflops=None, measured_tflops=None, **meta):
        _, per = time_steps(fn, self.steps, self.warmup, self.torch, self.dist)
        self.entries.append(dict(name=name, ms=float(np.median(per)), n=n, bpf=bytes_per_frame, fpf=flops_per_frame, nom=peak_tflops,
                                 meas=measured_tflops, meta=meta))

    def finish(self):
        torch = self.torch
        t = torch.tensor([e["ms"] for e in self.entries], dtype=torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        out = {}
        for e, ms in zip(self.entries, t.tolist()):
            sec, ntot = ms * 1e-3, e["n"] * self.world
            r = {"frames_per_s": ntot / sec, "ms": ms, "n_frames": ntot, "n_frames_per_gpu": e["n"]}
            if e["bpf"]:
                r["GBps_per_gpu"] = e["n"] * e["bpf"] / sec / 1e9
                r["hbm_frac"] = r["GBps_per_gpu"] / self.hbm
                r["bytes_per_frame"] = e["bpf"]
            if e["fpf"]:
                r["TFLOPs_per_gpu"] = e["n"] * e["fpf"] / sec / 1e12
                r["flops_per_frame"] = e["fpf"]
                if e["nom"]:
                    r["peak_tflops_nominal"] = e["nom"]
                    r["flop_frac_of_nominal"] = r["TFLOPs_per_gpu"] / e["nom"]
                if e["meas"]:
                    r["peak_tflops_measured"] = e["meas"]
                    r["flop_frac_of_measured"] = r["TFLOPs_per_gpu"] / e["meas"]
            r.update(e["meta"])
            out[e["name"]] = r
        return out


def extras_block(wifi, ctx, torch, dist, world, shard_lo, peaks, mp, n_frames, steps, warmup, full):
    """configs[1] (LT_LS + PS_Linear/Cubic/Sinc), configs[3] (per-frame solve, 256 Ki frames per GPU, FP32 and FP64 storage),
    configs[4] (all five estimators + the equalizer on whole frames) on EVERY rank's shard; `full` adds the remaining kernels
    (front-end, eigen-domain, closed form, pivoted solve, batched utils)."""
    X = Extras(torch, dist, world, peaks["hbm_gbs"], steps, warmup)
    for prec, cbytes in (("f32", 8), ("f64", 16)):
        n = n_frames
        fr = ctx.synth_frames(n, prec, first_frame=shard_lo, want=("tx_pre", "rx_pre", "tx_symb", "rx_symb"))
        H = torch.empty_like(fr["tx_pre"])
        X.rate("lt_ls_" + prec, lambda: ctx.lt_ls(fr["tx_pre"], fr["rx_pre"], out=H), n, 159 * cbytes, config="configs[1]")
        outs = {k: torch.empty_like(fr["tx_pre"]) for k in ("linear", "cubic", "sinc")}
        X.rate("ps_fused3_" + prec, lambda: ctx.ps(fr["tx_symb"], fr["rx_symb"], out=outs), n, 167 * cbytes, config="configs[1]")
        o1 = {"linear": outs["linear"]}
        X.rate("ps_linear_" + prec, lambda: ctx.ps(fr["tx_symb"], fr["rx_symb"], ("linear",), out=o1), n, 61 * cbytes, config="configs[1]",
               note="one estimator alone: 8 isolated pilot values per frame cost one 64-byte DRAM atom each (512 B against 64 B algorithmic)")
        eq = torch.empty_like(fr["rx_symb"])
        X.rate("equalize_" + prec, lambda: ctx.equalize(fr["rx_symb"], H, outs["linear"], out=eq), n, 1696 * cbytes, config="configs[4]")
        # BASELINE configs[4] per GPU: all five estimators + the equalizer on whole frames, in place (block 0 at stride 795)
        Hm5 = torch.empty_like(H)
        o5 = {"lt_ls": H, "linear": outs["linear"], "cubic": outs["cubic"], "sinc": outs["sinc"], "mmse": Hm5, "eq": eq}
        X.rate("all5_plus_equalizer_" + prec, lambda: ctx.estimate_all(fr["tx_pre"], fr["rx_pre"], fr["tx_symb"], fr["rx_symb"], out=o5), n,
               2014 * cbytes, config="configs[4]",
               note="LT_LS + PS_Linear/Cubic/Sinc + shared-filter PS_MMSE + equalizer on whole frames through wifi_estimate_all_batch: 4 launches, the "
                    "GEMM kernel hands the four pilot LS values of every frame to the interpolators (no pilot gather). hbm_frac is on the algorithmic "
                    "bytes of a single fused pass (read tx_pre, rx_pre, block 0 of tx, rx_symb = 954 c, write five estimates + eq = 1 060 c); the four "
                    "kernels move 2 181 c (the equalizer re-reads H_lt, H_linear; block 0 of rx is read twice)")
        del eq, Hm5, o5
        tx0 = fr["tx_symb"][:, 0, :].contiguous(); rx0 = fr["rx_symb"][:, 0, :].contiguous()
        del fr
        Hm = torch.empty_like(tx0)
        if prec == "f64":
            X.rate("mmse_shared_f64", lambda: ctx.mmse_shared(tx0, rx0, out=Hm), n, 159 * cbytes, 22472, 37.2, mp["fp64_dmma_tflops"], config="configs[2]")
        # per-frame solve: 256 Ki frames per GPU (configs[3]).  complex64 storage runs the solve in FP64 arithmetic (the mode
        # that meets the 1e-4 bound), so both storage types are measured against the FP64 peak.
        npf = min(n, 1 << 18)
        s2 = ctx.synth_frames(npf, prec, first_frame=shard_lo, per_frame_sigma=True, want=("sigma2",))["sigma2"]
        R = ctx.synth_covariance()
        Rp = R if prec == "f64" else R.to(torch.complex64)
        Hp = torch.empty_like(tx0[:npf])
        X.rate("mmse_perframe_hpd_" + prec, lambda: ctx.mmse_perframe(Rp, tx0[:npf], rx0[:npf], s2, flags=wifi.SOLVE_HPD, out=Hp), npf,
               159 * cbytes, 441949, 37.2, mp["fp64_dmma_tflops"], config="configs[3]",
               arithmetic="FP64 (blocked L D L^H, trailing updates on DMMA)" + ("; FP32 storage" if prec == "f32" else ""))
        if prec == "f32" and full:
            X.rate("mmse_perframe_hpd_f32_fast32", lambda: ctx.mmse_perframe(Rp, tx0[:npf], rx0[:npf], s2, flags=wifi.SOLVE_HPD | wifi.SOLVE_FAST32, out=Hp),
                   npf, 159 * cbytes, 441949, 74.4, mp["fp32_fma_tflops"], config="configs[3]",
                   note="WIFI_SOLVE_FAST32 opt-in: FP32 arithmetic, documented accuracy 4e-3 -- NOT a parity mode (north star: 1e-4)")
        # eigen-domain per-frame MMSE: the synthetic frames are BPSK, so |tx_k|^2 is shared
        absx2 = (tx0[0].abs().to(torch.float64)) ** 2
        ctx.mmse_eig_prepare(R, absx2)
        He = torch.empty_like(tx0)
        s2n = ctx.synth_frames(n, prec, first_frame=shard_lo, per_frame_sigma=True, want=("sigma2",))["sigma2"]
        X.rate("mmse_perframe_eig_" + prec, lambda: ctx.mmse_perframe_eig(tx0, rx0, s2n, out=He), n, 159 * cbytes + cbytes // 2, 2 * 22472,
               *((37.2, mp["fp64_dmma_tflops"]) if prec == "f64" else (None, None)), config="configs[3]",
               note="per-frame sigma2 for frames that share |tx_k|^2: two shared 53x53 complex products (2 x 22 472 flop) + a per-frame scaling "
                    "instead of the 4.4e5-flop solve; hbm_frac is on the algorithmic 159 c + sigma2 per frame")
        # low-rank per-frame MMSE (wifi_lowrank.cu): the synthetic covariance has rank 4 (4 channel taps), sigma2 per frame; one launch
        rk = ctx.mmse_lowrank_prepare(R)
        X.rate("mmse_perframe_lowrank_" + prec, lambda: ctx.mmse_perframe_lowrank(tx0, rx0, s2n, out=He), n, 159 * cbytes + cbytes // 2, 5874,
               *((74.4, mp["fp32_fma_tflops"]) if prec == "f32" else (37.2, mp["fp64_fma_tflops"])), config="configs[3]", rank=rk,
               arithmetic="FP32" if prec == "f32" else "FP64",
               note="per-frame sigma2 (and per-frame |tx_k|^2) for a covariance of rank r <= 8: H = U (sigma2 L^-1 + U^H diag(|x|^2) U)^-1 U^H (conj(x) rx), "
                    "an r x r solve per frame in registers, ONE launch on the frame's own 159 c + sigma2; flops at r = 4: 53 x (16 + 16 + 6) x 2 + 53 x 16 x 2 "
                    "+ ~150 for the 4 x 4 solve = 5 874 (the 53 x 53 solve: 441 949)")
        del He
        if full:
            # receiver front-end: 15 x 64 packet samples + 128 lptot samples in, 15 x 53 + 53 values + ow2 out
            nfe = min(n, 1 << 18)
            cdt, rdt = (torch.complex64, torch.float32) if prec == "f32" else (torch.complex128, torch.float64)
            pk = torch.randn(nfe, 1200, dtype=cdt, device=tx0.device); lp = torch.randn(nfe, 160, dtype=cdt, device=tx0.device)
            fe_out = (torch.empty(nfe, NBLK, NSC, dtype=cdt, device=tx0.device), torch.empty(nfe, NSC, dtype=cdt, device=tx0.device),
                      torch.empty(nfe, dtype=rdt, device=tx0.device))
            X.rate("frontend_" + prec, lambda: ctx.frontend(pk, lp, out=fe_out), nfe, (1088 + 848) * cbytes + cbytes // 2, config="8(f)1")
            del fe_out
            # fused receiver chain (wifi_rx_chain_batch): time samples of both sides -> 5 estimates + equalized symbols + ow2 in ONE
            # launch; 1 280 c in (rx: 15 x 64 + 128, tx: block 0 + 128), 5 x 53 + 795 c out.  The unfused path moves 5 886 c.
            pk2 = torch.randn(nfe, 1200, dtype=cdt, device=tx0.device); lp2 = torch.randn(nfe, 160, dtype=cdt, device=tx0.device)
            want = ("lt_ls", "linear", "cubic", "sinc", "mmse_cconv", "eq", "ow2")
            ch_out = ctx.rx_chain(pk, lp, pk2, lp2, want=want)
            X.rate("rx_chain_" + prec, lambda: ctx.rx_chain(pk, lp, pk2, lp2, want=want, out=ch_out), nfe, (1280 + 5 * 53 + 795) * cbytes + cbytes // 2,
                   config="8(f)1 fused: time samples -> LT_LS, PS_Linear/Cubic/Sinc, PS_MMSE (main.c:148 convention), equalizer, one launch",
                   note="the front-end's OFDM symbols never round-trip through HBM; the same planes through wifi_frontend_batch x 2 + "
                        "the stand-alone estimators and equalizer move 5 886 c per frame")
            del pk, lp, pk2, lp2, ch_out
            # PS_MMSE in main.c:148's calling convention (R_f = H_ls H_ls^H), rank-one closed form: 212 c + ow2 per frame
            Hc = torch.empty_like(tx0)
            X.rate("mmse_cconv_" + prec, lambda: ctx.mmse_cconv(tx0, rx0, s2n, H, out=Hc), n, 212 * cbytes + cbytes // 2, config="main.c:148 convention")
            del Hc
            npv = min(npf, 1 << 15)
            X.rate("mmse_perframe_pivot_" + prec, lambda: ctx.mmse_perframe(Rp, tx0[:npv], rx0[:npv], s2[:npv], flags=wifi.SOLVE_PIVOT, out=Hp[:npv]), npv,
                   159 * cbytes, 441949, 37.2, mp["fp64_dmma_tflops"], config="configs[3], general R", arithmetic="FP64 (pivoted LU, trailing updates on DMMA)")
            # utils.c routines, batched: 53 x 53 complex multiply() and inverse() through the mirror of the
            # reference interface (allocation of the result and, for inverse, the singularity check included)
            nb = 8192
            g = torch.Generator(device="cuda").manual_seed(7)
            A = torch.randn(nb, NSC, NSC, dtype=cdt, device="cuda", generator=g)
            A = A @ A.conj().transpose(1, 2) / NSC + torch.eye(NSC, dtype=cdt, device="cuda")       # Hermitian PD, well conditioned
            nom, meas = (74.4, mp["fp32_fma_tflops"]) if prec == "f32" else (37.2, mp["fp64_fma_tflops"])
            note = "frames_per_s = matrices/s; the reference: multiply() 9.9 ms, inverse() 13 s per 53 x 53 matrix on one core (SURVEY 8a)"
            X.rate("utils_multiply_53_" + prec, lambda: ctx.multiply(A, A), nb, 3 * NSC * NSC * cbytes, 8 * NSC ** 3, nom, meas, note=note)
            X.rate("utils_inverse_53_" + prec, lambda: ctx.inverse(A), nb, 2 * NSC * NSC * cbytes, 8 * NSC ** 3, nom, meas, note=note)
            del A
        del tx0, rx0, Hm, Hp, H, outs
        torch.cuda.empty_cache()
    out = X.finish()
    for prec in ("f32", "f64"):
        out["mmse_perframe_eig_" + prec]["speedup_vs_direct_solve"] = out["mmse_perframe_eig_" + prec]["frames_per_s"] / out["mmse_perframe_hpd_" + prec]["frames_per_s"]
        out["mmse_perframe_lowrank_" + prec]["speedup_vs_direct_solve"] = out["mmse_perframe_lowrank_" + prec]["frames_per_s"] / out["mmse_perframe_hpd_" + prec]["frames_per_s"]
    return out


def parity_sample(ctx, tx, rx, H, R, d, n=256):
    """accuracy.parity: the step's own output against the CPU oracle (long-double filter + long-double product on the same
    FP32-rounded inputs) on the first n frames of this rank, at the survey's floor 1e-3 and at the stated FP32 floor 1e-2."""
    import synth
    from oracle.pyoracle import Oracle
    o = Oracle()
    W = o.mmse_filter(R.cpu().numpy(), d.cpu().numpy())
    t, r = tx[:n].cpu().numpy().astype(np.complex128), rx[:n].cpu().numpy().astype(np.complex128)
    ref = o.mmse_apply(W, r / t)
    got = H[:n].cpu().numpy().astype(np.complex128)
    peak = np.abs(ref).max(axis=-1, keepdims=True)
    return {"frames": int(n), "max_rel_err_floor_1e-3": float(synth.rel_err(got, ref, 1e-3)), "max_rel_err_floor_1e-2": float(synth.rel_err(got, ref, 1e-2)),
            "max_err_over_frame_peak": float((np.abs(got - ref) / peak).max()), "bound": 1e-4,
            "definition": "max over sub-carriers of |d| / max(|ref_k|, floor * max_k |ref|); oracle = oracle/wifi_oracle.c (long double)"}


def run_ours(args):
    import torch
    wifi = importlib.import_module("80211parallelestimation_b200")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    ctx = wifi.WifiContext(local)
    peaks = measured_peaks()
    n_local = args.frames
    n_total = n_local * world
    shard = wifi.ShardedEstimator(n_total, rank, world)
    assert shard.n_local == n_local

    # ---- data resident in HBM: this rank's contiguous shard of the global synthetic sequence ----
    fr = ctx.synth_frames(n_local, "f32", first_frame=shard.lo, want=("tx_symb", "rx_symb", "H_true"))
    tx = fr["tx_symb"][:, 0, :].contiguous(); rx = fr["rx_symb"][:, 0, :].contiguous(); Htrue = fr["H_true"]
    del fr
    torch.cuda.empty_cache()
    H = torch.empty_like(tx)
    # shared filter: formed once on the device in FP64 (not part of the step)
    R = ctx.synth_covariance()
    d = torch.full((NSC,), OW2 / AMP ** 2, dtype=torch.float64, device=tx.device); d[26] = OW2 / 1e-8
    t0 = time.perf_counter(); ctx.mmse_filter_form(R, d, want_W=False); torch.cuda.synchronize(); filter_ms = 1e3 * (time.perf_counter() - t0)

    step = lambda: ctx.mmse_shared(tx, rx, out=H)
    sampler = ClockSampler(local); sampler.start()
    l0 = ctx.launches
    total_ms, per = time_steps(step, args.steps, args.warmup, torch, dist)
    launches = (ctx.launches - l0) * args.steps // (args.steps + args.warmup)
    t = torch.tensor([total_ms], dtype=torch.float64, device=tx.device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t[0])
    value = n_total * args.steps / (total_ms * 1e-3)
    kernel_ms = float(np.mean(per))

    # ---- statistics: the only (optional) collective, after the timed region ----
    stats = shard.reduce_stats(ctx.error_stats(H, Htrue))
    parity = parity_sample(ctx, tx, rx, H, R, d) if rank == 0 else None

    # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region ----
    def wall(fn, reps):
        """reps calls bracketed by synchronize + barrier on both sides; returns the MAX over ranks of the wall time (s)."""
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt_ = time.perf_counter() - t0
        tt = torch.tensor([dt_], dtype=torch.float64, device=tx.device)
        if dist is not None:
            dist.barrier()
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0])

    with numa_local(local) as numa:
        htx = pinned(wifi, (n_local, NSC), np.complex64); hrx = pinned(wifi, (n_local, NSC), np.complex64); hH = pinned(wifi, (n_local, NSC), np.complex64)
        htx[:] = tx.cpu().numpy(); hrx[:] = rx.cpu().numpy(); hH[:] = 0
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        ctx.mmse_shared(htx, hrx, out=hH)
    e2e_s = wall(lambda: ctx.mmse_shared(htx, hrx, out=hH), e2e_steps)          # each call returns after the D2H of the result has completed
    e2e_val = n_total * e2e_steps / e2e_s
    e2e_ok = bool(np.allclose(hH[:1024], H[:1024].cpu().numpy(), rtol=1e-5, atol=1e-8))
    # the same workload when the frames share their (known) tx block vector: the LS divide folds into the filter and only rx crosses the bus
    ctx.mmse_filter_fold_tx(tx[0])
    Hrx = ctx.mmse_shared_rx(rx)
    shared_tx_holds = bool((tx == tx[0]).all())      # false for the synthetic frames (random BPSK data): the variant is timed on the same bytes, not compared
    ctx.mmse_shared_rx(hrx, out=hH)
    rx_s = wall(lambda: ctx.mmse_shared_rx(hrx, out=hH), e2e_steps)
    rx_ok = bool(np.allclose(hH[:1024], Hrx[:1024].cpu().numpy(), rtol=1e-5, atol=1e-8))
    del Hrx
    # PCIe ceiling under the same conditions: the step's H2D and D2H volumes as ONE pinned copy each, issued together, on every rank at once
    h2d_b, d2h_b = 2 * n_local * NSC * 8, n_local * NSC * 8
    ctx.pcie_probe(htx, hH)                           # (allocates the probe's device buffers)
    pc_both = wall(lambda: (ctx.pcie_probe(htx, hH), ctx.pcie_probe(hrx, None)), 3) / 3       # 2 x 424 MB down, 424 MB up
    pc_h2d = wall(lambda: (ctx.pcie_probe(htx, None), ctx.pcie_probe(hrx, None)), 3) / 3
    pc_d2h = wall(lambda: ctx.pcie_probe(None, hH), 3) / 3
    pc_rx = wall(lambda: ctx.pcie_probe(hrx, hH), 3) / 3                                      # the shared-tx variant's volumes
    clocks = sampler.result()          # sampled every 5 ms from the first warm-up step to the end of the e2e region

    mp = ctx.measure_peaks()           # on-box FP32/FP64 FMA, DMMA and copy ceilings
    extras = None
    if not args.no_extras:
        del htx, hrx, hH
        extras = extras_block(wifi, ctx, torch, dist, world, shard.lo, peaks, mp, min(n_local, 1 << 20), max(3, min(args.steps, 10)), 3,
                              full=(world == 1))
    if rank == 0:
        bytes_per_frame = 159 * 8          # read tx 53c + rx 53c, write H 53c, FP32 complex (fused LS + filter)
        achieved = n_local * bytes_per_frame / (kernel_ms * 1e-3) / 1e9
        ceil_val = n_total / pc_both
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "configs[2]: batched PS_MMSE, shared Rhh/sigma2, %d frames per GPU as one complex GEMM (LS divide fused), FP32 I/O" % n_local,
                       "frames_per_gpu": n_local, "frames_total": n_total, "l2_policy": "inputs per pass (%.2f GB) >> 126 MB L2, no flush" % (n_local * bytes_per_frame / 1e9),
                       "parallelism": "frame-sharded x%d, no data-path collective" % world, "filter_form_ms": filter_ms,
                       "arithmetic": "FP32 I/O; products as 3xTF32 on tcgen05 with FP32 accumulation in TMEM; filter formed once in double-double"},
            "roofline": {"bound": "hbm", "kernel": "mmse_shared (fused LS divide + 53x53 complex filter GEMM)", "achieved": achieved, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": int(NCU_DRAM_BYTES_PER_MI_FRAMES * n_local / (1 << 20)),
                         "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of this kernel at 1 Mi frames "
                                           "(profiles/r02_ncu_kernels.txt), scaled to this launch's frame count",
                         "peak_source": peaks["source"],
                         "algorithmic_bytes_per_frame": bytes_per_frame, "kernel_ms": kernel_ms,
                         "tensor_TFLOPs_3xTF32": n_local * 3 * 2 * 112 * 112 / (kernel_ms * 1e-3) / 1e12},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
                    "steps": e2e_steps, "matches_device_result": e2e_ok, "api": "WifiContext.mmse_shared(numpy pinned) -> wifi_mmse_shared_host", "host_numa": numa.state,
                    "pcie_ceiling": {"value": ceil_val, "unit": UNIT, "ms_per_step": 1e3 * pc_both,
                                     "what": "the step's H2D (2 x 424 MB) and D2H (424 MB) volumes per GPU as plain pinned cudaMemcpyAsync copies issued together, "
                                             "all ranks at once, max over ranks (wifi_pcie_probe): no kernel, no chunking",
                                     "h2d_gbs_per_gpu_alone": h2d_b / pc_h2d / 1e9, "d2h_gbs_per_gpu_alone": d2h_b / pc_d2h / 1e9,
                                     "h2d_gbs_per_gpu_duplex": h2d_b / pc_both / 1e9, "aggregate_h2d_gbs_duplex": world * h2d_b / pc_both / 1e9},
                    "pcie_ceiling_gbs": world * (h2d_b + d2h_b) / pc_both / 1e9,
                    "frac_of_ceiling": e2e_val / ceil_val,
                    "shared_tx": {"value": n_total * e2e_steps / rx_s, "unit": UNIT, "h2d_bytes_per_step": int(d2h_b), "d2h_bytes_per_step": int(d2h_b),
                                  "api": "WifiContext.mmse_filter_fold_tx + mmse_shared_rx(numpy pinned) -> wifi_mmse_shared_rx_host",
                                  "matches_device_result": rx_ok, "frac_of_ceiling": (n_total * e2e_steps / rx_s) / (n_total / pc_rx),
                                  "note": "frames that share one known tx block vector (training symbols): rx/tx folds into the filter, 848 B instead of 1 272 B per frame "
                                          "cross the bus; the synthetic frames carry random BPSK data (shared tx holds: %s), so this line is a bytes/throughput "
                                          "measurement, the headline e2e above is the per-frame-tx call" % shared_tx_holds}},
            "gpu_launches": int(launches), "clocks": clocks,
            "measured_peaks": dict(mp, nominal={"fp32_fma_tflops": 74.4, "fp64_fma_tflops": 37.2, "fp64_dmma_tflops": 37.2,
                                                "note": "148 SMs x 128 (FP32) / 64 (FP64) FMA lanes x 2 x 1.965 GHz"}),
            "accuracy": {"nmse_vs_true_channel": stats["nmse"], "max_abs_err": stats["max_abs_err"], "count": stats["count"], "parity": parity},
        }
        if world == 1 and not args.no_cpu:
            cb, _ = cpu_reference_rate(1 << 21)
            cb["estimators"] = cpu_estimator_baselines()
            line["cpu_baseline"] = cb
        if extras is not None:
            line["extras"] = extras
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints to fd 1 (NCCL's version banner, ...)
    was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1 << 20, help="frames per GPU")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
