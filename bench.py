#!/usr/bin/env python
"""bench.py -- posterior-sample x input forwards / second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|1|3|4|5] [--scaling strong|weak]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload = BASELINE config 2, the one the metric is quoted on: BayesByBackprop MLP 784-512-512-10, B = 4096
input rows, n = 1024 posterior samples.  One "step" = one full Monte-Carlo posterior predictive pass: in-kernel Philox
sampling of all 669,706 weights of every sample, three dense layers, softmax, accumulation of sum p and sum p^2.

Multi-GPU = the north-star split: STRONG scaling -- the n = 1024 samples of one predict are cut into
contiguous global-id ranges, rank g of G evaluates [g*n/G, (g+1)*n/G), and ONE all-reduce of the [B,C] moment buffers
(this library's NVLink peer-memory kernel, csrc/collective.cu) combines them.  value = n * B forwards / (max-over-ranks
device time per step).  `--scaling weak` (n samples PER GPU) is kept as a secondary arm; under torchrun a short weak run
is also reported in the "weak" field of the line.

Legs of the default run (rank 0 prints ONE JSON line):
  value      inputs resident in HBM; K steps enqueued back to back, CUDA events on the launching stream
  e2e        the plugin call `PosteriorModel.predict_moments(x_numpy, n)` -- host array in, host arrays out, H2D of x and
             D2H of mean / E[p^2] inside every step, the same device work as the value leg
  roofline   algorithmic FLOPs of the fused-kernel launches / their CUDA-event time (recorded by the library on the
             stream that launches them)
  cpu_baseline  the oracle's loop-faithful NumPy port of posteriormodel.py:81-101 on the host cores (N = 1 only)
Before the timed region the GPU is pre-heated with the same steps for --preheat seconds (default 2 s) so that the
measurement is taken in the SUSTAINED clock / power regime; `roofline.peak` is therefore the sustained cuBLAS figure of
MEASURED_PEAKS.json (the burst figure and the fraction of it are printed beside it).
"""
from __future__ import annotations

import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # the CPU arm uses every host core; torchrun exports OMP_NUM_THREADS=1, which would halve it (BLAS reads these at import)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import argparse
import glob
import json
import math
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = [784, 512, 512, 10]
B_ROWS = 4096
N_SAMPLES = 1024
METRIC = "posterior-sample x input forwards/sec"
UNIT = "forwards/s"
WORKLOADS = {
    1: "BayesByBackprop MLP 784-512-10, predict n=100 samples, batch 10000 (BASELINE config 1)",
    2: "BayesByBackprop MLP 784-512-512-10, batch 4096, n=1024 posterior samples (BASELINE config 2)",
    3: "HMC stored-posterior predictive: 10k fp32 weight samples of MLP 784-1024-1024-10 streamed from HBM, batch 64 (BASELINE config 3)",
    4: "Bayesian CNN conv32-conv64-pool-dense128-dense10 on CIFAR-10 shape, batch 256, n=256 samples (BASELINE config 4)",
    5: "Chernoff-bound MC verification: 100k posterior samples x 128 perturbed inputs, MLP 784-512-512-10 (BASELINE config 5)",
}


def _ncu_traffic(pattern):
    """DRAM bytes (read + write) of ONE launch of the dominant kernel from the newest committed ncu --set full summary."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    for f in reversed(files):
        try:
            tot = 0.0
            for line in open(f):
                parts = line.split()
                if parts and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(parts[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[parts[2]]
            if tot:
                return tot, os.path.relpath(f, ROOT)
        except Exception:
            continue
    return None, None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback of B200_PROFILING.md"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  nvidia-smi needs ~100 ms before its first row
    and samples every 20 ms, while K steps of the 8-GPU strong split last ~15 ms -- so the sampler is started during the
    pre-heat (the same steps, back to back with the timed ones) and `mark()` notes where the timed region begins; if fewer
    than 3 rows fall inside it, the rows of the last second of pre-heat are used as well (`window` says which)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.t_mark, self.t_end = [], None, index, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        self.t_mark = time.time()

    def end(self):
        self.t_end = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t_end is None:
            self.end()
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        t0 = self.t_mark if self.t_mark is not None else 0.0
        inside = [r for (t, r) in self.rows if t0 <= t <= self.t_end + 0.03]
        window = "timed region"
        if len(inside) < 3:
            inside = [r for (t, r) in self.rows if t0 - 1.0 <= t <= self.t_end + 0.03]
            window = "timed region + the last second of pre-heat (identical steps, back to back)"
        sm, mx, reasons, pw = [], [], set(), []
        for r in inside:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's loop restated by the oracle (TensorFlow is not installable on this image)
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_pass(n_samples, B, dims=DIMS):
    """posteriormodel.py:94-101 restated by the oracle: per sample np.random.normal over every tensor (float64,
    single-threaded MT19937, as in the reference), cast to fp32, fp32 forward on all host cores, fp32 running sum, / float(n)."""
    from oracle import oracle as o
    L = o.mlp(dims)
    mean, std = o.synthetic_posterior(L, seed=1)
    x = o.synthetic_x((B, dims[0]), seed=0)
    np.random.seed(0)
    t0 = time.perf_counter()
    out = o.predict(L, mean, std, x, n_samples, eps=None)
    dt = time.perf_counter() - t0
    assert np.isfinite(out).all()
    return dt


def _blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(i.get("num_threads", 1)) for i in threadpool_info()] or [1])
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))


def run_reference(args):
    """--impl reference: the reference's own CPU path (oracle port) on the host cores, bounded sample of the same workload.
    Under torchrun rank 0 alone runs it."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(os.cpu_count() or 1)
    except Exception:
        pass
    n_ref = int(os.environ.get("DBX_REF_SAMPLES", "8"))
    for _ in range(args.warmup):
        cpu_reference_pass(1, B_ROWS)
    ts = [cpu_reference_pass(n_ref, B_ROWS) for _ in range(args.steps)]
    dt = float(np.mean(ts))
    val = n_ref * B_ROWS / dt
    cores = _blas_threads()
    sample = "%d of %d posterior samples x all %d rows per step" % (n_ref, N_SAMPLES, B_ROWS)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[2],
                   "reference_arm": "loop-faithful NumPy port of posteriormodel.py:81-101 (TensorFlow not installable here)",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def _setup_dist():
    import torch
    import torch.distributed as dist
    c = Ctx()
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    if c.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(c.local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", c.local))
    else:
        torch.cuda.set_device(0)
    c.dev = torch.device("cuda", torch.cuda.current_device())
    c.torch, c.dist = torch, dist
    return c


def _barrier(c):
    if c.world > 1:
        c.dist.barrier()
    c.torch.cuda.synchronize()


def _max_over_ranks(c, vals):
    t = c.torch.tensor(list(vals), dtype=c.torch.float64, device=c.dev)
    if c.world > 1:
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
    return [float(v) for v in t]


def timed_steps(c, step, steps, warm, preheat_s, first_index=0):
    """W untimed warm-up steps (+ pre-heat to the sustained regime), then EXACTLY `steps` steps between barrier +
    synchronize on both sides, one CUDA event per step boundary on the launching stream.  Returns (ms per step = max over
    ranks, per-step quartile summary of this rank, host enqueue ms per step, steps of pre-heat, next step index)."""
    torch = c.torch
    i = first_index
    step(i); i += 1                       # first call: workspace allocation, tensor maps, communicator rendezvous
    _barrier(c)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(warm):
        step(i); i += 1
    b.record()
    _barrier(c)
    est = _max_over_ranks(c, [a.elapsed_time(b) / max(warm, 1)])[0]
    pre = int(min(20000, math.ceil(preheat_s * 1e3 / max(est, 1e-3)))) if preheat_s > 0 else 0
    for _ in range(pre):                  # the same count on every rank (the steps hold a collective)
        step(i); i += 1
    _barrier(c)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    _barrier(c)
    t0 = time.perf_counter()
    for k in range(steps):
        evs[k].record()
        step(i); i += 1
    evs[steps].record()
    host_ms = (time.perf_counter() - t0) * 1e3 / steps
    _barrier(c)
    total = evs[0].elapsed_time(evs[steps]) / steps
    per = np.array([evs[k].elapsed_time(evs[k + 1]) for k in range(steps)])
    q = {"min": float(per.min()), "p25": float(np.percentile(per, 25)), "median": float(np.median(per)),
         "p75": float(np.percentile(per, 75)), "max": float(per.max()),
         "first_quarter_mean": float(per[:max(1, steps // 4)].mean()), "last_quarter_mean": float(per[-max(1, steps // 4):].mean())}
    return _max_over_ranks(c, [total])[0], q, host_ms, pre, i


def _mlp_posterior_model(c, dims, mode, tag, shard):
    """A saved synthetic Gaussian posterior reloaded as PosteriorModel (the user-facing object)."""
    import deepbayes_b200 as db
    from deepbayes_b200 import formats
    from oracle import oracle as o   # synthetic-input generators only
    L = o.mlp(dims)
    mean, std = o.synthetic_posterior(L, seed=1)
    layers = [db.Dense(dims[1], activation="relu", input_shape=(dims[0],))]
    for h in dims[2:-1]:
        layers.append(db.Dense(h, activation="relu"))
    layers.append(db.Dense(dims[-1], activation="softmax"))
    pdir = "/tmp/dbx_bench_posterior_%s_r%d" % (tag, c.rank)
    formats.save_gaussian_posterior(pdir, db.Sequential(layers), mean, std, {"input_upper": 1.0, "input_lower": 0.0})
    return db.PosteriorModel(pdir, mode=mode, seed=1234, shard=shard)


def _profile_read(lib):
    import ctypes as C
    pms, pfl, pn = C.c_double(), C.c_double(), C.c_int64()
    lib.dbx_profile_read(C.byref(pms), C.byref(pfl), C.byref(pn))
    return pms.value, pfl.value, pn.value


def run_config2(args, c):
    import deepbayes_b200 as db  # noqa: F401
    from deepbayes_b200 import _lib
    from deepbayes_b200 import distributed as dd
    from oracle import oracle as o   # synthetic inputs + the cpu_baseline leg only
    torch = c.torch
    lib = _lib.load()
    B, n, world, rank = args.batch, args.samples, c.world, c.rank
    warm = max(args.warmup, 3)
    strong = args.scaling == "strong"
    pm = _mlp_posterior_model(c, DIMS, args.mode, "cfg2", shard=True)
    eng = pm._engine
    x_np = o.synthetic_x((B, DIMS[0]), seed=0)
    x_dev = torch.from_numpy(x_np).to(c.dev)
    s1 = torch.empty((B, 10), dtype=torch.float32, device=c.dev)
    s2 = torch.empty_like(s1)

    def make_step(n_total, off, cnt, seed=1234):
        def step(i):
            # global sample ids: this rank owns [off, off + cnt) of step i's n_total draws
            eng.predict_sums(x_dev, cnt, seed=seed, s0=i * n_total + off, mode=args.mode, second_moment=True, out=(s1, s2))
            if world > 1:
                dd.allreduce_sum_(s1, s2)
        return step

    if strong:
        off, cnt = dd.shard_range(n, rank, world)
        n_total = n
    else:
        off, cnt, n_total = rank * n, n, n * world
    step = make_step(n_total, off, cnt)

    # ---- timed region 1: inputs resident in HBM
    step(0)
    _barrier(c)
    path = eng.last_path()
    if path != "fused_tcgen05" and args.mode == "bf16":
        raise RuntimeError("bench expected the fused tcgen05 path, got %r" % path)
    clocks = ClockSampler(torch.cuda.current_device())
    lib.dbx_profile(0)
    res = {}

    def run_value_leg():
        torch_ = torch
        i = 1
        step(i); i += 1
        _barrier(c)
        a, b = torch_.cuda.Event(enable_timing=True), torch_.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(warm):
            step(i); i += 1
        b.record()
        _barrier(c)
        est = _max_over_ranks(c, [a.elapsed_time(b) / warm])[0]
        pre = int(min(20000, math.ceil(args.preheat * 1e3 / max(est, 1e-3)))) if args.preheat > 0 else 0
        clocks.start()
        for _ in range(pre):
            step(i); i += 1
        _barrier(c)
        evs = [torch_.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        launches0 = _lib.launch_count()
        lib.dbx_profile(1)
        _barrier(c)
        clocks.mark()
        t0 = time.perf_counter()
        for k in range(args.steps):
            evs[k].record()
            step(i); i += 1
        evs[args.steps].record()
        host_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        _barrier(c)
        clocks.end()
        res["ms"] = evs[0].elapsed_time(evs[args.steps]) / args.steps
        per = np.array([evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)])
        res["quart"] = {"min": float(per.min()), "median": float(np.median(per)), "max": float(per.max()),
                        "first_quarter_mean": float(per[:max(1, args.steps // 4)].mean()),
                        "last_quarter_mean": float(per[-max(1, args.steps // 4):].mean())}
        res["host_ms"] = host_ms
        res["pre"] = pre
        res["prof"] = _profile_read(lib)
        lib.dbx_profile(0)
        res["launches"] = _lib.launch_count() - launches0
        res["clk"] = clocks.stop()
        return i

    next_i = run_value_leg()
    coll = dd.collective_kind() if world > 1 else "none (1 GPU)"

    # ---- timed region 2: end to end through the plugin call, host arrays in and out
    # x lives in page-locked host memory (a NumPy view of it): dbx_upload_f32 then issues one DMA per step
    x_pin = torch.from_numpy(x_np).pin_memory()
    x_host = x_pin.numpy()
    pm._next_sample = next_i * n_total

    def step_e2e(_i):
        return pm.predict_moments(x_host, n_total if strong else n)

    if strong or world == 1:
        for k in range(2):
            out = step_e2e(k)
        _barrier(c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            out = step_e2e(k)
        e1.record()
        _barrier(c)
        ms_e2e = e0.elapsed_time(e1) / args.steps
        mean_out, m2_out = out
        assert mean_out.shape == (B, 10) and np.isfinite(mean_out).all() and np.allclose(mean_out.sum(-1), 1, rtol=1e-3)
        assert m2_out.shape == (B, 10) and (m2_out <= mean_out + 1e-6).all()
    else:
        ms_e2e = float("nan")

    # ---- N ranks == 1 rank on the same global sample ids: rank 0 recomputes all n samples alone
    shard_check = None
    if world > 1:
        st = make_step(n_total, off, cnt, seed=4242)
        st(7)
        torch.cuda.synchronize()
        got1, got2 = s1.clone(), s2.clone()
        if rank == 0:
            f1, f2 = eng.predict_sums(x_dev, n_total, seed=4242, s0=7 * n_total, mode=args.mode, second_moment=True)
            d1 = float(((got1 - f1).abs() / n_total).max()); d2 = float(((got2 - f2).abs() / n_total).max())
            ok = torch.allclose(got1 / n_total, f1 / n_total, rtol=1e-5, atol=1e-6) and torch.allclose(got2 / n_total, f2 / n_total, rtol=1e-5, atol=1e-6)
            shard_check = "ok" if ok else "FAILED"
            shard_detail = {"max_abs_diff_mean": d1, "max_abs_diff_second_moment": d2, "ranks": world, "samples": n_total,
                            "tolerance": "rtol 1e-5, atol 1e-6 on the [B,C] means"}
        # every rank must hold the same bits (sums are taken in rank order on every rank)
        chk = torch.stack([got1.double().sum(), got2.double().sum()])
        lo, hi = chk.clone(), chk.clone()
        c.dist.all_reduce(lo, op=c.dist.ReduceOp.MIN); c.dist.all_reduce(hi, op=c.dist.ReduceOp.MAX)
        same_bits = bool((lo == hi).all())
        # the plugin call with the input rows sharded over the ranks' PCIe links + all-gather == the same call with the whole batch
        # uploaded by every rank (same sample ids)
        if strong and dd.sharded_upload_enabled(world):
            pm._next_sample = 9 * n_total
            a1, a2 = pm.predict_moments(x_host, n_total)
            os.environ["DBX_NO_SHARDED_UPLOAD"] = "1"
            pm._next_sample = 9 * n_total
            b1, b2 = pm.predict_moments(x_host, n_total)
            os.environ["DBX_NO_SHARDED_UPLOAD"] = "0"
            up_ok = bool(np.array_equal(a1, b1) and np.array_equal(a2, b2))
            if rank == 0:
                shard_detail["sharded_upload_equals_full_upload"] = up_ok
                if not up_ok:
                    shard_check = "FAILED"
        _barrier(c)

    # ---- secondary arm under torchrun: weak scaling (n samples per GPU)
    weak = None
    if world > 1 and strong and not args.no_weak:
        wstep = make_step(n * world, rank * n, n, seed=99)
        wms, _, _, _, _ = timed_steps(c, wstep, max(5, min(args.steps, 20)), 3, 0.0, first_index=100000)
        weak = {"scaling": "weak", "samples_per_gpu": n, "ms_per_step": wms, "value": world * n * B / (wms * 1e-3), "unit": UNIT}

    ms, ms_e2e = _max_over_ranks(c, [res["ms"], ms_e2e if ms_e2e == ms_e2e else -1.0])

    if rank == 0:
        peaks, peak_src = _peaks()
        flops_fwd = eng.flops_per_forward
        forwards = n_total * B
        value = forwards / (ms * 1e-3)
        pms, pfl, pn = res["prof"]
        tf_achieved = (pfl / (pms * 1e-3)) / 1e12 if pms > 0 else None
        sustained = res["pre"] > 0 or ms * args.steps >= 1000.0
        peak_tf = float(peaks["bf16_tflops_sustained"] if sustained else peaks["bf16_tflops"]) if args.mode == "bf16" else None
        traffic, traffic_file = _ncu_traffic("r*_ncu_fused_pair_summary.txt") if path == "fused_tcgen05" else (None, None)
        roof = {
            "bound": "tensor",
            "kernel": "fused_mlp2_kernel (tcgen05 cta_group::2, TMA, TMEM-resident activations)" if path == "fused_tcgen05" else "dense_kernel (generic FFMA)",
            "achieved": tf_achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": (tf_achieved / peak_tf) if (tf_achieved and peak_tf) else None,
            "peak_burst": float(peaks["bf16_tflops"]),
            "frac_of_burst": (tf_achieved / float(peaks["bf16_tflops"])) if tf_achieved else None,
            "regime": ("sustained: %d pre-heat steps (%.1f s) ran before the timed steps, peak = cuBLAS bf16 back to back for 4 s"
                       % (res["pre"], res["pre"] * ms / 1e3)) if sustained else "burst: short timed region without pre-heat, peak = best-of-10 cuBLAS bf16",
            "peak_source": peak_src,
            "traffic": traffic,
            "traffic_note": "DRAM read+write bytes of one 256-sample launch from %s; algorithmic bytes of that launch: 256 * 1.34 MB "
                            "sampled bf16 weights + 6.4 MB inputs" % traffic_file,
            "kernel_ms_per_step": pms / args.steps, "kernel_launches_per_step": pn / args.steps,
            "algorithmic_flops_per_forward": flops_fwd, "algorithmic_flops_per_launch": pfl / pn if pn else None,
            "whole_step_tflops": (forwards / world * flops_fwd / (ms * 1e-3)) / 1e12,
            "whole_step_frac": ((forwards / world * flops_fwd / (ms * 1e-3)) / 1e12 / peak_tf) if peak_tf else None,
            "whole_step_frac_of_burst": (forwards / world * flops_fwd / (ms * 1e-3)) / 1e12 / float(peaks["bf16_tflops"]),
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_reference_pass(1, B)  # warm the BLAS threads
            t1 = cpu_reference_pass(2, B) / 2
            n_ref = int(max(2, min(1024, round(15.0 / max(t1, 1e-3)))))
            dt = cpu_reference_pass(n_ref, B)
            cpu = {"value": n_ref * B / dt, "unit": UNIT, "cores": _blas_threads(), "kind": "port",
                   "sample": "%d of %d posterior samples x all %d rows (%.1f s); NumPy port of posteriormodel.py:81-101, "
                             "single-threaded np.random.normal + multi-threaded fp32 BLAS" % (n_ref, n, B, dt)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": args.mode if args.mode != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[2] if (B, n) == (B_ROWS, N_SAMPLES) else "MLP 784-512-512-10, batch %d, n=%d" % (B, n),
                       "global_samples_per_step": n_total, "samples_per_gpu": cnt, "path": path,
                       "preheat_steps": res["pre"], "step_ms_quartiles": res["quart"], "host_enqueue_ms_per_step": res["host_ms"],
                       "l2": "no explicit flush: each step streams %.2f GB of freshly sampled weights per GPU (> 126 MB L2)" % (cnt * eng.P * 2 / 1e9),
                       "parallelism": "posterior samples sharded x%d (contiguous global-id ranges), one all-reduce of the [B,C] moment "
                                      "buffers per step" % world,
                       "collective": coll},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": forwards / (ms_e2e * 1e-3) if ms_e2e > 0 else None, "unit": UNIT, "ms_per_step": ms_e2e if ms_e2e > 0 else None,
                    "call": "PosteriorModel.predict_moments(x_numpy, n): page-locked host array in, (mean, E[p^2]) host arrays out",
                    # summed over the ranks.  Sharded predict: every rank uploads 1 / world of the batch's rows and one all-gather
                    # over NVLink completes it (distributed.sharded_upload; DBX_NO_SHARDED_UPLOAD=1: every rank uploads all of it)
                    "h2d_bytes_per_step": int(x_np.nbytes) * (1 if (strong and world > 1 and dd.sharded_upload_enabled(world)) else world),
                    "h2d_split": ("rows sharded over ranks + all-gather" if (strong and world > 1 and dd.sharded_upload_enabled(world)) else "whole batch per rank"),
                    "d2h_bytes_per_step": int(2 * B * 10 * 4) * world},
            "gpu_launches": int(res["launches"]), "clocks": res["clk"],
        }
        if world > 1:
            line["shard_check"] = shard_check
            line["shard_check_detail"] = dict(shard_detail, all_ranks_hold_identical_bits=same_bits)
            if weak:
                line["weak"] = weak
        print(json.dumps(line), flush=True)
        if world > 1 and shard_check != "ok":
            raise SystemExit("shard_check failed: %r" % (shard_detail,))


# ------------------------------------------------------------------------------------------------------------------
# The other BASELINE configs as driver-runnable lines (python bench.py --config 1|3|4|5): same metric, own roofline
# ------------------------------------------------------------------------------------------------------------------
def run_other(args, c):
    import deepbayes_b200 as db
    from deepbayes_b200 import _lib
    from deepbayes_b200 import distributed as dd
    from deepbayes_b200.engine import Engine
    from oracle import oracle as o   # synthetic inputs only
    torch = c.torch
    lib = _lib.load()
    world, rank = c.world, c.rank
    warm = max(args.warmup, 3)
    peaks, peak_src = _peaks()
    cfg = args.config
    extra = {}
    if cfg == 1:
        pm = _mlp_posterior_model(c, [784, 512, 10], args.mode, "cfg1", shard=True)
        eng = pm._engine
        B, n = (args.batch if args.batch != B_ROWS else 10000), 100      # MNIST-test-shaped batch; also B = 1 and 128
        x_np = o.synthetic_x((B, 784), seed=0)
        units = n * B
        bound = "tensor"
    elif cfg == 5:
        pm = _mlp_posterior_model(c, DIMS, args.mode, "cfg5", shard=True)
        eng = pm._engine
        B, n = 128, 100000
        g = np.random.Generator(np.random.Philox(3))
        x_np = np.clip(o.synthetic_x((B, 784), seed=0) + g.uniform(-0.1, 0.1, (B, 784)).astype(np.float32), 0, 1)
        labels = g.integers(0, 10, size=B).astype(np.int32)
        units = n * B
        bound = "tensor"
    elif cfg == 4:
        m = db.Sequential([db.Conv2D(32, 3, activation="relu", input_shape=(32, 32, 3)), db.Conv2D(64, 3, activation="relu"),
                           db.MaxPooling2D(2), db.Flatten(), db.Dense(128, activation="relu"), db.Dense(10, activation="softmax")])
        eng = Engine(m)
        mean, std = o.synthetic_posterior(o.cnn_cfg4(), seed=1)
        eng.set_gaussian(mean, std)
        B, n = 256, 256
        x_np = o.synthetic_x((B, 32, 32, 3), seed=0)
        units = n * B
        bound = "tensor"
    elif cfg == 3:
        dims = [784, 1024, 1024, 10]
        layers = [db.Dense(1024, activation="relu", input_shape=(784,)), db.Dense(1024, activation="relu"), db.Dense(10, activation="softmax")]
        eng = Engine(db.Sequential(layers))
        mean, std = o.synthetic_posterior(o.mlp(dims), seed=1)
        eng.set_gaussian(mean, std)
        B, n = args.batch if args.batch != B_ROWS else 64, args.stored
        x_np = o.synthetic_x((B, 784), seed=0)
        units = n * B
        bound = "hbm"
    else:
        raise SystemExit("unknown --config %r" % cfg)
    off, cnt = dd.shard_range(n, rank, world)
    x_dev = torch.from_numpy(x_np).to(c.dev)
    outs = []

    if cfg == 3:
        # synthetic stored posterior = Gaussian draws written straight into this rank's rows [off, off+cnt) of the slab
        pitch = (eng.P + 3) // 4 * 4
        slab = torch.empty((cnt, pitch), dtype=torch.float32, device=c.dev)
        for s0 in range(0, cnt, 256):
            nc = min(256, cnt - s0)
            slab[s0:s0 + nc, :eng.P] = eng.sample(nc, seed=7, s0=off + s0)
        if pitch != eng.P:
            slab[:, eng.P:] = 0
        eng.set_samples(slab)
        freq = torch.full((cnt,), 1.0 / n, device=c.dev)
        extra["slab_gb_per_gpu"] = cnt * eng.P * 4 / 1e9

        def step(i):
            a, b = eng.predict_stored_sums(x_dev, cnt, j0=0, wts=freq, mode=args.mode, second_moment=True)
            if world > 1:
                dd.allreduce_sum_(a, b)
            outs[:] = [a, b]
    elif cfg == 5:
        lab_dev = labels

        def step(i):
            cts = eng.count_predicate(x_dev, lab_dev, cnt, seed=5, s0=i * n + off, mode=args.mode)
            if world > 1:
                c.dist.all_reduce(cts)
            outs[:] = [cts]
    else:
        s1 = torch.empty((B, 10), dtype=torch.float32, device=c.dev)
        s2 = torch.empty_like(s1)

        def step(i):
            eng.predict_sums(x_dev, cnt, seed=1234, s0=i * n + off, mode=args.mode, second_moment=True, out=(s1, s2))
            if world > 1:
                dd.allreduce_sum_(s1, s2)
            outs[:] = [s1, s2]

    step(0)
    _barrier(c)
    path = eng.last_path()
    # warm-up / pre-heat without the profiler, then the timed steps with it
    steps = args.steps
    clocks = ClockSampler(torch.cuda.current_device()); clocks.start()
    ms_w, _, _, pre, nxt = timed_steps(c, step, 1, warm, args.preheat, first_index=1)
    launches0 = _lib.launch_count()
    lib.dbx_profile(1)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    _barrier(c)
    clocks.mark()
    for k in range(steps):
        evs[k].record()
        step(nxt + k)
    evs[steps].record()
    _barrier(c)
    clocks.end()
    ms = _max_over_ranks(c, [evs[0].elapsed_time(evs[steps]) / steps])[0]
    pms, pfl, pn = _profile_read(lib)
    lib.dbx_profile(0)
    launches = _lib.launch_count() - launches0
    clk = clocks.stop()

    # ---- e2e: host arrays in, host arrays out, through the public call
    x_host = torch.from_numpy(x_np).pin_memory().numpy()
    if cfg in (1,):
        def e2e(i):
            return pm.predict_moments(x_host, n)
        d2h = 2 * B * 10 * 4
    elif cfg == 5:
        from deepbayes_b200.analyzers import verifiers

        def e2e(i):
            return verifiers.statistical_robustness(pm, x_host, labels, n=n)
        d2h = B * 8
    elif cfg == 3:
        def e2e(i):
            a, b = eng.predict_stored_sums(x_host, cnt, j0=0, wts=freq, mode=args.mode, second_moment=True)
            if world > 1:
                dd.allreduce_sum_(a, b)
            return a.cpu().numpy(), b.cpu().numpy()
        d2h = 2 * B * 10 * 4
    else:
        def e2e(i):
            a, b = eng.predict_sums(x_host, cnt, seed=99, s0=i * n + off, mode=args.mode, second_moment=True)
            if world > 1:
                dd.allreduce_sum_(a, b)
            return (a / float(n)).cpu().numpy(), (b / float(n)).cpu().numpy()
        d2h = 2 * B * 10 * 4
    e2e(0)
    _barrier(c)
    k_e = max(2, min(steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(k_e):
        e2e(1 + k)
    e1.record()
    _barrier(c)
    ms_e2e = _max_over_ranks(c, [e0.elapsed_time(e1) / k_e])[0]

    if rank == 0:
        if bound == "hbm":
            alg = 4.0 * eng.P * cnt           # bytes of stored fp32 samples one step streams on this GPU
            ach = alg / (pms / steps * 1e-3) / 1e9 if pms > 0 else None
            peak = float(peaks["hbm_gbs"])
            roof = {"bound": "hbm", "kernel": "dense_tc_f32w_kernel chain (fp32 slab by TMA, bf16 conversion in the SM)" if path == "generic_tc" else path,
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if ach else None, "traffic": None,
                    "algorithmic_bytes_per_step": alg, "algorithmic_bytes_per_sample": 4 * eng.P,
                    "kernel_ms_per_step": pms / steps, "kernel_launches_per_step": pn / steps, "peak_source": peak_src,
                    "whole_step_gbs": alg / (ms * 1e-3) / 1e9, "whole_step_frac": alg / (ms * 1e-3) / 1e9 / peak}
        else:
            ach = (pfl / (pms * 1e-3)) / 1e12 if pms > 0 else None
            peak = float(peaks["bf16_tflops_sustained"] if pre > 0 else peaks["bf16_tflops"])
            whole = (units / world * eng.flops_per_forward / (ms * 1e-3)) / 1e12
            roof = {"bound": "tensor", "kernel": "GEMM-carrying kernels of path %s (library CUDA events)" % path, "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak if ach else None, "traffic": None, "peak_burst": float(peaks["bf16_tflops"]),
                    "kernel_ms_per_step": pms / steps, "kernel_launches_per_step": pn / steps, "peak_source": peak_src,
                    "algorithmic_flops_per_forward": eng.flops_per_forward, "whole_step_tflops": whole, "whole_step_frac": whole / peak,
                    "whole_step_frac_of_burst": whole / float(peaks["bf16_tflops"])}
        line = {
            "metric": METRIC, "value": units / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.mode if args.mode != "fp32" else "f32", "data": "synthetic",
            "config": dict({"workload": WORKLOADS[cfg], "batch": B, "samples": n, "samples_per_gpu": cnt, "path": path, "preheat_steps": pre,
                            "l2": "inputs larger than L2 every step (sampled weights / stored slab)",
                            "collective": dd.collective_kind() if world > 1 else "none (1 GPU)"}, **extra),
            "roofline": roof, "cpu_baseline": None,
            "e2e": {"value": units / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(x_np.nbytes) * world,
                    "d2h_bytes_per_step": int(d2h) * world},
            "gpu_launches": int(launches), "clocks": clk,
        }
        if cfg in (3, 5):
            line["samples_per_s"] = n / (ms * 1e-3)
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--batch", type=int, default=B_ROWS)
    ap.add_argument("--samples", type=int, default=N_SAMPLES)
    ap.add_argument("--stored", type=int, default=10000, help="config 3: stored posterior samples (10,000 x 7.45 MB = 74.5 GB)")
    ap.add_argument("--preheat", type=float, default=2.0, help="seconds of untimed steps before the timed region (sustained regime)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    c = _setup_dist()
    try:
        if args.config == 2:
            run_config2(args, c)
        else:
            run_other(args, c)
    finally:
        if c.world > 1:
            from deepbayes_b200 import distributed as dd
            try:
                c.dist.barrier()
            except Exception:
                pass
            dd.shutdown()
            c.dist.destroy_process_group()


if __name__ == "__main__":
    main()
