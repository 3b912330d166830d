#!/usr/bin/env python
"""bench.py -- posterior-sample x input forwards / second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full Monte-Carlo posterior predictive pass of BASELINE config 2 on each GPU:
BayesByBackprop MLP 784-512-512-10, B = 4096 input rows, n = 1024 posterior samples PER GPU
(weak scaling: every rank evaluates its own 1024 global sample ids; one NCCL all-reduce of the
[B,C] moment buffers per step when N > 1).  Inside a step: in-kernel Philox sampling of all
669,706 weights per sample, three dense layers, softmax, accumulation of sum p and sum p^2.
value = N * n * B forwards / (max-over-ranks device time per step).
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

DIMS = [784, 512, 512, 10]
B_ROWS = 4096
N_SAMPLES = 1024
METRIC = "posterior-sample x input forwards/sec"
UNIT = "forwards/s"


def _ncu_traffic():
    """DRAM bytes (read + write) of ONE fused-kernel launch from the committed ncu --set full capture."""
    f = os.path.join(ROOT, "profiles", "r1_ncu_fused_pair_summary.txt")
    try:
        tot = 0.0
        for line in open(f):
            parts = line.split()
            if parts and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(parts[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[parts[2]]
        return tot or None
    except Exception:
        return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

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
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sm[len(sm) // 4:] or sm   # skip the ramp at the start of the timed region
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_pass(n_samples, B, threads=None):
    """The reference's CPU loop (posteriormodel.py:94-101) restated by the oracle: per sample
    np.random.normal over every tensor (float64, single-threaded MT19937, as in the reference),
    cast to fp32, fp32 forward on all host cores, fp32 running sum, / float(n)."""
    import torch
    from oracle import oracle as o
    if threads:
        torch.set_num_threads(threads)
    L = o.mlp(DIMS)
    mean, std = o.synthetic_posterior(L, seed=1)
    x = o.synthetic_x((B, DIMS[0]), seed=0)
    np.random.seed(0)
    t0 = time.perf_counter()
    out = o.predict(L, mean, std, x, n_samples, eps=None)
    dt = time.perf_counter() - t0
    assert np.isfinite(out).all()
    return dt


def run_reference(args):
    """--impl reference: the reference's own CPU path (oracle port; TensorFlow cannot be
    installed on this image) on the host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_ref = int(os.environ.get("DBX_REF_SAMPLES", "8"))
    for _ in range(args.warmup):
        cpu_reference_pass(1, B_ROWS)
    ts = [cpu_reference_pass(n_ref, B_ROWS) for _ in range(args.steps)]
    dt = float(np.mean(ts))
    val = n_ref * B_ROWS / dt
    sample = "%d of %d posterior samples x all %d rows per step" % (n_ref, N_SAMPLES, B_ROWS)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BayesByBackprop MLP 784-512-512-10, batch 4096, n=1024 posterior samples (BASELINE config 2)",
                   "reference_arm": "loop-faithful NumPy port of posteriormodel.py:81-101 (TensorFlow not installable here)",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=B_ROWS)
    ap.add_argument("--samples", type=int, default=N_SAMPLES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import deepbayes_b200 as db
    from deepbayes_b200 import _lib, formats
    from deepbayes_b200 import distributed as dd
    from oracle import oracle as o  # synthetic-input generators + the cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    B, n = args.batch, args.samples
    warm = max(args.warmup, 3)

    # ---- the user-facing objects: a saved synthetic posterior reloaded as PosteriorModel
    L = o.mlp(DIMS)
    mean, std = o.synthetic_posterior(L, seed=1)
    model = db.Sequential([db.Dense(512, activation="relu", input_shape=(784,)), db.Dense(512, activation="relu"),
                           db.Dense(10, activation="softmax")])
    pdir = "/tmp/dbx_bench_posterior_r%d" % rank
    formats.save_gaussian_posterior(pdir, model, mean, std, {"input_upper": 1.0, "input_lower": 0.0})
    pm = db.PosteriorModel(pdir, mode=args.mode, seed=1234, shard=False)
    eng = pm._engine
    x_host = o.synthetic_x((B, DIMS[0]), seed=0)
    x_dev = torch.from_numpy(x_host).to(dev)
    x_pinned = torch.from_numpy(x_host).pin_memory()   # the caller's input buffer of the e2e leg
    s1 = torch.empty((B, 10), dtype=torch.float32, device=dev)
    s2 = torch.empty_like(s1)
    lib = _lib.load()

    def step_resident(i):
        # global sample ids: rank r owns [r*n, (r+1)*n) of this step's N*n draws
        eng.predict_sums(x_dev, n, seed=1234, s0=(i * world + rank) * n, mode=args.mode, second_moment=True, out=(s1, s2))
        if world > 1:
            dd.allreduce_sum_(s1, s2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warm):
        step_resident(i)
    barrier()
    path = eng.last_path()

    # ---- timed region 1: inputs resident in HBM
    clocks = ClockSampler(torch.cuda.current_device())
    clocks.start()
    launches0 = _lib.launch_count()
    lib.dbx_profile(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step_resident(warm + i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    import ctypes as C
    pms, pfl, pn = C.c_double(), C.c_double(), C.c_int64()
    lib.dbx_profile_read(C.byref(pms), C.byref(pfl), C.byref(pn))
    lib.dbx_profile(0)
    launches = _lib.launch_count() - launches0
    clk = clocks.stop()

    # ---- timed region 2: end to end through PosteriorModel.predict-style host call
    # (pinned H2D of x every step, device loop, all-reduce, /n, D2H of the [B,C] mean)
    def step_e2e(i):
        xd = eng._dev_f32(x_pinned)  # pinned host memory -> async H2D on the compute stream, every step
        a, _ = eng.predict_sums(xd, n, seed=99, s0=(i * world + rank) * n, mode=args.mode, out=(s1, s2))
        if world > 1:
            dd.allreduce_sum_(a)
        return (a / float(n * world)).cpu().numpy()

    for i in range(2):
        step_e2e(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = step_e2e(2 + i)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    assert out.shape == (B, 10) and np.isfinite(out).all() and np.allclose(out.sum(-1), 1, rtol=1e-3)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks, peak_src = _peaks()
        flops_fwd = eng.flops_per_forward
        forwards = world * n * B
        value = forwards / (ms * 1e-3)
        tf_achieved = (pfl.value / (pms.value * 1e-3)) / 1e12 if pms.value > 0 else None
        peak_tf = float(peaks["bf16_tflops_sustained"]) if args.mode == "bf16" else None
        roof = {
            "bound": "tensor",
            "kernel": "fused_mlp2_kernel (tcgen05 cta_group::2, TMA, TMEM-resident activations)" if path == "fused_tcgen05" else "dense_kernel (generic FFMA)",
            "achieved": tf_achieved, "peak": peak_tf, "peak_burst": float(peaks["bf16_tflops"]), "unit": "TFLOP/s",
            "frac": (tf_achieved / peak_tf) if (tf_achieved and peak_tf) else None,
            "frac_of_burst": (tf_achieved / float(peaks["bf16_tflops"])) if tf_achieved else None,
            "peak_source": peak_src + " (sustained figure: kernel timed inside a long step)",
            "traffic": _ncu_traffic() if path == "fused_tcgen05" else None,
            "traffic_note": "DRAM read+write bytes of one launch (256 samples x 4096 rows) from profiles/r1_ncu_fused_pair_summary.txt; "
                            "algorithmic bytes of that launch: 256 * 1.34 MB sampled bf16 weights + 6.4 MB inputs",
            "kernel_ms_per_step": pms.value / args.steps, "kernel_launches_per_step": pn.value / args.steps,
            "algorithmic_flops_per_forward": flops_fwd,
            "whole_step_frac_of_burst_peak": (forwards * flops_fwd / (ms * 1e-3)) / 1e12 / float(peaks["bf16_tflops"]),
        }
        cpu = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cpu_reference_pass(1, B)  # warm the BLAS threads
            t1 = cpu_reference_pass(2, B) / 2
            n_ref = int(max(2, min(1024, round(15.0 / max(t1, 1e-3)))))
            dt = cpu_reference_pass(n_ref, B)
            cpu = {"value": n_ref * B / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d of %d posterior samples x all %d rows (%.1f s); NumPy port of posteriormodel.py:81-101, "
                             "single-threaded np.random.normal + multi-threaded fp32 BLAS" % (n_ref, n, B, dt)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.mode if args.mode != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": "BayesByBackprop MLP 784-512-512-10, batch %d, n=%d posterior samples per GPU "
                                   "(BASELINE config 2)" % (B, n),
                       "global_samples_per_step": world * n, "path": path,
                       "l2": "no explicit flush: each step streams %.2f GB of freshly sampled weights (> 126 MB L2)" % (n * eng.P * 2 / 1e9),
                       "parallelism": "sample-sharded x%d, one all-reduce of [B,C] moments per step" % world},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": forwards / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(x_host.nbytes), "d2h_bytes_per_step": int(B * 10 * 4)},
            "gpu_launches": int(launches), "clocks": clk,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
