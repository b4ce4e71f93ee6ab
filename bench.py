#!/usr/bin/env python
"""bench.py -- mode-solves/sec on the BASELINE.json headline workload (512x512, 4 modes, 256 freqs).

A "step" is one pass of the hot path over one batch of `--freqs-per-step` (default 64) frequency points (a contiguous
slice of the 256-point sweep C_0/linspace(1.5,1.6,256)); the sweep is sharded contiguously over ranks
(weak scaling: every GPU gets `--freqs-per-step` problems per step, no data-path collective; the only
collective is the final gather of n_complex).  `value` = mode-solves/s with the cross-section already resident
in HBM is not separable in this API (the C ABI takes host buffers), so `value` counts the device-timed solve
(CUDA events inside the library, max over ranks) and `e2e` the wall clock of the public call with HOST
buffers in and the (fields, n_complex) results back in host memory.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
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


def _clock_sampler(stop, samples):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", os.environ.get("LOCAL_RANK", "0")],
                                 capture_output=True, text=True, timeout=5).stdout.strip().splitlines()
            if out:
                samples.append([s.strip() for s in out[0].split(",")])
        except Exception:  # noqa: BLE001
            pass
        stop.wait(0.2)


def _clock_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    sm = sorted(float(s[0]) for s in samples if s[0].replace(".", "").isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(samples[0][1]) if samples[0][1].replace(".", "").isdigit() else None,
            "reasons": reasons, "samples": len(samples)}


def bench_config(args, world):
    """The `config` object shared by both arms (the driver compares them)."""
    return {"workload": f"headline Si strip {args.n}x{args.n}, num_modes=4, {args.freqs_per_step} freqs/step/GPU of the 256-pt sweep 1.5-1.6um",
            "l2": "inputs larger than L2 (per-step working set > 10 GB)", "parallelism": f"freq-shard x{world}"}


def reference_arm(args, rank, world):
    """CPU arm: the reference algorithm (oracle restatement == the reference's scipy ARPACK + SuperLU call,
    solver.py:744; the reference is pure Python and not installable here) on all host cores, on a bounded
    sample of the same workload.  One 512x512 solve costs 60-100 s of one core, so the K steps together are ONE round of
    `workers` concurrent solves (evenly spaced frequencies of the sweep, ceil(workers/K) per step)."""
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor

    cores = os.cpu_count() or 1
    # Each worker holds a SuperLU factorisation of the 524288 x 524288 operator (several GB); an unbounded pool of 128
    # workers took the GPU box down once (cgroup memory limit 200 GiB), so the pool is bounded by the container's memory
    # limit at 4 GB per worker, by its CPU quota and by 16 workers.
    limit = None
    for path in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(path).read().strip()
            if v.isdigit():
                limit = int(v)
                break
        except OSError:
            pass
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:  # noqa: BLE001
        avail = 64 << 30
    budget = min(avail, limit) if limit else avail
    mem_cap = max(1, int(0.6 * budget / (4 << 30)))  # measured peak RSS of one 512x512 solve: 2.5 GB
    try:  # CPU quota of the container (cgroup v2 cpu.max = "quota period"); the B200 box: 16 CPUs of 128, 200 GiB
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            cores = max(1, min(cores, int(int(q) / int(per))))
    except (OSError, ValueError):
        pass
    workers = max(1, min(cores, args.ref_workers or cores, mem_cap, 16))
    idx = np.linspace(0, 255, workers).round().astype(int)
    from tidy3d_b200 import workloads as W

    freqs = W.sweep_freqs(256)
    t0 = time.time()
    with ProcessPoolExecutor(workers) as ex:
        list(ex.map(_ref_freq, [(args.n, float(freqs[i])) for i in idx]))
    dt = time.time() - t0
    val = len(idx) / dt
    steps = max(1, args.steps)
    line = {
        "impl": "reference", "metric": "mode-solves/sec (512x512, 4 modes, 256 freqs)", "value": val, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 0, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, world),
        "cpu_baseline": {"value": val, "unit": "solves/s", "cores": workers, "kind": "port",
                         "sample": f"{len(idx)} of the 256 frequencies in one concurrent round ({-(-len(idx) // steps)} per step), one "
                                   f"per worker process, scipy eigs (ARPACK+SuperLU, tol=fp_eps) via oracle/restatement.py, {dt:.1f} s wall"},
        "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _ref_freq(a):
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import restatement as R
    from tidy3d_b200 import workloads as W

    n, freq = a
    wl = W.headline(nf=1, n=n)
    _, nc, _ = R.compute_modes(wl.eps_cross, wl.coords, freq, wl.mode_spec)
    return nc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--freqs-per-step", type=int, default=64)
    ap.add_argument("--ref-workers", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stencil-only", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g

    if rank == 0:
        g.build()  # no-op when the in-tree library is up to date; never let N ranks rebuild the same file at once
    if world > 1:
        dist.barrier()
    import ctypes as C

    from tidy3d_b200 import _cabi
    from tidy3d_b200 import workloads as W
    from tidy3d_b200.solver import compute_modes_batch, get_handle

    h = get_handle(local)
    h.set_options(max_batch=max(32, args.freqs_per_step))
    wl = W.headline(nf=256, n=args.n)
    fps = args.freqs_per_step
    # contiguous shard of the sweep per rank (independent problems, no exchange)
    per_rank = 256 // world
    my = wl.freqs[rank * per_rank : (rank + 1) * per_rank]

    def step_problems(s):
        fr = [my[(s * fps + i) % len(my)] for i in range(fps)]
        return [dict(eps_cross=wl.eps_cross, coords=wl.coords, freq=f, mode_spec=wl.mode_spec) for f in fr]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- roofline of the dominant kernel (fused stencil), measured live with CUDA events on the library stream
    pk = _cabi.PackedProblem(wl.eps_cross, wl.coords, wl.freqs[0], wl.mode_spec)
    ms, byts = C.c_double(), C.c_double()
    roof = {}
    for mode, nm in ((0, "apply"), (1, "jacobi")):
        rc = _cabi.lib().b200ms_bench_stencil(h._h, C.byref(pk.struct), fps, mode, 50, 0, None, None, C.byref(ms), C.byref(byts))
        assert rc == 0, h.last_error()
        roof[nm] = (byts.value / (ms.value * 1e-3) / 1e9, ms.value, byts.value)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = roof["apply"][0]
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                # dram read+write bytes per launch from one ncu --set full capture (profiles/r01_stencil_apply_b64_ncu.txt:
                # 955.3 MB at 64 x 512^2 = 56.94 B per cell), scaled to this launch's cell count
                "traffic": 56.94 * fps * args.n * args.n,
               
                "kernel": "stencil_march_kernel<double,double,MODE_APPLY> (fp64 operator apply y=(A-sigma)x, 56 B/cell)",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "apply_ms": roof["apply"][1], "bytes_per_launch": roof["apply"][2],
                "smoother_kernel": "stencil_march_kernel<float,float,MODE_JACOBI_D> (fp32 stored-diagonal sweep, 44 B/cell)",
                "smoother_GBps": roof["jacobi"][0], "smoother_ms": roof["jacobi"][1], "smoother_bytes_per_launch": roof["jacobi"][2],
                "working_set": f"{fps} problems x {args.n}^2: vectors + coefficient fields >> 126 MB L2"}
    if args.stencil_only:
        print(json.dumps(roofline))
        return

    for s in range(args.warmup):
        compute_modes_batch(step_problems(s), handle=h)  # results dropped immediately
    stop, samples = threading.Event(), []
    th = threading.Thread(target=_clock_sampler, args=(stop, samples), daemon=True)
    th.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, lib_ms, launches, results, h2d, d2h = 0.0, 0.0, 0, [], 0, 0
    first_step_n = None
    out = None
    for s in range(args.steps):
        out = None  # drop the previous step's results: their pinned buffers return to the pool
        out, info = compute_modes_batch(step_problems(args.warmup + s), handle=h, return_info=True)
        dev_ms += info[0]["solve_ms"]
        lib_ms += info[0]["total_ms"]
        launches += int(info[0]["stencil_applies"])
        results.append(np.array([o[1] for o in out]))
        if s == 0:
            first_step_n = results[0].copy()
            first_step_freqs = [p["freq"] for p in step_problems(args.warmup)]
        d2h = sum(o[0].nbytes + o[1].nbytes for o in out)
        h2d = 9 * 16 * args.n * args.n  # the shared cross-section is uploaded once per step
    barrier()
    wall = time.perf_counter() - t0
    stop.set()
    th.join(timeout=2)
    t = torch.tensor([dev_ms * 1e-3, wall], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the only collective of the path: gather n_complex of every rank's shard (NCCL over NVLink)
        mine = torch.from_numpy(np.concatenate(results).view(np.float64)).cuda()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
    dev_s, wall_s = float(t[0]), float(t[1])
    nsolves = args.steps * fps * world
    value = nsolves / dev_s
    e2e = nsolves / wall_s
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            # reference algorithm (scipy ARPACK + SuperLU through the oracle restatement) on the host cores, on a
            # bounded sample of the same workload: one frequency of the sweep per worker process
            from concurrent.futures import ProcessPoolExecutor

            cores = os.cpu_count() or 1
            try:
                q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
                if q != "max":
                    cores = max(1, min(cores, int(int(q) / int(per))))
            except (OSError, ValueError):
                pass
            workers = max(1, min(cores, 8))
            idx = np.linspace(0, fps - 1, workers).round().astype(int)
            tc = time.time()
            with ProcessPoolExecutor(workers) as ex:
                ncpu = list(ex.map(_ref_freq, [(args.n, float(first_step_freqs[i])) for i in idx]))
            tcpu = time.time() - tc
            dn = max(float(np.abs(np.asarray(ncpu[j]) - first_step_n[i]).max()) for j, i in enumerate(idx))
            cpu = {"value": workers / tcpu, "unit": "solves/s", "cores": workers, "kind": "port",
                   "sample": f"{workers} of the 256 frequencies ({args.n}x{args.n}, 4 modes), one per worker process, scipy eigs "
                             f"(ARPACK+SuperLU) via oracle/restatement.py: {tcpu:.1f} s wall",
                   "host_cores_total": cores, "max_abs_dn_gpu_vs_cpu": dn}
        line = {
            "metric": "mode-solves/sec (512x512, 4 modes, 256 freqs)", "value": value, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args, world),
            "e2e": {"value": e2e, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "lib_wall_ms_per_step": lib_ms / args.steps, "roofline": roofline, "cpu_baseline": cpu,
            "clocks": _clock_summary(samples),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
