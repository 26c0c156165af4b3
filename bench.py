#!/usr/bin/env python
"""bench.py — solve DOFs/s of the AMG-preconditioned CG solve (to 1e-8) on synthetic structured tet cubes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cube N]

Workload: Kuhn tet cube with `--cube` cells per side — on one GPU BASELINE.json configs[2] (default 118:
1 685 159 DOF, 9 858 192 tets), on 2/4/8 GPUs configs[3] (default 255: 16 777 216 DOF, 99 488 250 tets;
the line then also carries rank 0's single-GPU solve time of that same mesh) — operator K + M assembled on the
GPU, AMG hierarchy built on the GPU (aggregation seed 0), right-hand side b = A x*, x* the egg-carton
sin(2 pi x) sin(2 pi y) sin(2 pi z), initial guess 0, FEMSolver defaults + solverType_=1 (PCG),
tolerance_=1e-8, maxIters_=200.  One STEP = one complete PCG solve from the initial residual to
||r||/||b|| <= 1e-8.

  value   n / t_solve with b, x resident in HBM (fsb_solve_device), CUDA events on the solver's stream
  e2e     same metric through the host-buffer C-ABI call fsb_solve (pinned host b/x, H2D + D2H inside)
  roofline  the dominant kernel of the step (by CUDA-event time inside a profiled solve), algorithmic
            bytes per launch (DESIGN.md "algorithmic bytes") / its average launch duration
  cpu_baseline  the CPU oracle (a restatement of the reference: its CUDA build cannot be produced here)
                on this box's host cores, same workload, one full solve

`--impl reference` times that CPU oracle alone (all host threads) and prints the same line shape.
Under torchrun (N > 1) one cube is solved by all N GPUs together (strong scaling): setup is
replicated on every GPU, the fine level of the solve is sharded in contiguous partition ranges, halo
values and dot products travel over NVLink peer memory.
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

METRIC = "solve DOFs/sec (AMG-PCG to 1e-8)"
UNIT = "DOF/s"


# stdout carries exactly ONE JSON line: keep a private copy of fd 1 for it and point fd 1 at stderr so
# that library chatter (e.g. "NCCL version ..." under NCCL_DEBUG) cannot land in front of the line.
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_name(N):
    return f"Kuhn tet cube N={N} ({(N + 1) ** 3} DOF, {6 * N ** 3} tets), K+M, b=A*eggcarton, PCG+AMG V-cycle to 1e-8"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_problem(N):
    import sci_solver_fem_b200 as fsb
    verts, tets = fsb.meshio.kuhn_cube(N)
    xstar = np.sin(2 * np.pi * verts[:, 0]) * np.sin(2 * np.pi * verts[:, 1]) * np.sin(2 * np.pi * verts[:, 2])
    return verts, tets, xstar


# ----------------------------------------------------------------------------- algorithmic bytes
def algorithmic_bytes(kernel, n, nnz, nnzP=0, nc=0, nnz_in=None):
    """DESIGN.md 'algorithmic bytes per launch': fp64 values, int32 indices, every array once.
    The smoother stages touch only the intra-partition block A_in (nnz_in entries incl. the diagonal)."""
    if nnz_in is None:
        nnz_in = nnz
    if kernel in ("spmv", "spmv_dot"):
        return 12 * nnz + 4 * (n + 1) + 8 * n + 8 * n          # A, x gather once, y
    if kernel == "residual":
        return 12 * nnz + 4 * (n + 1) + 8 * n * 3              # A, x, b, r
    if kernel == "pre_smooth":
        return 12 * nnz_in + 4 * (n + 1) + 8 * n * 3           # A_in once (nu sweeps partition-resident), b, x out, r out
    if kernel == "post_smooth":
        return 12 * nnz_in + 4 * (n + 1) + 8 * n * 3           # A_in once, b', x in, x out
    if kernel == "restrict":
        return 12 * nnzP + 4 * (nc + 1) + 8 * n + 8 * nc
    if kernel == "prolong_add":
        return 12 * nnzP + 4 * (n + 1) + 8 * nc + 16 * n
    if kernel == "cg_update":
        return 48 * n
    if kernel == "dot":
        return 16 * n
    if kernel == "cg_pdir":
        return 24 * n
    return 0


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def run_oracle(N, steps, warmup, threads=None, single_thread_solve=False):
    from oracle import oracle as orc
    if threads:
        orc.set_threads(threads)
    verts, tets, xstar = build_problem(N)
    o = orc.Oracle(64, solverType=1, tolerance=1e-8, maxIters=200, seed=0)
    t0 = time.perf_counter(); o.pattern(len(verts), tets); t_pat = time.perf_counter() - t0
    t0 = time.perf_counter(); o.assemble(verts); t_asm = time.perf_counter() - t0
    t0 = time.perf_counter(); o.setup(); t_setup = time.perf_counter() - t0
    b = o.spmv(xstar)
    times, iters = [], 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        x, iters = o.solve(b)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    err = float(np.linalg.norm(x - xstar) / np.linalg.norm(xstar))
    nthreads = orc.lib().orc_max_threads()
    t_single = None
    if single_thread_solve:  # the same solve on ONE host thread (closest to the reference's serial host code)
        try:
            orc.set_threads(1)
            t0 = time.perf_counter(); o.solve(b); t_single = time.perf_counter() - t0
        except Exception:
            t_single = None
        finally:
            orc.set_threads(nthreads)
    return dict(n=len(verts), t_solve=float(np.mean(times)), iters=int(iters), relres=float(o.final_relres()), err=err,
                t_pattern=t_pat, t_assemble=t_asm, t_setup=t_setup, threads=nthreads, t_solve_1thread=t_single)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N = args.cube
    steps, warmup = args.steps, args.warmup
    if N > 160:  # bounded sample of the ~100M-tet workload: one complete solve (about 1.5 min of CPU work with the setup)
        steps, warmup = 1, 0
    r = run_oracle(N, steps, warmup)
    value = r["n"] / r["t_solve"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": r["t_solve"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(N), "iterations": r["iters"], "relres": r["relres"], "rel_l2_err_vs_exact": r["err"],
                   "setup_s": r["t_setup"], "assemble_s": r["t_assemble"], "pattern_s": r["t_pattern"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": "port",
                         "sample": f"full workload, {steps} complete PCG solve(s) (setup excluded), CPU oracle = restatement of the reference "
                                   "(its CUDA-only build needs CUSP + METIS 4 at configure time; not producible offline)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------- our arm
def ours(args):
    import torch
    import torch.distributed as dist
    import sci_solver_fem_b200 as fsb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = args.cube
    verts, tets, xstar = build_problem(N)
    n = len(verts)

    s = fsb.FEMSolver.from_arrays(verts, tets, None, device=local)
    s.solverType_, s.tolerance_, s.maxIters_, s.seed_ = 1, 1e-8, 200, 0
    t_pattern_cold = s.time_ms("pattern")   # first GPU work of the process: includes pool growth / module load
    t_pattern = t_assemble = t_setup = float("inf")
    rebuilds = 3 if N <= 160 else 1
    for _ in range(rebuilds):                # steady-state stage timings: best of 3 rebuilds (allocator pool warm)
        s.getMatrixFromMesh()
        t_pattern, t_assemble = min(t_pattern, s.time_ms("pattern")), min(t_assemble, s.time_ms("assemble"))
    for _ in range(rebuilds):
        s.setup()
        t_setup = min(t_setup, s.time_ms("setup"))
    nnz = s._L.fsb_matrix_nnz(s.handle)
    levels = [(s.level_rows(l), s.level_nnz(l)) for l in range(s.num_levels())]
    log(f"[rank {rank}] n={n} nnz={nnz} levels={levels} pattern {t_pattern:.1f} ms assemble {t_assemble:.1f} ms setup {t_setup:.1f} ms")

    # b = A x* on the device (outside every timed region)
    xs_dev = torch.from_numpy(xstar).cuda()
    b_dev = torch.empty(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    s.apply_matrix_device(xs_dev.data_ptr(), b_dev.data_ptr())
    b_host = b_dev.cpu().pin_memory()
    x_host = torch.zeros(n, dtype=torch.float64).pin_memory()
    x_dev = torch.zeros(n, dtype=torch.float64, device="cuda")
    stream = torch.cuda.ExternalStream(s._L.fsb_stream(s.handle))
    torch.cuda.synchronize()
    ms_single = None
    if world > 1:
        # the hierarchy is replicated, so every GPU can also solve the whole system alone: rank 0 times that
        # (same mesh, same build, same run) as the strong-scaling reference of this line
        if rank == 0:
            for i in range(2):
                with torch.cuda.stream(stream):
                    x_dev.zero_()
                s.solve_device(x_dev.data_ptr(), b_dev.data_ptr())
            ms_single = s.time_ms("solve")
        dist.barrier()
        s.dist_connect(rank, world, fsb.exchange_handles_torch)
        pb, rb, ab = s.dist_ranges()
        log(f"[rank {rank}] owns partitions [{pb[rank]},{pb[rank+1]}) rows [{rb[rank]},{rb[rank+1]})")
        dist.barrier()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def solve_device():
        with torch.cuda.stream(stream):
            x_dev.zero_()
        s.solve_device(x_dev.data_ptr(), b_dev.data_ptr())

    def solve_host():
        x_host.zero_()
        s.solve(x_host.numpy(), b_host.numpy())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms / steps, wall * 1e3 / steps

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, wall_dev = timed(solve_device, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    iters, relres, launches = s.iterations, s.relres, s.last_launches()
    xg = x_dev.cpu().numpy()
    err = float(np.linalg.norm(xg - xstar) / np.linalg.norm(xstar))
    ms_e2e, wall_e2e = timed(solve_host, max(1, args.steps), min(args.warmup, 3))
    # solveFEM-equivalent (hierarchy rebuilt + solve, host buffers), for the record (single GPU only:
    # a rebuild ends the sharded mode)
    t_solvefem = float("nan")
    if world == 1:
        t_solvefem = float("inf")
        for _ in range(2):  # steady state, like the stage timings above (the first rebuild after a solve re-grows pools)
            t0 = time.perf_counter(); x_host.zero_(); s.solveFEM(x_host.numpy(), b_host.numpy()); t_solvefem = min(t_solvefem, time.perf_counter() - t0)

    # roofline pass: one profiled solve (CUDA events around every kernel; graphs off)
    s.profile_ = 1
    solve_device()
    prof = s.profile_report()
    s.profile_ = 0
    tot = sum(ms for (_, ms) in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1][1])
    (kname, klev), (kcnt, kms) = top
    ln, lnnz = levels[klev]
    nnzP = nc = 0
    if klev + 1 < len(levels):
        nc = levels[klev + 1][0]
        nnzP = int(s._L.fsb_level_int(s.handle, klev, b"P_col", None, 0))
    kname_alg = "restrict" if (kname == "spmv" and klev < len(levels) - 1) else kname
    own = 1.0
    if world > 1 and klev == 0:  # the fine level is sharded: a launch touches this GPU's rows only
        own = float(rb[rank + 1] - rb[rank]) / n
    def nnz_in_of(lev):  # entries inside the partitions' diagonal blocks = nnz - nnz(A_out)
        if lev >= len(levels) - 1:
            return levels[lev][1]
        ap = s.level_int(lev, "Aout_ptr")
        return levels[lev][1] - int(ap[-1])
    bytes_per_launch = algorithmic_bytes(kname_alg, ln, lnnz, nnzP, nc, nnz_in_of(klev)) * own
    peak, peak_src = measured_peaks()
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of the same workload
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if tj.get("cube") == N and world == 1:
            traffic = tj.get("kernels", {}).get(f"{kname}@level{klev}")
    except Exception:
        pass
    achieved = bytes_per_launch / (kms / kcnt * 1e-3) / 1e9
    kernels = {f"{k}@L{l}": {"launches": c, "ms": round(ms, 4), "share": round(ms / tot, 4)} for (k, l), (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}
    # per-kernel achieved GB/s on the fine level (SpMV + smoother % of HBM peak is part of the metric)
    fine = {}
    nnz_in0 = nnz_in_of(0)
    for k in ("spmv_dot", "pre_smooth", "residual", "post_smooth", "cg_update", "dot", "cg_pdir"):
        if (k, 0) in prof:
            c, ms = prof[(k, 0)]
            gbs = algorithmic_bytes(k, levels[0][0], levels[0][1], nnz_in=nnz_in0) * (own if world > 1 else 1.0) / (ms / c * 1e-3) / 1e9
            fine[k] = {"us": round(ms / c * 1e3, 2), "GBps": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = run_oracle(N, 1, 0, single_thread_solve=True)
            cpu = {"value": r["n"] / r["t_solve"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                   "single_thread": ({"value": r["n"] / r["t_solve_1thread"], "unit": UNIT, "solve_s": r["t_solve_1thread"]}
                                     if r.get("t_solve_1thread") else None),
                   "sample": f"full workload, 1 complete PCG solve ({r['iters']} iterations, {r['t_solve']:.2f} s; setup {r['t_setup']:.1f} s excluded); "
                             "CPU oracle = restatement of the reference (no host solve and no offline CUDA build upstream)",
                   "iterations": r["iters"], "setup_s": r["t_setup"], "assemble_s": r["t_assemble"], "pattern_s": r["t_pattern"]}
        except Exception as e:  # the bench line must still be printed
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"oracle failed: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": n / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(N), "parallelism": "single GPU" if world == 1 else f"{world} GPUs: replicated setup, fine level sharded in contiguous partition ranges, NVLink peer-memory halo pushes + in-kernel all-reduce",
                       "l2_policy": "inputs larger than L2 (hierarchy + vectors ~%.0f MB per solve pass)" % ((12 * nnz + 40 * n) / 1e6),
                       "iterations": iters, "relres": relres, "rel_l2_err_vs_exact": err, "levels": levels,
                       "pattern_ms": t_pattern, "pattern_cold_ms": t_pattern_cold, "assemble_ms": t_assemble, "setup_ms": t_setup, "solveFEM_host_ms": None if t_solvefem != t_solvefem else t_solvefem * 1e3,
                       "wall_ms_per_step": wall_dev,
                       "single_gpu_same_mesh_ms": ms_single,
                       "speedup_vs_single_gpu_same_mesh": (ms_single / ms_dev) if ms_single else None,
                       "dofs_per_s_incl_assembly_and_setup": n / ((t_pattern + t_assemble + t_setup + ms_dev) * 1e-3)},
            "e2e": {"value": n / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 8 * n,
                    "ms_per_step": ms_e2e, "call": "fsb_solve (host b/x0 in pinned memory -> x)"},
            "gpu_launches": int(launches) * args.steps,
            "roofline": {"bound": "hbm", "kernel": f"{kname}@level{klev}", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "us_per_launch": kms / kcnt * 1e3,
                         "share_of_step": kms / tot, "fine_level_kernels": fine, "top_kernels": kernels},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cube", type=int, default=None,
                    help="cells per side; default 118 (BASELINE configs[2], ~10M tets) on one GPU, 255 (configs[3], ~100M tets) on 2/4/8 GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.cube is None:
        args.cube = 118 if args.gpus <= 1 and int(os.environ.get("WORLD_SIZE", "1")) <= 1 else 255
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
