#!/usr/bin/env python
"""bench.py — solve DOFs/s of the AMG-preconditioned CG solve (to 1e-8) on synthetic structured tet cubes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cube N]

Workload: Kuhn tet cube with `--cube` cells per side.  The default at EVERY GPU count is BASELINE.json
configs[3], the ~100 M-tet cube (N=255: 16 777 216 DOF, 99 488 250 tets) the metric's 1/2/4/8-GPU series is quoted
on — it fits one B200 (25 GB), so the 1 -> 8 series is one mesh and the driver's scaling efficiency is
like-for-like; on one GPU the line also carries a `secondary` block for configs[2] (N=118: 1 685 159 DOF,
9 858 192 tets), and on N > 1 GPUs rank 0's single-GPU solve time of the same mesh.  Operator K + M assembled on
the GPU, AMG hierarchy built on the GPU (aggregation seed 0), right-hand side b = A x*, x* the egg-carton
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
  parity    the solution against the committed oracle golden of the same cube (tests/golden/, iterations +-2,
            residual history and solution samples to 1e-6) and, on N > 1 GPUs, against rank 0's single-GPU
            solution of the same mesh; a failed check makes the process exit 3 after printing the line
Under torchrun (N > 1) one cube is solved by all N GPUs together (strong scaling): setup is
replicated on every GPU, every level with enough rows per GPU is sharded in contiguous partition ranges,
halo values and dot products travel over NVLink peer memory.
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


def setup_bytes(levels, nnzP, nnz_out):
    """Compulsory traffic of the AMG setup: per smoothing level the operator is read for the
    permutation and written back (24 nnz), split into slabs (10 B per in-partition entry) and A_out (12 B per entry),
    P and R are written (24 nnz_P), and the Galerkin product reads A, P, R once and writes A_c (12 (nnz + 2 nnz_P + nnz_c))."""
    tot = 0
    for l in range(len(levels) - 1):
        n, nnz = levels[l]
        nnzc = levels[l + 1][1]
        tot += 24 * nnz + 10 * (nnz - nnz_out[l]) + 12 * nnz_out[l] + 24 * nnzP[l] + 12 * (nnz + 2 * nnzP[l] + nnzc) + 16 * n
    return tot


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


def host_threads():
    """All the host threads this process may use — torchrun exports OMP_NUM_THREADS=1, which must not leak into the
    CPU arm."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N = args.cube
    steps, warmup = args.steps, args.warmup
    if N > 160:  # bounded sample of the ~100M-tet workload: one complete solve (setup excluded from the metric)
        steps, warmup = 1, 0
    r = run_oracle(N, steps, warmup, threads=host_threads())
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


# ----------------------------------------------------------------------------- parity
def golden_for(N):
    p = os.path.join(ROOT, "tests", "golden", f"oracle_cube{N}_pcg.json")
    if not os.path.exists(p):
        return None, None
    return json.load(open(p)), os.path.relpath(p, ROOT)


def parity_block(N, x, iters, hist, levels, x_single=None, iters_single=None):
    """Compares a solve of the bench workload with the committed oracle golden of the same cube and, for a sharded
    solve, with the single-GPU solution of the same mesh.  Bars: iterations +-2, residual history 1e-6 relative
    at equal iteration index, solution samples / norm 1e-6 (BASELINE.json north_star)."""
    out = {"ok": True, "checks": []}
    g, gpath = golden_for(N)
    if g is not None:
        ho, h = np.array(g["resid_history"]), np.array(hist)
        m = min(len(h), len(ho))
        hist_rel = float(np.max(np.abs(h[:m] / ho[:m] - 1))) if m else None
        xs, gs = x[g["sample_idx"]], np.array(g["x_samples"])
        samp = float(np.max(np.abs(xs - gs) / (1e-6 * np.abs(gs) + 1e-9)))   # <= 1 <=> |diff| <= 1e-6 |golden| + 1e-9 (np.allclose rule of the -m gpu test)
        nrm = float(abs(np.linalg.norm(x) - g["x_norm2"]) / g["x_norm2"])
        ok = abs(iters - g["iterations"]) <= 2 and (hist_rel is None or hist_rel <= 1e-6) and samp <= 1.0 and nrm <= 1e-6 \
            and [int(r) for r, _ in levels] == g["levels"]
        out["oracle_golden"] = {"file": gpath, "iterations": iters, "golden_iterations": g["iterations"], "iters_equal": iters == g["iterations"],
                                "hist_max_rel": hist_rel, "samples_worst_vs_tol": samp, "x_norm_rel": nrm, "levels_equal": [int(r) for r, _ in levels] == g["levels"], "ok": bool(ok)}
        out["ok"] = out["ok"] and bool(ok)
        out["checks"].append("oracle_golden")
    else:
        out["oracle_golden"] = None
    if x_single is not None:
        rel = float(np.linalg.norm(x - x_single) / np.linalg.norm(x_single))
        ok = rel <= 1e-6 and abs(iters - iters_single) <= 2
        out["vs_single_gpu"] = {"rel_l2": rel, "iterations": iters, "single_gpu_iterations": iters_single, "iters_equal": iters == iters_single, "ok": bool(ok)}
        out["ok"] = out["ok"] and bool(ok)
        out["checks"].append("vs_single_gpu")
    return out


# ----------------------------------------------------------------------------- our arm
def run_case(N, args, primary, world, rank, local, torch, dist, fsb):
    """Builds the cube-N problem on this rank's GPU, times the solve (device-resident and host-buffer paths), profiles
    one solve and returns the pieces of the bench line."""
    verts, tets, xstar = build_problem(N)
    n = len(verts)
    ne = len(tets)
    s = fsb.FEMSolver.from_arrays(verts, tets, None, device=local)
    del tets
    s.solverType_, s.tolerance_, s.maxIters_, s.seed_ = 1, 1e-8, 200, 0
    t_pattern_cold = s.time_ms("pattern")   # first GPU work of the process: includes pool growth / module load
    t_pattern = t_assemble = t_setup = float("inf")
    rebuilds = 3 if N <= 160 else 2
    for _ in range(rebuilds):                # steady-state stage timings: best of the rebuilds (allocator pool warm)
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
    ms_single = x_single = it_single = None
    dinfo = None
    lo, hi = 0, n
    if world > 1:
        # the hierarchy is replicated, so every GPU can also solve the whole system alone: rank 0 times that
        # (same mesh, same build, same run) as the strong-scaling reference of this line
        if rank == 0:
            for i in range(2):
                with torch.cuda.stream(stream):
                    x_dev.zero_()
                s.solve_device(x_dev.data_ptr(), b_dev.data_ptr())
            ms_single = s.time_ms("solve")
            x_single, it_single = x_dev.cpu().numpy(), s.iterations
        dist.barrier()
        s.dist_connect(rank, world, fsb.exchange_handles_torch)
        dinfo = s.dist_info()
        lo, hi = dinfo["user_range"]
        log(f"[rank {rank}] sharded levels {dinfo['sharded_levels']}, fine rows {[int(v) for v in s.dist_ranges(0)[1]]}, halo values {dinfo['halo_values']}, host slice [{lo},{hi})")
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
        x_host[lo:hi].zero_()
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
    if rank == 0 and primary:
        sampler.start()
    ms_dev, wall_dev = timed(solve_device, args.steps, args.warmup)
    clocks = sampler.stop() if (rank == 0 and primary) else None
    iters, relres, launches = s.iterations, s.relres, s.last_launches()
    hist = s.resid_history().copy()
    xg = x_dev.cpu().numpy()
    err = float(np.linalg.norm(xg - xstar) / np.linalg.norm(xstar))
    parity = parity_block(N, xg, iters, hist, levels, x_single, it_single) if rank == 0 else None
    ms_e2e, wall_e2e = timed(solve_host, max(1, args.steps), min(args.warmup, 3))
    if rank == 0 and parity is not None:  # the host-buffer path must deliver the same numbers (on its slice)
        xh = x_host.numpy()
        rel_h = float(np.linalg.norm(xh[lo:hi] - xg[lo:hi]) / np.linalg.norm(xg[lo:hi]))
        parity["host_path_rel_l2"] = rel_h
        parity["ok"] = bool(parity["ok"] and rel_h <= 1e-10)
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
    spread = None
    if world > 1:  # the same kernel on every GPU: shortest / longest mean launch (load balance of each phase)
        allp = [None] * world
        dist.all_gather_object(allp, {f"{k}@L{l}": ms / c * 1e3 for (k, l), (c, ms) in prof.items()})
        spread = {name: [round(min(p.get(name, 0.0) for p in allp), 1), round(max(p.get(name, 0.0) for p in allp), 1)]
                  for name in sorted(allp[0], key=lambda nm: -allp[0][nm])[:16]}
    nlev = len(levels)
    nnzP = [int(s._L.fsb_level_int(s.handle, l, b"P_col", None, 0)) for l in range(nlev - 1)]
    nnz_out = [int(s.level_int(l, "Aout_ptr")[-1]) for l in range(nlev - 1)]
    # fraction of a level's rows this GPU owns (sharded levels of a multi-GPU solve)
    own = [1.0] * nlev
    if world > 1:
        for l in range(dinfo["sharded_levels"]):
            rb = s.dist_ranges(l)[1]
            own[l] = float(rb[rank + 1] - rb[rank]) / levels[l][0]

    def kernel_bytes(kname, lev):
        ln, lnnz = levels[lev]
        nc = levels[lev + 1][0] if lev + 1 < nlev else 0
        nP = nnzP[lev] if lev < nlev - 1 else 0
        nin = lnnz - nnz_out[lev] if lev < nlev - 1 else lnnz   # entries inside the partitions' diagonal blocks
        k = "restrict" if (kname == "spmv" and lev < nlev - 1) else kname
        return algorithmic_bytes(k, ln, lnnz, nP, nc, nin) * own[lev]

    # dominant kernel = the largest time share among the kernels with an HBM byte model (exchange kernels and
    # waits measure peer latency, not HBM traffic)
    modeled = {k: v for k, v in prof.items() if kernel_bytes(*k) > 0}
    (kname, klev), (kcnt, kms) = max(modeled.items(), key=lambda kv: kv[1][1])
    bytes_per_launch = kernel_bytes(kname, klev)
    peak, peak_src = measured_peaks()
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of the same workload
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        traffic = tj.get(f"cube{N}", {}).get(f"{kname}@level{klev}")
        if traffic is not None and world > 1:
            traffic, traffic_note = traffic * own[klev], "single-GPU ncu capture scaled by this GPU's share of the level's rows (ncu cannot attach to a multi-rank run)"
    except Exception:
        pass
    if traffic is None:
        traffic_note = "no ncu --set full capture of this kernel on this cube under profiles/"
    achieved = bytes_per_launch / (kms / kcnt * 1e-3) / 1e9
    kernels = {f"{k}@L{l}": {"launches": c, "ms": round(ms, 4), "share": round(ms / tot, 4)} for (k, l), (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:12]}
    # per-kernel achieved GB/s on the fine level (SpMV + smoother % of HBM peak is part of the metric)
    fine = {}
    for k in ("spmv_dot", "pre_smooth", "residual_out", "restrict", "prolong_add", "bprime", "residual", "post_smooth", "cg_update", "dot", "cg_pdir"):
        if (k, 0) in prof:
            c, ms = prof[(k, 0)]
            kb = kernel_bytes(k, 0) if k not in ("residual_out", "bprime") else (12 * nnz_out[0] + 28 * levels[0][0]) * own[0]
            gbs = kb / (ms / c * 1e-3) / 1e9
            fine[k] = {"us": round(ms / c * 1e3, 2), "GBps": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4)}
    # whole-iteration roofline: algorithmic bytes of every modeled kernel of one iteration / time of one iteration
    it_bytes = sum(kernel_bytes(k, l) * c for (k, l), (c, ms) in prof.items()) + sum((12 * nnz_out[l] + 28 * levels[l][0]) * own[l] * prof[(k, l)][0]
                                                                                      for l in range(nlev - 1) for k in ("residual_out", "bprime") if (k, l) in prof)
    stages = {
        "assemble": {"bytes": 84 * ne + 24 * n + 8 * nnz, "ms": t_assemble},
        "pattern": {"bytes": 16 * ne + 64 * ne + 4 * (n + 1) + 4 * nnz, "ms": t_pattern},
        "setup": {"bytes": setup_bytes(levels, nnzP, nnz_out), "ms": t_setup},
        "solve": {"bytes": it_bytes, "ms": ms_dev, "note": "sum of the modeled kernels' algorithmic bytes of one solve on this GPU / device time of the solve"},
    }
    for st in stages.values():
        st["GBps"] = st["bytes"] / (st["ms"] * 1e-3) / 1e9
        st["frac_of_peak"] = st["GBps"] / peak
    out = dict(N=N, n=n, ne=ne, nnz=nnz, levels=levels, ms_dev=ms_dev, wall_dev=wall_dev, ms_e2e=ms_e2e, iters=iters, relres=relres, err=err,
               launches=launches, t_pattern=t_pattern, t_pattern_cold=t_pattern_cold, t_assemble=t_assemble, t_setup=t_setup, t_solvefem=t_solvefem,
               ms_single=ms_single, dinfo=dinfo, lo=lo, hi=hi, clocks=clocks, parity=parity,
               roofline={"bound": "hbm", "kernel": f"{kname}@level{klev}", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "bytes_per_launch": bytes_per_launch,
                         "us_per_launch": kms / kcnt * 1e3, "share_of_step": kms / tot, "fine_level_kernels": fine, "top_kernels": kernels,
                         "us_per_launch_min_max_over_ranks": spread},
               stages=stages)
    if rank == 0 and world == 1 and args.scipy_check and N <= 160:
        out["scipy"] = scipy_sanity(s, b_host.numpy(), xg)
    if world > 1:
        s.dist_disconnect()
    del s
    torch.cuda.empty_cache()
    return out


def scipy_sanity(s, b, x_ours):
    """Independent ground truth: SciPy CG + Jacobi on the same CSR, to 1e-8."""
    try:
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        ptr, col, val = s.matrix_csr()
        A = sp.csr_matrix((val, col, ptr))
        dinv = 1.0 / A.diagonal()
        M = spla.LinearOperator(A.shape, matvec=lambda v: dinv * v)
        cnt = [0]
        t0 = time.perf_counter()
        x, info = spla.cg(A, b, rtol=1e-8, atol=0.0, maxiter=3000, M=M, callback=lambda xk: cnt.__setitem__(0, cnt[0] + 1))
        dt = time.perf_counter() - t0
        return {"solver": "scipy.sparse.linalg.cg + Jacobi, rtol 1e-8", "iterations": cnt[0], "info": int(info), "seconds": dt,
                "dofs_per_s": A.shape[0] / dt, "relres": float(np.linalg.norm(b - A @ x) / np.linalg.norm(b)),
                "rel_l2_vs_ours": float(np.linalg.norm(x - x_ours) / np.linalg.norm(x_ours))}
    except Exception as e:
        return {"error": str(e)}


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
    r = run_case(N, args, True, world, rank, local, torch, dist, fsb)
    secondary = None
    if world == 1 and N != 118 and not args.no_secondary:
        q = run_case(118, args, False, world, rank, local, torch, dist, fsb)
        secondary = {"workload": workload_name(118) + " (BASELINE configs[2])", "value": q["n"] / (q["ms_dev"] * 1e-3), "unit": UNIT, "ms_per_step": q["ms_dev"],
                     "e2e": {"value": q["n"] / (q["ms_e2e"] * 1e-3), "ms_per_step": q["ms_e2e"]}, "iterations": q["iters"], "relres": q["relres"],
                     "rel_l2_err_vs_exact": q["err"], "levels": q["levels"], "pattern_ms": q["t_pattern"], "assemble_ms": q["t_assemble"], "setup_ms": q["t_setup"],
                     "solveFEM_host_ms": q["t_solvefem"] * 1e3, "dofs_per_s_incl_assembly_and_setup": q["n"] / ((q["t_pattern"] + q["t_assemble"] + q["t_setup"] + q["ms_dev"]) * 1e-3),
                     "roofline": q["roofline"], "roofline_stages": q["stages"], "parity": q["parity"], "scipy_cg_jacobi": q.get("scipy")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Nc = N if N <= 160 else 118   # bounded sample: the ~10 M-tet cube of the same family when the workload is the ~100 M-tet one
        try:
            c = run_oracle(Nc, 1, 0, threads=host_threads(), single_thread_solve=True)
            cpu = {"value": c["n"] / c["t_solve"], "unit": UNIT, "cores": c["threads"], "kind": "port",
                   "single_thread": ({"value": c["n"] / c["t_solve_1thread"], "unit": UNIT, "solve_s": c["t_solve_1thread"]}
                                     if c.get("t_solve_1thread") else None),
                   "sample": (f"Kuhn cube N={Nc} ({c['n']} DOF" + ("" if Nc == N else f", the same family at 1/{round(r['n'] / c['n'])} of the workload's DOFs") +
                              f"), 1 complete PCG solve ({c['iters']} iterations, {c['t_solve']:.2f} s; setup {c['t_setup']:.1f} s excluded); "
                              "CPU oracle = restatement of the reference (no host solve and no offline CUDA build upstream)"),
                   "iterations": c["iters"], "setup_s": c["t_setup"], "assemble_s": c["t_assemble"], "pattern_s": c["t_pattern"]}
        except Exception as e:  # the bench line must still be printed
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"oracle failed: {e}"}

    rc = 0
    if rank == 0:
        n, ms_dev = r["n"], r["ms_dev"]
        if world == 1:
            par = "single GPU"
        else:
            d = r["dinfo"]
            par = (f"{world} GPUs: replicated setup; levels 0..{d['sharded_levels'] - 1} sharded in contiguous partition ranges, the rest replicated after an "
                   "all-gather; NVLink peer-memory pushes with per-neighbour epoch flags (producers never wait, consumers wait in their prologue), in-kernel all-reduce")
        line = {
            "metric": METRIC, "value": n / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(N), "parallelism": par,
                       "l2_policy": "inputs larger than L2 (hierarchy + vectors ~%.0f MB per solve pass)" % ((12 * r["nnz"] + 40 * n) / 1e6),
                       "iterations": r["iters"], "relres": r["relres"], "rel_l2_err_vs_exact": r["err"], "levels": r["levels"],
                       "pattern_ms": r["t_pattern"], "pattern_cold_ms": r["t_pattern_cold"], "assemble_ms": r["t_assemble"], "setup_ms": r["t_setup"],
                       "solveFEM_host_ms": None if r["t_solvefem"] != r["t_solvefem"] else r["t_solvefem"] * 1e3,
                       "wall_ms_per_step": r["wall_dev"],
                       "single_gpu_same_mesh_ms": r["ms_single"],
                       "speedup_vs_single_gpu_same_mesh": (r["ms_single"] / ms_dev) if r["ms_single"] else None,
                       "sharding": r["dinfo"],
                       "dofs_per_s_incl_assembly_and_setup": n / ((r["t_pattern"] + r["t_assemble"] + r["t_setup"] + ms_dev) * 1e-3)},
            "e2e": {"value": n / (r["ms_e2e"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 16 * (r["hi"] - r["lo"]), "d2h_bytes_per_step": 8 * (r["hi"] - r["lo"]),
                    "ms_per_step": r["ms_e2e"], "call": "fsb_solve (host b/x0 in pinned memory -> x)" + ("" if world == 1 else "; every rank moves the user-numbering slice that covers its rows")},
            "gpu_launches": int(r["launches"]) * args.steps,
            "roofline": r["roofline"],
            "roofline_stages": r["stages"],
            "parity": r["parity"],
            "secondary": secondary,
            "cpu_baseline": cpu,
            "clocks": r["clocks"],
        }
        emit(line)
        if r["parity"] is not None and not r["parity"]["ok"]:
            log("PARITY FAILED:", json.dumps(r["parity"]))
            rc = 3
        if secondary and secondary["parity"] is not None and not secondary["parity"]["ok"]:
            log("PARITY FAILED (secondary):", json.dumps(secondary["parity"]))
            rc = 3
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cube", type=int, default=255,
                    help="cells per side; default 255 (BASELINE configs[3], ~100M tets: the mesh of the 1/2/4/8-GPU series); 118 = configs[2]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[2] (N=118) block of a single-GPU run")
    ap.add_argument("--no-scipy", dest="scipy_check", action="store_false", help="skip the SciPy CG sanity solve of the secondary block")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
