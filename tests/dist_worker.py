"""Worker for the multi-process tests (launched with torch.distributed.run).

  cpu mode: gloo, no GPU — exercises the host-side plumbing of the sharded solve (handle all-gather
            order, weight-balanced split agreed by every rank).
  gpu mode: nccl, one GPU per rank — sharded PCG solve vs the single-GPU solve of the same problem.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sci_solver_fem_b200 as fsb  # noqa: E402


def cpu_mode():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    payload = bytes([(rank * 37 + i) % 256 for i in range(64)])
    got = fsb.exchange_handles_torch(payload)
    assert len(got) == world
    for r in range(world):
        assert got[r] == bytes([(r * 37 + i) % 256 for i in range(64)]), "handles out of rank order"
    # every rank must derive the same split from the same weights
    rng = np.random.default_rng(5)
    w = rng.integers(1000, 9000, size=997)
    cut = fsb.split_by_weight(w, world)
    t = torch.from_numpy(cut.astype(np.int64))
    ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(t, ref)
    assert cut[0] == 0 and cut[-1] == w.size and np.all(np.diff(cut) > 0)
    loads = np.add.reduceat(w, cut[:-1])
    assert loads.max() <= 1.05 * w.sum() / world + w.max()
    dist.barrier()
    if rank == 0:
        print("CPU_DIST_OK")
    dist.destroy_process_group()


def gpu_mode(N):
    """Sharded PCG solve vs the single-GPU solve of the same problem (same process, same hierarchy).
    FSB_DIST_SAME_GPU=1: every rank uses cuda:0 (two processes time-slice one GPU; the peer arenas are still
    reached through CUDA IPC), process group on gloo — lets a one-GPU box run the sharded code path."""
    same_gpu = os.environ.get("FSB_DIST_SAME_GPU") == "1"
    local = 0 if same_gpu else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if same_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    v, t = fsb.meshio.kuhn_cube(N)
    n = len(v)
    s = fsb.FEMSolver.from_arrays(v, t, device=local)
    s.solverType_, s.tolerance_, s.maxIters_, s.seed_ = 1, 1e-8, 200, 0
    s.setup()
    rng = np.random.default_rng(1234)
    b = rng.uniform(-1, 1, n)
    x1 = s.solve(np.zeros_like(b), b).copy()          # replicated single-GPU solve on every rank
    it1, h1 = s.iterations, s.resid_history().copy()
    s.dist_connect(rank, world, fsb.exchange_handles_torch)
    info = s.dist_info()
    want_levels = int(os.environ.get("FSB_EXPECT_SHARDED_LEVELS", "0"))
    assert info["sharded_levels"] >= max(1, want_levels), info
    for l in range(info["sharded_levels"]):
        pb, rb, ab = s.dist_ranges(l)
        assert pb[0] == 0 and rb[-1] == s.level_rows(l) and np.all(np.diff(rb) > 0), (l, list(rb))
    want = os.environ.get("FSB_EXPECT_OVERLAP")
    if want:
        a = info["interior_ranges"][0]["operator_rows"]
        assert (a[1] > a[0]) == (want == "1" or rank % 2 == 1), (rank, info["interior_ranges"])
    lo, hi = info["user_range"]
    assert 0 <= lo < hi <= n
    dist.barrier()
    b_dev = torch.from_numpy(b).cuda()
    for rep in range(2):
        # host buffers: only the slice [lo, hi) crosses PCIe; x outside it stays as the caller left it
        x0 = np.full(n, 7.0)
        x0[lo:hi] = 0.0
        xd = s.solve(x0, b)
        assert abs(s.iterations - it1) <= 1, (s.iterations, it1)
        hd = s.resid_history()
        m = min(len(hd), len(h1))
        assert np.allclose(hd[:m], h1[:m], rtol=1e-6), np.abs(hd[:m] / h1[:m] - 1).max()
        err = np.linalg.norm(xd[lo:hi] - x1[lo:hi]) / np.linalg.norm(x1[lo:hi])
        assert err < 1e-8, err
        assert np.all(xd[:lo] == 7.0) and np.all(xd[hi:] == 7.0), "host x was written outside the rank's slice"
        # device buffers: every GPU ends up with the full solution
        x_dev = torch.zeros(n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        s.solve_device(x_dev.data_ptr(), b_dev.data_ptr())
        xf = x_dev.cpu().numpy()
        errf = np.linalg.norm(xf - x1) / np.linalg.norm(x1)
        assert errf < 1e-8, errf
    # all ranks hold the same full solution, bit for bit
    tt = torch.from_numpy(xf) if same_gpu else torch.from_numpy(xf).cuda()   # gloo moves host tensors, nccl device tensors
    ref = tt.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(tt, ref), "ranks disagree on the solution"
    dist.barrier()
    s.dist_disconnect()
    x2 = s.solve(np.zeros_like(b), b)                   # back to the single-GPU path
    assert np.array_equal(x2, x1)
    dist.barrier()
    if rank == 0:
        print(f"GPU_DIST_OK world={world} sharded_levels={info['sharded_levels']} iters={it1} rel_diff={errf:.2e} halo={info['halo_values']}")
    dist.destroy_process_group()


if __name__ == "__main__":
    if sys.argv[1] == "cpu":
        cpu_mode()
    else:
        gpu_mode(int(sys.argv[2]) if len(sys.argv) > 2 else 40)
