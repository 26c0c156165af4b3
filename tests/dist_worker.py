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
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    v, t = fsb.meshio.kuhn_cube(N)
    s = fsb.FEMSolver.from_arrays(v, t, device=local)
    s.solverType_, s.tolerance_, s.maxIters_, s.seed_ = 1, 1e-8, 200, 0
    s.setup()
    rng = np.random.default_rng(1234)
    b = rng.uniform(-1, 1, len(v))
    x1 = s.solve(np.zeros_like(b), b).copy()          # replicated single-GPU solve on every rank
    it1, h1 = s.iterations, s.resid_history().copy()
    s.dist_connect(rank, world, fsb.exchange_handles_torch)
    pb, rb, ab = s.dist_ranges()
    assert pb[0] == 0 and rb[-1] == len(v) and np.all(np.diff(rb) > 0)
    dist.barrier()
    for rep in range(2):
        xd = s.solve(np.zeros_like(b), b)
        assert s.iterations == it1, (s.iterations, it1)
        hd = s.resid_history()
        assert np.allclose(hd, h1, rtol=1e-9), np.abs(hd / h1 - 1).max()
        err = np.linalg.norm(xd - x1) / np.linalg.norm(x1)
        assert err < 1e-10, err
    # all ranks hold the same full solution
    t = torch.from_numpy(xd).cuda()
    ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(t, ref), "ranks disagree on the solution"
    s.dist_disconnect()
    x2 = s.solve(np.zeros_like(b), b)                   # back to the single-GPU path
    assert np.array_equal(x2, x1)
    dist.barrier()
    if rank == 0:
        print(f"GPU_DIST_OK world={world} iters={it1} rel_diff={err:.2e} rows={list(rb)}")
    dist.destroy_process_group()


if __name__ == "__main__":
    if sys.argv[1] == "cpu":
        cpu_mode()
    else:
        gpu_mode(int(sys.argv[2]) if len(sys.argv) > 2 else 40)
