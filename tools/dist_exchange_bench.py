"""Latency of the cross-GPU primitives without any compute in between (run under torchrun, one rank per GPU):
   torchrun --nproc-per-node N tools/dist_exchange_bench.py [--cube 255]
Prints, per channel of the sharded solve, microseconds per back-to-back exchange and the number of values this rank
sends / receives, and the latency of the in-kernel all-reduce."""
import argparse, ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sci_solver_fem_b200 as fsb
ap = argparse.ArgumentParser(); ap.add_argument("--cube", type=int, default=255); ap.add_argument("--reps", type=int, default=200); a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
v, t = fsb.meshio.kuhn_cube(a.cube)
s = fsb.FEMSolver.from_arrays(v, t, device=local)
s.solverType_, s.tolerance_, s.maxIters_, s.seed_ = 1, 1e-8, 200, 0
s.setup(); s.dist_connect(rank, world, fsb.exchange_handles_torch)
info = s.dist_info()
L = s._L; L.fsb_dist_bench_exchange.restype = C.c_double; L.fsb_dist_bench_exchange.argtypes = [C.c_void_p, C.c_int, C.c_int]
names = {0: "p halo", 1: "x0 halo"}
for l in range(info["sharded_levels"]):
    for k, nm in enumerate(("x-pre halo", "residual halo", "down", "up", "x-post halo")):
        names[2 + 5 * l + k] = f"L{l} {nm}"
dist.barrier()
for ch in [-1] + sorted(names):
    if ch >= 0 and names[ch].endswith("up") and int(names[ch][1]) + 1 >= info["sharded_levels"]:
        continue
    dist.barrier(); torch.cuda.synchronize()
    us = L.fsb_dist_bench_exchange(s.handle, ch, a.reps)
    tt = torch.tensor([us], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{'all-reduce (dot kernel, 1024 values)' if ch < 0 else names[ch]:38s} {float(tt[0]):8.2f} us", flush=True)
if rank == 0:
    print("halo values per exchange (rank 0):", info["halo_values"])
dist.barrier(); dist.destroy_process_group()
