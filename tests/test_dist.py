"""Multi-process tests of stage 4 (sharded solve): host plumbing on CPU with gloo (world_size 2), and —
when the box has >= 2 GPUs — the sharded PCG solve against the single-GPU solve."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(nproc, args, timeout, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), WORKER] + args
    e = dict(os.environ)
    e.update(env or {})
    for attempt in range(3):  # the probed port can be taken again before torchrun binds it: retry with a new one
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)
        if r.returncode == 0 or "EADDRINUSE" not in r.stderr:
            return r
        cmd[cmd.index("--master-port") + 1] = str(free_port())
    return r


def test_host_plumbing_gloo_world2():
    r = launch(2, ["cpu"], 300)
    assert r.returncode == 0 and "CPU_DIST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_split_by_weight_properties():
    import numpy as np
    import sci_solver_fem_b200 as fsb
    w = np.full(100, 10)
    assert list(fsb.split_by_weight(w, 4)) == [0, 25, 50, 75, 100]
    assert list(fsb.split_by_weight(w, 1)) == [0, 100]
    w = np.array([1000, 1, 1, 1, 1, 1, 1, 1])
    cut = fsb.split_by_weight(w, 2)
    assert cut[0] == 0 and cut[-1] == 8 and 1 <= cut[1] <= 7
    cut = fsb.split_by_weight(np.ones(3, dtype=np.int64), 8)   # more ranks than partitions: trailing ranks are empty
    assert cut[0] == 0 and cut[-1] == 3 and all(cut[i] <= cut[i + 1] for i in range(8))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("all_levels", [0, 1])
def test_sharded_solve_matches_single_gpu(world, all_levels):
    """all_levels=1 lowers the rows-per-GPU threshold so that every smoothing level of the small test mesh is
    sharded (exercises the down / up exchanges and the owned-range variants of the coarse smoothers)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    # (the dense tail is switched off so that level 1 of the small mesh is a smoothing level that can be sharded)
    env = {"FSB_SHARD_MINROWS": "1", "FSB_DENSE_TAIL": "0", "FSB_EXPECT_SHARDED_LEVELS": "2" if world == 2 else "1",
           "FSB_OVERLAP_MINROWS": "1"} if all_levels else {}
    r = launch(world, ["gpu", "40"], 900, env)
    assert r.returncode == 0 and "GPU_DIST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-6000:]


@pytest.mark.gpu
@pytest.mark.parametrize("all_levels", [0, 1, 2, 3])
def test_sharded_solve_two_ranks_on_one_gpu(all_levels):
    """The sharded code path on a ONE-GPU box: two processes time-slice cuda:0 and exchange through CUDA IPC
    exactly as they would over NVLink (slow — every cross-rank wait costs a time slice — but it is the same
    kernels, channels and push lists), compared with the single-GPU solve."""
    env = {"FSB_DIST_SAME_GPU": "1"}
    if all_levels:
        env.update({"FSB_SHARD_MINROWS": "1", "FSB_DENSE_TAIL": "0", "FSB_EXPECT_SHARDED_LEVELS": "2"})
    if all_levels >= 2:  # also the interior / boundary split: exchanges on their own stream next to the interior rows of their consumers
        env.update({"FSB_OVERLAP_MINROWS": "1", "FSB_EXPECT_OVERLAP": "1"})
    if all_levels == 3:  # ... on one of the two ranks only (GPUs decide independently; the exchange sequence must stay uniform)
        env.update({"FSB_OVERLAP_ODD_RANKS": "1", "FSB_EXPECT_OVERLAP": "odd"})
    r = launch(2, ["gpu", "40" if all_levels >= 2 else "32"], 1200, env)
    assert r.returncode == 0 and "GPU_DIST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-6000:]
