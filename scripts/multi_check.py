"""torchrun --nproc-per-node N scripts/multi_check.py [n]: the slab-partitioned solve on N GPUs
reproduces the single-GPU solve (pressure field, relative L2 <= 1e-8)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from mimpy_b200 import _lib, device, partition
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
ctx = _lib.Context(local_rank)
partition.init_nccl(ctx, rank, world)
comm = os.environ.get("MIMPY_B200_COMM", "peer")
if comm != "nccl":
    partition.init_peer(ctx, rank, world, n * n)
pb = bench.build_problem(ctx, n, rank, world)
dm, plan, rhs = pb["dm"], pb["plan"], pb["rhs"]
counts = [(k1 - k0) * n * n for k0, k1 in partition.slab_bounds(n, world)]


def gather_pressure(sol):
    p_loc = sol[plan.flux_dof:].contiguous()
    parts = [torch.empty(c, dtype=torch.float64, device=ctx.device) for c in counts]
    torch.cuda.synchronize()
    if len(set(counts)) == 1:
        dist.all_gather(parts, p_loc)
    else:
        for r in range(world):
            buf = p_loc if r == rank else parts[r]
            dist.broadcast(buf, src=r)
            parts[r] = buf
    return torch.cat(parts).cpu().numpy()


def solve_twice(make):
    hyb = make()
    sol, info = hyb.solve(rhs, rtol=1e-13, maxit=20000, check_every=25)
    sol = sol.clone()
    sol2, info2 = hyb.solve(rhs, rtol=1e-13, maxit=20000, check_every=25)   # run-to-run determinism
    repeat_ok = bool(torch.equal(sol, sol2)) and info2.iterations == info.iterations
    flags = [None] * world
    dist.all_gather_object(flags, repeat_ok)
    return gather_pressure(sol), info, flags


m_e, _, _ = device.local_matrices(dm, plan, 1)
# (a) blocks given: two-step set-up; (b) the bench path: M_E built inside the fused (brick) set-up kernel
runs = [("blocks", solve_twice(lambda: device.HybridSystem(dm, plan, m_e))),
        ("fused", solve_twice(lambda: device.HybridSystem(dm, plan, None, method=1)))]
ok = True
if rank == 0:
    ctx1 = _lib.Context(local_rank)
    pb1 = bench.build_problem(ctx1, n, 0, 1)
    m1, _, _ = device.local_matrices(pb1["dm"], pb1["plan"], 1)
    s1, i1 = device.HybridSystem(pb1["dm"], pb1["plan"], m1).solve(pb1["rhs"], rtol=1e-13, maxit=20000, check_every=25)
    p1 = device.to_host(ctx1, s1[pb1["plan"].flux_dof:])
    for name, (p_multi, info, flags) in runs:
        err = np.linalg.norm(p_multi - p1) / np.linalg.norm(p1)
        print("MULTI_CHECK %s comm=%s world=%d n=%d iters multi=%d single=%d rel_l2=%.3e bitwise_repeat=%s"
              % (name, comm, world, n, info.iterations, i1.iterations, err, flags))
        ok = ok and err < 1e-8 and (comm == "nccl" or all(flags))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
