"""Slab partition: host index logic on CPU (world_size-2 gloo) and the N-GPU solve (gpu marker)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_and_face_maps():
    from mimpy_b200 import partition
    from oracle import mimpy_oracle as orc

    assert partition.slab_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    nx, ny, nz = 4, 3, 5
    full = orc.hex_mesh(nx, ny, nz, 1., 1., 1.)
    for world in (1, 2, 3):
        seen_cells = 0
        for rank in range(world):
            s = partition.SlabPlan(nx, ny, nz, rank, world)
            loc = orc.hex_mesh(nx, ny, s.nzl, 1., 1., float(s.nzl) / nz)
            g = partition.local_to_global_faces(nx, ny, nz, s.k0, s.k1)
            assert len(g) == loc.nf == s.nf_local
            # the local cells are the global cells of the slab, face by face
            cells_glob = full.cell_faces.reshape(-1, 6)[s.k0 * nx * ny:s.k1 * nx * ny]
            assert np.array_equal(g[loc.cell_faces.reshape(-1, 6)], cells_glob)
            seen_cells += s.nc_local
            # flags: bottom/top layers are exactly the faces shared with the neighbour slabs
            flag = s.face_flag()
            f2c = full.face_to_cell[g]
            shared_up = (f2c >= s.k1 * nx * ny).any(axis=1)
            shared_dn = ((f2c >= 0) & (f2c < s.k0 * nx * ny)).any(axis=1)
            assert np.array_equal(flag == 1, shared_up) and np.array_equal(flag == 2, shared_dn)
            # Neumann mask = global no-flow faces (y and z boundaries) restricted to the slab
            glob_mask = np.zeros(full.nf, dtype=np.uint8)
            for b in (2, 3, 4, 5):
                glob_mask[[f for f, _ in full.boundary_faces[b]]] = 1
            assert np.array_equal(s.neumann_mask(), glob_mask[g])
            xm, xp = s.dirichlet_faces()
            assert set(g[xm]) == {f for f, _ in full.boundary_faces[0]} & set(g)
            assert set(g[xp]) == {f for f, _ in full.boundary_faces[1]} & set(g)
        assert seen_cells == nx * ny * nz


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from mimpy_b200 import partition

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny, nz = 5, 4, 6
    s = partition.SlabPlan(nx, ny, nz, rank, world)
    mask = s.neumann_mask()
    f2f = (np.cumsum(1 - mask) - 1).astype(np.int32)
    f2f[mask == 1] = -1
    n_owned, idx_lo, rank_lo, idx_hi, rank_hi = s.halo(f2f)
    g = partition.local_to_global_faces(nx, ny, nz, s.k0, s.k1)
    flux_to_face = np.nonzero(mask == 0)[0]
    # what I send up must be, entry by entry, the global faces my upper neighbour expects from below
    mine = {"hi": g[flux_to_face[idx_hi]].tolist(), "lo": g[flux_to_face[idx_lo]].tolist(), "owned": int(n_owned),
            "F": int((mask == 0).sum())}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = True
    if rank + 1 < world:
        ok &= gathered[rank]["hi"] == gathered[rank + 1]["lo"] and len(mine["hi"]) == nx * ny and rank_hi == rank + 1
    if rank > 0:
        ok &= rank_lo == rank - 1
    # every global flux row is owned exactly once
    total_owned = sum(x["owned"] for x in gathered)
    from oracle import mimpy_oracle as orc
    full = orc.hex_mesh(nx, ny, nz, 1., 1., 1.)
    n_neu = sum(len(full.boundary_faces[b]) for b in (2, 3, 4, 5))
    ok &= total_owned == full.nf - n_neu
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_halo_lists_agree_between_ranks_gloo():
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


@pytest.mark.gpu
def test_two_gpu_solve_matches_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "scripts", "multi_check.py"), "24"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI_CHECK" in r.stdout
