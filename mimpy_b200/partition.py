"""z-slab partition of a structured hex problem over the GPUs of one node.

Cell index i + nx*j + nx*ny*k and the interleaved face numbering of HexMesh make the cells and the
faces of a slab k in [k0, k1) contiguous, so every rank simply builds a STANDALONE local hex mesh
of nz_loc = k1 - k0 layers:

  * its bottom z-face layer (local faces j*(3nx+1)+3i) is, for rank > 0, the interface shared with
    the lower neighbour: interior faces of the global problem whose other cell lives there;
  * its top z-face layer (the LAST nx*ny local faces) is, for rank < world-1, a ghost copy of the
    interface rows owned by the upper neighbour.

The face system is sub-assembled (each rank holds its own cells' contributions); interface rows are
completed by an exchange-add with the neighbour, dot products by an all-reduce.  This module holds
the host-side index logic (testable on CPU with gloo) and the NCCL bootstrap.
"""
import ctypes as C

import numpy as np


def slab_bounds(nz, world):
    """Balanced contiguous layer ranges [(k0, k1)] for `world` ranks."""
    base, rem = divmod(int(nz), int(world))
    out, k = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((k, k + n))
        k += n
    return out


def hex_counts(nx, ny, nz):
    L = nx * ny + nx * (ny + 1) + (nx + 1) * ny
    return L, nz * L + nx * ny


def local_to_global_faces(nx, ny, nz, k0, k1):
    """Global face id of every local face of the slab mesh (closed form of hexmesh.py:172-228)."""
    L, _ = hex_counts(nx, ny, nz)
    nzl = k1 - k0
    n_low = nzl * L
    g = np.empty(n_low + nx * ny, dtype=np.int64)
    g[:n_low] = np.arange(n_low) + k0 * L
    j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    if k1 < nz:
        top = k1 * L + j * (3 * nx + 1) + 3 * i   # interleaved inside the upper neighbour's range
    else:
        top = nz * L + j * nx + i
    g[n_low:] = top.ravel()
    return g


class SlabPlan(object):
    """Host description of one rank's slab: interface layers, ownership, Neumann/Dirichlet faces
    for the C2/C3 boundary conditions (Dirichlet on x-/x+, no-flow elsewhere)."""

    def __init__(self, nx, ny, nz, rank, world):
        self.nx, self.ny, self.nz, self.rank, self.world = nx, ny, nz, rank, world
        self.k0, self.k1 = slab_bounds(nz, world)[rank]
        self.nzl = self.k1 - self.k0
        self.L, _ = hex_counts(nx, ny, nz)
        _, self.nf_local = hex_counts(nx, ny, self.nzl)
        self.nc_local = nx * ny * self.nzl
        j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
        self.bottom_faces = (j * (3 * nx + 1) + 3 * i).ravel()                       # local ids, (j,i) order
        self.top_faces = (self.nzl * self.L + j * nx + i).ravel()
        self.has_lower = rank > 0
        self.has_upper = rank < world - 1

    def face_flag(self):
        """uint8[nf_local]: 1 = top interface layer (this rank holds the lower-index cell),
        2 = bottom interface layer (the lower-index cell lives on the lower neighbour)."""
        f = np.zeros(self.nf_local, dtype=np.uint8)
        if self.has_upper:
            f[self.top_faces] = 1
        if self.has_lower:
            f[self.bottom_faces] = 2
        return f

    def neumann_mask(self):
        """No-flow on y-/y+ always, on z-/z+ only where the slab touches the global boundary."""
        nx, ny, nzl, L = self.nx, self.ny, self.nzl, self.L
        m = np.zeros(self.nf_local, dtype=np.uint8)
        k, i = np.meshgrid(np.arange(nzl), np.arange(nx), indexing="ij")
        m[(k * L + 3 * i + 1).ravel()] = 1                              # y- : yf(i, 0, k)
        m[(k * L + ny * (3 * nx + 1) + i).ravel()] = 1                  # y+ : yf(i, ny, k)
        if not self.has_lower:
            m[self.bottom_faces] = 1
        if not self.has_upper:
            m[self.top_faces] = 1
        return m

    def dirichlet_faces(self):
        """(x- faces, x+ faces) local ids."""
        nx, ny, nzl, L = self.nx, self.ny, self.nzl, self.L
        k, j = np.meshgrid(np.arange(nzl), np.arange(ny), indexing="ij")
        xm = (k * L + j * (3 * nx + 1) + 2).ravel()
        xp = (k * L + j * (3 * nx + 1) + 3 * nx).ravel()
        return xm, xp

    def halo(self, face_to_flux):
        """Exchange lists in the local flux numbering: (n_owned, idx_lo, rank_lo, idx_hi, rank_hi)."""
        f2f = np.asarray(face_to_flux)
        F = int((f2f[:self.nf_local] >= 0).sum())
        idx_lo = f2f[self.bottom_faces].astype(np.int32) if self.has_lower else np.zeros(0, np.int32)
        idx_hi = f2f[self.top_faces].astype(np.int32) if self.has_upper else np.zeros(0, np.int32)
        assert (idx_lo >= 0).all() and (idx_hi >= 0).all()
        n_owned = F - (len(idx_hi) if self.has_upper else 0)
        if self.has_upper:
            assert idx_hi.min() == n_owned  # ghost rows are the trailing ones
        return n_owned, idx_lo, self.rank - 1, idx_hi, self.rank + 1


def init_nccl(ctx, rank, world):
    """Creates this context's own NCCL communicator; the 128-byte id travels over torch.distributed."""
    import torch
    import torch.distributed as dist

    lib = ctx.lib
    ident = (C.c_uint8 * 128)()
    if rank == 0:
        from ._lib import check
        check(lib.mimpy_b200_nccl_unique_id(C.cast(ident, C.c_void_p)))
    t = torch.tensor(list(bytes(ident)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.to(ctx.device)
    dist.broadcast(t, src=0)
    buf = (C.c_uint8 * 128)(*t.cpu().tolist())
    ctx.call("mimpy_b200_nccl_init", C.cast(buf, C.c_void_p), rank, world)
    ctx.rank, ctx.world = rank, world


def init_peer(ctx, rank, world, halo_cap):
    """NVLink peer-memory path for the halo exchange and the scalar all-reduces (comm.cu): every
    rank allocates a symmetric buffer, the 64-byte cudaIpc handles are all-gathered over
    torch.distributed, every rank maps its peers' buffers.  halo_cap: doubles per interface layer."""
    import torch
    import torch.distributed as dist

    handle = (C.c_uint8 * 64)()
    ctx.call("mimpy_b200_peer_alloc", int(halo_cap), C.cast(handle, C.c_void_p))
    mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8)
    on_dev = dist.get_backend() == "nccl"
    if on_dev:
        mine = mine.to(ctx.device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    flat = torch.cat(parts).cpu().tolist()
    buf = (C.c_uint8 * (64 * world))(*flat)
    ctx.call("mimpy_b200_peer_attach", rank, world, C.cast(buf, C.c_void_p))
    dist.barrier()


def attach_halo(ctx, slab, plan):
    """Uploads the exchange lists of `slab` for the symbolic plan `plan` and registers them."""
    import torch
    from . import device
    from ._lib import ptr

    f2f = device.to_host(ctx, plan.face_to_flux)
    n_owned, idx_lo, rank_lo, idx_hi, rank_hi = slab.halo(f2f)
    with ctx.on_stream():
        d_lo = torch.as_tensor(idx_lo, device=ctx.device) if len(idx_lo) else None
        d_hi = torch.as_tensor(idx_hi, device=ctx.device) if len(idx_hi) else None
        buf = torch.empty(max(2 * (len(idx_lo) + len(idx_hi)), 1), dtype=torch.float64, device=ctx.device)
    ctx.call("mimpy_b200_set_halo", n_owned, len(idx_lo), ptr(d_lo), rank_lo, len(idx_hi), ptr(d_hi), rank_hi, ptr(buf))
    ctx._halo_keepalive = (d_lo, d_hi, buf)
    return n_owned
