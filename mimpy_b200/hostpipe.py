"""Host-buffer entry points of the hot path: the caller hands over PINNED host arrays and gets pinned host
arrays back; uploads, kernels and downloads are pipelined on three CUDA streams inside the library.

    assemble_to_host   cell_k (host) -> numeric assembly, slab by slab -> CSR values (host)
    solve_to_host      cell_k (host) -> fused hybrid set-up -> PCG -> [v; p] (host)

These are what bench.py's `e2e` figures time (BASELINE metric through the public API with host buffers)."""
import torch

from . import device


class AssemblyPipeline(object):
    """Slab-wise numeric assembly of a structured hex mesh with host in / host out.

    The mesh is cut into slabs of cell layers (multiples of the brick height 4).  Slab s: upload of its
    part of cell_k (stream A) -> mimpy_b200_mfd_numeric_layers on the context stream -> download of the
    CSR value ranges that became final (stream B).  PCIe is full duplex, so the download of slab s overlaps
    the upload of slab s+1 and the kernel of the slab in between."""

    def __init__(self, dmesh, plan, method, n_slabs=8):
        self.dm, self.plan, self.method = dmesh, plan, int(method)
        ctx = dmesh.ctx
        cx, cy, cz = dmesh.hex_dims
        n_slabs = max(1, min(int(n_slabs), cz // 4))
        per = ((cz + n_slabs - 1) // n_slabs + 3) // 4 * 4
        self.bounds = list(range(0, cz, per)) + [cz]
        self.layer = cx * cy
        self.ranges = device.layer_value_ranges(dmesh, plan, self.bounds)
        self.s_in = torch.cuda.Stream(device=ctx.device)
        self.s_out = torch.cuda.Stream(device=ctx.device)
        with ctx.on_stream():
            self.vals = torch.empty(plan.nnz, dtype=torch.float64, device=ctx.device)
        # the kernel that serves layer ranges must apply to this mesh (structured hex, even nx)
        self.piped = device.numeric_layers(dmesh, plan, self.method, self.vals, 0, cz)
        ctx.sync()

    def run(self, k_pinned, vals_pinned):
        """k_pinned: (nc, 9) pinned host tensor; vals_pinned: (nnz,) pinned host tensor, filled on return."""
        dm, plan, ctx = self.dm, self.plan, self.dm.ctx
        layer = self.layer
        if not self.piped:
            with ctx.on_stream():
                dm.cell_k.copy_(k_pinned, non_blocking=True)
                device.numeric(dm, plan, self.method, vals=self.vals)
                vals_pinned.copy_(self.vals, non_blocking=True)
            ctx.sync()
            return
        self.s_in.wait_stream(ctx.stream)
        for (k0, k1), ((f0, f1), (c0, c1)) in zip(zip(self.bounds[:-1], self.bounds[1:]), self.ranges):
            with torch.cuda.stream(self.s_in):
                dm.cell_k[k0 * layer:k1 * layer].copy_(k_pinned[k0 * layer:k1 * layer], non_blocking=True)
                ev_in = self.s_in.record_event()
            ctx.stream.wait_event(ev_in)
            device.numeric_layers(dm, plan, self.method, self.vals, k0, k1)
            ev_k = ctx.stream.record_event()
            self.s_out.wait_event(ev_k)
            with torch.cuda.stream(self.s_out):
                vals_pinned[f0:f1].copy_(self.vals[f0:f1], non_blocking=True)
                vals_pinned[c0:c1].copy_(self.vals[c0:c1], non_blocking=True)
        self.s_out.synchronize()
        ctx.sync()


def solve_to_host(dmesh, plan, method, k_pinned, rhs, sol_pinned, rtol=1e-8, precond="auto", saddle_vals=None,
                  **solve_kw):
    """cell_k from pinned host memory -> fused hybrid set-up -> PCG -> [v; p] into pinned host memory.
    Returns the solver info.  (The saddle CSR itself is not needed to solve; pass saddle_vals to have the
    stopping test verified in it.)"""
    ctx = dmesh.ctx
    with ctx.on_stream():
        dmesh.cell_k.copy_(k_pinned, non_blocking=True)
        hyb = device.HybridSystem(dmesh, plan, None, precond=precond, method=method)
        sol, info = hyb.solve(rhs, rtol=rtol, norm="saddle", saddle_vals=saddle_vals, **solve_kw)
        sol_pinned.copy_(sol, non_blocking=True)
    ctx.sync()
    return info
