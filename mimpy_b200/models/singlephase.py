"""Time-dependent slightly compressible single-phase flow with the API of
mimpy.models.singlephase.SinglePhase (reference singlephase.py:16-299).

Each implicit time step (singlephase.py:252-291) re-scales M by mu/rho(p) per cell (update_m),
sets the accumulation diagonal c phi |E| / dt, adds it times the old pressure (and the point
sources) to the right-hand side and solves the saddle system.  Here: mimpy_b200_singlephase_terms,
the fused numeric assembly is not even needed (only the hybridised face system is rebuilt), the
solve is the hybridised PCG.  No VTK output; start_solving_newton
(singlephase.py:301-376, a fixed 100-sweep Picard loop) is not provided.
"""
import ctypes as C

import numpy as np

from .. import _lib, device
from ..mfd import mfd as _mfd


class SinglePhase(object):
    _couplings_supported = True   # pointer couplings in the time loop (singlephase.py:199-223)

    def __init__(self):
        self.mesh = None
        self.mfd = None
        self.initial_pressure = None
        self.current_pressure = None
        self.current_velocity = None
        self.porosities = None
        self.viscosity = None
        self.ref_pressure = None
        self.ref_density = None
        self.compressibility = None
        self.m_x_coo_length = None
        self.c_start = None
        self.rhs_mfd = None
        self.delta_t = 1.
        self.number_of_time_steps = 1
        self.output_frequency = 1
        self.rate_wells = []
        self.rate_wells_rate = []
        self.rate_wells_name = []
        self.model_name = "model"
        self.pcg_rtol = 1.e-12

    def set_model_name(self, name):
        self.model_name = name

    def set_output_frequency(self, frequency):
        self.output_frequency = frequency

    def set_mesh(self, mesh, ctx=None):
        """singlephase.py:88-96: method 1, consistency check of M_E switched on."""
        self.mesh = mesh
        self.mfd = _mfd.MFD()
        self.mfd.set_mesh(mesh, ctx)
        self.mfd.set_m_e_construction_method(1)
        self.mfd.check_m_e = True
        self.ctx = self.mfd.ctx

    def set_compressibility(self, compressibility):
        self.compressibility = compressibility

    def set_ref_density(self, ref_density):
        self.ref_density = ref_density

    def set_ref_pressure(self, ref_pressure):
        self.ref_pressure = ref_pressure

    def set_viscosity(self, viscosity):
        self.viscosity = viscosity

    def set_initial_pressure(self, pressures):
        self.initial_pressure = pressures
        self.current_pressure = pressures
        self.current_velocity = np.zeros(self.mfd.flux_dof)

    def set_porosities(self, porosities):
        self.porosities = porosities

    def set_porosity_by_cell(self, cell_index, porosity):
        self.porosities[cell_index] = porosity

    def add_rate_well(self, injection_rate, cell_index, well_name):
        raise NameError("add_rate_well not yet implemented")  # singlephase.py:137-142

    def add_point_rate_well(self, injection_rate, cell_index, well_name):
        self.rate_wells.append(cell_index)
        self.rate_wells_rate.append(injection_rate)
        self.rate_wells_name.append(well_name)
        return len(self.rate_wells) - 1

    def set_time_step_size(self, delta_t):
        self.delta_t = delta_t

    def set_number_of_time_steps(self, number_of_time_steps):
        self.number_of_time_steps = number_of_time_steps

    def apply_pressure_boundary_from_function(self, boundary_marker, p_function):
        self.mfd.apply_dirichlet_from_function(boundary_marker, p_function)

    def apply_flux_boundary_from_function(self, boundary_marker, f_function):
        self.mfd.apply_neumann_from_function(boundary_marker, f_function)

    def initialize_system(self):
        """singlephase.py:182-237: blocks, numbering, static right-hand side."""
        import torch

        mfd = self.mfd
        if self.mesh.has_pointer_couplings() and not self._couplings_supported:
            raise NotImplementedError("Transport on a mesh with pointer couplings (transport.py:209-239): its time loop "
                                      "solves the uncoupled saddle system; MFD.solve and SinglePhase handle couplings")
        dm, plan = mfd._ensure_plan()
        self._dm, self._plan = dm, plan
        self._m_e = mfd._blocks()
        self.m_x_coo_length = int(device.coo_view(dm, plan, None, want_triples=False)[0][-1].item())
        n_div = int((mfd.face_to_flux[:, 0][self.mesh.cells.compact()[1]] >= 0).sum())
        self.c_start = self.m_x_coo_length + 2 * n_div
        self.c_end = dm.nc
        self.rhs_mfd = mfd.build_rhs()
        self._d_rhs_mfd = mfd.device_rhs
        self._coupling = None
        if self.mesh.has_pointer_couplings():
            # singlephase.py:199-223: the coupling terms L, L^T sit off the mesh pattern; they are kept as a CSR of their
            # own (reference numbering) next to the saddle matrix of the mesh, which is re-assembled every time step
            from scipy import sparse
            n = mfd.get_number_of_dof()
            cd, cr, cc = mfd.build_coupling_terms()
            e = sparse.coo_matrix((cd, (cr, cc)), shape=(n, n)).tocsr()
            e.sort_indices()
            dev = self.ctx.device
            with self.ctx.on_stream():
                self._coupling = (torch.as_tensor(e.indptr.astype(np.int32), device=dev),
                                  torch.as_tensor(e.indices.astype(np.int32), device=dev),
                                  torch.as_tensor(e.data, device=dev))
                self._mesh_rows = None
                if mfd._mult is not None:   # multiplier faces: device rows -> reference rows (mfd.reference_numbering)
                    self._mesh_rows = torch.as_tensor(mfd._mult[0], device=dev)
                    self._d_rhs_mfd = self._d_rhs_mfd[self._mesh_rows].contiguous()
        ctx = self.ctx
        with ctx.on_stream():
            self._d_mult = torch.empty(dm.nc, dtype=torch.float64, device=ctx.device)
            self._d_cdiag = torch.empty(dm.nc, dtype=torch.float64, device=ctx.device)
        self._hyb = None

    def start_solving_newton(self, inner_iterations=100):
        """singlephase.py:300-356: every time step repeats the implicit step 100 times, each pass starting from the
        pressure the previous pass returned (a fixed-point sweep towards the state where the accumulation term
        vanishes); output after the last pass of a step.  The same device loop as start_solving, 100 passes per step."""
        outer, frequency, user_output = self.number_of_time_steps, self.output_frequency, self.time_step_output
        assigned = "time_step_output" in self.__dict__   # set on the instance by the user (else: a method)

        def output_of_pass(current_time, k):
            if k == 0:
                user_output(0., 0)
            elif k % inner_iterations == 0 and (k // inner_iterations) % frequency == 0:
                user_output((k // inner_iterations) * self.delta_t, k // inner_iterations)

        self.number_of_time_steps, self.output_frequency = outer * inner_iterations, 1
        self.time_step_output = output_of_pass
        try:
            self.start_solving()
        finally:
            self.number_of_time_steps, self.output_frequency = outer, frequency
            if assigned:
                self.time_step_output = user_output
            else:
                del self.__dict__["time_step_output"]

    def _solve_coupled_step(self, rhs_mesh):
        """One time step on a mesh with pointer couplings: K = K0(multipliers, accumulation diagonal) + E, solved by
        mfd.coupled_gmres around the hybridised solve of K0 (self._hyb, just set up for this step)."""
        import torch
        from ..mfd.mfd import coupled_gmres

        ctx, dm, plan, mfd = self.ctx, self._dm, self._plan, self.mfd
        vals = device.numeric_from_blocks(dm, plan, self._m_e, multipliers=self._d_mult, c_diag=self._d_cdiag)
        ip, ix, ev = self._coupling
        n = int(ip.numel() - 1)
        rows = self._mesh_rows

        def apply_k(v):
            if rows is None:
                y = device.spmv(ctx, plan.indptr, plan.indices, vals, v, n=plan.ndof)
            else:
                y = torch.zeros_like(v)
                y[rows] = device.spmv(ctx, plan.indptr, plan.indices, vals, v[rows].contiguous(), n=plan.ndof)
            return y + device.spmv(ctx, ip, ix, ev, v, n=n)

        with ctx.on_stream():
            if rows is None:
                rhs = rhs_mesh
            else:
                rhs = torch.zeros(n, dtype=torch.float64, device=ctx.device)
                rhs[rows] = rhs_mesh
        return coupled_gmres(ctx, apply_k, self._hyb, rhs, self.pcg_rtol, 50000, mult=mfd._mult)

    def time_step_output(self, current_time, time_step):
        pass

    def start_solving(self):
        """Implicit time stepping (singlephase.py:244-299)."""
        import torch

        ctx, dm, plan = self.ctx, self._dm, self._plan
        nc, F = dm.nc, plan.flux_dof
        self.time_step_output(0., 0)
        with ctx.on_stream():
            d_p = torch.as_tensor(np.asarray(self.current_pressure, dtype=float).copy(), device=ctx.device)
            d_phi = torch.as_tensor(np.asarray(self.porosities, dtype=float), device=ctx.device)
            wells = None
            if self.rate_wells:
                w = np.zeros(nc)
                for idx, cell in enumerate(self.rate_wells):
                    w[cell] += self.rate_wells_rate[idx]
                wells = torch.as_tensor(w, device=ctx.device)
            for time_step in range(1, self.number_of_time_steps + 1):
                rhs = self._d_rhs_mfd.clone()
                ctx.call("mimpy_b200_singlephase_terms", nc, F, float(self.ref_pressure), float(self.compressibility),
                         float(self.ref_density), float(self.viscosity), float(self.delta_t), _lib.ptr(d_p),
                         _lib.ptr(d_phi), _lib.ptr(dm.cell_volume), _lib.ptr(rhs), _lib.ptr(self._d_mult),
                         _lib.ptr(self._d_cdiag))
                if wells is not None:
                    rhs[F:] += wells
                if self._hyb is None:
                    self._hyb = device.HybridSystem(dm, plan, self._m_e, multipliers=self._d_mult, c_diag=self._d_cdiag)
                else:
                    self._hyb.setup(self._m_e, self._d_mult, self._d_cdiag)
                if self._coupling is not None:
                    sol, info = self._solve_coupled_step(rhs)
                    F_ref = self.mfd.flux_dof
                    self.last_solver_info = info
                    d_p = sol[F_ref:].clone()
                    self.current_pressure = device.to_host(ctx, d_p)
                    self.current_velocity = device.to_host(ctx, sol[:F_ref])
                    if time_step % self.output_frequency == 0:
                        self.time_step_output(time_step * self.delta_t, time_step)
                    continue
                # warm start: the face multipliers of the previous time step
                sol, info = self._hyb.solve(rhs, rtol=self.pcg_rtol, maxit=50000,
                                            x0=getattr(self._hyb, "last_lambda", None))
                self.last_solver_info = info
                d_p = sol[F:].clone()
                self.current_pressure = device.to_host(ctx, d_p)
                self.current_velocity = device.to_host(ctx, sol[:F])
                if time_step % self.output_frequency == 0:
                    self.time_step_output(time_step * self.delta_t, time_step)
