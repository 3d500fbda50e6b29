"""Single-phase, single-component transport with the API of mimpy.models.transport.Transport
(reference transport.py:15-410).

Every time step (transport.py:363-391) is the implicit pressure step of SinglePhase
(transport.py:261-301 repeats singlephase.py:256-291) followed by an explicit upwind update of the
concentration (find_upwinding_direction :346-361, update_concentration :303-344).  The pressure step
reuses mimpy_b200_singlephase_terms + the hybridised PCG; the concentration step is one kernel,
mimpy_b200_transport_step, which evaluates the upwind cell of every face directly instead of keeping
the reference's (face, cell) list.

Like the reference, the model addresses fluxes with plain face indices (transport.py:205-209,
276-277), i.e. it assumes that no boundary is a Neumann boundary; this is checked here.
"""
import numpy as np

from .. import _lib, device
from ..mfd import mfd as _mfd
from .singlephase import SinglePhase


class Transport(SinglePhase):
    _couplings_supported = False  # transport.py:209-239 is not built: initialize_system refuses such meshes

    def __init__(self):
        SinglePhase.__init__(self)
        self.prev_pressure = None
        self.current_concentration = None
        self.concentration_boundaries = {}
        self.upwinded_face_cell = []

    def set_mesh_mfd(self, mesh, mfd, ctx=None):
        """transport.py:86-92: the caller supplies the MFD object (and with it the M_E method)."""
        self.mesh = mesh
        self.mfd = mfd
        self.mfd.set_mesh(mesh, ctx)
        self.ctx = self.mfd.ctx

    def set_mesh(self, mesh, ctx=None):
        self.set_mesh_mfd(mesh, _mfd.MFD(), ctx)

    def set_initial_pressure(self, pressures):
        """transport.py:118-126"""
        self.initial_pressure = pressures
        self.current_pressure = pressures
        self.prev_pressure = pressures
        self.current_velocity = np.zeros(self.mesh.get_number_of_faces())

    def set_initial_concentration(self, concentration):
        self.current_concentration = concentration

    def set_concentration_boundaries(self, boundary_marker, concentration_function):
        """Concentration entering through the pressure (Dirichlet) boundary `boundary_marker`."""
        self.concentration_boundaries[boundary_marker] = concentration_function

    def initialize_system(self):
        """transport.py:186-250"""
        SinglePhase.initialize_system(self)
        if self._plan.flux_dof != self.mesh.get_number_of_faces():
            raise ValueError("Transport numbers fluxes by face (transport.py:205-209): Neumann boundaries are not supported")
        self._state_ready = False

    def _ensure_state(self):
        """Uploads pressures, concentration, porosities, wells and boundary concentrations once (they
        may be set before or after initialize_system, like in the reference)."""
        import torch

        if self._state_ready:
            return
        ctx, nf = self.ctx, self.mesh.get_number_of_faces()
        bc = np.full(nf, np.nan)
        for marker, fn in self.concentration_boundaries.items():
            for face_index, _ in self.mesh.get_boundary_faces_by_marker(marker):
                bc[face_index] = fn(self.mesh.get_face_real_centroid(face_index))
        with ctx.on_stream():
            self._d_bc = torch.as_tensor(bc, device=ctx.device) if self.concentration_boundaries else None
            self._d_phi = torch.as_tensor(np.asarray(self.porosities, dtype=float), device=ctx.device)
            self._d_p = torch.as_tensor(np.asarray(self.current_pressure, dtype=float).copy(), device=ctx.device)
            self._d_p_prev = torch.as_tensor(np.asarray(self.prev_pressure, dtype=float).copy(), device=ctx.device)
            self._d_c = torch.as_tensor(np.asarray(self.current_concentration, dtype=float).copy(), device=ctx.device)
            self._d_v = torch.as_tensor(np.asarray(self.current_velocity, dtype=float).copy(), device=ctx.device)
            self._d_wells = None
            if self.rate_wells:
                w = np.zeros(self._dm.nc)
                for idx, cell in enumerate(self.rate_wells):
                    w[cell] += self.rate_wells_rate[idx]
                self._d_wells = torch.as_tensor(w, device=ctx.device)
        self._state_ready = True

    def update_pressure(self):
        """transport.py:261-301"""
        self._ensure_state()
        ctx, dm, plan = self.ctx, self._dm, self._plan
        nc, F = dm.nc, plan.flux_dof
        with ctx.on_stream():
            rhs = self._d_rhs_mfd.clone()
            ctx.call("mimpy_b200_singlephase_terms", nc, F, float(self.ref_pressure), float(self.compressibility),
                     float(self.ref_density), float(self.viscosity), float(self.delta_t), _lib.ptr(self._d_p),
                     _lib.ptr(self._d_phi), _lib.ptr(dm.cell_volume), _lib.ptr(rhs), _lib.ptr(self._d_mult),
                     _lib.ptr(self._d_cdiag))
            if self._d_wells is not None:
                rhs[F:] += self._d_wells
            if self._hyb is None:
                self._hyb = device.HybridSystem(dm, plan, self._m_e, multipliers=self._d_mult, c_diag=self._d_cdiag)
            else:
                self._hyb.setup(self._m_e, self._d_mult, self._d_cdiag)
            # warm start: the face multipliers of the previous time step
            sol, info = self._hyb.solve(rhs, rtol=self.pcg_rtol, maxit=50000,
                                        x0=getattr(self._hyb, "last_lambda", None))
            self.last_solver_info = info
            self._d_p_prev = self._d_p
            self._d_p = sol[F:].clone()
            self._d_v = sol[:F].clone()
        self._host_stale = True

    def find_upwinding_direction(self):
        """transport.py:346-361.  The device step evaluates the upwind cell of each face on the fly;
        the reference's list is rebuilt lazily by `get_upwinded_face_cell` for callers that read it."""
        self.upwinded_face_cell = None

    def get_upwinded_face_cell(self):
        self._sync_host()
        out = []
        for cell_index in range(self.mesh.get_number_of_cells()):
            for face_index, orientation in zip(self.mesh.get_cell(cell_index),
                                               self.mesh.get_cell_normal_orientation(cell_index)):
                if orientation * self.current_velocity[face_index] > 0:
                    out.append([face_index, cell_index])
        return out

    def update_concentration(self):
        """transport.py:303-344"""
        import torch

        self._ensure_state()
        ctx, dm, plan = self.ctx, self._dm, self._plan
        with ctx.on_stream():
            mv = dm.view()
            c_new = torch.empty_like(self._d_c)
            ctx.call("mimpy_b200_transport_step", _lib.C.byref(mv), _lib.C.byref(plan.s), _lib.ptr(self._d_v),
                     _lib.ptr(self._d_c), _lib.ptr(self._d_bc), _lib.ptr(self._d_p_prev), _lib.ptr(self._d_p),
                     _lib.ptr(self._d_phi), float(self.ref_pressure), float(self.compressibility),
                     float(self.ref_density), float(self.delta_t), _lib.ptr(c_new))
            self._d_c = c_new
        self._host_stale = True

    def _sync_host(self):
        if getattr(self, "_host_stale", False):
            ctx = self.ctx
            self.prev_pressure = device.to_host(ctx, self._d_p_prev)
            self.current_pressure = device.to_host(ctx, self._d_p)
            self.current_velocity = device.to_host(ctx, self._d_v)
            self.current_concentration = device.to_host(ctx, self._d_c)
            self._host_stale = False

    def start_solving(self):
        """transport.py:363-391 (VTK output left to time_step_output)."""
        self.time_step_output(0., 0)
        for time_step in range(1, self.number_of_time_steps + 1):
            self.update_pressure()
            self.find_upwinding_direction()
            self.update_concentration()
            if time_step % self.output_frequency == 0:
                self._sync_host()
                self.time_step_output(time_step * self.delta_t, time_step)
        self._sync_host()
