"""Two-phase, immiscible, slightly compressible flow (IMPES) with the API of
mimpy.models.twophase.TwoPhase (reference twophase.py:92-1092); the time loop lives on the GPU.

Per time step (twophase.py:703-742):
    update_pressure   Newton on the saddle system; M is rescaled by 1/lambda_t per cell (update_m,
                      mfd_cython.update_m_fast) inside the fused numeric assembly, the diagonal block
                      is the accumulation term, the linear solve is the hybridised PCG
    find_upwinding_direction (step 1 and every 10th)  -> mimpy_b200_twophase_upwind
    update_saturation -> mimpy_b200_twophase_saturation_step2
State arrays current_p_o / current_u_t / current_s_w are NumPy views refreshed after every step.
Relative permeabilities: Corey curves are evaluated inside the kernels; user callables given to
set_kro / set_krw (twophase.py:292-300) are evaluated on the host with NumPy, exactly as the reference
does, and the resulting per-cell mobilities are handed to the kernels.
Not covered: the PETSc solver branch, pointer couplings.
"""
import ctypes as C

import numpy as np

from .. import _lib, device
from ..mfd import mfd as _mfd


class TwoPhase(object):
    def __init__(self):
        self.mesh = None
        self.mfd = None
        self.initial_p_o = None
        self.previous_p_o = None
        self.current_p_o = None
        self.initial_u_t = None
        self.previous_u_t = None
        self.current_u_t = None
        self.initial_s_w = None
        self.current_s_w = None
        self.porosities = None
        self.newton_solution = None
        self.viscosity_water = None
        self.viscosity_oil = None
        self.ref_pressure_water = None
        self.ref_density_water = None
        self.compressibility_water = None
        self.ref_pressure_oil = None
        self.ref_density_oil = None
        self.compressibility_oil = None
        self.krw = None
        self.kro = None
        self._corey = None
        self.residual_saturation_water = None
        self.residual_saturation_oil = None
        self.saturation_boundaries = {}
        self.delta_t = 1.
        self.number_of_time_steps = 1
        self.saturation_time_steps = 1
        self.current_time = 0.
        self.rate_wells = []
        self.rate_wells_rate_water = []
        self.rate_wells_rate_oil = []
        self.rate_wells_name = []
        self.pressure_wells = []
        self.pressure_wells_bhp = []
        self.pressure_wells_pi = []
        self.pressure_wells_name = []
        self.newton_threshold = 1.e-6
        self.newton_step_max = 10
        self.output_frequency = 1
        self.prod_output_frequency = 1
        self.model_name = "model"
        self.solver = 0
        self.pcg_rtol = 1.e-12
        self.newton_history = []
        self.write_well_logs = False
        self.pressure_files = []      # <well>.inj, one per rate well      (twophase.py:419)
        self.production_files = []    # <well>.prod, one per pressure well (twophase.py:444)

    # ---- setters (twophase.py:198-474) -------------------------------------------------------
    def set_model_name(self, name):
        self.model_name = name

    def set_solver(self, tag):
        if tag == "PETSc":
            raise Exception("petsc4py not installed")

    def set_output_frequency(self, frequency):
        self.output_frequency = frequency

    def set_mesh(self, mesh, ctx=None):
        self.mesh = mesh
        self.mfd = _mfd.MFD()
        self.mfd.set_mesh(mesh, ctx)
        self.mfd.set_m_e_construction_method(6)  # twophase.py:231
        self.ctx = self.mfd.ctx

    def set_compressibility_water(self, v):
        self.compressibility_water = v

    def set_compressibility_oil(self, v):
        self.compressibility_oil = v

    def set_ref_density_oil(self, v):
        self.ref_density_oil = v

    def set_ref_density_water(self, v):
        self.ref_density_water = v

    def set_ref_pressure_oil(self, v):
        self.ref_pressure_oil = v

    def set_ref_pressure_water(self, v):
        self.ref_pressure_water = v

    def set_viscosity_water(self, v):
        self.viscosity_water = v

    def set_viscosity_oil(self, v):
        self.viscosity_oil = v

    def set_corey_relperm(self, n_o, n_w, k_rw_o=1.):
        """Corey-type relative permeabilities with clipping of the scaled saturation
        (twophase.py:268-291); evaluated inside the CUDA kernels."""
        self._corey = (float(n_o), float(n_w), float(k_rw_o))
        self.krw = lambda se: k_rw_o * (np.clip(se, 0., 1.) ** n_w)
        self.kro = lambda se: (1.0 - np.clip(se, 0., 1.)) ** n_o

    def set_kro(self, kro):
        """Oil relative permeability as a callable of the scaled saturation array (twophase.py:292-295).
        The callable runs on the host (NumPy) once per Newton iteration / saturation sub-step."""
        self.kro = kro
        self._corey = None

    def set_krw(self, krw):
        """Water relative permeability callable (twophase.py:297-300); see set_kro."""
        self.krw = krw
        self._corey = None

    def set_residual_saturation_water(self, v):
        self.residual_saturation_water = v

    def set_residual_saturation_oil(self, v):
        self.residual_saturation_oil = v

    def sefromsw(self, water_saturation):
        se = water_saturation - self.residual_saturation_water
        return se / (1. - self.residual_saturation_water - self.residual_saturation_oil)

    def water_mobility(self, water_saturation, water_pressure):
        mob = self.krw(self.sefromsw(np.asarray(water_saturation, dtype=float)))
        return mob / self.viscosity_water * self.ref_density_water * (1. + self.compressibility_water * water_pressure)

    def oil_mobility(self, water_saturation, oil_pressure):
        mob = self.kro(self.sefromsw(np.asarray(water_saturation, dtype=float)))
        return mob / self.viscosity_oil * self.ref_density_oil * (1. + self.compressibility_oil * oil_pressure)

    def set_initial_p_o(self, p_o):
        self.initial_p_o = p_o
        self.current_p_o = p_o

    def set_initial_u_t(self, u_t):
        raise Exception("deprecated use of intialize u_t")   # twophase.py:353-357: the reference refuses it as well

    def set_initial_s_w(self, s_w):
        self.initial_s_w = s_w
        self.current_s_w = s_w

    def set_cell_s_w(self, cell_index, s_w):
        self.initial_s_w[cell_index] = s_w
        self.current_s_w[cell_index] = s_w

    def set_porosities(self, porosities):
        self.porosities = porosities

    def set_cell_porosity(self, cell_index, porosity):
        self.porosities[cell_index] = porosity

    def apply_pressure_boundary_from_function(self, boundary_marker, p_function):
        self.mfd.apply_dirichlet_from_function(boundary_marker, p_function)

    def apply_flux_boundary_from_function(self, boundary_marker, f_function):
        self.mfd.apply_neumann_from_function(boundary_marker, f_function)

    def set_saturation_boundaries(self, boundary_marker, saturation_function):
        self.saturation_boundaries[boundary_marker] = saturation_function

    def add_rate_well(self, injection_rate_water, injection_rate_oil, cell_index, well_name):
        """Rate-specified point source in cell_index (twophase.py:407-420).  The reference opens
        <well_name>.inj here; logs are only written when write_well_logs is set."""
        self.rate_wells.append(cell_index)
        self.rate_wells_rate_water.append(injection_rate_water)
        self.rate_wells_rate_oil.append(injection_rate_oil)
        self.rate_wells_name.append(well_name)
        self.pressure_files.append(open(well_name + ".inj", "w") if self.write_well_logs else None)
        return len(self.rate_wells) - 1

    def add_pressure_well(self, bhp, PI, cell_index, well_name):
        self.pressure_wells.append(cell_index)
        self.pressure_wells_bhp.append(bhp)
        self.pressure_wells_pi.append(PI)
        self.pressure_wells_name.append(well_name)
        self.production_files.append(open(well_name + ".prod", "w") if self.write_well_logs else None)
        return len(self.pressure_wells) - 1

    def reset_wells(self):
        self.rate_wells, self.rate_wells_rate_water = [], []
        self.rate_wells_rate_oil, self.rate_wells_name = [], []
        self.pressure_wells, self.pressure_wells_bhp = [], []
        self.pressure_wells_pi, self.pressure_wells_name = [], []

    def get_well_rate_oil(self, well_index):
        return self.rate_wells_rate_oil[well_index]

    def get_well_rate_water(self, well_index):
        return self.rate_wells_rate_water[well_index]

    def set_time_step_size(self, delta_t):
        self.delta_t = delta_t

    def set_number_of_time_steps(self, n):
        self.number_of_time_steps = n

    def set_saturation_substeps(self, n):
        self.saturation_time_steps = n

    def time_step_output(self, current_time, time_step):
        pass

    # ---- device helpers -------------------------------------------------------------------------
    def _fluid(self):
        if self._corey is None and (self.krw is None or self.kro is None):
            raise Exception("relative permeabilities not set (set_corey_relperm, or set_krw and set_kro)")
        n_o, n_w, k0 = self._corey if self._corey is not None else (1., 1., 1.)   # unused with callables
        return device.make_fluid(self.viscosity_water, self.viscosity_oil, self.ref_density_water,
                                 self.ref_density_oil, self.compressibility_water, self.compressibility_oil,
                                 self.residual_saturation_water, self.residual_saturation_oil, n_w, n_o, k0)

    def _t(self, a, dtype=np.float64):
        import torch

        return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype), device=self.ctx.device)

    def _host_mobilities(self, d_s, d_p):
        """User relperm callables: (water, oil) mobilities per cell as device tensors, evaluated with
        NumPy from host copies of s_w and p (twophase.py:326-344)."""
        s = device.to_host(self.ctx, d_s)
        p = device.to_host(self.ctx, d_p)
        wm, om = self.water_mobility(s, p), self.oil_mobility(s, p)
        return self._t(wm), self._t(om)

    def _sync_host(self):
        H = lambda t: device.to_host(self.ctx, t)
        self.current_p_o = H(self._d_p)
        self.current_u_t = H(self._d_u)
        self.current_s_w = H(self._d_s)
        self.previous_p_o = H(self._d_p_prev)

    # ---- initialize_system (twophase.py:476-560) ---------------------------------------------------
    def initialize_system(self):
        import torch

        mfd = self.mfd
        mfd.set_mesh(self.mesh, self.ctx) if mfd.mesh is None else None
        if self.mesh.has_pointer_couplings():
            raise NotImplementedError("TwoPhase on a mesh with pointer couplings (twophase.py:950-955, 1039-1051): the IMPES "
                                      "pressure solve here is the hybridised PCG of the uncoupled saddle system; "
                                      "MFD.solve handles couplings (mfd.py), the time loop does not yet")
        dm, plan = mfd._ensure_plan()
        self._dm, self._plan = dm, plan
        ctx = self.ctx
        nc, F = dm.nc, plan.flux_dof
        self._m_e = mfd._blocks()                       # unscaled blocks: the reference's m_data_for_update
        self.m_x_coo_length = int(device.coo_view(dm, plan, None, want_triples=False)[0][-1].item())
        self.rhs_mfd = mfd.build_rhs()
        self._d_rhs_mfd = mfd.device_rhs
        with ctx.on_stream():
            self._d_p = self._t(np.asarray(self.current_p_o, dtype=float).copy())
            self._d_p_prev = self._d_p.clone()
            self._d_s = self._t(np.asarray(self.current_s_w, dtype=float).copy())
            self._d_u = torch.zeros(F, dtype=torch.float64, device=ctx.device)
            self._d_phi = self._t(self.porosities)
            self._d_fw = torch.zeros(F, dtype=torch.float64, device=ctx.device)
            self._d_upw = torch.full((F,), -1, dtype=torch.int32, device=ctx.device)
            # area of the face of every flux DOF: 8 B per face for the saturation update instead of a record sector
            self._d_flux_area = torch.empty(max(F, 1), dtype=torch.float64, device=ctx.device)
            mv0 = dm.view()
            ctx.call("mimpy_b200_flux_areas", C.byref(mv0), C.byref(plan.s), _lib.ptr(self._d_flux_area))
            self._d_vals = torch.empty(plan.nnz, dtype=torch.float64, device=ctx.device)
            self._d_invmob = torch.empty(nc, dtype=torch.float64, device=ctx.device)
            self._d_cdiag = torch.empty(nc, dtype=torch.float64, device=ctx.device)
            # rate wells: the reference looks the rate up with list.index(cell) (twophase.py:1025),
            # i.e. the FIRST well listed on a cell
            if self.rate_wells:
                first = {}
                for w, c in enumerate(self.rate_wells):
                    first.setdefault(c, w)
                self._d_wc = self._t(self.rate_wells, np.int32)
                self._d_wr = self._t([self.rate_wells_rate_water[first[c]] for c in self.rate_wells])
                tot = np.zeros(nc)
                for w, c in enumerate(self.rate_wells):
                    tot[c] += -self.rate_wells_rate_water[w]
                    tot[c] += -self.rate_wells_rate_oil[w]
                self._d_well_rhs = self._t(tot)
            else:
                self._d_wc = self._d_wr = self._d_well_rhs = None
            # saturation boundaries: f_w behind the boundary is static (twophase.py:974-994)
            bf, bs, bw = [], [], []
            for marker, sfun in self.saturation_boundaries.items():
                for (f, orient) in self.mesh.get_boundary_faces_by_marker(marker):
                    s = np.array([sfun(self.mesh.get_face_real_centroid(f))], dtype=float)
                    p = np.array([self.mfd.get_dirichlet_value(f)], dtype=float)
                    w = self.water_mobility(s, p)
                    bf.append(f)
                    bs.append(orient)
                    bw.append(float((w / (w + self.oil_mobility(s, p)))[0]))
            self._n_bnd = len(bf)
            self._d_bf = self._t(bf, np.int32) if bf else None
            self._d_bs = self._t(bs, np.int8) if bf else None
            self._d_bw = self._t(bw) if bf else None
        self._hyb = None
        self.current_u_t = np.zeros(F)
        self.initial_u_t = np.zeros(F)
        self.newton_solution = np.zeros(nc + F)
        self._fl = self._fluid()
        self.c_start = self.m_x_coo_length + 2 * int((mfd.face_to_flux[:, 0][self.mesh.cells.compact()[1]] >= 0).sum())
        self.c_end = self.c_start + nc

    # ---- update_pressure (twophase.py:744-926) ---------------------------------------------------------
    def update_pressure(self, time_step=0):
        import torch

        ctx, dm, plan = self.ctx, self._dm, self._plan
        nc, F = dm.nc, plan.flux_dof
        fl = self._fl
        mv = dm.view()
        with ctx.on_stream():
            po_k = self._d_p.clone()
            ut_k = self._d_u.clone()
            newton_residual = 100.
            newton_step = 0
            while abs(newton_residual > self.newton_threshold):
                if self._corey is not None:
                    ctx.call("mimpy_b200_twophase_mobility", nc, C.byref(fl), _lib.ptr(self._d_s), _lib.ptr(po_k),
                             _lib.ptr(self._d_invmob), None, None)
                else:
                    wm, om = self._host_mobilities(self._d_s, po_k)
                    torch.reciprocal(wm + om, out=self._d_invmob)
                x = torch.cat([ut_k, po_k])
                rhs = torch.empty(plan.ndof, dtype=torch.float64, device=ctx.device)
                # accumulation diagonal (c_diag) first; rhs terms after the matvec
                ctx.call("mimpy_b200_twophase_newton_terms", nc, F, C.byref(fl), float(self.delta_t),
                         _lib.ptr(self._d_s), _lib.ptr(self._d_p), _lib.ptr(self._d_phi),
                         _lib.ptr(dm.cell_volume), None, _lib.ptr(self._d_cdiag))
                for cell, pi in zip(self.pressure_wells, self.pressure_wells_pi):
                    self._d_cdiag[cell] += pi * 1. / self._d_invmob[cell]
                device.numeric(dm, plan, 6, multipliers=self._d_invmob, c_diag=self._d_cdiag, vals=self._d_vals)
                ctx.call("mimpy_b200_spmv", plan.ndof, _lib.ptr(plan.indptr), _lib.ptr(plan.indices),
                         _lib.ptr(self._d_vals), _lib.ptr(x), _lib.ptr(rhs))
                rhs -= self._d_rhs_mfd  # rhs = -build_rhs() + lhs.[u;p]        twophase.py:789-790
                ctx.call("mimpy_b200_twophase_newton_terms", nc, F, C.byref(fl), float(self.delta_t),
                         _lib.ptr(self._d_s), _lib.ptr(self._d_p), _lib.ptr(self._d_phi),
                         _lib.ptr(dm.cell_volume), _lib.ptr(rhs), None)
                if self._d_well_rhs is not None:
                    rhs[F:] += self._d_well_rhs
                for cell, bhp, pi in zip(self.pressure_wells, self.pressure_wells_bhp, self.pressure_wells_pi):
                    rhs[F + cell] -= pi * bhp * 1. / self._d_invmob[cell]
                newton_residual = float(torch.linalg.norm(rhs)) / float(plan.ndof)
                self.newton_history.append(newton_residual)
                if newton_residual > self.newton_threshold:
                    if self._hyb is None:
                        self._hyb = device.HybridSystem(dm, plan, self._m_e, multipliers=self._d_invmob,
                                                        c_diag=self._d_cdiag)
                    else:
                        self._hyb.setup(self._m_e, self._d_invmob, self._d_cdiag)
                    delta, info = self._hyb.solve(-rhs, rtol=self.pcg_rtol, maxit=50000)
                    self.last_solver_info = info
                    ut_k += delta[:F]
                    po_k += delta[F:]
                newton_step += 1
                if newton_step > self.newton_step_max:
                    raise ZeroDivisionError("Newton iteration did not converge (twophase.py:919-920)")
            self._d_p_prev = self._d_p
            self._d_p = po_k
            self._d_u_prev = self._d_u
            self._d_u = ut_k
        self.previous_u_t = self.current_u_t

    # ---- upwinding + saturation (twophase.py:928-1092) ----------------------------------------------------
    def find_upwinding_direction(self):
        ctx, dm, plan = self.ctx, self._dm, self._plan
        mv = dm.view()
        with ctx.on_stream():
            ctx.call("mimpy_b200_twophase_upwind", C.byref(mv), C.byref(plan.s), _lib.ptr(self._d_u), _lib.ptr(self._d_upw))
            self._d_fw.zero_()  # current_f_w = zeros (twophase.py:957)

    def update_saturation(self, time_step=0):
        ctx, dm, plan = self.ctx, self._dm, self._plan
        sat_delta_t = self.delta_t / float(self.saturation_time_steps)
        mv = dm.view()
        with ctx.on_stream():
            wm = om = None
            if self._corey is None:
                wm, om = self._host_mobilities(self._d_s, self._d_p)
            well = None
            if self.pressure_wells:
                # the well terms use the mobilities of the saturation BEFORE the update (twophase.py:965, 1059-1076)
                idx = self._t(self.pressure_wells, np.int64)
                well = (device.to_host(ctx, self._d_s[idx]), device.to_host(ctx, self._d_p[idx]), idx)
            ctx.call("mimpy_b200_twophase_saturation_step2", C.byref(mv), C.byref(plan.s), C.byref(self._fl),
                     float(sat_delta_t), _lib.ptr(self._d_upw), _lib.ptr(self._d_u), _lib.ptr(self._d_p),
                     _lib.ptr(self._d_p_prev), _lib.ptr(self._d_phi), self._n_bnd, _lib.ptr(self._d_bf),
                     _lib.ptr(self._d_bs), _lib.ptr(self._d_bw), len(self.rate_wells), _lib.ptr(self._d_wc),
                     _lib.ptr(self._d_wr), _lib.ptr(self._d_fw), _lib.ptr(self._d_s), _lib.ptr(wm), _lib.ptr(om),
                     _lib.ptr(self._d_flux_area))
            if well is not None:
                self._pressure_well_saturation(sat_delta_t, time_step, *well)

    def _pressure_well_saturation(self, dt, time_step, s_old, p, idx):
        """twophase.py:1053-1090, a handful of cells: evaluated with host scalars from the state before
        the saturation update; production logs as the reference writes them."""
        import torch

        wm = self.water_mobility(s_old, p)
        om = self.oil_mobility(s_old, p)
        add = np.zeros(len(self.pressure_wells))
        for w, cell in enumerate(self.pressure_wells):
            pi, bhp = self.pressure_wells_pi[w], self.pressure_wells_bhp[w]
            vol = self.mesh.get_cell_volume(cell)
            wp = pi * (bhp - p[w]) * dt * wm[w]
            wp = wp / self.porosities[cell] / self.ref_density_water / (1. + self.compressibility_water * p[w])
            wp /= vol
            op = pi * (bhp - p[w]) * dt * om[w]
            op = op / self.porosities[cell] / self.ref_density_oil / (1. + self.compressibility_oil * p[w])
            op /= vol
            out = self.production_files[w] if w < len(self.production_files) else None
            if out is not None and time_step % self.prod_output_frequency == 0:
                print(time_step, op * (-1), end=' ', file=out)
                print(wp * (-1), file=out)
            add[w] = wp
        # wells listed twice on a cell add twice (next_s_w[cell] += ..., twophase.py:1090)
        self._d_s.index_add_(0, idx, torch.as_tensor(add, device=self.ctx.device))

    def start_solving(self):
        """IMPES driver (twophase.py:703-742).  No VTK output."""
        for time_step in range(1, self.number_of_time_steps + 1):
            self.time_step = time_step
            self.update_pressure(time_step)
            if time_step == 1 or time_step % 10 == 0:
                self.find_upwinding_direction()
            for _ in range(self.saturation_time_steps):
                self.update_saturation(time_step)
            self._sync_host()
            if time_step % self.prod_output_frequency == 0:       # twophase.py:726-730
                for cell, out in zip(self.rate_wells, self.pressure_files):
                    if out is not None:
                        print(time_step, self.current_p_o[cell], end=' ', file=out)
                        print(self.current_s_w[cell], file=out)
            if time_step % self.output_frequency == 0:
                self.time_step_output(self.current_time, time_step)
            self.current_time = time_step * self.delta_t
