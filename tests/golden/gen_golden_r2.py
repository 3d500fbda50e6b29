"""Round-2 golden fixtures from the UNMODIFIED reference (see gen_golden.py for the conventions):

  twophase_pwell    6x3x3, 12 IMPES steps with pressure (BHP) wells, user relperm callables
                    (TwoPhase.set_krw / set_kro, twophase.py:292-300), compressible fluids
  twophase_c4_1233  BASELINE config 4 shape at 12x3x3, Corey relperms, rate wells

Usage (build container only):  python tests/golden/gen_golden_r2.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import ref_shim  # noqa: E402
from oracle import mimpy_oracle as orc  # noqa: E402
import gen_golden as g1  # noqa: E402
import twophase_cases as cases  # noqa: E402


def twophase_history(ref, name):
    nx, ny, nz, dims, dt, steps = cases.CASES[name]
    ktab = cases.permeability(name)
    mesh = g1.ref_hex(ref, nx, ny, nz, dims, ktab)
    tp = ref.twophase.TwoPhase()
    tp.set_mesh(mesh)
    cwd = os.getcwd()
    os.chdir("/tmp")          # the reference opens <well>.inj / <well>.prod in the cwd
    try:
        cases.configure(tp, mesh, name, "golden_")
        g1.quiet(tp.initialize_system)
        sw_hist, p_hist, u_hist = [], [], []
        for step in range(1, steps + 1):
            g1.quiet(tp.update_pressure, step)
            if step == 1 or step % 10 == 0:
                tp.find_upwinding_direction()
            for _ in range(tp.saturation_time_steps):
                tp.update_saturation(step)
            sw_hist.append(np.array(tp.current_s_w))
            p_hist.append(np.array(tp.current_p_o))
            u_hist.append(np.array(tp.current_u_t))
        for f in list(tp.production_files) + list(tp.pressure_files):
            f.flush()
        prod = {}
        for w, nme in enumerate(tp.pressure_wells_name):
            prod[nme] = np.loadtxt(nme + ".prod").reshape(-1, 3)
    finally:
        os.chdir(cwd)
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    d["delta_t"] = np.float64(dt)
    d["steps"] = np.int64(steps)
    d["s_w"], d["p_o"], d["u_t"] = np.array(sw_hist), np.array(p_hist), np.array(u_hist)
    if prod:
        d["prod_log"] = np.array([prod[n] for n in tp.pressure_wells_name])
    return d


def main():
    ref = ref_shim.load_reference()
    for name in cases.CASES:
        g1.save("twophase_" + name, twophase_history(ref, name))


if __name__ == "__main__":
    main()
