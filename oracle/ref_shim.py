"""Import the UNMODIFIED reference (ohinai/mimpy) from /root/reference in THIS container.

TEST INFRASTRUCTURE ONLY (golden-vector generation and oracle pinning). The reference's
Python sources are never copied: a synthetic package `mimpy` is registered whose __path__
points at /root/reference/mimpy, and the three Cython natives are taken from oracle/_ref
(built by oracle/build_ref.py from the .pyx files where they lie).

Runtime compatibility shim -- the reference targets Python 2.6-3.4:
  1. scipy.diag was removed from SciPy; mfd.py:10, twophase.py:5, singlephase.py:6 import it.
  2. three true-division sites now yield floats used as array sizes:
     variable_array.new_size (mesh.py:96), Mesh._memory_extension (mesh.py:347),
     Mesh.add_point (mesh.py:318)  -> monkeypatched with // versions (no source edit).
  3. Mesh.output_vtk_mesh opens 'wb' then print(str) (mesh.py:2039) -> stubbed (output only).
"""
import importlib.machinery
import importlib.util
import os
import sys
import sysconfig
import types

REF = os.environ.get("MIMPY_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
NATIVE = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "mimpy"))


def load_native(name, dotted=None):
    """Load one oracle/_ref Cython native (works on the GPU box too: needs only numpy)."""
    dotted = dotted or name
    if dotted in sys.modules:
        return sys.modules[dotted]
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    path = os.path.join(NATIVE, name + suffix)
    if not os.path.exists(path):
        raise ImportError("oracle/_ref/%s%s missing: run python oracle/build_ref.py" % (name, suffix))
    loader = importlib.machinery.ExtensionFileLoader(dotted, path)
    spec = importlib.util.spec_from_file_location(dotted, path, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    loader.exec_module(mod)
    return mod


def natives_available():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(NATIVE, n + suffix))
               for n in ("mfd_cython", "hexmesh_cython", "mesh_cython"))


_loaded = None


def load_reference():
    """Returns the reference package (mimpy) with shims applied. Container-only."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError("reference sources not present at %s" % REF)
    import numpy as np
    import scipy

    if not hasattr(scipy, "diag"):
        scipy.diag = np.diag  # shim 1

    def _pkg(dotted, sub):
        m = types.ModuleType(dotted)
        m.__path__ = [os.path.join(REF, sub)]
        m.__package__ = dotted
        sys.modules[dotted] = m
        return m

    if "mimpy" in sys.modules and not getattr(sys.modules["mimpy"], "_is_ref_shim", False):
        raise ImportError("another 'mimpy' is already imported")
    root = _pkg("mimpy", "mimpy")
    root._is_ref_shim = True
    root.__version__ = "0.1.1"
    for sub in ("mesh", "mfd", "models"):
        setattr(root, sub, _pkg("mimpy." + sub, os.path.join("mimpy", sub)))

    load_native("mfd_cython", "mimpy.mfd.mfd_cython")
    load_native("hexmesh_cython", "mimpy.mesh.hexmesh_cython")
    load_native("mesh_cython", "mimpy.mesh.mesh_cython")

    import importlib

    mesh = importlib.import_module("mimpy.mesh.mesh")
    hexmesh = importlib.import_module("mimpy.mesh.hexmesh")
    mfd = importlib.import_module("mimpy.mfd.mfd")
    twophase = importlib.import_module("mimpy.models.twophase")
    singlephase = importlib.import_module("mimpy.models.singlephase")
    transport = importlib.import_module("mimpy.models.transport")

    # shim 2: integer sizes
    def new_size(self, size, minimum=1000):
        return max(size + size // 2 + 2, minimum)

    def _memory_extension(self, size):
        return size + size // 2 + 1

    def add_point(self, new_point):
        if self.number_of_points < len(self.points):
            self.points[self.number_of_points] = new_point
            self.number_of_points += 1
        else:
            new_array_size = (len(self.points) + len(self.points) // 2 + 1, 3)
            self.points.resize(new_array_size, refcheck=False)
            self.points[self.number_of_points] = new_point
            self.number_of_points += 1
        return self.number_of_points - 1

    mesh.variable_array.new_size = new_size
    mesh.Mesh._memory_extension = _memory_extension
    mesh.Mesh.add_point = add_point
    # shim 3: output only
    mesh.Mesh.output_vtk_mesh = lambda self, *a, **k: None

    ns = types.SimpleNamespace(mesh=mesh, hexmesh=hexmesh, mfd=mfd, twophase=twophase,
                               singlephase=singlephase, transport=transport, root=root)
    _loaded = ns
    return ns
