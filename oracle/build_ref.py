"""Build recipe for oracle/_ref: the reference's OWN native operators, compiled where they lie.

TEST INFRASTRUCTURE ONLY. Nothing under mimpy_b200/ may import this.

The reference (ohinai/mimpy, /root/reference) ships three Cython modules which are its native
operator boundary for the hot path:

    mimpy/mfd/mfd_cython.pyx      update_m_fast                  (mfd_cython.pyx:6-22)
    mimpy/mesh/hexmesh_cython.pyx all_face_areas/all_face_normals (hexmesh_cython.pyx:30-118)
    mimpy/mesh/mesh_cython.pyx    all_cell_volumes_centroids      (mesh_cython.pyx:5-207)

This script translates each .pyx with `cython -2` (the -2 is required: mesh_cython.pyx:165,191
use the py2 print statement) into a C file under a temp dir (NOT into the repo, NOT into
/root/reference which is read-only), and compiles it with gcc straight into oracle/_ref/*.so.
No reference source is copied into the repository; oracle/_ref/ is git-ignored but travels to
the GPU box, so the `-m gpu` tests can check the CUDA geometry/update kernels against the real
reference natives there.

The pure-Python part of the reference (mfd.py, mesh.py, hexmesh.py, twophase.py) cannot travel;
it is imported from /root/reference in this container only (oracle/ref_shim.py) to generate the
golden vectors in tests/golden/ (tests/golden/gen_golden.py).
"""
import os
import subprocess
import sys
import sysconfig
import tempfile

REF = os.environ.get("MIMPY_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

PYX = [
    ("mimpy/mfd/mfd_cython.pyx", "mfd_cython"),
    ("mimpy/mesh/hexmesh_cython.pyx", "hexmesh_cython"),
    ("mimpy/mesh/mesh_cython.pyx", "mesh_cython"),
]


def have_reference():
    return all(os.path.exists(os.path.join(REF, p)) for p, _ in PYX)


def built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(OUT, name + suffix)) for _, name in PYX)


def build(force=False, verbose=False):
    """Compile the three reference natives into oracle/_ref. Returns True when they exist."""
    if built() and not force:
        return True
    if not have_reference():
        return built()
    import numpy

    os.makedirs(OUT, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    pyinc = sysconfig.get_paths()["include"]
    npinc = numpy.get_include()
    with tempfile.TemporaryDirectory(prefix="mimpy_ref_build_") as tmp:
        for rel, name in PYX:
            cfile = os.path.join(tmp, name + ".c")
            cmd = [sys.executable, "-m", "cython", "-2", os.path.join(REF, rel), "-o", cfile]
            subprocess.run(cmd, check=True, capture_output=not verbose)
            so = os.path.join(OUT, name + suffix)
            cmd = ["gcc", "-O2", "-fPIC", "-shared", "-w", "-I", pyinc, "-I", npinc, cfile, "-o", so]
            subprocess.run(cmd, check=True, capture_output=not verbose)
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref built:", ok)
    sys.exit(0 if ok else 1)
