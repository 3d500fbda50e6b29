"""mimpy_b200: B200-native (sm_100a) implementation of mimpy's mimetic-finite-difference hot path.

Sub-packages mirror the reference's layout so that user scripts switch by changing one import:
    mimpy.mesh.hexmesh  -> mimpy_b200.mesh.hexmesh
    mimpy.mfd.mfd       -> mimpy_b200.mfd.mfd
    mimpy.models.*      -> mimpy_b200.models.*
The compute path is libmimpy_b200.so (hand-written CUDA behind the C-ABI in include/mimpy_b200.h);
there is no CPU fallback.
"""
__version__ = "0.1.0"
