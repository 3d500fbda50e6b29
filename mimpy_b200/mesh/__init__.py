"""Mirror of mimpy.mesh (see mimpy_b200/__init__.py)."""
