"""Mirror of mimpy.mfd (see mimpy_b200/__init__.py)."""
