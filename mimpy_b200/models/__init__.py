"""Mirror of mimpy.models (see mimpy_b200/__init__.py)."""
