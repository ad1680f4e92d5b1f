"""Import-time stub for the ``pptk`` point-cloud viewer (datasets/flyingthings3d_subset.py:4)."""


def viewer(*args, **kwargs):
    raise RuntimeError("pptk is not available in this environment (stub)")
