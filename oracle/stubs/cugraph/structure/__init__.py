from . import symmetrize  # noqa: F401
