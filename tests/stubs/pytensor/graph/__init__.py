from . import basic, op, replace  # noqa: F401
