"""Container-only import shim: re-export the multipledispatch copy vendored in sympy."""
from sympy.multipledispatch import dispatch, Dispatcher  # noqa: F401
