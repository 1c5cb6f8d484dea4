from . import graph_step   # noqa: F401
