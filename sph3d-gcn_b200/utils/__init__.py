from . import dist_util, graph_step   # noqa: F401
