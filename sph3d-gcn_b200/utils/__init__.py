from . import data_util, dist_util, graph_step   # noqa: F401
