from . import data_util, dist_util, graph_step, train_step   # noqa: F401
