"""Host-side mirrors of the reference's Python op wrappers (/root/reference/tf_ops/*/tf_*.py):
same module names, function names, argument order and defaults, on torch CUDA tensors."""
from . import tf_nnquery, tf_buildkernel, tf_conv3d, tf_sample, tf_pool3d, tf_unpool3d, tf_sepconv, tf_rowsgemm  # noqa: F401
