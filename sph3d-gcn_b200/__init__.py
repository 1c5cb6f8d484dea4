"""sph3d-gcn_b200 -- B200-native (sm_100a) implementation of SPH3D-GCN's per-layer hot path.

The directory name carries a hyphen (it is the product name); import it through the
`sph3d_gcn_b200` shim at the repository root, or -- exactly like the reference, which is a tree of
scripts rather than a package -- put `sph3d-gcn_b200/utils` on sys.path and
`import sph3gcn_util as s3g_util`.

    tf_ops/            host mirrors of the reference's op wrappers (tf_ops/*/tf_*.py)
    utils/sph3gcn_util layer library with the reference's signatures (utils/sph3gcn_util.py)
    io/                TFRecord / tf.train.Example codec and the S3DIS block pipeline, without TensorFlow
    models/            the reference's model call graphs (models/SPH3D_*.py) on that layer library
    csrc/              hand-written CUDA kernels + the C ABI (include/sph3d_b200.h)
    lib/               built libsph3d_b200.so (git-ignored)

There is no CPU path: every op requires CUDA tensors and the compiled library.
"""
from . import build as _build_mod                      # noqa: F401
from . import _lib                                     # noqa: F401
from .tf_ops import tf_nnquery, tf_buildkernel, tf_conv3d, tf_sample, tf_pool3d, tf_unpool3d, tf_sepconv, tf_rowsgemm  # noqa: F401
from .utils import sph3gcn_util                        # noqa: F401
from . import models                                   # noqa: F401
from . import io                                       # noqa: F401

build = _build_mod.build
library_path = _lib.library_path
__all__ = ["tf_nnquery", "tf_buildkernel", "tf_conv3d", "tf_sample", "tf_pool3d", "tf_unpool3d", "tf_sepconv", "tf_rowsgemm",
           "sph3gcn_util", "models", "io", "build", "library_path"]
