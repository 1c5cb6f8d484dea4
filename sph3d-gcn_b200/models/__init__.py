"""Model call graphs of the reference (models/SPH3D_*.py) written against this package's sph3gcn_util mirror.

SURVEY.md section 8(f) row N1: these exist to prove that the layer library drops in at the reference's call sites and
to time BASELINE.json configs[1..3] end to end; they hold no kernels of their own.

    SPH3D_modelnet.get_model(points, is_training, config) -> (logits (B, num_cls), end_points)
    SPH3D_s3dis.get_model(points, is_training, config)    -> (logits (B, N, num_cls), end_points)
    SPH3D_shapenet.get_model(points, num_cls, is_training, config)
    get_loss(...) as in the reference (S3DIS masks the loss with inner_label)
"""
from . import SPH3D_modelnet, SPH3D_s3dis, SPH3D_shapenet, configs   # noqa: F401
