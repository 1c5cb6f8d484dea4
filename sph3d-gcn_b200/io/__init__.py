"""On-disk format and input pipeline of the reference's segmentation trainers, without TensorFlow:
TFRecord container + tf.train.Example codec (tfrecord.py), the S3DIS block pipeline (s3dis_input.py) and the block-overlap
evaluation loop and the scene merge of its block predictions (s3dis_eval.py, s3dis_merge.py, 8(f) N4)."""
from . import tfrecord, s3dis_input, s3dis_eval, s3dis_merge, shapenet_input, modelnet_input   # noqa: F401
