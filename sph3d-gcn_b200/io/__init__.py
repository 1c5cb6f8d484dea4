"""On-disk format and input pipeline of the reference's segmentation trainers, without TensorFlow:
TFRecord container + tf.train.Example codec (tfrecord.py) and the S3DIS block pipeline (s3dis_input.py)."""
from . import tfrecord, s3dis_input   # noqa: F401
