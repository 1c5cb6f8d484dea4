"""The reference's per-dataset configuration modules as plain objects (the reference imports a module of globals:
modelnet40_cls/modelnet_config.py, s3dis_seg/s3dis_config.py, shapenet_seg/shapenet_config.py).  `num_input` scales the
sampling pyramid so that small test clouds keep the same level structure."""
from types import SimpleNamespace


def _common(**kw):
    base = dict(kernel=[8, 2, 2], binSize=8 * 2 * 2 + 1, pool_method='max', unpool_method='mean', nnsearch='sphere',
                sample='FPS', with_bn=True, with_bias=False)
    base.update(kw)
    return SimpleNamespace(**base)


def modelnet(num_input=10000):
    """modelnet_config.py:3-37: levels = num_input / 4^(l+1) while above 100 points (2500, 625, 156 at 10000)."""
    levels = [num_input // 4 ** (i + 1) for i in range(10) if num_input // 4 ** (i + 1) > 100][:3]
    n = len(levels)
    return _common(num_input=num_input, num_cls=40, mlp=32, num_sample=levels, radius=[0.1, 0.2, 0.4][:n],
                   nn_uplimit=[64, 64, 64][:n], channels=[[64, 64], [64, 128], [128, 128]][:n],
                   multiplier=[[2, 1], [1, 2], [1, 1]][:n], global_channels=512, global_multiplier=2,
                   weight_decay=1e-5, normalize=True, use_raw=True)


def s3dis(num_input=8192):
    """s3dis_config.py:3-26 (2048, 768, 384, 128 at 8192 points)."""
    return _common(num_input=num_input, num_cls=13, mlp=64,
                   num_sample=[num_input // 4, num_input * 3 // 32, num_input * 3 // 64, num_input // 64],
                   radius=[0.1, 0.2, 0.4, 0.8], nn_uplimit=[64, 64, 64, 64],
                   channels=[[128, 128], [256, 256], [256, 256], [512, 512]],
                   multiplier=[[2, 2], [2, 2], [2, 2], [2, 2]], weight_decay=None, normalize=True)


def shapenet(num_input=2048, nn_uplimit=64):
    """shapenet_config.py:3-25 (1024, 768, 384, 128 at 2048 points); BASELINE.json configs[2] overrides K to 32."""
    return _common(num_input=num_input, mlp=64,
                   num_sample=[num_input // 2, num_input * 3 // 8, num_input * 3 // 16, num_input // 16],
                   radius=[0.08, 0.16, 0.32, 0.64], nn_uplimit=[nn_uplimit] * 4,
                   channels=[[128, 128], [256, 256], [256, 256], [512, 512]],
                   multiplier=[[2, 2], [2, 2], [2, 2], [2, 2]], weight_decay=None, normalize=False)
