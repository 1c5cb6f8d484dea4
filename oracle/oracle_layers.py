"""numpy float64 restatement of the TensorFlow graph nodes that surround the custom ops in the reference's layer library
(/root/reference/utils/sph3gcn_util.py): tf.matmul (:144-146, :203-205, :254-256), tf.nn.bias_add (:147-151), tf.nn.elu
(activation_fn default, :88-103) and tf.layers.batch_normalization(momentum=0.99, training=...) (:328-332).

TEST INFRASTRUCTURE ONLY (tests/): the checker for csrc/post.cu (sph3d_bias_act_bn[_grad]) and csrc/dense_*.cu
(sph3d_rows_gemm, sph3d_rows_wgrad).  PARITY UNPINNED: TensorFlow 1.12 is not in this image and the reference ships no golden vectors for
these nodes, so this module restates their published definitions --
  elu(x)   = x                    for x > 0,  exp(x) - 1 otherwise           (tensorflow/core/kernels/relu_op_functor.h)
  elu'(x)  = 1 resp. exp(x)       (EluGrad: (activations + 1) * gradients)
  batch normalisation over all axes but the last: mean, BIASED variance, y = gamma (x - mean) / sqrt(var + 1e-3) + beta,
  moving <- moving * 0.99 + batch * 0.01 while training; the moving statistics normalise otherwise
-- with the closed-form gradients written out (no autograd), and tests/test_host_logic_cpu.py checks those closed forms
against torch autograd of the same composition.
"""
import numpy as np

EPSILON, MOMENTUM = 1e-3, 0.99


def elu(z):
    return np.where(z > 0, z, np.expm1(np.minimum(z, 0)))


def elu_grad(z):
    return np.where(z > 0, 1.0, np.exp(np.minimum(z, 0)))


def bias_act_bn(x, bias=None, gamma=None, beta=None, moving_mean=None, moving_var=None, act=True, training=True):
    """x (R, C).  -> (out, new_moving_mean, new_moving_var, cache for the gradient)"""
    x = np.asarray(x, np.float64)
    z = x if bias is None else x + np.asarray(bias, np.float64)
    y = elu(z) if act else z
    if gamma is None:
        return y, None, None, (z, None, None, None)
    gamma, beta = np.asarray(gamma, np.float64), np.asarray(beta, np.float64)
    mm, mv = np.asarray(moving_mean, np.float64), np.asarray(moving_var, np.float64)
    if training:
        mean, var = y.mean(0), y.var(0)
        new_mm, new_mv = mm * MOMENTUM + mean * (1 - MOMENTUM), mv * MOMENTUM + var * (1 - MOMENTUM)
    else:
        mean, var, new_mm, new_mv = mm, mv, mm, mv
    invstd = 1.0 / np.sqrt(var + EPSILON)
    yhat = (y - mean) * invstd
    return gamma * yhat + beta, new_mm, new_mv, (z, yhat, invstd, gamma)


def bias_act_bn_grad(cache, grad_out, act=True, training=True, with_bias=True):
    """-> (grad_x, grad_bias or None, grad_gamma or None, grad_beta or None)"""
    z, yhat, invstd, gamma = cache
    g = np.asarray(grad_out, np.float64)
    R = g.shape[0]
    if yhat is None:
        dy, dgamma, dbeta = g, None, None
    else:
        dbeta, dgamma = g.sum(0), (g * yhat).sum(0)
        dy = gamma * invstd * (g - dbeta / R - yhat * dgamma / R) if training else gamma * invstd * g
    dz = dy * elu_grad(z) if act else dy
    return dz, (dz.sum(0) if with_bias else None), dgamma, dbeta


def dense(x, w):
    return np.asarray(x, np.float64) @ np.asarray(w, np.float64)


def dense_grad(x, w, grad_out):
    """-> (grad_x = g w^T, grad_w = x^T g)"""
    x, w, g = (np.asarray(t, np.float64) for t in (x, w, grad_out))
    return g @ w.T, x.T @ g
