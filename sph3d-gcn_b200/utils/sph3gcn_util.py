"""Layer library with the signatures of the reference's utils/sph3gcn_util.py, on PyTorch.

Every public function keeps the reference's name, positional order, defaults and return arity
(/root/reference/utils/sph3gcn_util.py:20-332) so the call sites in models/SPH3D_*.py carry over
unchanged; tensors are torch CUDA tensors and the six custom ops are the sm_100a kernels behind
include/sph3d_b200.h.  TensorFlow-isms are re-hosted as follows:

  tf.variable_scope(scope, reuse) + tf.get_variable  -> a module-level VariableStore keyed by
      "scope/name"; a variable is created on first use and re-used afterwards (eager execution
      calls the layer function every step, so `reuse` is accepted and ignored);
  tf.add_to_collection('losses', ...)                 -> get_collection('losses') (cleared by the caller
      once per step with clear_collections()); the L2 terms of all variables arrive as one fused entry;
  tf.contrib.layers.xavier_initializer()              -> Glorot uniform with TF's fan computation;
  tf.layers.batch_normalization(momentum=0.99)        -> batch norm over the last axis, eps 1e-3,
      biased batch variance, moving statistics updated with 0.99 decay, L2 terms of beta/gamma in
      get_collection('regularization_losses');
  is_training                                         -> Python bool (or 0-dim tensor).

Order of operations in the conv layers is the reference's: matmul -> bias -> activation -> BN.
"""
import contextlib
from collections import OrderedDict

import torch
import torch.nn.functional as F

if __package__ in (None, ""):           # flat import from sys.path, the way the reference scripts do it
    import importlib
    import os
    import sys
    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.insert(0, _root)
    _pkg = importlib.import_module("sph3d_gcn_b200")
    sys.modules[__name__] = _pkg.utils.sph3gcn_util
else:
    from ..tf_ops import tf_conv3d, tf_pool3d, tf_unpool3d, tf_sepconv, tf_rowsgemm
    from . import layer_tail
    from ..tf_ops.tf_nnquery import build_sphere_neighbor, build_cube_neighbor
    from ..tf_ops.tf_sample import farthest_point_sample, inverse_density_sample, random_sample
    from ..tf_ops.tf_buildkernel import spherical_kernel

    neighbor_fn = build_sphere_neighbor  # default nn search method

    elu = F.elu


    # ------------------------------------------------------------------ variable store / scopes
    class VariableStore(object):
        def __init__(self):
            self.params = OrderedDict()       # name -> torch.nn.Parameter
            self.buffers = OrderedDict()      # name -> tensor (BN moving statistics)
            self.collections = {"losses": [], "regularization_losses": []}
            # L2 terms (weight decay, BN regularisers) per collection, keyed by the variable's identity: a variable registers
            # its term once however often the layer runs (TensorFlow adds the regulariser when the variable is created)
            self.pending_l2 = {"losses": {}, "regularization_losses": {}}
            self.scope = []

        def full_name(self, name):
            return "/".join(self.scope + [name])


    _STORE = VariableStore()


    def get_variable_store():
        return _STORE


    def reset_variables():
        """Drop every variable (a fresh tf.Graph)."""
        global _STORE
        _STORE = VariableStore()
        return _STORE


    def trainable_variables():
        return list(_STORE.params.values())


    def named_variables():
        return OrderedDict(_STORE.params)


    class _L2Sum(torch.autograd.Function):
        """sum_i scale_i * tf.nn.l2_loss(v_i) = sum_i scale_i * |v_i|^2 / 2 over a LIST of variables with multi-tensor
        kernels: two or three launches per direction instead of four small kernels per variable and direction."""
        @staticmethod
        def forward(ctx, scales, *variables):
            ctx.scales = scales
            ctx.save_for_backward(*variables)
            norms = torch.stack(torch._foreach_norm(list(variables), 2))
            return 0.5 * (norms * norms * _device_constant(scales, norms.device)).sum()

        @staticmethod
        def backward(ctx, g):
            grads = torch._foreach_mul(list(ctx.saved_tensors), list(ctx.scales))
            torch._foreach_mul_(grads, g)
            return (None,) + tuple(grads)


    _CONSTANTS = {}

    def _device_constant(values, device):
        """small read-only device vector, uploaded once (a per-step H2D copy could not be captured in a CUDA graph)"""
        key = (tuple(values), device.type, device.index)
        if key not in _CONSTANTS:
            _CONSTANTS[key] = torch.tensor(values, dtype=torch.float32, device=device)
        return _CONSTANTS[key]


    def get_collection(name):
        """tf.get_collection.  The per-variable L2 terms (weight decay, BN regularizers) registered since the last
        clear_collections() are materialised here as ONE fused term: only their sum is ever consumed
        (tf.add_n(tf.get_collection('losses')) in the reference's train scripts)."""
        items = list(_STORE.collections.get(name, []))
        pending = list((_STORE.pending_l2.get(name) or {}).values())
        if pending:                     # built per call and never stored: calling get_collection twice counts nothing twice
            items.append(_L2Sum.apply(tuple(sc for _, sc in pending), *[v for v, _ in pending]))
        return items


    def clear_collections():
        for v in _STORE.collections.values():
            del v[:]
        for v in _STORE.pending_l2.values():
            v.clear()


    @contextlib.contextmanager
    def variable_scope(scope, reuse=None):
        _STORE.scope.append(str(scope))
        try:
            yield scope
        finally:
            _STORE.scope.pop()


    def _fans(shape):
        # tf.contrib.layers.xavier_initializer / variance_scaling fan computation
        if len(shape) < 1:
            return 1.0, 1.0
        if len(shape) == 1:
            return float(shape[0]), float(shape[0])
        receptive = 1.0
        for d in shape[:-2]:
            receptive *= d
        return float(shape[-2]) * receptive, float(shape[-1]) * receptive


    def get_variable(name, shape, initializer, device, trainable=True):
        full = _STORE.full_name(name)
        table = _STORE.params if trainable else _STORE.buffers
        if full in table:
            return table[full]
        t = torch.empty(tuple(int(s) for s in shape), dtype=torch.float32, device=device)
        initializer(t)
        if trainable:
            t = torch.nn.Parameter(t)
        table[full] = t
        return t


    def _variable_with_weight_decay(name, shape, stddev, with_decay, use_xavier=True, device=None):
        """Initialized variable with optional L2 weight decay (sph3gcn_util.py:61-85)."""
        if use_xavier:
            fan_in, fan_out = _fans(shape)
            limit = (6.0 / (fan_in + fan_out)) ** 0.5

            def initializer(t):
                with torch.no_grad():
                    t.uniform_(-limit, limit)
        else:
            def initializer(t):
                torch.nn.init.trunc_normal_(t, mean=0.0, std=stddev, a=-2 * stddev, b=2 * stddev)
        var = get_variable(name, shape, initializer, device)
        if with_decay is not None:
            _STORE.pending_l2["losses"][id(var)] = (var, float(with_decay))            # tf.nn.l2_loss * decay
        return var


    def _zeros_variable(name, shape, device):
        return get_variable(name, shape, lambda t: t.zero_(), device)


    # ------------------------------------------------------------------ graph builders
    def build_global_graph(xyz, query, radius):
        nn_uplimit = xyz.shape[1]
        nn_idx, nn_cnt, nn_dst = neighbor_fn(xyz, query, radius=radius, nnsample=nn_uplimit)
        return nn_idx, nn_cnt, nn_dst


    _SIDE_STREAMS = {}

    def _side_stream(device):
        key = (device.type, device.index)
        if key not in _SIDE_STREAMS:
            _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
        return _SIDE_STREAMS[key]


    # Deferred sampling.  FPS is `num_sample` sequential rounds on one SM per cloud (1.2 ms for 8192 -> 2048 points), far
    # longer than the ball query it is enqueued next to, and nothing needs its result before the level's pooling step.
    # Inside `with async_sampling():` build_graph leaves the sampler (and the construction of `indices`) running on the
    # side stream and returns at once; the current stream joins when `indices` is first consumed through gather_nd --
    # the only way the reference's models consume it (tf.gather_nd, models/SPH3D_s3dis.py:68-72) -- or, at the latest,
    # when the `with` block ends.  Outside the block build_graph joins before it returns (safe for any consumer).
    _ASYNC_SAMPLING = [False]
    _PENDING_SAMPLES = []          # (event, indices tensor) not yet joined


    @contextlib.contextmanager
    def async_sampling():
        old = _ASYNC_SAMPLING[0]
        _ASYNC_SAMPLING[0] = True
        try:
            yield
        finally:
            _ASYNC_SAMPLING[0] = old
            if not old:
                join_pending_samples()


    def join_pending_samples():
        """make the current stream wait for every sampler still running on the side stream"""
        _PREFETCHED.clear()
        while _PENDING_SAMPLES:
            event, indices = _PENDING_SAMPLES.pop()
            cur = torch.cuda.current_stream(indices.device)
            cur.wait_event(event)
            indices.record_stream(cur)


    def _join_sample(indices):
        event = getattr(indices, "_sph3d_ready", None)
        if event is not None:
            cur = torch.cuda.current_stream(indices.device)
            cur.wait_event(event)
            indices.record_stream(cur)
            indices._sph3d_ready = None
            for i, (ev, _) in enumerate(_PENDING_SAMPLES):
                if ev is event:
                    del _PENDING_SAMPLES[i]
                    break


    # Sampling pyramid ahead of the network.  FPS of level l+1 needs only the xyz picked at level l -- no features -- so the
    # whole chain (8192 -> 2048 -> 768 -> 384 -> 128 in SPH3D_s3dis: 3 328 sequential rounds, 1.5 ms on one SM per cloud) can
    # run on the side stream from the first instant instead of level by level behind the convolutions.  prefetch_samples
    # enqueues it; build_graph(xyz_l, ..., sample_method='FPS') then finds the picks of ITS xyz tensor instead of launching
    # FPS, and gather_nd(xyz_l, indices_l) returns the coarse cloud the chain already gathered (same values bit for bit).
    _PREFETCHED = {}               # id(xyz tensor) -> (xyz tensor, num_sample, indices)


    def prefetch_samples(xyz, num_samples, sample_method='FPS'):
        """start FPS for every level of the pyramid rooted at `xyz` on the side stream; no-op for other samplers / CPU"""
        _PREFETCHED.clear()
        if sample_method != 'FPS' or not xyz.is_cuda or not _ASYNC_SAMPLING[0]:
            return
        cur = torch.cuda.current_stream(xyz.device)
        side = _side_stream(xyz.device)
        side.wait_stream(cur)
        batch_size = xyz.shape[0]
        level_xyz = xyz
        with torch.cuda.stream(side):
            for num_sample in num_samples:
                if num_sample is None or num_sample <= 1 or num_sample > level_xyz.shape[1]:
                    break
                sample_index = farthest_point_sample(num_sample, level_xyz)
                indices = _sample_indices(sample_index, batch_size, num_sample)
                coarse = level_xyz[indices[..., 0].long(), indices[..., 1].long()].contiguous()
                event = torch.cuda.Event()
                event.record(side)
                indices._sph3d_ready = event
                indices._sph3d_src = level_xyz
                indices._sph3d_coarse = coarse
                _PENDING_SAMPLES.append((event, indices))
                _PREFETCHED[id(level_xyz)] = (level_xyz, int(num_sample), indices)
                level_xyz = coarse
        xyz.record_stream(side)


    def _sample_indices(sample_index, batch_size, num_sample):
        batch_indices = torch.arange(batch_size, device=sample_index.device, dtype=sample_index.dtype)
        batch_indices = batch_indices.view(-1, 1, 1).expand(-1, int(num_sample), 1)
        return torch.cat([batch_indices, sample_index.unsqueeze(2)], dim=2)       # (B,S,2) = [batch, point]


    def build_graph(xyz, radius, nn_uplimit, num_sample, sample_method=None):
        # FPS depends only on xyz and occupies one SM per cloud for `num_sample` sequential rounds, so it is
        # enqueued FIRST on a side stream and overlaps the ball query (which fills the other SMs); the main
        # stream waits for it before this function returns (or later, see async_sampling).  Same results, shorter
        # critical path.
        fps_event = None
        batch_size = xyz.shape[0]
        ahead = _PREFETCHED.get(id(xyz)) if (num_sample is not None and sample_method == 'FPS') else None
        if ahead is not None and ahead[0] is xyz and ahead[1] == int(num_sample):
            intra_idx, intra_cnt, intra_dst = neighbor_fn(xyz, xyz, radius=radius, nnsample=nn_uplimit)
            return intra_idx, intra_cnt, intra_dst, ahead[2]            # picks already on their way (prefetch_samples)
        if num_sample is not None and sample_method == 'FPS' and xyz.is_cuda:
            cur = torch.cuda.current_stream(xyz.device)
            side = _side_stream(xyz.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                sample_index = farthest_point_sample(num_sample, xyz)
                indices = _sample_indices(sample_index, batch_size, num_sample)
                fps_event = torch.cuda.Event()
                fps_event.record(side)
            xyz.record_stream(side)

        intra_idx, intra_cnt, intra_dst = neighbor_fn(xyz, xyz, radius=radius, nnsample=nn_uplimit)

        if num_sample is not None:
            if fps_event is not None:
                if _ASYNC_SAMPLING[0]:
                    indices._sph3d_ready = fps_event
                    _PENDING_SAMPLES.append((fps_event, indices))
                else:
                    cur.wait_event(fps_event)
                    indices.record_stream(cur)
                return intra_idx, intra_cnt, intra_dst, indices
            if sample_method == 'random':
                sample_index = random_sample(num_sample, xyz)
            elif sample_method == 'FPS':
                sample_index = farthest_point_sample(num_sample, xyz)
            elif sample_method == 'IDS':
                prob = intra_dst.sum(dim=-1) / intra_cnt.to(torch.float32)
                sample_index = inverse_density_sample(num_sample, prob)
            else:
                raise ValueError('Unknown sampling method.')
            indices = _sample_indices(sample_index, batch_size, num_sample)
        else:
            indices = None

        return intra_idx, intra_cnt, intra_dst, indices


    def build_graph_deconv(xyz, xyz_unpool, radius, nn_uplimit):
        intra_idx, intra_cnt, intra_dst = neighbor_fn(xyz, xyz, radius=radius, nnsample=nn_uplimit)
        inter_idx, inter_cnt, inter_dst = neighbor_fn(xyz, xyz_unpool, radius=radius, nnsample=nn_uplimit)
        return intra_idx, intra_cnt, intra_dst, inter_idx, inter_cnt, inter_dst


    def gather_nd(params, indices):
        """tf.gather_nd(params, indices) for the (B,S,2) = [batch, point] indices build_graph returns:
        the row selection the models apply to xyz / intra_idx / intra_cnt / intra_dst
        (models/SPH3D_s3dis.py:68-72)."""
        _join_sample(indices)
        if getattr(indices, "_sph3d_src", None) is params:                # the prefetched chain gathered this cloud already
            coarse = indices._sph3d_coarse
            coarse.record_stream(torch.cuda.current_stream(coarse.device))
            return coarse
        b = indices[..., 0].long()
        s = indices[..., 1].long()
        return params[b, s].contiguous()


    # ------------------------------------------------------------------ layers
    # The pointwise product (tf.matmul over the B*M rows, sph3gcn_util.py:144-146) and its two gradients run on hand-written
    # tcgen05 kernels at fp32 accuracy: y = x w and gx = g w^T through csrc/rowsgemm.cu (the weights re-packed per call into
    # the tensor core's operand image, the rows cross HBM once as fp32 and are split into three bf16 terms in registers),
    # gw = x^T g through csrc/rowswgrad.cu (both operands MN-major, one CTA per 128 x 128 block of gw and slab of rows,
    # partial blocks summed in slab order: deterministic).  Measured on the layer shapes of the three networks
    # (profiles/r2_rowsgemm.json): y / gx 1.3-1.9x the CUTLASS 9xBF16 collective instantiation of round 1 and 2-2.8x the
    # fp32 library GEMM from 3 072 rows up; gw 1.5-1.9x the split-K CUTLASS form and 3-5x the library GEMM (which runs the
    # 8-64 output tiles of x^T g on as many SMs without splitting the sum).  Anything the kernels do not take -- channel
    # counts that are not multiples of 4 (the 3-channel input layer, the 13 / 50-class logits), a few hundred rows, CPU
    # tensors -- stays on the library GEMM, the weight gradient then as an explicit split over row slabs.
    ROWS_GEMM = True
    ROWS_GEMM_MIN_ROWS = 1024
    ROWS_WGRAD_MIN_ROWS = 128      # x^T g over a few hundred rows: 32x32-tile library kernels take ~50 us (one cloud per rank)
    ROWS_GEMM_TERMS = 3            # 3: six cross products (error 1-2e-6 of the terms); 2: four (2-5e-6), ~1.25x faster
    SPLIT_K_WEIGHT_GRAD = True     # library fallback of the weight gradient: batched GEMM over row slabs + ordered sum
    _SPLIT_K_MIN_ROWS = 4096


    def _rows_ok(R, K, N, *tensors, min_rows=None):
        return (ROWS_GEMM and R >= (ROWS_GEMM_MIN_ROWS if min_rows is None else min_rows) and K % 4 == 0 and N % 4 == 0 and
                all(t.is_cuda and t.dtype == torch.float32 and t.data_ptr() % 16 == 0 for t in tensors))


    def _rows_gemm(x, w, trans, image=None):
        """x @ w (trans False) or x @ w.T (trans True) through sph3d_rows_gemm; None when the shape is not covered"""
        R, K = x.shape
        N = w.shape[0] if trans else w.shape[1]
        if not _rows_ok(R, K, N, x, w):
            return None
        return tf_rowsgemm.rows_gemm(x, w, trans=trans, terms=ROWS_GEMM_TERMS, image=image)


    def _weight_grad(x, g):
        """x (R, Cin), g (R, Cout) -> x^T g (Cin, Cout)"""
        R, cin = x.shape
        cout = g.shape[1]
        if _rows_ok(R, cin, cout, x, g, min_rows=ROWS_WGRAD_MIN_ROWS):
            return tf_rowsgemm.rows_wgrad(x, g, terms=ROWS_GEMM_TERMS)
        if not SPLIT_K_WEIGHT_GRAD or R < _SPLIT_K_MIN_ROWS:
            return x.t() @ g
        tiles = ((cin + 63) // 64) * ((cout + 63) // 64)
        sms = torch.cuda.get_device_properties(x.device).multi_processor_count if x.is_cuda else 148
        slabs = min(max((2 * sms + tiles - 1) // tiles, 1), R // 512)
        if slabs <= 1:
            return x.t() @ g
        rows = R // slabs
        main = rows * slabs
        out = torch.bmm(x[:main].view(slabs, rows, cin).transpose(1, 2), g[:main].view(slabs, rows, cout)).sum(dim=0)
        if main < R:
            out = out + x[main:].t() @ g[main:]
        return out


    # gx = g w^T and gw = x^T g of one layer depend on g only: the weight gradient is enqueued on a second stream and the
    # current stream joins after it has enqueued the input gradient, so the two kernels run side by side.  Deep levels
    # (1 000-6 000 rows) give each of them a few dozen CTAs and ~25 us of serial chunk chain; together they cost one chain.
    # Stream-ordered, so a captured step (GraphedStep) gets a fork / join in its graph.
    OVERLAP_WEIGHT_GRAD = True
    _GRAD_STREAMS = {}


    def _grad_stream(device):
        key = (device.type, device.index)
        if key not in _GRAD_STREAMS:
            _GRAD_STREAMS[key] = torch.cuda.Stream(device=device)
        return _GRAD_STREAMS[key]


    class _Dense(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):
            ctx.save_for_backward(x, w)
            ctx.image_t = None
            R, K = x.shape
            if ctx.needs_input_grad[0] and _rows_ok(R, K, w.shape[1], x, w):
                # the weights' operand images of both orientations in one launch: g w^T in backward finds its own
                image, ctx.image_t = tf_rowsgemm.pack_pair(w)
                return _rows_gemm(x, w, False, image)
            y = _rows_gemm(x, w, False)
            return y if y is not None else x @ w

        @staticmethod
        def backward(ctx, g):
            x, w = ctx.saved_tensors
            g = g.contiguous()
            gx = gw = None
            want_gx, want_gw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
            overlap = OVERLAP_WEIGHT_GRAD and want_gx and want_gw and g.is_cuda
            if overlap:
                cur, side = torch.cuda.current_stream(g.device), _grad_stream(g.device)
                side.wait_stream(cur)                      # g (and x) are ready where the current stream stands
                with torch.cuda.stream(side):
                    gw = _weight_grad(x, g)
                x.record_stream(side)
                g.record_stream(side)
            if want_gx:
                gx = _rows_gemm(g, w, True, ctx.image_t)
                if gx is None:
                    gx = g @ w.t()
            if overlap:
                cur.wait_stream(side)
                gw.record_stream(cur)
            elif want_gw:
                gw = _weight_grad(x, g)
            return gx, gw


    def _dense(x2d, w):
        return _Dense.apply(x2d.contiguous(), w.contiguous())


    # The separable layer as one kernel (csrc/sepconv.cu, row N2 of SURVEY.md section 8f): the depthwise result goes from
    # the gathering warps straight into the tensor core's operand and the pointwise product (and, without gradients, the
    # whole bias -> ELU -> folded-BN tail) happens before anything is written.  With gradients the depthwise result is
    # also kept (the weight gradient needs it) and the raw product feeds the training-mode layer tail.  Measured
    # (profiles/r2_sepconv.json, B200): without gradients the fused layer is 1.2-1.3x faster than the composition
    # (0.79 vs 1.01 ms at B=32, N=10^4, K=64, C=128 -> 128); with gradients it is 3-8 % slower than depthwise + product as
    # two kernels (0.78 vs 0.73 ms: the kernel is issue-bound and the operand staging adds instructions to the gather
    # loop's warps), so training keeps the composition unless FUSE_SEPARABLE_TRAINING is set.
    FUSE_SEPARABLE = True
    FUSE_SEPARABLE_TRAINING = False


    class _SeparableFused(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inputs, depthwise_kernel, kernel, nn_index, nn_count, filt_index):
            out, dw = tf_sepconv.separable_conv3d(inputs, depthwise_kernel, kernel, nn_index, nn_count, filt_index,
                                                  keep_depthwise=True)
            ctx.save_for_backward(inputs, depthwise_kernel, kernel, dw)
            ctx.graph = (nn_index, nn_count, filt_index)      # the tensor OBJECTS (a shared plan hangs off filt_index)
            return out

        @staticmethod
        def backward(ctx, g):
            inputs, depthwise_kernel, kernel, dw = ctx.saved_tensors
            B, M, cout = g.shape
            kp = kernel.shape[0]
            g2 = g.contiguous().reshape(-1, cout)
            gw = _weight_grad(dw.reshape(-1, kp), g2) if ctx.needs_input_grad[2] else None
            gi = gf = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                gdw = _rows_gemm(g2, kernel, True)
                if gdw is None:
                    gdw = g2 @ kernel.t()
                gi, gf = tf_conv3d.backward_on_graph(inputs, depthwise_kernel, gdw.reshape(B, M, kp), ctx.graph)
            return gi, gf, gw, None, None, None


    def _fused_separable(inputs, depthwise_kernel, kernel, nn_index, nn_count, filt_index, num_out_channels,
                         activation_fn, with_bn, with_bias, reuse, is_training):
        """the fused route of separable_conv3d, or None when the composition has to run"""
        if not (FUSE_SEPARABLE and inputs.is_cuda and tf_sepconv.supported(inputs, depthwise_kernel, nn_index,
                                                                         num_out_channels)):
            return None
        inputs, depthwise_kernel, kernel = inputs.contiguous(), depthwise_kernel.contiguous(), kernel.contiguous()
        needs_grad = torch.is_grad_enabled() and (inputs.requires_grad or depthwise_kernel.requires_grad or
                                                  kernel.requires_grad)
        tail_folds = (activation_fn is None or activation_fn is elu) and not (with_bn and _as_bool(is_training))
        if (needs_grad or not (FUSED_TAIL and tail_folds)) and not FUSE_SEPARABLE_TRAINING:
            return None
        if needs_grad or not (FUSED_TAIL and tail_folds):
            if needs_grad:
                outputs = _SeparableFused.apply(inputs, depthwise_kernel, kernel, nn_index, nn_count, filt_index)
            else:
                outputs, _ = tf_sepconv.separable_conv3d(inputs, depthwise_kernel, kernel, nn_index, nn_count, filt_index)
            return _post(outputs, num_out_channels, activation_fn, with_bn, with_bias, reuse, is_training)
        # inference: the whole tail rides in the epilogue (variables are created in _post's order)
        biases = _zeros_variable('biases', [num_out_channels], inputs.device) if with_bias else None
        scale = shift = None
        if with_bn:
            gamma, beta, moving_mean, moving_var = _bn_variables(num_out_channels, inputs.device, 'bn', reuse)
            scale = gamma * torch.rsqrt(moving_var + BN_EPSILON)
            shift = beta - moving_mean * scale
        act = tf_sepconv.ACT_ELU if activation_fn is elu else tf_sepconv.ACT_NONE
        outputs, _ = tf_sepconv.separable_conv3d(inputs, depthwise_kernel, kernel, nn_index, nn_count, filt_index,
                                                 bias=biases, scale=scale, shift=shift, act=act)
        return outputs


    # bias -> activation -> BN run as ONE op (csrc/post.cu) whenever the activation is the library's elu or None;
    # any other callable keeps the node-by-node composition.  FUSED_TAIL = False forces the composition (A/B runs).
    FUSED_TAIL = True


    def _post(outputs, num_out_channels, activation_fn, with_bn, with_bias, reuse, is_training):
        fused = FUSED_TAIL and outputs.is_cuda and (activation_fn is None or activation_fn is elu) and \
            (with_bias or with_bn or activation_fn is not None)
        biases = _zeros_variable('biases', [num_out_channels], outputs.device) if with_bias else None
        if fused:
            act = layer_tail.ACT_ELU if activation_fn is elu else layer_tail.ACT_NONE
            if with_bn:
                gamma, beta, moving_mean, moving_var = _bn_variables(num_out_channels, outputs.device, 'bn', reuse)
                return layer_tail.bias_act_bn(outputs, biases, gamma, beta, moving_mean, moving_var, act=act,
                                              training=_as_bool(is_training), eps=BN_EPSILON, momentum=BN_MOMENTUM)
            return layer_tail.bias_act_bn(outputs, biases, act=act)
        if with_bias:
            outputs = outputs + biases
        if activation_fn is not None:
            outputs = activation_fn(outputs)
        if with_bn:
            outputs = _batch_normalization(outputs, is_training, 'bn', reuse, fused=False)
        return outputs


    def separable_conv3d(inputs,
                         num_out_channels,
                         kernel_size,
                         depth_multiplier,
                         scope,
                         nn_index,
                         nn_count,
                         filt_index,
                         use_xavier=True,
                         stddev=1e-3,
                         weight_decay=None,
                         activation_fn=elu,
                         with_bn=False,
                         with_bias=False,
                         reuse=None,
                         is_training=None):
        """ 3D separable convolution with non-linear operation (sph3gcn_util.py:88-163).
            inputs BxNxC -> depthwise spherical conv (BxMxC*r) -> pointwise matmul -> B x M x num_out_channels
        """
        with variable_scope(scope, reuse=reuse):
            num_in_channels = inputs.shape[-1]
            depthwise_kernel_shape = [kernel_size, num_in_channels, depth_multiplier]
            depthwise_kernel = _variable_with_weight_decay('depthwise_weights', shape=depthwise_kernel_shape,
                                                           use_xavier=use_xavier, stddev=stddev,
                                                           with_decay=weight_decay, device=inputs.device)
            # Channel counts that are not a multiple of 4 (ModelNet's 32+3, 64+3, 128+3 after the raw-xyz concat)
            # would take the scalar-strip kernels; zero-padding the channel axis of the input and of both weight
            # tensors is exact (zeros contribute nothing, gradients flow through the pads) and keeps the op on
            # the 16-byte-vector kernels.  The VARIABLES keep the reference's shapes.
            pad = (-num_in_channels) % 4
            if pad:
                inputs, depthwise_kernel = F.pad(inputs, (0, pad)), F.pad(depthwise_kernel, (0, 0, 0, pad))
            batch_size = inputs.shape[0]
            num_in_channels = num_in_channels * depth_multiplier
            kernel = _variable_with_weight_decay('weights', shape=[num_in_channels, num_out_channels],
                                                 use_xavier=use_xavier, stddev=stddev,
                                                 with_decay=weight_decay, device=inputs.device)
            if pad:
                kernel = F.pad(kernel, (0, 0, 0, pad * depth_multiplier))
                num_in_channels += pad * depth_multiplier
            fused = _fused_separable(inputs, depthwise_kernel, kernel, nn_index, nn_count, filt_index, num_out_channels,
                                     activation_fn, with_bn, with_bias, reuse, is_training)
            if fused is not None:
                return fused
            outputs = tf_conv3d.depthwise_conv3d(inputs, depthwise_kernel, nn_index, nn_count, filt_index)
            outputs = _dense(outputs.reshape(-1, num_in_channels), kernel)
            outputs = outputs.reshape(batch_size, -1, num_out_channels)
            return _post(outputs, num_out_channels, activation_fn, with_bn, with_bias, reuse, is_training)


    def pointwise_conv3d(inputs,
                         num_out_channels,
                         scope,
                         use_xavier=True,
                         stddev=1e-3,
                         weight_decay=None,
                         activation_fn=elu,
                         with_bn=False,
                         with_bias=False,
                         reuse=None,
                         is_training=None):
        """ pointwise convolution with non-linear operation (sph3gcn_util.py:166-222). """
        with variable_scope(scope, reuse=reuse):
            batch_size = inputs.shape[0]
            num_in_channels = inputs.shape[-1]
            kernel = _variable_with_weight_decay('weights', shape=[num_in_channels, num_out_channels],
                                                 use_xavier=use_xavier, stddev=stddev,
                                                 with_decay=weight_decay, device=inputs.device)
            outputs = _dense(inputs.reshape(-1, num_in_channels), kernel)
            outputs = outputs.reshape(batch_size, -1, num_out_channels)
            return _post(outputs, num_out_channels, activation_fn, with_bn, with_bias, reuse, is_training)


    def fully_connected(inputs,
                        num_out_channels,
                        scope,
                        use_xavier=True,
                        stddev=1e-3,
                        weight_decay=None,
                        activation_fn=elu,
                        with_bn=False,
                        with_bias=False,
                        reuse=None,
                        is_training=None):
        """ Fully connected layer with non-linear operation (sph3gcn_util.py:225-273). inputs BxC. """
        with variable_scope(scope, reuse=reuse):
            num_in_channels = inputs.shape[-1]
            kernel = _variable_with_weight_decay('weights', shape=[num_in_channels, num_out_channels],
                                                 use_xavier=use_xavier, stddev=stddev,
                                                 with_decay=weight_decay, device=inputs.device)
            outputs = _dense(inputs, kernel)
            return _post(outputs, num_out_channels, activation_fn, with_bn, with_bias, reuse, is_training)


    def pool3d(inputs, nn_index, nn_count, scope, method):
        """ 3D pooling (sph3gcn_util.py:276-297). """
        with variable_scope(scope):
            if method == 'max':
                outputs, max_index = tf_pool3d.max_pool3d(inputs, nn_index, nn_count)
            elif method == 'avg':
                outputs = tf_pool3d.avg_pool3d(inputs, nn_index, nn_count)
            else:
                raise ValueError("Unknow pooling method %s." % method)
            return outputs


    def unpool3d(inputs, nn_index, nn_count, nn_dist, scope, method):
        """ 3D unpooling (sph3gcn_util.py:300-325). """
        with variable_scope(scope):
            if method == 'mean':
                outputs = tf_unpool3d.mean_interpolate(inputs, nn_index, nn_count)
            elif method == 'weighted':
                sum_nn_dist = nn_dist.sum(dim=-1, keepdim=True)
                epsilon = 1e-7
                weight = (nn_dist + epsilon) / (sum_nn_dist + epsilon)
                outputs = tf_unpool3d.weighted_interpolate(inputs, weight, nn_index, nn_count)
            else:
                raise ValueError("Unknow unpooling method %s." % method)
            return outputs


    BN_MOMENTUM, BN_EPSILON = 0.99, 1e-3


    def _as_bool(is_training):
        return bool(is_training) if is_training is not None else False


    def _bn_variables(C, device, name, reuse):
        """gamma / beta / moving statistics of one tf.layers.batch_normalization node + its L2 regularizers."""
        with variable_scope(name, reuse=reuse):
            gamma = get_variable('gamma', [C], lambda t: t.fill_(1.0), device)
            beta = get_variable('beta', [C], lambda t: t.zero_(), device)
            moving_mean = get_variable('moving_mean', [C], lambda t: t.zero_(), device, trainable=False)
            moving_var = get_variable('moving_variance', [C], lambda t: t.fill_(1.0), device, trainable=False)
        _STORE.pending_l2["regularization_losses"][id(beta)] = (beta, 1.0)
        _STORE.pending_l2["regularization_losses"][id(gamma)] = (gamma, 1.0)
        return gamma, beta, moving_mean, moving_var


    def batch_normalization(data, is_training, name, reuse=None):
        """tf.layers.batch_normalization(data, momentum=0.99, training=is_training, ...) over the last
        axis, with L2 regularizers (scale 1.0) on beta and gamma (sph3gcn_util.py:328-332)."""
        return _batch_normalization(data, is_training, name, reuse, fused=None)


    def _batch_normalization(data, is_training, name, reuse, fused):
        momentum, eps = BN_MOMENTUM, BN_EPSILON
        C = data.shape[-1]
        gamma, beta, moving_mean, moving_var = _bn_variables(C, data.device, name, reuse)
        training = _as_bool(is_training)
        if fused is None:
            fused = FUSED_TAIL and data.is_cuda
        if fused:
            return layer_tail.bias_act_bn(data, None, gamma, beta, moving_mean, moving_var, act=layer_tail.ACT_NONE,
                                          training=training, eps=eps, momentum=momentum)
        flat = data.reshape(-1, C)
        if training:
            mean = flat.mean(dim=0)
            var = flat.var(dim=0, unbiased=False)
            with torch.no_grad():
                moving_mean.mul_(momentum).add_(mean.detach(), alpha=1 - momentum)
                moving_var.mul_(momentum).add_(var.detach(), alpha=1 - momentum)
        else:
            mean, var = moving_mean, moving_var
        out = (flat - mean) * torch.rsqrt(var + eps) * gamma + beta
        return out.reshape(data.shape)
