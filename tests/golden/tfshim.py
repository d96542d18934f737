"""A minimal, eager `tensorflow` stand-in (torch-backed, fp32 or fp64) that is just large
enough to EXECUTE the hot-path source of /root/reference/holE.py unmodified:

    corrupt_heads / corrupt_tails / corrupt_entities / corrupt_batch     holE.py:97-158
    get_embedding                                                          holE.py:161-168
    evaluate_triples (sigmoid and --log_loss branches), summarize          holE.py:179-202, 237-246
    evaluate_batch (hinge and --log_loss branches)                         holE.py:205-234
    tf.train.inverse_time_decay + GradientDescentOptimizer.minimize        holE.py:291-296

TEST INFRASTRUCTURE ONLY (build container only: it imports /root/reference).  TensorFlow 1.2,
which holE.py needs, is not installable here; with this shim the *reference's own code* -- not
a restatement of it -- produces tests/golden/tfshim_step.npz (make_tfshim_golden.py), and the
oracle is pinned against that fixture.  What the shim itself has to get right is the semantics
of the ~30 TF ops the hot path calls.  Everything that is not plain elementwise arithmetic is
stated ONCE here, with its TF 1.2 source of truth:

  * embedding_lookup(max_norm): clip_by_norm over every axis but the first of the gathered
    [B,1,D] tensor, `t * clip_norm * minimum(rsqrt(sum t*t), 1/clip_norm)`  (TF 1.2
    embedding_ops.py + clip_ops.py; SURVEY App. B: graph.pbtxt:3107-3595, axes const :3197);
    its gradient w.r.t. the variable is an IndexedSlices (ids, rows);
  * Minimum / Maximum gradients route ties to the FIRST argument (`x <= y` / `x >= y`:
    math_grad.py _MaximumMinimumGrad; graph.pbtxt:16739, :40692) -- torch splits ties;
  * gradients of a non-scalar loss are seeded with ones (gradients_impl.py; graph.pbtxt:16484-16550);
  * minimize(): the variable's IndexedSlices are concatenated in the order
    [r+, r-, t+, t-, h+, h-] and applied by ONE ScatterSub(var, ids, lr * rows), duplicates
    applied sequentially (graph.pbtxt:47849-48049); if a dense gradient exists as well (the
    --log_loss branch's l2_loss(embeddings)) everything is densified and summed, then
    var -= lr * grad (gradients_impl._AggregatedGrads + ApplyGradientDescent);
  * inverse_time_decay: lr / (1 + rate * (float32(step) / decay_steps))  (graph.pbtxt:16151-16412);
  * tf.sigmoid = 1 / (1 + exp(-x)).

Random ops do not draw: tf.random_uniform pops the next value from `feed_random(...)`, so the
generator can make the reference's own corrupt_* code reproduce given corruption ids.
"""
import contextlib
import importlib.util
import sys
import types

import numpy as np
import torch

REF = "/root/reference"

# ----------------------------------------------------------------------------------------
# dtypes / state
# ----------------------------------------------------------------------------------------
float32, float64, int32, int64, string = torch.float32, torch.float64, torch.int32, torch.int64, "string"

_STATE = {"float": torch.float32, "random": [], "trainable": []}


def set_float(dtype):
    """Floating type of the whole graph: np.float32 mirrors the TF graph, np.float64 is the
    high-precision yardstick (TF's ops are the same functions at either width)."""
    _STATE["float"] = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32


def feed_random(values):
    """Values returned, in order, by the next tf.random_uniform calls."""
    _STATE["random"] = list(values)


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, np.ndarray) and x.dtype.kind in "OUS":
        return x
    return torch.as_tensor(np.asarray(x), dtype=dtype)


# ----------------------------------------------------------------------------------------
# the few ops whose gradient convention differs from torch's
# ----------------------------------------------------------------------------------------
class _MinimumTF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x <= y)                      # math_grad._MinimumGrad: LessEqual
        ctx.shapes = (x.shape, y.shape)
        return torch.minimum(x, y)

    @staticmethod
    def backward(ctx, g):
        (xmask,) = ctx.saved_tensors
        gx, gy = torch.where(xmask, g, torch.zeros_like(g)), torch.where(xmask, torch.zeros_like(g), g)
        return gx.sum_to_size(ctx.shapes[0]), gy.sum_to_size(ctx.shapes[1])


class _MaximumTF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x >= y)                      # math_grad._MaximumGrad: GreaterEqual
        ctx.shapes = (x.shape, y.shape)
        return torch.maximum(x, y)

    @staticmethod
    def backward(ctx, g):
        (xmask,) = ctx.saved_tensors
        gx, gy = torch.where(xmask, g, torch.zeros_like(g)), torch.where(xmask, torch.zeros_like(g), g)
        return gx.sum_to_size(ctx.shapes[0]), gy.sum_to_size(ctx.shapes[1])


class Variable:
    """tf.Variable / tf.get_variable: holds a torch leaf; gathers record IndexedSlices."""

    def __init__(self, value, name="Variable"):
        self.name = name + ":0"
        self.t = torch.as_tensor(value).clone().requires_grad_(torch.as_tensor(value).is_floating_point())
        self.slices = []          # (ids int64[B], rows [B, D]) in gather-creation order
        if self.t.requires_grad:
            _STATE["trainable"].append(self)

    def numpy(self):
        return self.t.detach().numpy()


class _Gather(torch.autograd.Function):
    """params[ids]; the gradient w.r.t. params is recorded as an IndexedSlices on the variable
    instead of being densified (array_grad._GatherGrad)."""

    @staticmethod
    def forward(ctx, params, ids, var, slot):
        ctx.var, ctx.slot, ctx.ids = var, slot, ids
        return params[ids]

    @staticmethod
    def backward(ctx, g):
        ctx.var.slices[ctx.slot] = (ctx.ids.reshape(-1), g.reshape(-1, g.shape[-1]).clone())
        return None, None, None, None


# ----------------------------------------------------------------------------------------
# module surface used by holE.py
# ----------------------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name=None):
    if dtype == string or isinstance(value, str):
        return np.asarray(value, dtype=object)
    return _t(value, dtype)


def cast(x, dtype, name=None):
    return _t(x).to(dtype)


def slice(x, begin, size, name=None):  # noqa: A001  (tf.slice)
    x = _t(x)
    idx = tuple(builtins_slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return x[idx]


import builtins as _b  # noqa: E402
builtins_slice = _b.slice


def reshape(x, shape, name=None):
    if isinstance(x, np.ndarray):
        return x.reshape([int(s) for s in shape])
    return _t(x).reshape([int(s) for s in shape])


def shape(x, name=None):
    return torch.tensor(list(x.shape), dtype=torch.int32)


def range(start, limit=None, delta=1, dtype=None, name=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    return torch.arange(int(start), int(limit), int(delta), dtype=torch.int32)


def gather(params, indices, name=None):
    return _t(params)[_t(indices).long()]


def concat(values, axis, name=None):
    return torch.cat([_t(v) for v in values], dim=axis)


def stack(values, axis=0, name=None):
    return torch.stack([_t(v) for v in values], dim=axis)


def random_uniform(shape, minval=0, maxval=None, dtype=float32, seed=None, name=None):  # noqa: A002
    if not _STATE["random"]:
        raise RuntimeError("tfshim: tf.random_uniform called with nothing fed (feed_random)")
    v = _STATE["random"].pop(0)
    want = [int(s) for s in (shape.tolist() if isinstance(shape, torch.Tensor) else shape)]
    out = _t(v, dtype if dtype in (torch.int32, torch.int64) else _STATE["float"])
    assert list(out.shape) == want, (out.shape, want)
    return out


def less(x, y, name=None):
    return _t(x) < y


def cond(pred, true_fn=None, false_fn=None, fn1=None, fn2=None, name=None):
    t, f = (true_fn or fn1), (false_fn or fn2)
    return t() if bool(pred) else f()


def complex(real, imag, name=None):  # noqa: A001
    return torch.complex(real, imag)


def conj(x, name=None):
    return torch.conj(x).resolve_conj()


def real(x, name=None):
    return x.real if x.is_complex() else x


def multiply(x, y, name=None):
    return x * y


def scalar_mul(scalar, x):
    return scalar * x


def reduce_sum(x, axis=None, keep_dims=False, name=None, reduction_indices=None):
    axis = reduction_indices if axis is None else axis
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keep_dims)


def reduce_mean(x, axis=None, keep_dims=False, name=None):
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keep_dims)


def reduce_max(x, axis=None, keep_dims=False, name=None):
    return x.max() if axis is None else x.amax(dim=axis, keepdim=keep_dims)


def reduce_min(x, axis=None, keep_dims=False, name=None):
    return x.min() if axis is None else x.amin(dim=axis, keepdim=keep_dims)


def sigmoid(x, name=None):
    return 1.0 / (1.0 + torch.exp(-x))          # Eigen scalar_sigmoid_op


def exp(x, name=None):
    return torch.exp(x)


def log(x, name=None):
    return torch.log(x)


def sqrt(x, name=None):
    return torch.sqrt(x)


def square(x, name=None):
    return x * x


def rsqrt(x, name=None):
    return 1.0 / torch.sqrt(x)


def minimum(x, y, name=None):
    x = _t(x)
    return _MinimumTF.apply(x, _t(y, x.dtype))


def maximum(x, y, name=None):
    x = _t(x)
    return _MaximumTF.apply(x, _t(y, x.dtype).expand_as(x) if not isinstance(y, torch.Tensor) else y)


@contextlib.contextmanager
def name_scope(name, default_name=None, values=None):
    yield name


def clip_by_norm(t, clip_norm, axes=None, name=None):
    """TF 1.2 clip_ops.clip_by_norm."""
    l2norm_inv = rsqrt(reduce_sum(t * t, axes, keep_dims=True))
    one = torch.ones((), dtype=t.dtype)
    return t * clip_norm * minimum(l2norm_inv, one / clip_norm)


def _embedding_lookup(params, ids, partition_strategy="mod", name=None, validate_indices=True, max_norm=None):
    """TF 1.2 embedding_ops.embedding_lookup, single-shard case."""
    var = params
    ids = _t(ids).long()
    if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= var.t.shape[0]):
        raise IndexError("InvalidArgumentError: indices out of range in embedding_lookup")   # Gather op
    var.slices.append(None)
    x = _Gather.apply(var.t, ids, var, len(var.slices) - 1)
    if max_norm is not None:
        x = clip_by_norm(x, max_norm, axes=list(_b.range(1, x.dim())))
    return x


def _l2_loss(t, name=None):
    t = t.t if isinstance(t, Variable) else t
    return (t * t).sum() / 2


nn = types.SimpleNamespace(embedding_lookup=_embedding_lookup, l2_loss=_l2_loss)

# summaries are side effects on a FileWriter: no-ops here
summary = types.SimpleNamespace(scalar=lambda *a, **k: None, histogram=lambda *a, **k: None,
                                merge_all=lambda: None)


class _MutableHashTable:
    """tf.contrib.lookup.MutableHashTable: insert(keys, values) / lookup(keys) with a default."""

    def __init__(self, key_dtype, value_dtype, default_value, shared_name=None, name=None):
        self.key_dtype, self.value_dtype, self.default = key_dtype, value_dtype, default_value
        self.d = {}

    @staticmethod
    def _key(k):
        return k.item() if isinstance(k, (torch.Tensor, np.generic)) else k

    def insert(self, keys, values, name=None):
        keys = keys.reshape(-1) if hasattr(keys, "reshape") else keys
        for k, v in zip(keys, values):
            self.d[self._key(k)] = v

    def lookup(self, keys, name=None):
        flat = keys.reshape(-1)
        vals = [self.d.get(self._key(k), self.default) for k in flat]
        if self.value_dtype == string:
            return np.asarray(vals, dtype=object).reshape(tuple(keys.shape))
        out = torch.stack([_t(v, self.value_dtype) for v in vals])
        return out.reshape(tuple(keys.shape) + tuple(out.shape[1:]))


contrib = types.SimpleNamespace(
    lookup=types.SimpleNamespace(MutableHashTable=_MutableHashTable),
    tensorboard=types.SimpleNamespace(plugins=types.SimpleNamespace(projector=types.SimpleNamespace())))


def _inverse_time_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
    """TF 1.2 learning_rate_decay.inverse_time_decay (always float32 for a Python-float lr)."""
    f = np.float32
    step = global_step.t.item() if isinstance(global_step, Variable) else global_step
    p = f(step) / f(decay_steps)
    return f(learning_rate) / (f(1.0) + f(decay_rate) * p)


#: position of each gather in the ScatterSub concat order [r+, r-, t+, t-, h+, h-]; gathers are
#: created as h+, t+, r+ (evaluate_triples on the positives) then h-, t-, r- for every corrupt batch
def _concat_order(n_slices):
    groups = n_slices // 3
    order = []
    for col in (2, 1, 0):                 # r, t, h
        for g in _b.range(groups):        # +, then each corrupt batch
            order.append(3 * g + col)
    return order


class _GradientDescentOptimizer:
    def __init__(self, learning_rate, use_locking=False, name="GradientDescent"):
        self.lr = learning_rate

    def minimize(self, loss, global_step=None, var_list=None, name=None):
        """Eager: differentiates `loss` (seed = ones) w.r.t. the variables in var_list and
        applies the update at once.  Returns the applied (ids, lr * rows) for inspection."""
        (var,) = var_list if var_list is not None else _STATE["trainable"][-1:]
        if var.t.grad is not None:
            var.t.grad = None
        loss.backward(torch.ones_like(loss))
        lr = torch.as_tensor(float(self.lr) if not isinstance(self.lr, torch.Tensor) else self.lr,
                             dtype=var.t.dtype)
        slices = [var.slices[k] for k in _concat_order(len(var.slices))]
        ids = torch.cat([s[0] for s in slices])
        rows = torch.cat([s[1] for s in slices])
        with torch.no_grad():
            if var.t.grad is not None:         # a dense gradient too: densify and add (--log_loss + l2)
                dense = var.t.grad.clone()
                dense.index_add_(0, ids, rows)
                var.t -= lr * dense
            else:                              # ScatterSub(var, ids, lr * rows), sequential duplicates
                E = var.t.numpy()
                np.subtract.at(E, ids.numpy(), (rows * lr).numpy())
        var.slices = []
        var.t.grad = None
        if isinstance(global_step, Variable):
            with torch.no_grad():
                global_step.t += 1
        return ids, rows


train = types.SimpleNamespace(GradientDescentOptimizer=_GradientDescentOptimizer,
                              inverse_time_decay=_inverse_time_decay)


def install():
    """Register this module as `tensorflow` (and the contrib sub-modules holE.py imports)."""
    me = sys.modules[__name__]
    sys.modules["tensorflow"] = me
    for name in ("tensorflow.contrib", "tensorflow.contrib.tensorboard", "tensorflow.contrib.tensorboard.plugins",
                 "tensorflow.contrib.tensorboard.plugins.projector"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["tensorflow.contrib.tensorboard.plugins"].projector = \
        sys.modules["tensorflow.contrib.tensorboard.plugins.projector"]
    return me


def import_reference_hole():
    """/root/reference/holE.py executed against this shim."""
    install()
    spec = importlib.util.spec_from_file_location("ref_holE_tfshim", REF + "/holE.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
