"""A minimal TensorFlow-1 graph-mode API over torch float64 -- TEST INFRASTRUCTURE for tests/golden/make_model_golden.py.

TensorFlow (the reference's un-vendored, un-pinned dependency: README says r1.0; model.py needs >= 1.2 for
``contrib.rnn.BasicLSTMCell``'s ``kernel`` / ``bias`` variables) is not installable in this container.  This package lets the
REFERENCE'S OWN ``model/model.py`` be imported and executed unmodified: its graph-building code (placeholders, variable
scopes and reuse, layer wiring, the loss formulas, the ``run_*`` methods and their feed dicts) runs as written; only the ~30
TensorFlow ops it calls on the vanilla path (segnet_mode == 0) are restated here from their published TF-1 definitions:

  tf.nn.conv2d            NHWC input, HWIO filter, strides [1,s,s,1], padding VALID
  tf.nn.conv2d_transpose  filter [kh, kw, out_channels, in_channels]; the gradient of conv2d w.r.t. its input, VALID
  contrib.rnn.BasicLSTMCell(n, forget_bias=1.0, state_is_tuple=True) under tf.nn.dynamic_rnn(time_major=False):
                          variables <scope>/basic_lstm_cell/kernel [in+n, 4n] (glorot-uniform) and bias [4n] (zeros);
                          i, j, f, o = split(concat([x, h], 1) @ kernel + bias, 4, axis=1);
                          c' = c * sigmoid(f + forget_bias) + sigmoid(i) * tanh(j);  h' = tanh(c') * sigmoid(o)
  tf.nn.l2_loss(t) = sum(t ** 2) / 2;  tf.nn.softmax over the last axis;  reduce_* with reduction_indices / keep_dims
  tf.variable_scope / get_variable: names joined by '/', reuse returns the existing variable, a scope OBJECT re-enters
                          its absolute scope (how dynamic_rnn(scope=scope) places the cell's variables)

Evaluation is lazy like a TF graph: every op returns a Node; Session.run(fetches, feed_dict) evaluates the needed sub-graph in
float64 with torch, so gradients of any scalar node w.r.t. the variables are available through autograd
(``gradients``).  Nothing under unreal_b200/ imports this package.
"""
import collections
import contextlib

import numpy as np
import torch

float32, float64, int32, int64 = "float32", "float64", "int32", "int64"
bool = "bool"
DT = torch.float64
_rng = np.random.RandomState(1234)


def _seed(s):
  global _rng
  _rng = np.random.RandomState(s)


class Dimension(object):
  def __init__(self, v):
    self.value = v


class TensorShape(object):
  def __init__(self, dims):
    self.dims = None if dims is None else list(dims)

  def as_list(self):
    if self.dims is None:
      raise ValueError("unknown static shape")
    return list(self.dims)

  def __getitem__(self, i):
    r = self.as_list()[i]
    return TensorShape(r) if isinstance(i, slice) else Dimension(r)

  def __len__(self):
    return len(self.as_list())


def _val(x, ev):
  """Evaluate an op argument: Nodes through `ev`, containers recursively, everything else as a constant."""
  if isinstance(x, Node):
    return ev(x)
  if isinstance(x, tuple) and hasattr(x, "_fields"):
    return type(x)(*[_val(v, ev) for v in x])
  if isinstance(x, (list, tuple)):
    return type(x)(_val(v, ev) for v in x)
  return x


def _t(x):
  if isinstance(x, torch.Tensor):
    return x
  a = np.asarray(x)
  return torch.as_tensor(a, dtype=DT if a.dtype.kind == "f" else None)


class Node(object):
  def __init__(self, fn, inputs, shape=None, name=None):
    self.fn, self.inputs, self._shape, self.name = fn, list(inputs), shape, name

  def get_shape(self):
    return TensorShape(self._shape)

  @property
  def shape(self):
    return TensorShape(self._shape)

  def __add__(self, o): return _binary(torch.add, self, o)
  def __radd__(self, o): return _binary(torch.add, o, self)
  def __sub__(self, o): return _binary(torch.sub, self, o)
  def __rsub__(self, o): return _binary(torch.sub, o, self)
  def __mul__(self, o): return _binary(torch.mul, self, o)
  def __rmul__(self, o): return _binary(torch.mul, o, self)
  def __truediv__(self, o): return _binary(torch.div, self, o)
  def __neg__(self): return Node(lambda a: -a, [self], self._shape)
  def __getitem__(self, i): return Node(lambda a: a[i], [self])
  __hash__ = object.__hash__


def _binary(f, a, b):
  sh = a._shape if isinstance(a, Node) and a._shape is not None else (b._shape if isinstance(b, Node) else None)
  return Node(lambda x, y: f(_t(x), _t(y)), [a, b], sh)


class Placeholder(Node):
  def __init__(self, dtype, shape, name):
    Node.__init__(self, None, [], None if shape is None else list(shape), name)
    self.dtype = dtype


class _Op(object):
  def __init__(self, name):
    self.name = name


class Variable(Node):
  def __init__(self, name, value):
    Node.__init__(self, None, [], list(value.shape), name + ":0")
    self.value = value
    self.op = _Op(name)
    self.dtype = float64
    self.device = "/cpu:0"

  def _ref(self):
    return self


def placeholder(dtype, shape=None, name=None):
  return Placeholder(dtype, shape, name)


def placeholder_with_default(default, shape, name=None):
  p = Placeholder(None, shape, name)
  p.default = default
  return p


def constant(v, dtype=None, shape=None, name=None):
  t = _t(np.asarray(v, dtype=np.float64 if isinstance(v, (float, list)) else None))
  if shape is not None:
    t = t.expand([int(d) for d in (shape.as_list() if isinstance(shape, TensorShape) else shape)]).clone()
  return Node(lambda: t, [], list(t.shape))


# ---- scopes, variables, collections ---------------------------------------------------------------
class GraphKeys(object):
  TRAINABLE_VARIABLES, GLOBAL_VARIABLES, LOCAL_VARIABLES = "trainable_variables", "variables", "local_variables"
  UPDATE_OPS = "update_ops"


class VariableScope(object):
  def __init__(self, name, reuse):
    self.name, self.reuse = name, reuse


class _Graph(object):
  def __init__(self):
    self.scopes = [VariableScope("", False)]
    self.variables = collections.OrderedDict()


_g = _Graph()


def reset_default_graph():
  global _g
  _g = _Graph()


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, reuse=None, **_):
  cur = _g.scopes[-1]
  if isinstance(name_or_scope, VariableScope):          # re-enter the scope object's absolute scope
    new = VariableScope(name_or_scope.name, name_or_scope.reuse or cur.reuse or bool_(reuse))
  else:
    name = name_or_scope if name_or_scope is not None else default_name
    new = VariableScope((cur.name + "/" if cur.name else "") + name, cur.reuse or bool_(reuse))
  _g.scopes.append(new)
  try:
    yield new
  finally:
    _g.scopes.pop()


def bool_(x):
  return x is True


@contextlib.contextmanager
def name_scope(name, default_name=None, values=None):
  yield (name or default_name or "") + "/"


@contextlib.contextmanager
def device(_):
  yield


def get_variable(name, shape=None, dtype=None, initializer=None, **_):
  sc = _g.scopes[-1]
  full = (sc.name + "/" if sc.name else "") + name
  if full in _g.variables:
    if not sc.reuse:
      raise ValueError("Variable %s already exists, disallowed (reuse not set)" % full)
    return _g.variables[full]
  if sc.reuse:
    raise ValueError("Variable %s does not exist (reuse set)" % full)
  shape = [int(s) for s in shape]
  if initializer is None:                                  # TF's default: glorot_uniform_initializer
    initializer = glorot_uniform_initializer()
  v = initializer(shape)
  v = torch.as_tensor(np.asarray(v, dtype=np.float64).reshape(shape), dtype=DT)
  var = Variable(full, v)
  _g.variables[full] = var
  return var


def get_collection(key, scope=None):
  if key in (GraphKeys.LOCAL_VARIABLES, GraphKeys.UPDATE_OPS):
    return []
  vs = list(_g.variables.values())
  return [v for v in vs if scope is None or v.name.startswith(scope)]


def random_uniform(shape, minval=0.0, maxval=1.0, dtype=None, seed=None, name=None):
  return _rng.uniform(minval, maxval, size=[int(s) for s in shape])      # eager: only ever used inside initialisers


def glorot_uniform_initializer():
  def init(shape, dtype=None, partition_info=None):
    fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (int(np.prod(shape[:-1])), shape[-1])
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return _rng.uniform(-lim, lim, size=shape)
  return init


def zeros_initializer():
  return lambda shape, dtype=None, partition_info=None: np.zeros(shape)


def variables_initializer(var_list, name=None):
  return Node(lambda: None, [])


def global_variables_initializer():
  return Node(lambda: None, [])


def assign(ref, value):
  def fn(v):
    ref.value = _t(v).detach().clone()
    return ref.value
  return Node(fn, [value])


def group(*ops, **_):
  return Node(lambda *a: None, [o.op_node if isinstance(o, _Op) else o for o in ops])


# ---- ops -------------------------------------------------------------------------------------------
def _axis(kw):
  for k in ("reduction_indices", "axis"):
    if kw.get(k) is not None:
      return kw[k]
  return None


def _keep(kw):
  return bool_(kw.get("keep_dims")) or bool_(kw.get("keepdims"))


def reduce_sum(x, axis=None, **kw):
  ax = axis if axis is not None else _axis(kw)
  return Node(lambda a: a.sum() if ax is None else a.sum(dim=ax, keepdim=_keep(kw)), [x])


def reduce_mean(x, axis=None, **kw):
  ax = axis if axis is not None else _axis(kw)
  return Node(lambda a: a.mean() if ax is None else a.mean(dim=ax, keepdim=_keep(kw)), [x])


def reduce_max(x, axis=None, **kw):
  ax = axis if axis is not None else _axis(kw)
  return Node(lambda a: a.max() if ax is None else a.max(dim=ax, keepdim=_keep(kw)).values, [x])


def _reshape_static(shape):
  try:
    return [None if int(s) == -1 else int(s) for s in shape]
  except TypeError:
    return None


def reshape(x, shape, name=None):
  return Node(lambda a, s: _t(a).reshape([int(v) for v in s]), [x, shape], _reshape_static(shape))


def matmul(a, b, name=None):
  sh = None
  if isinstance(a, Node) and isinstance(b, Node) and a._shape is not None and b._shape is not None:
    sh = [a._shape[0], b._shape[1]]
  return Node(lambda x, y: _t(x) @ _t(y), [a, b], sh)


def concat(values, axis, name=None):
  return Node(lambda vs: torch.cat([_t(v) for v in vs], dim=axis), [list(values)])


def stack(values, axis=0, name=None):
  return Node(lambda vs: [int(v) for v in vs], [list(values)])         # only used to assemble an output_shape


def shape(x, name=None):
  return Node(lambda a: list(a.shape), [x])


def multiply(a, b, name=None): return _binary(torch.mul, a, b)
def add(a, b, name=None): return _binary(torch.add, a, b)
def div(a, b, name=None): return _binary(torch.div, a, b)
def log(x, name=None): return Node(lambda a: torch.log(a), [x], getattr(x, "_shape", None))


def clip_by_value(x, lo, hi, name=None):
  return Node(lambda a: torch.clamp(a, lo, hi), [x], getattr(x, "_shape", None))


def argmax(x, axis=None, **kw):
  return Node(lambda a: a.argmax(dim=axis), [x])


def to_int32(x, name=None):
  return Node(lambda a: a.to(torch.int32), [x])


def gather(params, indices, **_):
  return Node(lambda p, i: p[i.long()], [params, indices])


def gradients(ys, xs, **_):
  """d ys / d xs for Variables xs, evaluated with autograd inside Session.run (a node like any other)."""
  return [_Grad(ys, x) for x in xs]


def convert_to_tensor(v, name=None, **_):
  return v if isinstance(v, Node) else constant(v)


@contextlib.contextmanager
def control_dependencies(_):
  yield


def global_norm(t_list):
  return Node(lambda ts: torch.sqrt(sum((_t(t) ** 2).sum() for t in ts)), [list(t_list)], [])


def clip_by_global_norm(t_list, clip_norm, use_norm=None, name=None):
  """t_list[i] * clip_norm / max(global_norm, clip_norm), global_norm = sqrt(sum_i l2norm(t_i)^2)  (TF's definition)."""
  norm = global_norm(t_list)
  return [Node(lambda t, n: _t(t) * clip_norm / torch.clamp(n, min=clip_norm), [t, norm], getattr(t, "_shape", None))
          for t in t_list], norm


class _Grad(Node):
  def __init__(self, y, x):
    Node.__init__(self, None, [], x._shape)
    self.y, self.x = y, x


class _NN(object):
  @staticmethod
  def relu(x, name=None):
    return Node(lambda a: torch.relu(a), [x], getattr(x, "_shape", None))

  @staticmethod
  def softmax(x, name=None):
    return Node(lambda a: torch.softmax(a, dim=-1), [x], getattr(x, "_shape", None))

  @staticmethod
  def l2_loss(x, name=None):
    return Node(lambda a: (a * a).sum() / 2.0, [x], [])

  @staticmethod
  def conv2d(x, W, strides, padding, name=None):
    assert padding == "VALID" and strides[0] == 1 and strides[3] == 1 and strides[1] == strides[2]
    s = strides[1]
    sh = None
    if x._shape is not None and W._shape is not None:
      n, h, w, _ = x._shape
      kh, kw, _, o = W._shape
      sh = [n, (h - kh) // s + 1, (w - kw) // s + 1, o]
    def fn(a, f):      # NHWC / HWIO -> torch's NCHW / OIHW
      return torch.nn.functional.conv2d(a.permute(0, 3, 1, 2), f.permute(3, 2, 0, 1), stride=s).permute(0, 2, 3, 1)
    return Node(fn, [x, W], sh)

  @staticmethod
  def conv2d_transpose(x, W, output_shape, strides, padding="SAME", name=None):
    assert padding == "VALID" and strides[0] == 1 and strides[3] == 1 and strides[1] == strides[2]
    s = strides[1]
    def fn(a, f, osh):  # filter [kh, kw, out, in] -> torch conv_transpose2d's [in, out, kh, kw]
      y = torch.nn.functional.conv_transpose2d(a.permute(0, 3, 1, 2), f.permute(3, 2, 0, 1), stride=s).permute(0, 2, 3, 1)
      assert list(y.shape) == [int(v) for v in osh], (y.shape, osh)
      return y
    return Node(fn, [x, W, output_shape])

  @staticmethod
  def dynamic_rnn(cell, inputs, initial_state=None, sequence_length=None, time_major=False, scope=None, dtype=None):
    assert not time_major
    with variable_scope(scope if scope is not None else "rnn"):
      in_dim = inputs._shape[-1]
      cell.build(in_dim)
    kernel, bias, n, fb = cell.kernel, cell.bias, cell.num_units, cell.forget_bias
    def fn(x, state, k, b, seq_len):
      c, h = _t(state[0]), _t(state[1])
      assert x.shape[0] == c.shape[0] and (seq_len is None or int(seq_len[0]) == x.shape[1])
      outs = []
      for t in range(x.shape[1]):
        z = torch.cat([x[:, t], h], dim=1) @ k + b
        i, j, f, o = z.split(n, dim=1)
        c = c * torch.sigmoid(f + fb) + torch.sigmoid(i) * torch.tanh(j)
        h = torch.tanh(c) * torch.sigmoid(o)
        outs.append(h)
      return torch.stack(outs, dim=1), c, h
    core = Node(fn, [inputs, initial_state, kernel, bias, sequence_length])
    outputs = Node(lambda r: r[0], [core], [inputs._shape[0], inputs._shape[1], n])
    return outputs, contrib.rnn.LSTMStateTuple(Node(lambda r: r[1], [core], [None, n]), Node(lambda r: r[2], [core], [None, n]))


nn = _NN()


class _Layers(object):
  @staticmethod
  def dropout(inputs, rate=0.5, training=False, **_):
    raise NotImplementedError("tf.layers.dropout: only on the segnet path")


layers = _Layers()


class _Losses(object):
  pass


losses = _Losses()
metrics = _Losses()


class test(object):
  TestCase = object


class RunOptions(object):
  FULL_TRACE = 3

  def __init__(self, *a, **k):
    pass


class RunMetadata(object):
  step_stats = None


# ---- session ---------------------------------------------------------------------------------------
class Session(object):
  def __init__(self, *a, **k):
    pass

  def __enter__(self):
    return self

  def __exit__(self, *a):
    return False

  def run(self, fetches, feed_dict=None, options=None, run_metadata=None):
    feed = {}
    for k, v in (feed_dict or {}).items():
      if isinstance(k, Node):
        feed[k] = v
      elif isinstance(k, tuple) and all(isinstance(x, Node) for x in k):      # a nested key (LSTMStateTuple of placeholders)
        for kk, vv in zip(k, v):
          feed[kk] = vv
    cache = {}
    want_grad = []
    seen = set()
    def collect(f):                      # every gradient node the fetches depend on, anywhere in the graph
      if isinstance(f, Node):
        if id(f) in seen:
          return
        seen.add(id(f))
        if isinstance(f, _Grad):
          want_grad.append(f)
          collect(f.y)
        for i in f.inputs:
          collect(i)
      elif isinstance(f, (list, tuple)):
        for v in f:
          collect(v)
    collect(fetches)
    gvars = []
    for g in want_grad:
      if g.x not in gvars:
        gvars.append(g.x)
    for v in gvars:
      v.value.requires_grad_(True)

    def ev(n):
      if id(n) in cache:
        return cache[id(n)]
      if isinstance(n, Placeholder):
        if n in feed:
          r = _t(feed[n])
          if n._shape is not None:
            assert len(r.shape) == len(n._shape) and all(s is None or s == d for s, d in zip(n._shape, r.shape)), \
                "feed for %s has shape %s, placeholder is %s" % (n.name, tuple(r.shape), n._shape)
        elif hasattr(n, "default"):
          r = _t(n.default)
        else:
          raise ValueError("placeholder %s was not fed" % n.name)
      elif isinstance(n, Variable):
        r = n.value
      elif isinstance(n, _Grad):
        if ("grad", id(n.y), id(n.x)) not in cache:        # one backward pass per differentiated node, for all variables
          y = ev(n.y)
          grads = torch.autograd.grad(y, [v.value for v in gvars], retain_graph=True, allow_unused=True)
          for v, gr in zip(gvars, grads):
            cache[("grad", id(n.y), id(v))] = torch.zeros_like(v.value) if gr is None else gr.detach()
        r = cache[("grad", id(n.y), id(n.x))]
      else:
        r = n.fn(*[_val(i, ev) for i in n.inputs])
      cache[id(n)] = r
      return r

    def out(f):
      if isinstance(f, Node):
        r = ev(f)
        if isinstance(r, torch.Tensor):
          return r.detach().numpy().copy()
        return r
      if isinstance(f, tuple) and hasattr(f, "_fields"):
        return type(f)(*[out(v) for v in f])
      if isinstance(f, (list, tuple)):
        return type(f)(out(v) for v in f) if isinstance(f, tuple) else [out(v) for v in f]
      return f
    with torch.enable_grad() if gvars else torch.no_grad():
      res = out(fetches)
    for v in gvars:
      v.value.requires_grad_(False)
    return res


from . import contrib  # noqa: E402  (tf.contrib.rnn.LSTMStateTuple)
