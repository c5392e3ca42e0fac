"""tensorflow.contrib of the TF-1 shim (see ../__init__.py): BasicLSTMCell / LSTMStateTuple are real, the slim-style layer
constructors model.py binds at import time exist only as names (they are called on the segnet path alone)."""
import collections


def _segnet_only(name):
  def f(*a, **k):
    raise NotImplementedError("contrib.%s: only used by the segnet path (segnet_mode != 0)" % name)
  f.__name__ = name
  return f


class _NS(object):
  pass


layers = _NS()
for _n in ("fully_connected", "conv2d", "conv2d_transpose", "max_pool2d", "batch_norm", "repeat", "l2_regularizer"):
  setattr(layers, _n, _segnet_only("layers." + _n))
layers.xavier_initializer = lambda *a, **k: None

framework = _NS()
framework.arg_scope = _segnet_only("framework.arg_scope")

rnn = _NS()
rnn.LSTMStateTuple = collections.namedtuple("LSTMStateTuple", ("c", "h"))


class BasicLSTMCell(object):
  """contrib.rnn.BasicLSTMCell(num_units, forget_bias=1.0, state_is_tuple=True): the cell's two variables are created on first
  use under <enclosing scope>/basic_lstm_cell/ (TF >= 1.2 names: kernel, bias) and shared by every later use."""

  def __init__(self, num_units, forget_bias=1.0, state_is_tuple=True, **_):
    assert state_is_tuple
    self.num_units, self.forget_bias = num_units, forget_bias
    self.kernel = self.bias = None

  def build(self, in_dim):
    import tensorflow as tf
    with tf.variable_scope("basic_lstm_cell"):
      self.kernel = tf.get_variable("kernel", [in_dim + self.num_units, 4 * self.num_units])
      self.bias = tf.get_variable("bias", [4 * self.num_units], initializer=tf.zeros_initializer())

  def zero_state(self, batch_size, dtype):
    import tensorflow as tf
    z = tf.constant([[0.0] * self.num_units] * batch_size)
    return rnn.LSTMStateTuple(z, z)


rnn.BasicLSTMCell = BasicLSTMCell
