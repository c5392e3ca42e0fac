"""tensorflow.python.training.training_ops of the TF-1 shim: the dense ApplyRMSProp kernel, from its documented definition
(tensorflow/core/ops/training_ops.cc):  ms <- rho * ms + (1 - rho) * grad^2;  mom <- momentum * mom + lr * grad / sqrt(ms +
epsilon);  var <- var - mom.  The three variables are updated in place when the returned op is run."""
import torch

import tensorflow as tf


def apply_rms_prop(var, ms, mom, lr, rho, momentum, epsilon, grad, use_locking=False, name=None):
  def fn(lr_, rho_, mo_, eps_, g):
    lr_, rho_, mo_, eps_, g = (tf._t(x).detach() for x in (lr_, rho_, mo_, eps_, g))
    ms.value = rho_ * ms.value + (1.0 - rho_) * g * g
    mom.value = mo_ * mom.value + lr_ * g / torch.sqrt(ms.value + eps_)
    var.value = (var.value.detach() - mom.value)
    return var.value
  node = tf.Node(fn, [lr, rho, momentum, epsilon, grad])
  op = tf._Op("ApplyRMSProp")
  op.op_node = node
  node.op = op
  return node
