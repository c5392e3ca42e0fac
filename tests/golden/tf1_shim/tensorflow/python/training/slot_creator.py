"""tensorflow.python.training.slot_creator of the TF-1 shim: a slot is a non-trainable variable shaped like its primary."""
import tensorflow as tf


def create_slot(primary, val, name, colocate_with_primary=True):
  sess = tf.Session()
  v = tf.Variable(primary.op.name + "/" + name, tf._t(sess.run(val)).to(tf.DT).reshape(primary.value.shape).clone())
  return v


def create_zeros_slot(primary, name, dtype=None, colocate_with_primary=True):
  import torch
  return tf.Variable(primary.op.name + "/" + name, torch.zeros_like(primary.value))
