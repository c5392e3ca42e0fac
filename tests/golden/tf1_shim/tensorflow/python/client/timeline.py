"""tensorflow.python.client.timeline of the TF-1 shim: imported by train/trainer.py, used only when GPU_LOG is set."""


class Timeline(object):
  def __init__(self, *a, **k):
    raise NotImplementedError("timeline: profiling is not part of the shim")
