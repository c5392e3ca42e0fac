"""RMSPropApplier: drop-in for train/rmsprop_applier.py, computed by K6 on the device.

Same constructor, `minimize_local`, `_apply_gradients`, `get_slot`.  TensorFlow graph handles
cannot exist here, so the documented deviation is: variables and gradients are torch CUDA
tensors and `_apply_gradients` performs the update eagerly, returning `(None, global_grad_norm)`
where the reference returned `(group_op, global_grad_norm)` (rmsprop_applier.py:129).

The variables are re-homed into one flat fp32 buffer the first time they are seen (their
`.data` become views of it) and the `rms` / `momentum` slots are flat buffers of the same
length, initialised to 1.0 / 0.0 like rmsprop_applier.py:38-43; the whole update is then two
launches (sum of squares, fused clip + RMSProp) instead of one per variable (:126-128).

`world_size > 1` (torch.distributed initialised): gradients are summed across ranks with
reduce-scatter, every rank updates the 1/world slice it owns (its slots only cover that
slice), and the updated parameters are all-gathered -- the synchronous stand-in for the
reference's lock-free shared update.
"""
import torch

from .. import _lib
from .. import kernels as K


class RMSPropApplier(object):
  def __init__(self, learning_rate, decay=0.9, momentum=0.0, epsilon=1e-10, clip_norm=40.0,
               device="/cpu:0", name="RMSPropApplier", process_group=None, average_gradients=True,
               keep_momentum_slot=True):
    self._name = name
    self._learning_rate = learning_rate
    self._decay = decay
    self._momentum = momentum
    self._epsilon = epsilon
    self._clip_norm = clip_norm
    self._device = device          # kept for signature compatibility; tensors decide the device
    self._slots = {}
    self._vars = None
    self._group = process_group
    self._average = average_gradients
    # TF's ApplyRMSProp writes the `momentum` slot even when momentum == 0 (mom = lr * g / sqrt(ms + eps), then
    # var -= mom; rmsprop_applier.py:86-93), and get_slot() can observe it: keep doing so by default (28 B per parameter
    # instead of 20 B; K6 is latency-bound at P = 1.9 M either way).  False: skip the dead store.
    self._keep_mom = bool(keep_momentum_slot)
    # the two device kernels (K6).  tests/test_sharded_applier.py swaps in a CPU stand-in to drive
    # the partition + collective plumbing under gloo; the product path is always `kernels`.
    self._ops = K

  # -- distributed helpers ---------------------------------------------------------------
  def _world(self):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
      return dist.get_world_size(self._group), dist.get_rank(self._group)
    return 1, 0

  # -- slots (rmsprop_applier.py:38-77) --------------------------------------------------
  def _create_slots(self, var_list):
    if self._vars is not None:
      if len(var_list) != len(self._vars) or any(a is not b for a, b in zip(var_list, self._vars)):
        raise _lib.UnrealError("RMSPropApplier is bound to one global variable list")
      return
    var_list = list(var_list)
    dev = var_list[0].device
    if dev.type != "cuda" and self._ops is K:
      raise _lib.UnrealError("RMSPropApplier needs CUDA variables; there is no CPU fallback")
    world, rank = self._world()
    sizes = [v.numel() for v in var_list]
    total = sum(sizes)
    quantum = 4 * world                     # float4 per rank
    padded = (total + quantum - 1) // quantum * quantum
    flat = torch.zeros(padded, dtype=torch.float32, device=dev)
    off = 0
    self._offsets = []
    for v, n in zip(var_list, sizes):
      if v.dtype != torch.float32:
        raise _lib.UnrealError("variables must be float32")
      flat[off:off + n].copy_(v.detach().reshape(-1))
      v.data = flat[off:off + n].view_as(v)   # variables now live in the flat buffer
      self._offsets.append((off, n))
      off += n
    self._vars = var_list
    self._total, self._padded = total, padded
    self._flat_var = flat
    self._flat_grad = torch.zeros(padded, dtype=torch.float32, device=dev)
    self._alloc_slots(dev, world, rank)

  def _alloc_slots(self, dev, world, rank):
    padded = self._padded
    self._shard = padded // world
    self._lo = rank * self._shard
    # slots only for the owned slice (the whole buffer when world == 1)
    self._rms = torch.ones(self._shard, dtype=torch.float32, device=dev)
    self._mom = torch.zeros(self._shard, dtype=torch.float32, device=dev)
    self._grad_shard = torch.zeros(self._shard, dtype=torch.float32, device=dev) if world > 1 else None
    self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
    self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
    self._slots = {"rms": self._rms, "momentum": self._mom}

  def get_slot(self, var, name):
    """View of the `rms` / `momentum` slot of `var` (rmsprop_applier.py:67-71).  With
    world_size > 1 only the part of the variable inside this rank's slice is returned."""
    if self._vars is None or name not in self._slots:
      return None
    for v, (off, n) in zip(self._vars, self._offsets):
      if v is var:
        lo, hi = max(off, self._lo), min(off + n, self._lo + self._shard)
        s = self._slots[name][lo - self._lo:hi - self._lo] if hi > lo else self._slots[name][:0]
        return s.view_as(var) if hi - lo == n else s
    return None

  def flat_parameters(self):
    return self._flat_var[:self._total]

  def bind_flat(self, flat_var):
    """Adopt a caller-owned flat fp32 parameter buffer (UnrealModel.flat) as the single global
    variable: no re-homing, the slots cover it element for element."""
    if self._vars is not None:
      if self._flat_var.data_ptr() != flat_var.data_ptr():
        raise _lib.UnrealError("RMSPropApplier is bound to one global variable list")
      return
    world, rank = self._world()
    if flat_var.dtype != torch.float32 or not (flat_var.is_cuda or self._ops is not K) or not flat_var.is_contiguous():
      raise _lib.UnrealError("bind_flat needs a contiguous float32 CUDA buffer; there is no CPU fallback")
    if flat_var.numel() % (4 * world) != 0:
      raise _lib.UnrealError("flat buffer length must be a multiple of 4 * world_size")
    self._vars = [flat_var]
    self._offsets = [(0, flat_var.numel())]
    self._total = self._padded = flat_var.numel()
    self._flat_var = flat_var.detach()
    self._flat_grad = torch.zeros_like(self._flat_var)
    self._alloc_slots(flat_var.device, world, rank)

  def apply_flat_to(self, flat_var, flat_grad, learning_rate=None):
    """Clip + RMSProp (+ the NCCL exchange when distributed) of `flat_grad` into `flat_var`."""
    self.bind_flat(flat_var)
    if flat_grad.data_ptr() != self._flat_grad.data_ptr():
      if flat_grad.numel() != self._padded or flat_grad.dtype != torch.float32 or not flat_grad.is_contiguous():
        raise _lib.UnrealError("flat_grad must be a contiguous float32 buffer of the parameters' length")
      self._flat_grad = flat_grad
    return self.apply_flat(learning_rate)

  # -- update -----------------------------------------------------------------------------
  def minimize_local(self, loss, global_var_list, local_var_list, thread_index, learning_rate=None):
    """rmsprop_applier.py:95-106: gradients of `loss` w.r.t. the local variables, applied to the
    global ones."""
    grads = torch.autograd.grad(loss, list(local_var_list), allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(v) for g, v in zip(grads, local_var_list)]
    return self._apply_gradients(global_var_list, grads, thread_index=thread_index, learning_rate=learning_rate)

  def _lr(self, learning_rate):
    lr = self._learning_rate if learning_rate is None else learning_rate
    if callable(lr):
      lr = lr()
    if isinstance(lr, torch.Tensor) and lr.is_cuda:
      return lr                             # device scalar: read by K6 when it runs (CUDA-graph replays)
    return float(lr)

  def _apply_gradients(self, global_var_list, local_grad_list, name=None, thread_index=None, learning_rate=None):
    """rmsprop_applier.py:109-132 -> (None, global_grad_norm as a 1-element CUDA tensor)."""
    self._create_slots(global_var_list)
    flat_g = self._flat_grad
    if isinstance(local_grad_list, torch.Tensor):          # already flat
      flat_g[:self._total].copy_(local_grad_list.reshape(-1))
    else:
      for g, (off, n) in zip(local_grad_list, self._offsets):
        flat_g[off:off + n].copy_(g.reshape(-1))
    return None, self.apply_flat(learning_rate)

  def flat_gradient(self):
    """The flat gradient buffer; producers may write into it directly and call apply_flat()."""
    return self._flat_grad

  def apply_flat(self, learning_rate=None):
    import torch.distributed as dist
    lr = self._lr(learning_rate)
    world, _ = self._world()
    self._sumsq.zero_()
    ops = self._ops
    if world == 1:
      ops.grad_sumsq(self._flat_grad, self._sumsq)
      ops.rmsprop_update(self._flat_var, self._rms, self._mom if (self._momentum != 0.0 or self._keep_mom) else None, self._flat_grad,
                       self._sumsq, lr, self._decay, self._momentum, self._epsilon, self._clip_norm,
                       grad_scale=1.0, grad_norm=self._norm)
      return self._norm
    dist.reduce_scatter_tensor(self._grad_shard, self._flat_grad, op=dist.ReduceOp.SUM, group=self._group)
    ops.grad_sumsq(self._grad_shard, self._sumsq)
    dist.all_reduce(self._sumsq, op=dist.ReduceOp.SUM, group=self._group)      # 8 bytes
    var_shard = self._flat_var[self._lo:self._lo + self._shard]
    ops.rmsprop_update(var_shard, self._rms, self._mom if (self._momentum != 0.0 or self._keep_mom) else None, self._grad_shard,
                     self._sumsq, lr, self._decay, self._momentum, self._epsilon, self._clip_norm,
                     grad_scale=(1.0 / world) if self._average else 1.0, grad_norm=self._norm)
    dist.all_gather_into_tensor(self._flat_var, var_shard, group=self._group)
    return self._norm
