"""PyTorch-CPU fp32 restatement of UnrealModel's vanilla path (model/model.py, segnet_mode == 0).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): imported only by ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py``.

**Parity: pinned to the reference's own model.py, executed; TensorFlow's kernels themselves unpinned.**  The arithmetic of
model.py lives in TensorFlow 1.x (un-vendored, unpinned; README says r1.0, ``BasicLSTMCell``'s ``kernel`` / ``bias`` names need
>= 1.2), which is neither under ``/root/reference`` nor installable here.  ``tests/golden/make_model_golden.py`` therefore
imports the reference's ``model/model.py`` UNMODIFIED over ``tests/golden/tf1_shim`` -- the ~30 TF-1 ops of the vanilla path
restated from their published definitions on torch float64 -- and records what the reference's own graph-building code, loss
formulas and ``run_*`` methods compute: ``tests/test_model_oracle.py::test_oracle_matches_the_references_model_py`` holds this
restatement to those vectors at 1e-9 (variable creation order and names, three acting steps with the carried LSTM state,
every ``run_*`` output, every loss term of a Trainer-shaped feed, the gradient of ``total_loss`` in all 20 variables).
What that does NOT pin is TensorFlow's implementation of the ops (conv2d, conv2d_transpose, BasicLSTMCell): the shim and this
file restate the same published semantics.  Those are anchored separately (same test file): every op against a direct loop
implementation of TF's documented definition; the LSTM unroll against ``torch.nn.LSTMCell`` fed the same weights with
permuted gate columns; autograd against float64 central differences of the total loss in all 20 variables; every head's loss
against numbers worked out by hand from model.py's formulas; and ``model/model_test.py``'s variable counts (20 / 18 / 12 /
14).  Everything below restates model.py line by line with the documented TF-1 semantics of the ops it calls:

  tf.nn.conv2d NHWC / HWIO / VALID                      model.py:283-289, :786-787
  tf.matmul + relu, flatten in NHWC order               model.py:332-340
  contrib.rnn.BasicLSTMCell(256, state_is_tuple=True)   model.py:110, :346-351
      one kernel [(in)+256, 1024] on concat([x, h]), gates i, j, f, o, forget_bias = 1.0,
      c' = c*sigmoid(f+1) + sigmoid(i)*tanh(j), h' = tanh(c')*sigmoid(o), state = (c, h)
  tf.nn.conv2d_transpose, filter [kh, kw, out, in]      model.py:418-430, :803-820
  dueling Q, max_a                                      model.py:431-441
  losses                                                model.py:490-598
  U(-1/sqrt(fan_in), 1/sqrt(fan_in)) init for W and b   model.py:31-42, :752-783

Batched form.  In the reference the leading axis of every training placeholder is TIME at batch
size 1 (one worker thread, one env).  Here tensors are time-major [T, N, ...]: N independent envs,
each of which is exactly one reference worker's unroll with its own LSTM state; ``mask[t, n]``
marks the steps that exist in env n's (possibly shorter) batch.  The total loss is the SUM over
envs of the reference's per-worker total loss.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


def variable_specs(action_size, objective_size=0, use_pixel_change=True, use_value_replay=True,
                   use_reward_prediction=True):
  """[(name, shape, fan_in)] in the reference's creation order (model.py:107-135; the LSTM kernel
  is created by dynamic_rnn after fc1).  VR adds no variables (it reuses the base weights)."""
  A, G = action_size, objective_size
  lstm_in = 256 + A + 1 + G
  specs = [
      ("W_base_conv1", (8, 8, 3, 16), 8 * 8 * 3), ("b_base_conv1", (16,), 8 * 8 * 3),
      ("W_base_conv2", (4, 4, 16, 32), 4 * 4 * 16), ("b_base_conv2", (32,), 4 * 4 * 16),
      ("W_base_fc1", (2592, 256), 2592), ("b_base_fc1", (256,), 2592),
      ("lstm_kernel", (lstm_in + 256, 1024), None), ("lstm_bias", (1024,), None),
      ("W_base_fc_p", (256, A), 256), ("b_base_fc_p", (A,), 256),
      ("W_base_fc_v", (256, 1), 256), ("b_base_fc_v", (1,), 256),
  ]
  if use_pixel_change:
    specs += [
        ("W_pc_fc1", (256, 2592), 256), ("b_pc_fc1", (2592,), 256),
        ("W_pc_deconv_v", (4, 4, 1, 32), 4 * 4 * 32), ("b_pc_deconv_v", (1,), 4 * 4 * 32),
        ("W_pc_deconv_a", (4, 4, A, 32), 4 * 4 * 32), ("b_pc_deconv_a", (A,), 4 * 4 * 32),
    ]
  if use_reward_prediction:
    specs += [("W_rp_fc1", (7776, 3), 7776), ("b_rp_fc1", (3,), 7776)]
  return specs


def init_params(action_size, objective_size=0, seed=0, **heads):
  """Reference initialisers (model.py:31-42): U(+-1/sqrt(fan_in)) for weights AND biases.
  BasicLSTMCell's kernel uses TF's default glorot-uniform and a zero bias."""
  rs = np.random.RandomState(seed)
  params = OrderedDict()
  for name, shape, fan_in in variable_specs(action_size, objective_size, **heads):
    if name == "lstm_kernel":
      lim = np.sqrt(6.0 / (shape[0] + shape[1]))
      v = rs.uniform(-lim, lim, size=shape)
    elif name == "lstm_bias":
      v = np.zeros(shape)
    else:
      d = 1.0 / np.sqrt(fan_in)
      v = rs.uniform(-d, d, size=shape)
    params[name] = torch.tensor(v, dtype=torch.float32)
  return params


def _bf16(x, on):
  return x.to(torch.bfloat16).to(torch.float32) if on else x


class ModelOracle(object):
  """Functional model over a dict of TF-layout parameters.  ``emulate_bf16`` rounds exactly the
  tensors the CUDA path stores in bf16 (GEMM operands: weights and the activations fed to them),
  so that the tight-tolerance tests compare accumulation order only."""

  def __init__(self, params, action_size, objective_size=0, pixel_change_lambda=0.05, entropy_beta=0.001,
               emulate_bf16=False, round_gates=True):
    self.p = params
    self.A = action_size
    self.G = objective_size
    self.pc_lambda = pixel_change_lambda
    self.entropy_beta = entropy_beta
    self.q = emulate_bf16
    # the LSTM gate pre-activations as bf16 too: the CUDA path's GEMM + cell kernel steps store them (small batches, the
    # acting step); its one-launch steps (UnrealModel.fused_lstm_step, large batches) take them from the fp32 accumulator
    self.qz = emulate_bf16 and round_gates
    # float32 like the reference's placeholders (model.py:141 "float"); float64 parameters switch every tensor to
    # double, which is what the finite-difference gradient check of tests/test_model_oracle.py runs in
    self.dtype = next(iter(params.values())).dtype

  def w(self, name):
    t = self.p[name]
    return _bf16(t, self.q) if name.startswith("W_") or name == "lstm_kernel" else t

  # model.py:281-289
  def encoder(self, images):
    """images [S,84,84,3] -> [S,9,9,32] (NHWC)."""
    x = _bf16(images.to(self.dtype), self.q).permute(0, 3, 1, 2)
    h1 = F.relu(F.conv2d(x, self.w("W_base_conv1").permute(3, 2, 0, 1), self.p["b_base_conv1"], stride=4))
    h1 = _bf16(h1, self.q)
    h2 = F.relu(F.conv2d(h1, self.w("W_base_conv2").permute(3, 2, 0, 1), self.p["b_base_conv2"], stride=2))
    return _bf16(h2.permute(0, 2, 3, 1), self.q)

  # model.py:321-355
  def lstm_layer(self, conv_out, lar, c0, h0):
    """conv_out [T,N,9,9,32], lar [T,N,A+1+G], (c0,h0) [N,256] -> outputs [T,N,256], (c,h)."""
    T, N = conv_out.shape[:2]
    flat = conv_out.reshape(T * N, 2592)
    fc = F.relu(flat @ self.w("W_base_fc1") + self.p["b_base_fc1"]).reshape(T, N, 256)
    x = _bf16(torch.cat([fc, lar.to(self.dtype)], dim=2), self.q)
    kernel, bias = self.w("lstm_kernel"), self.p["lstm_bias"]
    c, h = c0, h0
    outs = []
    for t in range(T):
      z = _bf16(torch.cat([x[t], _bf16(h, self.q)], dim=1) @ kernel + bias, self.qz)
      i, j, f, o = z.split(256, dim=1)
      c = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
      h = torch.tanh(c) * torch.sigmoid(o)
      outs.append(h)
    return torch.stack(outs), (c, h)

  # model.py:358-377
  def policy_value(self, lstm_out):
    pi = F.softmax(lstm_out @ self.p["W_base_fc_p"] + self.p["b_base_fc_p"], dim=-1)
    v = (lstm_out @ self.p["W_base_fc_v"] + self.p["b_base_fc_v"]).squeeze(-1)
    return pi, v

  def base_forward(self, images, lar, c0, h0):
    T, N = images.shape[:2]
    enc = self.encoder(images.reshape(T * N, 84, 84, 3)).reshape(T, N, 9, 9, 32)
    out, state = self.lstm_layer(enc, lar, c0, h0)
    pi, v = self.policy_value(out)
    return pi, v, state

  # model.py:411-443
  def pc_deconv(self, lstm_out):
    """[S,256] -> pc_q [S,20,20,A], pc_q_max [S,20,20]."""
    h = F.relu(_bf16(lstm_out, self.q) @ self.w("W_pc_fc1") + self.p["b_pc_fc1"])
    h = _bf16(h, self.q).reshape(-1, 9, 9, 32).permute(0, 3, 1, 2)
    wv = self.w("W_pc_deconv_v").permute(3, 2, 0, 1)   # [kh,kw,out,in] -> [in,out,kh,kw]
    wa = self.w("W_pc_deconv_a").permute(3, 2, 0, 1)
    v = F.relu(F.conv_transpose2d(h, wv, self.p["b_pc_deconv_v"], stride=2)).permute(0, 2, 3, 1)
    a = F.relu(F.conv_transpose2d(h, wa, self.p["b_pc_deconv_a"], stride=2)).permute(0, 2, 3, 1)
    q = v + a - a.mean(dim=3, keepdim=True)
    return q, q.max(dim=3).values

  def _zero_state(self, n):
    z = torch.zeros(n, 256, dtype=self.dtype)
    return z, z

  def pc_forward(self, images, lar):
    """PC tower from a zero LSTM state (model.py:393): images [L,N,84,84,3]."""
    L, N = images.shape[:2]
    enc = self.encoder(images.reshape(L * N, 84, 84, 3)).reshape(L, N, 9, 9, 32)
    out, _ = self.lstm_layer(enc, lar, *self._zero_state(N))
    q, qmax = self.pc_deconv(out.reshape(L * N, 256))
    return q.reshape(L, N, 20, 20, self.A), qmax.reshape(L, N, 20, 20)

  def vr_forward(self, images, lar):
    """VR tower from a zero LSTM state (model.py:459)."""
    L, N = images.shape[:2]
    enc = self.encoder(images.reshape(L * N, 84, 84, 3)).reshape(L, N, 9, 9, 32)
    out, _ = self.lstm_layer(enc, lar, *self._zero_state(N))
    return self.policy_value(out)[1]

  # model.py:473-488
  def rp_forward(self, images):
    """images [N,3,84,84,3] -> rp_c [N,3]."""
    N = images.shape[0]
    enc = self.encoder(images.reshape(N * 3, 84, 84, 3)).reshape(N, 7776)
    return F.softmax(enc @ self.w("W_rp_fc1") + self.p["b_rp_fc1"], dim=-1)

  # model.py:490-598 ------------------------------------------------------------------
  def base_loss(self, pi, v, a_onehot, adv, R, mask):
    log_pi = torch.log(pi.clamp(1e-20, 1.0))
    entropy = -(pi * log_pi).sum(-1)
    policy = -(((log_pi * a_onehot).sum(-1) * adv + entropy * self.entropy_beta) * mask).sum()
    value = 0.5 * 0.5 * (((R - v) ** 2) * mask).sum()          # 0.5 * tf.nn.l2_loss
    return policy, value

  def pc_loss(self, q, a_onehot, pc_r, mask):
    qa = (q * a_onehot[:, :, None, None, :]).sum(-1)
    return self.pc_lambda * 0.5 * (((pc_r - qa) ** 2) * mask[:, :, None, None]).sum()

  def vr_loss(self, v, R, mask):
    return 0.5 * (((R - v) ** 2) * mask).sum()

  def rp_loss(self, c, target):
    return -(target * torch.log(c.clamp(1e-20, 1.0))).sum()

  def total_loss(self, feed):
    """feed: dict of CPU tensors, time-major (see unreal_b200.model.model.UnrealModel.update).
    Returns (total, parts dict)."""
    parts = {}
    b = feed["base"]
    pi, v, _ = self.base_forward(b["images"], b["lar"], b["c0"], b["h0"])
    parts["policy"], parts["value"] = self.base_loss(pi, v, b["a"], b["adv"], b["R"], b["mask"])
    total = parts["policy"] + parts["value"]
    if "pc" in feed:
      f = feed["pc"]
      q, _ = self.pc_forward(f["images"], f["lar"])
      parts["pc"] = self.pc_loss(q, f["a"], f["R"], f["mask"])
      total = total + parts["pc"]
    if "vr" in feed:
      f = feed["vr"]
      parts["vr"] = self.vr_loss(self.vr_forward(f["images"], f["lar"]), f["R"], f["mask"])
      total = total + parts["vr"]
    if "rp" in feed:
      f = feed["rp"]
      parts["rp"] = self.rp_loss(self.rp_forward(f["images"]), f["c"])
      total = total + parts["rp"]
    return total, parts

  def loss_and_grads(self, feed):
    for t in self.p.values():
      t.requires_grad_(True)
      t.grad = None
    total, parts = self.total_loss(feed)
    total.backward()
    grads = OrderedDict((k, (t.grad if t.grad is not None else torch.zeros_like(t))) for k, t in self.p.items())
    for t in self.p.values():
      t.requires_grad_(False)
    return total.detach(), {k: v.detach() for k, v in parts.items()}, grads
