"""CPU checks of oracle/model_oracle.py (the PyTorch restatement of model/model.py).

Pinned to the reference: `test_oracle_matches_the_references_model_py` compares it with vectors produced by the reference's own
model/model.py running over a TF-1 op shim (tests/golden/make_model_golden.py).  TensorFlow itself cannot run here, so the op
semantics both rely on (TF conv2d NHWC/HWIO, conv2d_transpose with a [kh,kw,out,in] filter, BasicLSTMCell gate order and
forget bias) are checked against direct loop implementations of TF's documented definitions, torch.nn.LSTMCell, float64
finite differences and hand-computed heads; the reference's model_test.py pins the variable counts (20 / 18 / 12 / 14).
"""
import numpy as np
import torch

from oracle import model_oracle as M


def test_variable_counts_match_model_test():
  # model_test.py: all heads 20, pc only 18, vr only 12, rp only 14
  n = lambda **k: len(M.variable_specs(4, 0, **k))  # noqa: E731
  assert n(use_pixel_change=True, use_value_replay=True, use_reward_prediction=True) == 20
  assert n(use_pixel_change=True, use_value_replay=False, use_reward_prediction=False) == 18
  assert n(use_pixel_change=False, use_value_replay=True, use_reward_prediction=False) == 12
  assert n(use_pixel_change=False, use_value_replay=False, use_reward_prediction=True) == 14


def test_parameter_count_matches_survey():
  count = lambda A: sum(int(np.prod(s)) for _, s, _ in M.variable_specs(A, 0))  # noqa: E731
  assert count(4) == 1898877 and count(3) == 1897083 and count(6) == 1902465


def test_conv_is_tf_nhwc_hwio_valid():
  rs = np.random.RandomState(0)
  p = M.init_params(4, seed=1)
  o = M.ModelOracle(p, 4)
  img = torch.tensor(rs.rand(1, 84, 84, 3), dtype=torch.float32)
  h2 = o.encoder(img)
  assert tuple(h2.shape) == (1, 9, 9, 32)
  # direct definition for a few outputs of conv1 then conv2
  W1, b1 = p["W_base_conv1"].numpy(), p["b_base_conv1"].numpy()
  W2, b2 = p["W_base_conv2"].numpy(), p["b_base_conv2"].numpy()
  x = img.numpy()[0]
  h1 = np.zeros((20, 20, 16), np.float64)
  for oy in range(20):
    for ox in range(20):
      patch = x[4 * oy:4 * oy + 8, 4 * ox:4 * ox + 8, :]
      h1[oy, ox] = np.maximum(np.tensordot(patch, W1, axes=([0, 1, 2], [0, 1, 2])) + b1, 0)
  for (oy, ox) in ((0, 0), (3, 7), (8, 8)):
    patch = h1[2 * oy:2 * oy + 4, 2 * ox:2 * ox + 4, :]
    want = np.maximum(np.tensordot(patch, W2, axes=([0, 1, 2], [0, 1, 2])) + b2, 0)
    assert np.allclose(h2[0, oy, ox].numpy(), want, rtol=1e-4, atol=1e-5)


def test_deconv_is_tf_conv2d_transpose():
  p = M.init_params(4, seed=2)
  o = M.ModelOracle(p, 4)
  rs = np.random.RandomState(3)
  lstm_out = torch.tensor(rs.randn(1, 256) * 0.5, dtype=torch.float32)
  q, qmax = o.pc_deconv(lstm_out)
  h = np.maximum(lstm_out.numpy() @ p["W_pc_fc1"].numpy() + p["b_pc_fc1"].numpy(), 0).reshape(9, 9, 32)

  def deconv(W, b):   # tf.nn.conv2d_transpose, filter [kh,kw,out,in], stride 2, VALID
    out = np.zeros((20, 20, W.shape[2]))
    for i in range(9):
      for j in range(9):
        for kh in range(4):
          for kw in range(4):
            out[2 * i + kh, 2 * j + kw] += W[kh, kw] @ h[i, j]
    return np.maximum(out + b, 0)

  v = deconv(p["W_pc_deconv_v"].numpy(), p["b_pc_deconv_v"].numpy())
  a = deconv(p["W_pc_deconv_a"].numpy(), p["b_pc_deconv_a"].numpy())
  want = v + a - a.mean(axis=2, keepdims=True)
  assert np.allclose(q[0].numpy(), want, rtol=1e-4, atol=1e-5)
  assert np.allclose(qmax[0].numpy(), want.max(axis=2), rtol=1e-4, atol=1e-5)


def test_lstm_cell_is_basic_lstm_cell():
  p = M.init_params(4, seed=4)
  o = M.ModelOracle(p, 4)
  rs = np.random.RandomState(5)
  conv = torch.tensor(rs.rand(2, 1, 9, 9, 32), dtype=torch.float32)
  lar = torch.tensor(rs.rand(2, 1, 5), dtype=torch.float32)
  c0 = torch.tensor(rs.randn(1, 256) * 0.1, dtype=torch.float32)
  h0 = torch.tensor(rs.randn(1, 256) * 0.1, dtype=torch.float32)
  out, (c, h) = o.lstm_layer(conv, lar, c0, h0)
  sig = lambda z: 1 / (1 + np.exp(-z))  # noqa: E731
  cc, hh = c0.numpy().astype(np.float64), h0.numpy().astype(np.float64)
  for t in range(2):
    fc = np.maximum(conv[t].reshape(1, 2592).numpy() @ p["W_base_fc1"].numpy() + p["b_base_fc1"].numpy(), 0)
    z = np.concatenate([fc, lar[t].numpy(), hh], axis=1) @ p["lstm_kernel"].numpy() + p["lstm_bias"].numpy()
    i, j, f, g = z[:, :256], z[:, 256:512], z[:, 512:768], z[:, 768:]
    cc = cc * sig(f + 1.0) + sig(i) * np.tanh(j)
    hh = np.tanh(cc) * sig(g)
    assert np.allclose(out[t].numpy(), hh, rtol=1e-4, atol=1e-5)
  assert np.allclose(c.numpy(), cc, rtol=1e-4, atol=1e-5)


def test_losses_follow_model_py():
  o = M.ModelOracle(M.init_params(4, seed=6), 4, pixel_change_lambda=0.05, entropy_beta=0.001)
  pi = torch.tensor([[[0.1, 0.2, 0.3, 0.4]], [[0.25, 0.25, 0.25, 0.25]]])
  v = torch.tensor([[0.5], [-0.5]])
  a = torch.tensor([[[0., 1, 0, 0]], [[0., 0, 0, 1]]])
  adv = torch.tensor([[2.0], [-1.0]]); R = torch.tensor([[1.0], [0.0]]); mask = torch.ones(2, 1)
  pol, val = o.base_loss(pi, v, a, adv, R, mask)
  ent = [-(np.array(p_) * np.log(p_)).sum() for p_ in ([0.1, 0.2, 0.3, 0.4], [0.25] * 4)]
  want = -((np.log(0.2) * 2.0 + ent[0] * 0.001) + (np.log(0.25) * -1.0 + ent[1] * 0.001))
  assert abs(float(pol) - want) < 1e-5
  assert abs(float(val) - 0.25 * (0.25 + 0.25)) < 1e-6      # 0.5 * tf.nn.l2_loss = 0.25 * sum sq


# ----------------------------------------------------------------------------------------------------------------
# Independent anchors of the restatement (TensorFlow cannot run here, so parity with TF stays unpinned; these tie
# the oracle to (a) another implementation of the same cell, (b) the true derivative of its own forward pass, and
# (c) numbers worked out by hand from model.py's formulas).
# ----------------------------------------------------------------------------------------------------------------
def _feed(T, N, A, seed, dtype=torch.float32):
  rs = np.random.RandomState(seed)
  t = lambda *s: torch.tensor(rs.rand(*s), dtype=dtype)      # noqa: E731
  tn = lambda *s: torch.tensor(rs.randn(*s), dtype=dtype)    # noqa: E731
  onehot = lambda *s: torch.eye(A, dtype=dtype)[torch.from_numpy(rs.randint(0, A, size=s))]   # noqa: E731
  lar = lambda L: torch.cat([onehot(L, N), tn(L, N, 1)], dim=2)     # noqa: E731
  mask = torch.ones(T, N, dtype=dtype)
  mask[T - 1, 0] = 0.0                                               # one env's window is a step shorter
  return {"base": dict(images=t(T, N, 84, 84, 3), lar=lar(T), a=onehot(T, N), adv=tn(T, N), R=tn(T, N), mask=mask,
                       c0=tn(N, 256) * 0.1, h0=tn(N, 256) * 0.1),
          "pc": dict(images=t(T, N, 84, 84, 3), lar=lar(T), a=onehot(T, N), R=t(T, N, 20, 20), mask=mask),
          "vr": dict(images=t(T, N, 84, 84, 3), lar=lar(T), R=tn(T, N), mask=mask),
          "rp": dict(images=t(N, 3, 84, 84, 3), c=torch.eye(3, dtype=dtype)[torch.from_numpy(rs.randint(0, 3, size=N))])}


def test_lstm_matches_torch_lstmcell_with_permuted_gates():
  """BasicLSTMCell (gates i, j, f, o on concat([x, h]) @ kernel, forget_bias 1.0 added to f; model.py:110) against
  torch.nn.LSTMCell (gates i, f, g, o; separate input / hidden weights, no forget bias): an independently written
  cell fed the SAME weights, columns permuted, must give the same unroll."""
  A = 4
  p = M.init_params(A, seed=8)
  p["lstm_bias"] = torch.tensor(np.random.RandomState(9).randn(1024) * 0.1, dtype=torch.float32)
  o = M.ModelOracle(p, A)
  rs = np.random.RandomState(10)
  T, N = 5, 3
  conv = torch.tensor(rs.rand(T, N, 9, 9, 32), dtype=torch.float32)
  lar = torch.tensor(rs.rand(T, N, A + 1), dtype=torch.float32)
  c0 = torch.tensor(rs.randn(N, 256) * 0.3, dtype=torch.float32)
  h0 = torch.tensor(rs.randn(N, 256) * 0.3, dtype=torch.float32)
  out, (c, h) = o.lstm_layer(conv, lar, c0, h0)

  n_in = 256 + A + 1
  k, b = p["lstm_kernel"], p["lstm_bias"]
  i_, j_, f_, o_ = [slice(q * 256, (q + 1) * 256) for q in range(4)]
  order = (i_, f_, j_, o_)                                          # TF (i, j, f, o) -> torch (i, f, g = j, o)
  cell = torch.nn.LSTMCell(n_in, 256)
  with torch.no_grad():
    cell.weight_ih.copy_(torch.cat([k[:n_in, s] for s in order], dim=1).t())
    cell.weight_hh.copy_(torch.cat([k[n_in:, s] for s in order], dim=1).t())
    bias = torch.cat([b[s] for s in order]).clone()
    bias[256:512] += 1.0                                            # forget_bias
    cell.bias_ih.copy_(bias)
    cell.bias_hh.zero_()
    fc = torch.relu(conv.reshape(T * N, 2592) @ p["W_base_fc1"] + p["b_base_fc1"]).reshape(T, N, 256)
    hh, cc = h0, c0
    for t in range(T):
      hh, cc = cell(torch.cat([fc[t], lar[t]], dim=1), (hh, cc))
      assert torch.allclose(out[t], hh, rtol=1e-5, atol=1e-6), t
  assert torch.allclose(c, cc, rtol=1e-5, atol=1e-6) and torch.allclose(h, hh, rtol=1e-5, atol=1e-6)


def test_gradients_match_fp64_finite_differences():
  """The gradient every CUDA gradient test is compared with must be the derivative of the oracle's own forward
  pass: central differences of the TOTAL loss (all four heads, a ragged mask, non-zero start state) in float64
  against autograd, at random coordinates of every one of the 20 variables."""
  A, T, N = 4, 2, 2
  p32 = M.init_params(A, seed=12)
  p = {k: v.double() * (3.0 if k.startswith("W_base_conv") else 1.0) for k, v in p32.items()}
  p["lstm_bias"] = torch.tensor(np.random.RandomState(13).randn(1024) * 0.1, dtype=torch.float64)
  o = M.ModelOracle(p, A, 0, 0.05, 0.001)
  feed = _feed(T, N, A, seed=14, dtype=torch.float64)
  _, parts, grads = o.loss_and_grads(feed)
  assert set(parts) == {"policy", "value", "pc", "vr", "rp"}
  rs = np.random.RandomState(15)
  eps = 1e-6
  worst = 0.0
  for name, g in grads.items():
    flat = p[name].view(-1)
    # coordinates where the gradient is large enough to be measured, plus random ones
    top = torch.topk(g.view(-1).abs(), min(3, flat.numel())).indices.tolist()
    picks = set(top) | set(int(i) for i in rs.randint(0, flat.numel(), size=3))
    for i in picks:
      keep = float(flat[i])
      with torch.no_grad():
        flat[i] = keep + eps
        lp = float(o.total_loss(feed)[0])
        flat[i] = keep - eps
        lm = float(o.total_loss(feed)[0])
        flat[i] = keep
      fd = (lp - lm) / (2 * eps)
      an = float(g.view(-1)[i])
      err = abs(fd - an) / max(abs(an), abs(fd), 1e-3)
      worst = max(worst, err)
      assert err <= 2e-5, (name, i, fd, an)
  assert worst <= 2e-5


def test_every_head_against_a_hand_computed_answer():
  """model.py's formulas worked out by hand for parameters that make every unit of a layer carry the same value, so
  the whole network collapses to scalar arithmetic (python floats below; the oracle runs in float64).  Pins the
  constants a restatement can get wrong: VALID output sizes, NHWC flatten, the forget bias, the (c, h) order, the
  zero start state of PC / VR, conv2d_transpose's overlap counts, the dueling mean, 0.5 * l2_loss = 0.25 * sum of
  squares for the base value loss vs 0.5 * sum of squares for value replay, lambda * 0.5 for pixel control, the
  entropy sign, and the reward-prediction class order."""
  import math
  A = 4
  sg = lambda z: 1.0 / (1.0 + math.exp(-z))      # noqa: E731
  p = {k: torch.zeros(s, dtype=torch.float64) for k, s, _ in M.variable_specs(A, 0)}
  p["b_base_conv1"] += 0.5                       # conv1: W = 0 -> h1 = 0.5 everywhere (20x20x16)
  p["W_base_conv2"] += 0.01; p["b_base_conv2"] += 0.1
  h2 = 4 * 4 * 16 * 0.5 * 0.01 + 0.1             # = 1.38 on all 9x9x32
  p["W_base_fc1"] += 0.001
  fc = 2592 * h2 * 0.001                         # every one of the 256 units
  n_in = 256 + A + 1
  bi, bj, bf, bo = 0.3, 0.7, -0.2, 0.1
  p["lstm_bias"][0:256] = bi; p["lstm_bias"][256:512] = bj; p["lstm_bias"][512:768] = bf; p["lstm_bias"][768:] = bo
  p["lstm_kernel"][:256, 256:512] = 0.0005       # fc units -> candidate gate j
  p["lstm_kernel"][256 + A, 0:256] = 0.05        # last reward -> input gate i
  p["lstm_kernel"][256 + 2, 512:768] = 0.4       # last action == 2 -> forget gate f
  p["lstm_kernel"][n_in:, 768:] = 0.001          # h_{t-1} units -> output gate o
  p["W_base_fc_p"][:, 1] = 0.01; p["b_base_fc_p"][3] = 0.2
  p["W_base_fc_v"] += 0.002; p["b_base_fc_v"] += 0.1
  p["W_pc_fc1"] += 0.001
  p["W_pc_deconv_v"] += 0.01; p["b_pc_deconv_v"] += 0.05
  for a in range(A):
    p["W_pc_deconv_a"][:, :, a, :] = 0.01 * (a + 1)
  p["b_pc_deconv_a"] += -0.02
  p["W_rp_fc1"][:, 0] = 1e-4; p["W_rp_fc1"][:, 1] = 2e-4; p["W_rp_fc1"][:, 2] = -1e-4
  p["b_rp_fc1"][2] = 0.3
  o = M.ModelOracle(p, A, 0, pixel_change_lambda=0.05, entropy_beta=0.001)

  def unroll(last_actions, last_rewards, c, h):
    """-> list of h_t (a scalar: all 256 units are equal), final (c, h)."""
    hs = []
    for la, lr in zip(last_actions, last_rewards):
      i = bi + 0.05 * lr
      j = bj + 256 * fc * 0.0005
      f = bf + (0.4 if la == 2 else 0.0)
      g = bo + 256 * h * 0.001
      c = c * sg(f + 1.0) + sg(i) * math.tanh(j)
      h = math.tanh(c) * sg(g)
      hs.append(h)
    return hs, c, h

  def heads(h):
    z = [0.0, 256 * h * 0.01, 0.0, 0.2]
    m = max(z)
    e = [math.exp(v - m) for v in z]
    pi = [v / sum(e) for v in e]
    return pi, 256 * h * 0.002 + 0.1

  T, N = 3, 1
  img = torch.full((T, N, 84, 84, 3), 0.37, dtype=torch.float64)     # conv1's W is zero: any image gives h1 = 0.5
  la, lr = [2, 0, 2], [0.0, 1.0, -1.0]
  lar = torch.zeros(T, N, A + 1, dtype=torch.float64)
  for t in range(T):
    lar[t, 0, la[t]] = 1.0; lar[t, 0, A] = lr[t]
  acts, adv, R = [1, 3, 0], [0.5, -1.5, 2.0], [1.0, 0.2, -0.4]
  a1 = torch.zeros(T, N, A, dtype=torch.float64)
  for t in range(T):
    a1[t, 0, acts[t]] = 1.0
  c0, h0 = 0.2, -0.1
  ones = torch.ones(T, N, dtype=torch.float64)
  pc_R = torch.full((T, N, 20, 20), 0.3, dtype=torch.float64)
  feed = {"base": dict(images=img, lar=lar, a=a1, adv=torch.tensor(adv, dtype=torch.float64).view(T, N),
                       R=torch.tensor(R, dtype=torch.float64).view(T, N), mask=ones,
                       c0=torch.full((N, 256), c0, dtype=torch.float64), h0=torch.full((N, 256), h0, dtype=torch.float64)),
          "pc": dict(images=img, lar=lar, a=a1, R=pc_R, mask=ones),
          "vr": dict(images=img, lar=lar, R=torch.tensor(R, dtype=torch.float64).view(T, N), mask=ones),
          "rp": dict(images=img[:, 0].unsqueeze(0), c=torch.tensor([[0.0, 0.0, 1.0]], dtype=torch.float64))}
  total, parts = o.total_loss(feed)

  # base (model.py:499-517): start state fed, policy = -sum(log pi(a) * adv + beta * H), value = 0.5 * l2_loss(R - V)
  hs, c_end, h_end = unroll(la, lr, c0, h0)
  pol = val = 0.0
  for t in range(T):
    pi, v = heads(hs[t])
    H = -sum(q * math.log(q) for q in pi)
    pol -= math.log(pi[acts[t]]) * adv[t] + 0.001 * H
    val += 0.25 * (R[t] - v) ** 2
  assert abs(float(parts["policy"]) - pol) <= 1e-12 * max(1.0, abs(pol))
  assert abs(float(parts["value"]) - val) <= 1e-12
  _, _, state = o.base_forward(img, lar, feed["base"]["c0"], feed["base"]["h0"])
  assert abs(float(state[0][0, 7]) - c_end) <= 1e-12 and abs(float(state[1][0, 200]) - h_end) <= 1e-12   # (c, h)

  # value replay (:556-565): ZERO start state (:459), 0.5 * sum of squares
  hz, _, _ = unroll(la, lr, 0.0, 0.0)
  vr = sum(0.5 * (R[t] - heads(hz[t])[1]) ** 2 for t in range(T))
  assert abs(float(parts["vr"]) - vr) <= 1e-12

  # pixel control (:411-443, :531-546): zero start state (:393); conv2d_transpose 4x4 stride 2 over 9x9 -> 20x20:
  # an output row y receives cnt(y) = #{(i, kh): 2i + kh = y} kernel taps: 1 at the two border rows on each side, else 2
  cnt = [sum(1 for i in range(9) for k in range(4) if 2 * i + k == y) for y in range(20)]
  assert cnt == [1, 1] + [2] * 16 + [1, 1]
  pc = 0.0
  for t in range(T):
    hp = max(256 * hz[t] * 0.001, 0.0)
    for y in range(20):
      for x in range(20):
        taps = cnt[y] * cnt[x] * 32 * hp
        v = max(taps * 0.01 + 0.05, 0.0)
        ad = [max(taps * 0.01 * (a + 1) - 0.02, 0.0) for a in range(A)]
        q = v + ad[acts[t]] - sum(ad) / A
        pc += 0.05 * 0.5 * (0.3 - q) ** 2
  assert abs(float(parts["pc"]) - pc) <= 1e-10 * max(1.0, pc)

  # reward prediction (:473-488, :571-575): 3 frames' features concatenated, classes [zero, positive, negative]
  z = [7776 * h2 * 1e-4, 7776 * h2 * 2e-4, 7776 * h2 * -1e-4 + 0.3]
  rp = -(z[2] - math.log(sum(math.exp(v) for v in z)))
  assert abs(float(parts["rp"]) - rp) <= 1e-12 * max(1.0, abs(rp))
  assert abs(float(total) - (pol + val + vr + pc + rp)) <= 1e-9 * max(1.0, abs(float(total)))


import pytest  # noqa: E402


@pytest.mark.parametrize("fixture", ["model_reference_golden.npz", "model_reference_golden_a3g2.npz"], ids=["maze-A4", "indoor-A3-G2"])
def test_oracle_matches_the_references_model_py(fixture):
  """tests/golden/model_reference_golden.npz was produced by the REFERENCE'S OWN model/model.py, imported unmodified and
  executed over tests/golden/tf1_shim (TensorFlow cannot be installed: its ~30 ops on this path are restated from their
  published definitions; the graph building, variable order and reuse, layer wiring, loss formulas and the run_* methods
  are the reference's code).  The oracle must reproduce, in float64: the three acting steps of run_base_policy_and_value
  with their carried LSTM state, run_base_value / run_pc_q_max / run_vr_value / run_rp_c, every loss term of a training
  feed assembled like Trainer.process (trainer.py:500-541) and the gradient of total_loss in all 20 variables
  (rmsprop_applier.py:109-116)."""
  import os
  g = np.load(os.path.join(os.path.dirname(__file__), "golden", fixture))
  A, G, seed, T, Lp, Lv = (int(x) for x in g["meta"])
  p = M.init_params(A, G, seed=seed)
  p = type(p)((k, v.to(torch.float64)) for k, v in p.items())
  o = M.ModelOracle(p, A, G, 0.05, 0.001)
  # the reference created the oracle's 20 variables, in the oracle's order
  last = [str(n).split("/")[-1].split(":")[0] for n in g["variable_names"]]
  want = [n.replace("lstm_kernel", "kernel").replace("lstm_bias", "bias") for n, _, _ in M.variable_specs(A, G)]
  assert last == want
  assert str(g["variable_names"][6]) == "net_0/base_lstm_layer/basic_lstm_cell/kernel:0"
  f64 = lambda k: torch.from_numpy(np.asarray(g[k], dtype=np.float64))          # noqa: E731
  img = lambda k: torch.from_numpy(g[k].astype(np.float64) / 255.0)             # noqa: E731
  close = lambda a, b: np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=1e-9, atol=1e-11)   # noqa: E731
  with torch.no_grad():
    # acting: step size 1, the state carried from call to call (model.py:617-660)
    c = h = torch.zeros(1, 256, dtype=torch.float64)
    for t in range(3):
      pi, v, (c, h) = o.base_forward(img("act_frames")[t][None, None], f64("act_lar")[t][None, None], c, h)
      close(pi[0, 0], g["act_pi"][t]); close(v[0, 0], g["act_v"][t])
    close(c, g["act_state_c"]); close(h, g["act_state_h"])
    # run_base_value reads the carried state and leaves it alone (model.py:687-704)
    _, v, _ = o.base_forward(img("one_frame")[0][None, None], f64("one_lar")[0][None, None], c, h)
    close(v[0, 0], g["base_value"])
    _, qmax = o.pc_forward(img("one_frame")[0][None, None], f64("one_lar")[0][None, None])
    close(qmax[0, 0], g["pc_q_max"])
    close(o.vr_forward(img("one_frame")[0][None, None], f64("one_lar")[0][None, None])[0, 0], g["vr_value"])
    close(o.rp_forward(img("rp_frames")[None])[0], g["rp_c"])
  ones = lambda n: torch.ones(n, 1, dtype=torch.float64)                        # noqa: E731
  feed = {"base": dict(images=img("base_frames")[:, None], lar=f64("base_lar")[:, None], a=f64("base_a")[:, None],
                       adv=f64("base_adv")[:, None], R=f64("base_R")[:, None], mask=ones(T), c0=f64("base_c0"), h0=f64("base_h0")),
          "pc": dict(images=img("pc_frames")[:, None], lar=f64("pc_lar")[:, None], a=f64("pc_a")[:, None], R=f64("pc_R")[:, None],
                     mask=ones(Lp)),
          "vr": dict(images=img("vr_frames")[:, None], lar=f64("vr_lar")[:, None], R=f64("vr_R")[:, None], mask=ones(Lv)),
          "rp": dict(images=img("rp_train_frames")[None], c=f64("rp_c_target"))}
  total, parts, grads = o.loss_and_grads(feed)
  close(total, g["total_loss"])
  for k, name in (("policy", "policy_loss"), ("value", "value_loss"), ("pc", "pc_loss"), ("vr", "vr_loss"), ("rp", "rp_loss")):
    close(parts[k], g[name])
  with torch.no_grad():
    pi, v, _ = o.base_forward(feed["base"]["images"], feed["base"]["lar"], feed["base"]["c0"], feed["base"]["h0"])
    close(pi[:, 0], g["base_pi"]); close(v[:, 0], g["base_v"])
    close(-(pi * torch.log(pi.clamp(1e-20, 1.0))).sum(-1)[:, 0], g["entropy"])
    close(o.pc_forward(feed["pc"]["images"], feed["pc"]["lar"])[0][:, 0], g["pc_q"])
    close(o.vr_forward(feed["vr"]["images"], feed["vr"]["lar"])[:, 0], g["vr_v"])
    close(o.rp_forward(feed["rp"]["images"]), g["train_rp_c"])
  for name, _, _ in M.variable_specs(A, G):
    gr = grads[name].reshape(-1).numpy()
    scale = max(float(g["grad_norm_" + name]), 1e-30)
    assert abs(gr.sum() - float(g["grad_sum_" + name])) <= 1e-9 * scale * np.sqrt(gr.size), name
    assert abs(np.sqrt((gr * gr).sum()) - scale) <= 1e-9 * scale, name
    np.testing.assert_allclose(gr[g["grad_idx_" + name]], g["grad_val_" + name], rtol=1e-8, atol=1e-9 * scale, err_msg=name)
  # ---- the learner step: sync_from + RMSPropApplier.minimize_local, run by the reference's own classes (two updates, the
  # second with a gradient above the clip norm) against oracle gradients + the oracle's clip / ApplyRMSProp restatement
  from oracle import unreal_oracle as O
  names = [n for n, _, _ in M.variable_specs(A, G)]
  vars_ = [p[n].numpy().copy() for n in names]
  start = [v.copy() for v in vars_]
  rms = [np.ones_like(v) for v in vars_]
  mom = [np.zeros_like(v) for v in vars_]
  assert float(g["update_grad_norms"][0]) < 40.0 < float(g["update_grad_norms"][1])
  for step in range(2):
    cur = type(p)((n, torch.from_numpy(v.copy())) for n, v in zip(names, vars_))
    _, _, grads = M.ModelOracle(cur, A, G, 0.05, 0.001).loss_and_grads(feed)
    norm = O.rmsprop_step(vars_, rms, mom, [grads[n].numpy() for n in names], float(g["update_lrs"][step]), decay=0.99,
                          momentum=0.0, epsilon=0.1, clip_norm=40.0, dtype=np.float64)
    assert abs(float(norm) - float(g["update_grad_norms"][step])) <= 1e-9 * float(norm)
    feed["base"]["adv"] = feed["base"]["adv"] * 30.0
  for n, v, v0, r in zip(names, vars_, start, rms):
    idx = g["upd_idx_" + n]
    np.testing.assert_allclose(v.reshape(-1)[idx], g["upd_val_" + n], rtol=1e-10, atol=1e-13, err_msg=n)
    np.testing.assert_allclose(r.reshape(-1)[idx], g["upd_rms_" + n], rtol=1e-10, atol=1e-13, err_msg=n)
    step_sum = float((v - v0).sum())
    assert abs(step_sum - float(g["upd_sum_" + n])) <= 1e-9 * (abs(float(g["upd_sum_" + n])) + 1e-12 * v.size), n


class _OracleNet(object):
  """oracle.ModelOracle behind the call surface RolloutOracle expects of a network (model.py:617-728): the carried LSTM
  state lives here like `base_lstm_state_out` lives in the reference's model."""

  def __init__(self, oracle):
    self.o = oracle
    self.reset_state()

  def reset_state(self):
    self.c = torch.zeros(1, 256, dtype=torch.float64)
    self.h = torch.zeros(1, 256, dtype=torch.float64)

  @staticmethod
  def _x(s_t, lar):
    img = s_t['image'] if isinstance(s_t, dict) else s_t
    return (torch.from_numpy(np.asarray(img, np.float64))[None, None], torch.from_numpy(np.asarray(lar, np.float64))[None, None])

  def run_base_policy_and_value(self, sess, s_t, lar, mode=""):
    with torch.no_grad():
      pi, v, (self.c, self.h) = self.o.base_forward(*self._x(s_t, lar), self.c, self.h)
    return pi[0, 0].numpy(), float(v[0, 0]), None

  def run_base_value(self, sess, s_t, lar):
    with torch.no_grad():
      return float(self.o.base_forward(*self._x(s_t, lar), self.c, self.h)[1][0, 0])

  def run_pc_q_max(self, sess, s_t, lar):
    with torch.no_grad():
      return self.o.pc_forward(*self._x(s_t, lar))[1][0, 0].numpy()

  def run_vr_value(self, sess, s_t, lar):
    with torch.no_grad():
      return float(self.o.vr_forward(*self._x(s_t, lar))[0, 0])


def test_oracle_agent_loop_matches_the_references_trainer():
  """tests/golden/agent_reference_golden.npz: the reference's own Trainer.process -- _fill_experience until the buffer is
  full, then four learner iterations (sync_from, _process_base / _pc / _vr / _rp, the feed dict, minimize_local's update) --
  with the reference's own model, Experience, maze and RMSPropApplier, all over the TF-1 op shim
  (tests/golden/make_agent_golden.py).  The oracle replays the loop from the same seeds: RolloutOracle for the host logic,
  ModelOracle for the network, the oracle's clip + ApplyRMSProp for the update.  After the fill and after EVERY iteration
  the step counter, the returned (steps, score), the agent's cell, the position of the worker's RandomState stream, the
  carried LSTM state and sampled entries + sums of all 20 global variables must agree (1e-8: float64 both sides, four
  chained updates).  This pins the glue between the separately pinned pieces -- feed order, start state, bootstrap, the
  replay samplers' share of the random stream, the learning-rate anneal -- to the reference's code."""
  import os
  from oracle import unreal_oracle as O
  g = np.load(os.path.join(os.path.dirname(__file__), "golden", "agent_reference_golden.npz"))
  A, G, seed, net_seed, H, n_proc, max_t = (int(x) for x in g["meta"])
  names = [n for n, _, _ in M.variable_specs(A, G)]
  vars_ = [v.to(torch.float64).numpy().copy() for v in M.init_params(A, G, seed=net_seed).values()]
  p = dict((n, torch.from_numpy(v)) for n, v in zip(names, vars_))         # views: the update below is seen by the network
  rms = [np.ones_like(v) for v in vars_]
  mom = [np.zeros_like(v) for v in vars_]
  oracle = M.ModelOracle(p, A, G, 0.05, 0.001)
  net = _OracleNet(oracle)
  rs = np.random.RandomState(seed)
  w = O.RolloutOracle(H, rs, net, n_step_TD=20)

  def check(tag, ret):
    assert int(g[tag + "_local_t"]) == w.local_t, tag
    want_ret = g[tag + "_ret"]
    assert ret[0] == int(want_ret[0]) and ((ret[1] is None) == bool(np.isnan(want_ret[1]))), (tag, ret, want_ret)
    if ret[1] is not None:
      assert ret[1] == want_ret[1]
    assert (w.env.x, w.env.y) == tuple(int(v) for v in g[tag + "_pos"]), tag
    probe = np.random.RandomState(); probe.set_state(rs.get_state())
    assert int(probe.randint(0, 2 ** 31 - 1)) == int(g[tag + "_next_draw"]), tag + ": the RandomState streams have diverged"
    np.testing.assert_allclose(net.c.numpy(), g[tag + "_lstm_c"], rtol=1e-8, atol=1e-10, err_msg=tag)
    np.testing.assert_allclose(net.h.numpy(), g[tag + "_lstm_h"], rtol=1e-8, atol=1e-10, err_msg=tag)
    for n, v in zip(names, vars_):
      np.testing.assert_allclose(v.reshape(-1)[g["idx_" + n]], g[tag + "_val_" + n], rtol=1e-8, atol=1e-11, err_msg=tag + " " + n)
      assert abs(v.sum() - float(g[tag + "_sum_" + n])) <= 1e-8 * max(1.0, np.abs(v).sum()), (tag, n)

  n_fill = 0
  while not w.ring.is_full():
    w.fill_step()
    n_fill += 1
  assert n_fill == int(g["n_fill"])
  check("fill", (0, None))
  t64 = lambda x: torch.from_numpy(np.asarray(x, np.float64))                 # noqa: E731
  imgs = lambda states: t64(np.stack([s['image'] for s in states]))           # noqa: E731
  global_t = 0
  for it in range(n_proc):
    lr = O.anneal_learning_rate(float(g["initial_lr"]), global_t, max_t)
    t0 = w.local_t
    c0, h0 = net.c.clone(), net.h.clone()                                     # start_lstm_state (trainer.py:228)
    b = w.process_base()
    pc, vr, rp = w.process_pc(), w.process_vr(), w.process_rp()
    ones = lambda n: torch.ones(n, 1, dtype=torch.float64)                    # noqa: E731
    feed = {"base": dict(images=imgs(b["states"])[:, None], lar=t64(b["lar"])[:, None], a=t64(b["a"])[:, None],
                         adv=t64(b["adv"])[:, None], R=t64(b["R"])[:, None], mask=ones(len(b["states"])), c0=c0, h0=h0),
            "pc": dict(images=imgs(pc["states"])[:, None], lar=t64(pc["lar"])[:, None], a=t64(pc["a"])[:, None],
                       R=t64(pc["R"])[:, None], mask=ones(len(pc["states"]))),
            "vr": dict(images=imgs(vr["states"])[:, None], lar=t64(vr["lar"])[:, None], R=t64(vr["R"])[:, None],
                       mask=ones(len(vr["states"]))),
            "rp": dict(images=imgs(rp["states"])[None], c=t64(rp["c"])[None])}
    _, _, grads = oracle.loss_and_grads(feed)
    O.rmsprop_step(vars_, rms, mom, [grads[n].numpy() for n in names], lr, decay=0.99, momentum=0.0, epsilon=0.1,
                   clip_norm=40.0, dtype=np.float64)
    steps = w.local_t - t0
    global_t += steps
    check("it%d" % it, (steps, b["score"]))
