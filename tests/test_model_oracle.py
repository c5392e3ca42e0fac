"""CPU checks of oracle/model_oracle.py (the PyTorch restatement of model/model.py).

The reference pins only the variable count (model/model_test.py:8-77: 20 / 18 / 12 / 14); the op
semantics the restatement relies on (TF conv2d NHWC/HWIO, conv2d_transpose with a
[kh,kw,out,in] filter, BasicLSTMCell gate order and forget bias) are checked here against
direct loop implementations of TF's documented definitions.  Parity with TF itself is UNPINNED.
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
