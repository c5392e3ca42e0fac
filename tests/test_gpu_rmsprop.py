"""K6 parity: fused clip + RMSProp vs the reference's known-answer test and the oracle (C ABI)."""
import numpy as np
import pytest
import torch

from oracle import unreal_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-5


@pytest.fixture(scope="module")
def K():
  from unreal_b200 import kernels, _lib
  _lib.require_device()
  return kernels


def _t(a):
  return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_reference_known_answer(K):
  """train/rmsprop_applier_test.py:9-53: lr=2, decay=.9, momentum=0, eps=1, rms0=1."""
  var = _t(np.array([1.0, 2.0], np.float32)); rms = torch.ones(2, device="cuda"); mom = torch.zeros(2, device="cuda")
  norm = torch.zeros(1, device="cuda")
  for grad, want_var, want_ms in (([2.0, 4.0], [-1.6375218935831484, -2.2761798705987903], [1.3, 2.5]),
                                  ([3.0, 6.0], [-5.061902766795246, -6.861144190004011], [2.07, 5.85])):
    g = _t(np.array(grad, np.float32))
    ss = K.grad_sumsq(g)
    K.rmsprop_update(var, rms, mom, g, ss, lr=2.0, decay=0.9, momentum=0.0, eps=1.0, clip_norm=40.0, grad_norm=norm)
    np.testing.assert_allclose(var.cpu().numpy(), want_var, rtol=1e-6)
    np.testing.assert_allclose(rms.cpu().numpy(), want_ms, rtol=1e-6)
    np.testing.assert_allclose(norm.item(), np.sqrt(np.sum(np.square(grad))), rtol=1e-6)


@pytest.mark.parametrize("P", [1898877, 1, 3, 1027, 4096])
@pytest.mark.parametrize("momentum,gscale", [(0.0, 1.0), (0.9, 1.0), (0.0, 0.125)])
def test_matches_oracle(K, P, momentum, gscale):
  rs = np.random.RandomState(P % 1000 + int(momentum * 10))
  var = rs.randn(P).astype(np.float32); rms = (1 + rs.rand(P)).astype(np.float32)
  mom = (rs.randn(P) * 0.01).astype(np.float32) if momentum else np.zeros(P, np.float32)
  big = P > 100000
  grad = (rs.randn(P) * (0.2 if big else 0.01)).astype(np.float32)   # big: norm > 40 -> clipping active
  dvar, drms, dgrad = _t(var), _t(rms), _t(grad)
  dmom = _t(mom) if momentum else None
  norm = torch.zeros(1, device="cuda")
  ss = K.grad_sumsq(dgrad)
  K.rmsprop_update(dvar, drms, dmom, dgrad, ss, lr=7e-4, decay=0.99, momentum=momentum, eps=0.1, clip_norm=40.0,
                   grad_scale=gscale, grad_norm=norm)
  for dt, tol in ((np.float32, 2e-6), (np.float64, REL)):
    v, r, m = [var.astype(dt)], [rms.astype(dt)], [mom.astype(dt)]
    n = O.rmsprop_step(v, r, m, [grad.astype(dt) * dt(gscale)], 7e-4, 0.99, momentum, 0.1, 40.0, dt)
    np.testing.assert_allclose(norm.item(), n, rtol=1e-5)
    np.testing.assert_allclose(dvar.cpu().numpy(), v[0], rtol=tol, atol=tol * 1e-3)
    np.testing.assert_allclose(drms.cpu().numpy(), r[0], rtol=tol)
    if momentum:
      np.testing.assert_allclose(dmom.cpu().numpy(), m[0], rtol=tol, atol=1e-9)
  if big:
    assert gscale != 1.0 or norm.item() > 40.0   # the clip path is exercised


def test_sumsq_accumulates_and_no_clip(K):
  g = torch.full((1000,), 2.0, device="cuda")
  ss = K.grad_sumsq(g)
  K.grad_sumsq(g, ss)
  assert ss.item() == 8000.0
  var = torch.zeros(1000, device="cuda"); rms = torch.ones(1000, device="cuda")
  K.rmsprop_update(var, rms, None, g, None, lr=1.0, decay=0.5, momentum=0.0, eps=0.0, clip_norm=0.0)
  np.testing.assert_allclose(rms.cpu().numpy(), 2.5, rtol=1e-7)
  np.testing.assert_allclose(var.cpu().numpy(), -2.0 / np.sqrt(2.5), rtol=1e-6)
