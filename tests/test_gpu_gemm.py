"""K7 GEMM (tcgen05) against a float64 torch matmul of the same bf16-rounded operands.

Tolerance: operands are exactly representable, products are exact in fp32 and the accumulation
is fp32, so the only differences to the float64 reference are fp32 summation rounding
(<= K * 2^-24 relative to sum |a||b|) and, for bf16 outputs, the final rounding (2^-9).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, a_mn, b_mn, bias=None, add=None, relu=False):
  A = a.double().t() if a_mn else a.double()
  B = b.double() if b_mn else b.double().t()
  c = A @ B
  if bias is not None:
    c = c + bias.double()
  if add is not None:
    c = c + add.double()
  if relu:
    c = c.clamp_min(0)
  return c


def _mk(shape, seed, dev):
  g = torch.Generator(device=dev).manual_seed(seed)
  return (torch.randn(*shape, device=dev, generator=g) * 0.5).to(torch.bfloat16)


CASES = [
    # m, n, k, a_mn, b_mn
    (128, 128, 64, False, False),
    (256, 128, 256, False, False),
    (1000, 256, 2592, False, False),     # fc1 with a K-major weight shadow
    (1000, 256, 2592, False, True),      # fc1 with TF's [in,out] layout (model.py:337-340)
    (300, 1024, 264, False, True),       # LSTM input projection
    (777, 2592, 256, False, True),       # pc_fc1 (model.py:424)
    (4000, 16, 192, False, False),       # conv1 as im2col GEMM
    (810, 32, 256, False, False),        # conv2
    (810, 80, 32, False, False),         # deconv taps
    (555, 2592, 256, False, False),      # dgrad through fc1: dY [S,256] x W[2592,256]^T
    (2592, 256, 5000, True, True),       # wgrad fc1: X^T dY
    (16, 192, 4000, True, True),         # wgrad conv1 (small M padded by TMA)
    (264, 1024, 640, True, True),
    (96, 64, 200, True, False),
]


@pytest.mark.parametrize("m,n,k,a_mn,b_mn", CASES)
def test_gemm_matches_float64(m, n, k, a_mn, b_mn):
  from unreal_b200 import kernels as K
  dev = torch.device("cuda", 0)
  a = _mk((k, m) if a_mn else (m, k), 1, dev)
  b = _mk((k, n) if b_mn else (n, k), 2, dev)
  c = K.gemm_bf16(a, b, a_mn_major=a_mn, b_mn_major=b_mn)
  torch.cuda.synchronize()
  ref = _ref(a, b, a_mn, b_mn)
  scale = (a.double().abs().t() if a_mn else a.double().abs()) @ (b.double().abs() if b_mn else b.double().abs().t())
  err = ((c.double() - ref).abs() / scale.clamp_min(1e-6)).max().item()
  assert err <= k * 2.0 ** -23, err


def test_gemm_epilogue_bias_add_relu_bf16():
  from unreal_b200 import kernels as K
  dev = torch.device("cuda", 0)
  m, n, k = 700, 2592, 256
  a, b = _mk((m, k), 3, dev), _mk((k, n), 4, dev)
  bias = torch.randn(n, device=dev)
  add = torch.randn(m, n, device=dev)
  ref = _ref(a, b, False, True, bias, add, True)
  c32 = K.gemm_bf16(a, b, b_mn_major=True, bias=bias, add=add, relu=True)
  assert torch.allclose(c32.double(), ref, rtol=1e-4, atol=1e-4)
  c16 = K.gemm_bf16(a, b, b_mn_major=True, bias=bias, add=add, relu=True, out_dtype=torch.bfloat16)
  assert torch.allclose(c16.double(), ref, rtol=2.0 ** -7, atol=1e-2)


def test_gemm_strided_views_and_ragged_k():
  from unreal_b200 import kernels as K
  dev = torch.device("cuda", 0)
  m, n, k = 333, 1024, 261                      # LSTM input: 256 + A + 1 columns of a 264-wide buffer
  abuf = _mk((m, 264), 5, dev)
  wbuf = _mk((261 + 256, n), 6, dev)            # [in, out] kernel of BasicLSTMCell: x rows then h rows
  a = abuf[:, :k]
  b = wbuf[:k]
  out_buf = torch.zeros(m, n + 8, device=dev)
  c = K.gemm_bf16(a, b, out=out_buf[:, :n], b_mn_major=True)
  ref = _ref(a, b, False, True)
  assert torch.allclose(c.double(), ref, rtol=1e-4, atol=1e-4)
  assert float(out_buf[:, n:].abs().max()) == 0.0


def test_gemm_split_k_and_accumulate():
  from unreal_b200 import kernels as K
  dev = torch.device("cuda", 0)
  kk, m, n = 20000, 2592, 256
  x, dy = _mk((kk, m), 7, dev), _mk((kk, n), 8, dev)
  ref = _ref(x, dy, True, True)
  c = K.gemm_bf16(x, dy, a_mn_major=True, b_mn_major=True, split_k=7)
  assert torch.allclose(c.double(), ref, rtol=1e-3, atol=1e-2)
  K.gemm_bf16(x, dy, out=c, a_mn_major=True, b_mn_major=True, accumulate=True, split_k=3)
  assert torch.allclose(c.double(), 2 * ref, rtol=1e-3, atol=2e-2)


def test_gemm_rejects_bad_arguments():
  from unreal_b200 import _lib, kernels as K
  dev = torch.device("cuda", 0)
  a, b = _mk((64, 60), 1, dev), _mk((64, 60), 2, dev)    # ld 60 is not a multiple of 8
  with pytest.raises(_lib.UnrealError):
    K.gemm_bf16(a, b)
  with pytest.raises(_lib.UnrealError):
    K.gemm_bf16(_mk((64, 64), 1, dev), _mk((64, 72), 2, dev))
