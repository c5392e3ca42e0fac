"""K2 parity: generic pixel change vs the oracle's literal numpy restatement (C ABI)."""
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


@pytest.mark.parametrize("shape", [(5, 84, 84, 3), (1, 84, 84, 3), (3, 44, 60, 3), (2, 84, 84, 1), (2, 36, 84, 4),
                                   (1, 260, 260, 3), (2, 80, 16, 3)])
@pytest.mark.parametrize("dt", ["f32", "u8"])
def test_pixel_change_matches_oracle(K, shape, dt):
  rs = np.random.RandomState(sum(shape))
  if dt == "u8":
    cur8 = rs.randint(0, 256, size=shape).astype(np.uint8)
    prev8 = rs.randint(0, 256, size=shape).astype(np.uint8)
    cur, prev = cur8.astype(np.float32) / 255.0, prev8.astype(np.float32) / 255.0   # lab_environment.py:99-102
    got = K.pixel_change(torch.from_numpy(cur8).cuda(), torch.from_numpy(prev8).cuda()).cpu().numpy()
  else:
    cur = rs.rand(*shape).astype(np.float32)
    prev = rs.rand(*shape).astype(np.float32)
    got = K.pixel_change(torch.from_numpy(cur).cuda(), torch.from_numpy(prev).cuda()).cpu().numpy()
  assert got.shape == (shape[0], (shape[1] - 4) // 4, (shape[2] - 4) // 4)
  for m in range(shape[0]):
    want32 = O.pixel_change(cur[m], prev[m])                       # numpy float32 evaluation
    want64 = O.pixel_change(cur[m].astype(np.float64), prev[m].astype(np.float64))
    assert want32.dtype == np.float32
    assert np.max(np.abs(got[m] - want64) / np.maximum(np.abs(want64), 1e-3)) <= REL
    assert np.array_equal(got[m], want32)                          # same order of roundings: bit-equal


def test_maze_frames_through_generic_kernel_equal_closed_form(K, golden_dir):
  """The generic kernel on rendered maze frames reproduces the reference maps of all 136 pairs."""
  import os
  with np.load(os.path.join(golden_dir, "maze_golden.npz")) as z:
    tab, pcs = z["pair_table"], z["pair_pc"]
  p0 = torch.from_numpy(tab[:, 0:2].astype(np.int32)).cuda()
  p1 = torch.from_numpy(tab[:, 3:5].astype(np.int32)).cuda()
  got = K.pixel_change(K.maze_render(p1), K.maze_render(p0)).cpu().numpy()
  assert np.array_equal(got, pcs.astype(np.float32))
  got8 = K.pixel_change(K.maze_render(p1, dtype=torch.uint8), K.maze_render(p0, dtype=torch.uint8)).cpu().numpy()
  assert np.array_equal(got8, pcs.astype(np.float32))


@pytest.mark.parametrize("dt", ["f32", "u8"])
def test_stream_form_equals_pairwise(K, dt):
  S, L = 7, 20
  g = torch.Generator(device="cuda").manual_seed(0)
  if dt == "u8":
    frames = torch.randint(0, 256, (S, L + 1, 84, 84, 3), dtype=torch.uint8, device="cuda", generator=g)
  else:
    frames = torch.rand(S, L + 1, 84, 84, 3, device="cuda", generator=g)
  got = K.pixel_change_stream(frames)
  assert got.shape == (S, L, 20, 20)
  for s in range(S):
    want = K.pixel_change(frames[s, 1:].contiguous(), frames[s, :-1].contiguous())
    assert torch.equal(got[s], want)
  # idempotence / symmetry properties
  same = K.pixel_change(frames[0, :3].contiguous(), frames[0, :3].contiguous())
  assert (same == 0).all()
  ab = K.pixel_change(frames[0, 0:1].contiguous(), frames[0, 1:2].contiguous())
  ba = K.pixel_change(frames[0, 1:2].contiguous(), frames[0, 0:1].contiguous())
  assert torch.equal(ab, ba)


def test_rejects_bad_shapes(K):
  from unreal_b200 import _lib
  a = torch.zeros(1, 85, 84, 3, device="cuda")
  with pytest.raises(_lib.UnrealError):
    K.pixel_change(a, a)
  b = torch.zeros(1, 84, 84, 2, device="cuda")
  with pytest.raises(_lib.UnrealError):
    K.pixel_change(b, b)
  assert K.pixel_change(torch.zeros(0, 84, 84, 3, device="cuda"), torch.zeros(0, 84, 84, 3, device="cuda")).shape == (0, 20, 20)


@pytest.mark.parametrize("dt", ["f32", "u8"])
def test_fast_84_path_equals_generic_kernel_on_many_sequences(K, dt):
  """The TMA-staged 84x84x3 kernel (several work items per CTA, 3-deep frame ring) against the
  generic kernel, bit for bit, in stream and pair form."""
  from unreal_b200 import _lib
  S, L = 1300, 3          # > resident CTAs, so every CTA walks several sequences
  g = torch.Generator(device="cuda").manual_seed(5)
  if dt == "u8":
    frames = torch.randint(0, 256, (S, L + 1, 84, 84, 3), dtype=torch.uint8, device="cuda", generator=g)
  else:
    frames = torch.rand(S, L + 1, 84, 84, 3, device="cuda", generator=g)
  fast = K.pixel_change_stream(frames)
  fast_pair = K.pixel_change(frames[:, 1].contiguous(), frames[:, 0].contiguous())
  _lib.set_tunable("pc84", 0)
  try:
    slow = K.pixel_change_stream(frames)
  finally:
    _lib.set_tunable("pc84", 1)
  assert torch.equal(fast, slow)
  assert torch.equal(fast_pair, slow[:, 0])
