"""Out-of-bounds WRITE check of the round-2 kernels without a sanitizer (compute-sanitizer is closed on this pool): every
output lives inside a larger allocation whose margins hold a sentinel, sizes are deliberately off the CTA / warp / vector
granularity, and the margins must come back untouched."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PAD = 4096      # bytes of sentinel on each side (a multiple of 16: the views stay 16-byte aligned)


class Guarded(object):
  def __init__(self, shape, dtype):
    self.dtype = dtype
    self.n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    body = (self.n + 15) // 16 * 16
    self.raw = torch.full((PAD + body + PAD,), 0xA5, dtype=torch.uint8, device=DEV)
    self.t = self.raw[PAD:PAD + self.n].view(dtype).view(*shape)

  def ok(self):
    tail = self.raw[PAD + self.n:]
    return bool((self.raw[:PAD] == 0xA5).all()) and bool((tail == 0xA5).all())


def _check(*gs):
  torch.cuda.synchronize()
  for i, g in enumerate(gs):
    assert g.ok(), "output %d was written outside its bounds" % i


@pytest.fixture(scope="module")
def K():
  from unreal_b200 import kernels, _lib
  _lib.require_device()
  return kernels


@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8], ids=["f32", "u8"])
@pytest.mark.parametrize("variant", [-1, 0, 2, 3])
def test_maze_window_writes_stay_inside(K, dtype, variant):
  from unreal_b200 import _lib
  g = torch.Generator(device=DEV).manual_seed(1)
  for n, t in ((1, 1), (7, 13), (67, 32)):
    st = K.MazeState(n, DEV)
    acts = torch.randint(0, 4, (t, n), device=DEV, dtype=torch.int32, generator=g)
    obs, pc = Guarded((t, n, 84, 84, 3), dtype), Guarded((t, n, 20, 20), torch.float32)
    rew, term, rec = Guarded((t, n), torch.float32), Guarded((t, n), torch.uint8), Guarded((t, n), torch.int64)
    _lib.set_tunable("maze_render_variant", variant)
    try:
      K.maze_window(st, acts, obs=obs.t, pc=pc.t, reward=rew.t, terminal=term.t, frame_rec=rec.t, auto_reset=True)
    finally:
      _lib.set_tunable("maze_render_variant", -1)
    _check(obs, pc, rew, term, rec)


def test_pixel_change_u8_and_s2d_writes_stay_inside(K):
  g = torch.Generator(device=DEV).manual_seed(2)
  for s, l in ((1, 1), (3, 4), (151, 2)):           # fewer / more sequences than SMs
    frames = torch.randint(0, 256, (s, l + 1, 84, 84, 3), dtype=torch.uint8, device=DEV, generator=g)
    out = Guarded((s, l, 20, 20), torch.float32)
    K.pixel_change_stream(frames, out.t)
    pair = Guarded((s, 20, 20), torch.float32)
    K.pixel_change(frames[:, 1].contiguous(), frames[:, 0].contiguous(), pair.t)
    xs = Guarded((s, 6, 441, 8), torch.bfloat16)
    K.s2d_frames(frames[:, 0].contiguous(), xs.t)
    _check(out, pair, xs)
  sub = Guarded((3, 6, 4), torch.float32)
  K.subsample(torch.rand(3, 30, 20, device=DEV, generator=g), 5, sub.t)
  _check(sub)


def test_cell_table_rp_rows_select_writes_stay_inside(K):
  g = torch.Generator(device=DEV).manual_seed(3)
  for s in (1, 77, 4099):
    pos = torch.stack((torch.randint(0, 7, (s,), device=DEV, generator=g), torch.randint(0, 7, (s,), device=DEV, generator=g)),
                      1).int().contiguous()
    tab = torch.randn(49, 256, device=DEV, generator=g).to(torch.bfloat16)
    out = Guarded((s, 256), torch.bfloat16)
    K.cell_gather(tab, pos, out.t)
    seg = Guarded((49, 256), torch.float32)
    seg.t.zero_()
    K.cell_segment_sum(torch.randn(s, 256, device=DEV, generator=g).to(torch.bfloat16), pos, seg.t)
    _check(out, seg)
  n = 77
  lg = torch.randn(n, 8, device=DEV, generator=g)
  c = torch.nn.functional.one_hot(torch.randint(0, 3, (n,), device=DEV, generator=g), 3).float()
  res = K.rp_loss(lg, torch.zeros(3, device=DEV), c, want_p=True, want_loss=True, want_grad=True)
  assert all(torch.isfinite(v.float()).all() for v in res.values())
  frames = torch.randint(0, 256, (9, 84, 84, 3), dtype=torch.uint8, device=DEV, generator=g)
  sel = Guarded((5, 84, 84, 3), torch.uint8)
  sel.t.zero_()
  K.rows_select(sel.t, frames, torch.tensor([8, 0, 3, 3, 7], device=DEV), torch.tensor([1, 0, 1, 1, 1], dtype=torch.uint8, device=DEV))
  _check(sel)
  assert torch.equal(sel.t[0], frames[8]) and not sel.t[1].any() and torch.equal(sel.t[4], frames[7])


def test_lstm_bf16_cells_and_fused_pc_head_writes_stay_inside(K):
  from unreal_b200.model.model import UnrealModel
  g = torch.Generator(device=DEV).manual_seed(4)
  for n in (1, 37, 129):
    gates = Guarded((n, 1024), torch.bfloat16)
    gates.t.copy_(torch.randn(n, 1024, device=DEV, generator=g))
    cprev = torch.randn(n, 256, device=DEV, generator=g)
    cout, hout = Guarded((n, 256), torch.float32), Guarded((n, 256), torch.float32)
    xh = Guarded((n, 520), torch.bfloat16)
    xh.t.zero_()
    K.lstm_cell_fwd(gates.t, cprev, cout.t, hout.t, xh.t[:, 264:])
    assert not xh.t[:, :264].any()                      # only the h columns of the step operand are written
    dg, dc = Guarded((n, 1024), torch.bfloat16), Guarded((n, 256), torch.float32)
    dc.t.zero_()
    K.lstm_cell_bwd(gates.t, cprev, cout.t, hout.t, dc.t, dg.t, None)
    hact = Guarded((n, 256), torch.float32)
    K.lstm_cell_act(gates.t, cout.t, hout.t, hact.t, (torch.arange(n, device=DEV) % 3 != 0).to(torch.uint8))
    _check(gates, cout, hout, xh, dg, dc, hact)
  m = UnrealModel(4, 0, -1, True, True, True, True, 0.05, 0.001, DEV, {'segnet_mode': 0}, (84, 84), True, 0, 0.0, 0.0,
                  num_envs=3, seed=1)
  for s in (1, 7, 300):
    hp = torch.relu(torch.randn(s, 2592, device=DEV, generator=g)).to(torch.bfloat16)
    q = Guarded((s, 20, 20), torch.float32)
    K.pc_deconv_qmax(hp, m.pc_taps, m.pc_b8, 4, q.t)
    _check(q)
    loss, dy16, db8 = K.pc_deconv_loss(hp, m.pc_taps, m.pc_b8, torch.randint(0, 4, (s,), device=DEV, dtype=torch.int32, generator=g),
                                       torch.rand(s, 400, device=DEV, generator=g), torch.ones(s, device=DEV), 4, 0.05)
    assert torch.isfinite(loss).all() and not dy16[:, :, 8:].any()      # the padding channels are written as zeros


@pytest.mark.parametrize("s", [1, 5, 297, 601])
def test_pixel_control_backward_kernels_write_inside_their_buffers(K, s):
  """The last third of round 2: fused deconv + loss with its three gradient layouts (conv2's [S,400,16], [S,400,8], four
  parity planes), the backward convolutions with pc_fc1's ReLU mask (16 / 8 channels, planes), the filter gradients, and
  the one-pass maze Q-target scan -- through the C ABI with caller-owned, guarded outputs; sample counts below / off /
  above the 2 x 148 CTAs the persistent kernels launch."""
  from unreal_b200._lib import call, ptr, stream_ptr
  A = 4
  g = torch.Generator(device=DEV).manual_seed(100 + s)
  hp = torch.relu(torch.randn(s, 2592, device=DEV, generator=g)).to(torch.bfloat16)
  w8 = torch.zeros(4, 4, 8, 32, device=DEV, dtype=torch.bfloat16)
  w8[:, :, :1 + A] = (torch.randn(4, 4, 1 + A, 32, device=DEV, generator=g) * 0.05).to(torch.bfloat16)
  w16 = torch.zeros(4, 4, 16, 32, device=DEV, dtype=torch.bfloat16); w16[:, :, :8] = w8
  b8 = torch.zeros(8, device=DEV)
  taps = K.pc_deconv_taps(w8)
  act = torch.randint(0, A, (s,), device=DEV, generator=g, dtype=torch.int32)
  tgt = torch.rand(s, 400, device=DEV, generator=g)
  msk = torch.ones(s, device=DEV)
  sc = torch.tensor([0.7], device=DEV)
  for fn, shape in (("unreal_pc_deconv_loss", (s, 400, 16)), ("unreal_pc_deconv_loss_c8", (s, 400, 8)),
                    ("unreal_pc_deconv_loss_planes", (s, 4, 100, 8))):
    dy, loss, db = Guarded(shape, torch.bfloat16), Guarded((1,), torch.float64), Guarded((8,), torch.float32)
    loss.t.zero_(); db.t.zero_()
    call(fn, ptr(hp), ptr(taps), ptr(b8), ptr(act), ptr(tgt), ptr(msk), A, 0.05, s, ptr(loss.t), ptr(dy.t), ptr(db.t), stream_ptr())
    _check(dy, loss, db)
    assert bool(torch.isfinite(dy.t.float()).all())
    out, dbf = Guarded((s, 9, 9, 32), torch.bfloat16), Guarded((2592,), torch.float32)
    dbf.t.zero_()
    if shape[1] == 4:
      dw = Guarded((4, 4, 8, 32), torch.float32); dw.t.zero_()
      call("unreal_pc_planes_conv", ptr(dy.t), ptr(K.pc_w_planes(w8)), ptr(sc), ptr(hp), ptr(out.t), ptr(dbf.t), s, stream_ptr())
      call("unreal_pc_planes_wgrad", ptr(dy.t), ptr(hp), ptr(dw.t), s, stream_ptr())
    else:
      c = shape[2]
      dw = Guarded((4, 4, c, 32), torch.float32); dw.t.zero_()
      call("unreal_conv2_fwd_linear_masked", ptr(dy.t), c, ptr(K.conv_taps(w16 if c == 16 else w8, 2)), ptr(sc), ptr(hp), ptr(out.t),
           ptr(dbf.t), s, stream_ptr())
      call("unreal_conv2_wgrad" if c == 16 else "unreal_conv2_wgrad_c8", ptr(dy.t), ptr(hp), ptr(dw.t), s, stream_ptr())
    _check(out, dbf, dw)
    assert bool(torch.isfinite(out.t.float()).all()) and bool(torch.isfinite(dw.t).all()) and bool(torch.isfinite(dbf.t).all())
  t = 7
  p0 = torch.randint(0, 7, (t, s, 2), device=DEV, generator=g, dtype=torch.int32)
  p1 = torch.randint(0, 7, (t, s, 2), device=DEV, generator=g, dtype=torch.int32)
  ln = torch.randint(0, t + 1, (s,), device=DEV, generator=g, dtype=torch.int32)
  boot = torch.rand(s, 20, 20, device=DEV, generator=g)
  tg = Guarded((t, s, 20, 20), torch.float32)
  K.maze_pc_targets(p0, p1, ln, boot, 0.9, out=tg.t)
  _check(tg)
