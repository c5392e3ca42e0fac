"""K1 parity on the device: maze transitions / rewards / terminals / frames bit-exact,
pixel-change exactly equal to float32 of the reference's maps.  Calls go through the C ABI."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import unreal_oracle as O

pytestmark = pytest.mark.gpu


def _golden(golden_dir, name):
  with np.load(os.path.join(golden_dir, name)) as z:
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def K():
  from unreal_b200 import kernels, _lib
  _lib.require_device()
  kernels.maze_set_map(None)
  return kernels


@pytest.fixture(params=[0, 1, 2], ids=["direct", "tma", "warp"])
def variant(request, K):
  from unreal_b200 import _lib
  _lib.set_tunable("maze_render_variant", request.param)
  yield request.param
  _lib.set_tunable("maze_render_variant", -1)


def test_layout(K):
  start, goal, walls = K.maze_layout()
  assert start == O.START and goal == O.GOAL
  assert np.array_equal(np.frombuffer(walls, np.uint8).reshape(7, 7), O.WALLS.astype(np.uint8))


def test_golden_rollout_single_env(K, golden_dir, variant):
  """The reference's own 6000-step trajectory (seed 0xA3C), one env, no auto-reset."""
  g = _golden(golden_dir, "maze_golden.npz")
  dev = "cuda:0"
  st = K.MazeState(1, dev)
  obs = torch.empty(1, 84, 84, 3, device=dev)
  pc = torch.empty(1, 20, 20, device=dev)
  first = K.maze_render(st.pos)
  assert np.array_equal(first[0].cpu().numpy().astype(np.uint8), g["initial_frame"])
  h_tr, h_pc = hashlib.sha256(), hashlib.sha256()
  acts = torch.from_numpy(g["actions"].astype(np.int32)).to(dev)
  for i in range(len(g["actions"])):
    r, t = K.maze_step(st, acts[i:i + 1], obs=obs, pc=pc)
    x, y = st.pos[0].tolist()
    r, t = int(r.item()), int(t.item())
    assert (x, y, r, t) == (g["x"][i], g["y"][i], g["reward"][i], g["terminal"][i]), i
    h_tr.update(bytes([int(g["actions"][i]), x, y, r & 0xff, t]))
    h_pc.update(pc[0].cpu().numpy().astype(np.float32).tobytes())
    if i < 64:
      assert np.array_equal(pc[0].cpu().numpy(), g["first_pc"][i].astype(np.float32))
      assert np.array_equal(obs[0].cpu().numpy(), O.maze_render(x, y, np.float32))
    if t:
      K.maze_reset(st)
  assert h_tr.digest() == g["sha_transitions"].tobytes()
  assert h_pc.digest() == g["sha_pc_f32"].tobytes()


def test_all_cell_action_pairs(K, golden_dir, variant):
  g = _golden(golden_dir, "maze_golden.npz")
  tab = g["pair_table"].astype(np.int32)
  dev = "cuda:0"
  n = len(tab)
  st = K.MazeState(n, dev)
  st.pos.copy_(torch.from_numpy(tab[:, 0:2].copy()))
  obs = torch.full((n, 84, 84, 3), 7.0, device=dev)
  pc = torch.full((n, 20, 20), 7.0, device=dev)
  rec = torch.zeros(n, dtype=torch.int64, device=dev)
  r, t = K.maze_step(st, torch.from_numpy(tab[:, 2].copy()).to(dev), obs=obs, pc=pc, frame_rec=rec)
  assert np.array_equal(st.pos.cpu().numpy(), tab[:, 3:5])
  assert np.array_equal(r.cpu().numpy(), tab[:, 5].astype(np.float32))
  assert np.array_equal(t.cpu().numpy(), tab[:, 6].astype(np.uint8))
  assert np.array_equal(pc.cpu().numpy(), g["pair_pc"].astype(np.float32))
  want = np.stack([O.maze_render(int(x), int(y), np.float32) for x, y in tab[:, 3:5]])
  assert np.array_equal(obs.cpu().numpy(), want)
  assert np.array_equal(st.last_action.cpu().numpy(), tab[:, 2])
  assert np.array_equal(st.last_reward.cpu().numpy(), tab[:, 5].astype(np.float32))
  # frame records: x0|y0<<4, x1|y1<<4, action, reward(int8), terminal|valid
  rc = rec.cpu().numpy().view(np.uint8).reshape(n, 8)
  assert np.array_equal(rc[:, 0], tab[:, 0] | (tab[:, 1] << 4))
  assert np.array_equal(rc[:, 1], tab[:, 3] | (tab[:, 4] << 4))
  assert np.array_equal(rc[:, 2], tab[:, 2])
  assert np.array_equal(rc[:, 3].view(np.int8), tab[:, 5])
  assert np.array_equal(rc[:, 4], tab[:, 6] | 0x80)


@pytest.mark.parametrize("n", [1, 3, 147, 149, 1000])
def test_batched_random_walk_vs_oracle(K, variant, n):
  """N envs, ragged grid sizes, auto-reset on, u8 and f32 frames, 60 steps."""
  dev = "cuda:0"
  rs = np.random.RandomState(n)
  st = K.MazeState(n, dev)
  xs = np.full(n, O.START[0]); ys = np.full(n, O.START[1])
  la = np.zeros(n, np.int32); lr = np.zeros(n, np.float32)
  obs = torch.empty(n, 84, 84, 3, device=dev)
  obs8 = torch.empty(n, 84, 84, 3, dtype=torch.uint8, device=dev)
  pc = torch.empty(n, 20, 20, device=dev)
  for step in range(60):
    a = rs.randint(0, 4, size=n).astype(np.int32)
    if step % 7 == 0:
      a[::5] = 9          # out-of-range action: no move, reward 0
    use8 = step % 2 == 1
    r, t = K.maze_step(st, torch.from_numpy(a).to(dev), obs=obs8 if use8 else obs, pc=pc, auto_reset=True)
    want_pc = np.zeros((n, 20, 20), np.float32)
    want_r = np.zeros(n, np.float32); want_t = np.zeros(n, np.uint8)
    for e in range(n):
      nx, ny, rew, term = O.maze_step(int(xs[e]), int(ys[e]), int(a[e]))
      want_pc[e] = O.maze_pixel_change_closed_form(int(xs[e]), int(ys[e]), nx, ny)
      want_r[e] = rew; want_t[e] = term
      if term:
        xs[e], ys[e] = O.START; la[e] = 0; lr[e] = 0
      else:
        xs[e], ys[e] = nx, ny; la[e] = a[e]; lr[e] = rew
    assert np.array_equal(r.cpu().numpy(), want_r)
    assert np.array_equal(t.cpu().numpy(), want_t)
    assert np.array_equal(st.pos.cpu().numpy(), np.stack([xs, ys], 1))
    assert np.array_equal(st.last_action.cpu().numpy(), la)
    assert np.array_equal(st.last_reward.cpu().numpy(), lr)
    assert np.array_equal(pc.cpu().numpy(), want_pc)
    if step % 10 in (0, 1) or n <= 3:
      got = (obs8 if use8 else obs).cpu().numpy()
      for e in range(0, n, max(1, n // 37)):
        want = O.maze_render(int(xs[e]), int(ys[e]), np.float32)
        assert np.array_equal(got[e], want * 255 if use8 else want)


def test_active_mask_and_no_autoreset(K, variant):
  dev = "cuda:0"
  n = 64
  st = K.MazeState(n, dev)
  st.pos.copy_(torch.tensor([[5, 0]] * n, dtype=torch.int32))     # one RIGHT from the goal (6,0)
  active = torch.ones(n, dtype=torch.uint8, device=dev); active[::2] = 0
  obs = torch.full((n, 84, 84, 3), -1.0, device=dev)
  pc = torch.full((n, 20, 20), -1.0, device=dev)
  rec = torch.full((n,), -1, dtype=torch.int64, device=dev)
  a = torch.full((n,), 3, dtype=torch.int32, device=dev)
  r, t = K.maze_step(st, a, obs=obs, pc=pc, frame_rec=rec, active=active, auto_reset=False)
  r, t = r.cpu().numpy(), t.cpu().numpy()
  assert (r[1::2] == 1).all() and (t[1::2] == 1).all() and (r[::2] == 0).all() and (t[::2] == 0).all()
  pos = st.pos.cpu().numpy()
  assert (pos[1::2] == [6, 0]).all() and (pos[::2] == [5, 0]).all()     # no reset; skipped envs untouched
  assert (obs[::2] == -1).all() and (pc[::2] == -1).all() and (rec[::2] == 0).all()
  assert np.array_equal(obs[1].cpu().numpy(), O.maze_render(6, 0, np.float32))
  assert (st.last_action.cpu().numpy()[1::2] == 3).all() and (st.last_reward.cpu().numpy()[1::2] == 1).all()


def test_render_and_pc_pairs(K):
  dev = "cuda:0"
  cells = [(x, y) for y in range(7) for x in range(7) if not O.WALLS[y, x]]
  pos = torch.tensor(cells, dtype=torch.int32, device=dev)
  for dt in (torch.float32, torch.uint8):
    img = K.maze_render(pos, dtype=dt).cpu().numpy()
    for k, (x, y) in enumerate(cells):
      want = O.maze_render(x, y, np.float32)
      assert np.array_equal(img[k], want * 255 if dt == torch.uint8 else want)
  p1 = pos.roll(1, 0).contiguous()
  got = K.maze_pixel_change(pos, p1).cpu().numpy()
  for k in range(len(cells)):
    x0, y0 = cells[k]; x1, y1 = cells[k - 1]
    lit = O.pixel_change(O.maze_render(x1, y1), O.maze_render(x0, y0)).astype(np.float32)
    assert np.array_equal(got[k], lit)


def test_pixel_change_of_every_pair_of_free_cells_is_bit_exact(K):
  """Every value the closed form can take (k / 48 for the overlap sums k of any two cells, adjacent or not), computed on the
  device without an IEEE division (one FMA-corrected Newton step), against the oracle's float32 division: bit for bit."""
  dev = "cuda:0"
  cells = [(x, y) for y in range(7) for x in range(7) if not O.WALLS[y, x]]
  a = torch.tensor([c for c in cells for _ in cells], dtype=torch.int32, device=dev)
  b = torch.tensor([c for _ in cells for c in cells], dtype=torch.int32, device=dev)
  got = K.maze_pixel_change(a, b).cpu().numpy()
  an, bn = a.cpu().numpy(), b.cpu().numpy()
  for k in range(len(an)):
    want = O.maze_pixel_change_closed_form(int(an[k, 0]), int(an[k, 1]), int(bn[k, 0]), int(bn[k, 1]))
    assert np.array_equal(got[k], want), (an[k], bn[k])


def test_bad_arguments_fail_loudly(K):
  from unreal_b200 import _lib
  st = K.MazeState(4, "cuda:0")
  with pytest.raises(_lib.UnrealError):
    K.maze_step(st, torch.zeros(4, dtype=torch.int64, device="cuda:0"))       # wrong dtype
  with pytest.raises(_lib.UnrealError):
    K.maze_step(st, torch.zeros(4, dtype=torch.int32))                         # host tensor
  rc = _lib.lib.unreal_maze_step(None, None, None, None, None, None, None, None, 0, None, None, 4, 0, None)
  assert rc == -1 and "non-null" in _lib.last_error()


def test_s2d_render_equals_space_to_depth_of_the_f32_frame():
  """obs_dtype bf16: K1 renders straight into conv1's plane layout; it must equal
  unreal_s2d_frames of the f32 frame, for render and for step (with auto-reset)."""
  from unreal_b200 import kernels as K
  dev = torch.device("cuda", 0)
  cells = torch.tensor([(x, y) for y in range(7) for x in range(7)], dtype=torch.int32, device=dev)
  got = K.maze_render(cells, dtype=torch.bfloat16)
  assert tuple(got.shape) == (49, 6, 441, 8)
  want = K.s2d_frames(K.maze_render(cells, dtype=torch.float32))
  assert torch.equal(got, want)
  n = 300
  g = torch.Generator(device=dev).manual_seed(0)
  sa, sb = K.MazeState(n, dev), K.MazeState(n, dev)
  oa = torch.empty(n, 84, 84, 3, device=dev); ob = torch.empty(n, 6, 441, 8, dtype=torch.bfloat16, device=dev)
  pa = torch.empty(n, 20, 20, device=dev); pb = torch.empty(n, 20, 20, device=dev)
  for _ in range(60):
    act = torch.randint(0, 4, (n,), device=dev, dtype=torch.int32, generator=g)
    ra, ta = K.maze_step(sa, act, obs=oa, pc=pa, auto_reset=True)
    rb, tb = K.maze_step(sb, act, obs=ob, pc=pb, auto_reset=True)
    assert torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(pa, pb) and torch.equal(sa.pos, sb.pos)
    assert torch.equal(ob, K.s2d_frames(oa))


def test_observation_modes_agree_through_resets_and_masks():
  """K1 with f32 frames, x'' planes (bf16) and cell observations (int32, frames implicit) steps the same mazes:
  identical positions / rewards / terminals / records / pixel change, and every active env's observation is the
  render of its (post-auto-reset) cell -- over enough random steps to see hundreds of episode ends."""
  import torch
  from unreal_b200 import kernels as K
  from unreal_b200.environment.maze_environment import BatchedMazeEnvironment
  n = 2048
  envs = {m: BatchedMazeEnvironment(n, "cuda:0", obs_dtype=dt, auto_reset=True)
          for m, dt in (("f32", torch.float32), ("s2d", torch.bfloat16), ("cells", torch.int32))}
  g = torch.Generator(device="cuda").manual_seed(0)
  terms = 0
  for step in range(600):
    act = torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.int32)
    active = (torch.rand(n, device="cuda", generator=g) < 0.9).to(torch.uint8)
    outs = {m: e.process(act, active=active) for m, e in envs.items()}
    ref = envs["f32"]
    pos = ref.state.pos
    am = active.bool()
    for m, e in envs.items():
      assert torch.equal(e.state.pos, pos) and torch.equal(e.frame_rec, ref.frame_rec), (step, m)
      assert torch.equal(outs[m][1], outs["f32"][1]) and torch.equal(outs[m][2], outs["f32"][2]), (step, m)
      assert torch.equal(e.state.last_action, ref.state.last_action) and torch.equal(e.state.last_reward, ref.state.last_reward)
      assert torch.equal(outs[m][3][am], outs["f32"][3][am]), (step, m, "pixel change")
    if step % 20 == 0:
      assert torch.equal(outs["f32"][0]['image'][am], K.maze_render(pos, dtype=torch.float32)[am]), step
      assert torch.equal(outs["s2d"][0]['image'][am], K.maze_render(pos, dtype=torch.bfloat16)[am]), step
    assert torch.equal(outs["cells"][0]['image'][am], pos[am]), step
    terms += int((outs["f32"][2] & active).sum())
  assert terms > 100


@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8], ids=["f32", "u8"])
@pytest.mark.parametrize("auto_reset", [False, True], ids=["no-reset", "auto-reset"])
@pytest.mark.parametrize("wvariant", [-1, 0, 2, 3], ids=["default", "cta-per-item", "warp-per-item", "warp-per-env"])
def test_window_kernel_equals_step_by_step(K, dtype, auto_reset, wvariant):
  """unreal_maze_window (T process() calls per env in one launch: per-item re-simulation for f32, one warp per env for
  u8) against T unreal_maze_step calls, bit for bit on every output and on the carried state -- env counts off the
  CTA / warp granularity, window lengths 1 .. 32 (the maximum), frames / maps optional, two windows in a row."""
  dev = "cuda:0"
  rs = np.random.RandomState(7)
  for n, t in ((1, 1), (5, 7), (130, 20), (67, 32)):
    a_st, b_st = K.MazeState(n, dev), K.MazeState(n, dev)
    for rep in range(2):
      acts = torch.from_numpy(rs.randint(0, 4, size=(t, n)).astype(np.int32)).to(dev)
      if rep == 1 and n > 3:
        acts[:, 0] = 9                      # an out-of-range action is a no-op move with reward 0 (maze_environment.py:99-108)
      want = dict(obs=torch.empty(t, n, 84, 84, 3, dtype=dtype, device=dev), pc=torch.empty(t, n, 20, 20, device=dev),
                  reward=torch.empty(t, n, device=dev), terminal=torch.empty(t, n, dtype=torch.uint8, device=dev),
                  rec=torch.empty(t, n, dtype=torch.int64, device=dev))
      for i in range(t):
        K.maze_step(a_st, acts[i], obs=want["obs"][i], pc=want["pc"][i], reward=want["reward"][i],
                    terminal=want["terminal"][i], frame_rec=want["rec"][i], auto_reset=auto_reset)
      got = {k: torch.full_like(v, 3) for k, v in want.items()}
      from unreal_b200 import _lib as L
      L.set_tunable("maze_render_variant", wvariant)         # all three window kernels, whatever the default for the dtype is
      try:
        K.maze_window(b_st, acts, obs=got["obs"], pc=got["pc"], reward=got["reward"], terminal=got["terminal"],
                      frame_rec=got["rec"], auto_reset=auto_reset)
      finally:
        L.set_tunable("maze_render_variant", -1)
      for k in want:
        assert torch.equal(got[k], want[k]), (n, t, rep, k)
      for f in ("pos", "last_action", "last_reward"):
        assert torch.equal(getattr(a_st, f), getattr(b_st, f)), (n, t, rep, f)
  # outputs are optional: no frames / no maps, state still advances identically
  a_st, b_st = K.MazeState(33, dev), K.MazeState(33, dev)
  acts = torch.from_numpy(rs.randint(0, 4, size=(20, 33)).astype(np.int32)).to(dev)
  r1, t1 = K.maze_window(a_st, acts, auto_reset=auto_reset)
  obs = torch.empty(20, 33, 84, 84, 3, dtype=dtype, device=dev)
  r2, t2 = K.maze_window(b_st, acts, obs=obs, auto_reset=auto_reset)
  assert torch.equal(r1, r2) and torch.equal(t1, t2) and torch.equal(a_st.pos, b_st.pos)
  from unreal_b200 import _lib
  with pytest.raises(_lib.UnrealError):
    K.maze_window(K.MazeState(4, dev), torch.zeros(33, 4, dtype=torch.int32, device=dev))
