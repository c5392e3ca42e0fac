"""Trainer: drop-in for train/trainer.py of the reference, batched over num_envs mazes.

Constructor and method signatures are the reference's (trainer.py:31-59, :132-148, :176, :218,
:339, :383, :415, :438).  One Trainer drives `num_envs` environments in lock step, each of
them behaving exactly like one reference Trainer thread with its own RandomState stream:

  * rollouts stop at a terminal step and the env is reset (trainer.py:279-296) -- the env
    simply goes inactive for the rest of the window, so targets equal the reference's;
  * the per-env draw order is the reference's: T x choice (trainer.py:147-148) ->
    randint for the PC sequence -> randint for the VR sequence -> randint(2), randint(len) for
    RP (experience.py:103, :125, :137-141);
  * with num_envs == 1 the caller's RandomState object itself is advanced (its state is moved
    to the device and back around each call), so sharing it between objects keeps working.

All tensors stay on the device.  The network is any object with the batched call surface of
UnrealModel's run_* helpers (model/model.py:630-728):
    run_base_policy_and_value(sess, state, last_action_reward, active) -> pi [N,A] f32, v [N] f32, _
    run_base_value(sess, state, last_action_reward) -> v [N]
    run_pc_q_max(sess, state, last_action_reward)  -> q [N,20,20]
    run_vr_value(sess, state, last_action_reward)  -> v [N]
    reset_state(mask)          base_lstm_state_out
and optionally `update(feed, learning_rate) -> dict(losses..., grad_norm)`.

Summary writers / tf.summary plumbing of the reference (trainer.py:151-169, :577-632) is
observability only and is not reproduced; the arguments are accepted and ignored.
"""
import time
from collections import deque

import numpy as np
import torch

from .. import _lib
from .. import kernels as K
from ..environment.environment import Environment
from .experience import BatchedExperience

LOG_INTERVAL = 1000
PERFORMANCE_LOG_INTERVAL = 2000
LOSS_AND_EVAL_LOG_INTERVAL = 5000


class Trainer(object):
  def __new__(cls, *args, **kwargs):
    # env types without a closed form (synthetic / indoor-shaped frames, host simulators) are driven by
    # the framed-ring subclass; `Trainer(...)` stays the one constructor, as in the reference
    env_type = args[5] if len(args) > 5 else kwargs.get('env_type')
    if cls is Trainer and env_type != 'maze':
      from .frame_trainer import FrameTrainer
      return object.__new__(FrameTrainer)
    return object.__new__(cls)

  def _make_experience(self, streams):
    if self.env_type != 'maze':
      raise _lib.UnrealError("the compact-record Trainer drives env_type 'maze'")
    return BatchedExperience(self.num_envs, self.experience_history_size, None, self.device, streams=streams)

  def __init__(self, thread_index, global_network, initial_learning_rate, learning_rate_input, grad_applier,
               env_type, env_name, use_lstm, use_pixel_change, use_value_replay, use_reward_prediction,
               pixel_change_lambda, entropy_beta, local_t_max, n_step_TD, gamma, gamma_pc,
               experience_history_size, max_global_time_step, device, segnet_param_dict, image_shape,
               is_training, n_classes, random_state, termination_time, segnet_lambda, dropout,
               num_envs=1, seeds=None, verbose=False, use_graphs=False, obs_s2d=False, env_args=None, obs_cells=False):
    _lib.require_device()
    self.thread_index = thread_index
    self.learning_rate_input = learning_rate_input
    self.env_type = env_type
    self.env_name = env_name
    self.use_lstm = use_lstm
    self.use_pixel_change = use_pixel_change
    self.use_value_replay = use_value_replay
    self.use_reward_prediction = use_reward_prediction
    self.pixel_change_lambda = pixel_change_lambda
    self.entropy_beta = entropy_beta
    self.local_t_max = local_t_max
    self.n_step_TD = n_step_TD
    self.gamma = gamma
    self.gamma_pc = gamma_pc
    self.experience_history_size = experience_history_size
    self.max_global_time_step = max_global_time_step
    self.action_size = Environment.get_action_size(env_type, env_name)
    self.objective_size = Environment.get_objective_size(env_type, env_name)
    self.segnet_param_dict = segnet_param_dict or {}
    self.segnet_mode = self.segnet_param_dict.get("segnet_mode", 0)
    if self.segnet_mode:
      raise _lib.UnrealError("segnet_mode != 0 (ErfNet encoder/decoder) is outside the B200 hot path")
    self.is_training = is_training
    self.n_classes = n_classes
    self.segnet_lambda = segnet_lambda
    self.random_state = random_state
    self.termination_time = termination_time
    self.dropout = dropout
    self.verbose = verbose
    self.env_args = dict(env_args or {})     # extra create_environment arguments (framed env types)
    self.num_envs = int(num_envs)
    self.device = torch.device(device if str(device).startswith("cuda") else "cuda:0")
    if seeds is None and self.num_envs > 1:
      seeds = random_state.randint(0, 2 ** 31 - 1, size=self.num_envs)
    self._seeds = seeds

    # the reference builds a thread-local copy of the network and syncs it from the global one
    # every iteration (trainer.py:93-116, :457); the batched learner is synchronous, so the
    # global network is used directly.
    self.local_network = global_network
    self.grad_applier = grad_applier
    n, d = self.num_envs, self.device
    with torch.cuda.device(d):
      streams = K.MtStreams([0] * n if seeds is None else seeds, d)
    self.image_shape = tuple(image_shape) if image_shape is not None else (84, 84)
    self.experience = self._make_experience(streams)
    self.streams = streams
    self.local_t = 0
    self.initial_learning_rate = initial_learning_rate
    self.episode_reward = torch.zeros(n, dtype=torch.float32, device=d)
    self.episode_stats = torch.zeros(2, dtype=torch.float64, device=d)   # [episodes finished, sum of scores]
    self.prev_local_t = -1
    self.prev_local_t_loss = 0
    self.sr_size = 50
    self.success_rates = deque(maxlen=self.sr_size)
    self.environment = None
    self.last_feed = None
    self._ring_full = False
    self.use_graphs = bool(use_graphs)      # capture the rollout + sampling phase into one CUDA graph
    # obs_s2d: K1 renders frames straight into conv1's space-to-depth bf16 plane layout (42 KB per
    # frame, no f32 frame and no separate s2d pass); needs a network with the fused conv1 kernel
    self.obs_dtype = torch.bfloat16 if obs_s2d else torch.float32
    # obs_cells: observations are the agent CELLS (int32 [N,2]); frames stay implicit -- the network's conv1
    # kernels render their input tiles in shared memory (maze only; needs UnrealModel's fused encoder)
    if obs_cells:
      self.obs_dtype = torch.int32
    self._graph = None
    self.graph_update = True                # with use_graphs: capture the learner update as well (single process)
    self._ugraph = None
    self._ugraph_b = None
    self._ugraph_out = None
    import os
    self.nccl_in_graph = os.environ.get("UNREAL_NCCL_IN_GRAPH", "0") == "1"   # capture the gradient exchange too
    self._lr_dev = None

  # -- RandomState hand-over for the single-env drop-in case ---------------------------------
  def _rng_in(self):
    if self.num_envs == 1 and self._seeds is None:
      self.streams.load_numpy_state(self.random_state)

  def _rng_out(self):
    if self.num_envs == 1 and self._seeds is None:
      self.streams.store_numpy_state(self.random_state)

  # -- reference surface ---------------------------------------------------------------------
  def prepare(self, termination_time=50.0, termination_dist_value=-10.0):
    """trainer.py:132-135."""
    self.environment = Environment.create_environment(
        self.env_type, self.env_name, self.termination_time,
        env_args={'num_envs': self.num_envs, 'device': self.device, 'auto_reset': True, 'obs_dtype': self.obs_dtype},
        thread_index=self.thread_index)
    n, d, T = self.num_envs, self.device, self.n_step_TD
    A = self.action_size
    # rollout buffers: obs has T+1 slots (state before each action + the bootstrap state)
    self._obs = torch.zeros(T + 1, n, *K.obs_shape(self.obs_dtype), dtype=self.obs_dtype, device=d)
    self._pos = torch.zeros(T + 1, n, 2, dtype=torch.int32, device=d)
    self._lar = torch.zeros(T, n, A + 1, dtype=torch.float32, device=d)
    self._act = torch.zeros(T, n, dtype=torch.int32, device=d)
    self._val = torch.zeros(T, n, dtype=torch.float32, device=d)
    self._rew = torch.zeros(T, n, dtype=torch.float32, device=d)
    self._term = torch.zeros(T, n, dtype=torch.uint8, device=d)
    self._active = torch.zeros(T, n, dtype=torch.uint8, device=d)

  def stop(self):
    if self.environment is not None:
      self.environment.stop()

  def _anneal_learning_rate(self, global_time_step):
    """trainer.py:140-144."""
    learning_rate = self.initial_learning_rate * (self.max_global_time_step - global_time_step) / \
        self.max_global_time_step
    if learning_rate < 0.0:
      learning_rate = 0.0
    return learning_rate

  def choose_action(self, pi_values, active=None, out=None):
    """trainer.py:147-148, per env on its own device RandomState stream.  A 1-D numpy/torch
    vector is accepted for the scalar case and returns a python int.  `out` (int32 [N]): rows of active envs are
    written there, the others left as they are."""
    scalar = not isinstance(pi_values, torch.Tensor) or pi_values.dim() == 1
    pi = torch.as_tensor(np.asarray(pi_values, dtype=np.float32) if scalar else pi_values)
    pi = pi.to(self.device, torch.float32).reshape(-1, pi.shape[-1]).contiguous()
    a = self.streams.choose_action(pi, active, out)
    return int(a[0]) if scalar else a

  def set_start_time(self, start_time):
    self.start_time = start_time

  def _last_action_reward(self, last_action, last_reward):
    """ExperienceFrame.concat_action_and_reward (experience.py:34-46), batched: [N, A+1]."""
    n = last_action.shape[0]
    out = torch.zeros(n, self.action_size + 1, dtype=torch.float32, device=self.device)
    out.scatter_(1, last_action.to(torch.int64).clamp_(0, self.action_size - 1).unsqueeze(1), 1.0)
    out[:, -1] = last_reward
    return out

  def _fill_mask(self):
    """Warm-up mask: an env whose ring is full is one reference worker that has left `_fill_experience`
    (trainer.py:446-448) -- it is frozen (no step, no RNG draw, no frame) until the slowest ring is full."""
    if getattr(self, "_fill_active", None) is None:
      self._fill_active = torch.ones(self.num_envs, dtype=torch.uint8, device=self.device)
    return self._fill_active

  def _fill_done(self, env):
    """trainer.py:203-205, without a host round trip: reset the envs whose ring became full in THIS step
    (once, like the reference) and take them out of the warm-up."""
    full = self.experience.ring.state()["full"]
    newly = full & self._fill_active
    env.reset(newly)
    self._fill_active = self._fill_active & (1 - full)

  def _fill_experience(self, sess):
    """trainer.py:176-205: one policy forward, one env step, one add_frame, for every env still warming up."""
    env = self.environment
    act = self._fill_mask()
    lar = self._last_action_reward(env.last_action, env.last_reward)
    pi, _, _ = self.local_network.run_base_policy_and_value(sess, env.last_state, lar, act)
    action = self.choose_action(pi, act)
    env.process(action, active=act)           # auto-reset covers `if terminal: reset()` (:201-202)
    self.experience.add_frames(env.frame_rec)  # frozen envs carry an invalid record: nothing is added
    self._fill_done(env)

  def _set_pending(self, lengths, stats0):
    """What process() returns, as ONE small device tensor (a single read-back per iteration):
    [env steps taken in this window = sum of the rollout lengths (trainer.py:277 `local_t += 1` per step of every
    worker; main.py:125 adds each worker's diff to global_t), episodes finished in it, sum of their scores]. """
    self._pending = torch.cat((lengths.sum().to(torch.float64).view(1), self.episode_stats - stats0))

  def _print_log(self, global_t):
    if (self.thread_index == 0) and (self.local_t - self.prev_local_t >= PERFORMANCE_LOG_INTERVAL):
      self.prev_local_t += PERFORMANCE_LOG_INTERVAL
      elapsed_time = time.time() - self.start_time
      steps_per_sec = global_t / elapsed_time
      print("### Performance : {} STEPS in {:.0f} sec. {:.0f} STEPS/sec. {:.2f}M STEPS/hour".format(
          global_t, elapsed_time, steps_per_sec, steps_per_sec * 3600 / 1000000.))

  # -- [Base A3C]  trainer.py:218-336 ----------------------------------------------------------
  def _process_base(self, sess, global_t, summary_writer, summary_op_dict, summary_dict):
    env, net = self.environment, self.local_network
    n, T, d = self.num_envs, self.n_step_TD, self.device
    # the model updates its state buffers in place: keep a copy of the state the rollout started from
    state = net.base_lstm_state_out if self.use_lstm else None
    start_lstm_state = None if state is None else tuple(x.clone() if isinstance(x, torch.Tensor) else x for x in state)
    active = torch.ones(n, dtype=torch.uint8, device=d)
    ended = torch.zeros(n, dtype=torch.uint8, device=d)
    stats0 = self.episode_stats.clone()       # [episodes finished, sum of their scores] before this window
    self._obs[0].copy_(env.last_state['image'])
    self._pos[0].copy_(env.state.pos)
    self._active.zero_(); self._rew.zero_(); self._term.zero_(); self._act.zero_()
    import inspect
    if getattr(self, "_net_v_out", None) is None:     # networks that can write the values straight into the history row
      self._net_v_out = "v_out" in inspect.signature(net.run_base_policy_and_value).parameters
    last_rec = torch.zeros(n, dtype=torch.int64, device=d)
    # networks with the persistent [N,256] LSTM state buffers (UnrealModel) get them zeroed by the bookkeeping kernel
    lstm_c, lstm_h = getattr(net, "_lstm_c", None), getattr(net, "_lstm_h", None)
    fused_state = isinstance(lstm_c, torch.Tensor) and isinstance(lstm_h, torch.Tensor)
    for t in range(T):
      lar = K.rollout_lar(env.last_action, env.last_reward, self.action_size, out=self._lar[t])   # straight into the feed
      if self._net_v_out:
        pi, v, _ = net.run_base_policy_and_value(sess, {'image': self._obs[t]}, lar, active, v_out=self._val[t])
      else:
        pi, v, _ = net.run_base_policy_and_value(sess, {'image': self._obs[t]}, lar, active)
        self._val[t].copy_(v)
      # the history row is the kernel's output: finished rollouts draw nothing and keep the zero the row was reset to
      action = self.choose_action(pi, active, out=self._act[t])
      self._active[t].copy_(active)
      # the new frame lands in obs[t+1]; envs whose rollout already ended are skipped by the kernel
      # (no pixel-change map: replayed frames re-derive theirs from the record's two cells, _process_pc)
      env.process(action, active=active, out_obs=self._obs[t + 1], out_pc=False,
                  out_reward=self._rew[t], out_terminal=self._term[t])
      self._pos[t + 1].copy_(env.state.pos)
      self.experience.add_frames(env.frame_rec)
      if not fused_state:
        net.reset_state(self._term[t] & active)      # :293
      # everything else of :265-296 in one launch, terminal handling as masked arithmetic (no host round trip, the
      # CPU keeps enqueueing ahead of the GPU; an all-terminated batch just idles through the steps): last record,
      # episode reward / statistics (the reference prints every finished episode's score, :283-291), `ended`,
      # LSTM state reset, `active`
      K.rollout_post(self._rew[t], self._term[t], env.frame_rec, active, ended, last_rec, self.episode_reward,
                     lstm_c if fused_state else None, lstm_h if fused_state else None, self.episode_stats)
    lengths = self._active.sum(0).to(torch.int32)
    self._set_pending(lengths, stats0)         # read back once, after the update has been enqueued
    # bootstrap: V(new_state) with frame.get_action_reward for envs that did not end (:298-300)
    rec = K.frame_unpack(last_rec, fields=("action", "reward"))
    boot_lar = self._last_action_reward(rec["action"], rec["reward"])
    boot_obs = self._obs[1:].gather(
        0, (lengths.to(torch.int64) - 1).clamp_(min=0).view(1, n, *([1] * (self._obs.dim() - 2))).expand(
            1, n, *self._obs.shape[2:]))[0]
    # every env's current frame (the reset frame for ended envs) goes back into the env's own
    # persistent frame buffer, which is where the next rollout starts reading (graph-replay safe)
    env._obs.copy_(boot_obs)
    env.last_state = {'image': env._obs}
    boot = net.run_base_value(sess, {'image': boot_obs}, boot_lar)
    boot = torch.where(ended.bool(), torch.zeros_like(boot), boot).contiguous()
    R, adv = K.nstep_returns(self._rew, self._val, self._term, boot, self.gamma)
    batch_a = torch.nn.functional.one_hot(self._act.to(torch.int64), self.action_size).to(torch.float32)
    return dict(si=self._obs[:T], pos=self._pos[:T], last_action_rewards=self._lar, a=batch_a, adv=adv, R=R,
                start_lstm_state=start_lstm_state, length=lengths, active=self._active, terminal_end=ended)

  # -- replayed sequences, shared by PC and VR -----------------------------------------------
  def _sample_sequence(self):
    L = self.local_t_max + 1
    start, length, f = self.experience.sample_sequence(L)
    n_batch = (length - 1).clamp_(min=0).to(torch.int32)    # the last frame is only the bootstrap state
    idx_last = (length.to(torch.int64) - 1).clamp_(min=0)
    take = lambda x: x.gather(1, idx_last.view(-1, *([1] * (x.dim() - 1))).expand(-1, 1, *x.shape[2:]))[:, 0]  # noqa: E731
    boot_pos = take(f["pos0"]).contiguous()
    boot_lar = self._last_action_reward(take(f["last_action"]), take(f["last_reward"]))
    boot_state = {'image': K.maze_render(boot_pos, dtype=self.obs_dtype), 'pos': boot_pos}
    lar = self._last_action_reward(f["last_action"].reshape(-1), f["last_reward"].reshape(-1))
    lar = lar.view(self.num_envs, L, -1)
    return start, length, n_batch, f, boot_state, boot_lar, lar

  # -- [Pixel change]  trainer.py:339-380 ------------------------------------------------------
  def _process_pc(self, sess):
    n, L = self.num_envs, self.local_t_max + 1
    start, length, n_batch, f, boot_state, boot_lar, lar = self._sample_sequence()
    # the `frames[1].terminal` guard (:352) can never fire (only the last sampled frame can be
    # terminal), so the bootstrap is always taken -- as in the reference
    pc_boot = self.local_network.run_pc_q_max(sess, boot_state, boot_lar).contiguous()
    pos0 = f["pos0"][:, :L - 1].transpose(0, 1).contiguous()      # [L-1, N, 2] time-major
    pos1 = f["pos1"][:, :L - 1].transpose(0, 1).contiguous()
    # maze_pixel_change + pc_targets in one pass: the [L-1,N,20,20] maps themselves are never written
    pc_R = K.maze_pc_targets(pos0, pos1, n_batch, pc_boot, self.gamma_pc)
    a = torch.nn.functional.one_hot(f["action"][:, :L - 1].to(torch.int64), self.action_size).to(torch.float32)
    return dict(pos=f["pos0"][:, :L - 1], last_action_reward=lar[:, :L - 1], a=a, R=pc_R.transpose(0, 1),
                length=n_batch, start=start)

  # -- [Value replay]  trainer.py:383-412 ------------------------------------------------------
  def _process_vr(self, sess):
    L = self.local_t_max + 1
    start, length, n_batch, f, boot_state, boot_lar, lar = self._sample_sequence()
    vr_boot = self.local_network.run_vr_value(sess, boot_state, boot_lar).contiguous()
    vr_R = K.sequence_returns(f["reward"].contiguous(), n_batch, vr_boot, self.gamma)
    return dict(pos=f["pos0"][:, :L - 1], last_action_reward=lar[:, :L - 1], R=vr_R[:, :L - 1], length=n_batch,
                start=start)

  # -- [Reward prediction]  trainer.py:415-436 -------------------------------------------------
  def _process_rp(self):
    start, f = self.experience.sample_rp_sequence()
    r = f["reward"][:, 3]
    c = torch.zeros(self.num_envs, 3, dtype=torch.float32, device=self.device)
    zero = r.abs() < 1e-10
    c[:, 0] = zero.float(); c[:, 1] = (~zero & (r > 0)).float(); c[:, 2] = (~zero & (r < 0)).float()
    return dict(pos=f["pos0"][:, :3], c=c, start=start)

  # -- rollout + replay sampling + targets: everything that feeds one update -------------------
  def _data_phase(self, sess):
    feed = {'base': self._process_base(sess, 0, None, None, {'placeholders': {}, 'values': {}})}
    if self.use_pixel_change:
      feed['pc'] = self._process_pc(sess)
    if self.use_value_replay:
      feed['vr'] = self._process_vr(sess)
    if self.use_reward_prediction:
      feed['rp'] = self._process_rp()
    return feed

  def _data_phase_graphed(self, sess):
    """The data phase is ~1000 small launches with no host decision inside (terminal handling is
    masked arithmetic): capture it ONCE into a CUDA graph and replay it.  Every buffer it touches
    is persistent -- env / ring / RNG / LSTM state are updated in place, the filters' bf16 shadows
    are refreshed in place -- so the replay acts on the live state and the returned feed tensors
    keep their addresses."""
    if self._graph is None:
      # first call: run the phase eagerly (this IS this iteration's data, and it performs every lazy
      # initialisation), then record the same code into a graph without executing it
      feed = self._data_phase(sess)
      pending = self._pending
      torch.cuda.synchronize(self.device)
      g = torch.cuda.CUDAGraph()
      with _lib.graph_capture(g):
        self._graph_feed = self._data_phase(sess)
      self._graph_pending = self._pending
      self._pending = pending
      self._graph = g
      return feed
    self._pending = self._graph_pending
    self._graph.replay()
    return dict(self._graph_feed)

  def _update(self, feed, learning_rate):
    """The learner step (`sess.run(apply_gradients)`, trainer.py:543-559).  With use_graphs the whole update -- feed
    conversion, forward, backward, clip + RMSProp, shadow refresh, some hundreds of launches -- is captured ONCE into
    CUDA graphs over the data phase's static feed tensors and replayed; the annealed learning rate travels through a
    device scalar that K6 reads when it runs.  Single process: one graph.  Under NCCL: graph A (forward +
    backward -> flat gradient), then the exchange step launched eagerly -- reduce-scatter, sum of squares of the shard,
    8-byte all-reduce, fused clip + RMSProp on the shard, all-gather: five launches -- then graph B (bf16 / tap-major
    shadow refresh); `nccl_in_graph` captures the collectives as well (one graph, like the single-process case)."""
    net, ap = self.local_network, self.grad_applier
    distributed = ap is not None and getattr(ap, "_world", None) is not None and ap._world()[0] > 1
    if not (self.use_graphs and self.graph_update) or self._graph is None:
      return net.update(feed, learning_rate, ap)
    if self._lr_dev is None:
      self._lr_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
    self._lr_dev.fill_(float(learning_rate))
    split = distributed and not self.nccl_in_graph
    if self._ugraph is None:
      if any(feed[k] is not self._graph_feed[k] for k in self._graph_feed):
        return net.update(feed, learning_rate, ap)     # the capturing iteration of the data phase: eager feed
      static_feed = dict(self._graph_feed)
      out = net.update(static_feed, self._lr_dev, ap)  # this iteration's update, eagerly (also warms every lazy path)
      torch.cuda.synchronize(self.device)
      g = torch.cuda.CUDAGraph()
      if not split:
        with _lib.graph_capture(g):                    # recorded, not executed
          self._ugraph_out = net.update(static_feed, self._lr_dev, ap)
        self._ugraph = g
        return out
      with _lib.graph_capture(g):
        total, parts, grad = net.update_gradient(static_feed)
      g2 = torch.cuda.CUDAGraph()
      with _lib.graph_capture(g2):
        net.refresh_shadow()
      self._ugraph, self._ugraph_b = g, g2
      self._ugraph_out = dict(parts)
      self._ugraph_out["total"] = total
      self._ugraph_grad = grad
      return out
    self._ugraph.replay()
    if split:
      self._ugraph_out["grad_norm"] = ap.apply_flat_to(net.flat, self._ugraph_grad, self._lr_dev)
      self._ugraph_b.replay()
    return self._ugraph_out

  # -- one iteration  trainer.py:438-636 -------------------------------------------------------
  def process(self, sess=None, global_t=0, summary_writer=None, summary_op_dict=None, score_input=None,
              sr_input=None, eval_input=None, entropy_input=None, term_global_t=None, losses_input=None):
    with torch.cuda.device(self.device):
      self._rng_in()
      try:
        if not self._ring_full:
          self._ring_full = self.experience.is_full()     # sticky: a full ring stays full
        if not self._ring_full:
          self._fill_experience(sess)
          return 0, None
        start_local_t = self.local_t
        cur_learning_rate = self._anneal_learning_rate(global_t)
        feed = self._data_phase_graphed(sess) if self.use_graphs else self._data_phase(sess)
        feed['learning_rate'] = cur_learning_rate
      finally:
        self._rng_out()
      self.last_feed = feed
      if hasattr(self.local_network, 'update'):
        self.last_losses = self._update(feed, cur_learning_rate)
      steps, finished, score_sum = self._pending.tolist()       # the iteration's only host read-back
      self.local_t += int(steps)
      if hasattr(self, 'start_time'):
        self._print_log(global_t)
      # trainer.py:451, :481: the score of the episode that ended in this rollout, else None.  With num_envs > 1
      # several workers' episodes can end in one window: their MEAN score (each is one reference worker's return).
      episode_score = None
      if finished > 0:
        episode_score = score_sum / finished
        if self.num_envs == 1 and float(episode_score).is_integer():
          episode_score = int(episode_score)                      # maze rewards are ints, like the reference's sum
        if self.verbose:
          print("Trainer {}>>> score={}".format(self.thread_index, episode_score))     # :281
      # :635-636: env steps actually taken (sum over envs -- what main.py:125 accumulates into global_t)
      return self.local_t - start_local_t, episode_score
