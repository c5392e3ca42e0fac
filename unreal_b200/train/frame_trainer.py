"""FrameTrainer: the batched Trainer for env types whose frames have no closed form (SURVEY.md 8f-4:
indoor / lab / gym process() shape; here 'synthetic' frames of the MINOS observation shape, or host
simulators behind environment/frame_environment.py:HostEnvProducer).

Same reference surface and the same per-env semantics as `Trainer` (train/trainer.py:176-436): the
only difference is where a replayed frame comes from.  The maze Trainer stores 8-byte records and
re-renders frames from cells; here the states, pixel-change maps, float rewards and objective vectors
sit in payload rings beside the records (`FramedExperience`) and sampled sequences are gathered from
them (`unreal_replay_gather`), so `_process_pc / _process_vr / _process_rp` hand the network real
frames: feed entries carry `images` u8 [L,N,84,84,3] instead of `pos`.
"""
import torch

from .. import _lib
from .. import kernels as K
from ..environment.environment import Environment
from .experience import FramedExperience
from .trainer import Trainer


class FrameTrainer(Trainer):
  def _make_experience(self, streams):
    h, w = self.image_shape
    return FramedExperience(self.num_envs, self.experience_history_size, None, self.device, streams=streams,
                            frame_shape=(h, w, 3), objective_size=self.objective_size)

  def prepare(self, termination_time=50.0, termination_dist_value=-10.0):
    """trainer.py:132-135."""
    args = dict(self.env_args)
    args.update({'num_envs': self.num_envs, 'device': self.device})
    args.setdefault('height', self.image_shape[0]); args.setdefault('width', self.image_shape[1])
    self.environment = Environment.create_environment(self.env_type, self.env_name, self.termination_time,
                                                      env_args=args, thread_index=self.thread_index)
    n, d, T = self.num_envs, self.device, self.n_step_TD
    A, G = self.action_size, self.objective_size
    h, w = self.image_shape
    self.obs_dtype = torch.uint8
    self._obs = torch.zeros(T + 1, n, h, w, 3, dtype=torch.uint8, device=d)
    self._lar = torch.zeros(T, n, A + 1 + G, dtype=torch.float32, device=d)
    self._act = torch.zeros(T, n, dtype=torch.int32, device=d)
    self._val = torch.zeros(T, n, dtype=torch.float32, device=d)
    self._rew = torch.zeros(T, n, dtype=torch.float32, device=d)
    self._term = torch.zeros(T, n, dtype=torch.uint8, device=d)
    self._active = torch.zeros(T, n, dtype=torch.uint8, device=d)
    self._pc = torch.zeros(T, n, (h - 4) // 4, (w - 4) // 4, dtype=torch.float32, device=d)

  def _last_action_reward(self, last_action, last_reward, objective=None):
    """ExperienceFrame.concat_action_and_reward (experience.py:34-46), batched: [N, A+1+G]."""
    lar = Trainer._last_action_reward(self, last_action, last_reward)
    if self.objective_size:
      lar = torch.cat((lar, objective.to(torch.float32).reshape(lar.shape[0], self.objective_size)), dim=1)
    return lar

  def _fill_experience(self, sess):
    """trainer.py:176-205 for every env still warming up (a full ring = a worker that has left the warm-up)."""
    env = self.environment
    act = self._fill_mask()
    obj = env.objective.clone() if env.objective is not None else None
    last_action, last_reward = env.last_action.clone(), env.last_reward.clone()
    lar = self._last_action_reward(last_action, last_reward, obj)
    prev = env.last_state['image']
    pi, _, _ = self.local_network.run_base_policy_and_value(sess, env.last_state, lar, act)
    action = self.choose_action(pi, act)
    _, reward, _, pc = env.process(action, active=act)        # terminal envs are reset inside (:201-202)
    self.experience.add_frames(env.frame_rec, frame=prev, pixel_change=pc, reward=reward, last_reward=last_reward,
                               objective=obj)
    self._fill_done(env)                                      # :203-205

  # -- [Base A3C]  trainer.py:218-336 ----------------------------------------------------------
  def _process_base(self, sess, global_t, summary_writer, summary_op_dict, summary_dict):
    env, net = self.environment, self.local_network
    n, T, d = self.num_envs, self.n_step_TD, self.device
    state = net.base_lstm_state_out if self.use_lstm else None
    start_lstm_state = None if state is None else tuple(x.clone() if isinstance(x, torch.Tensor) else x for x in state)
    active = torch.ones(n, dtype=torch.uint8, device=d)
    ended = torch.zeros(n, dtype=torch.uint8, device=d)
    stats0 = self.episode_stats.clone()
    self._obs[0].copy_(env.last_state['image'])
    env._cur = self._obs[0]
    self._active.zero_(); self._rew.zero_(); self._term.zero_()
    last_action = torch.zeros(n, dtype=torch.int32, device=d)     # of each env's LAST step (bootstrap input)
    last_reward = torch.zeros(n, dtype=torch.float32, device=d)
    last_obj = env.objective.clone() if env.objective is not None else None
    for t in range(T):
      obj = env.objective.clone() if env.objective is not None else None
      prev_reward = env.last_reward.clone()
      lar = self._last_action_reward(env.last_action, prev_reward, obj)
      pi, v, _ = net.run_base_policy_and_value(sess, {'image': self._obs[t]}, lar, active)
      action = self.choose_action(pi, active)
      self._lar[t].copy_(lar); self._val[t].copy_(v); self._act[t].copy_(action); self._active[t].copy_(active)
      env.process(action, active=active, out_obs=self._obs[t + 1], out_pc=self._pc[t], out_reward=self._rew[t],
                  out_terminal=self._term[t])
      self.experience.add_frames(env.frame_rec, frame=self._obs[t], pixel_change=self._pc[t], reward=self._rew[t],
                                 last_reward=prev_reward, objective=obj)
      am = active.bool()
      last_action = torch.where(am, action, last_action)
      last_reward = torch.where(am, self._rew[t], last_reward)
      if obj is not None:
        last_obj = torch.where(am.unsqueeze(1), obj, last_obj)    # frame.state's objective (:300)
      self.episode_reward += self._rew[t]
      term_now = self._term[t] & active
      ended |= term_now
      tn = term_now.to(torch.float32)
      self.episode_stats[0] += tn.sum()
      self.episode_stats[1] += (self.episode_reward * tn).sum()
      net.reset_state(term_now)                # :293
      self.episode_reward.mul_(1 - tn)
      active = active & (1 - term_now)
    lengths = self._active.sum(0).to(torch.int32)
    self._set_pending(lengths, stats0)
    # bootstrap: V(new_state) with frame.get_action_reward for envs that did not end (:298-300)
    boot_lar = self._last_action_reward(last_action, last_reward, last_obj)
    boot_obs = self._obs[1:].gather(
        0, (lengths.to(torch.int64) - 1).clamp_(min=0).view(1, n, 1, 1, 1).expand(1, n, *self._obs.shape[2:]))[0]
    env.set_current(boot_obs)                  # each env's current frame (the reset frame for ended envs)
    boot = net.run_base_value(sess, {'image': boot_obs}, boot_lar)
    boot = torch.where(ended.bool(), torch.zeros_like(boot), boot).contiguous()
    R, adv = K.nstep_returns(self._rew, self._val, self._term, boot, self.gamma)
    batch_a = torch.nn.functional.one_hot(self._act.to(torch.int64), self.action_size).to(torch.float32)
    return dict(si=self._obs[:T], last_action_rewards=self._lar, a=batch_a, adv=adv, R=R,
                start_lstm_state=start_lstm_state, length=lengths, active=self._active, terminal_end=ended)

  # -- replayed sequences, shared by PC and VR -----------------------------------------------
  def _lar_seq(self, f, L):
    n = self.num_envs
    obj = f["objective"].reshape(n * L, -1) if self.objective_size else None
    return self._last_action_reward(f["last_action"].reshape(-1), f["last_reward"].reshape(-1), obj).view(n, L, -1)

  def _sample_sequence(self):
    L = self.local_t_max + 1
    n = self.num_envs
    ex = self.experience
    start, length, f = ex.sample_sequence(L)
    n_batch = (length - 1).clamp_(min=0).to(torch.int32)    # the last frame is only the bootstrap state
    images = ex.gather_frames(start, length, L)             # [L,N,h,w,3] u8, zero past len
    idx_last = (length.to(torch.int64) - 1).clamp_(min=0)
    lar = self._lar_seq(f, L)
    boot_lar = lar.gather(1, idx_last.view(n, 1, 1).expand(n, 1, lar.shape[2]))[:, 0]
    boot_img = images.gather(0, idx_last.view(1, n, 1, 1, 1).expand(1, n, *images.shape[2:]))[0]
    return start, length, n_batch, f, {'image': boot_img}, boot_lar, lar, images

  # -- [Pixel change]  trainer.py:339-380 ------------------------------------------------------
  def _process_pc(self, sess):
    L = self.local_t_max + 1
    start, length, n_batch, f, boot_state, boot_lar, lar, images = self._sample_sequence()
    pc_boot = self.local_network.run_pc_q_max(sess, boot_state, boot_lar).contiguous()
    pc = self.experience.gather_pixel_change(start, length, L)[:L - 1].contiguous()       # [L-1,N,20,20]
    pc_R = K.pc_targets(pc, None, n_batch, pc_boot, self.gamma_pc)
    a = torch.nn.functional.one_hot(f["action"][:, :L - 1].to(torch.int64), self.action_size).to(torch.float32)
    return dict(images=images[:L - 1], last_action_reward=lar[:, :L - 1], a=a, R=pc_R.transpose(0, 1),
                length=n_batch, start=start)

  # -- [Value replay]  trainer.py:383-412 ------------------------------------------------------
  def _process_vr(self, sess):
    L = self.local_t_max + 1
    start, length, n_batch, f, boot_state, boot_lar, lar, images = self._sample_sequence()
    vr_boot = self.local_network.run_vr_value(sess, boot_state, boot_lar).contiguous()
    vr_R = K.sequence_returns(f["reward"].contiguous(), n_batch, vr_boot, self.gamma)
    return dict(images=images[:L - 1], last_action_reward=lar[:, :L - 1], R=vr_R[:, :L - 1], length=n_batch,
                start=start)

  # -- [Reward prediction]  trainer.py:415-436 -------------------------------------------------
  def _process_rp(self):
    start, f = self.experience.sample_rp_sequence()
    images = self.experience.gather_frames(start, None, 4, time_major=False)            # [N,4,h,w,3]
    r = f["reward"][:, 3]
    c = torch.zeros(self.num_envs, 3, dtype=torch.float32, device=self.device)
    zero = r.abs() < 1e-10
    c[:, 0] = zero.float(); c[:, 1] = (~zero & (r > 0)).float(); c[:, 2] = (~zero & (r < 0)).float()
    return dict(images=images[:, :3], c=c, start=start)
