#!/usr/bin/env python
"""Golden vectors of the WHOLE learner loop, produced by the reference's own classes wired as main.py wires them.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_agent_golden.py [/root/reference]

Everything that runs is the reference's code, unmodified: ``environment.maze_environment.MazeEnvironment``,
``train.experience.Experience``, ``model.model.UnrealModel`` (a global network and the worker's local one),
``train.rmsprop_applier.RMSPropApplier`` and ``train.trainer.Trainer`` -- ``_fill_experience`` until the replay buffer is full,
then ``Trainer.process`` (sync_from, _process_base / _pc / _vr / _rp, the feed dict, apply_gradients) several times.
``tensorflow`` resolves to ``tests/golden/tf1_shim`` (TensorFlow cannot be installed here: its ops on this path are restated
from their published definitions, float64).  ``Trainer.__init__`` itself is not run (it needs TF summaries, options and the
environment registry); the worker object is assembled from its attributes with the reference's methods bound to it, and
the maze needs the three-line ``MazeShim`` of make_golden.py (the fork's Trainer passes ``flag=`` and reads
``_last_full_state``, which only IndoorEnvironment has).

Written: ``agent_reference_golden.npz`` -- after the fill and after every ``process()`` call: the worker's step counter, the
returned (steps, score), the agent's cell, the next draw of the worker's RandomState (the stream position), 64 sampled
entries + the sum of every GLOBAL variable, and the local network's carried LSTM state.
tests/test_model_oracle.py::test_oracle_agent_loop_matches_the_references_trainer replays the same loop with the oracle.
"""
import contextlib
import io
import os
import sys
import types
from collections import deque

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "tf1_shim"))

import tensorflow as tf  # noqa: E402  (the shim)
import torch  # noqa: E402
from environment.maze_environment import MazeEnvironment  # noqa: E402  (reference)
from model.model import UnrealModel  # noqa: E402  (reference)
from train.experience import Experience  # noqa: E402  (reference)
from train.rmsprop_applier import RMSPropApplier  # noqa: E402  (reference)
from train.trainer import Trainer  # noqa: E402  (reference)
from oracle import model_oracle as MO  # noqa: E402

A, G, SEED, NET_SEED = 4, 0, 3, 11
H, N_PROCESS, INITIAL_LR, MAX_T = 60, 4, 7e-4, 10 ** 6


class MazeShim(MazeEnvironment):
  def process(self, action, flag=0):
    image, r, t, pc = MazeEnvironment.process(self, action)
    self._last_full_state = {'success': bool(t)}
    return {'image': image}, r, t, pc


def main():
  specs = MO.variable_specs(A, G)
  params = MO.init_params(A, G, seed=NET_SEED)
  mk = lambda idx: UnrealModel(A, G, idx, True, True, True, True, 0.05, 0.001, "/cpu:0", {'segnet_mode': 0}, [84, 84],  # noqa: E731
                               True, 0, 0.0, 0.0)
  with contextlib.redirect_stdout(io.StringIO()):
    glob = mk(-1)
    for v, (name, _, _) in zip(glob.get_vars(), specs):
      v.value = params[name].to(torch.float64).clone()
    lr_in = tf.placeholder("float")
    applier = RMSPropApplier(learning_rate=lr_in, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0, device="/cpu:0")
    local = mk(1)
    local.prepare_loss()
    rs = np.random.RandomState(SEED)
    me = types.SimpleNamespace(
        thread_index=1, learning_rate_input=lr_in, use_lstm=True, use_pixel_change=True, use_value_replay=True,
        use_reward_prediction=True, local_t_max=20, n_step_TD=20, gamma=0.99, gamma_pc=0.9, experience_history_size=H,
        max_global_time_step=MAX_T, action_size=4, objective_size=0, segnet_param_dict={'segnet_mode': 0}, segnet_mode=0,
        is_training=True, n_classes=0, segnet_lambda=0.0, random_state=rs, local_network=local,
        experience=Experience(H, random_state=rs), local_t=0, initial_learning_rate=INITIAL_LR, episode_reward=0,
        prev_local_t=-1, prev_local_t_loss=0, sr_size=50, success_rates=deque(maxlen=50), environment=MazeShim(),
        start_time=0.0)
    me.apply_gradients = applier.minimize_local(local.total_loss, glob.get_vars(), local.get_vars(), 1)   # trainer.py:120-122
    me.sync = local.sync_from(glob)                                                                       # :124
  for name in ("choose_action", "_anneal_learning_rate", "_fill_experience", "_process_base", "_process_pc", "_process_vr",
               "_process_rp", "concat_action_and_reward"):
    if hasattr(Trainer, name):
      setattr(me, name, (lambda fn: (lambda *a, **k: fn(me, *a, **k)))(getattr(Trainer, name)))
  me._print_log = lambda *a, **k: None                     # wall-clock logging
  me._record_all = me._record_one = lambda *a, **k: None    # TF summaries
  sess = tf.Session()
  sess.run(me.sync)            # main.py:456 starts every worker from the global weights (process() syncs again each call)
  pick = np.random.RandomState(SEED + 100)
  idx = {name: pick.randint(0, int(np.prod(shape)), size=64) for name, shape, _ in specs}
  out = {"meta": np.array([A, G, SEED, NET_SEED, H, N_PROCESS, MAX_T]), "initial_lr": np.array(INITIAL_LR)}
  out.update({"idx_" + k: v for k, v in idx.items()})

  def snapshot(tag, ret=(0, None)):
    probe = np.random.RandomState()
    probe.set_state(rs.get_state())
    out[tag + "_local_t"] = np.array(me.local_t)
    out[tag + "_ret"] = np.array([ret[0], np.nan if ret[1] is None else ret[1]], np.float64)
    out[tag + "_pos"] = np.array([me.environment.x, me.environment.y])
    out[tag + "_next_draw"] = np.array(probe.randint(0, 2 ** 31 - 1))
    out[tag + "_lstm_c"] = np.asarray(local.base_lstm_state_out[0], np.float64)
    out[tag + "_lstm_h"] = np.asarray(local.base_lstm_state_out[1], np.float64)
    for (name, _, _), v in zip(specs, glob.get_vars()):
      val = v.value.detach().numpy().reshape(-1)
      out[tag + "_val_" + name] = val[idx[name]]
      out[tag + "_sum_" + name] = val.sum()

  global_t, n_fill = 0, 0
  with contextlib.redirect_stdout(io.StringIO()):
    while not me.experience.is_full():
      ret = Trainer.process(me, sess, global_t, None, {'losses_input': None}, None, None, None, None, None, {})
      assert ret == (0, None)
      n_fill += 1
    out["n_fill"] = np.array(n_fill)
    snapshot("fill")
    for it in range(N_PROCESS):
      ret = Trainer.process(me, sess, global_t, None, {'losses_input': None, 'score_input': None, 'sr_input': None}, None, None,
                            None, None, None, {})
      global_t += ret[0]                  # main.py:125
      snapshot("it%d" % it, ret)
  path = os.path.join(HERE, "agent_reference_golden.npz")
  np.savez_compressed(path, **out)
  print("wrote %s (%d bytes); fill calls %d; local_t after each process:" % (path, os.path.getsize(path), n_fill),
        [int(out["it%d_local_t" % i]) for i in range(N_PROCESS)], "returns", [out["it%d_ret" % i].tolist() for i in range(N_PROCESS)])


if __name__ == "__main__":
  main()
