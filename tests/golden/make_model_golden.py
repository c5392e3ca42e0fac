#!/usr/bin/env python
"""Golden vectors of the model path, produced by THE REFERENCE'S OWN model/model.py.

Run only in the authoring container (needs the read-only reference tree):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_model_golden.py [/root/reference]

TensorFlow is the reference's un-vendored dependency and cannot be installed here, so ``tensorflow`` resolves to
``tests/golden/tf1_shim`` -- ~30 TF-1 ops restated from their published definitions over torch float64 (see its docstring).
Everything else is the reference, unmodified: ``UnrealModel.__init__`` builds its graph (placeholders, scopes, variable
creation order and reuse across the four towers), ``prepare_loss()`` builds the losses, ``run_base_policy_and_value /
run_base_value / run_pc_q_max / run_vr_value / run_rp_c`` run with their own feed dicts, the gradient of
``total_loss`` w.r.t. ``get_vars()`` is what ``RMSPropApplier.minimize_local`` asks TF for (rmsprop_applier.py:109-116), and
the learner step itself is the reference's ``train/rmsprop_applier.py`` too: a global network, ``sync_from``, then
``minimize_local`` (clip_by_global_norm + ApplyRMSProp on slots it creates) run twice with the learning rate fed through
its placeholder -- the second time with a gradient above the clip norm.

The variables are set from ``oracle.model_oracle.init_params`` (by creation ORDER, after checking that the reference created
the same 20 shapes in the same order); the test (tests/test_model_oracle.py::test_oracle_matches_the_references_model_py)
rebuilds them from the seed.  Written: ``model_reference_golden.npz`` (inputs as uint8 / small arrays, outputs float64,
gradients as per-variable sums, norms and 64 sampled entries).
"""
import os
import sys

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
from model.model import UnrealModel  # noqa: E402  (reference, unmodified)
from train.rmsprop_applier import RMSPropApplier  # noqa: E402  (reference, unmodified)
from oracle import model_oracle as MO  # noqa: E402

A, G, SEED = 4, 0, 7            # rebound by main(): the maze's action space, then the indoor pointgoal one with its objective
T_BASE, L_PC, L_VR = 5, 4, 3


def frames(rs, n):
  """Blocky random frames (12 x 12 blocks of one colour, so the .npz compresses): uint8 [n,84,84,3]."""
  blocks = rs.randint(0, 256, size=(n, 7, 7, 3)).astype(np.uint8)
  return np.repeat(np.repeat(blocks, 12, axis=1), 12, axis=2)


def lar(rs, n):
  out = np.zeros((n, A + 1 + G), np.float64)
  out[np.arange(n), rs.randint(0, A, size=n)] = 1.0
  out[:, A] = rs.randint(-1, 2, size=n)
  if G:
    out[:, A + 1:] = rs.randn(n, G)           # the objective vector (indoor_environment.py: distance / direction to the goal)
  return out


def onehot(idx, k):
  out = np.zeros((len(idx), k), np.float64)
  out[np.arange(len(idx)), idx] = 1.0
  return out


def main(a=4, g=0, seed=7, fname="model_reference_golden.npz"):
  global A, G, SEED
  A, G, SEED = a, g, seed
  tf.reset_default_graph()
  rs = np.random.RandomState(SEED)
  net = UnrealModel(A, G, 0, True, True, True, True, 0.05, 0.001, "/cpu:0", {'segnet_mode': 0}, [84, 84], True, 0, 0.0, 0.0)
  net.prepare_loss()
  ref_vars = net.get_vars()
  specs = MO.variable_specs(A, G)
  assert len(ref_vars) == len(specs) == 20
  names = []
  for v, (name, shape, _) in zip(ref_vars, specs):
    last = v.name.split("/")[-1].split(":")[0]
    assert tuple(v.get_shape().as_list()) == tuple(shape), (v.name, shape)
    assert last == name or (last, name) in (("kernel", "lstm_kernel"), ("bias", "lstm_bias")), (v.name, name)
    names.append(v.name)
  params = MO.init_params(A, G, seed=SEED)
  for v, (name, _, _) in zip(ref_vars, specs):
    v.value = params[name].to(torch.float64).clone()
  params["lstm_bias"] = params["lstm_bias"] + 0.0          # (zeros; the test perturbs nothing)

  sess = tf.Session()
  out = {"variable_names": np.array(names)}
  # ---- acting: three steps of run_base_policy_and_value carry the LSTM state; the other run_* leave it alone
  f_act, l_act = frames(rs, 3), lar(rs, 3)
  net.reset_state()
  pis, vs = [], []
  for t in range(3):
    pi, v, _ = net.run_base_policy_and_value(sess, {'image': f_act[t] / 255.0}, l_act[t])
    pis.append(pi); vs.append(v)
  out.update(act_frames=f_act, act_lar=l_act, act_pi=np.array(pis), act_v=np.array(vs),
             act_state_c=np.asarray(net.base_lstm_state_out[0]), act_state_h=np.asarray(net.base_lstm_state_out[1]))
  f_one, l_one = frames(rs, 1), lar(rs, 1)
  out.update(one_frame=f_one, one_lar=l_one,
             base_value=net.run_base_value(sess, {'image': f_one[0] / 255.0}, l_one[0]),
             pc_q_max=net.run_pc_q_max(sess, {'image': f_one[0] / 255.0}, l_one[0]),
             vr_value=net.run_vr_value(sess, {'image': f_one[0] / 255.0}, l_one[0]))
  f_rp = frames(rs, 3)
  out.update(rp_frames=f_rp, rp_c=net.run_rp_c(sess, [{'image': f / 255.0} for f in f_rp]))
  # ---- one training feed, as Trainer.process assembles it (trainer.py:500-541)
  fb, lb = frames(rs, T_BASE), lar(rs, T_BASE)
  ab = onehot(rs.randint(0, A, size=T_BASE), A)
  adv, R = rs.randn(T_BASE), rs.randn(T_BASE)
  c0, h0 = rs.randn(1, 256) * 0.3, rs.randn(1, 256) * 0.3
  fp, lp = frames(rs, L_PC), lar(rs, L_PC)
  ap = onehot(rs.randint(0, A, size=L_PC), A)
  pcr = rs.rand(L_PC, 20, 20)
  fv, lv = frames(rs, L_VR), lar(rs, L_VR)
  vrr = rs.randn(L_VR)
  frp = frames(rs, 3)
  rpc = onehot([2], 3)
  feed = {net.base_input: fb / 255.0, net.base_last_action_reward_input: lb, net.base_a: ab, net.base_adv: adv,
          net.base_r: R, net.base_initial_lstm_state0: c0, net.base_initial_lstm_state1: h0,
          net.pc_input: fp / 255.0, net.pc_last_action_reward_input: lp, net.pc_a: ap, net.pc_r: pcr,
          net.vr_input: fv / 255.0, net.vr_last_action_reward_input: lv, net.vr_r: vrr,
          net.rp_input: frp / 255.0, net.rp_c_target: rpc}
  grads = tf.gradients(net.total_loss, ref_vars)            # rmsprop_applier.py:109-116 (minimize_local)
  res = sess.run([net.total_loss, net.policy_loss, net.value_loss, net.pc_loss, net.vr_loss, net.rp_loss, net.entropy,
                  net.base_pi, net.base_v, net.pc_q, net.vr_v, net.rp_c] + grads, feed_dict=feed)
  keys = ("total_loss", "policy_loss", "value_loss", "pc_loss", "vr_loss", "rp_loss", "entropy", "base_pi", "base_v", "pc_q",
          "vr_v", "train_rp_c")
  out.update({k: np.asarray(v) for k, v in zip(keys, res[:len(keys)])})
  out.update(base_frames=fb, base_lar=lb, base_a=ab, base_adv=adv, base_R=R, base_c0=c0, base_h0=h0, pc_frames=fp, pc_lar=lp,
             pc_a=ap, pc_R=pcr, vr_frames=fv, vr_lar=lv, vr_R=vrr, rp_train_frames=frp, rp_c_target=rpc)
  pick = np.random.RandomState(SEED + 1)
  for (name, shape, _), g in zip(specs, res[len(keys):]):
    g = np.asarray(g).reshape(-1)
    idx = pick.randint(0, g.size, size=64)
    out["grad_sum_" + name] = g.sum()
    out["grad_norm_" + name] = np.sqrt((g * g).sum())
    out["grad_idx_" + name] = idx
    out["grad_val_" + name] = g[idx]
  # ---- the learner step as main.py / trainer.py wire it (main.py:299-318, trainer.py:120-133): a GLOBAL network, the worker's
  # local copy synced from it (sync_from), gradients of the LOCAL loss applied to the GLOBAL variables by the reference's
  # RMSPropApplier.minimize_local (global-norm clip at 40 + shared RMSProp, slots rms = 1 / momentum = 0); two updates with
  # the annealed learning rate fed through its placeholder
  glob = UnrealModel(A, G, -1, True, True, True, True, 0.05, 0.001, "/cpu:0", {'segnet_mode': 0}, [84, 84], True, 0, 0.0, 0.0)
  for v, (name, _, _) in zip(glob.get_vars(), specs):
    v.value = params[name].to(torch.float64).clone()
  lr_in = tf.placeholder("float")
  applier = RMSPropApplier(learning_rate=lr_in, decay=0.99, momentum=0.0, epsilon=0.1, clip_norm=40.0, device="/cpu:0")
  apply_op, grad_norm = applier.minimize_local(net.total_loss, glob.get_vars(), net.get_vars(), 0)
  sync = net.sync_from(glob)
  norms = []
  for step, lr_now in enumerate((7e-4, 6.5e-4)):
    sess.run(sync)
    feed[lr_in] = lr_now
    norms.append(float(sess.run([apply_op, grad_norm], feed_dict=feed)[1]))
    feed[net.base_adv] = feed[net.base_adv] * 30.0        # the second update's gradient exceeds the clip norm
  out["update_grad_norms"] = np.array(norms)
  out["update_lrs"] = np.array([7e-4, 6.5e-4])
  for (name, _, _), v in zip(specs, glob.get_vars()):
    val = v.value.detach().numpy().reshape(-1)
    idx = pick.randint(0, val.size, size=64)
    out["upd_sum_" + name] = (val - params[name].to(torch.float64).numpy().reshape(-1)).sum()
    out["upd_idx_" + name] = idx
    out["upd_val_" + name] = val[idx]
    rms = applier.get_slot(v, "rms").value.numpy().reshape(-1)
    out["upd_rms_" + name] = rms[idx]
  out["meta"] = np.array([A, G, SEED, T_BASE, L_PC, L_VR])
  path = os.path.join(HERE, fname)
  np.savez_compressed(path, **out)
  print("wrote %s (%d bytes): total_loss %.9f" % (path, os.path.getsize(path), float(out["total_loss"])))
  print({k: float(out[k]) for k in ("policy_loss", "value_loss", "pc_loss", "vr_loss", "rp_loss")})


if __name__ == "__main__":
  main()                                                           # the maze: 4 actions, no objective
  main(3, 2, 8, "model_reference_golden_a3g2.npz")                 # indoor pointgoal: 3 actions + a 2-vector objective
