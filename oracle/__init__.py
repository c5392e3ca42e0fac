"""CPU oracle for the UNREAL rollout-and-target hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import it, and there only as the checker / the timed CPU arm, never as a
fallback for the CUDA path (``unreal_b200`` raises when its CUDA library is
missing instead of routing here).

Parity status: PINNED.  Every function here is checked against outputs of the
reference's own Python (``/root/reference``, imported in the authoring
container by ``tests/golden/make_golden.py``) stored as fixtures under
``tests/golden/`` and against the one known-answer test the reference holds
(``train/rmsprop_applier_test.py``).  The TensorFlow arithmetic behind
``model/model.py`` and ``training_ops.apply_rms_prop`` is a third-party
dependency that is absent from the reference tree (``tensorflow``, version
unpinned, README.unreal.md:33 says r1.0); for those the published algorithm is
restated and the reference's own call sites and test vectors anchor it: the
reference's ``model/model.py``, ``train/rmsprop_applier.py`` and ``Trainer.process``
run UNMODIFIED over ``tests/golden/tf1_shim`` (a TF-1 graph API on torch float64)
in ``tests/golden/make_model_golden.py`` / ``make_agent_golden.py``, and
``oracle/model_oracle.py`` + the RMSProp / rollout restatements here reproduce
their outputs at 1e-9 / 1e-8 (``tests/test_model_oracle.py``).
"""
