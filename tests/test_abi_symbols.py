"""CPU: the C-ABI library builds, loads, and exports every symbol include/unreal_b200.h
declares; the ctypes table mirrors the header; argument errors are reported without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_mod():
  import __graft_entry__ as ge
  ge.build()                      # nvcc cross-compiles for sm_100a without a GPU
  from unreal_b200 import _lib
  return _lib


def _declared():
  src = open(os.path.join(ROOT, "include", "unreal_b200.h")).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  return sorted(set(re.findall(r"\b(unreal_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib_mod):
  names = _declared()
  assert len(names) >= 28
  raw = ctypes.CDLL(lib_mod.LIB_PATH)
  for n in names:
    assert hasattr(raw, n), "header declares %s but the library does not export it" % n
  assert sorted(lib_mod.SIGNATURES) == names, "ctypes table and header differ"
  assert lib_mod.MISSING == []


def test_library_is_sm100a_only(lib_mod):
  import shutil
  import subprocess
  cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
  if not os.path.exists(cuobjdump):
    pytest.skip("cuobjdump not available")
  out = subprocess.run([cuobjdump, "-lelf", lib_mod.LIB_PATH], capture_output=True, text=True).stdout
  archs = set(re.findall(r"sm_(\d+a?)", out))
  assert archs == {"100a"}, archs


def test_argument_errors_without_gpu(lib_mod):
  L = lib_mod.lib
  assert L.unreal_abi_version() == 1
  assert L.unreal_maze_step(None, None, None, None, None, None, None, None, 0, None, None, -1, 0, None) == -1
  assert b"n < 0" in L.unreal_last_error()
  assert L.unreal_nstep_returns(None, None, None, None, 0.99, None, None, 4, -2, None) == -1
  assert L.unreal_pixel_change(None, None, 0, None, 1, 85, 84, 3, None) == -1
  assert b"multiples of 4" in L.unreal_last_error()
  assert L.unreal_replay_create(None, 4, 100) == -1
  assert L.unreal_set_tunable(b"x", 3) == 0
  v = ctypes.c_int()
  assert L.unreal_get_tunable(b"x", ctypes.byref(v)) == 0 and v.value == 3
  # empty batches are legal no-ops even with null pointers
  assert L.unreal_nstep_returns(None, None, None, None, 0.99, None, None, 0, 0, None) == 0
  assert L.unreal_pc_targets(None, None, None, None, 0.9, None, 0, 5, None) == 0


def test_product_fails_loudly_without_device(lib_mod):
  import torch
  if torch.cuda.is_available():
    pytest.skip("a GPU is present")
  with pytest.raises(lib_mod.UnrealError):
    lib_mod.require_device()
  from unreal_b200.environment.environment import Environment
  with pytest.raises(lib_mod.UnrealError):
    Environment.create_environment('maze', '')
  with pytest.raises(lib_mod.UnrealError):
    Environment.create_environment('lab', 'nav_maze_static_01')


def test_no_product_module_imports_the_oracle():
  bad = []
  for dp, _, files in os.walk(os.path.join(ROOT, "unreal_b200")):
    for f in files:
      if f.endswith(".py"):
        txt = open(os.path.join(dp, f)).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
          bad.append(os.path.join(dp, f))
  assert not bad, "product code must never import the oracle: %s" % bad
