"""CPU checks of the host+device core headers (maze / MT19937 / ring index logic) that the
CUDA kernels are built from, compiled with g++ (tests/cpu_harness).  Test infrastructure only:
the package never loads this library."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import unreal_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_harness", "core_harness.cpp")
OUT = os.path.join(HERE, "cpu_harness", "libcore_harness.so")
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def H():
  if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
    pytest.skip("g++ or CUDA headers not available")
  subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CUDA_INC, SRC, "-o", OUT])
  lib = ctypes.CDLL(OUT)
  lib.h_frame_pack.restype = ctypes.c_uint64
  return lib


def _load(golden_dir, name):
  with np.load(os.path.join(golden_dir, name)) as z:
    return {k: z[k] for k in z.files}


def test_maze_core_all_pairs(H, golden_dir):
  g = _load(golden_dir, "maze_golden.npz")
  out = (ctypes.c_int * 4)()
  for x, y, a, nx, ny, r, t in g["pair_table"]:
    H.h_maze_step(int(x), int(y), int(a), out)
    assert tuple(out) == (nx, ny, r, t)
  H.h_maze_step(0, 2, 9, out)
  assert tuple(out) == (0, 2, 0, 0)
  for c in range(7):
    assert [H.h_pc_overlap(i, c) for i in range(20)] == list(O._overlap_weights(c))


def test_mt_core_matches_numpy(H):
  for seed in (0xA3C, 0, 2 ** 32 - 1, 99):
    out = np.zeros(1500, np.int32)
    H.h_mt_randint(ctypes.c_uint32(seed), ctypes.c_uint32(1978), 1500, out.ctypes.data_as(ctypes.c_void_p))
    rs = np.random.RandomState(seed)
    assert list(out) == [rs.randint(0, 1978) for _ in range(1500)]
  gen = np.random.RandomState(1)
  pi = gen.rand(500, 4).astype(np.float32)
  pi /= pi.sum(1, keepdims=True)
  out = np.zeros(500, np.int32)
  H.h_mt_choice(ctypes.c_uint32(5), pi.ctypes.data_as(ctypes.c_void_p), 4, 500, out.ctypes.data_as(ctypes.c_void_p))
  rs = np.random.RandomState(5)
  assert list(out) == [rs.choice(4, p=pi[i]) for i in range(500)]


@pytest.mark.parametrize("name", ["h2000", "h64", "h16", "h40neg"])
def test_ring_core_matches_reference(H, golden_dir, name):
  g = _load(golden_dir, "experience_golden.npz")
  Hs, L, seed, n, every = [int(v) for v in g[name + "_cfg"]]
  want = g[name + "_log"]
  log = np.zeros((len(want) + 4, 7), np.int64)
  rew = np.ascontiguousarray(g[name + "_reward"]); term = np.ascontiguousarray(g[name + "_terminal"])
  rows = H.h_ring_protocol(Hs, L, ctypes.c_uint32(seed), n, every, rew.ctypes.data_as(ctypes.c_void_p),
                           term.ctypes.data_as(ctypes.c_void_p), log.ctypes.data_as(ctypes.c_void_p), len(log))
  assert rows == len(want)
  assert np.array_equal(log[:rows], want)
