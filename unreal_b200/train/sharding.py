"""Env partition across the GPUs of one box: rank g owns the contiguous env range
[g*N/W, (g+1)*N/W) and its rings / RNG streams / rollout buffers; nothing of K1-K5 crosses ranks.
An env's RandomState seed depends only on its GLOBAL index, so a run is the same set of per-env
streams whatever the world size."""
import os


def env_shard(total_envs, world_size, rank):
  """-> (lo, hi): the global env indices this rank owns.  Remainders go to the low ranks."""
  if world_size < 1 or not (0 <= rank < world_size):
    raise ValueError("bad rank %d / world %d" % (rank, world_size))
  if total_envs < 0:
    raise ValueError("negative env count")
  base, rem = divmod(total_envs, world_size)
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def env_seeds(base_seed, lo, hi):
  """numpy-legacy RandomState seeds of envs lo..hi-1 (32-bit, like np.random.RandomState(seed))."""
  return [(int(base_seed) + e) & 0xFFFFFFFF for e in range(lo, hi)]


def dist_env():
  """(rank, world, local_rank) from the torchrun environment (1-process defaults)."""
  return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
          int(os.environ.get("LOCAL_RANK", "0")))
