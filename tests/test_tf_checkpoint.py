"""TF-1 checkpoint reader / writer (unreal_b200/train/tf_checkpoint.py; the reference's tf.train.Saver, main.py:356,
:363-381) without TensorFlow: known-answer CRC-32C, a hand-assembled LevelDB block with prefix-compressed keys and a
snappy block, writer -> reader round trips over many blocks, corruption detection, and the reference's pairing of
checkpoint keys with variables by their last path component."""
import struct

import numpy as np
import pytest

from unreal_b200.train import tf_checkpoint as T


def test_crc32c_known_answers():
  # RFC 3720 B.4 test vectors
  assert T.crc32c(b"\x00" * 32) == 0x8A9136AA
  assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
  assert T.crc32c(bytes(range(32))) == 0x46DD794E
  assert T.crc32c(b"123456789") == 0xE3069283
  # LevelDB's mask is a rotation plus a constant, and not the identity
  c = T.crc32c(b"foo")
  assert T._mask(c) != c and T._mask(c) == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_hand_assembled_block_with_shared_prefixes():
  # entries: (shared, unshared, value length, key suffix, value); one restart at 0
  block = (bytes([0, 5, 1]) + b"net/a" + b"1" + bytes([4, 2, 2]) + b"bc" + b"22" + bytes([5, 1, 0]) + b"d" +
           struct.pack("<I", 0) + struct.pack("<I", 1))
  assert list(T._block_entries(block)) == [(b"net/a", b"1"), (b"net/bc", b"22"), (b"net/bd", b"")]
  with pytest.raises(T.TFCheckpointError):
    list(T._block_entries(bytes([9, 1, 0]) + b"x" + struct.pack("<II", 0, 1)))     # shares more than the previous key has


def test_snappy_block():
  # literal "abcd", copy (1-byte offset) of 8 bytes from 4 back, literal "Z"
  comp = bytes([13]) + bytes([3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) + bytes([0 << 2]) + b"Z"
  assert T._snappy_decompress(comp) == b"abcdabcdabcdZ"
  with pytest.raises(T.TFCheckpointError):
    T._snappy_decompress(bytes([5]) + bytes([((8 - 4) << 2) | 1, 9]))               # copy from before the start


def test_bundle_entry_proto_round_trip():
  e = T._parse_entry(T._build_entry(1, (4, 4, 16, 32), 123456, 32768, 0xDEADBEEF))
  assert e == dict(dtype=1, shape=(4, 4, 16, 32), shard_id=0, offset=123456, size=32768, crc32c=0xDEADBEEF)
  e = T._parse_entry(T._build_entry(3, (), 0, 4, 7))
  assert e["shape"] == () and e["offset"] == 0 and e["dtype"] == 3
  # bytes of a BundleEntryProto written field by field: dtype DT_FLOAT, shape [2,3], offset 8, size 24, crc
  raw = b"\x08\x01" + b"\x12\x08" + b"\x12\x02\x08\x02" + b"\x12\x02\x08\x03" + b"\x20\x08" + b"\x28\x18" + b"\x35" + struct.pack("<I", 5)
  assert T._parse_entry(raw) == dict(dtype=1, shape=(2, 3), shard_id=0, offset=8, size=24, crc32c=5)


def _variables(rs, n_extra=0):
  named = {
      "net_-1/base_encoder/W_base_conv1": rs.randn(8, 8, 3, 16).astype(np.float32),
      "net_-1/base_encoder/b_base_conv1": rs.randn(16).astype(np.float32),
      "net_-1/base_lstm_layer/basic_lstm_cell/kernel": rs.randn(261, 1024).astype(np.float32),
      "net_-1/base_lstm_layer/basic_lstm_cell/bias": rs.randn(1024).astype(np.float32),
      "global_step": np.array(1234567, np.int64),
      "some/double": rs.randn(3, 2),
  }
  for i in range(n_extra):                      # enough keys for several 4 KB index blocks
    named["net_-1/extra/var_%04d/with_a_long_shared_prefix" % i] = rs.randn(i % 5 + 1).astype(np.float32)
  return named


@pytest.mark.parametrize("n_extra", [0, 400])
def test_write_then_read(tmp_path, n_extra):
  named = _variables(np.random.RandomState(3), n_extra)
  prefix = str(tmp_path / "checkpoint-0.912345-7000000")
  T.write_tf_checkpoint(prefix, named)
  got = T.read_tf_checkpoint(prefix)
  assert sorted(got) == sorted(named)
  for k, v in named.items():
    assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape and np.array_equal(got[k], v), k
  index = open(prefix + ".index", "rb").read()
  assert struct.unpack("<Q", index[-8:])[0] == 0xdb4775248b80fb57
  if n_extra:
    assert len(index) > 3 * 4096                # the index really spans several data blocks


def test_corruption_is_detected(tmp_path):
  named = _variables(np.random.RandomState(4))
  prefix = str(tmp_path / "checkpoint-10")
  T.write_tf_checkpoint(prefix, named)
  data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
  data[100] ^= 1
  open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
  with pytest.raises(T.TFCheckpointError, match="tensor CRC"):
    T.read_tf_checkpoint(prefix)
  assert "global_step" in T.read_tf_checkpoint(prefix, verify=False)
  index = bytearray(open(prefix + ".index", "rb").read())
  index[10] ^= 1
  open(prefix + ".index", "wb").write(bytes(index))
  with pytest.raises(T.TFCheckpointError, match="CRC mismatch"):
    T.read_tf_checkpoint(prefix)
  open(prefix + ".index", "wb").write(b"not a table" * 10)
  with pytest.raises(T.TFCheckpointError, match="magic"):
    T.read_tf_checkpoint(prefix)


class _Net(object):
  """The two members of UnrealModel the checkpoint functions use."""

  def __init__(self, shapes):
    import torch
    self.vars = {k: torch.zeros(s) for k, s in shapes.items()}

  def named_vars(self):
    return self.vars

  def load_vars(self, named):
    import torch
    for k, v in named.items():
      self.vars[k].copy_(torch.as_tensor(np.asarray(v)))


SHAPES = {"W_base_conv1": (8, 8, 3, 16), "b_base_conv1": (16,), "lstm_kernel": (261, 1024), "lstm_bias": (1024,)}


def test_model_round_trip_under_the_reference_names(tmp_path):
  rs = np.random.RandomState(5)
  src = _Net(SHAPES)
  for v in src.vars.values():
    v.copy_(__import__("torch").from_numpy(rs.randn(*v.shape).astype(np.float32)))
  prefix = str(tmp_path / "checkpoint-0.5-4200000")
  T.save_tf_checkpoint(src, prefix)
  raw = T.read_tf_checkpoint(prefix)
  assert "net_-1/base_encoder/W_base_conv1" in raw and "net_-1/base_lstm_layer/basic_lstm_cell/kernel" in raw
  dst = _Net(SHAPES)
  assert T.load_tf_checkpoint(dst, prefix) == 4200000       # main.py:382-412: the step is the file name's last token
  for k in SHAPES:
    assert np.array_equal(dst.vars[k].numpy(), src.vars[k].numpy()), k
  # an older TF-1 release's cell names, another worker scope
  named = {k.replace("net_-1", "net_3").replace("kernel", "weights").replace("/bias", "/biases"): v for k, v in raw.items()}
  T.write_tf_checkpoint(str(tmp_path / "checkpoint-77"), named)
  dst2 = _Net(SHAPES)
  assert T.load_tf_checkpoint(dst2, str(tmp_path / "checkpoint-77")) == 77
  assert np.array_equal(dst2.vars["lstm_kernel"].numpy(), src.vars["lstm_kernel"].numpy())


def test_restore_is_all_or_nothing(tmp_path):
  rs = np.random.RandomState(6)
  named = {"net_-1/base_encoder/W_base_conv1": rs.randn(8, 8, 3, 16).astype(np.float32),
           "net_-1/base_encoder/b_base_conv1": rs.randn(16).astype(np.float32),
           "net_-1/base_lstm_layer/basic_lstm_cell/kernel": rs.randn(263, 1024).astype(np.float32),     # another action count
           "net_-1/base_lstm_layer/basic_lstm_cell/bias": rs.randn(1024).astype(np.float32)}
  T.write_tf_checkpoint(str(tmp_path / "checkpoint-1"), named)
  dst = _Net(SHAPES)
  with pytest.raises(T.TFCheckpointError, match="shape"):
    T.load_tf_checkpoint(dst, str(tmp_path / "checkpoint-1"))
  assert all(float(v.abs().sum()) == 0.0 for v in dst.vars.values())
  del named["net_-1/base_encoder/b_base_conv1"]
  named["net_-1/base_lstm_layer/basic_lstm_cell/kernel"] = rs.randn(261, 1024).astype(np.float32)
  T.write_tf_checkpoint(str(tmp_path / "checkpoint-2"), named)
  with pytest.raises(T.TFCheckpointError, match="no variable for b_base_conv1"):
    T.load_tf_checkpoint(dst, str(tmp_path / "checkpoint-2"))
