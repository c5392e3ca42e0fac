"""TensorFlow-1 checkpoints without TensorFlow: read the variables the reference's `tf.train.Saver` wrote
(main.py:356 -- `tf.train.Saver(global_network.get_global_vars())`, restored at main.py:363-381) into an
`UnrealModel`, and write an `UnrealModel`'s variables in the same format under the reference's names.

Format (TF "tensor bundle", checkpoint V2; restated from the published layout, TensorFlow itself is not in this image
and the reference tree holds no checkpoint file, so the byte layout is **unpinned** against a TF-written file -- the
tests pin the reader against this module's writer and against hand-assembled blocks):

  <prefix>.index                    a LevelDB-format table: key "" -> BundleHeaderProto, key <variable name> ->
                                    BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}
  <prefix>.data-00000-of-00001      the tensors' little-endian bytes at those offsets

  table  = data blocks | metaindex block | index block | footer (two block handles, padding to 40 bytes, magic)
  block  = entries | restart offsets (u32 each) | number of restarts (u32), followed in the file by a type byte
           (0 raw, 1 snappy) and the masked CRC-32C of block + type byte
  entry  = varint shared key bytes | varint unshared key bytes | varint value bytes | key suffix | value

Variable names.  The reference's global network lives in scope `net_-1` with sub-scopes per tower
(`net_-1/base_encoder/W_base_conv1`, `net_-1/base_lstm_layer/basic_lstm_cell/kernel`, ...); its own restore path pairs
variables with checkpoint keys by the LAST path component (main.py:366-378), and so does `load_tf_checkpoint`
(`kernel` / `weights` and `bias` / `biases` of the BasicLSTMCell, whose names changed across TF-1 releases, map to
`lstm_kernel` / `lstm_bias`).  Layouts are TF's own -- the flat parameter buffer already stores them that way.
"""
import os
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 6: np.int8, 9: np.int64, 19: np.float16}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}

# scope of each variable in the reference's graph (model/model.py:108, :190, :323, :359, :369, :412, :482; the cell's
# variables live where tf.nn.dynamic_rnn(scope=base_lstm_layer) creates them)
_SCOPE = {
    "W_base_conv1": "base_encoder", "b_base_conv1": "base_encoder", "W_base_conv2": "base_encoder", "b_base_conv2": "base_encoder",
    "W_base_fc1": "base_lstm_layer", "b_base_fc1": "base_lstm_layer", "lstm_kernel": "base_lstm_layer/basic_lstm_cell",
    "lstm_bias": "base_lstm_layer/basic_lstm_cell",
    "W_base_fc_p": "base_policy_layer", "b_base_fc_p": "base_policy_layer", "W_base_fc_v": "base_value_layer",
    "b_base_fc_v": "base_value_layer",
    "W_pc_fc1": "pc_deconv_layers", "b_pc_fc1": "pc_deconv_layers", "W_pc_deconv_v": "pc_deconv_layers",
    "b_pc_deconv_v": "pc_deconv_layers", "W_pc_deconv_a": "pc_deconv_layers", "b_pc_deconv_a": "pc_deconv_layers",
    "W_rp_fc1": "rp_fc", "b_rp_fc1": "rp_fc",
}
_LSTM_ENDINGS = {"kernel": "lstm_kernel", "weights": "lstm_kernel", "bias": "lstm_bias", "biases": "lstm_bias"}


class TFCheckpointError(ValueError):
  pass


# ---- CRC-32C (Castagnoli), masked as LevelDB / TensorFlow store it -------------------------------------------------
def _crc_table():
  t = []
  for i in range(256):
    c = i
    for _ in range(8):
      c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    t.append(c)
  return t


_CRC = _crc_table()


def crc32c(data, crc=0):
  c = crc ^ 0xFFFFFFFF
  tab = _CRC
  for b in bytes(data):
    c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
  return c ^ 0xFFFFFFFF


def _mask(c):
  return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints, protobuf wire format (only what the two bundle messages use) -----------------------------------------
def _varint(buf, pos):
  out = shift = 0
  while True:
    if pos >= len(buf):
      raise TFCheckpointError("truncated varint")
    b = buf[pos]
    pos += 1
    out |= (b & 0x7F) << shift
    if not b & 0x80:
      return out, pos
    shift += 7
    if shift > 63:
      raise TFCheckpointError("varint longer than 64 bits")


def _put_varint(v):
  out = bytearray()
  while True:
    b = v & 0x7F
    v >>= 7
    if v:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _proto_fields(buf):
  """-> list of (field number, wire type, value): varints as ints, length-delimited as bytes, fixed32 as int."""
  pos, out = 0, []
  while pos < len(buf):
    tag, pos = _varint(buf, pos)
    f, wt = tag >> 3, tag & 7
    if wt == 0:
      v, pos = _varint(buf, pos)
    elif wt == 2:
      n, pos = _varint(buf, pos)
      v = bytes(buf[pos:pos + n])
      if len(v) != n:
        raise TFCheckpointError("truncated protobuf field")
      pos += n
    elif wt == 5:
      v = struct.unpack_from("<I", buf, pos)[0]
      pos += 4
    elif wt == 1:
      v = struct.unpack_from("<Q", buf, pos)[0]
      pos += 8
    else:
      raise TFCheckpointError("unsupported protobuf wire type %d" % wt)
    out.append((f, wt, v))
  return out


def _parse_entry(buf):
  """BundleEntryProto: 1 dtype, 2 shape (TensorShapeProto: 2 = repeated Dim {1 size}), 3 shard_id, 4 offset, 5 size,
  6 crc32c (fixed32), 7 slices."""
  e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None)
  for f, _, v in _proto_fields(buf):
    if f == 1:
      e["dtype"] = v
    elif f == 2:
      dims = []
      for f2, _, v2 in _proto_fields(v):
        if f2 == 2:
          size = 0
          for f3, _, v3 in _proto_fields(v2):
            if f3 == 1:
              size = v3 - (1 << 64) if v3 >= (1 << 63) else v3
          dims.append(size)
        elif f2 == 3 and v2:
          raise TFCheckpointError("tensor of unknown rank in the checkpoint")
      e["shape"] = tuple(dims)
    elif f == 3:
      e["shard_id"] = v
    elif f == 4:
      e["offset"] = v
    elif f == 5:
      e["size"] = v
    elif f == 6:
      e["crc32c"] = v
    elif f == 7:
      raise TFCheckpointError("partitioned (sliced) variables are not supported")
  return e


def _build_entry(dtype_id, shape, offset, size, crc):
  dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
  out = b"\x08" + _put_varint(dtype_id) + b"\x12" + _put_varint(len(dims)) + dims
  if offset:
    out += b"\x20" + _put_varint(offset)
  out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
  return out


# ---- snappy (block format) decompression: index blocks may be compressed ----------------------------------------------
def _snappy_decompress(buf):
  n, pos = _varint(buf, 0)
  out = bytearray()
  while pos < len(buf):
    tag = buf[pos]
    pos += 1
    kind = tag & 3
    if kind == 0:
      ln = tag >> 2
      if ln >= 60:
        nb = ln - 59
        ln = int.from_bytes(buf[pos:pos + nb], "little")
        pos += nb
      ln += 1
      out += buf[pos:pos + ln]
      pos += ln
      continue
    if kind == 1:
      ln = ((tag >> 2) & 7) + 4
      off = ((tag >> 5) << 8) | buf[pos]
      pos += 1
    elif kind == 2:
      ln = (tag >> 2) + 1
      off = buf[pos] | (buf[pos + 1] << 8)
      pos += 2
    else:
      ln = (tag >> 2) + 1
      off = int.from_bytes(buf[pos:pos + 4], "little")
      pos += 4
    if off == 0 or off > len(out):
      raise TFCheckpointError("corrupt snappy block")
    for _ in range(ln):
      out.append(out[-off])
  if len(out) != n:
    raise TFCheckpointError("snappy block decompressed to %d bytes, header says %d" % (len(out), n))
  return bytes(out)


# ---- LevelDB table ---------------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify):
  end = offset + size
  if end + 5 > len(data):
    raise TFCheckpointError("block handle (%d, %d) runs past the end of the index file" % (offset, size))
  raw, kind = data[offset:end], data[end]
  if verify:
    want = struct.unpack_from("<I", data, end + 1)[0]
    got = _mask(crc32c(data[offset:end + 1]))
    if want != got:
      raise TFCheckpointError("index block at %d: CRC mismatch (stored %08x, computed %08x)" % (offset, want, got))
  if kind == 1:
    raw = _snappy_decompress(raw)
  elif kind != 0:
    raise TFCheckpointError("index block at %d: unknown compression type %d" % (offset, kind))
  return raw


def _block_entries(block):
  if len(block) < 4:
    raise TFCheckpointError("block shorter than its restart count")
  n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
  limit = len(block) - 4 * (n_restarts + 1)
  if limit < 0:
    raise TFCheckpointError("block restart array larger than the block")
  pos, key = 0, b""
  while pos < limit:
    shared, pos = _varint(block, pos)
    unshared, pos = _varint(block, pos)
    vlen, pos = _varint(block, pos)
    if shared > len(key) or pos + unshared + vlen > limit:
      raise TFCheckpointError("corrupt block entry")
    key = key[:shared] + bytes(block[pos:pos + unshared])
    pos += unshared
    yield key, bytes(block[pos:pos + vlen])
    pos += vlen


def _read_table(data, verify=True):
  if len(data) < 48:
    raise TFCheckpointError("index file shorter than a table footer")
  footer = data[-48:]
  if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
    raise TFCheckpointError("not a TensorFlow V2 checkpoint index (bad table magic)")
  _, pos = _varint(footer, 0)                 # metaindex handle
  _, pos = _varint(footer, pos)
  ioff, pos = _varint(footer, pos)
  isize, pos = _varint(footer, pos)
  out = []
  for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
    boff, p = _varint(handle, 0)
    bsize, p = _varint(handle, p)
    out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
  return out


class _BlockBuilder(object):
  def __init__(self, restart_interval=16):
    self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b"", restart_interval

  def add(self, key, value):
    shared = 0
    if self.count % self.interval == 0:
      if self.count:
        self.restarts.append(len(self.buf))
    else:
      while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
        shared += 1
    self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
    self.last = key
    self.count += 1

  def finish(self):
    return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _write_table(items, block_size=4096):
  """items: sorted (key, value) pairs -> table bytes (uncompressed blocks)."""
  out, index = bytearray(), []

  def emit(block):
    off = len(out)
    out.extend(block)
    out.append(0)
    out.extend(struct.pack("<I", _mask(crc32c(block + b"\x00"))))
    return _put_varint(off) + _put_varint(len(block))

  b = _BlockBuilder()
  for key, value in items:
    b.add(key, value)
    if len(b.buf) >= block_size:
      index.append((b.last, emit(b.finish())))
      b = _BlockBuilder()
  if b.count:
    index.append((b.last, emit(b.finish())))
  meta = emit(_BlockBuilder().finish())
  ib = _BlockBuilder(restart_interval=1)
  for key, handle in index:
    ib.add(key, handle)
  ih = emit(ib.finish())
  footer = meta + ih
  out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))
  return bytes(out)


# ---- public API ---------------------------------------------------------------------------------------------------------
def read_tf_checkpoint(prefix, verify=True):
  """<prefix>.index + data shards -> {variable name: numpy array}.  `verify`: check the blocks' and tensors' CRC-32C."""
  with open(prefix + ".index", "rb") as f:
    data = f.read()
  entries = _read_table(data, verify)
  if not entries or entries[0][0] != b"":
    raise TFCheckpointError("the index has no bundle header (key \"\")")
  num_shards, endian = 1, 0
  for f_, _, v in _proto_fields(entries[0][1]):
    if f_ == 1:
      num_shards = v
    elif f_ == 2:
      endian = v
  if endian != 0:
    raise TFCheckpointError("big-endian checkpoints are not supported")
  shards, out = {}, {}
  for key, value in entries[1:]:
    e = _parse_entry(value)
    if e["dtype"] not in _DTYPES:
      raise TFCheckpointError("%s: unsupported dtype enum %d" % (key.decode(), e["dtype"]))
    if e["shard_id"] not in shards:
      with open("%s.data-%05d-of-%05d" % (prefix, e["shard_id"], num_shards), "rb") as f:
        shards[e["shard_id"]] = f.read()
    raw = shards[e["shard_id"]][e["offset"]:e["offset"] + e["size"]]
    dt = np.dtype(_DTYPES[e["dtype"]])
    if len(raw) != e["size"] or e["size"] != int(np.prod(e["shape"], dtype=np.int64)) * dt.itemsize:
      raise TFCheckpointError("%s: %d bytes for a %s tensor of shape %s" % (key.decode(), len(raw), dt, e["shape"]))
    if verify and e["crc32c"] is not None and _mask(crc32c(raw)) != e["crc32c"]:
      raise TFCheckpointError("%s: tensor CRC mismatch" % key.decode())
    out[key.decode()] = np.frombuffer(raw, dtype=dt.newbyteorder("<")).reshape(e["shape"]).astype(dt)
  return out


def write_tf_checkpoint(prefix, named):
  """{variable name: array} -> <prefix>.index + <prefix>.data-00000-of-00001 (one shard, uncompressed blocks)."""
  items, blob = [], bytearray()
  for name in sorted(named, key=lambda s: s.encode()):
    a = np.asarray(named[name])
    a = a if a.ndim == 0 else np.ascontiguousarray(a)        # (ascontiguousarray would turn a scalar into [1])
    if a.dtype not in _DTYPE_IDS:
      raise TFCheckpointError("%s: dtype %s cannot be written" % (name, a.dtype))
    raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
    items.append((name.encode(), _build_entry(_DTYPE_IDS[a.dtype], a.shape, len(blob), len(raw), _mask(crc32c(raw)))))
    blob += raw
  header = b"\x08\x01" + b"\x1a\x02\x08\x01"          # num_shards = 1, (little endian = default), version {producer 1}
  os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
  with open(prefix + ".data-00000-of-00001", "wb") as f:
    f.write(bytes(blob))
  with open(prefix + ".index", "wb") as f:
    f.write(_write_table([(b"", header)] + items))


def tf_variable_names(net, scope="net_-1"):
  """{variable name of `UnrealModel.named_vars()`: its name in the reference's graph} (without the ":0" suffix)."""
  out = {}
  for k in net.named_vars():
    leaf = {"lstm_kernel": "kernel", "lstm_bias": "bias"}.get(k, k)
    out[k] = "%s/%s/%s" % (scope, _SCOPE.get(k, "base"), leaf)
  return out


def match_variables(net, ckpt):
  """Pair checkpoint keys with the model's variables the way the reference's restore does (main.py:366-378): by the last
  path component.  -> {model variable: checkpoint key}; raises when a variable has no or several candidates, or a
  shape differs."""
  want = {k: tuple(v.shape) for k, v in net.named_vars().items()}
  found = {}
  for key in ckpt:
    leaf = key.split("/")[-1]
    name = _LSTM_ENDINGS.get(leaf, leaf) if "lstm" in key else leaf
    if name in want:
      if name in found:
        raise TFCheckpointError("variable %s matches both %s and %s" % (name, found[name], key))
      found[name] = key
  missing = sorted(set(want) - set(found))
  if missing:
    raise TFCheckpointError("the checkpoint has no variable for %s" % ", ".join(missing))
  for name, key in found.items():
    if tuple(ckpt[key].shape) != want[name]:
      raise TFCheckpointError("%s: checkpoint shape %s, model shape %s" % (key, tuple(ckpt[key].shape), want[name]))
  return found


def load_tf_checkpoint(net, prefix, verify=True):
  """Restore an `UnrealModel` from the reference's checkpoint (main.py:363-381); nothing is written unless every variable
  matched.  -> the global step parsed from the file name the way main.py:382-412 does (`checkpoint-<step>` or
  `checkpoint-<score>-<step>`), or None."""
  ckpt = read_tf_checkpoint(prefix, verify)
  pairs = match_variables(net, ckpt)
  net.load_vars({name: ckpt[key] for name, key in pairs.items()})
  tokens = os.path.basename(prefix).split("-")
  try:
    return int(tokens[2]) if len(tokens) == 3 else int(tokens[1])
  except (IndexError, ValueError):
    return None


def save_tf_checkpoint(net, prefix, scope="net_-1"):
  """Write an `UnrealModel`'s variables as a TF-1 checkpoint under the reference's variable names (what its
  `tf.train.Saver.restore` binds; main.py:356, :469-519)."""
  names = tf_variable_names(net, scope)
  write_tf_checkpoint(prefix, {names[k]: v.detach().cpu().numpy() for k, v in net.named_vars().items()})
