"""TensorFlow-free reader for the TF2 object-graph checkpoints the reference saves with
``actor_model.save_weights("{ep}pickle_base/AgentK_Actor_pickle")``
(``train/code/master_DDPG_truss2D_MO.py:885-906``; loaded at ``test/*/code/master_DDPG_truss2D_MO.py:778-800``).

On-disk format (tensor bundle):
  ``<prefix>.index``                an uncompressed LevelDB-style table: 48-byte footer (metaindex and index
                                    BlockHandles as varints, magic 0xdb4775248b80fb57), prefix-compressed
                                    key/value blocks each followed by a 1-byte compression tag and a 4-byte
                                    CRC; values are ``BundleEntryProto`` messages
                                    (1 dtype, 2 shape, 3 shard_id, 4 offset, 5 size, 6 crc32c)
  ``<prefix>.data-00000-of-00001``  raw little-endian tensor bytes

Only what the actor needs is implemented: float32 / int64 tensors in a single shard, no compression.
Host logic only (no GPU involved).
"""
from __future__ import annotations

import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _read_block(data, offset, size):
    if data[offset + size] != 0:
        raise ValueError("compressed checkpoint index blocks are not supported")
    block = data[offset:offset + size]
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def _parse_entry(buf):
    """BundleEntryProto -> dict(dtype, shape, shard, offset, size)"""
    pos, ent = 0, {"dtype": 0, "shape": (), "shard": 0, "offset": 0, "size": 0}
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            val, pos = _varint(buf, pos)
            if field == 1:
                ent["dtype"] = val
            elif field == 3:
                ent["shard"] = val
            elif field == 4:
                ent["offset"] = val
            elif field == 5:
                ent["size"] = val
        elif wire == 2:
            ln, pos = _varint(buf, pos)
            sub = buf[pos:pos + ln]
            pos += ln
            if field == 2:                           # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }
                dims, sp = [], 0
                while sp < len(sub):
                    t2, sp = _varint(sub, sp)
                    if t2 >> 3 == 2 and t2 & 7 == 2:
                        dl, sp = _varint(sub, sp)
                        dim, dp, size = sub[sp:sp + dl], 0, 0
                        sp += dl
                        while dp < len(dim):
                            t3, dp = _varint(dim, dp)
                            if t3 & 7 == 0:
                                v, dp = _varint(dim, dp)
                                if t3 >> 3 == 1:
                                    size = v
                            elif t3 & 7 == 2:
                                l3, dp = _varint(dim, dp)
                                dp += l3
                        dims.append(size)
                    elif t2 & 7 == 0:
                        _, sp = _varint(sub, sp)
                    elif t2 & 7 == 2:
                        l2, sp = _varint(sub, sp)
                        sp += l2
                ent["shape"] = tuple(dims)
        elif wire == 5:
            pos += 4
        elif wire == 1:
            pos += 8
        else:
            raise ValueError("unexpected wire type %d" % wire)
    return ent


def list_entries(prefix: str):
    data = open(prefix + ".index", "rb").read()
    footer = data[-48:]
    if struct.unpack("<Q", footer[-8:])[0] != _MAGIC:
        raise ValueError("%s.index: bad table magic" % prefix)
    pos = 0
    _, pos = _varint(footer, pos)      # metaindex offset
    _, pos = _varint(footer, pos)      # metaindex size
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _read_block(data, idx_off, idx_size):
        off, hp = _varint(handle, 0)
        size, hp = _varint(handle, hp)
        for key, val in _read_block(data, off, size):
            if key == b"":
                continue                                  # BundleHeaderProto
            entries[key.decode("utf-8", "replace")] = _parse_entry(val)
    return entries


def load_checkpoint(prefix: str):
    """-> {variable key: ndarray} for every float32/int tensor stored in shard 0"""
    entries = list_entries(prefix)
    shard = prefix + ".data-00000-of-00001"
    if not os.path.exists(shard):
        raise FileNotFoundError(shard)
    raw = open(shard, "rb").read()
    out = {}
    for key, ent in entries.items():
        dt = _DTYPES.get(ent["dtype"])
        if dt is None or ent["shard"] != 0:
            continue
        n = int(np.prod(ent["shape"])) if ent["shape"] else 1
        if n * np.dtype(dt).itemsize != ent["size"]:
            continue
        arr = np.frombuffer(raw, dtype=dt, count=n, offset=ent["offset"]).reshape(ent["shape"]).copy()
        out[key] = arr
    return out


ACTOR_LAYERS = ("gcn_l1_1", "gcn_l1_2", "gcn_l1_3", "gcn_l1_4", "gcn_l2_1", "gcn_l2_2", "gcn_l2_3", "gcn_l2_4",
                "gcn_l2_5", "gcn_l3_1", "gcn_l3_2", "gcn_l4_1", "gcn_l4_2")


def load_actor_weights(prefix: str):
    """-> {layer: (kernel [in,out] float32, bias [out] float32)} for the 13 GCNConv layers of
    ``multimodes_actor`` (``truss2D_RL.py:49-72``)."""
    ck = load_checkpoint(prefix)
    out = {}
    for name in ACTOR_LAYERS:
        k = ck.get("%s/kernel/.ATTRIBUTES/VARIABLE_VALUE" % name)
        b = ck.get("%s/bias/.ATTRIBUTES/VARIABLE_VALUE" % name)
        if k is None or b is None:
            raise KeyError("checkpoint %s has no %s kernel/bias" % (prefix, name))
        out[name] = (np.ascontiguousarray(k, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32))
    return out


def random_actor_weights(seed: int = 0, hidden: int = 200):
    """Glorot-normal kernels / zero biases of the reference's architecture (``truss2D_RL.py:26, 53-68``),
    for synthetic benchmarks when no checkpoint is shipped."""
    rng = np.random.RandomState(seed)
    shapes = {"gcn_l1_1": (13, hidden), "gcn_l1_2": (13, hidden), "gcn_l1_3": (13, hidden), "gcn_l1_4": (4, hidden),
              "gcn_l4_1": (hidden, 2), "gcn_l4_2": (hidden, 3)}
    out = {}
    for name in ACTOR_LAYERS:
        fi, fo = shapes.get(name, (hidden, hidden))
        std = np.sqrt(2.0 / (fi + fo))
        out[name] = ((rng.randn(fi, fo) * std).astype(np.float32), np.zeros(fo, dtype=np.float32))
    return out
