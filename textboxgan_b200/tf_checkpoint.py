"""TensorFlow checkpoint / SavedModel-variables import and export without TensorFlow.

The reference stores everything it trains or loads as TensorFlow "tensor bundles": ``tf.train.Checkpoint`` files of
``train.py:94-108`` / ``models/model_loader.py:57-81`` (``ckpt-N.index`` + ``ckpt-N.data-00000-of-00001``), and the
``variables/variables.{index,data-*}`` pair inside the ASTER SavedModel (``aster_ocr_utils/aster_inferer.py:24-26``,
``README.md:60-66``).  TensorFlow is not installable offline, so this module reads and writes the format itself:

* the ``.index`` file is a LevelDB-format sorted string table (prefix-compressed key blocks with restart arrays, an
  index block, a 48-byte footer ending in the magic ``0xdb4775248b80fb57``), uncompressed as TensorFlow's
  ``BundleWriter`` writes it; key ``""`` holds a ``BundleHeaderProto``, every other key a ``BundleEntryProto``
  ``{dtype, shape, shard_id, offset, size, crc32c}`` pointing into a data shard;
* object-based checkpoints name variables ``<attribute path>/.ATTRIBUTES/VARIABLE_VALUE`` and store the
  ``TrackableObjectGraph`` under ``_CHECKPOINTABLE_OBJECT_GRAPH``; :class:`ObjectGraph` walks it by attribute names, which
  is how the reference's Keras objects are addressed (``generator.synthesis.synth_blocks[i].conv_0.w`` ...).

PARITY NOTE: no TensorFlow-written file exists in this environment (none in /root/reference, no network), so the reader
is pinned only against this module's own writer and the published format; the attribute paths in
:func:`generator_variable_paths` / :func:`discriminator_variable_paths` are read off the reference sources.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
VAR_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
           6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("bool"), 17: np.dtype("<u2"), 19: np.dtype("<f2"),
           22: np.dtype("<u4"), 23: np.dtype("<u8")}
_DT_STRING, _DT_BFLOAT16 = 7, 14
_NP_TO_DT = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("int64"): 9,
             np.dtype("bool"): 10, np.dtype("float16"): 19}


# ----------------------------------------------------------------------------------------------
# crc32c (Castagnoli), masked as LevelDB / TensorFlow store it
# ----------------------------------------------------------------------------------------------
_CRC_TABLE: Optional[np.ndarray] = None


def _crc_table() -> np.ndarray:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros((8, 256), dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
            t[0, i] = c
        for k in range(1, 8):
            t[k] = (t[k - 1] >> 8) ^ t[0, t[k - 1] & 0xFF]
        _CRC_TABLE = t
    return _CRC_TABLE


def _native_crc():
    """``tbg_crc32c`` of libtbg.so (host-only helper) when the library is built; None otherwise."""
    global _NATIVE
    if _NATIVE is False:
        try:
            from . import lib as _lib

            _NATIVE = _lib.load().tbg_crc32c
        except Exception:
            _NATIVE = None
    return _NATIVE


_NATIVE = False


def crc32c(data: bytes, crc: int = 0, pure_python: bool = False) -> int:
    """CRC-32C of ``data``: the library's slicing-by-8 routine, or the same algorithm in Python when libtbg.so is absent."""
    fn = None if pure_python else _native_crc()
    if fn is not None:
        buf = bytes(data)
        return int(fn(buf, len(buf), crc))
    t = _crc_table()
    t0, t1, t2, t3, t4, t5, t6, t7 = (t[k].tolist() for k in range(8))
    c = crc ^ 0xFFFFFFFF
    mv = memoryview(data)
    n8 = len(mv) // 8
    if n8:
        words = struct.unpack_from(f"<{n8}Q", mv, 0)
        for w in words:
            w ^= c
            c = (t7[w & 0xFF] ^ t6[(w >> 8) & 0xFF] ^ t5[(w >> 16) & 0xFF] ^ t4[(w >> 24) & 0xFF] ^
                 t3[(w >> 32) & 0xFF] ^ t2[(w >> 40) & 0xFF] ^ t1[(w >> 48) & 0xFF] ^ t0[(w >> 56) & 0xFF])
    for b in mv[n8 * 8:]:
        c = t0[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------
# varints and a minimal protobuf wire codec
# ----------------------------------------------------------------------------------------------
def _read_varint(buf, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """(field number, wire type, value) of one message; length-delimited values as bytes."""
    pos = 0
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, v


def _pb_tag(field: int, wt: int) -> bytes:
    return _varint((field << 3) | wt)


def _pb_bytes(field: int, payload: bytes) -> bytes:
    return _pb_tag(field, 2) + _varint(len(payload)) + payload


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


# ----------------------------------------------------------------------------------------------
# LevelDB-format table (tensorflow/core/lib/io/table*.cc, format.cc, block.cc)
# ----------------------------------------------------------------------------------------------
def _parse_block(contents: bytes) -> List[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", contents, len(contents) - 4)[0]
    end = len(contents) - 4 - 4 * n_restarts
    out, pos, key = [], 0, b""
    while pos < end:
        shared, pos = _read_varint(contents, pos)
        non_shared, pos = _read_varint(contents, pos)
        vlen, pos = _read_varint(contents, pos)
        key = key[:shared] + bytes(contents[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(contents[pos:pos + vlen])))
        pos += vlen
    return out


def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    contents = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if mask_crc(crc32c(data[offset:offset + size + 1])) != stored:
            raise ValueError("tensor bundle index: block checksum mismatch")
    if ctype != 0:
        raise NotImplementedError("tensor bundle index block is snappy-compressed; TensorFlow's BundleWriter writes "
                                  "uncompressed tables — re-save the checkpoint with TensorFlow")
    return contents


def read_table(path: str, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a tensor bundle index (bad table magic)")
    footer = data[-48:]
    _, pos = _read_varint(footer, 0)           # metaindex handle (unused)
    _, pos = _read_varint(footer, pos)
    idx_off, pos = _read_varint(footer, pos)
    idx_size, pos = _read_varint(footer, pos)
    out = []
    for _, handle in _parse_block(_read_block(data, idx_off, idx_size, verify)):
        off, p = _read_varint(handle, 0)
        size, _ = _read_varint(handle, p)
        out.extend(_parse_block(_read_block(data, off, size, verify)))
    return out


def _build_block(entries: Iterable[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    buf, restarts, last, n = bytearray(), [], b"", 0
    for key, value in entries:
        shared = 0
        if n % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        last, n = key, n + 1
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


def write_table(path: str, items: List[Tuple[bytes, bytes]], block_size: int = 4096) -> None:
    items = sorted(items)
    out = bytearray()
    index_entries = []

    def emit(block: bytes) -> bytes:
        off = len(out)
        out.extend(block)
        out.append(0)                                                       # kNoCompression
        out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _varint(off) + _varint(len(block))

    cur: List[Tuple[bytes, bytes]] = []
    cur_bytes = 0
    for kv in items:
        cur.append(kv)
        cur_bytes += len(kv[0]) + len(kv[1]) + 8
        if cur_bytes >= block_size:
            index_entries.append((cur[-1][0], emit(_build_block(cur))))
            cur, cur_bytes = [], 0
    if cur or not items:
        index_entries.append((cur[-1][0] if cur else b"", emit(_build_block(cur))))
    meta = emit(_build_block([]))
    index = emit(_build_block(index_entries, restart_interval=1))
    footer = meta + index
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(out))


# ----------------------------------------------------------------------------------------------
# tensor bundle (tensorflow/core/util/tensor_bundle/tensor_bundle.cc, protobuf/tensor_bundle.proto)
# ----------------------------------------------------------------------------------------------
def _parse_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": False}
    for field, _, v in _pb_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            for f2, _, dim in _pb_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, d in _pb_fields(dim):
                        if f3 == 1:
                            size = _signed64(d)
                    e["shape"].append(size)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = v
        elif field == 7:
            e["slices"] = True
    return e


def _shard_path(prefix: str, shard: int, num_shards: int) -> str:
    return f"{prefix}.data-{shard:05d}-of-{num_shards:05d}"


def checkpoint_prefix(path: str) -> str:
    """Accepts a prefix (``.../ckpt-10``), its ``.index`` file, a SavedModel directory (``variables/variables``) or a
    checkpoint directory (the newest ``*.index`` inside, like ``CheckpointManager.latest_checkpoint``)."""
    if path.endswith(".index"):
        return path[:-6]
    if os.path.isdir(path):
        sm = os.path.join(path, "variables", "variables")
        if os.path.exists(sm + ".index"):
            return sm
        cands = [f[:-6] for f in os.listdir(path) if f.endswith(".index")]
        if not cands:
            raise FileNotFoundError(f"no tensor bundle (*.index) in {path}")

        def num(name: str) -> int:
            tail = name.rsplit("-", 1)[-1]
            return int(tail) if tail.isdigit() else -1

        return os.path.join(path, max(cands, key=num))
    return path


def list_variables(path: str) -> List[Tuple[str, List[int], int]]:
    """(key, shape, dtype enum) of every tensor — ``tf.train.list_variables``."""
    out = []
    for key, value in read_table(checkpoint_prefix(path) + ".index"):
        if key:
            e = _parse_entry(value)
            out.append((key.decode(), e["shape"], e["dtype"]))
    return out


def load_tensor_bundle(path: str, verify_crc: bool = True, keys: Optional[Iterable[str]] = None) -> Dict[str, object]:
    """All tensors of a bundle as NumPy arrays (bfloat16 widened to float32; DT_STRING entries as ``bytes`` payloads of
    scalar strings, which is what the object graph is)."""
    prefix = checkpoint_prefix(path)
    table = read_table(prefix + ".index")
    num_shards = 1
    for key, value in table:
        if key == b"":
            for field, _, v in _pb_fields(value):
                if field == 1:
                    num_shards = v
                elif field == 2 and v != 0:
                    raise NotImplementedError("big-endian tensor bundle")
    want = set(keys) if keys is not None else None
    shards: Dict[int, bytes] = {}
    out: Dict[str, object] = {}
    for key, value in table:
        name = key.decode()
        if not name or (want is not None and name not in want):
            continue
        e = _parse_entry(value)
        if e["slices"]:
            raise NotImplementedError(f"{name}: partitioned (sliced) variables are not supported")
        if e["shard_id"] not in shards:
            shards[e["shard_id"]] = open(_shard_path(prefix, e["shard_id"], num_shards), "rb").read()
        raw = shards[e["shard_id"]][e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise ValueError(f"{name}: data shard is truncated")
        if verify_crc and e["crc32c"] is not None and e["dtype"] != _DT_STRING and mask_crc(crc32c(raw)) != e["crc32c"]:
            raise ValueError(f"{name}: tensor checksum mismatch")
        if e["dtype"] == _DT_STRING:
            # scalar string: varint length(s) + 4-byte checksum of the lengths + bytes (tensor_bundle.cc WriteStringTensor)
            n = int(np.prod(e["shape"])) if e["shape"] else 1
            pos, lens = 0, []
            for _ in range(n):
                ln, pos = _read_varint(raw, pos)
                lens.append(ln)
            pos += 4
            vals = []
            for ln in lens:
                vals.append(bytes(raw[pos:pos + ln]))
                pos += ln
            out[name] = vals[0] if n == 1 else vals
        elif e["dtype"] == _DT_BFLOAT16:
            u16 = np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16
            out[name] = u16.view(np.float32).reshape(e["shape"]).copy()
        elif e["dtype"] in _DTYPES:
            out[name] = np.frombuffer(raw, dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
        else:
            raise NotImplementedError(f"{name}: unsupported dtype enum {e['dtype']}")
    return out


def save_tensor_bundle(prefix: str, tensors: Dict[str, object]) -> None:
    """Write ``prefix.index`` + ``prefix.data-00000-of-00001`` in TensorFlow's format (one shard, little-endian);
    ``bytes`` values become scalar DT_STRING tensors."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)) or ".", exist_ok=True)
    data = bytearray()
    items = [(b"", _pb_tag(1, 0) + _varint(1) + _pb_bytes(3, _pb_tag(1, 0) + _varint(1)))]   # num_shards = 1, version.producer = 1
    for name in sorted(tensors):
        v = tensors[name]
        if isinstance(v, (bytes, bytearray)):
            lens = _varint(len(v))
            raw = lens + struct.pack("<I", mask_crc(crc32c(lens))) + bytes(v)
            dt, shape, crc = _DT_STRING, [], mask_crc(crc32c(bytes(v), crc32c(lens)))
        else:
            arr = np.asarray(v)
            if not arr.flags.c_contiguous:
                arr = arr.copy(order="C")         # (np.ascontiguousarray would turn a 0-d scalar into shape [1])
            if arr.dtype not in _NP_TO_DT:
                raise TypeError(f"{name}: dtype {arr.dtype} not supported")
            raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
            dt, shape, crc = _NP_TO_DT[arr.dtype], list(arr.shape), mask_crc(crc32c(raw))
        shape_pb = b"".join(_pb_bytes(2, _pb_tag(1, 0) + _varint(d)) for d in shape)
        entry = _pb_tag(1, 0) + _varint(dt) + _pb_bytes(2, shape_pb)
        if len(data):
            entry += _pb_tag(4, 0) + _varint(len(data))
        entry += _pb_tag(5, 0) + _varint(len(raw)) + _pb_tag(6, 5) + struct.pack("<I", crc)
        items.append((name.encode(), entry))
        data += raw
    with open(_shard_path(prefix, 0, 1), "wb") as f:
        f.write(bytes(data))
    write_table(prefix + ".index", items)


# ----------------------------------------------------------------------------------------------
# object graph (tensorflow/core/protobuf/trackable_object_graph.proto)
# ----------------------------------------------------------------------------------------------
class ObjectGraph:
    """``TrackableObjectGraph``: nodes[i] = {children: {local_name: node_id}, attributes: {name: checkpoint_key}}."""

    def __init__(self, serialized: bytes):
        self.nodes: List[dict] = []
        for field, _, node in _pb_fields(serialized):
            if field != 1:
                continue
            children, attrs = {}, {}
            for f2, _, v in _pb_fields(node):
                if f2 == 1:        # ObjectReference {node_id = 1, local_name = 2}
                    nid, lname = 0, ""
                    for f3, _, x in _pb_fields(v):
                        if f3 == 1:
                            nid = x
                        elif f3 == 2:
                            lname = x.decode()
                    children[lname] = nid
                elif f2 == 2:      # SerializedTensor {name = 1, full_name = 2, checkpoint_key = 3}
                    nm, key = "", ""
                    for f3, _, x in _pb_fields(v):
                        if f3 == 1:
                            nm = x.decode()
                        elif f3 == 3:
                            key = x.decode()
                    attrs[nm] = key
            self.nodes.append({"children": children, "attributes": attrs})

    def resolve(self, path: Iterable[str], attribute: str = "VARIABLE_VALUE") -> str:
        """Checkpoint key of the variable reached from the root by following attribute names (list entries are "0", "1", ...)."""
        node = 0
        walked = []
        for name in path:
            ch = self.nodes[node]["children"]
            if name not in ch:
                raise KeyError(f"object graph: '{'/'.join(walked) or '<root>'}' has no child '{name}' (has: {sorted(ch)[:12]})")
            node = ch[name]
            walked.append(name)
        attrs = self.nodes[node]["attributes"]
        if attribute not in attrs:
            raise KeyError(f"object graph: '{'/'.join(walked)}' holds no {attribute}")
        return attrs[attribute]

    @staticmethod
    def build(paths_to_keys: Dict[Tuple[str, ...], str]) -> bytes:
        """Serialise a graph whose variables sit at the given attribute paths (export / tests)."""
        nodes: List[dict] = [{"children": {}, "attributes": {}}]
        for path, key in paths_to_keys.items():
            node = 0
            for name in path:
                ch = nodes[node]["children"]
                if name not in ch:
                    nodes.append({"children": {}, "attributes": {}})
                    ch[name] = len(nodes) - 1
                node = ch[name]
            nodes[node]["attributes"]["VARIABLE_VALUE"] = key
        out = b""
        for n in nodes:
            body = b""
            for lname, nid in n["children"].items():
                body += _pb_bytes(1, _pb_tag(1, 0) + _varint(nid) + _pb_bytes(2, lname.encode()))
            for nm, key in n["attributes"].items():
                body += _pb_bytes(2, _pb_bytes(1, nm.encode()) + _pb_bytes(3, key.encode()))
            out += _pb_bytes(1, body)
        return out


# ----------------------------------------------------------------------------------------------
# the reference's objects -> this repository's variable names
# ----------------------------------------------------------------------------------------------
def _modconv_paths(prefix: str, base: Tuple[str, ...]) -> Dict[str, Tuple[str, ...]]:
    """ModulatedConv2D attributes (modulated_conv2d.py:52-64): w, mod_dense.w, mod_bias.b."""
    return {prefix + "/w": base + ("w",), prefix + "/mod_dense/w": base + ("mod_dense", "w"),
            prefix + "/mod_bias/b": base + ("mod_bias", "b")}


def generator_variable_paths(cfg) -> Dict[str, Tuple[str, ...]]:
    """This repository's generator variable name -> attribute path below the reference's ``Generator`` object
    (generator.py:14-17, synthesis_block.py:26-60,97-135, to_rgb.py:13-26, latent_encoder.py:20-37,
    mapping_block.py:20-33, word_encoder.py:17-37)."""
    res = cfg.generator_resolutions
    out: Dict[str, Tuple[str, ...]] = {
        "word_encoder/w0_embedding": ("word_encoder", "w0_embedding"),
        "word_encoder/w_embedding": ("word_encoder", "w_embedding"),
        "word_encoder/fc/kernel": ("word_encoder", "fc", "kernel"),
        "word_encoder/fc/bias": ("word_encoder", "fc", "bias"),
        "latent_encoder/w_avg": ("latent_encoder", "w_avg"),
    }
    t0 = f"synthesis/{res[0][0]}x{res[0][1]}/ToRGB"
    out.update(_modconv_paths(t0 + "/conv", ("synthesis", "initial_torgb", "conv")))
    out[t0 + "/bias/b"] = ("synthesis", "initial_torgb", "apply_bias", "b")
    for i, (h, w) in enumerate(res[1:]):
        pb = f"synthesis/{h}x{w}/block"
        blk = ("synthesis", "synth_blocks", str(i))
        for j in (0, 1):
            out.update(_modconv_paths(f"{pb}/conv_{j}", blk + (f"conv_{j}",)))
            out[f"{pb}/noise_{j}/w"] = blk + (f"apply_noise_{j}", "noise_strength")
            out[f"{pb}/bias_{j}/b"] = blk + (f"apply_bias_act_{j}", "b")
        tr = ("synthesis", "torgbs", str(i))
        out.update(_modconv_paths(f"synthesis/{h}x{w}/ToRGB/conv", tr + ("conv",)))
        out[f"synthesis/{h}x{w}/ToRGB/bias/b"] = tr + ("apply_bias", "b")
    for i in range(cfg.n_mapping):
        out[f"latent_encoder/g_mapping/dense_{i}/w"] = ("latent_encoder", "g_mapping", "dense_layers", str(i), "w")
        out[f"latent_encoder/g_mapping/bias_{i}/b"] = ("latent_encoder", "g_mapping", "bias_act_layers", str(i), "b")
    return out


def discriminator_variable_paths(cfg) -> Dict[str, Tuple[str, ...]]:
    """Discriminator variable name -> attribute path (discriminator.py:30-66,113-130,174-200; conv.py:41-49;
    from_rgb.py:14-24)."""
    res = cfg.discrim_resolutions
    r0 = res[0]
    out = {f"{r0[0]}x{r0[1]}/FromRGB/conv/w": ("initial_fromrgb", "conv", "w"),
           f"{r0[0]}x{r0[1]}/FromRGB/bias/b": ("initial_fromrgb", "apply_bias_act", "b")}
    for i, (h, w) in enumerate(res[:-1]):
        pb, blk = f"{h}x{w}", ("blocks", str(i))
        out[pb + "/conv_0/w"] = blk + ("conv_0", "w")
        out[pb + "/bias_0/b"] = blk + ("apply_bias_act_0", "b")
        out[pb + "/conv_1/w"] = blk + ("conv_1", "w")
        out[pb + "/bias_1/b"] = blk + ("apply_bias_act_1", "b")
        out[pb + "/skip/w"] = blk + ("conv_skip", "w")
    rf = res[-1]
    pl, last = f"{rf[0]}x{rf[1]}/last", ("last_block",)
    out[pl + "/conv_0/w"] = last + ("conv_0", "w")
    out[pl + "/bias_0/b"] = last + ("apply_bias_act_0", "b")
    out[pl + "/dense_1/w"] = last + ("dense_1", "w")
    out[pl + "/bias_1/b"] = last + ("apply_bias_act_1", "b")
    out["last_dense/w"] = ("last_dense", "w")
    out["last_bias/b"] = ("last_bias", "b")
    return out


def _load_model_state(path: str, root: str, paths: Dict[str, Tuple[str, ...]], shapes: Dict[str, Tuple[int, ...]]):
    import torch

    tensors = load_tensor_bundle(path)
    if OBJECT_GRAPH_KEY not in tensors:
        raise KeyError(f"{path}: not an object-based checkpoint (no {OBJECT_GRAPH_KEY})")
    graph = ObjectGraph(tensors[OBJECT_GRAPH_KEY])
    state = {}
    for name, rel in paths.items():
        key = graph.resolve((root,) + rel)
        arr = np.asarray(tensors[key], dtype=np.float32)
        if tuple(arr.shape) != tuple(shapes[name]):
            if int(np.prod(arr.shape)) != int(np.prod(shapes[name])):
                raise ValueError(f"{name}: checkpoint shape {arr.shape} != expected {tuple(shapes[name])}")
            arr = arr.reshape(shapes[name])       # e.g. the scalar noise strength stored as shape [] vs [1]
        state[name] = torch.from_numpy(arr.copy())
    return state


def load_generator_from_tf_checkpoint(generator, path: str, is_g_clone: bool = False) -> None:
    """Restore a :class:`~textboxgan_b200.generator.Generator` from the authors' "trained model" checkpoint
    (README.md:60-66): the ``g_clone`` (EMA) or ``generator`` object of train.py:94-108."""
    shapes = {k: tuple(v.shape) for k, v in generator.params.items()}
    state = _load_model_state(path, "g_clone" if is_g_clone else "generator", generator_variable_paths(generator.cfg), shapes)
    generator.load_state_dict(state)


def load_discriminator_from_tf_checkpoint(discriminator, path: str) -> None:
    shapes = {k: tuple(v.shape) for k, v in discriminator.params.items()}
    discriminator.load_state_dict(_load_model_state(path, "discriminator", discriminator_variable_paths(discriminator.cfg),
                                                    shapes))


def export_tf_checkpoint(prefix: str, generator=None, g_clone=None, discriminator=None) -> None:
    """Write an object-based checkpoint of the models in the reference's own layout, restorable by its
    ``ModelLoader.load_checkpoint`` (models/model_loader.py:57-81)."""
    tensors: Dict[str, object] = {}
    graph: Dict[Tuple[str, ...], str] = {}
    for root, model, paths_fn in (("generator", generator, generator_variable_paths),
                                  ("g_clone", g_clone, generator_variable_paths),
                                  ("discriminator", discriminator, discriminator_variable_paths)):
        if model is None:
            continue
        for name, rel in paths_fn(model.cfg).items():
            path = (root,) + rel
            key = "/".join(path) + VAR_SUFFIX
            tensors[key] = model.params[name].detach().cpu().numpy().astype(np.float32)
            graph[path] = key
    tensors[OBJECT_GRAPH_KEY] = ObjectGraph.build(graph)
    save_tensor_bundle(prefix, tensors)


# ----------------------------------------------------------------------------------------------
# ASTER SavedModel variables
# ----------------------------------------------------------------------------------------------
def aster_weights_from_checkpoint(path: str) -> Dict[str, "object"]:
    """Weight dictionary for :class:`~textboxgan_b200.aster_inferer.AsterInferer` from the ASTER SavedModel's
    ``variables/variables`` bundle (aster_inferer.py:24-26).  The SavedModel is not in the reference repository and its
    variable names are known only through the rename table of aster_ocr_utils/weigths_tf1_to_tf2.py:3-19, so tensors are
    matched to this repository's names by those name fragments and by shape; anything ambiguous or missing raises with
    the list of variables found (parity of the recogniser stays UNPINNED until such a file is available)."""
    import torch

    from .aster_inferer import init_aster_params

    tensors = {k: v for k, v in load_tensor_bundle(path).items() if isinstance(v, np.ndarray) and v.dtype.kind == "f"}
    want = init_aster_params(0)
    fragments = {
        "dec/attention_v": ("attention_v",), "dec/query_layer/w": ("BahdanauAttention", "query"),
        "dec/memory_layer/w": ("BahdanauAttention", "memory"), "dec/lstm_cell/w": ("Predictor/lstm_cell", "kernel"),
        "dec/lstm_cell/b": ("Predictor/lstm_cell", "bias"), "dec/dense/w": ("Predictor/dense", "kernel"),
        "dec/dense/b": ("Predictor/dense", "bias"),
    }
    out, used = {}, set()
    for name, ref in want.items():
        frs = fragments.get(name, tuple(p for p in name.split("/") if p))
        cands = [k for k, v in tensors.items() if k not in used and tuple(v.shape) == tuple(ref.shape)
                 and all(f.lower() in k.lower() for f in frs) and "backward" not in k.lower()]
        if len(cands) != 1:
            raise KeyError(f"aster weight '{name}' {tuple(ref.shape)}: {len(cands)} candidates among {len(tensors)} variables "
                           f"(first names: {sorted(tensors)[:8]})")
        used.add(cands[0])
        out[name] = torch.from_numpy(np.asarray(tensors[cands[0]], dtype=np.float32).copy())
    return out
