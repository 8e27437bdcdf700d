"""A minimal HDF5 writer, from the published format specification (superblock 0, version-1 object headers, symbol-table
groups: "TREE" / "SNOD" / "HEAP", simple dataspaces, IEEE little-endian floats, contiguous or compact layout). h5py is
not in this image, so this is what stands behind model.save(path.h5) (model_training.py:302, 346) — write_keras_weights()
lays the weights out as Keras does (/model_weights/<layer>/<layer>/<weight>:0) — and behind the reader's tests
(lisec_b200/h5weights.py). PARITY UNPINNED: the files are read back by lisec_b200's own reader; they carry no
model_config / layer_names attributes, so Keras's load_model() would not accept them, and nothing here proves
compatibility with libhdf5's output."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _Buf:
    def __init__(self):
        self.b = bytearray()

    def alloc(self, n, align=8):
        while len(self.b) % align:
            self.b.append(0)
        a = len(self.b)
        self.b.extend(b"\0" * n)
        return a

    def put(self, addr, data):
        self.b[addr:addr + len(data)] = data


def _msg(mtype, body, flags=0):
    body = bytes(body) + b"\0" * (-len(body) % 8)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(buf, msgs, split_at=None):
    """Version-1 object header; with split_at the messages from that index on go into a continuation block."""
    first = msgs if split_at is None else msgs[:split_at]
    rest = [] if split_at is None else msgs[split_at:]
    cont_addr = None
    if rest:
        blob = b"".join(rest)
        cont_addr = buf.alloc(len(blob))
        buf.put(cont_addr, blob)
        first = first + [_msg(0x0000, b"\0" * 8), _msg(0x0010, struct.pack("<QQ", cont_addr, len(blob)))]
    blob = b"".join(first)
    n = len(first) + len(rest)
    addr = buf.alloc(16 + len(blob))
    buf.put(addr, struct.pack("<BBHII4x", 1, 0, n, 1, len(blob)) + blob)
    return addr


def _dataset(buf, arr, compact=False, split=False):
    arr = np.ascontiguousarray(arr)
    raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
    space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
    if arr.dtype.kind == "f":
        size = arr.dtype.itemsize
        exp_bits, man_bits = {4: (8, 23), 8: (11, 52), 2: (5, 10)}[size]
        bits = 0x20 | (((size * 8 - 1) & 0xFF) << 8)  # mantissa normalisation = implied, sign location in byte 1
        props = struct.pack("<HHBBBBI", 0, size * 8, man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
        dtype = struct.pack("<B", 0x10 | 1) + struct.pack("<I", bits)[:3] + struct.pack("<I", size) + props
    else:
        size = arr.dtype.itemsize
        bits = 0x08 if arr.dtype.kind == "i" else 0
        dtype = struct.pack("<B", 0x10 | 0) + struct.pack("<I", bits)[:3] + struct.pack("<I", size) + struct.pack("<HH", 0, size * 8)
    if compact:
        layout = struct.pack("<BBH", 3, 0, len(raw)) + raw
    else:
        data_addr = UNDEF
        if raw:
            data_addr = buf.alloc(len(raw))
            buf.put(data_addr, raw)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, len(raw))
    msgs = [_msg(0x0001, space), _msg(0x0003, dtype, flags=1), _msg(0x0008, layout)]
    return _object_header(buf, msgs, split_at=2 if split else None)


def _group(buf, entries, snod_entries, tree_children):
    """entries {name: object header address} -> address of the group's object header."""
    names = sorted(entries, key=lambda s: s.encode())
    heap_data = bytearray(b"\0" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        e = n.encode() + b"\0"
        heap_data.extend(e + b"\0" * (-len(e) % 8))
    seg = buf.alloc(len(heap_data))
    buf.put(seg, heap_data)
    heap = buf.alloc(32)
    buf.put(heap, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, seg))
    # leaves
    nodes = []  # (address, heap offset of the largest name below)
    for i in range(0, max(len(names), 1), snod_entries):
        chunk = names[i:i + snod_entries]
        a = buf.alloc(8 + 40 * snod_entries)
        body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
        for n in chunk:
            body += struct.pack("<QQII16x", name_off[n], entries[n], 0, 0)
        buf.put(a, body)
        nodes.append((a, name_off[chunk[-1]] if chunk else 0))
    level = 0
    while True:
        parents = []
        for i in range(0, len(nodes), tree_children):
            chunk = nodes[i:i + tree_children]
            a = buf.alloc(24 + 16 * tree_children + 8)
            body = b"TREE" + struct.pack("<BBHQQ", 0, level, len(chunk), UNDEF, UNDEF) + struct.pack("<Q", 0)
            for child, key in chunk:
                body += struct.pack("<QQ", child, key)
            buf.put(a, body)
            parents.append((a, chunk[-1][1]))
        nodes, level = parents, level + 1
        if len(nodes) == 1:
            break
    return _object_header(buf, [_msg(0x0011, struct.pack("<QQ", nodes[0][0], heap))]), nodes[0][0], heap


def write_h5(path, datasets, snod_entries=8, tree_children=32, compact=(), split=(), superblock_version=0):
    """datasets: {"a/b/c": ndarray}. compact / split: paths stored with compact layout / with the layout message in a
    continuation block."""
    buf = _Buf()
    sb_size = 24 + (4 if superblock_version == 1 else 0) + 32 + 40
    buf.alloc(sb_size)
    tree = {}
    for p, arr in datasets.items():
        node = tree
        parts = p.split("/")
        for q in parts[:-1]:
            node = node.setdefault(q, {})
        node[parts[-1]] = (p, np.asarray(arr))

    def build(node):
        entries = {}
        for name, val in node.items():
            if isinstance(val, dict):
                entries[name] = build(val)[0]
            else:
                p, arr = val
                entries[name] = _dataset(buf, arr, compact=p in compact, split=p in split)
        return _group(buf, entries, snod_entries, tree_children)

    root_header, root_btree, root_heap = build(tree)
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", superblock_version, 0, 0, 0, 0, 8, 8, 0)
    sb += struct.pack("<HHI", 4, 16, 0)
    if superblock_version == 1:
        sb += struct.pack("<HH", 32, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(buf.b), UNDEF)
    sb += struct.pack("<QQII", 0, root_header, 1, 0) + struct.pack("<QQ", root_btree, root_heap)
    buf.put(0, sb)
    with open(path, "wb") as f:
        f.write(bytes(buf.b))


def write_keras_weights(path, pack):
    """pack: {"dense/kernel": array, "batch_normalization_3/moving_mean": array, ...} (lisec_b200/weights.py names) ->
    an .h5 with the datasets where Keras's model.save() puts them."""
    write_h5(path, {"model_weights/%s/%s:0" % (k.split("/")[0], k): np.asarray(v, dtype=np.float32) for k, v in pack.items()})
