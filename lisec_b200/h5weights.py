"""Keras .h5 weight ingestion without h5py — load_model(path) / model.save(path) as the reference uses them
(Predict.py:51-52, model_training.py:302, 337-338; north_star: "same SampleModel .h5 weights").

h5py / libhdf5 are not in this image, so this is a small reader of the HDF5 file format itself, restricted to what
`model.save()` of Keras 2.x / tf.keras 2.0-2.x writes through h5py's defaults ("earliest" library version bounds):

    superblock version 0 or 1; groups as symbol tables (version-1 B-tree "TREE" nodes -> "SNOD" symbol-table nodes ->
    names in a local "HEAP"); version-1 object headers with continuation blocks; datasets with a simple dataspace, an
    IEEE little-endian float32 / float64 (or fixed-point) datatype and CONTIGUOUS or COMPACT layout (Keras neither
    chunks nor compresses its weights).

Keras stores a model as  /model_weights/<layer name>/<weight name>  with weight names like "dense/kernel:0",
"batch_normalization_3/moving_mean:0" (so the path is /model_weights/dense/dense/kernel:0); `model.save_weights()` files
have the layer groups at the root. read_keras_weights() walks every group, takes every dataset below model_weights (or
the root) and keys it by its weight name without the ":0" — the names of lisec_b200/weights.py.

Anything else (superblock 2/3, version-2 "OHDR" object headers, chunked / filtered datasets, big-endian data) raises
H5FormatError with the reason. STATUS: written from the published format specification (HDF5 File Format Specification
version 1.1/2.0); there is not one HDF5 file in this image and the reference's SampleModel/*.h5 blobs are absent
(.MISSING_LARGE_BLOBS), so the reader has only been exercised against files produced by the independent minimal writer in
tests/h5_writer.py — PARITY UNPINNED until it has met a file written by h5py.
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


class _File:
    def __init__(self, data: bytes):
        self.d = data
        base = -1
        off = 0
        while off < len(data):  # the superblock sits at 0, 512, 1024, ...
            if data[off:off + 8] == SIGNATURE:
                base = off
                break
            off = 512 if off == 0 else off * 2
        if base < 0:
            raise H5FormatError("not an HDF5 file (no superblock signature)")
        ver = data[base + 8]
        if ver not in (0, 1):
            raise H5FormatError("superblock version %d: only the version 0/1 layout h5py writes by default is supported" % ver)
        self.so, self.sl = data[base + 13], data[base + 14]  # size of offsets / lengths
        if self.so != 8 or self.sl != 8:
            raise H5FormatError("offsets/lengths of %d/%d bytes: only 8/8 is supported" % (self.so, self.sl))
        p = base + 16
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", data, p)
        p += 4 + 4  # K values, file consistency flags
        if ver == 1:
            p += 4  # indexed storage internal node K + reserved
        self.base_addr, _, self.eof, _ = struct.unpack_from("<QQQQ", data, p)
        p += 32
        # root group symbol table entry
        _, self.root_header, cache, _ = struct.unpack_from("<QQII", data, p)
        self.root_scratch = struct.unpack_from("<QQ", data, p + 24) if cache == 1 else None

    def u(self, fmt: str, addr: int):
        return struct.unpack_from("<" + fmt, self.d, self.base_addr + addr)

    # ---- object headers (version 1) ----------------------------------------------------------------------------------
    def messages(self, addr: int):
        """[(type, flags, payload bytes)] of the version-1 object header at addr, continuation blocks followed."""
        if self.d[self.base_addr + addr:self.base_addr + addr + 4] == b"OHDR":
            raise H5FormatError("version-2 object header at %#x (file written with libver='latest'): not supported" % addr)
        ver, _, nmsg, _, size = self.u("BBHII", addr)
        if ver != 1:
            raise H5FormatError("object header version %d at %#x" % (ver, addr))
        out, blocks = [], [(addr + 16, size)]  # the first message is 8-byte aligned behind the 12-byte prefix
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self.u("HHB", p)
                body = self.d[self.base_addr + p + 8:self.base_addr + p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:  # continuation: offset, length
                    blocks.append(struct.unpack("<QQ", body[:16]))
                out.append((mtype, flags, body))
        return out

    # ---- groups (symbol tables) ----------------------------------------------------------------------------------------
    def _heap_name(self, heap_addr: int, off: int) -> str:
        if self.d[self.base_addr + heap_addr:self.base_addr + heap_addr + 4] != b"HEAP":
            raise H5FormatError("local heap signature missing at %#x" % heap_addr)
        (seg,) = self.u("Q", heap_addr + 24)
        a = self.base_addr + seg + off
        return self.d[a:self.d.index(b"\0", a)].decode("utf-8")

    def _btree_entries(self, node: int, heap: int, out: Dict[str, int]):
        sig = self.d[self.base_addr + node:self.base_addr + node + 4]
        if sig == b"TREE":
            ntype, level, used = self.u("BBH", node + 4)
            if ntype != 0:
                raise H5FormatError("B-tree node type %d inside a group" % ntype)
            p = node + 24  # signature, type, level, entries, left and right sibling
            for i in range(used):
                (child,) = self.u("Q", p + 8)  # key i (8), child i (8), ..., key `used`
                self._btree_entries(child, heap, out)
                p += 16
        elif sig == b"SNOD":
            _, _, nsym = self.u("BBH", node + 4)
            p = node + 8
            for _ in range(nsym):
                name_off, header = self.u("QQ", p)
                out[self._heap_name(heap, name_off)] = header
                p += 40
        else:
            raise H5FormatError("neither a B-tree node nor a symbol-table node at %#x" % node)

    def children(self, header_addr: int):
        """{name: object header address} of a group, or None when the object is not an old-style group."""
        for mtype, _, body in self.messages(header_addr):
            if mtype == 0x0011:
                btree, heap = struct.unpack("<QQ", body[:16])
                out: Dict[str, int] = {}
                self._btree_entries(btree, heap, out)
                return out
            if mtype in (0x0002, 0x0006):
                raise H5FormatError("new-style group (link messages) at %#x: not supported" % header_addr)
        return None

    # ---- datasets ------------------------------------------------------------------------------------------------------
    def dataset(self, header_addr: int):
        shape = dtype = layout = None
        for mtype, _, body in self.messages(header_addr):
            if mtype == 0x0001:  # dataspace
                ver, rank, flags = body[0], body[1], body[2]
                p = 8 if ver == 1 else 4
                shape = struct.unpack_from("<%dQ" % rank, body, p) if rank else ()
            elif mtype == 0x0003:  # datatype
                cls, bits0 = body[0] & 0x0F, body[1]
                (size,) = struct.unpack_from("<I", body, 4)
                if bits0 & 1:
                    raise H5FormatError("big-endian dataset")
                if cls == 1:
                    dtype = {2: np.float16, 4: np.float32, 8: np.float64}.get(size)
                elif cls == 0:
                    signed = bool(bits0 & 0x08)
                    dtype = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}.get(size) if signed else \
                        {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}.get(size)
                if dtype is None:
                    dtype = ("unsupported", cls, size)
            elif mtype == 0x0008:  # data layout
                ver, lclass = body[0], body[1]
                if ver != 3:
                    raise H5FormatError("data layout message version %d" % ver)
                if lclass == 1:
                    layout = ("contiguous",) + struct.unpack_from("<QQ", body, 2)
                elif lclass == 0:
                    (n,) = struct.unpack_from("<H", body, 2)
                    layout = ("compact", body[4:4 + n])
                else:
                    raise H5FormatError("chunked dataset (Keras does not chunk weights): not supported")
            elif mtype == 0x000B:
                raise H5FormatError("filtered (compressed) dataset: not supported")
        if shape is None or dtype is None or layout is None:
            return None
        if isinstance(dtype, tuple):
            raise H5FormatError("datatype class %d of %d bytes" % dtype[1:])
        count = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            raw = layout[1]
        else:
            addr, size = layout[1], layout[2]
            raw = b"" if addr == UNDEF else self.d[self.base_addr + addr:self.base_addr + addr + size]
        need = count * np.dtype(dtype).itemsize
        if len(raw) < need:
            raise H5FormatError("dataset data truncated (%d of %d bytes)" % (len(raw), need))
        return np.frombuffer(raw[:need], dtype=np.dtype(dtype).newbyteorder("<")).reshape(shape).astype(dtype)

    def walk(self, header_addr: int, prefix: str, out: Dict[str, np.ndarray], depth: int = 0):
        if depth > 16:
            raise H5FormatError("group nesting deeper than 16 (cycle?)")
        kids = self.children(header_addr)
        if kids is None:
            arr = self.dataset(header_addr)
            if arr is not None:
                out[prefix] = arr
            return
        for name, addr in kids.items():
            self.walk(addr, prefix + "/" + name if prefix else name, out, depth + 1)


def read_datasets(path: str) -> Dict[str, np.ndarray]:
    """Every dataset of the file, keyed by its path ("model_weights/dense/dense/kernel:0")."""
    with open(path, "rb") as f:
        h5 = _File(f.read())
    out: Dict[str, np.ndarray] = {}
    h5.walk(h5.root_header, "", out)
    return out


def read_keras_weights(path: str) -> Dict[str, np.ndarray]:
    """Keras-named weight pack of a model.save() / save_weights() file: {"dense/kernel": ..., "batch_normalization/gamma":
    ..., ...} — optimizer state (optimizer_weights/...) is skipped."""
    pack: Dict[str, np.ndarray] = {}
    for p, arr in read_datasets(path).items():
        parts = p.split("/")
        if parts[0] == "optimizer_weights":
            continue
        if parts[0] == "model_weights":
            parts = parts[1:]
        if len(parts) < 2:
            continue
        name = "/".join(parts[1:])  # drop the layer group; the weight name carries the layer's name again
        if name.endswith(":0"):
            name = name[:-2]
        pack[name] = arr
    if not pack:
        raise H5FormatError("no Keras weights found in %s" % path)
    return pack
