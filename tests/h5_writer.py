"""A minimal HDF5 *writer* for the tests of keras_ocr_b200.hdf5 (the image has no h5py): superblock version 0,
old-style groups (symbol table message -> v1 B-tree -> SNOD nodes + local heap) and float32 / float64 / int32 datasets,
contiguous or chunked with shuffle + deflate -- the structures h5py's default ``libver='earliest'`` produces for a
Keras ``save_weights`` file.  Written from the HDF5 File Format Specification, like the reader, but independently: the
reader is additionally checked against a file produced by the HDF5 C library (tests/test_hdf5.py)."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 4                                    # a symbol table node holds up to 2 * LEAF_K entries


class _Out:
    def __init__(self):
        self.buf = bytearray()

    def tell(self):
        return len(self.buf)

    def align(self, n=8):
        self.buf += bytes(-len(self.buf) % n)

    def put(self, data):
        self.align()
        pos = len(self.buf)
        self.buf += data
        return pos


def _msg(mtype, body):
    body += bytes(-len(body) % 8)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _header(messages):
    blob = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(blob)) + blob


def _datatype(dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        exp, mant, bias = {4: (8, 23, 127), 8: (11, 52, 1023)}[dtype.itemsize]
        bits = dtype.itemsize * 8
        return (struct.pack("<BBBBI", 0x11, 0x20, bits - 1, 0, dtype.itemsize)
                + struct.pack("<HHBBBBI", 0, bits, mant, exp, 0, mant, bias))
    return struct.pack("<BBBBI", 0x10, 0x08 if dtype.kind == "i" else 0, 0, 0, dtype.itemsize) + struct.pack("<HH", 0, dtype.itemsize * 8)


def _dataset(out, arr, chunks=None):
    arr = np.asarray(arr)
    arr = arr if arr.ndim == 0 else np.ascontiguousarray(arr)
    space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
    msgs = [_msg(0x0001, space), _msg(0x0003, _datatype(arr.dtype))]
    if chunks is None:
        data_at = out.put(arr.tobytes())
        msgs.append(_msg(0x0008, struct.pack("<BBQQ", 3, 1, data_at, arr.nbytes)))
    else:
        item = arr.dtype.itemsize
        entries = []
        grid = [range(0, s, c) for s, c in zip(arr.shape, chunks)]
        for offs in np.array(np.meshgrid(*grid, indexing="ij")).reshape(arr.ndim, -1).T:
            block = np.zeros(chunks, arr.dtype)
            region = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, arr.shape))
            part = arr[region]
            block[tuple(slice(0, n) for n in part.shape)] = part
            raw = np.frombuffer(block.tobytes(), np.uint8).reshape(-1, item).T.tobytes()     # shuffle
            raw = zlib.compress(raw, 4)                                                        # deflate
            entries.append((tuple(int(o) for o in offs), out.put(raw), len(raw)))
        key = lambda size, offs: struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)
        node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF)
        for offs, at, size in entries:
            node += key(size, offs) + struct.pack("<Q", at)
        node += key(0, tuple(s for s in arr.shape))
        tree_at = out.put(node)
        layout = struct.pack("<BBB", 3, 2, arr.ndim + 1) + struct.pack("<Q", tree_at)
        layout += b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", item)
        msgs.append(_msg(0x0008, layout))
        # filter pipeline v1: shuffle (id 2, one client value = element size) then deflate (id 1, level)
        pipe = struct.pack("<BB6x", 1, 2)
        pipe += struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<I", item) + bytes(4)
        pipe += struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<I", 4) + bytes(4)
        msgs.append(_msg(0x000B, pipe))
    return out.put(_header(msgs))


def _group(out, tree, chunked):
    """tree: {name: ndarray | dict}.  Returns the address of the group's object header."""
    children = {}
    for name, value in tree.items():
        if isinstance(value, dict):
            children[name] = _group(out, value, chunked)
        else:
            children[name] = _dataset(out, value, chunked.get(id(value)))
    names = sorted(children)                               # B-tree order = byte order of the names
    heap = bytearray(b"\x00" * 8)                          # offset 0 = the empty string
    name_off = {}
    for name in names:
        name_off[name] = len(heap)
        heap += name.encode() + b"\x00"
        heap += bytes(-len(heap) % 8)
    heap_data_at = out.put(bytes(heap))
    heap_at = out.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), UNDEF, heap_data_at))
    nodes = []
    for i in range(0, max(len(names), 1), 2 * LEAF_K):
        part = names[i:i + 2 * LEAF_K]
        node = b"SNOD" + struct.pack("<BxH", 1, len(part))
        for name in part:
            node += struct.pack("<QQII16x", name_off[name], children[name], 0, 0)
        node += bytes((2 * LEAF_K - len(part)) * 40)
        nodes.append((out.put(node), name_off[part[-1]] if part else 0))
    tree_node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(nodes), UNDEF, UNDEF) + struct.pack("<Q", 0)
    for at, last in nodes:
        tree_node += struct.pack("<QQ", at, last)
    tree_at = out.put(tree_node)
    return out.put(_header([_msg(0x0011, struct.pack("<QQ", tree_at, heap_at))]))


def write(path, tree, chunked=None, userblock=0):
    """tree: nested dict of ndarrays.  ``chunked``: {id(array): chunk shape} for arrays to store chunked + filtered.
    ``userblock``: bytes of zeros in front of the superblock (512, 1024, ...)."""
    out = _Out()
    out.buf += bytes(96)                                   # superblock, patched below
    root = _group(out, tree, chunked or {})
    out.align()
    sb = SIGNATURE = b"\x89HDF\r\n\x1a\n"
    sb += struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, 16, 0)
    sb += struct.pack("<QQQQ", userblock, UNDEF, len(out.buf), UNDEF)
    sb += struct.pack("<QQII16x", 0, root, 0, 0)
    assert len(sb) == 96
    out.buf[:96] = sb
    with open(path, "wb") as f:
        f.write(bytes(userblock) + bytes(out.buf))
