"""A dependency-free reader for the subset of HDF5 that Keras weight files use.

The reference loads ``crnn_kurapan.h5`` with ``model.load_weights`` (reference recognition.py:386-392), i.e. through
h5py, which this offline image does not have.  A Keras ``save_weights`` file is a tree of groups (one per layer) whose
leaves are small float32 datasets; this module walks that tree straight from the bytes of the file, following the HDF5
File Format Specification (version 3.0):

* superblock versions 0-3 (a user block of 512 * 2^k bytes in front is skipped, as h5py does);
* object headers version 1 and version 2 (``OHDR`` / ``OCHK``), continuation messages;
* old-style groups (symbol table message -> v1 B-tree -> ``SNOD`` nodes + local heap) and new-style groups with
  compact storage (link messages); dense link storage (fractal heaps) is NOT read and raises ``Hdf5Error``;
* datasets with compact, contiguous or chunked (v1 B-tree index) layout, fixed-point and IEEE floating-point types of
  either byte order, filters deflate / shuffle / fletcher32;
* attributes are ignored -- tensors are mapped by their path (``weights.map_keras_datasets``), not by the order in the
  ``layer_names`` / ``weight_names`` attributes.

    tensors = hdf5.read_datasets("crnn_kurapan.h5")   # {"conv_1/conv_1/kernel:0": ndarray, ...}
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEFINED = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(ValueError):
    pass


class _File:
    def __init__(self, data):
        self.data = memoryview(data)
        self.base = self.superblock = self._find_superblock()
        self._read_superblock()

    # ------------------------------------------------------------------ primitives
    def u(self, pos, size):
        return int.from_bytes(self.data[pos:pos + size], "little")

    def offset(self, pos):
        return self.u(pos, self.so)

    def length(self, pos):
        return self.u(pos, self.sl)

    def addr(self, value):
        """File address -> position in the buffer (addresses are relative to the base address)."""
        if value >= (1 << (8 * self.so)) - 1:
            raise Hdf5Error("undefined address")
        return self.base + value

    def _find_superblock(self):
        pos = 0
        while pos + 8 <= len(self.data):
            if bytes(self.data[pos:pos + 8]) == SIGNATURE:
                return pos
            pos = 512 if pos == 0 else pos * 2
        raise Hdf5Error("not an HDF5 file (no superblock signature)")

    def _read_superblock(self):
        p = self.base
        version = self.data[p + 8]
        if version in (0, 1):
            self.so, self.sl = self.data[p + 13], self.data[p + 14]
            q = p + 24 + (4 if version == 1 else 0)
            self.base = self.offset(q)                           # every address in the file is relative to this one
            q += 4 * self.so                                     # base, free-space info, end of file, driver info
            # root group symbol table entry: name offset, object header address, cache type, reserved, scratch pad
            self.root_header = self.offset(q + self.so)
        elif version in (2, 3):
            self.so, self.sl = self.data[p + 9], self.data[p + 10]
            self.base = self.offset(p + 12)
            self.root_header = self.offset(p + 12 + 3 * self.so)
        else:
            raise Hdf5Error(f"unsupported superblock version {version}")
        if self.so not in (4, 8) or self.sl not in (4, 8):
            raise Hdf5Error("unsupported offset / length size")

    # ------------------------------------------------------------------ object headers
    def messages(self, address):
        """All (type, flags, body position, body size) header messages of the object at ``address``."""
        p = self.addr(address)
        if bytes(self.data[p:p + 4]) == b"OHDR":
            return self._messages_v2(p)
        return self._messages_v1(p)

    def _messages_v1(self, p):
        if self.data[p] != 1:
            raise Hdf5Error(f"unsupported object header version {self.data[p]}")
        total = self.u(p + 2, 2)
        blocks = [(p + 16, self.u(p + 8, 4))]                     # first block follows the 16-byte (padded) prefix
        out = []
        while blocks and len(out) < total:
            q, size = blocks.pop(0)
            end = q + size
            while q + 8 <= end and len(out) < total:
                mtype, msize, flags = self.u(q, 2), self.u(q + 2, 2), self.data[q + 4]
                body = q + 8
                if mtype == 0x0010:                              # continuation
                    blocks.append((self.addr(self.offset(body)), self.length(body + self.so)))
                out.append((mtype, flags, body, msize))
                q = body + msize
        return out

    def _messages_v2(self, p):
        flags = self.data[p + 5]
        q = p + 6
        if flags & 0x20:
            q += 16                                              # access / modification / change / birth times
        if flags & 0x10:
            q += 4                                               # max compact / min dense attributes
        width = 1 << (flags & 3)
        chunk0 = self.u(q, width)
        q += width
        track_order = bool(flags & 0x04)
        blocks, out = [(q, chunk0)], []
        while blocks:
            q, size = blocks.pop(0)
            end = q + size
            head = 4 + (2 if track_order else 0)
            while q + head <= end:
                mtype, msize, mflags = self.data[q], self.u(q + 1, 2), self.data[q + 3]
                body = q + head
                if body + msize > end:
                    break
                if mtype == 0x10:
                    cont, clen = self.addr(self.offset(body)), self.length(body + self.so)
                    if bytes(self.data[cont:cont + 4]) != b"OCHK":
                        raise Hdf5Error("bad object header continuation block")
                    blocks.append((cont + 4, clen - 8))          # minus signature and checksum
                out.append((mtype, mflags, body, msize))
                q = body + msize
        return out

    # ------------------------------------------------------------------ groups
    def children(self, address):
        """{name: object header address} of a group; None if the object is not a group."""
        msgs = self.messages(address)
        links = {}
        is_group = False
        for mtype, _flags, body, size in msgs:
            if mtype == 0x0011:                                  # symbol table: v1 B-tree + local heap
                is_group = True
                self._walk_group_btree(self.offset(body), self.offset(body + self.so), links)
            elif mtype == 0x0006:                                # link message (compact new-style group)
                is_group = True
                name, target = self._link(body)
                if target is not None:
                    links[name] = target
            elif mtype == 0x0002:                                # link info
                is_group = True
                fractal = self.offset(body + 2 + (8 if self.data[body + 1] & 1 else 0))
                if fractal < (1 << (8 * self.so)) - 1:
                    raise Hdf5Error("dense link storage (fractal heap) is not supported; re-save the file with "
                                    "h5py's default libver='earliest'")
        return links if is_group else None

    def _link(self, body):
        flags = self.data[body + 1]
        q = body + 2
        ltype = 0
        if flags & 0x08:
            ltype = self.data[q]; q += 1
        if flags & 0x04:
            q += 8
        if flags & 0x10:
            q += 1
        width = 1 << (flags & 3)
        nlen = self.u(q, width); q += width
        name = bytes(self.data[q:q + nlen]).decode("utf-8"); q += nlen
        return name, (self.offset(q) if ltype == 0 else None)   # soft / external links are skipped

    def _heap_string(self, heap_address, off):
        p = self.addr(heap_address)
        if bytes(self.data[p:p + 4]) != b"HEAP":
            raise Hdf5Error("bad local heap")
        segment = self.addr(self.offset(p + 8 + 2 * self.sl))
        q = segment + off
        end = q
        while self.data[end] != 0:
            end += 1
        return bytes(self.data[q:end]).decode("utf-8")

    def _walk_group_btree(self, btree_address, heap_address, links):
        p = self.addr(btree_address)
        if bytes(self.data[p:p + 4]) != b"TREE" or self.data[p + 4] != 0:
            raise Hdf5Error("bad group B-tree node")
        level, used = self.data[p + 5], self.u(p + 6, 2)
        q = p + 8 + 2 * self.so                                  # past the sibling pointers
        for i in range(used):
            child = self.offset(q + self.sl + i * (self.sl + self.so))      # key, child, key, child, ..., key
            if level > 0:
                self._walk_group_btree(child, heap_address, links)
                continue
            s = self.addr(child)
            if bytes(self.data[s:s + 4]) != b"SNOD":
                raise Hdf5Error("bad symbol table node")
            count = self.u(s + 6, 2)
            entry = 2 * self.so + 24
            for k in range(count):
                e = s + 8 + k * entry
                links[self._heap_string(heap_address, self.offset(e))] = self.offset(e + self.so)

    # ------------------------------------------------------------------ datasets
    def dataset(self, address):
        """ndarray of the dataset at ``address``, or None if the object is not a dataset."""
        shape = dtype = layout = None
        filters = []
        for mtype, _flags, body, size in self.messages(address):
            if mtype == 0x0001:
                shape = self._dataspace(body)
            elif mtype == 0x0003:
                dtype = self._datatype(body)
            elif mtype == 0x0008:
                layout = (body, size)
            elif mtype == 0x000B:
                filters = self._filters(body)
        if shape is None or dtype is None or layout is None:
            return None
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        raw = self._read_layout(layout[0], shape, dtype, filters, count)
        return np.frombuffer(raw, dtype=dtype, count=count).reshape(shape).copy()

    def _dataspace(self, body):
        version, rank = self.data[body], self.data[body + 1]
        if version == 1:
            q = body + 8
        elif version == 2:
            if self.data[body + 3] == 2:
                raise Hdf5Error("null dataspace")
            q = body + 4
        else:
            raise Hdf5Error(f"unsupported dataspace version {version}")
        return tuple(self.length(q + i * self.sl) for i in range(rank))

    def _datatype(self, body):
        cls, bits0, size = self.data[body] & 0x0F, self.data[body + 1], self.u(body + 4, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 1:                                             # IEEE floating point
            if size not in (2, 4, 8):
                raise Hdf5Error(f"unsupported float size {size}")
            return np.dtype(f"{order}f{size}")
        if cls == 0:                                             # fixed point
            signed = "i" if bits0 & 0x08 else "u"
            if size not in (1, 2, 4, 8):
                raise Hdf5Error(f"unsupported integer size {size}")
            return np.dtype(f"{order}{signed}{size}")
        raise Hdf5Error(f"unsupported datatype class {cls} (only numeric datasets are read)")

    def _filters(self, body):
        version, n = self.data[body], self.data[body + 1]
        q = body + (8 if version == 1 else 2)
        out = []
        for _ in range(n):
            fid = self.u(q, 2); q += 2
            nlen = 0
            if version == 1 or fid >= 256:
                nlen = self.u(q, 2); q += 2
            q += 2                                               # flags
            ncd = self.u(q, 2); q += 2
            q += (nlen + 7) // 8 * 8 if version == 1 else nlen
            cd = [self.u(q + 4 * i, 4) for i in range(ncd)]
            q += 4 * ncd
            if version == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    def _read_layout(self, body, shape, dtype, filters, count):
        version = self.data[body]
        nbytes = count * dtype.itemsize
        if version in (1, 2):
            rank, cls = self.data[body + 1], self.data[body + 2]
            q = body + 8
            address = None
            if cls != 0:
                address = self.offset(q); q += self.so
            dims = [self.u(q + 4 * i, 4) for i in range(rank)]
            q += 4 * rank
            if cls == 0:
                size = self.u(q, 4)
                return bytes(self.data[q + 4:q + 4 + size])
            if cls == 1:
                return self._contiguous(address, nbytes)
            return self._chunked(address, dims, shape, dtype, filters)
        if version == 3:
            cls = self.data[body + 1]
            if cls == 0:
                size = self.u(body + 2, 2)
                return bytes(self.data[body + 4:body + 4 + size])
            if cls == 1:
                return self._contiguous(self.offset(body + 2), nbytes)
            if cls == 2:
                rank = self.data[body + 2]
                address = self.offset(body + 3)
                dims = [self.u(body + 3 + self.so + 4 * i, 4) for i in range(rank)]
                return self._chunked(address, dims, shape, dtype, filters)
        raise Hdf5Error(f"unsupported data layout (version {version})")

    def _contiguous(self, address, nbytes):
        if address >= (1 << (8 * self.so)) - 1:                  # never written: fill value (zeros)
            return bytes(nbytes)
        p = self.addr(address)
        return bytes(self.data[p:p + nbytes])

    def _chunked(self, btree_address, dims, shape, dtype, filters):
        chunk = tuple(dims[:-1])                                 # the last "dimension" is the element size
        if len(chunk) != len(shape):
            raise Hdf5Error("chunk rank does not match the dataspace")
        out = np.zeros(shape, dtype=dtype)
        if btree_address >= (1 << (8 * self.so)) - 1:
            return out.tobytes()
        for offsets, address, size, mask in self._chunk_leaves(btree_address, len(shape)):
            raw = bytes(self.data[self.addr(address):self.addr(address) + size])
            for index in range(len(filters) - 1, -1, -1):        # undo the pipeline in reverse order
                if mask & (1 << index):
                    continue
                fid, _cd = filters[index]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    item = dtype.itemsize
                    raw = np.frombuffer(raw, np.uint8).reshape(item, -1).T.tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise Hdf5Error(f"unsupported filter {fid}")
            block = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk))).reshape(chunk)
            region = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offsets, chunk, shape))
            out[region] = block[tuple(slice(0, r.stop - r.start) for r in region)]
        return out.tobytes()

    def _chunk_leaves(self, address, rank):
        p = self.addr(address)
        if bytes(self.data[p:p + 4]) != b"TREE" or self.data[p + 4] != 1:
            raise Hdf5Error("bad chunk B-tree node")
        level, used = self.data[p + 5], self.u(p + 6, 2)
        q = p + 8 + 2 * self.so
        key = 8 + 8 * (rank + 1)
        for i in range(used):
            k = q + i * (key + self.so)
            size, mask = self.u(k, 4), self.u(k + 4, 4)
            offsets = tuple(self.u(k + 8 + 8 * d, 8) for d in range(rank))
            child = self.offset(k + key)
            if level > 0:
                yield from self._chunk_leaves(child, rank)
            else:
                yield offsets, child, size, mask


def read_datasets(path_or_bytes):
    """{"group/sub/dataset": ndarray} for every numeric dataset of the file (depth-first, names in B-tree order)."""
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        data = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            data = f.read()
    f = _File(data)
    out, seen = {}, set()

    def visit(address, prefix):
        if address in seen:                                      # hard links may form cycles
            return
        seen.add(address)
        kids = f.children(address)
        if kids is not None:
            for name, target in kids.items():
                visit(target, f"{prefix}/{name}" if prefix else name)
            return
        try:
            arr = f.dataset(address)
        except Hdf5Error as exc:
            if "datatype class" in str(exc):                     # strings, references, compounds: not weights
                return
            raise
        if arr is not None:
            out[prefix] = arr

    visit(f.root_header, "")
    return out
