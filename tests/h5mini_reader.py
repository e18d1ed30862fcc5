"""A small, independent HDF5 reader — TEST INFRASTRUCTURE ONLY.

Written from the HDF5 File Format Specification (version 0 superblock, version 1 object headers, symbol-table groups with
their local heap / v1 B-tree / SNOD nodes, contiguous and compact layouts, fixed-point and IEEE floating-point datatypes).
It shares no code with the writer in voronoirt_b200/csrc/outfile.cu and is itself pinned against a file written by the real
HDF5 library (tests/golden/libhdf5_sample.mat: a MATLAB v7.3 file, i.e. a 512-byte user block followed by a libhdf5 file).
The image has neither libhdf5 nor h5py, so this reader is what checks the output files of vrt_output_* (SURVEY §8 f4:
create_output_file / write_to_file, io.jl:159-225).
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


class H5Error(ValueError):
    pass


class File:
    def __init__(self, path):
        self.buf = open(path, "rb").read()
        # the superblock may sit at 0, 512, 1024, ... (user block)
        off = 0
        while True:
            if self.buf[off:off + 8] == SIG:
                break
            off = 512 if off == 0 else off * 2
            if off >= len(self.buf):
                raise H5Error("no HDF5 signature")
        self.sb_off = off
        b = self.buf
        ver = b[off + 8]
        if ver != 0:
            raise H5Error(f"superblock version {ver} not supported by this reader")
        if b[off + 13] != 8 or b[off + 14] != 8:
            raise H5Error("only 8-byte offsets and lengths are supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, off + 16)
        self.base, self.freespace, self.eof, self.driver = struct.unpack_from("<4Q", b, off + 24)
        if self.eof > len(b):     # (libhdf5 stores the end-of-file address including the user block)
            raise H5Error(f"end-of-file address {self.eof} beyond the file ({len(b)} bytes)")
        name_off, ohdr, cache, _ = struct.unpack_from("<QQII", b, off + 56)
        self.root_header = ohdr
        if cache == 1:
            self.root_btree, self.root_heap = struct.unpack_from("<QQ", b, off + 80)
        else:
            self.root_btree = self.root_heap = None
        self.root = Group(self, ohdr)

    def at(self, addr):
        return self.base + addr

    def __getitem__(self, name):
        return self.root[name]

    def keys(self):
        return self.root.keys()


def _messages(f, addr):
    """version 1 object header -> list of (type, flags, payload bytes), continuation blocks followed"""
    b = f.buf
    p = f.at(addr)
    ver, _, nmsg, refcnt, hsize = struct.unpack_from("<BBHII", b, p)
    if ver != 1:
        raise H5Error(f"object header version {ver} at {addr}")
    blocks = [(p + 16, hsize)]
    out = []
    while blocks and len(out) < nmsg:
        q, size = blocks.pop(0)
        end = q + size
        while q + 8 <= end and len(out) < nmsg:
            mtype, msize, flags = struct.unpack_from("<HHB", b, q)
            data = b[q + 8:q + 8 + msize]
            if msize % 8:
                raise H5Error("message size not a multiple of 8 in a version 1 header")
            out.append((mtype, flags, data))
            if mtype == 0x0010:     # continuation
                caddr, clen = struct.unpack_from("<QQ", data, 0)
                blocks.append((f.at(caddr), clen))
            q += 8 + msize
    if len(out) != nmsg:
        raise H5Error(f"object header at {addr}: {len(out)} of {nmsg} messages found")
    return out


class Group:
    def __init__(self, f, header_addr):
        self.f = f
        self.header = header_addr
        msgs = _messages(f, header_addr)
        st = [m for m in msgs if m[0] == 0x0011]
        if not st:
            raise H5Error("group without a symbol table message")
        self.btree, self.heap = struct.unpack_from("<QQ", st[0][2], 0)
        self.entries = {}
        self._read_heap()
        self._walk(self.btree)

    def _read_heap(self):
        b, p = self.f.buf, self.f.at(self.heap)
        if b[p:p + 4] != b"HEAP":
            raise H5Error("local heap signature")
        self.heap_size, self.heap_free, self.heap_data = struct.unpack_from("<QQQ", b, p + 8)

    def _name(self, off):
        b, p = self.f.buf, self.f.at(self.heap_data) + off
        if off >= self.heap_size:
            raise H5Error("link name offset outside the heap")
        e = b.index(b"\0", p)
        return b[p:e].decode()

    def _walk(self, addr):
        b, p = self.f.buf, self.f.at(addr)
        if b[p:p + 4] != b"TREE":
            raise H5Error("B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, p + 4)
        if ntype != 0:
            raise H5Error("not a group B-tree")
        q = p + 24
        keys, kids = [], []
        for i in range(used):
            keys.append(struct.unpack_from("<Q", b, q)[0])
            kids.append(struct.unpack_from("<Q", b, q + 8)[0])
            q += 16
        keys.append(struct.unpack_from("<Q", b, q)[0])
        last = ""
        for i, kid in enumerate(kids):
            if level > 0:
                self._walk(kid)
                continue
            s = self.f.at(kid)
            if b[s:s + 4] != b"SNOD":
                raise H5Error("symbol table node signature")
            nsym = struct.unpack_from("<H", b, s + 6)[0]
            if nsym > 2 * self.f.leaf_k:
                raise H5Error("too many symbols in a node for the file's leaf K")
            names = []
            for j in range(nsym):
                noff, ohdr, cache, _ = struct.unpack_from("<QQII", b, s + 8 + 40 * j)
                nm = self._name(noff)
                names.append(nm)
                self.entries[nm] = (ohdr, cache, b[s + 8 + 40 * j + 24:s + 8 + 40 * j + 40])
            if names != sorted(names) or (names and names[0] <= last and last):
                raise H5Error("symbol table entries are not in increasing name order")
            # key[i] < every name of child i <= key[i+1]
            if names:
                if not (self._name(keys[i]) < names[0] or (i == 0 and self._name(keys[i]) == "")):
                    raise H5Error("B-tree left key does not bound its child")
                if self._name(keys[i + 1]) != names[-1]:
                    raise H5Error("B-tree right key is not the child's largest name")
                last = names[-1]

    def keys(self):
        return sorted(self.entries)

    def __getitem__(self, name):
        ohdr, cache, scratch = self.entries[name]
        msgs = _messages(self.f, ohdr)
        if any(m[0] == 0x0011 for m in msgs):
            return Group(self.f, ohdr)
        return Dataset(self.f, ohdr, msgs)


class Dataset:
    def __init__(self, f, addr, msgs):
        self.f = f
        self.shape = self.dtype = None
        self.layout = None
        self.messages = [m[0] for m in msgs]
        for mtype, flags, d in msgs:
            if mtype == 0x0001:
                ver, rank, fl = struct.unpack_from("<BBB", d, 0)
                if ver != 1:
                    raise H5Error(f"dataspace version {ver}")
                self.shape = tuple(struct.unpack_from(f"<{rank}Q", d, 8))
            elif mtype == 0x0003:
                cls, ver = d[0] & 0x0F, d[0] >> 4
                bits = d[1] | (d[2] << 8) | (d[3] << 16)
                size = struct.unpack_from("<I", d, 4)[0]
                order = ">" if bits & 1 else "<"
                if cls == 0:
                    boff, prec = struct.unpack_from("<HH", d, 8)
                    if boff != 0 or prec != 8 * size:
                        raise H5Error("fixed-point type with padding bits")
                    self.dtype = np.dtype(f"{order}{'i' if bits & 8 else 'u'}{size}")
                elif cls == 1:
                    boff, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", d, 8)
                    ieee = {4: (32, 23, 8, 0, 23, 127), 8: (64, 52, 11, 0, 52, 1023)}.get(size)
                    if ieee is None or (prec, eloc, esize, mloc, msize, bias) != ieee or boff != 0:
                        raise H5Error("floating-point type is not IEEE binary32/64")
                    if (bits >> 8) & 0xFF != prec - 1 or (bits >> 4) & 3 != 2:
                        raise H5Error("floating-point sign position / mantissa normalisation")
                    self.dtype = np.dtype(f"{order}f{size}")
                else:
                    self.dtype = ("class", cls, size)
            elif mtype == 0x0008:
                ver = d[0]
                if ver == 3:
                    lclass = d[1]
                    if lclass == 1:
                        a, s = struct.unpack_from("<QQ", d, 2)
                        self.layout = ("contiguous", a, s)
                    elif lclass == 0:
                        s = struct.unpack_from("<H", d, 2)[0]
                        self.layout = ("compact", d[4:4 + s])
                    else:
                        self.layout = ("chunked",)
                elif ver in (1, 2):
                    rank, lclass = d[1], d[2]
                    if lclass == 1:
                        a = struct.unpack_from("<Q", d, 8)[0]
                        dims = struct.unpack_from(f"<{rank}I", d, 16)
                        self.layout = ("contiguous", a, int(np.prod(dims)))
                    else:
                        self.layout = ("other", lclass)
                else:
                    raise H5Error(f"layout version {ver}")

    def read(self):
        if not isinstance(self.dtype, np.dtype):
            raise H5Error(f"unsupported datatype {self.dtype}")
        count = int(np.prod(self.shape)) if self.shape else 1
        nbytes = count * self.dtype.itemsize
        if self.layout[0] == "contiguous":
            _, a, s = self.layout
            if a == UNDEF:
                raise H5Error("dataset storage not allocated")
            if s != nbytes:
                raise H5Error(f"layout size {s} != {nbytes}")
            p = self.f.at(a)
            if p + nbytes > len(self.f.buf):
                raise H5Error("dataset data beyond the end of the file")
            raw = self.f.buf[p:p + nbytes]
        elif self.layout[0] == "compact":
            raw = self.layout[1][:nbytes]
        else:
            raise H5Error(f"unsupported layout {self.layout}")
        return np.frombuffer(raw, dtype=self.dtype).reshape(self.shape)
