"""HDF5 on-disk encoding / decoding for `h5lite` (the HDF5 File Format Specification, version 3.0 subset).

What is written (everything little endian, 8-byte offsets and lengths):
  * superblock version 2 (48 bytes, Jenkins lookup3 checksum);
  * one version-2 object header ("OHDR", checksummed) per group / dataset, in a single chunk;
  * groups: Link Info (0x02) + Group Info (0x0A) + one hard Link message (0x06) per child - "compact" link storage, no
    B-trees or heaps;
  * datasets: Dataspace v2 (0x01), Datatype (0x03), Fill Value v3 (0x05), Data Layout v3 contiguous (0x08), raw data
    8-byte aligned;
  * attributes: Attribute message v3 (0x0C) in the object header; Python `str` values are variable-length UTF-8 strings
    (datatype class 9) held in global heap collections ("GCOL") - what h5py writes for `obj.attrs['repr'] = "..."`, so
    they read back as `str`; `bytes` are fixed-length strings; numbers / arrays are numeric datasets;
  * datatypes: signed / unsigned integers, IEEE float32 / float64, fixed-length strings (numpy 'S'), numpy bool as h5py's
    enum {FALSE = 0, TRUE = 1} over int8.
The reader parses exactly this subset back (`read_into`) and refuses anything else loudly.

libhdf5 itself is not in this image, so the encoder is validated here by its own decoder (tests/test_hdf.py) and by the
field-by-field layout checks in that test; files are meant to open with h5py / libhdf5 >= 1.8.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
_M32 = 0xFFFFFFFF


# ------------------------------------------------------------------------------------------ checksum
def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & _M32


def lookup3(data, initval=0):
    """Bob Jenkins' lookup3 `hashlittle` - the metadata checksum of HDF5 (H5_checksum_lookup3)."""
    data = bytes(data)
    n = len(data)
    a = b = c = (0xDEADBEEF + n + initval) & _M32
    i = 0
    while n > 12:
        a = (a + int.from_bytes(data[i:i + 4], "little")) & _M32
        b = (b + int.from_bytes(data[i + 4:i + 8], "little")) & _M32
        c = (c + int.from_bytes(data[i + 8:i + 12], "little")) & _M32
        a = (a - c) & _M32; a ^= _rot(c, 4); c = (c + b) & _M32
        b = (b - a) & _M32; b ^= _rot(a, 6); a = (a + c) & _M32
        c = (c - b) & _M32; c ^= _rot(b, 8); b = (b + a) & _M32
        a = (a - c) & _M32; a ^= _rot(c, 16); c = (c + b) & _M32
        b = (b - a) & _M32; b ^= _rot(a, 19); a = (a + c) & _M32
        c = (c - b) & _M32; c ^= _rot(b, 4); b = (b + a) & _M32
        i += 12
        n -= 12
    if n == 0:
        return c
    tail = data[i:] + b"\x00" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & _M32
    b = (b + int.from_bytes(tail[4:8], "little")) & _M32
    c = (c + int.from_bytes(tail[8:12], "little")) & _M32
    c ^= b; c = (c - _rot(b, 14)) & _M32
    a ^= c; a = (a - _rot(c, 11)) & _M32
    b ^= a; b = (b - _rot(a, 25)) & _M32
    c ^= b; c = (c - _rot(b, 16)) & _M32
    a ^= c; a = (a - _rot(c, 4)) & _M32
    b ^= a; b = (b - _rot(a, 14)) & _M32
    c ^= b; c = (c - _rot(b, 24)) & _M32
    return c


def _pad8(n):
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------ datatypes
_VLEN_STR = "vlen-utf8"


def encode_datatype(dt):
    """Datatype message body for a numpy dtype (or the vlen-string marker)."""
    if dt == _VLEN_STR:
        base = struct.pack("<BBBBI", 0x10, 0x00, 0, 0, 1) + struct.pack("<HH", 0, 8)        # unsigned 8-bit integer
        return struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + base                         # class 9: string, UTF-8
    dt = np.dtype(dt)
    if dt.kind == "b":    # h5py's bool: enum over int8, version-3 datatype (names not padded)
        base = struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 1) + struct.pack("<HH", 0, 8)
        return struct.pack("<BBBBI", 0x38, 2, 0, 0, 1) + base + b"FALSE\x00TRUE\x00" + b"\x00\x01"
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 8:
            prop = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            return struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + prop
        prop = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + prop
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, max(dt.itemsize, 1))                  # null-padded, ASCII
    raise TypeError("h5lite cannot store dtype %r" % (dt,))


def decode_datatype(buf, pos):
    """-> (numpy dtype or _VLEN_STR, bytes consumed)."""
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, pos)
    cls, ver = cv & 0x0F, cv >> 4
    if cls == 0:
        return np.dtype(("<i%d" if b0 & 0x08 else "<u%d") % size), 12
    if cls == 1:
        return np.dtype("<f%d" % size), 20
    if cls == 3:
        return np.dtype("S%d" % size), 8
    if cls == 9:
        if (b0 & 0x0F) != 1:
            raise NotImplementedError("variable-length sequences")
        _, n = decode_datatype(buf, pos + 8)
        return _VLEN_STR, 8 + n
    if cls == 8:
        nmem = b0 | (b1 << 8)
        base, n = decode_datatype(buf, pos + 8)
        p = pos + 8 + n
        names = []
        for _ in range(nmem):
            e = p + bytes(buf[p:p + 256]).index(b"\x00")
            names.append(bytes(buf[p:e]))
            p = e + 1 if ver >= 3 else p + _pad8(e + 1 - p)
        p += nmem * base.itemsize
        if names == [b"FALSE", b"TRUE"] and base.itemsize == 1:
            return np.dtype(bool), p - pos
        return base, p - pos
    raise NotImplementedError("HDF5 datatype class %d" % cls)


def encode_dataspace(shape):
    if shape == ():
        return struct.pack("<BBBB", 2, 0, 0, 0)
    return struct.pack("<BBBB", 2, len(shape), 0, 1) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def decode_dataspace(buf, pos):
    ver, rank, flags, typ = struct.unpack_from("<BBBB", buf, pos)
    if ver != 2:
        raise NotImplementedError("dataspace message version %d" % ver)
    shape = tuple(struct.unpack_from("<Q", buf, pos + 4 + 8 * i)[0] for i in range(rank))
    return shape, 4 + 8 * rank * (2 if flags & 1 else 1)


# ------------------------------------------------------------------------------------------ writer
class _Heap:
    """Global heap collections holding the variable-length strings of the file."""

    PER_COLLECTION = 2000

    def __init__(self):
        self.items = []   # utf-8 bytes

    def add(self, s):
        self.items.append(s)
        return len(self.items) - 1

    def collections(self):
        for c0 in range(0, len(self.items), self.PER_COLLECTION):
            yield self.items[c0:c0 + self.PER_COLLECTION]

    def size_of(self, objs):
        used = 16 + sum(16 + _pad8(len(o)) for o in objs)
        return max(4096, used + 16)

    def place(self, addr):
        """Assign addresses: item i -> (collection address, object index); returns the address after the last collection."""
        self.where, self.blocks = [], []
        for objs in self.collections():
            size = self.size_of(objs)
            self.blocks.append((addr, size, objs))
            self.where += [(addr, j + 1) for j in range(len(objs))]
            addr += size
        return addr

    def encode(self, addr, size, objs):
        out = bytearray(b"GCOL" + struct.pack("<BBBBQ", 1, 0, 0, 0, size))
        for j, o in enumerate(objs):
            out += struct.pack("<HHIQ", j + 1, 1, 0, len(o)) + o + b"\x00" * (_pad8(len(o)) - len(o))
        out += struct.pack("<HHIQ", 0, 0, 0, size - len(out))
        return bytes(out) + b"\x00" * (size - len(out))


def _attr_payload(value, heap):
    """-> (datatype body, dataspace body, data bytes or ('vlen', heap item index))."""
    if isinstance(value, (str, np.str_)):
        return encode_datatype(_VLEN_STR), encode_dataspace(()), ("vlen", [heap.add(str(value).encode("utf-8"))])
    if isinstance(value, (bytes, np.bytes_)):
        a = np.asarray(value, dtype="S%d" % max(len(value), 1))
    else:
        a = np.asarray(value)
        if a.dtype.kind == "U" or (a.dtype == object and all(isinstance(x, str) for x in a.reshape(-1))):
            # a list / array of str: an array of variable-length UTF-8 strings (what h5py writes; TdemSystem's .stm lines)
            idx = [heap.add(str(x).encode("utf-8")) for x in a.reshape(-1)]
            return encode_datatype(_VLEN_STR), encode_dataspace(a.shape), ("vlen", idx)
        if a.dtype == object:
            raise TypeError("h5lite cannot store attribute value %r" % (value,))
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    return encode_datatype(a.dtype), encode_dataspace(a.shape), np.ascontiguousarray(a).tobytes()


def _msg(mtype, body):
    return struct.pack("<BHB", mtype, len(body), 0) + body


class _Plan:
    """Sizes first, then addresses, then bytes: an object header's length does not depend on any address."""

    def __init__(self, node, heap):
        from .h5lite import Dataset
        self.node, self.is_dataset = node, isinstance(node, Dataset)
        self.attrs = [(k, _attr_payload(v, heap)) for k, v in node.attrs.items()]
        self.children = [] if self.is_dataset else [(k, _Plan(c, heap)) for k, c in node._children.items()]
        n = 0
        for k, (dt, ds, data) in self.attrs:
            n += 4 + 9 + len(k.encode("utf-8")) + 1 + len(dt) + len(ds) + (16 * len(data[1]) if isinstance(data, tuple) else len(data))
        if self.is_dataset:
            a = node._a
            self.dt, self.ds = encode_datatype(a.dtype), encode_dataspace(a.shape)
            n += (4 + len(self.ds)) + (4 + len(self.dt)) + (4 + 2) + (4 + 18)
        else:
            n += (4 + 18) + (4 + 2)
            for k, _ in self.children:
                kb = k.encode("utf-8")
                n += 4 + 2 + (1 if not kb.isascii() else 0) + (1 if len(kb) < 256 else 2) + len(kb) + 8
        self.msg_bytes = n
        self.header_bytes = 10 + n + 4

    def place(self, addr):
        self.addr = addr
        addr = _pad8(addr + self.header_bytes)
        if self.is_dataset:
            nbytes = self.node._a.nbytes
            self.data_addr = addr if nbytes else UNDEF
            addr = _pad8(addr + nbytes)
        for _, c in self.children:
            addr = c.place(addr)
        return addr

    def encode(self, heap):
        body = bytearray()
        if self.is_dataset:
            body += _msg(0x01, self.ds) + _msg(0x03, self.dt) + _msg(0x05, struct.pack("<BB", 3, 0x0A))
            body += _msg(0x08, struct.pack("<BBQQ", 3, 1, self.data_addr, self.node._a.nbytes))
        else:
            body += _msg(0x02, struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF)) + _msg(0x0A, struct.pack("<BB", 0, 0))
            for k, c in self.children:
                kb = k.encode("utf-8")
                flags = (0 if len(kb) < 256 else 1) | (0 if kb.isascii() else 0x10)
                link = struct.pack("<BB", 1, flags) + (b"" if kb.isascii() else b"\x01")
                link += struct.pack("<B" if len(kb) < 256 else "<H", len(kb)) + kb + struct.pack("<Q", c.addr)
                body += _msg(0x06, link)
        for k, (dt, ds, data) in self.attrs:
            kb = k.encode("utf-8") + b"\x00"
            if isinstance(data, tuple):
                data = b"".join(struct.pack("<IQI", len(heap.items[i]), *heap.where[i]) for i in data[1])
            body += _msg(0x0C, struct.pack("<BBHHHB", 3, 0, len(kb), len(dt), len(ds), 0 if kb.isascii() else 1) + kb + dt + ds + data)
        assert len(body) == self.msg_bytes, (len(body), self.msg_bytes)
        head = b"OHDR" + struct.pack("<BBI", 2, 0x02, len(body)) + bytes(body)
        return head + struct.pack("<I", lookup3(head))


def write(filename, root):
    """Serialise the in-memory tree under `root` (an h5lite Group) to `filename`."""
    heap = _Heap()
    plan = _Plan(root, heap)
    addr = heap.place(48)
    eof = plan.place(_pad8(addr))
    with open(filename, "wb") as f:
        sb = SIGNATURE + struct.pack("<BBBBQQQQ", 2, 8, 8, 0, 0, UNDEF, eof, plan.addr)
        f.write(sb + struct.pack("<I", lookup3(sb)))
        for caddr, size, objs in heap.blocks:
            f.seek(caddr)
            f.write(heap.encode(caddr, size, objs))

        def emit(p):
            f.seek(p.addr)
            f.write(p.encode(heap))
            if p.is_dataset and p.node._a.nbytes:
                a = p.node._a
                if a.dtype.byteorder == ">":
                    a = a.astype(a.dtype.newbyteorder("<"))
                f.seek(p.data_addr)
                f.write(np.ascontiguousarray(a).view(np.uint8).reshape(-1).data)
            for _, c in p.children:
                emit(c)
        emit(plan)
        f.truncate(eof)


# ------------------------------------------------------------------------------------------ reader
def _read_header(buf, addr):
    """-> list of (message type, body offset, body size) of the version-2 object header at `addr` (checksum verified)."""
    if bytes(buf[addr:addr + 4]) != b"OHDR":
        raise NotImplementedError("object header at %d is not version 2 (h5lite reads the files it writes)" % addr)
    ver, flags = struct.unpack_from("<BB", buf, addr + 4)
    p = addr + 6
    if flags & 0x20:
        p += 16
    if flags & 0x10:
        p += 4
    nsz = 1 << (flags & 3)
    size = int.from_bytes(buf[p:p + nsz], "little")
    p += nsz
    end = p + size
    if struct.unpack_from("<I", buf, end)[0] != lookup3(buf[addr:end]):
        raise ValueError("object header checksum mismatch at %d" % addr)
    msgs, extra = [], 2 if flags & 0x04 else 0
    while p + 4 + extra <= end:
        mtype, msize, _ = struct.unpack_from("<BHB", buf, p)
        p += 4 + extra
        if mtype == 0x10:
            raise NotImplementedError("object header continuation blocks")
        msgs.append((mtype, p, msize))
        p += msize
    return msgs


def _heap_object(buf, caddr, idx):
    if bytes(buf[caddr:caddr + 4]) != b"GCOL":
        raise ValueError("no global heap collection at %d" % caddr)
    size = struct.unpack_from("<Q", buf, caddr + 8)[0]
    p = caddr + 16
    while p + 16 <= caddr + size:
        i, _, _, n = struct.unpack_from("<HHIQ", buf, p)
        if i == 0:
            break
        if i == idx:
            return bytes(buf[p + 16:p + 16 + n])
        p += 16 + _pad8(n)
    raise KeyError("global heap object %d not found" % idx)


def _read_attr(buf, p):
    ver, _, nlen, dlen, slen = struct.unpack_from("<BBHHH", buf, p)
    if ver != 3:
        raise NotImplementedError("attribute message version %d" % ver)
    p += 9
    name = bytes(buf[p:p + nlen - 1]).decode("utf-8")
    p += nlen
    dt, _ = decode_datatype(buf, p)
    shape, _ = decode_dataspace(buf, p + dlen)
    p += dlen + slen
    if dt == _VLEN_STR:
        n = int(np.prod(shape)) if shape else 1
        out = []
        for j in range(n):
            ln, caddr, idx = struct.unpack_from("<IQI", buf, p + 16 * j)
            out.append(_heap_object(buf, caddr, idx)[:ln].decode("utf-8"))
        return name, (out[0] if shape == () else np.asarray(out, dtype=object).reshape(shape))
    a = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape)) if shape else 1, offset=p).reshape(shape).copy()
    if dt.kind == "S" and shape == ():
        return name, np.bytes_(a[()])
    return name, (a[()] if shape == () else a)


def _read_node(buf, addr, node):
    from .h5lite import Dataset, Group
    msgs = _read_header(buf, addr)
    types = [m[0] for m in msgs]
    links, shape, dt, layout = [], None, None, None
    for mtype, p, n in msgs:
        if mtype == 0x0C:
            k, v = _read_attr(buf, p)
            node.attrs[k] = v
        elif mtype == 0x06:
            ver, flags = struct.unpack_from("<BB", buf, p)
            q = p + 2
            if flags & 0x08:
                if buf[q] != 0:
                    raise NotImplementedError("soft / external links")
                q += 1
            if flags & 0x04:
                q += 8
            if flags & 0x10:
                q += 1
            nsz = 1 << (flags & 3)
            ln = int.from_bytes(buf[q:q + nsz], "little")
            q += nsz
            links.append((bytes(buf[q:q + ln]).decode("utf-8"), struct.unpack_from("<Q", buf, q + ln)[0]))
        elif mtype == 0x01:
            shape, _ = decode_dataspace(buf, p)
        elif mtype == 0x03:
            dt, _ = decode_datatype(buf, p)
        elif mtype == 0x08:
            ver, cls = struct.unpack_from("<BB", buf, p)
            if ver != 3 or cls != 1:
                raise NotImplementedError("data layout version %d class %d (contiguous only)" % (ver, cls))
            layout = struct.unpack_from("<QQ", buf, p + 2)
    return types, links, shape, dt, layout


def read_into(filename, root):
    """Parse `filename` (written by `write`) into the in-memory tree under `root`."""
    from .h5lite import Dataset, Group
    buf = np.fromfile(filename, dtype=np.uint8)
    mv = memoryview(buf)
    if bytes(mv[:8]) != SIGNATURE:
        raise OSError("%s is not an HDF5 file" % filename)
    ver = mv[8]
    if ver not in (2, 3) or mv[9] != 8 or mv[10] != 8:
        raise NotImplementedError("superblock version %d / offset size %d: h5lite reads the files it writes" % (ver, mv[9]))
    base, ext, eof, root_addr = struct.unpack_from("<QQQQ", mv, 12)
    if struct.unpack_from("<I", mv, 44)[0] != lookup3(mv[:44]):
        raise ValueError("superblock checksum mismatch")

    def rec(addr, group):
        types, links, shape, dt, layout = _read_node(mv, addr, group)
        for name, caddr in links:
            ctypes_, clinks, cshape, cdt, clayout = _read_node(mv, caddr, Group())   # peek: dataset or group?
            if 0x08 in ctypes_:
                if cdt == _VLEN_STR:
                    raise NotImplementedError("variable-length string datasets")
                n = int(np.prod(cshape)) if cshape else 1
                a = (np.frombuffer(buf, dtype=cdt, count=n, offset=clayout[0]).reshape(cshape).copy() if clayout[0] != UNDEF and n
                     else np.zeros(cshape, dtype=cdt))
                d = Dataset(group, name, a)
                _read_node(mv, caddr, d)
                group._children[name] = d
            else:
                g = Group(group, name)
                group._children[name] = g
                rec(caddr, g)
    rec(root_addr, root)
