"""Minimal, dependency-free HDF5 reader/writer for the file subset VaSP's hemodynamics path touches.

No libhdf5 / h5py exists in the target image, so the drop-in reads and writes the
files itself.  The subset is exactly what legacy dolfin (HDF5 1.12, ``libver=earliest``) emits for

* ``Mesh/mesh_fluid.h5`` / ``mesh_refined_fluid.h5``  (read at reference
  ``compute_hemodynamics.py:187-197``),
* ``Visualization_separate_domain/u.h5``              (written at ``create_hdf5.py:172-174``, read at
  ``compute_hemodynamics.py:176-179,269,274,277``),
* ``Hemodynamic_indices/<Name>.h5``                    (``write_checkpoint`` at
  ``compute_hemodynamics.py:286,361``; layout documented by the reference itself in
  ``postprocessing_h5py/postprocessing_h5py_common.py:234-242,639-662``):

superblock v0 (8-byte offsets/lengths), v1 object headers with continuation blocks, old-style groups
(symbol-table message -> v1 B-tree ``TREE`` of any depth + ``HEAP`` local heap + ``SNOD`` nodes),
datasets with dataspace v1/v2, fixed-point / IEEE float / fixed-string datatypes, layout v3
contiguous or compact, attribute messages v1-v3.  Chunked or filtered datasets are rejected loudly.

Every contiguous dataset is reported with its absolute ``(offset, nbytes)`` so the snapshot streamer can
``pread`` raw little-endian bytes straight into pinned staging buffers without any HDF5 library in the
loop.
"""
from __future__ import annotations

import mmap
import os
import struct
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Tuple, Union

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(RuntimeError):
    """The file uses an HDF5 feature outside the supported subset."""


# --------------------------------------------------------------------------------------------------
# reader
# --------------------------------------------------------------------------------------------------
def _pad8(n: int) -> int:
    return (n + 7) & ~7


def _parse_datatype(buf: bytes, pos: int) -> Tuple[np.dtype, int]:
    """Datatype message body -> (numpy dtype, bytes consumed)."""
    cls_ver = buf[pos]
    cls = cls_ver & 0x0F
    bits0 = buf[pos + 1]
    size = struct.unpack_from("<I", buf, pos + 4)[0]
    if cls == 0:  # fixed point
        order = ">" if (bits0 & 1) else "<"
        signed = bool(bits0 & 0x08)
        return np.dtype(f"{order}{'i' if signed else 'u'}{size}"), 8 + 4
    if cls == 1:  # IEEE float
        order = ">" if (bits0 & 1) else "<"
        return np.dtype(f"{order}f{size}"), 8 + 12
    if cls == 3:  # fixed-length string
        return np.dtype(f"S{size}"), 8
    raise H5FormatError(f"unsupported HDF5 datatype class {cls}")


def _parse_dataspace(buf: bytes, pos: int) -> Tuple[Tuple[int, ...], int]:
    ver = buf[pos]
    rank = buf[pos + 1]
    flags = buf[pos + 2]
    if ver == 1:
        p = pos + 8
    elif ver == 2:
        if buf[pos + 3] == 2:  # null dataspace
            return (0,), 4
        p = pos + 4
    else:
        raise H5FormatError(f"unsupported dataspace version {ver}")
    dims = struct.unpack_from(f"<{rank}Q", buf, p) if rank else ()
    p += 8 * rank
    if flags & 1:
        p += 8 * rank
    if ver == 1 and flags & 2:
        p += 8 * rank
    return tuple(int(d) for d in dims), p - pos


@dataclass
class _Msg:
    type: int
    flags: int
    pos: int
    size: int


class H5Object:
    """A group or dataset, addressed by its object-header offset."""

    def __init__(self, f: "H5File", addr: int, name: str, msgs: Optional[List["_Msg"]] = None):
        self._f = f
        self.addr = addr
        self.name = name
        self._msgs = f._read_header(addr) if msgs is None else msgs
        self._attrs: Optional[Dict[str, np.ndarray]] = None

    # -- attributes --------------------------------------------------------------------------------
    def attr_value_offset(self, name: str) -> Optional[int]:
        """Absolute file offset of the value of attribute ``name`` (``None`` if absent)."""
        self.attrs
        return self._attr_pos.get(name)

    @property
    def attrs(self) -> Dict[str, np.ndarray]:
        if self._attrs is None:
            self._attrs = {}
            self._attr_pos: Dict[str, int] = {}
            buf = self._f._buf
            for m in self._msgs:
                if m.type != 0x000C:
                    continue
                p = m.pos
                ver = buf[p]
                nsz, tsz, ssz = struct.unpack_from("<HHH", buf, p + 2)
                p += 8
                if ver == 3:
                    p += 1
                pad = _pad8 if ver == 1 else (lambda n: n)
                aname = bytes(buf[p:p + nsz]).split(b"\0", 1)[0].decode()
                p += pad(nsz)
                dt, _ = _parse_datatype(buf, p)
                p += pad(tsz)
                shape, _ = _parse_dataspace(buf, p)
                p += pad(ssz)
                count = int(np.prod(shape)) if shape else 1
                val = np.frombuffer(buf, dtype=dt, count=count, offset=p).copy()
                self._attrs[aname] = val.reshape(shape) if shape else val.reshape(())
                self._attr_pos[aname] = p
        return self._attrs

    def _find(self, mtype: int) -> Optional[_Msg]:
        for m in self._msgs:
            if m.type == mtype:
                return m
        return None

    @property
    def is_group(self) -> bool:
        return self._find(0x0011) is not None

    @property
    def is_dataset(self) -> bool:
        return self._find(0x0008) is not None


class H5Dataset(H5Object):
    def __init__(self, f: "H5File", addr: int, name: str, msgs: Optional[List["_Msg"]] = None):
        super().__init__(f, addr, name, msgs)
        buf = f._buf
        ms, mt, ml = self._find(0x0001), self._find(0x0003), self._find(0x0008)
        if ms is None or mt is None or ml is None:
            raise H5FormatError(f"{name}: not a dataset")
        self.shape, _ = _parse_dataspace(buf, ms.pos)
        self.dtype, _ = _parse_datatype(buf, mt.pos)
        ver = buf[ml.pos]
        if ver != 3:
            raise H5FormatError(f"{name}: data layout version {ver} unsupported (need 3)")
        lclass = buf[ml.pos + 1]
        self.compact = False
        if lclass == 1:
            self.offset, self.nbytes = struct.unpack_from("<QQ", buf, ml.pos + 2)
            self.offset += f.base
        elif lclass == 0:
            self.nbytes = struct.unpack_from("<H", buf, ml.pos + 2)[0]
            self.offset = ml.pos + 4
            self.compact = True
        else:
            raise H5FormatError(f"{name}: chunked/filtered datasets are not supported "
                                "(dolfin writes contiguous data; SURVEY.md §5.9)")
        expect = int(np.prod(self.shape)) * self.dtype.itemsize if self.shape else self.dtype.itemsize
        if self.offset == _UNDEF:  # never written
            self.nbytes = 0
        elif self.nbytes < expect:
            raise H5FormatError(f"{name}: layout size {self.nbytes} < {expect}")
        else:
            self.nbytes = expect

    def read(self) -> np.ndarray:
        count = int(np.prod(self.shape)) if self.shape else 1
        if self.nbytes == 0:
            return np.zeros(self.shape, self.dtype)
        a = np.frombuffer(self._f._buf, dtype=self.dtype, count=count, offset=self.offset)
        return a.reshape(self.shape).copy()

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, idx):
        return self.read()[idx]


class H5Group(H5Object):
    def __init__(self, f: "H5File", addr: int, name: str, msgs: Optional[List["_Msg"]] = None):
        super().__init__(f, addr, name, msgs)
        m = self._find(0x0011)
        if m is None:
            raise H5FormatError(f"{name}: not an old-style group")
        self._btree, self._heap = struct.unpack_from("<QQ", f._buf, m.pos)
        self._links: Optional[Dict[str, int]] = None

    def _load(self) -> Dict[str, int]:
        if self._links is None:
            f = self._f
            buf = f._buf
            hp = f.base + self._heap
            if bytes(buf[hp:hp + 4]) != b"HEAP":
                raise H5FormatError("bad local heap signature")
            heap_data = f.base + struct.unpack_from("<Q", buf, hp + 24)[0]
            links: Dict[str, int] = {}

            def walk(addr: int) -> None:
                p = f.base + addr
                sig = bytes(buf[p:p + 4])
                if sig == b"TREE":
                    ntype, level, used = struct.unpack_from("<BBH", buf, p + 4)
                    if ntype != 0:
                        raise H5FormatError("expected group B-tree node")
                    q = p + 24
                    for i in range(used):
                        child = struct.unpack_from("<Q", buf, q + 8 + 16 * i)[0]
                        walk(child)
                elif sig == b"SNOD":
                    nsym = struct.unpack_from("<H", buf, p + 6)[0]
                    for i in range(nsym):
                        e = p + 8 + 40 * i
                        noff, oaddr = struct.unpack_from("<QQ", buf, e)
                        s = heap_data + noff
                        end = buf.find(b"\0", s)
                        links[bytes(buf[s:end]).decode()] = oaddr
                else:
                    raise H5FormatError(f"unexpected node signature {sig!r} in group B-tree")

            if self._btree != _UNDEF:
                walk(self._btree)
            self._links = links
        return self._links

    def keys(self) -> List[str]:
        return list(self._load().keys())

    def __contains__(self, name: str) -> bool:
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __iter__(self) -> Iterator[str]:
        return iter(self.keys())

    def __getitem__(self, path: str) -> Union["H5Group", H5Dataset]:
        node: Union[H5Group, H5Dataset] = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, H5Group):
                raise KeyError(path)
            links = node._load()
            if part not in links:
                raise KeyError(f"{path!r}: no member {part!r} in {node.name!r}")
            node = self._f._open(links[part], f"{node.name.rstrip('/')}/{part}")
        return node


def dataset_table(group: "H5Group", names: List[str], attr: str) -> Tuple[np.ndarray, np.ndarray, List["H5Dataset"]]:
    """``(data offsets, float64 values of scalar attribute attr)`` of many like datasets of one group, plus the
    first dataset parsed in full.

    A time series holds thousands of datasets written by the same code path: their object headers are byte-for-byte
    equal except for the data address, the attribute's value and the modification time.  The first header is parsed
    normally; the others are compared with it as one numpy block and only the two fields are lifted out.  Any header
    that differs anywhere else (another shape, another attribute set, a continuation block) is parsed the slow way.
    """
    f = group._f
    links = group._load()
    first = group[names[0]]
    if not isinstance(first, H5Dataset):
        raise KeyError(f"{names[0]}: not a dataset")
    n = len(names)
    offsets = np.empty(n, dtype=np.int64)
    values = np.empty(n, dtype=np.float64)
    offsets[0], values[0] = first.offset, float(first.attrs[attr])
    slow = list(range(1, n))
    p0 = f.base + links[names[0]]
    hsize = struct.unpack_from("<I", f._buf, p0 + 8)[0]
    ml, at = first._find(0x0008), first.attr_value_offset(attr)
    simple = (not first.compact and all(m.type != 0x0010 for m in first._msgs) and at is not None
              and first.attrs[attr].dtype == np.dtype("<f8") and first.attrs[attr].shape == ())
    if simple and n > 1:
        span = 16 + hsize
        raw = np.frombuffer(f._buf, dtype=np.uint8)
        starts = np.array([f.base + links[nm] for nm in names], dtype=np.int64)
        ok = starts + span <= raw.size
        H = raw[np.where(ok, starts, 0)[:, None] + np.arange(span)[None, :]]
        rel_addr, rel_val = ml.pos + 2 - p0, at - p0
        mask = np.ones(span, dtype=bool)
        mask[rel_addr:rel_addr + 8] = False
        mask[rel_val:rel_val + 8] = False
        mt = first._find(0x0012)
        if mt is not None:
            mask[mt.pos - p0 + 4:mt.pos - p0 + 8] = False
        same = ok & (H[:, mask] == H[0, mask]).all(axis=1)
        offsets[same] = np.ascontiguousarray(H[same, rel_addr:rel_addr + 8]).view("<u8").ravel().astype(np.int64) + f.base
        values[same] = np.ascontiguousarray(H[same, rel_val:rel_val + 8]).view("<f8").ravel()
        slow = [i for i in np.nonzero(~same)[0] if i != 0]
    for i in slow:
        d = group[names[i]]
        if d.shape != first.shape or d.dtype != first.dtype:
            raise H5FormatError(f"{names[i]}: shape/type {d.shape} {d.dtype} differs from {names[0]}")
        offsets[i], values[i] = d.offset, float(d.attrs[attr])
    return offsets, values, [first]


class H5File(H5Group):
    """Read-only view of an HDF5 file (the subset described in the module docstring)."""

    def __init__(self, path: Union[str, Path]):
        self.path = Path(path)
        self._fh = open(self.path, "rb")
        size = os.fstat(self._fh.fileno()).st_size
        if size < 96:
            raise H5FormatError(f"{path}: too small to be HDF5")
        self._buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        buf = self._buf
        sb = -1
        off = 0
        while off < size:  # superblock may sit at 0, 512, 1024, ...
            if bytes(buf[off:off + 8]) == _SIG:
                sb = off
                break
            off = 512 if off == 0 else off * 2
        if sb < 0:
            raise H5FormatError(f"{path}: HDF5 signature not found")
        ver = buf[sb + 8]
        if ver not in (0, 1):
            raise H5FormatError(f"{path}: superblock version {ver} unsupported (need 0/1, libver=earliest)")
        so, sl = buf[sb + 13], buf[sb + 14]
        if so != 8 or sl != 8:
            raise H5FormatError("only 8-byte offsets/lengths supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", buf, sb + 16)
        p = sb + 24 + (4 if ver == 1 else 0)
        base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", buf, p)
        self.base = base
        p += 32
        _noff, root_addr = struct.unpack_from("<QQ", buf, p)
        self._cache: Dict[int, H5Object] = {}
        self._f = self
        H5Group.__init__(self, self, root_addr, "/")

    # -- object headers ------------------------------------------------------------------------------
    def _read_header(self, addr: int) -> List[_Msg]:
        buf = self._buf
        p = self.base + addr
        if buf[p] != 1:
            if bytes(buf[p:p + 4]) == b"OHDR":
                raise H5FormatError("v2 object headers unsupported (file not written with libver=earliest)")
            raise H5FormatError(f"bad object header version {buf[p]} at {addr}")
        nmsg = struct.unpack_from("<H", buf, p + 2)[0]
        hsize = struct.unpack_from("<I", buf, p + 8)[0]
        blocks = [(p + 16, hsize)]
        msgs: List[_Msg] = []
        while blocks and len(msgs) < nmsg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and len(msgs) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", buf, q)
                body = q + 8
                if mtype == 0x0010:
                    coff, clen = struct.unpack_from("<QQ", buf, body)
                    blocks.append((self.base + coff, clen))
                if mflags & 0x02:
                    raise H5FormatError("shared header messages unsupported")
                msgs.append(_Msg(mtype, mflags, body, msize))
                q = body + msize
        return msgs

    def _open(self, addr: int, name: str) -> Union[H5Group, H5Dataset]:
        if addr not in self._cache:
            msgs = self._read_header(addr)  # parsed once, handed to the object
            types = {m.type for m in msgs}
            if 0x0011 in types:
                obj: H5Object = H5Group(self, addr, name, msgs)
            elif 0x0008 in types:
                obj = H5Dataset(self, addr, name, msgs)
            else:
                raise H5FormatError(f"{name}: neither old-style group nor dataset")
            self._cache[addr] = obj
        return self._cache[addr]  # type: ignore[return-value]

    def close(self) -> None:
        self._cache.clear()
        try:
            self._buf.close()
        except (BufferError, ValueError):
            pass
        self._fh.close()

    def __enter__(self) -> "H5File":
        return self

    def __exit__(self, *exc) -> None:
        self.close()


# --------------------------------------------------------------------------------------------------
# writer
# --------------------------------------------------------------------------------------------------
def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10 | 0, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize == 8:
        # IEEE double, little endian: sign bit 63, exponent 52..62, mantissa 0..51, bias 1023
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, 0x3F, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, 0x1F, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x10 | 3, 0x00, 0, 0, dt.itemsize)  # null-terminated, ASCII
    raise H5FormatError(f"cannot write dtype {dt}")


def _space_msg(shape: Tuple[int, ...]) -> bytes:
    """Dataspace message v1.  Like the library dolfin links (HDF5 1.12, earliest format) a simple dataspace of rank
    >= 1 carries its maximum dimensions too (flag bit 0; equal to the current ones: nothing here is extendible)."""
    rank = len(shape)
    dims = b"".join(struct.pack("<Q", int(d)) for d in shape)
    return struct.pack("<BBBB4x", 1, rank, 1 if rank else 0, 0) + dims + (dims if rank else b"")


# object modification time written into dataset headers (message 0x0012 v1, seconds since the epoch): taken once per
# process, or from SOURCE_DATE_EPOCH, so that the files of one run are reproducible byte for byte
_MTIME = int(os.environ.get("SOURCE_DATE_EPOCH", 0) or 0) or int(__import__("time").time())


def _attr_msg(name: str, value) -> bytes:
    if isinstance(value, (str, bytes)):
        raw = value.encode() if isinstance(value, str) else value
        arr = np.array(raw, dtype=f"S{max(len(raw), 1)}")  # dolfin stores the exact length, no terminator
    else:
        arr = np.asarray(value)
        if arr.dtype.kind == "f":
            arr = arr.astype("<f8")
        elif arr.dtype.kind in "iub":
            arr = arr.astype("<u8" if arr.dtype.kind == "u" else "<i8")
    nm = name.encode() + b"\0"
    dtm = _dtype_msg(arr.dtype)
    spm = _space_msg(arr.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(spm))
    body += nm.ljust(_pad8(len(nm)), b"\0") + dtm.ljust(_pad8(len(dtm)), b"\0") + spm.ljust(_pad8(len(spm)), b"\0")
    body += arr.tobytes()
    return body


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = body.ljust(_pad8(len(body)), b"\0")
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(msgs: List[bytes], refcount: int = 1) -> bytes:
    payload = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), refcount, len(payload)) + payload


@dataclass
class _WNode:
    name: str
    attrs: Dict[str, object] = field(default_factory=dict)
    children: Dict[str, "_WNode"] = field(default_factory=dict)
    # dataset fields
    is_dataset: bool = False
    shape: Tuple[int, ...] = ()
    dtype: Optional[np.dtype] = None
    data_addr: int = 0
    data_size: int = 0
    nlinks: int = 1          # hard links to this object (its header is written once, with this reference count)
    # filled at close
    header_addr: int = 0
    written: bool = False


class H5Writer:
    """Streaming writer: raw dataset bytes go to disk as they are created, metadata at ``close()``.

    Produces superblock v0 / v1 object headers / symbol-table groups with standard node sizes
    (group leaf K = 4, internal K = 16) and multi-level B-trees, so files are readable by libhdf5-based
    tools (dolfin, h5py, ParaView) as well as by :class:`H5File`.
    """

    LEAF_K = 4
    INTERNAL_K = 16
    _SB_SIZE = 96

    def __init__(self, path: Union[str, Path]):
        self.path = Path(path)
        self._fh = open(self.path, "wb")
        self._fh.write(b"\0" * self._SB_SIZE)
        self._pos = self._SB_SIZE
        self._root = _WNode("/")
        self._closed = False
        self._meta: Optional[bytearray] = None   # metadata is assembled in memory at close() and written in one piece
        self._meta_base = 0
        self._hdr_cache: Dict[tuple, Tuple[bytes, int]] = {}
        self._heap_cache: Dict[tuple, Tuple[bytes, Dict[str, int], int]] = {}

    # -- tree helpers ----------------------------------------------------------------------------------
    def _node(self, path: str, create: bool = True) -> _WNode:
        node = self._root
        for part in [p for p in path.split("/") if p]:
            if part not in node.children:
                if not create:
                    raise KeyError(path)
                node.children[part] = _WNode(part)
            node = node.children[part]
            if node.is_dataset:
                raise ValueError(f"{path}: {part} is a dataset")
        return node

    def create_group(self, path: str, attrs: Optional[Dict[str, object]] = None) -> None:
        node = self._node(path)
        if attrs:
            node.attrs.update(attrs)

    def _alloc(self, nbytes: int) -> int:
        pad = (-self._pos) % 8
        if pad:
            self._fh.write(b"\0" * pad)
            self._pos += pad
        addr = self._pos
        self._pos += nbytes
        return addr

    def create_dataset(self, path: str, data, dtype=None, attrs: Optional[Dict[str, object]] = None,
                       alias_of: Optional[str] = None) -> Tuple[int, int]:
        """Write ``data`` contiguously; returns ``(file offset, nbytes)``.

        ``alias_of`` names an already-written dataset whose raw bytes this one shares (used for the
        per-step boundary-mesh copies of ``write_checkpoint`` series, which are all identical).
        """
        parent, _, leaf = path.rstrip("/").rpartition("/")
        pnode = self._node(parent)
        if leaf in pnode.children:
            raise ValueError(f"{path} already exists")
        if alias_of is not None and not attrs:
            # same bytes, same header: a hard link to the object already written (one header, reference count + 1)
            src = self._lookup(alias_of)
            src.nlinks += 1
            pnode.children[leaf] = src
            return src.data_addr, src.data_size
        node = _WNode(leaf, is_dataset=True)
        if alias_of is not None:
            src = self._lookup(alias_of)
            node.shape, node.dtype, node.data_addr, node.data_size = src.shape, src.dtype, src.data_addr, src.data_size
        else:
            arr = np.ascontiguousarray(data, dtype=dtype)
            if arr.dtype.byteorder == ">":
                arr = arr.astype(arr.dtype.newbyteorder("<"))
            node.shape, node.dtype, node.data_size = arr.shape, arr.dtype, arr.nbytes
            node.data_addr = self._alloc(arr.nbytes)
            self._fh.write(memoryview(arr).cast("B") if arr.nbytes else b"")
        if attrs:
            node.attrs.update(attrs)
        pnode.children[leaf] = node
        return node.data_addr, node.data_size

    def link(self, path: str, target: str) -> None:
        """Hard link: ``path`` becomes another name of the existing object ``target`` (group or dataset)."""
        parent, _, leaf = path.rstrip("/").rpartition("/")
        pnode = self._node(parent)
        if leaf in pnode.children:
            raise ValueError(f"{path} already exists")
        src = self._lookup(target)
        src.nlinks += 1
        pnode.children[leaf] = src

    def create_dataset_block(self, paths: List[str], rows: np.ndarray, shape: Optional[Tuple[int, ...]] = None) -> None:
        """``len(paths)`` datasets whose raw data are the consecutive rows of ``rows`` -- ONE write for the whole
        block (the per-step vectors of a checkpoint series arrive from the GPU as such a block).  ``shape``: shape of
        each dataset (default: the row shape)."""
        arr = np.ascontiguousarray(rows)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        n = len(paths)
        if arr.shape[0] != n:
            raise ValueError("one row per dataset")
        row_bytes = arr.nbytes // max(n, 1)
        if row_bytes % 8:
            raise ValueError("rows must be a multiple of 8 bytes (dataset addresses are 8-byte aligned)")
        base = self._alloc(arr.nbytes)
        self._fh.write(memoryview(arr).cast("B") if arr.nbytes else b"")
        shp = tuple(shape) if shape is not None else arr.shape[1:]
        for k, path in enumerate(paths):
            parent, _, leaf = path.rstrip("/").rpartition("/")
            pnode = self._node(parent)
            if leaf in pnode.children:
                raise ValueError(f"{path} already exists")
            pnode.children[leaf] = _WNode(leaf, is_dataset=True, shape=shp, dtype=arr.dtype,
                                          data_addr=base + k * row_bytes, data_size=row_bytes)

    def _lookup(self, path: str) -> _WNode:
        node = self._root
        for part in [p for p in path.split("/") if p]:
            node = node.children[part]
        return node

    def set_attrs(self, path: str, attrs: Dict[str, object]) -> None:
        self._lookup(path).attrs.update(attrs)

    # -- serialisation -----------------------------------------------------------------------------------
    def _emit(self, blob: bytes) -> int:
        meta = self._meta
        pad = (-len(meta)) % 8
        if pad:
            meta += b"\0" * pad
        addr = self._meta_base + len(meta)
        meta += blob
        return addr

    def _write_group_index(self, node: _WNode) -> Tuple[int, int]:
        """Local heap + SNODs + B-tree for one group; returns (btree addr, heap addr)."""
        names = sorted(node.children.keys(), key=lambda s: s.encode())
        key = tuple(names)
        cached = self._heap_cache.get(key)   # the step groups of a checkpoint series all have the same member names
        if cached is None:
            heap = bytearray(8)  # offset 0: empty string (B-tree key 0)
            offs: Dict[str, int] = {}
            for n in names:
                offs[n] = len(heap)
                b = n.encode() + b"\0"
                heap += b.ljust(_pad8(len(b)), b"\0")
            # libhdf5 wants a free block it can describe (>= 16 bytes) or H5HL_FREE_NULL (=1)
            free_off = len(heap)
            heap += struct.pack("<QQ", 1, 16)
            cached = (bytes(heap), offs, free_off)
            if len(names) <= 16:
                self._heap_cache[key] = cached
        heap_blob, offs, free_off = cached
        data_addr = self._emit(heap_blob)
        heap_addr = self._emit(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_blob), free_off, data_addr))

        cap = 2 * self.LEAF_K
        snod_size = 8 + cap * 40
        leaves: List[Tuple[int, int]] = []  # (addr, heap offset of largest name)
        for i in range(0, max(len(names), 1), cap):
            chunk = names[i:i + cap]
            body = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk)))
            for n in chunk:
                ch = node.children[n]
                if ch.is_dataset:
                    body += struct.pack("<QQII16x", offs[n], ch.header_addr, 0, 0)
                else:
                    bt, hp = ch._index  # type: ignore[attr-defined]
                    body += struct.pack("<QQIIQQ", offs[n], ch.header_addr, 1, 0, bt, hp)
            leaves.append((self._emit(bytes(body).ljust(snod_size, b"\0")), offs[chunk[-1]] if chunk else 0))

        level = 0
        fan = 2 * self.INTERNAL_K
        node_size = 24 + (2 * fan + 1) * 8
        kids = leaves
        while True:
            groups = [kids[i:i + fan] for i in range(0, len(kids), fan)]
            addr0 = self._emit(b"")  # the level's nodes follow each other from the next 8-byte boundary
            addrs = [addr0 + gi * node_size for gi in range(len(groups))]  # siblings link to each other
            nxt: List[Tuple[int, int]] = []
            first_key = 0
            blob = bytearray()
            for gi, grp in enumerate(groups):
                left = addrs[gi - 1] if gi > 0 else _UNDEF
                right = addrs[gi + 1] if gi + 1 < len(groups) else _UNDEF
                body = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, level, len(grp), left, right))
                body += struct.pack("<Q", first_key)
                for caddr, ckey in grp:
                    body += struct.pack("<QQ", caddr, ckey)
                first_key = grp[-1][1]
                blob += bytes(body).ljust(node_size, b"\0")
                nxt.append((addrs[gi], grp[-1][1]))
            got = self._emit(bytes(blob))
            assert got == addr0
            if len(nxt) == 1:
                return nxt[0][0], heap_addr
            kids = nxt
            level += 1

    def _write_node(self, node: _WNode) -> None:
        if node.written:   # reached through another hard link already
            return
        node.written = True
        if node.is_dataset and not node.attrs:
            # attribute-less datasets of one shape and type (the vectors of a series) differ in their data address only
            key = (node.shape, node.dtype.str, node.data_size, node.nlinks)
            tpl = self._hdr_cache.get(key)
            if tpl is not None:
                blob, at = tpl
                hdr = bytearray(blob)
                struct.pack_into("<Q", hdr, at, node.data_addr if node.data_size else _UNDEF)
                node.header_addr = self._emit(bytes(hdr))
                return
        attr_msgs = [_message(0x000C, _attr_msg(k, v)) for k, v in node.attrs.items()]
        if node.is_dataset:
            # message set, order, versions and flags of a dataset header as dolfin's HDF5 writes them (checked against
            # the reference's own dolfin-written files: tests/test_h5_structure.py)
            msgs = [
                _message(0x0001, _space_msg(node.shape)),
                _message(0x0003, _dtype_msg(node.dtype), flags=1),
                # fill value v2 as a serial dolfin run leaves it: space allocated late, fill written "if set", defined
                # with size 0 (a file written under MPI differs in two places: early allocation, layout flagged constant)
                _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0), flags=1),
                _message(0x0008, struct.pack("<BBQQ", 3, 1, node.data_addr if node.data_size else _UNDEF,
                                             node.data_size)),
                _message(0x0012, struct.pack("<B3xI", 1, _MTIME)),
            ] + attr_msgs
        else:
            for ch in node.children.values():
                self._write_node(ch)
            bt, hp = self._write_group_index(node)
            node._index = (bt, hp)  # type: ignore[attr-defined]
            msgs = [_message(0x0011, struct.pack("<QQ", bt, hp))] + attr_msgs
        blob = _object_header(msgs, node.nlinks)
        node.header_addr = self._emit(blob)
        if node.is_dataset and not node.attrs:
            at = blob.index(struct.pack("<BBQ", 3, 1, node.data_addr if node.data_size else _UNDEF)) + 2
            self._hdr_cache[(node.shape, node.dtype.str, node.data_size, node.nlinks)] = (blob, at)

    def close(self) -> None:
        if self._closed:
            return
        self._meta_base = self._alloc(0)
        self._meta = bytearray()
        self._write_node(self._root)
        bt, hp = self._root._index  # type: ignore[attr-defined]
        self._fh.write(self._meta)
        self._pos = self._meta_base + len(self._meta)
        eof = self._alloc(0)
        sb = _SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
        sb += struct.pack("<HHI", self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
        sb += struct.pack("<QQIIQQ", 0, self._root.header_addr, 1, 0, bt, hp)
        assert len(sb) == self._SB_SIZE
        self._fh.seek(0)
        self._fh.write(sb)
        self._fh.truncate(eof)
        self._fh.close()
        self._closed = True

    def __enter__(self) -> "H5Writer":
        return self

    def __exit__(self, *exc) -> None:
        self.close()
