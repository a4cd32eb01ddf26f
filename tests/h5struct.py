"""Structural fingerprint of an HDF5 file (TEST INFRASTRUCTURE; independent of ``vasp_b200.h5lite``'s reader).

Walks the on-disk structures directly -- superblock, object headers, header messages, group B-trees, symbol-table
nodes, local heaps -- and records everything about them that is *format*, not *content*: versions, flags, K values,
message types in order, datatype / dataspace / fill-value / layout / attribute encodings.  Addresses, sizes of raw
data and the data itself are left out.  ``tests/test_h5_structure.py`` compares the fingerprint of files written by
``H5Writer`` with the fingerprint of the files legacy dolfin (HDF5 1.12, ``libver=earliest``) wrote for the reference's
own test-suite (``/root/reference/tests/test_data``; stored under ``tests/golden/h5_structure.json`` by
``tests/golden/make_h5_structure_golden.py`` so that nothing is read from the reference at test time).
"""
from __future__ import annotations

import struct
from typing import Dict, List

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


def _u(buf, fmt, pos):
    return struct.unpack_from("<" + fmt, buf, pos)


def _datatype(buf, p) -> Dict:
    cv, b0, b1, b2 = buf[p], buf[p + 1], buf[p + 2], buf[p + 3]
    size = _u(buf, "I", p + 4)[0]
    out = {"class": cv & 0x0F, "version": cv >> 4, "bits": [b0, b1, b2], "size": size}
    cls = cv & 0x0F
    if cls == 0:
        out["props"] = list(_u(buf, "HH", p + 8))                       # bit offset, precision
    elif cls == 1:
        out["props"] = list(_u(buf, "HHBBBBI", p + 8))                  # offset, precision, exp loc/size, mant loc/size, bias
    return out


def _dataspace(buf, p) -> Dict:
    ver, rank, flags = buf[p], buf[p + 1], buf[p + 2]
    out = {"version": ver, "rank": rank, "flags": flags}
    if ver == 2:
        out["space_type"] = buf[p + 3]
    return out


def _message(buf, mtype, flags, p, size) -> Dict:
    m: Dict = {"type": mtype, "msg_flags": flags, "msg_size": size}
    if mtype == 0x0001:
        m.update(_dataspace(buf, p))
    elif mtype == 0x0003:
        m.update(_datatype(buf, p))
    elif mtype == 0x0005:  # fill value
        ver = buf[p]
        m["version"] = ver
        if ver in (1, 2):
            m.update(alloc_time=buf[p + 1], write_time=buf[p + 2], defined=buf[p + 3])
        elif ver == 3:
            m["fv_flags"] = buf[p + 1]
    elif mtype == 0x0008:  # layout
        m.update(version=buf[p], layout_class=buf[p + 1])
    elif mtype == 0x000C:  # attribute
        ver = buf[p]
        nsz, tsz, ssz = _u(buf, "HHH", p + 2)
        q = p + 8 + (1 if ver == 3 else 0)
        pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
        name = bytes(buf[q:q + nsz]).split(b"\0", 1)[0].decode()
        q += pad(nsz)
        dt = _datatype(buf, q)
        q += pad(tsz)
        sp = _dataspace(buf, q)
        m.update(version=ver, attr_flags=buf[p + 1], name=name, name_size=nsz, datatype=dt, datatype_size=tsz,
                 dataspace=sp, dataspace_size=ssz)
    elif mtype == 0x0011:  # symbol table
        pass
    elif mtype == 0x0012:  # object modification time
        m["version"] = buf[p]
    return m


def _object_header(buf, base, addr) -> Dict:
    p = base + addr
    ver = buf[p]
    if ver != 1:
        return {"header_version": "v2" if bytes(buf[p:p + 4]) == b"OHDR" else int(ver)}
    nmsg = _u(buf, "H", p + 2)[0]
    refcount, hsize = _u(buf, "II", p + 4)
    blocks = [(p + 16, hsize)]
    msgs: List[Dict] = []
    raw = []
    seen = 0
    while blocks and seen < nmsg:
        q, left = blocks.pop(0)
        end = q + left
        while q + 8 <= end and seen < nmsg:
            mtype, msize, mflags = _u(buf, "HHB", q)
            body = q + 8
            seen += 1
            if mtype == 0x0010:
                coff, clen = _u(buf, "QQ", body)
                blocks.append((base + coff, clen))
            elif mtype != 0x0000:
                msgs.append(_message(buf, mtype, mflags, body, msize))
                raw.append((mtype, body, msize))
            q = body + msize
    return {"header_version": 1, "refcount": refcount, "messages": msgs, "_raw": raw,
            "first_message_aligned": (p + 16) % 8 == 0}


def fingerprint(path) -> Dict:
    buf = open(path, "rb").read()
    sb = buf.find(SIG)
    assert sb in (0, 512, 1024, 2048), "no HDF5 signature"
    ver = buf[sb + 8]
    out: Dict = {"superblock": {"version": ver, "free_space_version": buf[sb + 9], "root_symtab_version": buf[sb + 10],
                                "shared_header_version": buf[sb + 12], "size_of_offsets": buf[sb + 13],
                                "size_of_lengths": buf[sb + 14]}}
    leaf_k, internal_k, flags = _u(buf, "HHI", sb + 16)
    out["superblock"].update(group_leaf_k=leaf_k, group_internal_k=internal_k, consistency_flags=flags)
    p = sb + 24 + (4 if ver == 1 else 0)
    base, free, eof, drv = _u(buf, "QQQQ", p)
    out["superblock"].update(base_address=base, free_space_undefined=free == UNDEF, driver_undefined=drv == UNDEF,
                             eof_is_file_size=eof == len(buf) - sb)
    noff, root, cache_type = _u(buf, "QQI", p + 32)
    out["superblock"]["root_cache_type"] = cache_type
    objects: Dict[str, Dict] = {}

    def walk_group(name: str, addr: int) -> None:
        oh = _object_header(buf, base, addr)
        raw = oh.pop("_raw", [])
        objects[name] = oh
        st = [r for r in raw if r[0] == 0x0011]
        if not st:
            oh["kind"] = "dataset" if any(r[0] == 0x0008 for r in raw) else "other"
            return
        oh["kind"] = "group"
        bt, hp = _u(buf, "QQ", st[0][1])
        h = base + hp
        assert bytes(buf[h:h + 4]) == b"HEAP"
        hver = buf[h + 4]
        dsize, free_head, daddr = _u(buf, "QQQ", h + 8)
        oh["heap"] = {"version": hver, "data_size_multiple_of_8": dsize % 8 == 0,
                      "free_list": "none" if free_head == UNDEF or free_head == 1 else "block",
                      "first_name_offset": 8}
        names: Dict[str, tuple] = {}
        levels = []
        snods = []

        def walk_tree(a: int) -> None:
            q = base + a
            sig = bytes(buf[q:q + 4])
            if sig == b"TREE":
                ntype, level, used = _u(buf, "BBH", q + 4)
                assert ntype == 0
                levels.append(level)
                for i in range(used):
                    walk_tree(_u(buf, "Q", q + 24 + 8 + 16 * i)[0])
            else:
                assert sig == b"SNOD", sig
                sver, nsym = buf[q + 4], _u(buf, "H", q + 6)[0]
                snods.append((sver, nsym))
                for i in range(nsym):
                    e = q + 8 + 40 * i
                    no, oa, ct = _u(buf, "QQI", e)
                    s = base + daddr + no
                    nm = bytes(buf[s:buf.find(b"\0", s)]).decode()
                    names[nm] = (oa, ct)

        if bt != UNDEF:
            walk_tree(bt)
        oh["btree"] = {"depth": (max(levels) + 1) if levels else 0,
                       "snod_version": sorted({s[0] for s in snods}),
                       "snod_max_entries": max((s[1] for s in snods), default=0),
                       "snod_within_2k": all(s[1] <= 2 * leaf_k for s in snods),
                       "names_sorted": list(names) == sorted(names, key=lambda s: s.encode())}
        oh["child_cache_types"] = sorted({ct for _, ct in names.values()})
        for nm, (oa, ct) in names.items():
            walk_group(f"{name.rstrip('/')}/{nm}", oa)

    walk_group("/", root)
    out["objects"] = objects
    return out


def object_signature(obj: Dict) -> Dict:
    """What must agree between a dolfin-written object and ours: kind, header version, and per message type the
    encoding (order of messages is recorded separately)."""
    sig = {"kind": obj.get("kind"), "header_version": obj["header_version"],
           "message_order": [m["type"] for m in obj.get("messages", []) if m["type"] != 0x000C],
           "messages": {}, "attributes": {}}
    for m in obj.get("messages", []):
        if m["type"] == 0x000C:
            a = {k: m[k] for k in ("version", "attr_flags", "datatype", "dataspace", "msg_flags")}
            a["name_size_includes_nul"] = m["name_size"] == len(m["name"]) + 1
            sig["attributes"][m["name"]] = a
        else:
            sig["messages"][m["type"]] = {k: v for k, v in m.items() if k not in ("type", "msg_size")}
            if m["type"] in (0x0005, 0x0008, 0x0012):
                sig["messages"][m["type"]]["msg_size"] = m["msg_size"]
    for k in ("heap", "btree", "child_cache_types"):
        if k in obj:
            sig[k] = obj[k]
    return sig
