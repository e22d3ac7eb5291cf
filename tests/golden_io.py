"""Readers for the golden fixture packs written by tests/golden/make_golden.py.

traces.pack.gz (gzip, little endian), magic b"MRTSGOLD1":
  u32 n_traces; per trace:
    str name, str mapkey            (str = u16 length + utf8)
    u8 conflict_policy; 7 unit types: 12 x i16 fields (cost hp minDamage maxDamage attackRange produceTime moveTime
      attackTime harvestTime returnTime harvestAmount sightRadius), u8 flags (isResource|isStockpile<<1|canHarvest<<2|
      canMove<<3|canAttack<<4), u8 n_produces, u8[n] produced type ids
    u32 n_entries; per entry: i32 time, i32 res0, i32 res1, u16 n_units, units (u8 type, i8 player, u8 x, u8 y, i16 res,
      i16 hp), u16 n_actions, actions (u16 unit list index in this snapshot, u8 type, i16 parameter(-1 = absent), i16 x,
      i16 y, i8 unitType(-1 = none))
"""
import gzip
import json
import os
import struct

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_maps():
    with gzip.open(os.path.join(HERE, "maps.pack.gz"), "rb") as f:
        return json.loads(f.read().decode())


class _R:
    def __init__(self, b):
        self.b, self.o = b, 0

    def u(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v

    def s(self):
        (n,) = self.u("H")
        v = self.b[self.o:self.o + n].decode()
        self.o += n
        return v


def load_traces(filter_fn=None):
    with gzip.open(os.path.join(HERE, "traces.pack.gz"), "rb") as f:
        r = _R(f.read())
    assert r.b[:9] == b"MRTSGOLD1"
    r.o = 9
    (nt,) = r.u("I")
    out = []
    for _ in range(nt):
        name, mapkey = r.s(), r.s()
        (conflict,) = r.u("B")
        types = []
        for _t in range(7):
            fields = list(r.u("12h"))
            flags, npd = r.u("BB")
            prod = list(r.b[r.o:r.o + npd]); r.o += npd
            types.append((fields, flags, prod))
        (ne,) = r.u("I")
        entries = []
        for _e in range(ne):
            time, r0, r1, nu = r.u("iiiH")
            units = [r.u("BbBBhh") for _ in range(nu)]
            (na,) = r.u("H")
            acts = [r.u("HBhhhb") for _ in range(na)]
            entries.append(dict(time=time, res=(r0, r1), units=units, actions=acts))
        t = dict(name=name, mapkey=mapkey, conflict=conflict, types=types, entries=entries)
        if filter_fn is None or filter_fn(t):
            out.append(t)
    assert r.o == len(r.b)
    return out
