#!/usr/bin/env python3
"""Build the golden fixture packs from the reference checkout (run in the build container only).

Inputs (read-only, never copied verbatim):
  /root/reference/maps/**/*.xml            -- every map the reference ships (TestLoadingMaps fixture set)
  /root/reference/data/traces/**/trace_0.zip -- the 280 recorded games replayed by
                                               test/microrts/TestTracesIntegrity.java:72-127
Outputs (committed):
  tests/golden/maps.pack.gz    gzip(JSON): {key: {w,h,terrain,players,units}}  key = path under maps/ without .xml
  tests/golden/traces.pack.gz  gzip(binary), layout documented in tests/golden_io.py (MRTSGOLD1)

The packs carry parsed *data* only (unit tuples, action tuples); no reference source text.
Usage: python tests/golden/make_golden.py [/root/reference]
"""
import gzip, io, json, os, shutil, struct, sys, zipfile
import xml.etree.ElementTree as ET

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
TYPE_NAMES = ["Resource", "Base", "Barracks", "Worker", "Light", "Heavy", "Ranged"]


def parse_pgs(e):
    w = int(e.get("width")); h = int(e.get("height"))
    terrain = e.find("terrain").text.strip()
    assert len(terrain) == w * h and set(terrain) <= {"0", "1"}, (w, h, len(terrain))
    players = [[int(p.get("ID")), int(p.get("resources"))] for p in e.find("players")]
    units = []
    for u in e.find("units"):
        units.append([u.get("type"), int(u.get("ID")), int(u.get("player")), int(u.get("x")), int(u.get("y")),
                      int(u.get("resources")), int(u.get("hitpoints"))])
    return dict(w=w, h=h, terrain=terrain, players=players, units=units)


def build_maps():
    maps = {}
    root = os.path.join(REF, "maps")
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if not f.endswith(".xml"):
                continue
            p = os.path.join(d, f)
            key = os.path.relpath(p, root)[:-4]
            e = ET.parse(p).getroot()
            if e.tag != "rts.PhysicalGameState":
                continue
            maps[key] = parse_pgs(e)
    with gzip.GzipFile(os.path.join(OUT, "maps.pack.gz"), "wb", mtime=0) as g:
        g.write(json.dumps(maps, sort_keys=True, separators=(",", ":")).encode())
    # the same pack ships with the package (microrts_b200/maps.py: standard_map) so that bench.py needs nothing under tests/
    shutil.copyfile(os.path.join(OUT, "maps.pack.gz"), os.path.join(OUT, "..", "..", "microrts_b200", "data", "maps.pack.gz"))
    print("maps:", len(maps))
    return maps


UTT_FIELDS = ["cost", "hp", "minDamage", "maxDamage", "attackRange", "produceTime", "moveTime", "attackTime",
              "harvestTime", "returnTime", "harvestAmount", "sightRadius"]
UTT_FLAGS = ["isResource", "isStockpile", "canHarvest", "canMove", "canAttack"]


def pack_str(b, s):
    s = s.encode()
    b.write(struct.pack("<H", len(s))); b.write(s)


def build_traces():
    root = os.path.join(REF, "data", "traces")
    paths = []
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".zip"):
                paths.append(os.path.join(d, f))
    paths.sort()
    b = io.BytesIO()
    b.write(b"MRTSGOLD1")
    b.write(struct.pack("<I", len(paths)))
    n_entries = 0
    for p in paths:
        rel = os.path.relpath(os.path.dirname(p), root)          # <mapdir>/<mapname>/<agent>
        mapkey = os.path.dirname(rel)                            # TestTracesIntegrity.java:55-58
        z = zipfile.ZipFile(p)
        names = z.namelist(); assert len(names) == 1
        tr = ET.fromstring(z.read(names[0]))
        assert tr.tag == "rts.Trace"
        ue = tr.find("rts.units.UnitTypeTable")
        pack_str(b, rel); pack_str(b, mapkey)
        b.write(struct.pack("<B", int(ue.get("moveConflictResolutionStrategy"))))
        types = list(ue)
        assert [t.get("name") for t in types] == TYPE_NAMES
        for t in types:
            assert int(t.get("ID")) == TYPE_NAMES.index(t.get("name"))
            b.write(struct.pack("<12h", *[int(t.get(f)) for f in UTT_FIELDS]))
            flags = 0
            for i, f in enumerate(UTT_FLAGS):
                flags |= (t.get(f) == "true") << i
            prod = [TYPE_NAMES.index(c.get("type")) for c in t if c.tag == "produces"]
            b.write(struct.pack("<BB", flags, len(prod))); b.write(bytes(prod))
        entries = list(tr.find("entries"))
        b.write(struct.pack("<I", len(entries)))
        for e in entries:
            n_entries += 1
            pgs = parse_pgs(e.find("rts.PhysicalGameState"))
            b.write(struct.pack("<iiiH", int(e.get("time")), pgs["players"][0][1], pgs["players"][1][1], len(pgs["units"])))
            idx = {}
            for i, (tn, uid, pl, x, y, res, hp) in enumerate(pgs["units"]):
                assert uid not in idx
                idx[uid] = i
                b.write(struct.pack("<BbBBhh", TYPE_NAMES.index(tn), pl, x, y, res, hp))
            acts = list(e.find("actions"))
            b.write(struct.pack("<H", len(acts)))
            for a in acts:
                ua = a.find("UnitAction")
                ut = ua.get("unitType")
                b.write(struct.pack("<HBhhhb", idx[int(a.get("unitID"))], int(ua.get("type")),
                                    int(ua.get("parameter", "-1")), int(ua.get("x", "0")), int(ua.get("y", "0")),
                                    TYPE_NAMES.index(ut) if ut is not None else -1))
    with gzip.GzipFile(os.path.join(OUT, "traces.pack.gz"), "wb", mtime=0) as g:
        g.write(b.getvalue())
    print("traces:", len(paths), "entries:", n_entries, "raw bytes:", b.tell())


if __name__ == "__main__":
    build_maps()
    build_traces()
