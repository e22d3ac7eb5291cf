"""The reference's map set as shipped with the package, and the map text format.

  load_maps()      {key: {w, h, terrain, players, units}} for every map under the reference's maps/ directory (key = path without
                   ".xml"), re-encoded by tests/golden/make_golden.py into microrts_b200/data/maps.pack.gz (gzip JSON)
  map_to_xml(m)    the reference's map file text (PhysicalGameState.toxml, src/rts/PhysicalGameState.java:600-640 layout)
  standard_map()   a PhysicalGameState handle for one of them (what PhysicalGameState.load("maps/<key>.xml", utt) gives the reference)
"""
import gzip
import json
import os

_PACK = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "maps.pack.gz")
_maps = None


def load_maps():
    global _maps
    if _maps is None:
        with gzip.open(_PACK, "rb") as f:
            _maps = json.loads(f.read().decode())
    return _maps


def map_to_xml(m):
    s = ['<rts.PhysicalGameState width="%d" height="%d">' % (m["w"], m["h"]), "  <terrain>%s</terrain>" % m["terrain"], "  <players>"]
    for pid, res in m["players"]:
        s.append('    <rts.Player ID="%d" resources="%d">\n    </rts.Player>' % (pid, res))
    s.append("  </players>\n  <units>")
    for (tn, uid, pl, x, y, res, hp) in m["units"]:
        s.append('    <rts.units.Unit type="%s" ID="%d" player="%d" x="%d" y="%d" resources="%d" hitpoints="%d" >\n    </rts.units.Unit>'
                 % (tn, uid, pl, x, y, res, hp))
    s.append("  </units>\n</rts.PhysicalGameState>")
    return "\n".join(s) + "\n"


def standard_map(key, utt):
    from .api import PhysicalGameState
    return PhysicalGameState.fromXML(map_to_xml(load_maps()[key]), utt)
