"""CPU-side checks of the product library: it loads, exports every symbol include/microrts_cuda.h declares, and the
host-only entry points (unit type table, map loading) behave like the reference's.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import parity as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from microrts_b200 import _ffi
    return _ffi.lib()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "microrts_cuda.h")).read()
    names = set(re.findall(r"\b(mrts_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 40
    raw = ctypes.CDLL(os.path.join(ROOT, "microrts_b200", "libmicrorts_cuda.so"))
    missing = [n for n in sorted(names) if not hasattr(raw, n)]
    assert not missing, missing
    assert lib.mrts_abi_version() == 2


def test_no_oracle_or_cpu_path_in_product():
    """The product must not route through the oracle: nothing under microrts_b200/ may import, link or load it."""
    for d, _, files in os.walk(os.path.join(ROOT, "microrts_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|mrts_oracle|liboracle|o_game_", txt), f


def test_unit_type_table_matches_golden_utt(lib, traces):
    import microrts_b200 as M
    utt = M.UnitTypeTable(1, 1)
    for tid, (fields, flags, prod) in enumerate(traces[0]["types"]):
        t = utt.getUnitType(tid)
        assert [getattr(t, f) for f in t.FIELDS] == list(fields)
        got = t.isResource | t.isStockpile << 1 | t.canHarvest << 2 | t.canMove << 3 | t.canAttack << 4
        assert got == flags and t.produces == list(prod)
    assert utt.getMaxAttackRange() == 3
    v2, v3 = M.UnitTypeTable(2, 1), M.UnitTypeTable(3, 2)
    assert v2.getUnitType("Base").produceTime == 200 and v3.getUnitType("Base").produceTime == 10
    assert v2.getUnitType("Heavy").hp == 8 and v3.getUnitType("Worker").maxDamage == 2
    assert v3.getMoveConflictResolutionStrategy() == 2


def test_utt_versions_match_oracle(lib):
    import microrts_b200 as M
    from oracle import oracle as O
    for v in (1, 2, 3):
        a, b = M.UnitTypeTable(v, 1), O.Utt(v, 1)
        for tid in range(7):
            t = a.getUnitType(tid)
            assert [getattr(t, f) for f in t.FIELDS] == [b.field(tid, k) for k in range(12)]
            assert t.produces == b.produces(tid)


def test_every_map_loads(lib, maps, tmp_path):
    """TestLoadingMaps (test/microrts/TestLoadingMaps.java): every map parses; contents equal the golden pack."""
    import microrts_b200 as M
    utt = M.UnitTypeTable(1, 1)
    assert len(maps) == 140
    for key, m in maps.items():
        path = tmp_path / (key.replace("/", "_") + ".xml")
        path.write_text(P.map_to_xml(m))
        pgs = M.PhysicalGameState.load(str(path), utt)
        assert (pgs.getWidth(), pgs.getHeight()) == (m["w"], m["h"])
        exp = np.array([[P.TYPE_NAMES.index(u[0])] + list(u[1:]) for u in m["units"]], dtype=np.int64).reshape(-1, 7)
        got = pgs.getUnits().astype(np.int64)
        assert got.shape == exp.shape and (got == exp).all(), key
        assert "".join(str(v) for v in pgs.getTerrain().reshape(-1)) == m["terrain"]
        assert [pgs.getPlayerResources(0), pgs.getPlayerResources(1)] == [p[1] for p in m["players"]]


@pytest.mark.skipif(not os.path.isdir("/root/reference/maps"), reason="reference checkout not present")
def test_loader_on_reference_files(lib, maps):
    """The loader on the reference's own XML files (raw formatting), when the checkout is available."""
    import microrts_b200 as M
    utt = M.UnitTypeTable(1, 1)
    for key, m in maps.items():
        pgs = M.PhysicalGameState.load("/root/reference/maps/%s.xml" % key, utt)
        assert pgs.getUnits().shape[0] == len(m["units"]) and pgs.getWidth() == m["w"]


def test_map_errors(lib):
    import microrts_b200 as M
    utt = M.UnitTypeTable(1, 1)
    with pytest.raises(M.MicroRTSError):
        M.PhysicalGameState.load("/nonexistent/map.xml", utt)
    with pytest.raises(M.MicroRTSError):
        M.PhysicalGameState.fromXML("<rts.PhysicalGameState width=\"2\" height=\"2\"><terrain>000</terrain></rts.PhysicalGameState>", utt)
    two = {"w": 2, "h": 2, "terrain": "0000", "players": [[0, 5], [1, 5]],
           "units": [["Base", 1, 0, 0, 0, 0, 10], ["Base", 2, 1, 0, 0, 0, 10]]}
    with pytest.raises(M.MicroRTSError):  # PhysicalGameState.addUnit: two units in one cell
        M.PhysicalGameState.fromXML(P.map_to_xml(two), utt)


def test_utt_from_json(lib):
    import json
    import microrts_b200 as M
    base = M.UnitTypeTable(2, 3)
    doc = {"moveConflictResolutionStrategy": 3, "unitTypes": []}
    for t in base.getUnitTypes():
        d = {f: getattr(t, f) for f in t.FIELDS}
        d.update(ID=t.ID, name=t.name, isResource=t.isResource, isStockpile=t.isStockpile, canHarvest=t.canHarvest,
                 canMove=t.canMove, canAttack=t.canAttack, produces=[base.getUnitType(p).name for p in t.produces], producedBy=[])
        doc["unitTypes"].append(d)
    u = M.UnitTypeTable.fromJSON(json.dumps(doc))
    assert u.getMoveConflictResolutionStrategy() == 3
    w = u.getUnitType("Worker")
    # the reference reads harvestTime from the "produceTime" key and leaves returnTime at 10 (UnitType.java:224-228)
    assert w.harvestTime == w.produceTime == 50 and w.returnTime == 10 and w.produces == [1, 2]
    with pytest.raises(M.MicroRTSError):
        M.UnitTypeTable.fromJSON("{not json")


def test_compute_fails_loudly_without_gpu(lib, maps):
    """No CPU fallback: creating a batch without a CUDA device is an error, not a silent host path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import microrts_b200 as M
    utt = M.UnitTypeTable(1, 1)
    pgs = M.PhysicalGameState.fromXML(P.map_to_xml(maps["8x8/basesWorkers8x8"]), utt)
    with pytest.raises(M.MicroRTSError):
        M.BatchedGameState(utt, pgs, 4)
