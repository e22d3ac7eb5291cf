"""Reference-format traces (microrts_b200.trace): the reader and writer are pinned against the reference's own trace files,
and games recorded on the device replay state-for-state through the oracle (TestTracesIntegrity protocol)."""
import glob
import os
import zipfile

import numpy as np
import pytest

import microrts_b200 as M
import parity as P
from microrts_b200 import trace as T
from oracle import oracle as O

REF_TRACES = "/root/reference/data/traces"


def as_golden(name, mapkey, uttd, tr):
    """A trace in the dict form of tests/golden_io.load_traces (what the replay harnesses take)."""
    entries = []
    for e in tr.entries:
        idx = {u[1]: i for i, u in enumerate(e.units)}
        entries.append(dict(time=e.time, res=e.resources, units=[(u[0], u[2], u[3], u[4], u[5], u[6]) for u in e.units],
                            actions=[(idx[a[0]], a[1], a[2], a[3], a[4], a[5]) for a in e.actions]))
    return dict(name=name, mapkey=mapkey, conflict=uttd["conflict"], types=uttd["types"], entries=entries)


@pytest.mark.skipif(not os.path.isdir(REF_TRACES), reason="reference checkout not present")
def test_reader_and_writer_against_reference_files(traces):
    """read_trace_zip parses the reference's files to exactly the committed golden data, and trace_to_xml reproduces the
    reference's XML text byte for byte (Trace.toxml / TraceEntry.toxml / UnitAction.toxml / XMLWriter layout)."""
    by_name = {t["name"]: t for t in traces}
    utt = M.UnitTypeTable(1, 1)
    paths = sorted(glob.glob(os.path.join(REF_TRACES, "**", "trace_0.zip"), recursive=True))
    assert len(paths) == 280
    for p in paths[::9]:
        rel = os.path.relpath(os.path.dirname(p), REF_TRACES)
        uttd, tr = T.read_trace_zip(p)
        g = by_name[rel]
        assert len(tr.entries) == len(g["entries"])
        for e, ge in zip(tr.entries, g["entries"]):
            assert e.time == ge["time"] and e.resources == tuple(ge["res"])
            assert [(u[0], u[2], u[3], u[4], u[5], u[6]) for u in e.units] == [tuple(u) for u in ge["units"]]
            idx = {u[1]: i for i, u in enumerate(e.units)}
            assert [(idx[a[0]],) + a[1:] for a in e.actions] == [tuple(a) for a in ge["actions"]]
        with zipfile.ZipFile(p) as z:
            original = z.read(z.namelist()[0]).decode()
        assert T.trace_to_xml(utt, tr) == original, rel


def test_zip_round_trip(traces, tmp_path):
    utt = M.UnitTypeTable(1, 1)
    t = traces[3]
    tr = T.Trace(4, 4, "0" * 16)
    for e in t["entries"][:5]:
        units = [(u[0], 100 + i, u[1], u[2], u[3], u[4], u[5]) for i, u in enumerate(e["units"])]
        tr.entries.append(T.TraceEntry(e["time"], e["res"], units, [(100 + a[0],) + tuple(a[1:]) for a in e["actions"]]))
    path = str(tmp_path / "trace_0.zip")
    T.write_trace_zip(path, utt, tr)
    uttd, back = T.read_trace_zip(path)
    assert uttd["conflict"] == 1 and [x["name"] for x in uttd["types"]] == O.TYPE_NAMES
    assert len(back.entries) == len(tr.entries)
    for a, b in zip(tr.entries, back.entries):
        assert (a.time, a.resources, a.units, a.actions) == (b.time, b.resources, b.units, b.actions)


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["8x8/basesWorkers8x8", "16x16/basesWorkers16x16"])
def test_device_games_recorded_as_reference_traces(backend, traces, maps, tmp_path, key):
    """LightRush mirror matches recorded on the device: the trace file replays state-for-state through the oracle, and its
    entries sit at the same times with the same unit lists as the reference's recorded game of the same setup."""
    golden = [t for t in traces if t["mapkey"] == key and "LightRush" in t["name"]][0]
    utt = M.UnitTypeTable(1, 1)
    b = M.BatchedGameState(utt, M.PhysicalGameState.fromXML(P.map_to_xml(maps[key]), utt), 2, scripted_ai=True)
    b.set_policy(0, M.POLICY_LIGHT_RUSH)
    b.set_policy(1, M.POLICY_LIGHT_RUSH)
    limit = 120 if backend == "emu" else 500
    recorded = T.record_traces(b, [1], limit)[1]
    b.close()
    path = str(tmp_path / "trace_0.zip")
    T.write_trace_zip(path, utt, recorded)
    uttd, tr = T.read_trace_zip(path)
    types = [([f for f in x["fields"]], sum(int(v) << i for i, v in enumerate(x["flags"])), [O.TYPE_NAMES.index(n) for n in x["produces"]]) for x in uttd["types"]]
    mine = as_golden("device", key, dict(conflict=uttd["conflict"], types=types), tr)
    from test_oracle_golden import replay
    g = replay(mine, maps)
    assert g.time == tr.entries[-1].time
    # same decision times and unit lists as the reference's own recording (IDs differ: the reference's are global counters)
    ref_entries = [e for e in golden["entries"] if e["time"] < limit or limit == 500]
    got = mine["entries"] if limit == 500 else mine["entries"][:-1]
    ref_entries = ref_entries if limit == 500 else ref_entries[:len(got)]
    assert [e["time"] for e in got] == [e["time"] for e in ref_entries]
    for a, r in zip(got, ref_entries):
        assert [tuple(u) for u in a["units"]] == [tuple(u) for u in r["units"]], a["time"]
