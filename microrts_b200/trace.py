"""Reference-format game traces (rts.Trace / rts.TraceEntry, src/rts/Trace.java:95-181, src/rts/TraceEntry.java:106-164) for
games played on the device, so that the Java GUI and test/microrts/TestTracesIntegrity.java can consume them.

  record_traces(batch, games, max_cycles)  play the batch one cycle per launch and collect, per recorded game, the
                                           TraceEntry sequence of src/tests/GenerateTestTraces.java:101-134: the initial
                                           state, one entry per cycle in which actions were issued (state before the
                                           cycle + the actions), and the final state
  write_trace_zip(path, utt, trace)        rts.Trace.toZip layout: a zip with one XML entry, written with the reference's
                                           XMLWriter formatting (two-space indent, every element closed on its own line)
  read_trace_zip(path)                     the inverse (also reads the reference's own data/traces/**/trace_0.zip)

What a recorded entry holds: the assignments that exist after the cycle with issue time == the entry's time, in insertion
order.  Those are the actions as GameState.issue left them (a pair cancelled by the conflict policy appears as the two NONE
actions it became), which replays to the same states through issueSafe.  An action whose ETA is 1 is complete after the
cycle and is not seen (only NONE padding has ETA 1 under the stock unit type tables, and it has no effect).
This is host-side tooling on top of mrts_batch_export; it is not part of the stepping hot path.
"""
import io
import zipfile
import xml.etree.ElementTree as ET

import numpy as np

UTT_FIELDS = ["cost", "hp", "minDamage", "maxDamage", "attackRange", "produceTime", "moveTime", "attackTime", "harvestTime",
              "returnTime", "harvestAmount", "sightRadius"]
UTT_FLAGS = ["isResource", "isStockpile", "canHarvest", "canMove", "canAttack"]


class TraceEntry:
    def __init__(self, time, resources, units, actions):
        self.time = int(time)
        self.resources = (int(resources[0]), int(resources[1]))
        self.units = [tuple(int(v) for v in u) for u in units]        # (type id, ID, player, x, y, resources, hitpoints), list order
        self.actions = [tuple(int(v) for v in a) for a in actions]    # (unit ID, type, parameter(-1 none), x, y, unit type id(-1 none))


class Trace:
    def __init__(self, width, height, terrain, entries=None):
        self.width, self.height, self.terrain = int(width), int(height), terrain   # terrain: str of '0'/'1', row major
        self.entries = entries or []


def _entry_from_export(ex, g, time=None, with_actions_at=None):
    hdr = ex["header"][g]
    n = int(hdr[3])
    u = ex["units"][g, :n]
    units = [(r[0], r[6], r[1], r[2], r[3], r[4], r[5]) for r in u]
    actions = []
    if with_actions_at is not None:
        a = ex["actions"][g, :n]
        order = sorted((int(a[i][6]), i) for i in range(n) if u[i][7] and int(a[i][5]) == with_actions_at)
        for _rank, i in order:
            ty = int(a[i][0])
            actions.append((int(u[i][6]), ty, int(a[i][1]) if ty != 5 else -1, int(a[i][2]), int(a[i][3]), int(a[i][4])))
    return TraceEntry(hdr[0] if time is None else time, (hdr[1], hdr[2]), units, actions)


def record_traces(batch, games, max_cycles, terrain=None):
    """Play `batch` (policies already set) to game over / max_cycles, one cycle per launch, and return {game: Trace} for
    the listed games.  Every game of the batch advances; only the listed ones are exported each cycle."""
    games = list(games)
    lo, hi = min(games), max(games) + 1
    maps = batch.maps
    traces = {}
    for g in games:
        m = maps[g % len(maps)]
        ter = "".join(str(int(v)) for v in m.getTerrain().reshape(-1))
        traces[g] = Trace(m.getWidth(), m.getHeight(), ter)
    prev = batch.export(lo, hi - lo)
    live = set()
    for g in games:
        traces[g].entries.append(_entry_from_export(prev, g - lo))
        if not (prev["header"][g - lo][5] and prev["header"][g - lo][0] > 0):
            live.add(g)
    t_done = {}
    while live:
        batch.step(1, max_cycles)
        cur = batch.export(lo, hi - lo)
        for g in sorted(live):
            t_before = int(prev["header"][g - lo][0])
            e = _entry_from_export(cur, g - lo, with_actions_at=t_before)
            if e.actions:  # state BEFORE the cycle + the actions issued in it (GenerateTestTraces.java:111-116)
                pe = _entry_from_export(prev, g - lo)
                uid = {u[1] for u in pe.units}
                pe.actions = [a for a in e.actions if a[0] in uid]
                traces[g].entries.append(pe)
            h = cur["header"][g - lo]
            if h[5] or h[0] >= max_cycles or int(h[0]) == t_before:
                traces[g].entries.append(_entry_from_export(cur, g - lo))
                t_done[g] = int(h[0])
                live.discard(g)
        prev = cur
    return traces


# ---- XML in the reference's XMLWriter layout (src/util/XMLWriter.java) ---------------------------------------------------
class _W:
    def __init__(self):
        self.out, self.depth = io.StringIO(), 0

    def tag(self, name, attrs=None):
        if name.startswith("/"):
            self.depth -= 1
            self.out.write("  " * self.depth + "<" + name + ">\n")
        else:
            self.out.write("  " * self.depth + "<" + name + ((" " + attrs) if attrs else "") + ">\n")
            self.depth += 1

    def inline(self, name, text):
        self.out.write("  " * self.depth + "<%s>%s</%s>\n" % (name, text, name))


def _utt_xml(w, utt):
    types = utt.getUnitTypes()
    w.tag("rts.units.UnitTypeTable", 'moveConflictResolutionStrategy="%d"' % utt.getMoveConflictResolutionStrategy())
    for t in types:
        attrs = 'ID="%d" name="%s" ' % (t.ID, t.name) + " ".join('%s="%d"' % (f, getattr(t, f)) for f in UTT_FIELDS) + " " + \
                " ".join('%s="%s"' % (f, "true" if getattr(t, f) else "false") for f in UTT_FLAGS)
        w.tag("rts.units.UnitType", attrs)
        for p in t.produces:
            w.tag("produces", 'type="%s"' % types[p].name)
            w.tag("/produces")
        for o in types:
            if t.ID in o.produces:
                w.tag("producedBy", 'type="%s"' % o.name)
                w.tag("/producedBy")
        w.tag("/rts.units.UnitType")
    w.tag("/rts.units.UnitTypeTable")


def trace_to_xml(utt, trace):
    names = [t.name for t in utt.getUnitTypes()]
    w = _W()
    w.tag("rts.Trace")
    _utt_xml(w, utt)
    w.tag("entries")
    for e in trace.entries:
        w.tag("rts.TraceEntry", 'time = "%d"' % e.time)
        w.tag("rts.PhysicalGameState", 'width="%d" height="%d"' % (trace.width, trace.height))
        w.inline("terrain", trace.terrain)
        w.tag("players")
        for p in (0, 1):
            w.tag("rts.Player", 'ID="%d" resources="%d"' % (p, e.resources[p]))
            w.tag("/rts.Player")
        w.tag("/players")
        w.tag("units")
        for (ty, uid, pl, x, y, res, hp) in e.units:
            w.tag("rts.units.Unit", 'type="%s" ID="%d" player="%d" x="%d" y="%d" resources="%d" hitpoints="%d" ' % (names[ty], uid, pl, x, y, res, hp))
            w.tag("/rts.units.Unit")
        w.tag("/units")
        w.tag("/rts.PhysicalGameState")
        w.tag("actions")
        for (uid, ty, par, x, y, ut) in e.actions:
            w.tag("action", 'unitID="%d"' % uid)
            attrs = 'type="%d" ' % ty                                  # UnitAction.toxml, src/rts/UnitAction.java:544-561
            if ty == 5:
                attrs += 'x="%d" y="%d"' % (x, y)
            else:
                if par != -1:
                    attrs += 'parameter="%d"' % par + (" " if ut >= 0 else "")
                if ut >= 0:
                    attrs += 'unitType="%s"' % names[ut]
            w.tag("UnitAction", attrs)
            w.tag("/UnitAction")
            w.tag("/action")
        w.tag("/actions")
        w.tag("/rts.TraceEntry")
    w.tag("/entries")
    w.tag("/rts.Trace")
    return w.out.getvalue()


def write_trace_zip(path, utt, trace):
    import os
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr(os.path.basename(path), trace_to_xml(utt, trace))   # Trace.toZip names the entry after the file


def read_trace_zip(path):
    """-> (utt description dict, Trace).  utt description: conflict policy + per type the 12 fields, 5 flags, produced names."""
    with zipfile.ZipFile(path) as z:
        root = ET.fromstring(z.read(z.namelist()[0]))
    assert root.tag == "rts.Trace"
    ue = root.find("rts.units.UnitTypeTable")
    names = [t.get("name") for t in ue]
    utt = dict(conflict=int(ue.get("moveConflictResolutionStrategy")), types=[])
    for t in ue:
        utt["types"].append(dict(name=t.get("name"), fields=[int(t.get(f)) for f in UTT_FIELDS], flags=[t.get(f) == "true" for f in UTT_FLAGS],
                                 produces=[c.get("type") for c in t if c.tag == "produces"]))
    trace = None
    for e in root.find("entries"):
        pg = e.find("rts.PhysicalGameState")
        if trace is None:
            trace = Trace(pg.get("width"), pg.get("height"), pg.find("terrain").text.strip())
        res = {int(p.get("ID")): int(p.get("resources")) for p in pg.find("players")}
        units = [(names.index(u.get("type")), int(u.get("ID")), int(u.get("player")), int(u.get("x")), int(u.get("y")), int(u.get("resources")),
                  int(u.get("hitpoints"))) for u in pg.find("units")]
        acts = []
        for a in e.find("actions"):
            ua = a.find("UnitAction")
            ut = ua.get("unitType")
            acts.append((int(a.get("unitID")), int(ua.get("type")), int(ua.get("parameter", "-1")), int(ua.get("x", "0")), int(ua.get("y", "0")),
                         names.index(ut) if ut is not None else -1))
        trace.entries.append(TraceEntry(int(e.get("time")), (res.get(0, 0), res.get(1, 0)), units, acts))
    return utt, trace
