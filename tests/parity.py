"""Shared helpers for the parity tests: compare the batched engine (through the C ABI) with the CPU oracle."""
import numpy as np

from oracle import oracle as O

TYPE_NAMES = O.TYPE_NAMES


from microrts_b200.maps import map_to_xml  # noqa: E402,F401  (the map text format lives in the package)


def export_game(ex, g):
    """(header, units[n,8], actions[n,8]) of game g from BatchedGameState.export()."""
    n = int(ex["header"][g, 3])
    return ex["header"][g], ex["units"][g, :n], ex["actions"][g, :n]


def oracle_snapshot(game):
    """Same shape as export_game from an oracle Game: units [n,8] = type,player,x,y,res,hp,id,has; actions [n,8]."""
    u = game.units()
    a = game.assignments()
    n = len(u)
    units = np.zeros((n, 8), dtype=np.int32)
    acts = np.zeros((n, 8), dtype=np.int32)
    if n:
        units[:, :6] = u[:, :6]
        units[:, 6] = u[:, 6]
        units[:, 7] = a[:, 0]
        acts[:, 0:5] = a[:, 1:6]
        acts[:, 5] = a[:, 6]
        acts[:, 6] = a[:, 7]
        acts[a[:, 0] == 0] = 0
    hdr = np.array([game.time, game.resources(0), game.resources(1), n, game.winner, int(game.gameover), 0, 0], dtype=np.int32)
    return hdr, units, acts


def assert_same_state(ex, g, game, ctx="", check_ids=True, check_actions=True):
    h, u, a = export_game(ex, g)
    oh, ou, oa = oracle_snapshot(game)
    msg = "%s game %d time dev=%d oracle=%d" % (ctx, g, h[0], oh[0])
    assert (h[:6] == oh[:6]).all(), "header mismatch %s: dev=%s oracle=%s" % (msg, h[:6], oh[:6])
    cols = slice(0, 8) if check_ids else [0, 1, 2, 3, 4, 5, 7]
    if not check_actions:
        cols = slice(0, 6)
    assert u.shape == ou.shape and (u[:, cols] == ou[:, cols]).all(), "unit mismatch %s\ndev=\n%s\noracle=\n%s" % (msg, u, ou)
    if check_actions:
        assert (a[:, :7] == oa[:, :7]).all(), "assignment mismatch %s\ndev=\n%s\noracle=\n%s" % (msg, a[:, :7], oa[:, :7])


def raw_rows(pairs, units, w, max_k):
    """(unit_list_index, (type,param,x,y,utype)) pairs -> RAW rows [max_k][8] addressed by the unit's cell."""
    rows = np.zeros((max_k, 8), dtype=np.int32)
    for k, (ui, (ty, par, x, y, ut)) in enumerate(pairs):
        rows[k] = [units[ui][2] + units[ui][3] * w, ty, par, x, y, ut, 0, 0]
    return rows
