// player_actions.hpp -- the host half of GameState.getPlayerActions / PlayerActionGenerator ("cartesian product
// stays host-side") on top of the ordered unit action lists the device enumerates (mrts_batch_unit_actions).
//
//   UnitAction.resourceUsage        src/rts/UnitAction.java:246-296      -> act_usage
//   ResourceUsage.consistentWith    src/rts/ResourceUsage.java:31-50     -> HRu::consistent_with
//   PlayerAction.cartesianProduct   src/rts/PlayerAction.java:180-195    -> player_actions
//   GameState.getPlayerActions      src/rts/GameState.java:493-524       -> player_actions
//   PlayerActionGenerator           src/rts/PlayerActionGenerator.java:56-106 (constructor), :114-121 (randomizeOrder),
//                                   :128-140 (incrementCurrentChoice), :148-195 (getNextAction), :201-222 (getRandom), :229-252 (getActionIndex)
//   java.util.Random                JDK (Java SE specification: 48-bit LCG)  -> JavaRandom
// The reference's generators are unseeded statics; here every draw comes from a JavaRandom the caller seeds.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "host_model.hpp"

namespace mrts {

struct JavaRandom {
    uint64_t s;
    explicit JavaRandom(int64_t seed = 0) : s(jr_scramble(seed)) {}
    int next(int bits) { s = (s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1); return (int)((int64_t)s >> (48 - bits)); }
    int nextInt(int bound) {
        int r = next(31), m = bound - 1;
        if ((bound & m) == 0) return (int)(((int64_t)bound * (int64_t)r) >> 31);
        for (int u = r; (int)((unsigned)u - (unsigned)(r = u % bound) + (unsigned)m) < 0; u = next(31)) {}
        return r;
    }
    double nextDouble() { int64_t a = next(26), b = next(27); return (double)((a << 27) + b) * 0x1.0p-53; }
    float nextFloat() { return next(24) / (float)(1 << 24); }
};

struct HAct { int type = 0, param = 0, x = 0, y = 0, utype = -1; }; // param: direction, or the duration of NONE
struct HChoice { int slot = 0, uid = 0, type = 0, x = 0, y = 0, owner = 0; std::vector<HAct> acts; };
// what PlayerActionGenerator's constructor sees of one game
struct HView {
    int W = 0, time = 0, winner = -1;
    bool gameover = false, can[2] = {false, false};
    int res[2] = {0, 0}, ru[2] = {0, 0};
    std::vector<int> pos;          // positions used by the assignments in flight
    std::vector<HChoice> choices;  // the player's idle units in unit-list order with their ordered action lists
};

// one game's slice of the arrays mrts_batch_unit_actions fills (include/microrts_cuda.h); false if a list did not fit
inline bool decode_view(const int32_t *hdr, const int32_t *pos, const int32_t *ch, const int32_t *ls, int K, int MA, int W, int none_duration, HView &v) {
    v.W = W; v.time = hdr[6]; v.gameover = hdr[7] & 1; v.winner = ((hdr[7] >> 1) & 3) - 1; v.can[0] = hdr[7] & 8; v.can[1] = hdr[7] & 16;
    v.res[0] = hdr[3]; v.res[1] = hdr[4]; v.ru[0] = hdr[1]; v.ru[1] = hdr[2];
    v.pos.assign(pos, pos + hdr[5]);
    v.choices.clear();
    if (hdr[0] > K) return false;
    for (int c = 0; c < hdr[0]; c++) {
        HChoice h;
        h.slot = ch[c * 4]; h.uid = ch[c * 4 + 1];
        int packed = ch[c * 4 + 2], cnt = ch[c * 4 + 3];
        h.type = packed & 255; h.x = (packed >> 8) & 255; h.y = (packed >> 16) & 255; h.owner = ((packed >> 24) & 255) - 1;
        if (cnt > MA) return false;
        for (int k = 0; k < cnt; k++) {
            uint32_t a = (uint32_t)ls[c * MA + k];
            HAct ua;
            ua.type = a & 15; ua.param = ua.type == 0 ? none_duration : (int)((a >> 4) & 15) - 1; ua.x = (a >> 8) & 255; ua.y = (a >> 16) & 255; ua.utype = (int)((a >> 24) & 255) - 1;
            h.acts.push_back(ua);
        }
        v.choices.push_back(std::move(h));
    }
    return true;
}

// UnitAction.resourceUsage: at most one position (linear arithmetic on x + y * W) and the cost of a PRODUCE
struct Usage { int pos = -1, res[2] = {0, 0}; bool has_pos = false; };
inline Usage act_usage(const HChoice &c, const HAct &a, const UttH &utt, int W) {
    Usage u;
    if (a.type == MRTS_MOVE || a.type == MRTS_PRODUCE) {
        if (a.type == MRTS_PRODUCE && c.owner >= 0 && a.utype >= 0 && a.utype < (int)utt.types.size()) u.res[c.owner] += utt.types[a.utype].cost;
        int p = c.x + c.y * W;
        switch (a.param) { case 0: p -= W; break; case 1: p++; break; case 2: p += W; break; case 3: p--; break; }
        u.pos = p; u.has_pos = true;
    }
    return u;
}
struct HRu {
    std::vector<int> pos; int res[2] = {0, 0};
    // this.consistentWith(another, gs): `another` = one action's usage
    bool consistent_with(const Usage &o, const int player_res[2]) const {
        if (o.has_pos) for (int p : pos) if (p == o.pos) return false;
        for (int i = 0; i < 2; i++) {
            if (o.res[i] == 0) continue;
            if (res[i] + o.res[i] > 0 && res[i] + o.res[i] > player_res[i]) return false;
        }
        return true;
    }
    void merge(const Usage &o) { if (o.has_pos) pos.push_back(o.pos); res[0] += o.res[0]; res[1] += o.res[1]; }
};
inline HRu base_usage(const HView &v) { HRu r; r.pos = v.pos; r.res[0] = v.ru[0]; r.res[1] = v.ru[1]; return r; }

typedef std::vector<std::pair<int, int>> HPlayerAction; // (choice index, action index) in the order the pairs were added

// GameState.getPlayerActions: the reference's order (the last unit's action varies fastest); stops at max_out (the full count is returned)
inline int64_t player_actions(const HView &v, const UttH &utt, std::vector<HPlayerAction> &out, int64_t max_out) {
    struct PA { HRu r; HPlayerAction a; };
    std::vector<PA> l(1);
    l[0].r = base_usage(v);
    for (int c = 0; c < (int)v.choices.size(); c++) {
        std::vector<PA> l2;
        for (const PA &pa : l)
            for (int k = 0; k < (int)v.choices[c].acts.size(); k++) {
                Usage u = act_usage(v.choices[c], v.choices[c].acts[k], utt, v.W);
                if (!pa.r.consistent_with(u, v.res)) continue;
                PA q; q.r = pa.r; q.r.merge(u); q.a = pa.a; q.a.emplace_back(c, k);
                l2.push_back(std::move(q));
            }
        l.swap(l2);
    }
    out.clear();
    for (int64_t i = 0; i < (int64_t)l.size() && i < max_out; i++) out.push_back(l[i].a);
    return (int64_t)l.size();
}

class PlayerActionGenerator {
  public:
    HView view; UttH utt; HRu base_ru;
    long long size = 1, generated = 0;
    std::vector<int> sizes, cur;
    bool more = true;

    // false: "Move generator created with no units that can execute actions"
    bool init(const HView &v, const UttH &u) {
        view = v; utt = u; base_ru = base_usage(v);
        size = 1; generated = 0; more = true;
        for (const HChoice &c : view.choices) {
            long long n = (long long)c.acts.size();
            if (INT64_MAX / size <= n) size = INT64_MAX; else size *= n;
        }
        if (view.choices.empty()) return false;
        sizes.clear(); for (const HChoice &c : view.choices) sizes.push_back((int)c.acts.size());
        cur.assign(sizes.size(), 0);
        return true;
    }
    void randomize_order(JavaRandom &r) {
        for (HChoice &c : view.choices) {
            std::vector<HAct> tmp = c.acts;
            c.acts.clear();
            while (!tmp.empty()) { int j = r.nextInt((int)tmp.size()); c.acts.push_back(tmp[j]); tmp.erase(tmp.begin() + j); }
        }
    }
    void increment(int start) {
        for (int i = 0; i < start; i++) cur[i] = 0;
        cur[start]++;
        if (cur[start] >= sizes[start]) { if (start < (int)cur.size() - 1) increment(start + 1); else more = false; }
    }
    // getNextAction(-1): false when exhausted
    bool next(HPlayerAction &pa) {
        while (more) {
            bool consistent = true;
            pa.clear();
            HRu r = base_ru;
            int i = (int)view.choices.size();
            while (i > 0) {
                i--;
                Usage u = act_usage(view.choices[i], view.choices[i].acts[cur[i]], utt, view.W);
                if (r.consistent_with(u, view.res)) { r.merge(u); pa.emplace_back(i, cur[i]); }
                else { consistent = false; break; }
            }
            increment(i);
            if (consistent) { generated++; return true; }
        }
        return false;
    }
    void random(JavaRandom &r, HPlayerAction &pa) {
        pa.clear();
        HRu ru = base_ru;
        for (int i = 0; i < (int)view.choices.size(); i++) {
            std::vector<int> l(view.choices[i].acts.size());
            for (int k = 0; k < (int)l.size(); k++) l[k] = k;
            bool consistent = false;
            do {
                int j = r.nextInt((int)l.size()), k = l[j];
                l.erase(l.begin() + j);
                Usage u = act_usage(view.choices[i], view.choices[i].acts[k], utt, view.W);
                if (ru.consistent_with(u, view.res)) { ru.merge(u); pa.emplace_back(i, k); consistent = true; }
            } while (!consistent);
        }
    }
    // getActionIndex for a PlayerAction given as (choice, action) pairs
    long long action_index(const HPlayerAction &pa) const {
        std::vector<int> choice(sizes.size(), 0);
        for (auto &p : pa) { if (p.first < 0 || p.first >= (int)sizes.size()) return -1; choice[p.first] = p.second; }
        long long index = 0, mult = 1;
        for (size_t i = 0; i < choice.size(); i++) { index += choice[i] * mult; mult *= sizes[i]; }
        return index;
    }
};

// a (choice, action) pair as a RAW action row of the C ABI: {cell, type, parameter, x, y, unit type, 0, 0}
inline void raw_row(const HView &v, const std::pair<int, int> &p, int32_t *row) {
    const HChoice &c = v.choices[p.first]; const HAct &a = c.acts[p.second];
    row[0] = c.x + c.y * v.W; row[1] = a.type; row[2] = a.param; row[3] = a.x; row[4] = a.y; row[5] = a.utype; row[6] = 0; row[7] = 0;
}

} // namespace mrts
