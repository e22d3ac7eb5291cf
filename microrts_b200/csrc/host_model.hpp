// host_model.hpp -- host-side world model of the C-ABI: unit type table, map (PhysicalGameState) loading, and the
// packing of both into the device layout of layout.h.  Plain C++17, no CUDA.
#pragma once
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "layout.h"

namespace mrts {

struct UnitTypeH {
    std::string name;
    int cost = 1, hp = 1, minDamage = 1, maxDamage = 1, attackRange = 1;
    int produceTime = 10, moveTime = 10, attackTime = 10, harvestTime = 10, returnTime = 10;
    int harvestAmount = 1, sightRadius = 4;
    bool isResource = false, isStockpile = false, canHarvest = false, canMove = true, canAttack = true;
    std::vector<int> produces;
    int flags() const {
        return (isResource ? UF_RESOURCE : 0) | (isStockpile ? UF_STOCKPILE : 0) | (canHarvest ? UF_HARVEST : 0) |
               (canMove ? UF_MOVE : 0) | (canAttack ? UF_ATTACK : 0);
    }
};

struct UttH {
    int conflict = 1;
    std::vector<UnitTypeH> types;
    int find(const std::string &n) const {
        for (size_t i = 0; i < types.size(); i++) if (types[i].name == n) return (int)i;
        return -1;
    }
    int maxAttackRange() const { // UnitTypeTable.getMaxAttackRange
        int m = 0;
        for (auto &t : types) if (t.attackRange > m) m = t.attackRange;
        return m;
    }
};

// UnitTypeTable.setUnitTypeTable(version, crs): src/rts/units/UnitTypeTable.java:104-289.
// Defaults for unset fields are UnitType's field initialisers (src/rts/units/UnitType.java:23-110).
inline UttH make_utt(int version, int conflict) {
    UttH u; u.conflict = conflict;
    auto add = [&](const char *name) -> UnitTypeH & { u.types.emplace_back(); u.types.back().name = name; return u.types.back(); };
    { auto &t = add("Resource"); t.isResource = true; t.canMove = false; t.canAttack = false; t.sightRadius = 0; }
    { auto &t = add("Base"); t.cost = 10; t.hp = 10;
      if (version == 1) t.produceTime = 250; else if (version == 2) t.produceTime = 200; // v3: default 10
      t.isStockpile = true; t.canMove = false; t.canAttack = false; t.sightRadius = 5; }
    { auto &t = add("Barracks"); t.cost = 5; t.hp = 4;
      if (version == 1) t.produceTime = 200; else if (version == 2 || version == 3) t.produceTime = 100;
      t.canMove = false; t.canAttack = false; t.sightRadius = 3; }
    { auto &t = add("Worker"); t.cost = 1; t.hp = 1;
      if (version == 3) { t.minDamage = 0; t.maxDamage = 2; } else { t.minDamage = t.maxDamage = 1; }
      t.attackRange = 1; t.produceTime = 50; t.moveTime = 10; t.attackTime = 5; t.harvestTime = 20; t.returnTime = 10;
      t.canHarvest = true; t.sightRadius = 3; }
    { auto &t = add("Light"); t.cost = 2; t.hp = 4;
      if (version == 3) { t.minDamage = 1; t.maxDamage = 3; } else { t.minDamage = t.maxDamage = 2; }
      t.attackRange = 1; t.produceTime = 80; t.moveTime = 8; t.attackTime = 5; t.sightRadius = 2; }
    { auto &t = add("Heavy");
      if (version == 3) { t.minDamage = 0; t.maxDamage = 6; } else { t.minDamage = t.maxDamage = 4; }
      t.attackRange = 1; t.produceTime = 120;
      if (version == 1) { t.moveTime = 12; t.hp = 4; t.cost = 2; }
      else if (version == 2 || version == 3) { t.moveTime = 10; t.hp = 8; t.cost = 3; }
      t.attackTime = 5; t.sightRadius = 2; }
    { auto &t = add("Ranged"); t.cost = 2; t.hp = 1;
      if (version == 3) { t.minDamage = 1; t.maxDamage = 2; } else { t.minDamage = t.maxDamage = 1; }
      t.attackRange = 3; t.produceTime = 100; t.moveTime = 10; t.attackTime = 5; t.sightRadius = 3; }
    u.types[1].produces = {3};
    u.types[2].produces = {4, 5, 6};
    u.types[3].produces = {1, 2};
    return u;
}

// ---- a tiny JSON reader, enough for UnitTypeTable.toJSON output (src/rts/units/UnitTypeTable.java:340-378) ----------
struct Json {
    enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
    bool b = false; double num = 0; std::string str;
    std::vector<Json> arr; std::vector<std::pair<std::string, Json>> obj;
    const Json *get(const std::string &k) const { for (auto &p : obj) if (p.first == k) return &p.second; return nullptr; }
    int geti(const std::string &k, int def) const { auto *j = get(k); return (j && j->kind == NUM) ? (int)j->num : def; }
    bool getb(const std::string &k, bool def) const { auto *j = get(k); return (j && j->kind == BOOL) ? j->b : def; }
};
struct JsonParser {
    const char *p; bool ok = true;
    explicit JsonParser(const char *s) : p(s) {}
    void ws() { while (*p && isspace((unsigned char)*p)) p++; }
    Json value() {
        Json j; ws();
        if (*p == '{') { p++; j.kind = Json::OBJ; ws(); if (*p == '}') { p++; return j; }
            for (;;) { ws(); Json k = value(); if (k.kind != Json::STR) { ok = false; return j; } ws(); if (*p != ':') { ok = false; return j; } p++;
                j.obj.emplace_back(k.str, value()); ws(); if (*p == ',') { p++; continue; } if (*p == '}') { p++; return j; } ok = false; return j; } }
        if (*p == '[') { p++; j.kind = Json::ARR; ws(); if (*p == ']') { p++; return j; }
            for (;;) { j.arr.push_back(value()); ws(); if (*p == ',') { p++; continue; } if (*p == ']') { p++; return j; } ok = false; return j; } }
        if (*p == '"') { p++; j.kind = Json::STR; while (*p && *p != '"') { if (*p == '\\' && p[1]) p++; j.str.push_back(*p++); } if (*p == '"') p++; else ok = false; return j; }
        if (!strncmp(p, "true", 4)) { p += 4; j.kind = Json::BOOL; j.b = true; return j; }
        if (!strncmp(p, "false", 5)) { p += 5; j.kind = Json::BOOL; j.b = false; return j; }
        if (!strncmp(p, "null", 4)) { p += 4; return j; }
        char *e; j.num = strtod(p, &e); if (e == p) { ok = false; return j; } p = e; j.kind = Json::NUM; return j;
    }
};

// UnitTypeTable.fromJSON + UnitType.updateFromJSON (src/rts/units/UnitTypeTable.java:413-431, UnitType.java:217-248)
inline bool utt_from_json(const char *text, UttH &u, std::string &err) {
    JsonParser jp(text); Json root = jp.value();
    if (!jp.ok || root.kind != Json::OBJ) { err = "malformed UnitTypeTable JSON"; return false; }
    u = UttH(); u.conflict = root.geti("moveConflictResolutionStrategy", 1);
    const Json *a = root.get("unitTypes");
    if (!a || a->kind != Json::ARR) { err = "missing unitTypes"; return false; }
    for (auto &o : a->arr) { UnitTypeH t; auto *n = o.get("name"); if (!n) { err = "unit type without name"; return false; } t.name = n->str; u.types.push_back(t); }
    for (size_t i = 0; i < a->arr.size(); i++) {
        auto &o = a->arr[i]; auto &t = u.types[i];
        t.cost = o.geti("cost", 1); t.hp = o.geti("hp", 1); t.minDamage = o.geti("minDamage", 1); t.maxDamage = o.geti("maxDamage", 1);
        t.attackRange = o.geti("attackRange", 1); t.produceTime = o.geti("produceTime", 10); t.moveTime = o.geti("moveTime", 10);
        // sic: the reference reads harvestTime from the "produceTime" key and never reads returnTime (UnitType.java:224-228)
        t.attackTime = o.geti("attackTime", 10); t.harvestTime = o.geti("produceTime", 10); t.returnTime = 10;
        t.harvestAmount = o.geti("harvestAmount", 10); t.sightRadius = o.geti("sightRadius", 10);
        t.isResource = o.getb("isResource", false); t.isStockpile = o.getb("isStockpile", false); t.canHarvest = o.getb("canHarvest", false);
        t.canMove = o.getb("canMove", false); t.canAttack = o.getb("canAttack", false);
        if (auto *pr = o.get("produces")) for (auto &s : pr->arr) { int id = u.find(s.str); if (id < 0) { err = "unknown produced type " + s.str; return false; } t.produces.push_back(id); }
    }
    return true;
}

inline bool utt_check_limits(const UttH &u, std::string &err) {
    if (u.types.empty() || u.types.size() > MRTS_MAX_TYPES) { err = "unit type table must have 1.." + std::to_string(MRTS_MAX_TYPES) + " types"; return false; }
    for (auto &t : u.types) {
        if (t.cost < 0 || t.cost > 255 || t.hp < 0 || t.hp > 255 || t.minDamage < 0 || t.minDamage > 255 || t.maxDamage < t.minDamage || t.maxDamage > 255 ||
            t.attackRange < 0 || t.attackRange > 127 || t.sightRadius < 0 || t.sightRadius > 255 || t.harvestAmount < 0 || t.harvestAmount > 255 ||
            t.produceTime < 0 || t.produceTime > 65535 || t.moveTime < 0 || t.moveTime > 65535 || t.attackTime < 0 || t.attackTime > 65535 ||
            t.harvestTime < 0 || t.harvestTime > 65535 || t.returnTime < 0 || t.returnTime > 65535 || t.produces.size() > 8) {
            err = "unit type '" + t.name + "' exceeds an engine field limit"; return false;
        }
    }
    return true;
}

// device constants: packed unit types + LCG jump table (layout.h)
inline void build_const_words(const UttH &u, std::vector<uint32_t> &w) {
    w.assign(MRTS_CONST_WORDS, 0);
    for (size_t i = 0; i < u.types.size(); i++) {
        const UnitTypeH &t = u.types[i];
        uint32_t *q = &w[i * MRTS_UTT_WORDS];
        q[0] = (uint32_t)t.cost | ((uint32_t)t.hp << 8) | ((uint32_t)t.minDamage << 16) | ((uint32_t)t.maxDamage << 24);
        q[1] = (uint32_t)t.attackRange | ((uint32_t)t.sightRadius << 8) | ((uint32_t)t.harvestAmount << 16) | ((uint32_t)t.flags() << 24);
        q[2] = (uint32_t)t.produceTime | ((uint32_t)t.moveTime << 16);
        q[3] = (uint32_t)t.attackTime | ((uint32_t)t.harvestTime << 16);
        q[4] = (uint32_t)t.returnTime | ((uint32_t)t.produces.size() << 16);
        for (size_t k = 0; k < t.produces.size(); k++) q[5 + (k >> 2)] |= (uint32_t)t.produces[k] << ((k & 3) * 8);
        for (size_t k = 0; k < t.produces.size() && k < 4; k++) q[7] |= (uint32_t)(u.types[t.produces[k]].cost & 0xff) << (k * 8);
    }
    // jump[d] = (A_d, C_d): s_{+2d} = A_d * s + C_d  (mod 2^48), java.util.Random's LCG
    const uint64_t A = 0x5DEECE66DULL, C = 0xBULL, M = (1ULL << 48) - 1;
    uint64_t a = 1, c = 0;
    uint64_t *j = (uint64_t *)&w[MRTS_MAX_TYPES * MRTS_UTT_WORDS];
    for (int d = 0; d < MRTS_JUMP_ENTRIES; d++) {
        j[2 * d] = a; j[2 * d + 1] = c;
        for (int s = 0; s < 2; s++) { a = (a * A) & M; c = (c * A + C) & M; }
    }
    uint16_t *eta = (uint16_t *)&w[MRTS_ETA_OFFSET];
    for (size_t i = 0; i < u.types.size(); i++) {
        const UnitTypeH &t = u.types[i];
        eta[i * 8 + 1] = (uint16_t)t.moveTime; eta[i * 8 + 2] = (uint16_t)t.harvestTime; eta[i * 8 + 3] = (uint16_t)t.moveTime;
        eta[i * 8 + 4] = (uint16_t)t.produceTime; eta[i * 8 + 5] = (uint16_t)t.attackTime;
    }
}

struct MapUnit { int type; long long id; int player, x, y, res, hp; };
struct MapH {
    int w = 0, h = 0;
    std::vector<uint8_t> terrain;
    int res[2] = {0, 0};
    std::vector<MapUnit> units;
};

// ---- minimal XML reader for the map format (PhysicalGameState.fromXML, src/rts/PhysicalGameState.java:700-726) ---------
struct XmlTag { std::string name; std::map<std::string, std::string> attr; bool closing = false; size_t end = 0; };
inline bool xml_next_tag(const std::string &s, size_t from, XmlTag &t, size_t &start) {
    for (;;) {
        size_t lt = s.find('<', from);
        if (lt == std::string::npos) return false;
        if (s.compare(lt, 4, "<!--") == 0) { size_t e = s.find("-->", lt); if (e == std::string::npos) return false; from = e + 3; continue; }
        if (lt + 1 < s.size() && (s[lt + 1] == '?' || s[lt + 1] == '!')) { size_t e = s.find('>', lt); if (e == std::string::npos) return false; from = e + 1; continue; }
        size_t gt = s.find('>', lt);
        if (gt == std::string::npos) return false;
        start = lt; t = XmlTag(); t.end = gt + 1;
        size_t p = lt + 1;
        if (s[p] == '/') { t.closing = true; p++; }
        size_t q = p;
        while (q < gt && !isspace((unsigned char)s[q]) && s[q] != '/') q++;
        t.name = s.substr(p, q - p);
        p = q;
        while (p < gt) {
            while (p < gt && (isspace((unsigned char)s[p]) || s[p] == '/')) p++;
            if (p >= gt) break;
            size_t k = p;
            while (k < gt && !isspace((unsigned char)s[k]) && s[k] != '=') k++;
            std::string key = s.substr(p, k - p);
            p = k;
            while (p < gt && isspace((unsigned char)s[p])) p++;
            if (p >= gt || s[p] != '=') continue;
            p++;
            while (p < gt && isspace((unsigned char)s[p])) p++;
            if (p >= gt || (s[p] != '"' && s[p] != '\'')) continue;
            char qc = s[p++];
            size_t v = s.find(qc, p);
            if (v == std::string::npos || v > gt) break;
            t.attr[key] = s.substr(p, v - p);
            p = v + 1;
        }
        return true;
    }
}

inline bool map_from_xml(const std::string &xml, const UttH &utt, MapH &m, std::string &err) {
    m = MapH();
    size_t pos = 0, st = 0; XmlTag t;
    bool have_root = false;
    std::vector<long long> ids;
    while (xml_next_tag(xml, pos, t, st)) {
        pos = t.end;
        if (t.closing) { if (t.name == "rts.PhysicalGameState") break; continue; }
        if (t.name == "rts.PhysicalGameState") {
            if (have_root) break;
            have_root = true;
            m.w = atoi(t.attr["width"].c_str()); m.h = atoi(t.attr["height"].c_str());
            if (m.w <= 0 || m.h <= 0) { err = "bad map size"; return false; }
        } else if (t.name == "terrain") {
            size_t e = xml.find("</terrain>", pos);
            if (e == std::string::npos) { err = "unterminated <terrain>"; return false; }
            for (size_t i = pos; i < e; i++) { char ch = xml[i]; if (ch == '0' || ch == '1') m.terrain.push_back((uint8_t)(ch - '0')); else if (!isspace((unsigned char)ch)) { err = "unsupported terrain encoding"; return false; } }
            pos = e;
        } else if (t.name == "rts.Player") {
            int id = atoi(t.attr["ID"].c_str());
            if (id < 0 || id > 1) { err = "only two players are supported"; return false; }
            m.res[id] = atoi(t.attr["resources"].c_str());
        } else if (t.name == "rts.units.Unit") {
            MapUnit u;
            u.type = utt.find(t.attr["type"]);
            if (u.type < 0) { err = "unknown unit type '" + t.attr["type"] + "'"; return false; }
            u.id = atoll(t.attr["ID"].c_str()); u.player = atoi(t.attr["player"].c_str());
            u.x = atoi(t.attr["x"].c_str()); u.y = atoi(t.attr["y"].c_str());
            u.res = atoi(t.attr["resources"].c_str()); u.hp = atoi(t.attr["hitpoints"].c_str());
            for (long long o : ids) if (o == u.id) { err = "Repeated unit ID " + std::to_string(u.id) + " in map"; return false; } // PhysicalGameState.java:719-722
            ids.push_back(u.id);
            m.units.push_back(u);
        }
    }
    if (!have_root) { err = "no <rts.PhysicalGameState> element"; return false; }
    if ((int)m.terrain.size() != m.w * m.h) { err = "terrain length does not match width*height"; return false; }
    return true;
}

inline bool map_check(const MapH &m, const UttH &utt, std::string &err) {
    if (m.w < 1 || m.h < 1 || m.w > 254 || m.h > 254) { err = "map dimensions out of range (1..254)"; return false; }
    std::vector<uint8_t> occ((size_t)m.w * m.h, 0);
    for (auto &u : m.units) {
        if (u.type < 0 || u.type >= (int)utt.types.size()) { err = "unit with unknown type"; return false; }
        if (u.x < 0 || u.y < 0 || u.x >= m.w || u.y >= m.h) { err = "unit outside the map"; return false; }
        if (u.player < -1 || u.player > 1) { err = "unit owner must be -1, 0 or 1"; return false; }
        if (u.x + u.y * m.w < (int)m.terrain.size() && m.terrain[u.x + u.y * m.w]) { err = "unit placed on a wall cell: (" + std::to_string(u.x) + ", " + std::to_string(u.y) + ")"; return false; } // the occupancy grid holds one byte per cell: wall OR unit
        if (occ[u.x + u.y * m.w]) { err = "PhysicalGameState.addUnit: added two units in position: (" + std::to_string(u.x) + ", " + std::to_string(u.y) + ")"; return false; }
        if (u.res < -32768 || u.res > 32767 || u.hp < -32768 || u.hp > 32767) { err = "unit resources/hitpoints out of range"; return false; }
        occ[u.x + u.y * m.w] = 1;
    }
    return true;
}

// upper bound on simultaneously live units: every resource point can become at most one unit of cost >= 1
inline int map_unit_bound(const MapH &m) {
    long long b = (long long)m.units.size() + m.res[0] + m.res[1];
    for (auto &u : m.units) if (u.player < 0) b += u.res > 0 ? u.res : 0;
    long long cells = (long long)m.w * m.h;
    return (int)(b < cells ? b : cells);
}

inline void build_map_blob(const MapH &m, int cap, std::vector<uint32_t> &blob) {
    int W = m.w, H = m.h, P = W + 2;
    blob.assign(mrts_map_blob_words(W, H, cap), 0);
    int pcb = (((W + 2) * (H + 2)) + 15) & ~15;
    uint8_t *grid = (uint8_t *)blob.data();
    memset(grid, 0xFF, pcb);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) grid[(y + 1) * P + x + 1] = m.terrain[x + y * W] ? 0xFF : 0;
    int32_t *hdr = (int32_t *)(blob.data() + pcb / 4);
    uint32_t *un = blob.data() + pcb / 4 + MRTS_HDR_WORDS;
    long long next_id = 0;
    int n = 0;
    for (auto &u : m.units) {
        if (n >= cap) break;
        un[UW_W0 * cap + n] = (uint32_t)u.type | ((uint32_t)(u.player + 1) << 8) | ((uint32_t)u.x << 16) | ((uint32_t)u.y << 24);
        un[UW_W1 * cap + n] = ((uint32_t)u.hp & 0xffffu) | ((uint32_t)u.res << 16);
        un[UW_A0 * cap + n] = AT_IDLE | (0xFFu << 8);
        un[UW_ID * cap + n] = (uint32_t)u.id;
        if (u.id >= next_id) next_id = u.id + 1;
        n++;
    }
    hdr[H_RES0] = m.res[0]; hdr[H_RES1] = m.res[1]; hdr[H_NUNITS] = n; hdr[H_NEXTID] = (int32_t)next_id;
    uint8_t *ter = (uint8_t *)(blob.data() + mrts_map_terrain_offset_words(W, H, cap));
    for (int i = 0; i < W * H; i++) ter[i] = m.terrain[i] ? 1 : 0;
}

inline uint64_t jr_scramble(int64_t seed) { return ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
static const int64_t SEED_XOR_CONFLICT = 0x5851F42D4C957F2DLL; // GameState.r stream
static const int64_t SEED_XOR_DAMAGE = 0x14057B7EF767814FLL;   // UnitAction.r stream

} // namespace mrts
