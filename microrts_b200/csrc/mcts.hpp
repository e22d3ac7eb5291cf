// mcts.hpp -- host-side NaiveMCTS over the batched engine: many searches in lockstep, one per root game.
//
//   NaiveMCTS.startNewComputation / iteration / getBestActionSoFar   src/ai/mcts/naivemcts/NaiveMCTS.java:140-158,195-223,226-262
//   NaiveMCTSNode (constructor, selectLeaf, selectFromAlreadySampled*, selectLeafUsingLocalMABs, propagateEvaluation)
//                                                                    src/ai/mcts/naivemcts/NaiveMCTSNode.java:39-105,108-188,191-330,341-368
//   Sampler.weighted                                                 src/util/Sampler.java:116-137,141-161
//   UCT.startNewComputation / monteCarloRun / getBestActionSoFar     src/ai/mcts/uct/UCT.java:103-110,140-168,171-199
//   UCTNode (constructor, UCTSelectLeaf, childValue)                 src/ai/mcts/uct/UCTNode.java:37-68,70-109,112-125
// The trees (visit counts, the local multi-armed bandits, the children maps) live on the host; every game-rule operation runs on
// the device for all searches at once: a node's state is a game of a pool batch; creating a node is clone (mrts_batch_copy_games) +
// issue of the sampled PlayerAction + the node's cycle loop (mrts_batch_cycle_to_decision) + its move generator's lists
// (mrts_batch_unit_actions); a playout is mrts_batch_rollout of the leaf's state.
// The reference draws from unseeded static generators (MCTSNode.r, and util.Sampler.generator which the playout policy shares), so
// it cannot be replayed; here search t owns a seeded MCTSNode.r stream and a seeded Sampler stream, and its k-th playout is seeded
// with seed_t * 1000003 + k.  With equal seeds the trees equal the CPU oracle's restatement node for node (tests/test_mcts.py).
#pragma once
#include <cmath>
#include <map>

#include "player_actions.hpp"

namespace mrts {

struct MctsParams { int lookahead = 100, max_depth = 10; float e_l = 0.3f, e_g = 0.0f, e_0 = 0.4f; int strategy = 0, fensa = 1, eval_fn = 0; };

struct MNode {
    int type = -1, parent = -1, depth = 0, time = 0; // type: 0 max, 1 min, -1 game over
    double accum = 0; int visits = 0;
    bool has_gen = false;
    HView view;                                     // moveGenerator.getChoices() + the state's resource usage
    std::vector<int> children;
    std::vector<std::vector<int>> codes;            // per child: one action index per choice (the BigInteger action code)
    std::vector<HPlayerAction> pas;                 // per child: the PlayerAction in the order it was sampled
    std::map<std::vector<int>, int> children_map;
    std::vector<std::vector<double>> ate_accum; std::vector<std::vector<int>> ate_visits; // unitActionTable
    // UCT: the node's move generator (shuffled at construction), UCTNode.hasMoreActions and the float accumulator of UCTNode
    PlayerActionGenerator gen; bool has_more = true; float accum_f = 0;
};

struct MTree {
    std::vector<MNode> nodes;
    JavaRandom r, sampler;
    int64_t seed = 0, runs = 0;
    int root_time = 0;
    // the iteration in flight
    int leaf = -1; bool creating = false; int new_parent = -1; HPlayerAction new_pa; std::vector<int> new_code;
};

class NaiveMctsHost {
  public:
    MctsParams P; int player = 0; UttH utt; double bound = 1.0;
    int algorithm = 0; // 0 = NaiveMCTS, 1 = UCT
    std::vector<MTree> trees;

    int sampler_weighted(MTree &t, const std::vector<double> &dist) {
        double total = 0, accum = 0, tmp;
        for (double f : dist) total += f;
        if (total == 0) return t.sampler.nextInt((int)dist.size());
        tmp = t.sampler.nextDouble() * total;
        for (int i = 0; i < (int)dist.size(); i++) { accum += dist[i]; if (accum >= tmp) return i; }
        return (int)dist.size() - 1;
    }
    int select_egreedy(MTree &t, MNode &nd) {
        if (t.r.nextFloat() >= P.e_g) {
            int best = -1;
            for (int ci : nd.children) {
                MNode &c = t.nodes[ci];
                if (best < 0) { best = ci; continue; }
                MNode &b = t.nodes[best];
                if (nd.type == 0 ? (c.accum / c.visits) > (b.accum / b.visits) : (c.accum / c.visits) < (b.accum / b.visits)) best = ci;
            }
            return best;
        }
        return nd.children[t.r.nextInt((int)nd.children.size())];
    }
    int select_ucb1(MTree &t, MNode &nd) {
        int best = -1; double best_score = 0; const float C = 0.05f;
        for (int ci : nd.children) {
            MNode &c = t.nodes[ci];
            double exploitation = ((double)c.accum) / c.visits, exploration = std::sqrt(std::log((double)nd.visits) / c.visits);
            exploitation = nd.type == 0 ? (bound + exploitation) / (2 * bound) : (bound - exploitation) / (2 * bound);
            double tmp = C * exploitation + exploration;
            if (best < 0 || tmp > best_score) { best = ci; best_score = tmp; }
        }
        return best;
    }
    // selectLeaf: sets t.leaf, or t.creating with the parent / PlayerAction / code of the node to create
    void select_leaf(MTree &t, int ni) {
        for (;;) {
            MNode &nd = t.nodes[ni];
            if (!nd.has_gen || nd.depth >= P.max_depth) { t.leaf = ni; return; }
            if (!nd.children.empty() && t.r.nextFloat() >= P.e_0) { ni = P.strategy == 0 ? select_egreedy(t, nd) : select_ucb1(t, nd); continue; }
            // selectLeafUsingLocalMABs
            const int nc = (int)nd.view.choices.size();
            std::vector<std::vector<double>> dists(nc);
            std::vector<int> not_sampled;
            for (int e = 0; e < nc; e++) {
                const int na = (int)nd.view.choices[e].acts.size();
                std::vector<double> &dist = dists[e]; dist.resize(na);
                const std::vector<double> &acc = nd.ate_accum[e]; const std::vector<int> &vc = nd.ate_visits[e];
                int best_idx = -1, visits = 0; double best_eval = 0;
                for (int i = 0; i < na; i++) {
                    bool take = best_idx == -1 || (visits != 0 && vc[i] == 0) ||
                                (visits != 0 && (nd.type == 0 ? (acc[i] / vc[i]) > best_eval : (acc[i] / vc[i]) < best_eval));
                    if (take) { best_idx = i; best_eval = vc[i] > 0 ? acc[i] / vc[i] : 0; visits = vc[i]; }
                    dist[i] = P.e_l / na; // float arithmetic, as in the reference
                }
                if (vc[best_idx] != 0) dist[best_idx] = (1 - P.e_l) + (P.e_l / na);
                else if (P.fensa) { for (int j = 0; j < na; j++) if (vc[j] > 0) dist[j] = 0; }
                not_sampled.push_back(e);
            }
            HRu ru = base_usage(nd.view);
            HPlayerAction pa; std::vector<int> code_v(nc, 0);
            while (!not_sampled.empty()) {
                int k = t.r.nextInt((int)not_sampled.size());
                int i = not_sampled[k];
                not_sampled.erase(not_sampled.begin() + k);
                const HChoice &c = nd.view.choices[i];
                int code = sampler_weighted(t, dists[i]);
                Usage u = act_usage(c, c.acts[code], utt, nd.view.W);
                if (!ru.consistent_with(u, nd.view.res)) {
                    std::vector<double> dl = dists[i]; std::vector<int> outs(dl.size());
                    for (int j = 0; j < (int)outs.size(); j++) outs[j] = j;
                    do {
                        int idx = 0; while (outs[idx] != code) idx++;
                        dl.erase(dl.begin() + idx); outs.erase(outs.begin() + idx);
                        double total = 0, accum = 0, tmp;
                        for (double f : dl) total += f;
                        if (total == 0) code = outs[t.sampler.nextInt((int)outs.size())];
                        else {
                            tmp = t.sampler.nextDouble() * total; code = outs.back();
                            for (int j = 0; j < (int)dl.size(); j++) { accum += dl[j]; if (accum >= tmp) { code = outs[j]; break; } }
                        }
                        u = act_usage(c, c.acts[code], utt, nd.view.W);
                    } while (!ru.consistent_with(u, nd.view.res));
                }
                ru.merge(u);
                pa.emplace_back(i, code);
                code_v[i] = code;
            }
            auto it = nd.children_map.find(code_v);
            if (it != nd.children_map.end()) { ni = it->second; continue; }
            t.creating = true; t.new_parent = ni; t.new_pa = pa; t.new_code = code_v; t.leaf = -1;
            return;
        }
    }
    // UCTSelectLeaf
    void select_leaf_uct(MTree &t, int ni) {
        for (;;) {
            MNode &nd = t.nodes[ni];
            if (nd.depth >= P.max_depth) { t.leaf = ni; return; }
            if (nd.has_more) {
                if (!nd.has_gen) { t.leaf = ni; return; }
                HPlayerAction pa;
                if (nd.gen.next(pa)) { t.creating = true; t.new_parent = ni; t.new_pa = pa; t.new_code.clear(); t.leaf = -1; return; }
                nd.has_more = false;
            }
            int best = -1; double best_score = 0; const float C = 0.05f; const float fbound = 1.0f;
            for (int ci : nd.children) {
                MNode &c = t.nodes[ci];
                double exploitation = ((double)c.accum_f) / c.visits, exploration = std::sqrt(std::log((double)nd.visits) / c.visits);
                exploitation = nd.type == 0 ? (fbound + exploitation) / (2 * fbound) : (fbound - exploitation) / (2 * fbound);
                double tmp = C * exploitation + exploration;
                if (best < 0 || tmp > best_score) { best = ci; best_score = tmp; }
            }
            if (best < 0) { t.leaf = ni; return; }
            ni = best;
        }
    }
    // what a node's constructor derives from its (already cycled) state: `type` and `view` come from the device
    void init_node(MTree &t, MNode &nn, int type, int time, const HView *view) {
        nn.type = type; nn.time = time;
        if (!view) return;
        nn.has_gen = true; nn.view = *view;
        if (algorithm == 0) {
            for (const HChoice &c : view->choices) { nn.ate_accum.emplace_back(c.acts.size(), 0.0); nn.ate_visits.emplace_back(c.acts.size(), 0); }
        } else {
            nn.gen.init(*view, utt);
            nn.gen.randomize_order(t.r); // moveGenerator.randomizeOrder()
            nn.view = nn.gen.view;       // the rows of a sampled PlayerAction index the shuffled lists
        }
    }
    // the node created for the iteration in flight
    int attach_new_node(MTree &t, int type, int time, const HView *view) {
        MNode nn;
        nn.parent = t.new_parent; nn.depth = t.nodes[t.new_parent].depth + 1;
        init_node(t, nn, type, time, view);
        int id = (int)t.nodes.size();
        t.nodes.push_back(std::move(nn));
        MNode &par = t.nodes[t.new_parent];
        par.children.push_back(id); par.codes.push_back(t.new_code); par.pas.push_back(t.new_pa);
        if (algorithm == 0) par.children_map[t.new_code] = id;
        t.leaf = id; t.creating = false;
        return id;
    }
    void propagate(MTree &t, int ni, double evaluation) {
        if (algorithm == 1) { // UCT.monteCarloRun: float accumulators up the parents
            while (ni >= 0) { MNode &nd = t.nodes[ni]; nd.accum_f += evaluation; nd.accum = nd.accum_f; nd.visits++; ni = nd.parent; }
            return;
        }
        int child = -1;
        while (ni >= 0) {
            MNode &nd = t.nodes[ni];
            nd.accum += evaluation; nd.visits++;
            if (child >= 0) {
                int idx = 0; while (nd.children[idx] != child) idx++;
                const std::vector<int> &code = nd.codes[idx];
                for (size_t i = 0; i < code.size(); i++) { nd.ate_accum[i][code[i]] += evaluation; nd.ate_visits[i][code[i]]++; }
            }
            child = ni; ni = nd.parent;
        }
    }
    int most_visited(const MTree &t) const {
        const MNode &root = t.nodes[0]; int best = -1;
        for (int i = 0; i < (int)root.children.size(); i++) {
            const MNode &c = t.nodes[root.children[i]];
            if (best == -1) { best = i; continue; }
            const MNode &b = t.nodes[root.children[best]];
            if (c.visits > b.visits || (algorithm == 1 && c.visits == b.visits && c.accum_f > b.accum_f)) best = i; // UCT breaks ties by evaluation
        }
        return best;
    }
};

} // namespace mrts
