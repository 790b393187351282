// Native forward simulator emitting flattened forests directly.
//
// Restates the LAW of the reference's `rand_tree(model, present_time, init_type, n; reject_stubs, min_leaves,
// max_rejections)` (src/multitype_branching_process.jl:318-402) with `sample_child!` (:432-461) and `mutate!`
// (:411-420), including the pruning of unsampled subtrees and the child re-ordering caused by `delete!`
// (src/treenode.jl:38-48: a spliced-out birth re-attaches its child at the END of the grandparent's children).
// Julia's RNG stream cannot be reproduced; every tree draws from its own splitmix64/xoshiro256** stream
// keyed by (seed, tree index), so a forest does not depend on thread count or on how many trees are drawn.
// Output = the post-order struct-of-arrays of include/gcdyn_b200.h (nodes of PostOrderTraversal(tree.children[1])).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace gcdyn_simu {

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    Rng(uint64_t seed, uint64_t stream) {
        uint64_t x = seed * 0x9e3779b97f4a7c15ULL + stream * 0xd1342543de82ef95ULL + 0x632be59bd9b4e019ULL;
        for (auto& v : s) v = splitmix(x);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {   // xoshiro256**
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }          // [0,1)
    double exponential(double mean) { return -mean * std::log1p(-uniform()); }
};

struct Model {                 // canonical parameter set (one model), Γ row-major here
    int K;
    const double* lambda;      // [K]
    double mu, delta, rho, sigma;
    const double* Gamma;       // [K*K] row-major
};

struct Node {
    uint8_t event;             // index into EVENTS (types.jl:6)
    int32_t type;              // type index
    double time;
    int32_t up;
    std::vector<int32_t> children;
    bool keep = false;
};

enum { EV_ROOT = 0, EV_BIRTH, EV_SDEATH, EV_UDEATH, EV_TC, EV_SSURV, EV_USURV };

// StatsBase `sample(Weights(w))`: one uniform scaled by the total, cumulative walk (mbp.jl:419,442,445).
inline int weighted_index(Rng& rng, const double* w, int n, double total) {
    const double t = rng.uniform() * total;
    int i = 0;
    double cw = w[0];
    while (cw < t && i < n - 1) { ++i; cw += w[i]; }
    return i;
}

struct FlatOut {
    std::vector<int64_t> tree_off{0}, child_off{0};
    std::vector<uint8_t> event;
    std::vector<int32_t> type_idx, btype_idx, parent, child_idx, root_type_idx;
    std::vector<double> t_start, t_end, root_time;
    int64_t rejections = 0;
};

// One accepted tree appended to `out`.  Returns false when max_rejections is exceeded (mbp.jl:389-391).
inline bool simulate_tree(const Model& m, double T, int init_type, uint64_t seed, uint64_t tree_index, bool reject_stubs,
                          int min_leaves, int max_rejections, FlatOut& out, std::vector<Node>& nodes,
                          std::vector<double>& wbuf) {
    Rng rng(seed, tree_index);
    const int K = m.K;
    int rejections = 0;
    for (;;) {
        nodes.clear();
        nodes.push_back(Node{EV_ROOT, init_type, 0.0, -1, {}, false});
        std::vector<int32_t> stack{0};
        // evolve a fully observed tree, LIFO (mbp.jl:346-356)
        while (!stack.empty()) {
            const int32_t pi = stack.back();
            stack.pop_back();
            const int x = nodes[pi].type;
            const double lx = m.lambda[x], mx = m.mu, gx = m.delta * -m.Gamma[(size_t)x * K + x];
            const double wait = rng.exponential(1.0 / (lx + mx + gx));
            Node c{};
            c.up = pi;
            c.type = x;
            if (nodes[pi].time + wait > T) {
                const double w2[2] = {m.rho, 1.0 - m.rho};
                c.event = weighted_index(rng, w2, 2, 1.0) == 0 ? EV_SSURV : EV_USURV;
                c.time = T;
            } else {
                const double w3[3] = {lx, mx, gx};
                const int k = weighted_index(rng, w3, 3, lx + mx + gx);
                c.event = k == 0 ? EV_BIRTH : (k == 1 ? EV_UDEATH : EV_TC);
                if (c.event == EV_UDEATH && rng.uniform() < m.sigma) c.event = EV_SDEATH;
                c.time = nodes[pi].time + wait;
                if (c.event == EV_TC) {                     // mutate!, mbp.jl:411-420
                    double total = 0.0;
                    for (int j = 0; j < K; ++j) {
                        wbuf[j] = j == x ? 0.0 : m.Gamma[(size_t)x * K + j] / -m.Gamma[(size_t)x * K + x];
                        total += wbuf[j];
                    }
                    c.type = weighted_index(rng, wbuf.data(), K, total);
                }
            }
            const int32_t ci = (int32_t)nodes.size();
            nodes.push_back(c);
            nodes[pi].children.push_back(ci);
            if (c.event == EV_BIRTH) { stack.push_back(ci); stack.push_back(ci); }
            else if (c.event == EV_TC) stack.push_back(ci);
        }
        // flag every node with a sampled descendant (mbp.jl:359-372)
        for (int32_t i = 0; i < (int32_t)nodes.size(); ++i) {
            if (!nodes[i].children.empty() || (nodes[i].event != EV_SSURV && nodes[i].event != EV_SDEATH)) continue;
            for (int32_t j = i; j >= 0 && !nodes[j].keep; j = nodes[j].up) nodes[j].keep = true;
        }
        // prune in post-order; a birth left with one child is spliced out (mbp.jl:375-383, treenode.jl:38-48)
        {
            std::vector<int32_t> order;
            order.reserve(nodes.size());
            std::vector<std::pair<int32_t, int32_t>> st{{0, 0}};
            while (!st.empty()) {
                auto& top = st.back();
                if (top.second < (int32_t)nodes[top.first].children.size()) {
                    const int32_t ch = nodes[top.first].children[top.second++];
                    st.push_back({ch, 0});
                } else { order.push_back(top.first); st.pop_back(); }
            }
            for (int32_t i : order) {
                Node& n = nodes[i];
                if (n.event != EV_BIRTH && n.event != EV_ROOT) continue;
                std::vector<int32_t> kept;
                for (int32_t c : n.children) if (nodes[c].keep) kept.push_back(c);
                n.children.swap(kept);
                if (n.event == EV_BIRTH && n.children.size() == 1) {        // delete!(node)
                    const int32_t par = n.up, child = n.children[0];
                    auto& pc = nodes[par].children;
                    for (size_t k = 0; k < pc.size();) { if (pc[k] == i) pc.erase(pc.begin() + k); else ++k; }   // detach!
                    n.children.clear();
                    nodes[child].up = par;
                    pc.push_back(child);                                     // attach! appends at the END
                    n.up = -1;
                }
            }
        }
        // leaves of the pruned tree (LeafTraversal(root) counts the root itself when it has no children)
        int leaves = 0;
        {
            std::vector<int32_t> st{0};
            while (!st.empty()) {
                const int32_t i = st.back(); st.pop_back();
                if (nodes[i].children.empty()) ++leaves;
                else for (int32_t c : nodes[i].children) st.push_back(c);
            }
        }
        if ((reject_stubs && nodes[0].children.empty()) || leaves < min_leaves) {
            if (++rejections > max_rejections) return false;
            continue;
        }
        break;
    }
    out.rejections += rejections;
    // flatten PostOrderTraversal(root.children[1]) (treenode.jl:159-175)
    const int64_t base = (int64_t)out.event.size();
    std::vector<int32_t> order;
    if (!nodes[0].children.empty()) {
        std::vector<std::pair<int32_t, int32_t>> st{{nodes[0].children[0], 0}};
        while (!st.empty()) {
            auto& top = st.back();
            if (top.second < (int32_t)nodes[top.first].children.size()) {
                const int32_t ch = nodes[top.first].children[top.second++];
                st.push_back({ch, 0});
            } else { order.push_back(top.first); st.pop_back(); }
        }
    }
    std::vector<int32_t> pos(nodes.size(), -1);
    for (size_t k = 0; k < order.size(); ++k) pos[order[k]] = (int32_t)k;
    for (int32_t i : order) {
        const Node& n = nodes[i];
        out.event.push_back(n.event);
        out.type_idx.push_back(n.type);
        out.btype_idx.push_back(nodes[n.up].type);
        out.t_start.push_back(nodes[n.up].time);
        out.t_end.push_back(n.time);
        out.parent.push_back(n.up == 0 ? -1 : pos[n.up]);
        for (int32_t c : n.children) out.child_idx.push_back(pos[c]);
        out.child_off.push_back((int64_t)out.child_idx.size());
    }
    out.tree_off.push_back(base + (int64_t)order.size());
    out.root_type_idx.push_back(init_type);
    out.root_time.push_back(0.0);
    return true;
}

}  // namespace gcdyn_simu
