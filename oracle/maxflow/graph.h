// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// A from-scratch stand-in for the third-party max-flow library the reference includes but does not vendor:
// `#include "../maxflow/graph.h"` (FastGCStereo.h:7) = Boykov-Kolmogorov maxflow v3.01/3.04 (README.md:42-48, maxflow/README.TXT;
// .gitignore:1-5).  Only the interface the reference uses is provided, with the published semantics of that library:
//
//   Graph<captype, tcaptype, flowtype>(node_num_max, edge_num_max)
//   add_node(n)                        nodes 0..n-1
//   add_tweights(i, cap_source, cap_sink)   accumulates; only the difference tr_cap = sum(source) - sum(sink) shapes the cut, the
//                                      common part min(source, sink) is added to the flow value (BK's `flow += min(..)`)
//   add_edge(i, j, cap, rev_cap)       arc i->j with capacity cap and its sister j->i with rev_cap
//   maxflow()                          value of the maximum flow (= minimum cut) including the add_tweights constants
//   what_segment(i, default = SOURCE)  SINK for the nodes of the sink search tree, SOURCE for the source tree, `default` for free
//                                      nodes.  When BK terminates its trees are maximal: the sink tree is exactly the set of nodes from
//                                      which the sink can still be reached through arcs with residual capacity, so with the default
//                                      argument (the only form the reference uses, FastGCStereo.h:366,553) the answer is
//                                      SINK  <=>  the sink is reachable from i in the residual graph of a maximum flow.
//                                      That set (the smallest sink side of any minimum cut) does not depend on which maximum flow an
//                                      algorithm finds, so any exact max-flow algorithm reproduces BK's segmentation -- up to floating-
//                                      point ties in `captype` arithmetic (documented residual freedom).
//
// Algorithm here: Dinic (BFS level graph + iterative DFS blocking flow) on the explicit graph with a source and a sink node; residual
// capacities are kept in `captype` / `tcaptype` like BK does, so a saturated arc is exactly 0.
#pragma once
#include <algorithm>
#include <cstddef>
#include <limits>
#include <vector>

template <typename captype, typename tcaptype, typename flowtype> class Graph {
public:
    typedef enum { SOURCE = 0, SINK = 1 } termtype;
    typedef int node_id;

    Graph(int node_num_max, int edge_num_max, void (*err_function)(const char*) = NULL) : flow_(0), solved_(false) {
        tr_cap_.reserve((size_t)std::max(node_num_max, 0));
        arcs_.reserve(2 * (size_t)std::max(edge_num_max, 0));
        (void)err_function;
    }
    node_id add_node(int num = 1) {
        const node_id first = (node_id)tr_cap_.size();
        tr_cap_.resize(tr_cap_.size() + (size_t)num, (tcaptype)0);
        head_.resize(tr_cap_.size(), -1);
        return first;
    }
    void add_edge(node_id i, node_id j, captype cap, captype rev_cap) {
        arcs_.push_back(Arc{j, head_[i], cap}); head_[i] = (int)arcs_.size() - 1;
        arcs_.push_back(Arc{i, head_[j], rev_cap}); head_[j] = (int)arcs_.size() - 1;
    }
    void add_tweights(node_id i, tcaptype cap_source, tcaptype cap_sink) {
        const tcaptype delta = tr_cap_[i];
        if (delta > 0) cap_source += delta; else cap_sink -= delta;
        flow_ += (cap_source < cap_sink) ? cap_source : cap_sink;
        tr_cap_[i] = cap_source - cap_sink;
    }
    flowtype maxflow(bool reuse_trees = false, void* changed_list = NULL) {
        (void)reuse_trees; (void)changed_list;
        const int n = (int)tr_cap_.size();
        std::vector<int> level(n), it(n), queue(n), stack, path;
        for (;;) {
            // BFS from the source side: nodes with residual source capacity are level 1
            std::fill(level.begin(), level.end(), 0);
            int qh = 0, qt = 0;
            for (int v = 0; v < n; v++) if (tr_cap_[v] > 0) { level[v] = 1; queue[qt++] = v; }
            bool sink_seen = false;
            while (qh < qt) {
                const int v = queue[qh++];
                if (tr_cap_[v] < 0) sink_seen = true;
                for (int a = head_[v]; a >= 0; a = arcs_[a].next)
                    if (arcs_[a].r > 0 && !level[arcs_[a].to]) { level[arcs_[a].to] = level[v] + 1; queue[qt++] = arcs_[a].to; }
            }
            if (!sink_seen) break;
            for (int v = 0; v < n; v++) it[v] = head_[v];
            // blocking flow: DFS from every node that still has source capacity
            for (int s = 0; s < n; s++) {
                while (tr_cap_[s] > 0 && level[s] == 1) {
                    // find a path s -> ... -> t (tr_cap < 0) in the level graph
                    path.clear();
                    int v = s;
                    bool found = false;
                    for (;;) {
                        if (tr_cap_[v] < 0) { found = true; break; }
                        bool advanced = false;
                        for (int& a = it[v]; a >= 0; a = arcs_[a].next) {
                            const int u = arcs_[a].to;
                            if (arcs_[a].r > 0 && level[u] == level[v] + 1) { path.push_back(a); v = u; advanced = true; break; }
                        }
                        if (advanced) continue;
                        level[v] = -1;   // dead end: remove from the level graph
                        if (path.empty()) break;
                        const int a = path.back(); path.pop_back();
                        v = arcs_[a ^ 1].to;
                    }
                    if (!found) break;
                    tcaptype f = tr_cap_[s];
                    if (-tr_cap_[v] < f) f = -tr_cap_[v];
                    for (int a : path) if ((tcaptype)arcs_[a].r < f) f = (tcaptype)arcs_[a].r;
                    for (int a : path) { arcs_[a].r -= (captype)f; arcs_[a ^ 1].r += (captype)f; }
                    tr_cap_[s] -= f; tr_cap_[v] += f;
                    flow_ += f;
                }
            }
        }
        // sink side: nodes from which the sink is reachable through residual arcs (backward BFS from the nodes with sink capacity)
        reach_sink_.assign(n, 0);
        int qh = 0, qt = 0;
        for (int v = 0; v < n; v++) if (tr_cap_[v] < 0) { reach_sink_[v] = 1; queue[qt++] = v; }
        while (qh < qt) {
            const int v = queue[qh++];
            for (int a = head_[v]; a >= 0; a = arcs_[a].next) {   // arc v->u; its sister u->v carries the residual towards v
                const int u = arcs_[a].to;
                if (arcs_[a ^ 1].r > 0 && !reach_sink_[u]) { reach_sink_[u] = 1; queue[qt++] = u; }
            }
        }
        solved_ = true;
        return flow_;
    }
    termtype what_segment(node_id i, termtype default_segm = SOURCE) {
        if (!solved_) return default_segm;
        if (reach_sink_[i]) return SINK;
        return default_segm == SOURCE ? SOURCE : (tr_cap_[i] > 0 ? SOURCE : default_segm);
    }

private:
    struct Arc { int to, next; captype r; };
    std::vector<tcaptype> tr_cap_;
    std::vector<int> head_;
    std::vector<Arc> arcs_;
    std::vector<char> reach_sink_;
    flowtype flow_;
    bool solved_;
};
