// bvh_quality.cpp — CPU experiment (not product code): how many BVH4 node visits / triangle tests per ray do
// different BVH2 topologies and 4-wide collapses need on the benchmark scene?  Builders: Morton LBVH, PLOC (parallel
// locally-ordered clustering, Meister & Bittner 2018) and a binned-SAH top-down build; optional reference splitting
// (primitives longer than argv[3] are clipped into several references); collapse to BVH4 by "grandchildren at even
// depth", greedily by area, or by dynamic programming over the cuts of <= 4 nodes; traversal like rtb_wavefront.cu
// (sorted children, t_best culling).  Every tree change of round 1 was measured here before it was built on the GPU.
//
//   python tools/dump_corners.py /tmp/bq/corners.bin
//   g++ -O2 -o /tmp/bq/bq tools/experiments/bvh_quality.cpp && /tmp/bq/bq /tmp/bq/corners.bin 6320 0.556
//   (argv[2] = primitives below this index bounce like the Matte teapot, the others mirror; argv[3] = split length)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <random>
#include <vector>

struct V { float x, y, z; };
static V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V operator*(V a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static V unit(V a) { return a * (1.0f / std::sqrt(dot(a, a))); }
struct Box {
    V lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
    void grow(V p) { lo = {std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z)}; hi = {std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z)}; }
    void grow(const Box& b) { grow(b.lo); grow(b.hi); }
    float area() const { V d = hi - lo; return 2.f * (d.x * d.y + d.y * d.z + d.z * d.x); }
};
struct Tri { V a, b, c; };
struct Node { Box box; int l = -1, r = -1; int first = 0, count = 0; };   // leaf iff l < 0

static std::vector<Tri> tris;
static std::vector<Box> tbox;
static std::vector<int> origid;
static float split_len = 0;   // references longer than this along an axis are split
static int split_max_depth = 8;
// clip polygon against axis plane
static std::vector<V> clipPoly(const std::vector<V>& p, int ax, float pos, bool keep_less) {
    std::vector<V> out; size_t n = p.size();
    for (size_t i = 0; i < n; ++i) {
        V a = p[i], b = p[(i + 1) % n]; float da = (&a.x)[ax] - pos, db = (&b.x)[ax] - pos;
        bool ia = keep_less ? da <= 0 : da >= 0, ib = keep_less ? db <= 0 : db >= 0;
        if (ia) out.push_back(a);
        if (ia != ib) { float t = da / (da - db); V c = a + (b - a) * t; (&c.x)[ax] = pos; out.push_back(c); }
    }
    return out;
}
static void split_ref(const std::vector<V>& poly, const Box& clipbox, int depth, std::vector<Box>& out) {
    Box b; for (auto& p : poly) b.grow(p);
    // intersect with clipbox
    b.lo = {std::max(b.lo.x, clipbox.lo.x), std::max(b.lo.y, clipbox.lo.y), std::max(b.lo.z, clipbox.lo.z)};
    b.hi = {std::min(b.hi.x, clipbox.hi.x), std::min(b.hi.y, clipbox.hi.y), std::min(b.hi.z, clipbox.hi.z)};
    V d = b.hi - b.lo; int ax = d.x > d.y ? (d.x > d.z ? 0 : 2) : (d.y > d.z ? 1 : 2);
    float len = (&d.x)[ax];
    if (depth >= split_max_depth || len <= split_len) { out.push_back(b); return; }
    float pos = (&b.lo.x)[ax] + len * 0.5f;
    auto L = clipPoly(poly, ax, pos, true), R = clipPoly(poly, ax, pos, false);
    Box bl = b, br = b; (&bl.hi.x)[ax] = pos; (&br.lo.x)[ax] = pos;
    if (L.size() >= 3) split_ref(L, bl, depth + 1, out);
    if (R.size() >= 3) split_ref(R, br, depth + 1, out);
}


// ---------------- builders: all return a BVH2 over `order` (a permutation of primitive ids) ----------------
struct Bvh2 { std::vector<Node> n; std::vector<int> order; int root = 0; };

static uint64_t spread21(uint32_t x) {
    uint64_t v = x & 0x1fffffu;
    v = (v | (v << 32)) & 0x1f00000000ffffull; v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full; v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull; return v;
}
static std::vector<uint64_t> morton_sorted(std::vector<int>& order) {
    Box sb; for (auto& b : tbox) sb.grow(b);
    int n = (int)tris.size();
    std::vector<uint64_t> key(n);
    for (int i = 0; i < n; ++i) {
        V c = (tbox[i].lo + tbox[i].hi) * 0.5f;
        float u[3] = {(c.x - sb.lo.x) / (sb.hi.x - sb.lo.x), (c.y - sb.lo.y) / (sb.hi.y - sb.lo.y), (c.z - sb.lo.z) / (sb.hi.z - sb.lo.z)};
        uint32_t q[3]; for (int k = 0; k < 3; ++k) q[k] = (uint32_t)std::min(std::max(u[k] * 2097152.f, 0.f), 2097151.f);
        key[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    }
    order.resize(n); for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
    std::vector<uint64_t> ks(n); for (int i = 0; i < n; ++i) ks[i] = key[order[i]];
    return ks;
}

static int collapse_leafmax = 4;
static int lamb_below = 6320;

// generic: given a binary topology over sorted leaves expressed as a recursive function, emit nodes with leaf collapse
static Bvh2 build_lbvh() {
    Bvh2 t; auto ks = morton_sorted(t.order);
    std::function<int(int, int)> rec = [&](int lo, int hi) -> int {   // [lo, hi]
        int id = (int)t.n.size(); t.n.emplace_back();
        Box b; for (int i = lo; i <= hi; ++i) b.grow(tbox[t.order[i]]);
        t.n[id].box = b;
        if (hi - lo + 1 <= collapse_leafmax) { t.n[id].first = lo; t.n[id].count = hi - lo + 1; return id; }
        // split at the highest differing bit (Karras), ties (equal keys) -> middle by index bits
        uint64_t a = ks[lo], z = ks[hi];
        int split;
        if (a == z) split = (lo + hi) / 2;
        else {
            int pre = __builtin_clzll(a ^ z);
            int s = lo, step = hi - lo;
            do { step = (step + 1) >> 1; int ns = s + step; if (ns < hi && __builtin_clzll(a ^ ks[ns]) > pre) s = ns; } while (step > 1);
            split = s;
        }
        int l = rec(lo, split), r = rec(split + 1, hi);
        t.n[id].l = l; t.n[id].r = r; return id;
    };
    t.root = rec(0, (int)tris.size() - 1);
    return t;
}

static Bvh2 build_sah() {
    Bvh2 t; int n = (int)tris.size(); t.order.resize(n); for (int i = 0; i < n; ++i) t.order[i] = i;
    std::function<int(int, int)> rec = [&](int lo, int hi) -> int {
        int id = (int)t.n.size(); t.n.emplace_back();
        Box b, cb; for (int i = lo; i <= hi; ++i) { b.grow(tbox[t.order[i]]); cb.grow((tbox[t.order[i]].lo + tbox[t.order[i]].hi) * 0.5f); }
        t.n[id].box = b;
        int cnt = hi - lo + 1;
        if (cnt <= collapse_leafmax) {
            // leaf unless splitting pays (cost model: node 1, tri 1.5)
            if (cnt <= 1) { t.n[id].first = lo; t.n[id].count = cnt; return id; }
        }
        float best = 1e30f; int baxis = -1; float bpos = 0;
        const int NB = 32;
        for (int ax = 0; ax < 3; ++ax) {
            float l = (&cb.lo.x)[ax], h = (&cb.hi.x)[ax]; if (h <= l) continue;
            Box bb[NB]; int bc[NB] = {0};
            for (int i = lo; i <= hi; ++i) { V c = (tbox[t.order[i]].lo + tbox[t.order[i]].hi) * 0.5f; int k = std::min(NB - 1, (int)(((&c.x)[ax] - l) / (h - l) * NB)); bb[k].grow(tbox[t.order[i]]); bc[k]++; }
            float la[NB], ra[NB]; int lc[NB], rc[NB]; Box acc; int c = 0;
            for (int k = 0; k < NB; ++k) { if (bc[k]) acc.grow(bb[k]); c += bc[k]; la[k] = c ? acc.area() : 0; lc[k] = c; }
            acc = Box(); c = 0;
            for (int k = NB - 1; k >= 0; --k) { if (bc[k]) acc.grow(bb[k]); c += bc[k]; ra[k] = c ? acc.area() : 0; rc[k] = c; }
            for (int k = 0; k + 1 < NB; ++k) { if (!lc[k] || !rc[k + 1]) continue; float cost = la[k] * lc[k] + ra[k + 1] * rc[k + 1]; if (cost < best) { best = cost; baxis = ax; bpos = l + (h - l) * (k + 1) / NB; } }
        }
        float leaf_cost = b.area() * cnt * 1.0f;
        if (baxis < 0 || (cnt <= collapse_leafmax && 1.0f * b.area() + best >= leaf_cost)) {
            if (cnt <= collapse_leafmax) { t.n[id].first = lo; t.n[id].count = cnt; return id; }
            // cannot split by centroid: median
            int mid = (lo + hi) / 2; int l2 = rec(lo, mid), r2 = rec(mid + 1, hi); t.n[id].l = l2; t.n[id].r = r2; return id;
        }
        int mid = (int)(std::partition(t.order.begin() + lo, t.order.begin() + hi + 1, [&](int p) { V c = (tbox[p].lo + tbox[p].hi) * 0.5f; return (&c.x)[baxis] < bpos; }) - t.order.begin());
        if (mid == lo || mid == hi + 1) mid = (lo + hi + 1) / 2;
        int l2 = rec(lo, mid - 1), r2 = rec(mid, hi); t.n[id].l = l2; t.n[id].r = r2; return id;
    };
    t.root = rec(0, n - 1);
    return t;
}

// PLOC: clusters in Morton order, nearest neighbour within +-R by merged surface area, mutual pairs merge.
static Bvh2 build_ploc(int R, bool sah_collapse) {
    std::vector<int> order; morton_sorted(order);
    int n = (int)tris.size();
    struct C { Box b; int l, r, prim, size; };
    std::vector<C> nodes;    // full binary tree, leaves first
    std::vector<int> cur(n);
    for (int i = 0; i < n; ++i) { nodes.push_back({tbox[order[i]], -1, -1, order[i], 1}); cur[i] = i; }
    while (cur.size() > 1) {
        int m = (int)cur.size();
        std::vector<int> nn(m);
        for (int i = 0; i < m; ++i) {
            float best = 1e30f; int bj = -1;
            for (int j = std::max(0, i - R); j <= std::min(m - 1, i + R); ++j) {
                if (j == i) continue;
                Box b = nodes[cur[i]].b; b.grow(nodes[cur[j]].b); float a = b.area();
                if (a < best) { best = a; bj = j; }
            }
            nn[i] = bj;
        }
        std::vector<int> next;
        for (int i = 0; i < m; ++i) {
            int j = nn[i];
            if (nn[j] == i) { if (i < j) { Box b = nodes[cur[i]].b; b.grow(nodes[cur[j]].b); nodes.push_back({b, cur[i], cur[j], -1, nodes[cur[i]].size + nodes[cur[j]].size}); next.push_back((int)nodes.size() - 1); } }
            else next.push_back(cur[i]);
        }
        cur.swap(next);
    }
    // emit with collapse: DFS assigns contiguous primitive ranges
    Bvh2 t;
    std::function<void(int, std::vector<int>&)> gather = [&](int c, std::vector<int>& out) { if (nodes[c].l < 0) out.push_back(nodes[c].prim); else { gather(nodes[c].l, out); gather(nodes[c].r, out); } };
    std::function<int(int)> rec = [&](int c) -> int {
        int id = (int)t.n.size(); t.n.emplace_back(); t.n[id].box = nodes[c].b;
        bool leaf = nodes[c].size <= collapse_leafmax;
        if (leaf && sah_collapse && nodes[c].size > 1) {
            // keep the split if SAH says the two children are cheaper than one leaf
            float cl = nodes[nodes[c].l].b.area() * nodes[nodes[c].l].size + nodes[nodes[c].r].b.area() * nodes[nodes[c].r].size;
            if (1.0f * nodes[c].b.area() + cl < nodes[c].b.area() * nodes[c].size) leaf = false;
        }
        if (leaf) { t.n[id].first = (int)t.order.size(); gather(c, t.order); t.n[id].count = nodes[c].size; return id; }
        int l = rec(nodes[c].l), r = rec(nodes[c].r); t.n[id].l = l; t.n[id].r = r; return id;
    };
    t.root = rec(cur[0]);
    return t;
}


// ---------------- insertion-based optimisation (Bittner et al. 2013; the one-subtree form of Meister & Bittner 2018) ----------------
// Every node N (with its whole subtree) is taken out of the tree — its parent P disappears, its sibling takes P's place — and put
// back as the sibling of the node X for which the sum of the inner nodes' surface areas grows least (branch and bound from the
// root); the move is kept when the tree's summed area shrinks.  Leaves keep their primitives.
static void reinsertion_opt(Bvh2& t, int passes) {
    int N = (int)t.n.size();
    std::vector<int> parent(N, -1);
    for (int i = 0; i < N; ++i) if (t.n[i].l >= 0) { parent[t.n[i].l] = i; parent[t.n[i].r] = i; }
    auto uni = [](const Box& a, const Box& b) { Box c = a; c.grow(b); return c; };
    auto refit_up = [&](int v) { while (v >= 0) { t.n[v].box = uni(t.n[t.n[v].l].box, t.n[t.n[v].r].box); v = parent[v]; } };
    auto total = [&]() { double c = 0; for (int i = 0; i < N; ++i) if (t.n[i].l >= 0 && (parent[i] >= 0 || i == t.root)) c += t.n[i].box.area(); return c; };
    std::mt19937 rng(7);
    for (int pass = 0; pass < passes; ++pass) {
        double before = total();
        std::vector<int> cand; for (int i = 0; i < N; ++i) if (i != t.root && parent[i] >= 0 && parent[i] != t.root) cand.push_back(i);
        std::sort(cand.begin(), cand.end(), [&](int a, int b) { return t.n[parent[a]].box.area() > t.n[parent[b]].box.area(); });
        int moved = 0;
        for (int n : cand) {
            int P = parent[n]; if (P < 0 || P == t.root) continue;
            int G = parent[P]; int S = t.n[P].l == n ? t.n[P].r : t.n[P].l;
            // cost of the inner nodes on the path before
            double path_before = t.n[P].box.area(); for (int a = G; a >= 0; a = parent[a]) path_before += t.n[a].box.area();
            // remove
            if (t.n[G].l == P) t.n[G].l = S; else t.n[G].r = S; parent[S] = G;
            std::vector<std::pair<int, Box>> saved; for (int a = G; a >= 0; a = parent[a]) saved.push_back({a, t.n[a].box});
            refit_up(G);
            double path_after = 0; for (int a = G; a >= 0; a = parent[a]) path_after += t.n[a].box.area();
            double gain = path_before - path_after;
            // best position: branch and bound on the induced cost
            const Box nb = t.n[n].box; float na = nb.area();
            double best = 1e300; int bx = -1;
            std::vector<std::pair<double, int>> pq; pq.push_back({0.0, t.root});
            auto cmp = [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first > b.first; };
            while (!pq.empty()) {
                std::pop_heap(pq.begin(), pq.end(), cmp); auto [ind, x] = pq.back(); pq.pop_back();
                if (ind + na >= best) break;
                double direct = uni(t.n[x].box, nb).area();
                if (ind + direct < best) { best = ind + direct; bx = x; }
                if (t.n[x].l >= 0) {
                    double ci = ind + direct - t.n[x].box.area();
                    if (ci + na < best) { pq.push_back({ci, t.n[x].l}); std::push_heap(pq.begin(), pq.end(), cmp); pq.push_back({ci, t.n[x].r}); std::push_heap(pq.begin(), pq.end(), cmp); }
                }
            }
            if (bx >= 0 && best < gain * (1.0 - 1e-6) && bx != S) {
                int XP = parent[bx];
                t.n[P].l = bx; t.n[P].r = n; parent[P] = XP; parent[bx] = P; parent[n] = P;
                if (XP < 0) t.root = P; else { if (t.n[XP].l == bx) t.n[XP].l = P; else t.n[XP].r = P; }
                refit_up(P); ++moved;
            } else {   // put it back
                for (auto& sv : saved) t.n[sv.first].box = sv.second;
                if (t.n[G].l == S) t.n[G].l = P; else t.n[G].r = P; parent[P] = G; parent[S] = P;
                t.n[P].l = S; t.n[P].r = n;
            }
        }
        double after = total();
        printf("  reinsertion pass %d: moved %d, inner area %.1f -> %.1f (%.2f %%)\n", pass, moved, before, after, 100.0 * (after - before) / before);
        if (moved == 0 || (before - after) < 1e-4 * before) break;
    }
}

// ---------------- BVH4 collapse + traversal ----------------
struct N4 { Box b[4]; int code[4]; /* 0 empty, <0 leaf: -(idx2+1), >0: n4 index */ };
struct Bvh4 { std::vector<N4> n; const Bvh2* src; };
static int greedy4 = 0;
static Bvh4 collapse4(const Bvh2& t) {
    Bvh4 q; q.src = &t;
    if (greedy4 == 2) {
        // DP: cost4[v] = A(v) + min over cuts (<=4 nodes) of sum of child costs; leaf cost = 1.2 * A(v) * count
        int N = (int)t.n.size();
        std::vector<double> cost(N, 0.0);
        std::vector<std::vector<int>> cut(N);
        // post-order
        std::vector<int> order; order.reserve(N);
        { std::vector<int> st{t.root}; while (!st.empty()) { int v = st.back(); st.pop_back(); order.push_back(v); if (t.n[v].l >= 0) { st.push_back(t.n[v].l); st.push_back(t.n[v].r); } } }
        for (int k = N - 1; k >= 0; --k) {
            int v = order[k];
            if (t.n[v].l < 0) { cost[v] = 1.2 * t.n[v].box.area() * t.n[v].count; continue; }
            // enumerate cuts by BFS expansion: state = list of nodes; expand any internal node while size < 4
            double best = 1e300; std::vector<int> bestc;
            std::function<void(std::vector<int>&, size_t)> rec = [&](std::vector<int>& c, size_t start) {
                double sum = 0; for (int x : c) sum += cost[x];
                if (sum < best) { best = sum; bestc = c; }
                if (c.size() >= 4) return;
                for (size_t i = start; i < c.size(); ++i) {
                    int x = c[i]; if (t.n[x].l < 0) continue;
                    std::vector<int> d = c; d[i] = t.n[x].l; d.push_back(t.n[x].r);
                    rec(d, i);
                }
            };
            std::vector<int> c0{t.n[v].l, t.n[v].r};
            rec(c0, 0);
            cost[v] = t.n[v].box.area() * 1.0 + best;
            cut[v] = bestc;
        }
        std::function<int(int)> rec2 = [&](int c) -> int {
            int id = (int)q.n.size(); q.n.emplace_back(); for (int k = 0; k < 4; ++k) q.n[id].code[k] = 0;
            std::vector<int> ent = t.n[c].l < 0 ? std::vector<int>{c} : cut[c];
            for (size_t k = 0; k < ent.size(); ++k) {
                q.n[id].b[k] = t.n[ent[k]].box;
                if (t.n[ent[k]].l < 0) q.n[id].code[k] = -(ent[k] + 1);
                else { int ci = rec2(ent[k]); q.n[id].code[k] = ci; }
            }
            return id;
        };
        rec2(t.root);
        return q;
    }
    if (greedy4) {
        // expand the entry with the largest surface area until four slots are filled
        std::function<int(int)> rec = [&](int c) -> int {
            int id = (int)q.n.size(); q.n.emplace_back(); for (int k = 0; k < 4; ++k) q.n[id].code[k] = 0;
            std::vector<int> ent;
            if (t.n[c].l < 0) ent.push_back(c); else { ent.push_back(t.n[c].l); ent.push_back(t.n[c].r); }
            while (ent.size() < 4) {
                int bi = -1; float ba = -1;
                for (size_t k = 0; k < ent.size(); ++k) if (t.n[ent[k]].l >= 0 && t.n[ent[k]].box.area() > ba) { ba = t.n[ent[k]].box.area(); bi = (int)k; }
                if (bi < 0) break;
                int e = ent[bi]; ent[bi] = t.n[e].l; ent.push_back(t.n[e].r);
            }
            for (size_t k = 0; k < ent.size(); ++k) {
                q.n[id].b[k] = t.n[ent[k]].box;
                if (t.n[ent[k]].l < 0) q.n[id].code[k] = -(ent[k] + 1);
                else { int ci = rec(ent[k]); q.n[id].code[k] = ci; }
            }
            return id;
        };
        rec(t.root);
        return q;
    }
    std::function<int(int)> rec = [&](int c) -> int {
        int id = (int)q.n.size(); q.n.emplace_back(); for (int k = 0; k < 4; ++k) q.n[id].code[k] = 0;
        std::vector<int> ent;
        const Node& nd = t.n[c];
        if (nd.l < 0) ent.push_back(c);
        else for (int ch : {nd.l, nd.r}) { if (t.n[ch].l < 0) ent.push_back(ch); else { ent.push_back(t.n[ch].l); ent.push_back(t.n[ch].r); } }
        for (size_t k = 0; k < ent.size(); ++k) {
            q.n[id].b[k] = t.n[ent[k]].box;
            if (t.n[ent[k]].l < 0) q.n[id].code[k] = -(ent[k] + 1);
            else { int ci = rec(ent[k]); q.n[id].code[k] = ci; }
        }
        return id;
    };
    rec(t.root);
    return q;
}
static bool mt(const Tri& tr, V o, V d, float& t) {
    V e1 = tr.b - tr.a, e2 = tr.c - tr.a, p = cross(d, e2); float det = dot(e1, p); if (std::fabs(det) < 1e-12f) return false;
    float inv = 1.f / det; V s = o - tr.a; float u = dot(s, p) * inv; if (u < 0 || u > 1) return false;
    V q = cross(s, e1); float v = dot(d, q) * inv; if (v < 0 || u + v > 1) return false; t = dot(e2, q) * inv; return t > 1e-4f;
}
struct Cnt { double nodes = 0, tris = 0, rays = 0; };
static int trace(const Bvh4& q, V o, V d, float& tbest, Cnt& c) {
    V inv = {1.f / d.x, 1.f / d.y, 1.f / d.z};
    int hit = -1; tbest = 1e30f;
    int stack[128]; int sp = 0; int cur = 0;   // cur >= 0: n4 node; cur < 0 leaf
    c.rays++;
    for (;;) {
        if (cur >= 0) {
            const N4& nd = q.n[cur]; c.nodes++;
            float key[4]; int code[4]; int nh = 0;
            for (int k = 0; k < 4; ++k) {
                if (!nd.code[k]) continue;
                float x0 = (nd.b[k].lo.x - o.x) * inv.x, x1 = (nd.b[k].hi.x - o.x) * inv.x, y0 = (nd.b[k].lo.y - o.y) * inv.y, y1 = (nd.b[k].hi.y - o.y) * inv.y, z0 = (nd.b[k].lo.z - o.z) * inv.z, z1 = (nd.b[k].hi.z - o.z) * inv.z;
                float tn = std::max(std::max(std::min(x0, x1), std::min(y0, y1)), std::max(std::min(z0, z1), 0.f));
                float tf = std::min(std::min(std::max(x0, x1), std::max(y0, y1)), std::max(z0, z1));
                if (tn <= tf && tn <= tbest) { key[nh] = tn; code[nh] = nd.code[k]; nh++; }
            }
            for (int a = 1; a < nh; ++a) for (int b = a; b > 0 && key[b] < key[b - 1]; --b) { std::swap(key[b], key[b - 1]); std::swap(code[b], code[b - 1]); }
            if (!nh) { if (!sp) break; cur = stack[--sp]; if (cur > 0) {} continue; }
            for (int k = nh - 1; k >= 1; --k) stack[sp++] = code[k];
            cur = code[0];
            if (cur > 0) continue;
        }
        // leaf (cur < 0)
        const Node& lf = q.src->n[-cur - 1];
        for (int k = lf.first; k < lf.first + lf.count; ++k) { float t; c.tris++; if (mt(tris[q.src->order[k]], o, d, t) && t < tbest) { tbest = t; hit = q.src->order[k]; } }
        if (!sp) break;
        cur = stack[--sp];
    }
    return hit;
}

static float sah(const Bvh2& t) { double c = 0; float ra = t.n[t.root].box.area(); for (auto& n : t.n) c += n.box.area() / ra * (n.l < 0 ? 1.5 * n.count : 1.0); return (float)c; }

int main(int argc, char** argv) {
    if (argc > 2) lamb_below = atoi(argv[2]);
    FILE* f = fopen(argv[1], "rb"); std::vector<float> buf; float tmp[9];
    while (fread(tmp, 4, 9, f) == 9) { tris.push_back({{tmp[0], tmp[1], tmp[2]}, {tmp[3], tmp[4], tmp[5]}, {tmp[6], tmp[7], tmp[8]}}); }
    fclose(f);
    if (argc > 3) split_len = atof(argv[3]);
    {
        std::vector<Tri> nt; int id = 0;
        for (auto& t : tris) {
            Box b; b.grow(t.a); b.grow(t.b); b.grow(t.c);
            if (split_len > 0) {
                std::vector<Box> parts; Box inf; inf.lo = {-1e30f,-1e30f,-1e30f}; inf.hi = {1e30f,1e30f,1e30f};
                split_ref({t.a, t.b, t.c}, inf, 0, parts);
                for (auto& pb : parts) { nt.push_back(t); tbox.push_back(pb); origid.push_back(id); }
            } else { nt.push_back(t); tbox.push_back(b); origid.push_back(id); }
            ++id;
        }
        printf("%zu refs from %zu tris\n", nt.size(), tris.size());
        tris.swap(nt);
    }
    printf("%zu triangles\n", tris.size());
    struct Cfg { const char* name; Bvh2 t; };
    std::vector<Cfg> cfgs;
    cfgs.push_back({"LBVH (GPU builder today)", build_lbvh()});

    cfgs.push_back({"PLOC r=16", build_ploc(16, false)});

    cfgs.push_back({"PLOC r=16 + SAH leaf split", build_ploc(16, true)});
    cfgs.push_back({"binned SAH top-down", build_sah()});
    { Bvh2 o = build_sah(); reinsertion_opt(o, 8); cfgs.push_back({"binned SAH + reinsertion", o}); }
    { Bvh2 o = build_ploc(16, true); reinsertion_opt(o, 8); cfgs.push_back({"PLOC + SAH leaf + reinsertion", o}); }
    // camera of main.rs at 4K, every 6th pixel
    V orig{2.28125f, -0.5f, 0.f}, cam{2.f, 0.f, -0.5f}, vu{0, 1, 0}, vv{-0.5625f, 0, 0};
    int W = 3840, H = 2160;
    for (int pass = 0; pass < 3; ++pass)
    for (auto& c : cfgs) {
        greedy4 = pass;
        if (greedy4 == 1) printf("[greedy BVH4 collapse] "); if (greedy4 == 2) printf("[DP BVH4 collapse] ");
        Bvh4 q = collapse4(c.t);
        Cnt prim, bnc, bnd, bnt; std::mt19937 rng(1); std::uniform_real_distribution<float> U(-0.5f, 0.5f);
        int nleaf = 0, maxleaf = 0; for (auto& n : c.t.n) if (n.l < 0) { nleaf++; maxleaf = std::max(maxleaf, n.count); }
        for (int r = 0; r < H; r += 12) for (int col = 0; col < W; col += 12) {
            V p = orig + vu * ((col + 0.5f) / W) + vv * ((r + 0.5f) / H); V d = unit(p - cam); float t;
            int h = trace(q, p, d, t, prim);
            int depth = 1;
            while (h >= 0 && depth < 5) {   // bounce: teapot (h < 6320) lambertian, disks mirror
                V hp = p + d * t; const Tri& tr = tris[h]; V n = unit(cross(tr.b - tr.a, tr.c - tr.a)); if (dot(n, d) > 0) n = n * -1.f;
                V nd;
                if (origid[h] < lamb_below) { V rv = unit(V{U(rng), U(rng), U(rng)}); nd = unit(n + rv); } else nd = unit(d - n * (2.f * dot(d, n)));
                bool fromdisk = origid[h] >= lamb_below; p = hp + nd * 0.001f; d = nd; { Cnt one; h = trace(q, p, d, t, one); bnc.nodes += one.nodes; bnc.tris += one.tris; bnc.rays++; Cnt& w = fromdisk ? bnd : bnt; w.nodes += one.nodes; w.tris += one.tris; w.rays++; } depth++;
            }
        }
        printf("%-30s nodes2 %5zu leaves %4d maxleaf %d n4 %4zu SAH %.1f | primary: %.2f visits %.2f tris | bounce (%.0f rays): %.2f visits %.2f tris\n", c.name, c.t.n.size(), nleaf, maxleaf,
               q.n.size(), sah(c.t), prim.nodes / prim.rays, prim.tris / prim.rays, bnc.rays, bnc.nodes / bnc.rays, bnc.tris / bnc.rays);
        printf("      from disk (%.0f): %.2f visits %.2f tris | from teapot (%.0f): %.2f visits %.2f tris\n", bnd.rays, bnd.nodes / bnd.rays, bnd.tris / bnd.rays, bnt.rays, bnt.nodes / bnt.rays, bnt.tris / bnt.rays);
    }
    return 0;
}
