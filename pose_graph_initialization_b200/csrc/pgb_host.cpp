// pgb_host.cpp — host side of the B200 path (include/pgb.h): similarity queue, pose graph, visibility
// table, A* and the sequential commit of the reference, reorganised as speculative waves.
//
// Exactness argument (SURVEY §7 hard part 3): the only graph reads of AStarTraversal::getPath are
// getEdgesByVertex(v) for every EXPANDED vertex v (graph_traversal.h:817) and the (immutable) edges of the
// found path (graph_traversal.h:323-328).  Committing an edge (u,v) only appends to the edge lists of u
// and v.  Every vertex carries the commit stamp of its last modification; a speculative search made at
// stamp s is still what the sequential reference would compute at commit time iff no expanded vertex has a
// stamp > s and the hasLink answer is unchanged (hasLink is monotone false->true).  Otherwise the search is
// re-run on the exact sequential state; if it yields a different hypothesis whose verdict is not cached the
// commit stops there and the rest of the wave is re-queued in order.
//
// Data structures are flat (arena A* nodes, bit-matrix visibility, per-vertex edge index vectors) but the
// heap is driven by the same std::push_heap/std::pop_heap calls std::priority_queue makes, so ties break
// exactly as in the reference.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <queue>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/pgb.h"

namespace {

// ---- pose algebra (Sophus/Eigen semantics; scalar evaluation order as documented in DESIGN.md) --------
struct SE3 {
    double q[4];  // x y z w
    double t[3];
};
inline SE3 se3Identity() { return SE3{{0, 0, 0, 1}, {0, 0, 0}}; }
inline void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
inline void quatRotate(const double q[4], const double p[3], double o[3])
{
    double uv[3], c2[3];
    cross3(q, p, uv);
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    cross3(q, uv, c2);
    o[0] = (p[0] + q[3] * uv[0]) + c2[0];
    o[1] = (p[1] + q[3] * uv[1]) + c2[1];
    o[2] = (p[2] + q[3] * uv[2]) + c2[2];
}
inline SE3 se3Mul(const SE3 &a, const SE3 &b)  // graph_traversal.h:341-344
{
    SE3 r;
    const double *A = a.q, *B = b.q;
    double w = A[3] * B[3] - A[0] * B[0] - A[1] * B[1] - A[2] * B[2];
    double x = A[3] * B[0] + A[0] * B[3] + A[1] * B[2] - A[2] * B[1];
    double y = A[3] * B[1] + A[1] * B[3] + A[2] * B[0] - A[0] * B[2];
    double z = A[3] * B[2] + A[2] * B[3] + A[0] * B[1] - A[1] * B[0];
    const double n = std::sqrt((x * x + y * y) + (z * z + w * w));
    r.q[0] = x / n; r.q[1] = y / n; r.q[2] = z / n; r.q[3] = w / n;
    double rt[3];
    quatRotate(a.q, b.t, rt);
    r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
    return r;
}
inline SE3 se3Inverse(const SE3 &a)
{
    SE3 r;
    r.q[0] = -a.q[0]; r.q[1] = -a.q[1]; r.q[2] = -a.q[2]; r.q[3] = a.q[3];
    const double nt[3] = {a.t[0] * -1.0, a.t[1] * -1.0, a.t[2] * -1.0};
    quatRotate(r.q, nt, r.t);
    return r;
}

struct Edge {
    uint32_t src, dst;
    SE3 T;
    double score;
    uint32_t inlierNumber, nCorr;
    uint8_t branch;
};

inline uint64_t edgeKey(uint32_t s, uint32_t d) { return ((uint64_t)s << 32) | d; }

struct Graph {
    std::vector<Edge> edges;                       // commit order (= PoseGraph::edges_ids)
    std::vector<std::vector<uint32_t>> byVertex;   // edge indices per vertex in insertion order (pose_graph.h:219-220)
    std::unordered_map<uint64_t, uint32_t> lookup; // (src,dst) -> edge index
    bool hasEdge(uint32_t s, uint32_t d) const { return lookup.find(edgeKey(s, d)) != lookup.end(); }
    const Edge *find(uint32_t s, uint32_t d) const
    {
        auto it = lookup.find(edgeKey(s, d));
        return it == lookup.end() ? nullptr : &edges[it->second];
    }
};

// visibility_table.h:45-171 with the map as a bit matrix and the neighbour std::sets as sorted bit rows
// (ascending iteration == std::set order).  pair2 = (third,third) entries (:101) can never be queried by
// hasLink (from == to returns early, :152), so they are not materialised; `!hasPair1 || !hasPair1` (:108)
// is kept as is.
struct Visibility {
    uint32_t V = 0, words = 0;
    std::vector<uint64_t> link, nb;
    std::vector<uint32_t> fifo;
    void init(uint32_t v)
    {
        V = v;
        words = (v + 63) / 64;
        link.assign((size_t)V * words, 0);
        nb.assign((size_t)V * words, 0);
    }
    bool get(const std::vector<uint64_t> &m, uint32_t r, uint32_t c) const { return (m[(size_t)r * words + (c >> 6)] >> (c & 63)) & 1; }
    void set(std::vector<uint64_t> &m, uint32_t r, uint32_t c) { m[(size_t)r * words + (c >> 6)] |= 1ull << (c & 63); }
    bool hasLink(uint32_t a, uint32_t b) const
    {
        if (a == b) return false;
        return get(link, std::min(a, b), std::max(a, b));
    }
    void pushRow(uint32_t r)
    {
        const uint64_t *row = &nb[(size_t)r * words];
        for (uint32_t w = 0; w < words; w++) {
            uint64_t x = row[w];
            while (x) {
                fifo.push_back(w * 64 + (uint32_t)__builtin_ctzll(x));
                x &= x - 1;
            }
        }
    }
    bool addLink(uint32_t from_, uint32_t to_)
    {
        if (from_ == to_) return false;
        const uint32_t from = std::min(from_, to_), to = std::max(from_, to_);
        set(nb, from, to);
        set(nb, to, from);
        if (get(link, from, to)) return false;
        set(link, from, to);
        fifo.clear();
        pushRow(from);
        pushRow(to);
        for (size_t head = 0; head < fifo.size(); ++head) {
            const uint32_t v = fifo[head];
            if (v == from || v == to) continue;
            const uint32_t first = std::min(v, from), second = std::max(v, from);
            if (!get(link, first, second)) {
                set(link, first, second);
                pushRow(v);
            }
            set(nb, v, from);
            set(nb, v, to);
        }
        return true;
    }
};

struct AStarOut {
    bool found = false;
    SE3 pose = se3Identity();
    uint32_t touched = 0;
    std::vector<uint32_t> expanded;  // vertices whose edge lists were read
};

struct HeapItem {
    double f;
    uint32_t node;
};
struct ArenaNode {
    uint32_t vertex, parent, depth;
    double c0, c1;
};

struct AStarScratch {
    std::vector<HeapItem> heap;
    std::vector<ArenaNode> arena;
    std::vector<uint32_t> mark;  // per-vertex stamp of "expanded in this search"
    uint32_t epoch = 0;
    std::vector<uint32_t> path;
};

// AStarTraversal<ImageSimilarityHeuristics>::getPath with the arguments of pose_graph_builder.h:834-841.
void aStar(const Graph &g, const double *sim, uint32_t V, uint32_t from, uint32_t to, size_t maxDepth, double weight,
           AStarScratch &S, AStarOut &out)
{
    out.found = false;
    out.touched = 0;
    out.expanded.clear();
    if (S.mark.size() != V) { S.mark.assign(V, 0); S.epoch = 0; }
    if (++S.epoch == 0) { std::fill(S.mark.begin(), S.mark.end(), 0); S.epoch = 1; }
    S.heap.clear();
    S.arena.clear();
    auto cmp = [](const HeapItem &a, const HeapItem &b) { return a.f < b.f; };  // graph_traversal.h:669-673
    S.arena.push_back(ArenaNode{from, UINT32_MAX, 0, 1.0, 0.0});  // cost (1,0,0)  :721
    S.heap.push_back(HeapItem{0.0, 0});
    std::push_heap(S.heap.begin(), S.heap.end(), cmp);
    const double oneMinusWeight = 1.0 - weight;
    const std::vector<uint32_t> *edges = nullptr;  // `edges` keeps its previous content if the vertex has none (pose_graph.h:145-146)
    while (!S.heap.empty()) {
        const uint32_t ni = S.heap.front().node;
        ++out.touched;  // :750
        std::pop_heap(S.heap.begin(), S.heap.end(), cmp);
        S.heap.pop_back();
        const ArenaNode node = S.arena[ni];
        if (node.depth > maxDepth) continue;  // :755
        const uint32_t v = node.vertex;
        if (v == to) {  // :766
            S.path.clear();
            for (uint32_t k = ni; k != UINT32_MAX; k = S.arena[k].parent) S.path.push_back(S.arena[k].vertex);
            std::reverse(S.path.begin(), S.path.end());
            SE3 pose = se3Identity();  // recoverPath :304
            bool ok = true;
            for (size_t i = 1; i < S.path.size(); ++i) {
                const uint32_t s = S.path[i - 1], d = S.path[i];
                if (const Edge *e = g.find(s, d))
                    pose = se3Mul(e->T, pose);  // :344
                else if (const Edge *e2 = g.find(d, s))
                    pose = se3Mul(se3Inverse(e2->T), pose);  // :342
                else { ok = false; break; }
            }
            if (ok) {
                out.found = true;
                out.pose = pose;
                break;  // one path is tested, then the search ends (:792-800 with kMaximumPathNumber = 1)
            }
            continue;
        }
        S.mark[v] = S.epoch;  // nodeStates[v] = Open  :814
        out.expanded.push_back(v);
        if (!g.byVertex[v].empty()) edges = &g.byVertex[v];  // :817
        if (node.depth < maxDepth && edges) {  // :820
            for (uint32_t ei : *edges) {
                const Edge &e = g.edges[ei];
                if (e.score < 0.0) continue;  // kMinimumInlierRatio is `const bool` receiving 0.0 (:613, :839)
                const uint32_t next = (v == e.dst) ? e.src : e.dst;  // :838-840
                const double edgeCost = node.c0 > e.score ? e.score : node.c0;  // MIN :843
                double h = sim[(size_t)next * V + to];
                h = h < 0.0 ? 0.0 : (1.0 < h ? 1.0 : h);  // std::clamp(similarity, 0, 1)  :594
                const double nextToDest = node.c1 < h ? h : node.c1;  // MAX :847
                const double combined = weight * edgeCost + oneMinusWeight * nextToDest;  // :851-852
                if (S.mark[next] != S.epoch) {  // nodeStates.find(next) == end  :855-856
                    S.arena.push_back(ArenaNode{next, ni, node.depth + 1, edgeCost, nextToDest});
                    S.heap.push_back(HeapItem{combined, (uint32_t)S.arena.size() - 1});
                    std::push_heap(S.heap.begin(), S.heap.end(), cmp);
                }
            }
        }
    }
}

struct Item {
    uint32_t pairId = UINT32_MAX;  // UINT32_MAX: queued pair without registered correspondences
    uint32_t src = 0, dst = 0;
    uint32_t nCorr = 0;
    bool specDone = false;
    bool visible = false;
    bool hasHyp = false;
    SE3 hyp = se3Identity();
    uint64_t stamp = 0;  // commit counter when the speculation was made
    std::vector<uint32_t> expanded;
    uint32_t touched = 0;
    bool needGpu = false;
    int verdictSlot = -1;
};

struct HypKey {
    uint32_t pair;
    uint64_t bits[7];
    bool operator==(const HypKey &o) const { return pair == o.pair && !memcmp(bits, o.bits, sizeof bits); }
};
struct HypKeyHash {
    size_t operator()(const HypKey &k) const
    {
        uint64_t h = 1469598103934665603ull ^ k.pair;
        for (int i = 0; i < 7; i++) { h ^= k.bits[i]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
struct PathVerdict {
    bool ok;
    pgi_verdict v;
};

inline HypKey makeKey(uint32_t pair, const SE3 &h)
{
    HypKey k;
    k.pair = pair;
    memcpy(k.bits, h.q, 32);
    memcpy(k.bits + 4, h.t, 24);
    return k;
}

double nowSec() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

struct pgb_builder {
    pgb_config cfg;
    uint32_t V = 0;
    std::vector<double> sim;
    uint64_t P = 0;
    std::vector<uint32_t> pairViews;
    std::vector<uint64_t> mOffset;
    std::unordered_map<uint64_t, uint32_t> pairIndex;
    // similarity-ordered queue, materialised in pop order (the queue is static: imagesimilarity_graph.h:149-157)
    std::vector<std::pair<uint32_t, uint32_t>> order;
    size_t nextInOrder = 0;
    std::deque<Item> pending;  // re-queued items, in queue order, ahead of `order[nextInOrder..]`
    Graph graph;
    Visibility vis;
    std::vector<uint64_t> vertexStamp;
    uint64_t commitStamp = 0;
    std::vector<Item> wave;
    std::vector<pgi_verdict> fbCache;
    std::vector<uint8_t> fbHave;
    std::unordered_map<HypKey, PathVerdict, HypKeyHash> pathCache;
    std::vector<pgb_log> log;
    pgb_counters ctr;
    std::vector<AStarScratch> scratch;
};

namespace {

void speculate(pgb_builder *b, Item &it, AStarScratch &S)
{
    it.visible = b->vis.hasLink(it.src, it.dst);  // pose_graph_builder.h:456-457
    it.hasHyp = false;
    it.expanded.clear();
    it.touched = 0;
    it.stamp = b->commitStamp;
    if (b->cfg.use_path_finding && it.visible && it.pairId != UINT32_MAX && it.nCorr >= b->cfg.minimum_point_number) {
        AStarOut o;
        aStar(b->graph, b->sim.data(), b->V, it.src, it.dst, (size_t)b->cfg.maximum_search_depth,
              b->cfg.traversal_heuristics_weight, S, o);
        it.hasHyp = o.found;
        it.hyp = o.pose;
        it.expanded.swap(o.expanded);
        it.touched = o.touched;
    }
    it.specDone = true;
}

bool speculationHolds(const pgb_builder *b, const Item &it)
{
    if (!it.specDone) return false;
    if (!it.visible && b->vis.hasLink(it.src, it.dst)) return false;
    for (uint32_t v : it.expanded)
        if (b->vertexStamp[v] > it.stamp) return false;
    return true;
}

bool needsGpu(const pgb_builder *b, const Item &it)
{
    if (it.pairId == UINT32_MAX || it.nCorr < b->cfg.minimum_point_number) return false;
    if (b->graph.hasEdge(it.src, it.dst) || b->graph.hasEdge(it.dst, it.src)) return false;
    if (it.hasHyp) {
        auto pc = b->pathCache.find(makeKey(it.pairId, it.hyp));
        if (pc == b->pathCache.end()) return true;
        if (pc->second.ok) return false;
    }
    return !b->fbHave[it.pairId];
}

}  // namespace

extern "C" {

int32_t pgb_create(const pgb_config *cfg, uint64_t n_views, const double *sim, uint64_t n_pairs,
                   const uint32_t *pair_views, const uint64_t *m_offset, pgb_builder **out)
{
    if (!cfg || !sim || !out || (n_pairs && (!pair_views || !m_offset)) || n_views > 0x7fffffffULL) return -1;
    pgb_builder *b = new pgb_builder();
    b->cfg = *cfg;
    if (b->cfg.host_threads <= 0) b->cfg.host_threads = (int32_t)std::max(1u, std::thread::hardware_concurrency());
    b->V = (uint32_t)n_views;
    b->sim.assign(sim, sim + n_views * n_views);
    b->P = n_pairs;
    b->pairViews.assign(pair_views, pair_views + 2 * n_pairs);
    b->mOffset.assign(m_offset, m_offset + n_pairs + 1);
    for (uint64_t p = 0; p < n_pairs; p++) b->pairIndex[edgeKey(pair_views[2 * p], pair_views[2 * p + 1])] = (uint32_t)p;
    // SimilarityTable::loadFromFile queue construction (imagesimilarity_graph.h:141-163): the table is pre-filled
    // with 1.0f (:58-65); (i,j) is queued while row i is filled iff i != j, thr <= s_ij and s_ji != s_ij at that time.
    {
        const size_t V = n_views;
        std::vector<double> table(V * V, (double)1.0f);
        std::priority_queue<std::tuple<double, size_t, size_t>> q;
        for (size_t i = 0; i < V; i++)
            for (size_t j = 0; j < V; j++) {
                table[i * V + j] = sim[i * V + j];
                if (i != j && cfg->similarity_threshold <= table[i * V + j] && table[j * V + i] != table[i * V + j])
                    q.emplace(std::make_tuple(table[i * V + j], i, j));
            }
        b->order.reserve(q.size());
        while (!q.empty()) {  // std::less on the tuple: larger similarity, then larger i, then larger j first (App. A.1)
            b->order.emplace_back((uint32_t)std::get<1>(q.top()), (uint32_t)std::get<2>(q.top()));
            q.pop();
        }
    }
    b->graph.byVertex.resize(n_views);
    b->vis.init((uint32_t)n_views);
    b->vertexStamp.assign(n_views, 0);
    b->fbCache.resize(n_pairs);
    b->fbHave.assign(n_pairs, 0);
    memset(&b->ctr, 0, sizeof b->ctr);
    b->scratch.resize(b->cfg.host_threads);
    *out = b;
    return 0;
}

void pgb_destroy(pgb_builder *b) { delete b; }

uint64_t pgb_remaining(pgb_builder *b) { return b ? b->pending.size() + (b->order.size() - b->nextInOrder) + b->wave.size() : 0; }

int32_t pgb_set_fallback_verdicts(pgb_builder *b, const pgi_verdict *verdicts, uint64_t n_pairs)
{
    if (!b || !verdicts || n_pairs != b->P) return -1;
    for (uint64_t p = 0; p < n_pairs; p++) {
        b->fbCache[p] = verdicts[p];
        b->fbHave[p] = 1;
    }
    return 0;
}

uint32_t pgb_next_wave(pgb_builder *b, uint32_t max_items, pgb_item *items)
{
    if (!b || !items || !b->wave.empty()) return 0;
    const double t0 = nowSec();
    // 1. gather items in queue order: re-queued ones first, then fresh pops
    while (b->wave.size() < max_items) {
        if (!b->pending.empty()) {
            b->wave.push_back(std::move(b->pending.front()));
            b->pending.pop_front();
        } else if (b->nextInOrder < b->order.size()) {
            Item it;
            it.src = b->order[b->nextInOrder].first;
            it.dst = b->order[b->nextInOrder].second;
            ++b->nextInOrder;
            auto pi = b->pairIndex.find(edgeKey(it.src, it.dst));
            if (pi != b->pairIndex.end()) {
                it.pairId = pi->second;
                it.nCorr = (uint32_t)(b->mOffset[it.pairId + 1] - b->mOffset[it.pairId]);
            }
            b->wave.push_back(std::move(it));
        } else
            break;
    }
    const uint32_t n = (uint32_t)b->wave.size();
    // 2. (re-)speculate where needed, in parallel on the current snapshot
    std::vector<uint32_t> todo;
    for (uint32_t i = 0; i < n; i++)
        if (!speculationHolds(b, b->wave[i])) todo.push_back(i);
    b->ctr.astar_runs += todo.size();
    const int T = std::max(1, std::min<int>(b->cfg.host_threads, (int)((todo.size() + 7) / 8)));
    if (T <= 1) {
        for (uint32_t i : todo) speculate(b, b->wave[i], b->scratch[0]);
    } else {
        std::atomic<size_t> next(0);
        auto worker = [&](int tid) {
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= todo.size()) break;
                speculate(b, b->wave[todo[k]], b->scratch[tid]);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < T; t++) pool.emplace_back(worker, t);
        worker(0);
        for (auto &th : pool) th.join();
    }
    // 3. emit
    for (uint32_t i = 0; i < n; i++) {
        Item &it = b->wave[i];
        it.needGpu = needsGpu(b, it);
        pgb_item &o = items[i];
        o.pair_id = it.pairId;
        o.src = it.src; o.dst = it.dst;
        o.has_hyp = it.hasHyp; o.need_gpu = it.needGpu; o.visible = it.visible; o.pad = 0;
        memcpy(o.hyp, it.hyp.q, 32);
        memcpy(o.hyp + 4, it.hyp.t, 24);
    }
    b->ctr.waves++;
    b->ctr.items_speculated += n;
    b->ctr.sec_astar += nowSec() - t0;
    return n;
}

uint32_t pgb_commit_wave(pgb_builder *b, const pgi_verdict *verdicts, uint32_t n_verdicts)
{
    if (!b) return 0;
    const double t0 = nowSec();
    const uint32_t n = (uint32_t)b->wave.size();
    // absorb the engine's verdicts into the caches (pure functions of (pair, hypothesis) / of the pair)
    uint32_t vi = 0;
    for (uint32_t i = 0; i < n; i++) {
        Item &it = b->wave[i];
        if (!it.needGpu) continue;
        if (vi >= n_verdicts || !verdicts) break;
        const pgi_verdict &v = verdicts[vi++];
        if (it.hasHyp) {
            PathVerdict pv;
            pv.ok = v.branch == 1;
            pv.v = v;
            b->pathCache[makeKey(it.pairId, it.hyp)] = pv;
        }
        if (v.status & 1u) {  // the fallback ran: its verdict is a function of the pair alone
            b->fbCache[it.pairId] = v;
            b->fbHave[it.pairId] = 1;
        }
    }
    uint32_t done = 0;
    AStarScratch &S = b->scratch[0];
    for (; done < n; ++done) {
        Item &it = b->wave[done];
        pgb_log lg;
        memset(&lg, 0, sizeof lg);
        lg.src = it.src; lg.dst = it.dst;
        lg.pair_index = it.pairId == UINT32_MAX ? -1 : (int64_t)it.pairId;
        lg.n_corr = it.nCorr;
        if (b->graph.hasEdge(it.src, it.dst) || b->graph.hasEdge(it.dst, it.src)) {  // pose_graph_builder.h:438-443
            b->ctr.skipped++;
            b->ctr.pairs_popped++;
            b->log.push_back(lg);
            continue;
        }
        if (!speculationHolds(b, it)) {  // exact sequential state: re-run (this IS the reference's search)
            speculate(b, it, S);
            b->ctr.astar_reruns++;
        }
        lg.visible = it.visible;
        if (it.pairId == UINT32_MAX || it.nCorr < b->cfg.minimum_point_number) {  // :550-551
            b->ctr.skipped++;
            b->ctr.pairs_popped++;
            b->log.push_back(lg);
            continue;
        }
        // resolve the verdict of (pair, exact hypothesis)
        const pgi_verdict *final = nullptr;
        const pgi_verdict *pathV = nullptr;
        if (it.hasHyp) {
            auto pc = b->pathCache.find(makeKey(it.pairId, it.hyp));
            if (pc == b->pathCache.end()) break;  // unknown: re-queue from here
            pathV = &pc->second.v;
            if (pc->second.ok) final = pathV;
        }
        if (!final) {
            if (!b->fbHave[it.pairId]) break;
            final = &b->fbCache[it.pairId];
        }
        lg.had_path = it.hasHyp;
        lg.touched_nodes = it.touched;
        if (pathV) { lg.test_passed = pathV->test_passed; lg.test_count = pathV->test_count; }
        lg.branch = final->accepted ? final->branch : 0;
        lg.inlier_number = final->inlier_count;
        memcpy(lg.E, final->E, sizeof lg.E);
        b->ctr.pairs_popped++;
        if (!final->accepted) {  // :641-642
            b->ctr.rejected++;
            b->log.push_back(lg);
            continue;
        }
        Edge e;
        e.src = it.src; e.dst = it.dst;
        memcpy(e.T.q, final->q, 32);
        memcpy(e.T.t, final->t, 24);
        e.score = (double)final->inlier_count / (double)it.nCorr;  // :645-646
        e.inlierNumber = final->inlier_count; e.nCorr = it.nCorr; e.branch = final->branch;
        const uint32_t ei = (uint32_t)b->graph.edges.size();
        b->graph.edges.push_back(e);
        b->graph.lookup[edgeKey(e.src, e.dst)] = ei;
        b->graph.byVertex[e.src].push_back(ei);  // pose_graph.h:219-220
        b->graph.byVertex[e.dst].push_back(ei);
        ++b->commitStamp;
        b->vertexStamp[e.src] = b->commitStamp;
        b->vertexStamp[e.dst] = b->commitStamp;
        const double tv = nowSec();
        b->vis.addLink(e.src, e.dst);  // :692
        b->ctr.sec_visibility += nowSec() - tv;
        lg.committed = 1;
        memcpy(lg.q, final->q, 32);
        memcpy(lg.t, final->t, 24);
        lg.score = e.score;
        b->log.push_back(lg);
        b->ctr.committed++;
        if (final->branch == 1) b->ctr.path_accepted++; else b->ctr.fallback_accepted++;
    }
    // re-queue the tail in order
    for (uint32_t i = n; i > done; --i) b->pending.push_front(std::move(b->wave[i - 1]));
    b->ctr.items_requeued += n - done;
    b->wave.clear();
    b->ctr.sec_commit += nowSec() - t0;
    return done;
}

uint64_t pgb_edge_count(pgb_builder *b) { return b ? b->graph.edges.size() : 0; }
void pgb_copy_edges(pgb_builder *b, pgb_edge *out)
{
    for (size_t i = 0; i < b->graph.edges.size(); i++) {
        const Edge &e = b->graph.edges[i];
        pgb_edge &o = out[i];
        memset(&o, 0, sizeof o);
        o.src = e.src; o.dst = e.dst;
        memcpy(o.q, e.T.q, 32);
        memcpy(o.t, e.T.t, 24);
        o.score = e.score; o.inlier_number = e.inlierNumber; o.n_corr = e.nCorr; o.branch = e.branch;
    }
}
uint64_t pgb_log_count(pgb_builder *b) { return b ? b->log.size() : 0; }
void pgb_copy_log(pgb_builder *b, pgb_log *out) { memcpy(out, b->log.data(), b->log.size() * sizeof(pgb_log)); }
void pgb_get_counters(pgb_builder *b, pgb_counters *out) { *out = b->ctr; }

int32_t pgb_astar(pgb_builder *b, uint32_t src, uint32_t dst, double *hyp_q_t, uint32_t *touched_nodes)
{
    if (!b || src >= b->V || dst >= b->V) return -1;
    AStarOut o;
    aStar(b->graph, b->sim.data(), b->V, src, dst, (size_t)b->cfg.maximum_search_depth, b->cfg.traversal_heuristics_weight,
          b->scratch[0], o);
    if (hyp_q_t) { memcpy(hyp_q_t, o.pose.q, 32); memcpy(hyp_q_t + 4, o.pose.t, 24); }
    if (touched_nodes) *touched_nodes = o.touched;
    return o.found ? 1 : 0;
}

}  // extern "C"
