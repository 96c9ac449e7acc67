// ORACLE — TEST INFRASTRUCTURE ONLY (see pgo_geom.hpp header).
//
// pgo_host.hpp — literal, sequential (core_number = 1) restatement of the host loop that
// drives the hot path: SURVEY §3.2/§3.3, App. A.1-A.4, A.10-A.12.
//   SimilarityTable queue        imagesimilarity_graph.h:51-66, :108-171
//   PoseGraph                    pose_graph.h:136-224 (insertion-ordered per-vertex edge lists)
//   VisibilityTable              visibility_table.h:45-171 (bug-for-bug, :101 and :108)
//   AStarTraversal::getPath      graph_traversal.h:679-870 ; recoverPath :290-348
//   PoseGraphBuilder::findPath   pose_graph_builder.h:785-862
//   PoseGraphBuilder::processImages  pose_graph_builder.h:352-715 (HDF5/image I/O replaced by the
//                                    synthetic scene arrays; epipolar hashing off)
// Deliberately uses the same std:: containers/algorithms as the reference so tie-breaking
// (std::priority_queue over std::vector, std::set iteration) is inherited, not re-derived.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <queue>
#include <set>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

#include "pgo_estimate.hpp"
#include "pgo_geom.hpp"

namespace pgo {
namespace host {

typedef size_t ViewId;
typedef std::pair<ViewId, ViewId> EdgeId;
struct EdgeIdHash {  // types.h:16-22
    std::size_t operator()(const EdgeId &p) const { return std::hash<ViewId>()(p.first) ^ std::hash<ViewId>()(p.second); }
};

struct Scene {
    size_t V = 0;
    const double *focal = nullptr;    // V
    const double *size = nullptr;     // V x 2 (width, height)
    const double *sim = nullptr;      // V x V row-major
    const uint64_t *kpOffset = nullptr;  // V+1
    const float *kp = nullptr;        // sum K x 2 (pixel x,y; cv::KeyPoint.pt is float)
    size_t P = 0;
    const uint32_t *pairViews = nullptr;  // P x 2 (src, dst) as queued (i<j for a symmetric matrix)
    const uint64_t *mOffset = nullptr;    // P+1
    const uint32_t *matches = nullptr;    // sum N x 2 (srcIdx, dstIdx)
};

struct Config {  // examples/cpp_example.cpp:32-66 defaults in comments
    double similarityThreshold = 0.5;     // similarity_threshold
    double inlierOutlierThreshold = 0.4;  // inlier_outlier_threshold (px)
    size_t minimumInlierNumber = 20;      // minimum_inlier_number
    size_t minimumPointNumber = 50;       // minimum_point_number
    size_t maximumSearchDepth = 5;        // maximum_search_depth
    double traversalHeuristicsWeight = 0.8;  // traversal_heuristics_weight
    bool usePathFinding = true;           // use_path_finding
};

struct Edge {
    ViewId src, dst;
    SE3 T;         // T_dst_src
    double score;  // inlierNumber / matches.size()  pose_graph_builder.h:645-646
};

class PoseGraph {  // pose_graph.h:62-224
public:
    bool hasEdge(ViewId s, ViewId d) const { return edges.find(EdgeId(s, d)) != edges.end(); }
    bool addEdge(ViewId s, ViewId d, const Edge &e)
    {
        if (hasEdge(s, d)) return false;
        const EdgeId id(s, d);
        edges.insert({id, e});
        edgeIds.push_back(id);
        edgesOfVertices[s].emplace_back(id);  // :219-220
        edgesOfVertices[d].emplace_back(id);
        return true;
    }
    bool getEdgesByVertex(ViewId v, std::vector<EdgeId> &out) const
    {
        auto it = edgesOfVertices.find(v);
        if (it == edgesOfVertices.end()) return false;  // :145-146 — `out` keeps its previous content
        out = it->second;
        return true;
    }
    const Edge &getEdgeById(const EdgeId &id) const { return edges.find(id)->second; }
    size_t numEdges() const { return edgeIds.size(); }
    std::map<ViewId, std::vector<EdgeId>> edgesOfVertices;
    std::unordered_map<EdgeId, Edge, EdgeIdHash> edges;
    std::vector<EdgeId> edgeIds;
};

class VisibilityTable {  // visibility_table.h:12-172
public:
    bool addLink(ViewId from_, ViewId to_)
    {
        if (from_ == to_) return false;
        const ViewId from = std::min(from_, to_), to = std::max(from_, to_);
        if (neighbors[from].find(to) == neighbors[from].end()) neighbors[from].insert(to);
        if (neighbors[to].find(from) == neighbors[to].end()) neighbors[to].insert(from);
        if (hasLinkOrdered(from, to)) return false;
        visibilityMap[std::make_pair(from, to)] = true;
        std::queue<ViewId> views;
        for (const ViewId &v : neighbors[from]) views.emplace(v);
        for (const ViewId &v : neighbors[to]) views.emplace(v);
        while (!views.empty()) {
            const ViewId view_id = views.front();
            views.pop();
            if (view_id == from || view_id == to) continue;
            const ViewId first = std::min(view_id, from), second = std::max(view_id, from);
            const ViewId third = std::min(view_id, to);
            const auto pair1 = std::make_pair(first, second);
            const auto pair2 = std::make_pair(third, third);  // :101 (sic)
            const bool hasPair1 = visibilityMap.find(pair1) != visibilityMap.end();
            if (!hasPair1 || !hasPair1) {  // :108 (sic)
                visibilityMap[pair1] = true;
                visibilityMap[pair2] = true;
                // :113 — the range expression uses the OUTER view_id
                std::vector<ViewId> nb(neighbors[view_id].begin(), neighbors[view_id].end());
                for (const ViewId &v : nb) views.emplace(v);
            }
            neighbors[view_id].insert(from);
            neighbors[view_id].insert(to);
        }
        return true;
    }
    bool hasLink(ViewId from_, ViewId to_) const
    {
        if (from_ == to_) return false;
        return hasLinkOrdered(std::min(from_, to_), std::max(from_, to_));
    }
    size_t size() const { return visibilityMap.size(); }

private:
    bool hasLinkOrdered(ViewId a, ViewId b) const { return visibilityMap.find(std::make_pair(a, b)) != visibilityMap.end(); }
    std::map<std::pair<ViewId, ViewId>, bool> visibilityMap;
    std::unordered_map<ViewId, std::set<ViewId>> neighbors;
};

struct AStarResult {
    bool pathFound = false;    // a path reached the destination and was composed (foundPaths_ == 1)
    SE3 pose = se3Identity();  // the composed hypothesis
    std::vector<ViewId> path;
    size_t touchedNodes = 0;
    size_t depth = 0;
};

// PoseGraphTraversal::recoverPath  graph_traversal.h:290-348
inline bool recoverPath(const PoseGraph &g, const std::vector<ViewId> &path, SE3 &pose)
{
    pose = se3Identity();  // :304
    for (size_t i = 1; i < path.size(); ++i) {
        const ViewId s = path[i - 1], d = path[i];
        if (g.hasEdge(s, d))
            pose = se3Mul(g.getEdgeById(EdgeId(s, d)).T, pose);  // :344
        else if (g.hasEdge(d, s))
            pose = se3Mul(se3Inverse(g.getEdgeById(EdgeId(d, s)).T), pose);  // :342
        else
            return false;
    }
    return true;
}

// AStarTraversal<ImageSimilarityHeuristics>::getPath  graph_traversal.h:679-870 with the arguments of
// pose_graph_builder.h:834-841: returnMultiple=true, minInlierRatio=(bool)0.0, maxPaths=1.
inline AStarResult aStar(const PoseGraph &g, const Scene &sc, ViewId from, ViewId to, size_t maxDepth, double weight)
{
    typedef std::tuple<double, double, double> Cost;
    typedef std::tuple<ViewId, std::vector<ViewId>, size_t> Node;
    typedef std::pair<Cost, Node> Item;
    struct Cmp {
        bool operator()(const Item &a, const Item &b) const { return std::get<2>(a.first) < std::get<2>(b.first); }  // :669-673
    };
    AStarResult res;
    std::priority_queue<Item, std::vector<Item>, Cmp> open;
    std::set<ViewId> states;  // nodeStates: only membership is ever queried (:855-856)
    open.push(std::make_pair(std::make_tuple(1.0, 0.0, 0.0), std::make_tuple(from, std::vector<ViewId>(), (size_t)0)));  // :721
    const double oneMinusWeight = 1.0 - weight;
    const double minimumInlierRatio = 0.0;  // `const bool kMinimumInlierRatio_` receiving 0.0 (:613, :839)
    std::vector<EdgeId> edges;
    while (!open.empty()) {
        Node node = open.top().second;
        const Cost cost = open.top().first;
        const size_t depth = std::get<2>(node);
        ++res.touchedNodes;  // :750
        open.pop();
        if (depth > maxDepth) continue;  // :755
        const ViewId v = std::get<0>(node);
        std::vector<ViewId> parents = std::get<1>(node);
        if (v == to) {  // :766
            std::vector<ViewId> path = parents;
            path.push_back(v);
            SE3 pose;
            if (recoverPath(g, path, pose)) {
                res.pathFound = true;
                res.pose = pose;
                res.path = path;
                res.depth = depth;
                break;  // :792-800 — one path is tested, then the search stops whatever the verdict
            }
            continue;  // :803-806 with kReturnMultiplePaths = true
        }
        parents.emplace_back(v);
        states.insert(v);  // :814
        g.getEdgesByVertex(v, edges);  // :817
        if (depth < maxDepth) {  // :820
            for (const EdgeId &eid : edges) {
                const Edge &e = g.getEdgeById(eid);
                if (e.score < minimumInlierRatio) continue;  // :830
                const ViewId next = (v == e.dst) ? e.src : e.dst;  // :838-840
                const double edgeCost = std::get<0>(cost) > e.score ? e.score : std::get<0>(cost);  // MIN :843
                double h = sc.sim[next * sc.V + to];  // getSimilarity imagesimilarity_graph.h:95-106
                h = std::clamp(h, 0.0, 1.0);          // :594
                const double nextToDest = std::get<1>(cost) < h ? h : std::get<1>(cost);  // MAX :847
                const double combined = weight * edgeCost + oneMinusWeight * nextToDest;  // :851-852
                if (states.find(next) == states.end())  // :855-856
                    open.push(std::make_pair(std::make_tuple(edgeCost, nextToDest, combined),
                                             std::make_tuple(next, parents, depth + 1)));
            }
        }
    }
    return res;
}

// SimilarityTable::loadFromFile queue construction  imagesimilarity_graph.h:141-163 on an in-memory
// matrix: table pre-filled with 1.0f (:58-65), pair (i,j) queued while filling row i.
inline std::priority_queue<std::tuple<double, ViewId, ViewId>> buildPairQueue(const Scene &sc, double threshold)
{
    std::priority_queue<std::tuple<double, ViewId, ViewId>> q;
    const size_t V = sc.V;
    std::vector<double> table(V * V, (double)1.0f);
    for (size_t i = 0; i < V; i++)
        for (size_t j = 0; j < V; j++) {
            table[i * V + j] = sc.sim[i * V + j];
            if (i != j && threshold <= table[i * V + j] && table[j * V + i] != table[i * V + j])
                q.emplace(std::make_tuple(table[i * V + j], i, j));
        }
    return q;
}

struct PairLog {
    uint32_t src, dst;
    int64_t pairIndex;    // index into the scene's pair list, -1 if the queued pair has no matches
    uint8_t visible;      // visibilityTable.hasLink
    uint8_t hadPath;      // A* composed a hypothesis
    uint8_t testPassed;   // InTraversalPoseTester::test verdict
    uint8_t branch;       // 0 rejected/skipped, 1 path, 2 fallback
    uint8_t committed;
    uint32_t testCount;
    uint32_t inlierNumber;
    uint32_t nCorr;
    uint32_t touchedNodes;
    double E[9];
    double q[4];
    double t[3];
    double score;
};

struct RunResult {
    std::vector<PairLog> log;      // one record per popped pair, in processing order
    std::vector<Edge> edges;       // committed edges in commit order
    size_t fallbackRuns = 0, pathAccepted = 0, fallbackAccepted = 0, rejected = 0, skipped = 0;
    size_t corrEvals = 0;          // hypothesis x correspondence Sampson evaluations (test + getInliers, as executed)
};

// processImages, sequential  pose_graph_builder.h:352-715.  maxPairs = 0 -> whole queue.
inline RunResult run(const Scene &sc, const Config &cfg, size_t maxPairs = 0)
{
    RunResult out;
    PoseGraph graph;
    VisibilityTable vis;
    auto queue = buildPairQueue(sc, cfg.similarityThreshold);
    std::map<std::pair<uint32_t, uint32_t>, size_t> pairIndex;
    for (size_t p = 0; p < sc.P; p++) pairIndex[{sc.pairViews[2 * p], sc.pairViews[2 * p + 1]}] = p;

    std::vector<double> corr;
    std::vector<uint8_t> mask;
    size_t processed = 0;
    while (!queue.empty()) {
        if (maxPairs && processed >= maxPairs) break;
        const auto vp = queue.top();
        queue.pop();
        ++processed;
        const ViewId src = std::get<1>(vp), dst = std::get<2>(vp);
        PairLog lg{};
        lg.src = (uint32_t)src; lg.dst = (uint32_t)dst; lg.pairIndex = -1;
        if (graph.hasEdge(src, dst) || graph.hasEdge(dst, src)) {  // :438-443
            out.skipped++;
            out.log.push_back(lg);
            continue;
        }
        const bool visible = vis.hasLink(src, dst);  // :456-457
        lg.visible = visible;
        auto pit = pairIndex.find({(uint32_t)src, (uint32_t)dst});
        size_t n = 0;
        size_t p = 0;
        if (pit != pairIndex.end()) {
            p = pit->second;
            lg.pairIndex = (int64_t)p;
            n = (size_t)(sc.mOffset[p + 1] - sc.mOffset[p]);
        }
        lg.nCorr = (uint32_t)n;
        if (n < cfg.minimumPointNumber) {  // :550-551
            out.skipped++;
            out.log.push_back(lg);
            continue;
        }
        // createCorrespondenceMatrix :553-565 (source intrinsics for both images, §0.8)
        corr.resize(n * 4);
        double thrNorm = 0.0;
        const double f = sc.focal[src], cx = sc.size[2 * src] / 2.0, cy = sc.size[2 * src + 1] / 2.0;  // :284-286
        createCorrespondenceMatrix(sc.kp + 2 * sc.kpOffset[src], sc.kp + 2 * sc.kpOffset[dst],
                                   sc.matches + 2 * sc.mOffset[p], n, f, f, cx, cy, cfg.inlierOutlierThreshold,
                                   corr.data(), thrNorm);
        std::vector<SE3> poses;
        if (cfg.usePathFinding && visible) {  // :569-603
            AStarResult ar = aStar(graph, sc, src, dst, cfg.maximumSearchDepth, cfg.traversalHeuristicsWeight);
            lg.touchedNodes = (uint32_t)ar.touchedNodes;
            lg.hadPath = ar.pathFound;
            if (ar.pathFound) {
                size_t cnt = 0;
                const bool ok = inTraversalTest(corr.data(), n, ar.pose, 1.5 * thrNorm, 5, cnt);  // :798-811, GT:790
                lg.testPassed = ok;
                lg.testCount = (uint32_t)cnt;
                out.corrEvals += n;  // upper bound of the early-exit loop; the GPU path evaluates all n
                if (ok) poses.push_back(ar.pose);
            }
        }
        mask.clear();
        EstimateResult er = estimatePose(corr.data(), n, thrNorm, cfg.minimumInlierNumber, poses.data(), poses.size(), mask);
        if (!poses.empty()) out.corrEvals += n;
        if (er.fallbackIters > 0) out.fallbackRuns++;
        lg.branch = (uint8_t)er.branch;
        lg.inlierNumber = (uint32_t)er.inlierNumber;
        for (int k = 0; k < 9; k++) lg.E[k] = er.E[k];
        if (!er.success) {  // :641-642
            out.rejected++;
            out.log.push_back(lg);
            continue;
        }
        const double ratio = (double)er.inlierNumber / (double)n;  // :645-646
        Edge e{src, dst, er.pose, ratio};
        graph.addEdge(src, dst, e);  // :649-654
        vis.addLink(src, dst);       // :692
        out.edges.push_back(e);
        lg.committed = 1;
        lg.q[0] = er.pose.q.x; lg.q[1] = er.pose.q.y; lg.q[2] = er.pose.q.z; lg.q[3] = er.pose.q.w;
        lg.t[0] = er.pose.t.x; lg.t[1] = er.pose.t.y; lg.t[2] = er.pose.t.z;
        lg.score = ratio;
        if (er.branch == 1) out.pathAccepted++; else out.fallbackAccepted++;
        out.log.push_back(lg);
    }
    return out;
}

// ---- replay of a logged run (bench.py --verify, tests/test_gpu_fullsize.py) ------------------------------------
// The product logs, per queue position, what its speculative-wave host decided: visibility, the hypothesis its A*
// composed, the touched-node count, and the verdict the engine returned.  `replay` walks the queue with THIS file's
// sequential host — same PoseGraph / VisibilityTable / aStar as run() — but takes the per-pair verdict from the log
// instead of computing it (the verdicts themselves are checked separately, tuple by tuple, against estimatePose), and
// compares everything the host logic produces on the way: queue order, duplicate skips, hasLink, whether a path was
// found, the composed pose bit for bit, touched nodes.  A run whose log passes both checks committed exactly the graph
// the sequential reference semantics commit.
struct ReplayLog {
    size_t n = 0;
    const uint32_t *src = nullptr, *dst = nullptr;
    const int64_t *pairIndex = nullptr;
    const uint8_t *visible = nullptr, *hadPath = nullptr, *committed = nullptr;
    const uint32_t *touchedNodes = nullptr, *nCorr = nullptr;
    const double *hyp = nullptr;    // n x 7 (q xyzw, t): the hypothesis handed to the engine (valid if hadPath)
    const double *q = nullptr, *t = nullptr, *score = nullptr;  // committed edge (valid if committed)
};
struct ReplayResult {
    size_t checked = 0, mismatches = 0, edges = 0, searches = 0;
    int64_t firstBad = -1;
    int firstBadField = 0;  // 1 queue order, 2 skip, 3 visible, 4 hadPath, 5 touched, 6 hypothesis, 7 length
};

inline ReplayResult replay(const Scene &sc, const Config &cfg, const ReplayLog &lg, size_t maxPairs = 0)
{
    ReplayResult out;
    PoseGraph graph;
    VisibilityTable vis;
    auto queue = buildPairQueue(sc, cfg.similarityThreshold);
    auto bad = [&](size_t k, int field) {
        ++out.mismatches;
        if (out.firstBad < 0) { out.firstBad = (int64_t)k; out.firstBadField = field; }
    };
    size_t k = 0;
    while (!queue.empty()) {
        if (maxPairs && k >= maxPairs) break;
        if (k >= lg.n) { bad(k, 7); break; }
        const auto vp = queue.top();
        queue.pop();
        const ViewId src = std::get<1>(vp), dst = std::get<2>(vp);
        const size_t pos = k++;
        ++out.checked;
        if (lg.src[pos] != src || lg.dst[pos] != dst) { bad(pos, 1); break; }
        const bool dup = graph.hasEdge(src, dst) || graph.hasEdge(dst, src);  // :438-443
        if (dup) {
            if (lg.committed[pos]) bad(pos, 2);
            continue;
        }
        const bool visible = vis.hasLink(src, dst);  // :456-457
        if ((lg.visible[pos] != 0) != visible) bad(pos, 3);
        const size_t n = lg.pairIndex[pos] >= 0 ? lg.nCorr[pos] : 0;
        if (n < cfg.minimumPointNumber) {  // :550-551
            if (lg.committed[pos]) bad(pos, 2);
            continue;
        }
        if (cfg.usePathFinding && visible) {  // :569-603
            AStarResult ar = aStar(graph, sc, src, dst, cfg.maximumSearchDepth, cfg.traversalHeuristicsWeight);
            ++out.searches;
            if ((lg.hadPath[pos] != 0) != ar.pathFound) bad(pos, 4);
            if (lg.touchedNodes[pos] != (uint32_t)ar.touchedNodes) bad(pos, 5);
            if (ar.pathFound && lg.hadPath[pos]) {
                const double h[7] = {ar.pose.q.x, ar.pose.q.y, ar.pose.q.z, ar.pose.q.w, ar.pose.t.x, ar.pose.t.y, ar.pose.t.z};
                if (std::memcmp(h, lg.hyp + 7 * pos, sizeof h) != 0) bad(pos, 6);
            }
        } else if (lg.hadPath[pos])
            bad(pos, 4);
        if (!lg.committed[pos]) continue;  // :641-642
        Edge e{src, dst, SE3{Quat{lg.q[4 * pos], lg.q[4 * pos + 1], lg.q[4 * pos + 2], lg.q[4 * pos + 3]},
                             Vec3{lg.t[3 * pos], lg.t[3 * pos + 1], lg.t[3 * pos + 2]}},
               lg.score[pos]};
        graph.addEdge(src, dst, e);  // :649-654
        vis.addLink(src, dst);       // :692
        ++out.edges;
    }
    if (!maxPairs && k != lg.n) bad(k, 7);
    return out;
}

}  // namespace host
}  // namespace pgo
