"""The fine staleness rule of the wave host (pgb_host.cpp: scoreChangeIsInert), checked directly: build a dense graph,
run searches, change the score of random edges at vertices the search expanded, and whenever the rule calls every such
change inert the search repeated on the changed graph must return exactly what it returned before — same path pose, same
pop and push counts, same expansions with the same cost tuples.  (And the rule must not be vacuous: it has to spare a
good share of the changes, and changes it does not spare do alter some searches.)"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r"""
#include "%s/pose_graph_initialization_b200/csrc/pgb_host.cpp"
#include <cstdio>
#include <random>
static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
static bool same(const AStarOut &a, const AStarOut &b)
{
    return a.found == b.found && a.touched == b.touched && a.pushes == b.pushes && a.expanded == b.expanded && a.expCost == b.expCost &&
           a.minPoppedF == b.minPoppedF && (!a.found || !memcmp(&a.pose, &b.pose, sizeof(SE3)));
}
int main()
{
    const uint32_t V = 160;
    std::mt19937_64 g(99);
    std::vector<double> sim((size_t)V * V);
    for (uint32_t i = 0; i < V; i++)
        for (uint32_t j = i; j < V; j++) {
            const double d = std::min((j - i) %% V, (i + V - j) %% V) / (double)V;   // ring cameras: similarity falls with distance
            const double s = i == j ? 1.0 : std::floor(std::max(0.0, std::cos(d * 6.283185307179586)) * 1000.0) / 1000.0;
            sim[(size_t)i * V + j] = sim[(size_t)j * V + i] = std::min(s, i == j ? 1.0 : 0.999);
        }
    std::vector<uint32_t> pv;
    for (uint32_t i = 0; i < V; i++)
        for (uint32_t j = i + 1; j < V; j++) { pv.push_back(i); pv.push_back(j); }
    const uint64_t P = pv.size() / 2;
    std::vector<uint64_t> mo(P + 1);
    for (uint64_t p = 0; p <= P; p++) mo[p] = p * 2000;
    pgb_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.minimum_inlier_number = 20; cfg.minimum_point_number = 50; cfg.maximum_search_depth = 5;
    cfg.traversal_heuristics_weight = 0.8; cfg.use_path_finding = 1; cfg.host_threads = 1;
    pgb_builder *b = nullptr;
    if (pgb_create(&cfg, V, sim.data(), P, pv.data(), mo.data(), &b) != 0) return 2;
    const size_t M = b->order.size() * 6 / 10;   // 60 %% of the queue committed: a dense graph
    for (size_t p = 0; p < M; p++) {
        Edge e;
        e.src = b->order[p].first; e.dst = b->order[p].second;
        const uint64_t h = mix(p * 0x9e3779b97f4a7c15ULL + 7);
        const bool path = h %% 100 < 29;
        const uint32_t inl = path ? 120 + (uint32_t)((h >> 8) %% 380) : 580 + (uint32_t)((h >> 8) %% 120);  // the measured score mix
        e.T = se3Identity();
        e.T.t[0] = (double)(h %% 1000) * 1e-3;  // (poses only matter to recoverPath; make them distinguishable)
        e.score = (double)inl / 2000.0;
        e.inlierNumber = inl; e.nCorr = 2000; e.branch = path ? 1 : 2;
        const uint32_t ei = (uint32_t)b->graph.edges.size();
        b->graph.edges.push_back(e);
        b->graph.lookup[edgeKey(e.src, e.dst)] = ei;
        b->graph.byVertex[e.src].push_back(Adj{e.dst, ei, e.score});
        b->graph.byVertex[e.dst].push_back(Adj{e.src, ei, e.score});
    }
    const double wgt = 0.8, omw = 1.0 - wgt;
    long inertChanges = 0, liveChanges = 0, inertTrials = 0, altered = 0, liveTrials = 0;
    AStarScratch S;
    for (size_t q = 0; q < 700; q++) {
        const uint32_t from = b->order[M + q].first, to = b->order[M + q].second;
        GraphView gv{&b->graph, nullptr, 0};
        AStarOut ref;
        aStar(gv, b->sim.data(), V, from, to, 5, wgt, S, ref);
        if (ref.expanded.empty()) continue;
        const double *simTo = b->sim.data() + (size_t)to * V;
        for (int trial = 0; trial < 8; trial++) {
            // change the scores of a few random edges at expanded vertices
            struct Ch { uint32_t ei; double sOld, sNew; };
            std::vector<Ch> chs;
            bool allInert = true;
            const int nChanges = trial %% 2 ? 3 : 1;
            for (int c = 0; c < nChanges; c++) {
                const uint32_t v = ref.expanded[g() %% ref.expanded.size()];
                const std::vector<Adj> &l = b->graph.byVertex[v];
                const Adj &a = l[g() %% l.size()];
                const uint32_t ei = a.edge;
                bool dup = false;
                for (const Ch &x : chs) dup |= x.ei == ei;
                if (dup) continue;
                const Edge &e = b->graph.edges[ei];
                // half of the changes lower the score (a path-branch edge replacing the predicted fallback edge: the common
                // case on the benchmark scenes), the others are arbitrary
                const uint32_t oldInl = (uint32_t)(e.score * 2000.0 + 0.5);
                const double sNew = (g() %% 2) ? (double)(40 + g() %% (oldInl > 41 ? oldInl - 40 : 1)) / 2000.0 : (double)(60 + g() %% 1300) / 2000.0;
                chs.push_back(Ch{ei, e.score, sNew});
                // the rule, for every expansion of either endpoint of the edge
                for (size_t x = 0; x < ref.expanded.size(); x++) {
                    const uint32_t u = ref.expanded[x];
                    if (u != e.src && u != e.dst) continue;
                    const uint32_t other = u == e.src ? e.dst : e.src;
                    const bool inert = scoreChangeIsInert(ref.expCost[2 * x], ref.expCost[2 * x + 1], simTo[other], e.score, sNew, wgt, omw, ref.minPoppedF);
                    allInert &= inert;
                    (inert ? inertChanges : liveChanges)++;
                }
            }
            auto apply = [&](bool forward) {
                for (const Ch &x : chs) {
                    Edge &e = b->graph.edges[x.ei];
                    e.score = forward ? x.sNew : x.sOld;
                    for (uint32_t v : {e.src, e.dst})
                        for (Adj &a : b->graph.byVertex[v])
                            if (a.edge == x.ei) a.score = e.score;
                }
            };
            apply(true);
            AStarOut now;
            aStar(gv, b->sim.data(), V, from, to, 5, wgt, S, now);
            apply(false);
            if (allInert) {
                ++inertTrials;
                if (!same(ref, now)) { printf("rule violated at query %%zu trial %%d\n", q, trial); return 1; }
            } else {
                ++liveTrials;
                altered += !same(ref, now);
            }
        }
    }
    printf("ok %%ld %%ld %%ld %%ld %%ld\n", inertChanges, liveChanges, inertTrials, liveTrials, altered);
    return 0;
}
"""


def test_changes_the_rule_calls_inert_leave_the_search_bit_identical(tmp_path):
    src = tmp_path / "s.cpp"
    src.write_text(SRC % ROOT)
    exe = tmp_path / "s"
    inc = ["-I", os.path.join(ROOT, "include")]
    if os.path.isdir("/usr/local/cuda/include"):
        inc += ["-I", "/usr/local/cuda/include"]
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-w"] + inc + ["-o", str(exe), str(src), "-ldl"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    tag, inert_changes, live_changes, inert_trials, live_trials, altered = r.stdout.split()
    assert tag == "ok"
    assert int(inert_trials) > 400 and int(inert_changes) > 1500  # the rule spares a good share of the changes ...
    assert int(live_trials) > 100 and int(altered) > 10           # ... and the ones it does not spare do alter searches
