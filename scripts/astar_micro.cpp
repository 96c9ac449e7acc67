// Host A* micro-benchmark: fast-forwards the committed graph to queue position M with synthetic edges (29 % "path"
// scores ~0.62, the rest "fallback" scores 0.35-0.45: the mix of the 40 %-outlier scenes), then searches the next K
// queue positions on that fixed graph with T threads and reports time, pops and pushes, plus a checksum of the results
// so that two builds can be compared.
//   g++ -O3 -ffp-contract=off -pthread -I include -o /tmp/astar_micro scripts/astar_micro.cpp
//   /tmp/astar_micro sim1000.bin 1000 300000 4096 8 [margin]     (margin >= 0: cost floor = last floor seen for the
//                                                                  source / destination view - margin; default: no floor)
#include "../pose_graph_initialization_b200/csrc/pgb_host.cpp"
#include <cstdio>
#include <cstdlib>

static uint64_t mix(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

int main(int argc, char **argv)
{
    if (argc < 6) return 1;
    const uint32_t V = (uint32_t)atoi(argv[2]);
    const size_t M = (size_t)atol(argv[3]), K = (size_t)atol(argv[4]);
    const int T = atoi(argv[5]);
    const double margin = argc > 6 ? atof(argv[6]) : -1.0;
    std::vector<double> srcFloor(V, -1.0), dstFloor(V, -1.0);
    std::atomic<uint64_t> retries(0);
    std::vector<double> sim((size_t)V * V);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(sim.data(), 8, sim.size(), f) != sim.size()) return 2;
    fclose(f);
    std::vector<uint32_t> pv;
    for (uint32_t i = 0; i < V; i++)
        for (uint32_t j = i + 1; j < V; j++) { pv.push_back(i); pv.push_back(j); }
    const uint64_t P = pv.size() / 2;
    std::vector<uint64_t> mo(P + 1);
    for (uint64_t p = 0; p <= P; p++) mo[p] = p * 2000;
    pgb_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.similarity_threshold = 0.0; cfg.minimum_inlier_number = 20; cfg.minimum_point_number = 50;
    cfg.maximum_search_depth = 5; cfg.traversal_heuristics_weight = 0.8; cfg.use_path_finding = 1; cfg.host_threads = T;
    pgb_builder *b = nullptr;
    if (pgb_create(&cfg, V, sim.data(), P, pv.data(), mo.data(), &b) != 0) return 3;
    const size_t Mx = std::min(M, b->order.size());
    for (size_t p = 0; p < Mx; p++) {
        Edge e;
        e.src = b->order[p].first; e.dst = b->order[p].second;
        const uint64_t h = mix(p * 0x9e3779b97f4a7c15ULL + 12345);
        const bool path = h % 100 < 29;
        const uint32_t inl = path ? 1200 + (uint32_t)((h >> 8) % 80) : 700 + (uint32_t)((h >> 8) % 200);
        e.T = se3Identity();
        e.score = (double)inl / 2000.0;
        e.inlierNumber = inl; e.nCorr = 2000; e.branch = path ? 1 : 2;
        const uint32_t ei = (uint32_t)b->graph.edges.size();
        b->graph.edges.push_back(e);
        b->graph.lookup[edgeKey(e.src, e.dst)] = ei;
        b->graph.byVertex[e.src].push_back(Adj{e.dst, ei, e.score});
        b->graph.byVertex[e.dst].push_back(Adj{e.src, ei, e.score});
    }
    const size_t Kx = std::min(K, b->order.size() - Mx);
    std::vector<uint64_t> pops(T, 0), pushes(T, 0), sum(T, 0);
    std::vector<AStarOut> outs(Kx);
    const double t0 = nowSec();
    std::atomic<size_t> next(0);
    b->pool.run([&](int tid) {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= Kx) break;
            GraphView gv{&b->graph, nullptr, 0};
            const uint32_t s = b->order[Mx + i].first, d = b->order[Mx + i].second;
            double floorF = -1.0;
            if (margin >= 0.0 && srcFloor[s] >= 0.0 && dstFloor[d] >= 0.0) floorF = std::min(srcFloor[s], dstFloor[d]) - margin;
            aStar(gv, b->sim.data(), V, s, d, 5, 0.8, b->scratch[tid], outs[i], floorF);
            if (!outs[i].valid) {
                ++retries;
                aStar(gv, b->sim.data(), V, s, d, 5, 0.8, b->scratch[tid], outs[i]);
            }
            if (outs[i].minPoppedF <= 1.0) srcFloor[s] = dstFloor[d] = outs[i].minPoppedF;
            pops[tid] += outs[i].touched;
            pushes[tid] += outs[i].pushes;
        }
    });
    const double dt = nowSec() - t0;
    uint64_t tp = 0, tq = 0, cs = 0;
    for (int t = 0; t < T; t++) { tp += pops[t]; tq += pushes[t]; }
    for (size_t i = 0; i < Kx; i++) {
        cs = mix(cs ^ outs[i].touched) ^ mix(outs[i].pushes + 77) ^ (outs[i].found ? 1 : 0);
        for (uint32_t v : outs[i].expanded) cs = mix(cs + v);
    }
    printf("M=%zu K=%zu T=%d  %.3f s  searches/s %.0f  pops %llu (%.0f/search)  pushes %llu (%.0f/search)  %.2f ns/push/thread  retries %llu  checksum %016llx\n",
           Mx, Kx, T, dt, Kx / dt, (unsigned long long)tp, (double)tp / Kx, (unsigned long long)tq, (double)tq / Kx,
           dt * T / (double)tq * 1e9, (unsigned long long)retries.load(), (unsigned long long)cs);
    return 0;
}
