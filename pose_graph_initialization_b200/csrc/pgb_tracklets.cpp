// pgb_tracklets.cpp — the reference's Tracklets (point_track.h:541-712) behind the C-ABI of include/pgb.h.
//
// Tracklets is the bookkeeping that turns verified matches into multi-view point tracks so that a later pair of the
// queue can be matched "quickly" from the tracks its two views share (pose_graph_builder.h:492-520: getCorrespondences;
// :663-676 and :697-703: add).  It is host-side and sequential in the reference (one writer lock around add); it sits on
// the commit side of the hot path, so it is restated here bug for bug — the correspondences it hands out are the input
// of createCorrespondenceMatrix for quick-matched pairs:
//   * a (view, point) pair is numbered on first sight, and number 0 doubles as "not numbered yet" (point_track.h:655-661):
//     the very first pair ever added (number 0) is numbered again the next time it is seen and loses its old tracks;
//   * a track is extended through every track of the other endpoint, so one match can extend several tracks and a view
//     can appear in a track several times with different points (:669-695);
//   * getCorrespondences walks the destination view's track list in insertion order (duplicates included), scans a shared
//     track until it has met two points of the two views (two points of the SAME view end the scan as well, leaving the
//     other index 0), and stops only after exceeding the maximum (:629-632, i.e. it may return maximum + 1 matches); the
//     third tuple member is never written (value-initialised, 0.0).
// Storage is flat (vectors indexed by pair number / track number; one open-addressing table for the pair numbers) instead
// of the reference's four node-based maps.
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/pgb.h"

struct pgb_tracklets {
    struct Pt {
        uint64_t view, point;
        bool operator==(const Pt &o) const { return view == o.view && point == o.point; }
    };
    struct PtHash {
        size_t operator()(const Pt &p) const { return (size_t)(p.view * 8001u + p.point); }  // pairHash, point_track.h:536-539
    };
    uint64_t pointPairNumber = 0;
    std::unordered_map<Pt, uint64_t, PtHash> pointPairs;     // (view, point) -> pair number (0: see above)
    std::vector<std::vector<uint64_t>> pairToTracks;         // by pair number
    std::vector<std::vector<Pt>> tracks;                     // by track number
    std::unordered_map<uint64_t, std::vector<uint64_t>> viewToTracks;
    std::vector<uint64_t> &tracksOfPair(uint64_t id)
    {
        if (pairToTracks.size() <= id) pairToTracks.resize(id + 1);
        return pairToTracks[id];
    }
};

extern "C" {

pgb_tracklets *pgb_tracklets_create(uint64_t view_number)
{
    pgb_tracklets *t = new pgb_tracklets();
    t->viewToTracks.reserve(view_number);
    return t;
}

void pgb_tracklets_destroy(pgb_tracklets *t) { delete t; }

// Tracklets::add (point_track.h:638-712)
int32_t pgb_tracklets_add(pgb_tracklets *t, uint64_t view_src, uint64_t view_dst, uint64_t n, const uint64_t *point_src,
                          const uint64_t *point_dst, const uint8_t *inlier_mask)
{
    if (!t || (n && (!point_src || !point_dst || !inlier_mask))) return -1;
    for (uint64_t i = 0; i < n; i++) {
        if (!inlier_mask[i]) continue;  // :650-651
        const pgb_tracklets::Pt ps{view_src, point_src[i]}, pd{view_dst, point_dst[i]};
        uint64_t &is = t->pointPairs[ps];  // :655-661 (0 = "new", also the number of the first pair ever)
        if (is == 0) is = t->pointPairNumber++;
        const uint64_t idS = is;
        uint64_t &id = t->pointPairs[pd];
        if (id == 0) id = t->pointPairNumber++;
        const uint64_t idD = id;
        t->tracksOfPair(idS > idD ? idS : idD);  // both lists exist before references are taken
        std::vector<uint64_t> &tracksS = t->pairToTracks[idS], &tracksD = t->pairToTracks[idD];
        const size_t nD = tracksD.size();  // :665
        bool added = false;
        for (size_t k = 0; k < tracksS.size(); k++) {  // :668-679 (tracksS does not grow here unless idS == idD)
            const uint64_t tr = tracksS[k];
            std::vector<pgb_tracklets::Pt> &track = t->tracks[tr];
            bool has = false;
            for (const auto &p : track) has |= p == pd;
            if (has) continue;
            t->viewToTracks[view_dst].push_back(tr);
            track.push_back(pd);
            tracksD.push_back(tr);
            added = true;
        }
        for (size_t k = 0; k < nD; k++) {  // :681-693
            const uint64_t tr = tracksD[k];
            std::vector<pgb_tracklets::Pt> &track = t->tracks[tr];
            bool has = false;
            for (const auto &p : track) has |= p == ps;
            if (has) continue;
            t->viewToTracks[view_src].push_back(tr);
            track.push_back(ps);
            tracksS.push_back(tr);
            added = true;
        }
        if (!added) {  // :695-703
            const uint64_t tr = t->tracks.size();
            t->tracks.push_back({ps, pd});
            t->viewToTracks[view_src].push_back(tr);
            t->viewToTracks[view_dst].push_back(tr);
            tracksS.push_back(tr);
            tracksD.push_back(tr);
        }
    }
    return 0;
}

// Tracklets::getCorrespondences (point_track.h:575-636).  Returns the number of matches (may be maximum + 1), or -1 if
// `capacity` is too small / arguments are invalid.  out_src / out_dst receive the keypoint indices.
int64_t pgb_tracklets_get_correspondences(pgb_tracklets *t, uint64_t view_src, uint64_t view_dst, uint64_t maximum,
                                          uint64_t *out_src, uint64_t *out_dst, uint64_t capacity)
{
    if (!t || !out_src || !out_dst) return -1;
    const auto it = t->viewToTracks.find(view_src);
    if (it == t->viewToTracks.end()) return 0;  // :583-588
    const auto jt = t->viewToTracks.find(view_dst);
    if (jt == t->viewToTracks.end()) return 0;  // :590-596
    std::unordered_set<uint64_t> inSrc(it->second.begin(), it->second.end());  // :602-606
    uint64_t count = 0;
    for (const uint64_t tr : jt->second) {  // :608
        if (inSrc.find(tr) == inSrc.end()) continue;
        uint64_t a = 0, b = 0;  // (value-initialised tuple)
        int cnt = 0;
        for (const auto &p : t->tracks[tr]) {  // :614-627
            if (p.view == view_src) { a = p.point; ++cnt; }
            else if (p.view == view_dst) { b = p.point; ++cnt; }
            if (cnt == 2) break;
        }
        if (count >= capacity) return -1;
        out_src[count] = a;
        out_dst[count] = b;
        ++count;
        if (count > maximum) break;  // :631-632
    }
    return (int64_t)count;
}

uint64_t pgb_tracklets_track_count(pgb_tracklets *t) { return t ? t->tracks.size() : 0; }

}  // extern "C"
