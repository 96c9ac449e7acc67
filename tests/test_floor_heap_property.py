"""The lemma behind the host A*'s cost floor (pgb_host.cpp: aStar), checked directly on random operation sequences:
take libstdc++'s binary heap (std::push_heap / std::pop_heap, which is what std::priority_queue runs) and a second copy
in which entries below a floor are appended WITHOUT std::push_heap's climb.  As long as no entry below the floor reaches
the top, every entry at or above the floor must sit in the same slot in both arrays after every operation, and the popped
sequences must be identical — with plenty of equal keys, where the pop order depends on the layout."""
import os
import subprocess

SRC = r"""
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>
struct It { double f; uint32_t id; };
static bool cmp(const It &a, const It &b) { return a.f < b.f; }
int main() {
    std::mt19937_64 g(12345);
    long checked = 0, pops = 0, trials = 0;
    for (int trial = 0; trial < 400; trial++) {
        const int levels = 3 + (int)(g() % 40);           // few distinct keys: ties everywhere
        const double floorF = (double)(g() % levels) / levels;
        std::vector<It> A, B;                              // A: fully ordered; B: entries below the floor never climb
        uint32_t next = 0;
        bool valid = true;
        for (int op = 0; op < 3000 && valid; op++) {
            const bool push = A.empty() || (g() % 100) < 70;
            if (push) {
                const int batch = 1 + (int)(g() % 12);     // an expansion pushes several children in a row
                for (int k = 0; k < batch; k++) {
                    // biased towards low keys, as the children of a dense graph are
                    const double f = (double)(std::min(g() % levels, g() % levels)) / levels;
                    It it{f, next++};
                    A.push_back(it); std::push_heap(A.begin(), A.end(), cmp);
                    B.push_back(it);
                    if (f >= floorF) std::push_heap(B.begin(), B.end(), cmp);  // (climb of the last element only)
                }
            } else {
                if (B.front().f < floorF) { valid = false; break; }           // the floor was a bad guess: the caller retries
                if (A.front().id != B.front().id) { printf("pop mismatch trial %d op %d\n", trial, op); return 1; }
                std::pop_heap(A.begin(), A.end(), cmp); A.pop_back();
                std::pop_heap(B.begin(), B.end(), cmp); B.pop_back();
                pops++;
            }
            for (size_t s = 0; s < A.size(); s++) {
                const bool la = A[s].f >= floorF, lb = B[s].f >= floorF;
                if (la != lb || (la && A[s].id != B[s].id)) { printf("slot mismatch trial %d op %d slot %zu\n", trial, op, s); return 1; }
                checked += la;
            }
        }
        trials += valid;
    }
    printf("ok %ld %ld %ld\n", checked, pops, trials);
    return 0;
}
"""


def test_entries_at_or_above_the_floor_keep_their_slots(tmp_path):
    src = tmp_path / "h.cpp"
    src.write_text(SRC)
    exe = tmp_path / "h"
    subprocess.run(["g++", "-O2", "-o", str(exe), str(src)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    tag, checked, pops, trials = r.stdout.split()
    assert tag == "ok" and int(checked) > 1_000_000 and int(pops) > 50_000 and int(trials) > 20
