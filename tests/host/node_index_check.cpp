// Host-logic check (CPU only): the product's node index (trg-planner_b200/host/node_index.h)
// against the oracle's kd-tree restatement (oracle/kdtree_port.h, itself pinned to the reference's
// kdtree.c by tests/test_oracle_cpu.py): nearest incl. exact ties, range sets AND result order,
// bulk build == sequential insertion.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "kdtree_port.h"
#include "node_index.h"

static int fails = 0;
static long n_ties = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } ++fails; } } while (0)

// ChunkTable::refine (nodes created since the last hand-over) against a brute-force merge: same
// float arithmetic as the table (dx*dx then += dy*dy), candidates at any distance (pruned probe,
// ring fallback, exhaustive fallback), exact ties on a lattice, prune() and clear().
static void check_chunk_table(std::mt19937& gen) {
  for (int round = 0; round < 6; ++round) {
    const bool lattice = round >= 4;
    const float ext = round % 2 ? 300.f : 30.f;
    const float cell = round % 3 == 0 ? 0.3f : (round % 3 == 1 ? 0.6f : 0.9f);
    const int n = 6000;
    std::uniform_real_distribution<float> U(0.f, ext);
    std::vector<float> xs(n), ys(n);
    for (int i = 0; i < n; ++i) { xs[i] = U(gen); ys[i] = U(gen); }
    if (lattice) for (int i = 0; i < n; ++i) { xs[i] = 3.f + (float)(i % 80) * 0.25f; ys[i] = 5.f + (float)((i / 80) % 75) * 0.25f; }
    const int handed = n / 2;  // seq < handed: "on the device"; the rest is in the table
    trg_b200::ChunkTable tab;
    tab.configure(-1.f, -2.f, cell);
    for (int i = handed; i < n; ++i) tab.insert(xs[i], ys[i], i);
    auto d2f = [&](int i, float qx, float qy) { const float dx = xs[i] - qx, dy = ys[i] - qy; float v = 0.f; v += dx * dx; v += dy * dy; return v; };
    for (int pass = 0; pass < 2; ++pass) {
      int lo = handed;
      if (pass == 1) { lo = handed + n / 4; tab.prune(lo); }  // forget the older half of the table
      for (int q = 0; q < 4000; ++q) {
        float qx = U(gen), qy = U(gen);
        if (lattice && q % 2) { qx = 3.f + 0.125f + 0.25f * (float)(q % 79); qy = 5.f + 0.125f + 0.25f * (float)((q / 79) % 74); }  // cell centres: 4-way ties
        // device candidate: brute force over seq < handed (every third query: none at all)
        float d2 = std::numeric_limits<float>::infinity(); int seq = -1; bool tie = false;
        if (q % 3) for (int i = 0; i < handed; ++i) { const float v = d2f(i, qx, qy); if (v < d2) { d2 = v; seq = i; tie = false; } else if (v == d2) tie = true; }
        float wd = d2; int ws = seq; bool wt = tie;
        for (int i = lo; i < n; ++i) { const float v = d2f(i, qx, qy); if (v < wd) { wd = v; ws = i; wt = false; } else if (v == wd && i != ws) wt = true; }
        tab.refine(qx, qy, d2, seq, tie);
        CHECK(d2 == wd, "chunk table d2 %g want %g (round %d pass %d q %d)", d2, wd, round, pass, q);
        CHECK(tie == wt, "chunk table tie %d want %d (round %d pass %d q %d)", (int)tie, (int)wt, round, pass, q);
        if (!wt) CHECK(seq == ws, "chunk table seq %d want %d (round %d pass %d q %d)", seq, ws, round, pass, q);
        if (wt) ++n_ties;
      }
    }
    tab.clear();
    float d2 = 4.f; int seq = 7; bool tie = false;
    tab.refine(1.f, 1.f, d2, seq, tie);
    CHECK(d2 == 4.f && seq == 7 && !tie && tab.empty(), "cleared table must not change the candidate");
  }
}

int main() {
  std::mt19937 gen(7);
  check_chunk_table(gen);
  for (int round = 0; round < 6; ++round) {
    const int n = round == 0 ? 1 : (round == 1 ? 17 : 20000);
    const float ext = round < 4 ? 40.f : 400.f;
    std::uniform_real_distribution<float> U(0.f, ext);
    std::vector<float> xs(n), ys(n);
    for (int i = 0; i < n; ++i) { xs[i] = U(gen); ys[i] = U(gen); }
    if (round == 3)  // lattice with exact duplicates / exact ties
      for (int i = 0; i < n; ++i) { xs[i] = (float)(i % 100) * 0.25f; ys[i] = (float)((i / 100) % 100) * 0.25f; }
    kdport::Tree2 ref;
    trg_b200::OrderTree2D seq, bulk;
    trg_b200::NodeGrid grid;
    grid.configure(0.f, 0.f, ext, ext, 0.6f);
    for (int i = 0; i < n; ++i) {
      ref.insert(xs[i], ys[i], i);
      seq.insert(xs[i], ys[i], i);
      grid.insert(xs[i], ys[i]);
    }
    bulk.build_bulk(xs.data(), ys.data(), n);
    CHECK(bulk.low() == seq.low() && bulk.high() == seq.high() && bulk.axis() == seq.axis(), "bulk != sequential (n=%d)", n);
    // a tree grown elsewhere and adopted (the device build hands over lo / hi / parent / axis of k_kd_build)
    trg_b200::OrderTree2D adopted;
    std::vector<float> xy2(2 * (size_t)n);
    for (int i = 0; i < n; ++i) { xy2[2 * i] = xs[i]; xy2[2 * i + 1] = ys[i]; }
    {
      std::vector<int> lo = seq.low(), hi = seq.high(), par = seq.parent();
      std::vector<uint8_t> ax = seq.axis();
      adopted.adopt(xy2.data(), n, std::move(lo), std::move(hi), std::move(par), std::move(ax));
    }
    // the hash grid filled in one go by banded helper threads (after a build), and re-used after clear()
    trg_b200::NodeGrid grid2;
    grid2.configure(0.f, 0.f, ext, ext, 0.6f);
    for (int i = 0; i < std::min(n, 50); ++i) grid2.insert(xs[n - 1 - i], ys[n - 1 - i]);  // stale content
    grid2.clear();
    CHECK(grid2.nearest(1.f, 1.f).entry == -1 && grid2.count_in_range(1.f, 1.f, 5.f) == 0, "cleared grid still answers");
    grid2.rebuild(xy2.data(), n, 5);
    trg_b200::NodeGrid grid3;
    grid3.configure(0.f, 0.f, ext, ext, 0.6f);
    for (int i = 0; i < 7 && i < n; ++i) grid3.insert(ys[i], xs[i]);
    grid3.clear();
    for (int i = 0; i < n; ++i) grid3.insert(xs[i], ys[i]);   // the lazy clear happens on this first insert
    std::vector<int64_t> want;
    std::vector<int> got, got2;
    std::uniform_real_distribution<float> Q(-1.f, ext + 1.f);
    for (int k = 0; k < 4000; ++k) {
      float qx = Q(gen), qy = Q(gen);
      if (k % 5 == 0) { int j = gen() % n; qx = xs[j]; qy = ys[j]; }
      if (round == 3 && k % 3 == 0) { qx = 0.125f + 0.25f * (float)(gen() % 99); qy = 0.125f + 0.25f * (float)(gen() % 99); }  // 4-way ties
      const int64_t wn = ref.nearest(qx, qy);
      CHECK(seq.nearest(qx, qy) == (int)wn, "tree nearest");
      CHECK(bulk.nearest(qx, qy) == (int)wn, "bulk tree nearest");
      CHECK(adopted.nearest(qx, qy) == (int)wn, "adopted tree nearest");
      auto g = grid.nearest(qx, qy);
      for (const trg_b200::NodeGrid* gg : {&grid2, &grid3}) {
        const auto g2 = gg->nearest(qx, qy);
        CHECK(g2.d2 == g.d2 && g2.tie == g.tie && (g.tie || g2.entry == g.entry), "rebuilt / re-used grid nearest");
        CHECK(gg->count_in_range(qx, qy, 0.6f) == grid.count_in_range(qx, qy, 0.6f), "rebuilt / re-used grid range count");
      }
      const float dxw = xs[wn] - qx, dyw = ys[wn] - qy;
      float d2w = 0.f; d2w += dxw * dxw; d2w += dyw * dyw;
      CHECK(g.entry >= 0 && g.d2 == d2w, "grid nearest distance %g vs %g", g.d2, d2w);
      if (!g.tie) CHECK(g.entry == (int)wn, "grid nearest entry without tie");
      if (g.tie) {  // tree-free tie resolution must pick what the kd-tree traversal picks
        std::vector<int> cand;
        grid.for_each_within_d2(qx, qy, g.d2, std::sqrt(g.d2) * 1.001f + 1e-4f, [&](int e) { cand.push_back(e); });
        std::sort(cand.begin(), cand.end());
        ++n_ties;
        CHECK(cand.size() >= 2, "tie without candidates");
        CHECK(trg_b200::first_visited_of(xs.data(), ys.data(), cand, qx, qy) == (int)wn, "tie resolution (%zu candidates)", cand.size());
      }
      // 2.45, not 2.5: on the exact lattice of round 3 a point at distance == r along a split axis is
      // dropped by the reference traversal's strict `fabs(dx) < range` pruning (kdtree.c:289) although it
      // passes `dist_sq <= range^2` (:281) - a measure-zero quirk the grid does not reproduce
      for (float r : {0.3f, 0.6f, 2.45f}) {
        ref.range(qx, qy, r, &want);
        seq.range(qx, qy, r, got);
        bulk.range(qx, qy, r, got2);
        bool same = want.size() == got.size() && got == got2;
        for (size_t i = 0; same && i < want.size(); ++i) same = (want[i] == got[i]);
        CHECK(same, "range order r=%g", r);
        if (r < 1.0f) {  // grid candidates + root-path ordering must reproduce the traversal's result order
          std::vector<int> c2;
          grid.for_each_in_range(qx, qy, r, [&](int e) { c2.push_back(e); });
          seq.order_like_range(c2, qx, qy);
          CHECK(c2 == got, "order_like_range r=%g (%zu candidates)", r, c2.size());
          std::vector<int> c3 = got;
          std::reverse(c3.begin(), c3.end());
          bulk.order_like_range(c3, qx, qy);
          CHECK(c3 == got, "order_like_range on the bulk-built tree");
          std::vector<int> c4;
          grid2.for_each_in_range(qx, qy, r, [&](int e) { c4.push_back(e); });
          adopted.order_like_range(c4, qx, qy);
          CHECK(c4 == got, "order_like_range: rebuilt grid + adopted tree");
        }
        CHECK(grid.count_in_range(qx, qy, r) == (int)want.size(), "grid range count round=%d n=%d q=(%g,%g) r=%g got=%d want=%d", round, n, qx, qy, r, grid.count_in_range(qx, qy, r), (int)want.size());
      }
    }
  }
  if (n_ties < 1000) { printf("too few ties exercised: %ld\n", n_ties); ++fails; }
  printf(fails ? "FAILED %d\n" : "OK\n", fails);
  return fails ? 1 : 0;
}
