// tilechol_plan.h — host-side symbolic analysis of the sparse tile Cholesky that preconditions the PCG solve of the
// reduced camera system (K3).
//
// What it replaces: the reference solves the reduced camera system with Ceres' SPARSE_SCHUR, i.e. a fill-reducing
// ordering + supernodal sparse Cholesky in CHOLMOD (src/base3d/bundle_adjustment.cc:555).  Here the structure of the
// system is fixed for the life of a BA session, so everything symbolic is done ONCE on the host at session creation:
//   1. nested-dissection ordering of the image graph (geometric bisection on the camera centres where they separate
//      the graph well, breadth-first level sets otherwise),
//   2. images grouped into tiles of TC_TI images (TC_T = 6 * TC_TI unknowns), every dissection node padded to whole tiles,
//      tiles numbered bottom-up by the height of their node in the dissection tree (a topological order of the
//      elimination tree, so independent subtrees are factored side by side),
//   3. symbolic factorisation at tile granularity (column structures, elimination tree),
//   4. the left-looking task list: one task per tile L(i,j) with the list of products L(i,k) L(j,k)' it subtracts,
//   5. scatter maps from the stored 6 x 6 blocks of S (and the intrinsics border) into the tiles.
// The numeric phase (tilechol.cuh) is a persistent kernel that walks the task list with per-tile ready flags.
// Plain C++ (no CUDA) so that the CPU test suite can check the plan against a dense factorisation.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <cmath>
#include <atomic>
#include <memory>
#include <numeric>
#include <thread>
#include <vector>

namespace mm {

constexpr int TC_TI = 8;            // images per tile
constexpr int TC_T = 6 * TC_TI;     // unknowns per tile (48)
constexpr int TC_TT = TC_T * TC_T;  // doubles per tile (2304 = 18 KB)
constexpr int TC_BORDER_PER_TILE = TC_T / 9;   // cameras (9 intrinsics each) per border tile (5, 3 unknowns padded)

struct TileCholPlan {
  int n_img = 0, n_cam_border = 0;
  int nt_pose = 0, nt = 0;                 // pose tiles, all tiles (pose + border)
  int64_t n_l = 0;                         // tiles of L (lower triangle incl. diagonal), = number of tasks
  int64_t n_upd = 0;                       // tile products over all tasks
  int n_nodes = 0, max_height = 0;
  std::vector<int> img_tile, img_slot;     // image -> tile, slot within the tile
  std::vector<int> tile_nunk;              // real unknowns per tile (the rest is padding: unit diagonal)
  std::vector<int> tile_height;            // height of the tile's dissection node (border: max + 1)
  std::vector<int64_t> col_ptr;            // [nt + 1] CSC over L tiles; the diagonal tile is the first of its column
  std::vector<int> row_idx;                // [n_l] tile row of every L tile
  std::vector<int> col_idx;                // [n_l] tile column of every L tile
  std::vector<int> has_a;                  // [n_l] 1 = receives entries of S (must be zeroed + scattered before the factorisation)
  std::vector<int64_t> upd_ptr;            // [n_l + 1]
  std::vector<int> upd_a, upd_b;           // [n_upd] L tile ids of (i,k) and (j,k)
  std::vector<int64_t> rowp_ptr;           // [nt + 1] row patterns (strictly lower): L tile ids (j,k), k ascending
  std::vector<int> rowp_tile, rowp_col;    // [n_l - nt]
  std::vector<int> unk_of;                 // [nt * TC_T] original unknown (6 * img + r | 6 * n_img + 9 * cam + r) or -1 (padding)
  // scatter of the stored S blocks (n_img diagonal blocks, then the off-diagonal blocks): destination tile and offset
  std::vector<int> sc_tile, sc_off;        // sc_off = (row0 * TC_T... see tile_elem) | transpose flag in bit 30
  std::vector<int> a_tiles;                // L tile ids with has_a
  double flops = 0.0;                      // 2 * TC_T^3 per product (diagnostic)
  // ---- dense node blocks.  The tiles of one dissection node are consecutive; W = inverse of the node's (lower triangular,
  // dense) diagonal block of L, computed tile by tile after the factorisation of the node:  W(j,j) = inv(L(j,j)),
  // W(i,j) = -inv(L(i,i)) sum_{k=j}^{i-1} L(i,k) W(k,j).  With W the substitutions treat a whole node in one step instead
  // of tile row by tile row (the dependency chain of a solve shrinks from ~#tile rows on the root path to ~#levels).
  int n_tnodes = 0;
  std::vector<int> tile_node, node_first, node_nt;
  std::vector<int64_t> w_row_ptr;          // [nt + 1]  W(i, j), j = node_first .. i, stored at w_row_ptr[i] + (j - node_first)
  int64_t n_w = 0;                         // tiles of W (two copies on the device: column-major and row-major)
  int64_t n_wtask = 0, n_wupd = 0;         // W tasks = off-diagonal tiles of W, task ids n_l .. n_l + n_wtask - 1
  std::vector<int> wt_row, wt_col, wt_store;
  std::vector<int64_t> wupd_ptr;           // [n_wtask + 1]
  std::vector<int> wupd_l, wupd_w, wupd_flag;   // L tile id (i,k) | W storage index (k,j) | task whose flag says W(k,j) is ready
  std::vector<int> task_order;             // [n_l + n_wtask] static list schedule (a topological order)
  // ---- substitution tasks (forward then backward, one list).  Every task writes one 48-vector slot; flag index = slot.
  //   kind 0 MV   out = sum_items M v        (M column-major; items: matrix = L | WC | WR tile, v = slot)
  //   kind 1 MVT  out = sum_items M' v       (M column-major tile of L)
  //   kind 2 SUM  out = base - sum_items v   (base: rhs of tile row `tile` (forward) or a slot (backward))
  //   kind 3 MV + scatter of the result to the caller's unknown order (final x)
  int n_slots = 0, n_stasks = 0;
  std::vector<int> st_kind, st_out, st_base, st_tile;
  std::vector<int64_t> st_item_ptr;
  std::vector<int> it_mat, it_src;          // it_mat = selector << 28 | tile index (selector 0 L, 1 WC, 2 WR); it_src = slot
};
constexpr int TC_SOLVE_CHUNK = 4;           // tiles per accumulation task of the substitutions
enum { TC_ST_MV = 0, TC_ST_MVT = 1, TC_ST_SUM = 2, TC_ST_MV_OUT = 3 };
enum { TC_MAT_L = 0, TC_MAT_WC = 1, TC_MAT_WR = 2 };

// element (r, c) of a tile: column-major, so that both operands of  C -= A B'  are read along contiguous runs
#ifdef __CUDACC__
__host__ __device__
#endif
inline int tile_elem(int r, int c) { return c * TC_T + r; }

namespace tc_detail {

struct Graph {
  int n = 0;
  std::vector<int> ptr, adj;
};

inline Graph build_graph(int n, int n_off, const int* a, const int* b) {
  Graph g; g.n = n; g.ptr.assign((size_t)n + 1, 0);
  for (int e = 0; e < n_off; ++e) if (a[e] != b[e]) { g.ptr[a[e] + 1]++; g.ptr[b[e] + 1]++; }
  for (int i = 0; i < n; ++i) g.ptr[i + 1] += g.ptr[i];
  g.adj.resize((size_t)g.ptr[n]);
  std::vector<int> fill(g.ptr.begin(), g.ptr.end() - 1);
  for (int e = 0; e < n_off; ++e) if (a[e] != b[e]) { g.adj[fill[a[e]]++] = b[e]; g.adj[fill[b[e]]++] = a[e]; }
  return g;
}

struct NDNode { std::vector<int> own; int parent = -1; int height = 0; };

// Nested dissection.  The two parts of a separated subset are dissected side by side on host threads (top levels only); every
// call works on its own vertices, reads foreign vertices only through their stamp (relaxed atomics, marks are drawn from one
// global counter and therefore never collide), and the tree is flattened in the order a sequential run creates its nodes, so
// the plan does not depend on the number of threads.
struct Dissector {
  struct Tree { std::vector<int> own; std::vector<Tree*> kids; ~Tree() { for (Tree* k : kids) delete k; } };
  const Graph& g; const double* pos; int leaf;
  std::unique_ptr<std::atomic<int>[]> stamp, stamp_visit; std::vector<int> level; std::atomic<int> cur{0};
  std::vector<NDNode> nodes;
  int par_depth = 3;
  Dissector(const Graph& g_, const double* pos_, int leaf_) : g(g_), pos(pos_), leaf(leaf_), stamp(new std::atomic<int>[(size_t)std::max(g_.n, 1)]),
      stamp_visit(new std::atomic<int>[(size_t)std::max(g_.n, 1)]), level((size_t)g_.n, 0) {
    for (int i = 0; i < g.n; ++i) { stamp[i].store(0, std::memory_order_relaxed); stamp_visit[i].store(0, std::memory_order_relaxed); }
    if (const char* e = getenv("MM_TC_PLAN_THREADS")) par_depth = atoi(e) <= 1 ? 0 : 3;
    if (std::thread::hardware_concurrency() < 4) par_depth = 0;
  }
  int st(int v) const { return stamp[v].load(std::memory_order_relaxed); }
  int sv(int v) const { return stamp_visit[v].load(std::memory_order_relaxed); }
  int next_mark() { return cur.fetch_add(1, std::memory_order_relaxed) + 1; }

  // breadth-first levels inside the subset marked with `mark`; returns the visit order (one component from `start`)
  void bfs(int start, int mark, std::vector<int>& order, std::vector<int>& lvl_start) {
    order.clear(); lvl_start.clear();
    const int visited = next_mark();
    order.push_back(start); stamp_visit[start].store(visited, std::memory_order_relaxed); level[start] = 0; lvl_start.push_back(0);
    size_t head = 0; int cur_level = 0;
    while (head < order.size()) {
      const int v = order[head++];
      if (level[v] > cur_level) { cur_level = level[v]; lvl_start.push_back((int)head - 1); }
      for (int e = g.ptr[v]; e < g.ptr[v + 1]; ++e) {
        const int u = g.adj[e];
        if (st(u) == mark && sv(u) != visited) { stamp_visit[u].store(visited, std::memory_order_relaxed); level[u] = level[v] + 1; order.push_back(u); }
      }
    }
    lvl_start.push_back((int)order.size());
  }

  void flatten(Tree* t, int parent) {
    // a Tree without vertices only groups the components of a disconnected subset: its kids hang under `parent`
    int id = parent;
    if (!t->own.empty()) { NDNode nd; nd.own.swap(t->own); nd.parent = parent; nodes.push_back(std::move(nd)); id = (int)nodes.size() - 1; }
    for (Tree* k : t->kids) flatten(k, id);
  }
  void run(std::vector<int>& all) {
    Tree root;
    dissect(all, &root, 0);
    flatten(&root, -1);
    // heights bottom-up (children were created after their parent)
    for (int i = (int)nodes.size() - 1; i >= 0; --i) if (nodes[i].parent >= 0) nodes[nodes[i].parent].height = std::max(nodes[nodes[i].parent].height, nodes[i].height + 1);
  }

  // appends the dissection of `sub` to `out->kids`
  void dissect(std::vector<int>& sub, Tree* out, int depth) {
    if (sub.empty()) return;
    // connected components of the subset become siblings
    const int mark = next_mark();
    for (int v : sub) stamp[v].store(mark, std::memory_order_relaxed);
    std::vector<int> order, lvl;
    std::vector<std::vector<int>> comps;
    {
      const int seen = next_mark();
      for (int v : sub) {
        if (sv(v) >= seen) continue;       // reached by a sweep of this call (their stamps are > seen)
        bfs(v, mark, order, lvl);
        comps.push_back(order);
      }
    }
    if (comps.size() > 1) {
      for (auto& c : comps) dissect(c, out, depth);
      return;
    }
    Tree* nd = new Tree(); out->kids.push_back(nd);
    if ((int)sub.size() <= leaf) { nd->own = sub; return; }
    std::vector<int> S, A, B;
    if (!separate(sub, S, A, B)) { nd->own = sub; return; }
    nd->own = S;
    if (depth < par_depth && (int)A.size() > 256 && (int)B.size() > 256) {
      Tree ta;                                   // A's nodes come first in the sequential order
      std::thread th; bool spawned = false;
      try { th = std::thread([&] { dissect(A, &ta, depth + 1); }); spawned = true; } catch (...) {}
      if (!spawned) dissect(A, &ta, depth + 1);
      Tree tb; dissect(B, &tb, depth + 1);
      if (spawned) th.join();
      for (Tree* k : ta.kids) nd->kids.push_back(k);
      for (Tree* k : tb.kids) nd->kids.push_back(k);
      ta.kids.clear(); tb.kids.clear();
    } else {
      dissect(A, nd, depth + 1); dissect(B, nd, depth + 1);
    }
  }

  // vertex separator of a connected subset: the best of (a) the boundary of a median cut along each coordinate axis and
  // (b) a breadth-first level set from a pseudo-peripheral vertex.  cost = |S| with a penalty for unbalanced parts.
  bool separate(const std::vector<int>& sub, std::vector<int>& S, std::vector<int>& A, std::vector<int>& B) {
    const int n = (int)sub.size();
    double best = 1e300; std::vector<int> side_best;      // side: 0 = A, 1 = B, 2 = separator (indexed like sub)
    std::vector<int> side((size_t)n);
    const int mark = next_mark();
    for (int v : sub) stamp[v].store(mark, std::memory_order_relaxed);
    auto evaluate = [&](const std::vector<int>& sd) {
      int na = 0, nb = 0, ns = 0;
      for (int s : sd) { if (s == 0) ++na; else if (s == 1) ++nb; else ++ns; }
      if (na == 0 || nb == 0 || ns == 0) return 1e300;
      const double bal = (double)std::max(na, nb) / (double)(na + nb);
      return (double)ns * (bal > 0.7 ? 4.0 : 1.0);
    };
    // index of every vertex inside `sub`
    for (int i = 0; i < n; ++i) level[sub[i]] = i;        // (level[] is free here: bfs is not running)
    auto cut_to_separator = [&](std::vector<int>& sd) {
      // sd holds 0/1 from an edge cut: the smaller of the two boundaries becomes the separator
      int ba = 0, bb = 0;
      std::vector<char> touch((size_t)n, 0);
      for (int i = 0; i < n; ++i) {
        const int v = sub[i];
        for (int e = g.ptr[v]; e < g.ptr[v + 1]; ++e) { const int u = g.adj[e]; if (st(u) == mark && sd[level[u]] != sd[i]) { touch[i] = 1; break; } }
        if (touch[i]) { if (sd[i] == 0) ++ba; else ++bb; }
      }
      const int pick = ba <= bb ? 0 : 1;
      for (int i = 0; i < n; ++i) if (sd[i] == pick && touch[i]) sd[i] = 3;     // (marked first, renamed below: the test above reads the 0/1 sides)
      for (int i = 0; i < n; ++i) if (sd[i] == 3) sd[i] = 2;
    };
    if (pos) {
      std::vector<int> idx((size_t)n);
      double ext[3]; double ext_max = 0.0;
      for (int ax = 0; ax < 3; ++ax) {
        double lo = 1e300, hi = -1e300;
        for (int v : sub) { lo = std::min(lo, pos[3 * (size_t)v + ax]); hi = std::max(hi, pos[3 * (size_t)v + ax]); }
        ext[ax] = hi - lo; ext_max = std::max(ext_max, ext[ax]);
      }
      for (int ax = 0; ax < 3; ++ax) {
        if (!(ext[ax] > 0.0) || ext[ax] < 0.05 * ext_max) continue;      // (a cut across a thin direction, e.g. the flying height, never separates well)
        std::iota(idx.begin(), idx.end(), 0);
        std::sort(idx.begin(), idx.end(), [&](int x, int y) { const double px = pos[3 * (size_t)sub[x] + ax], py = pos[3 * (size_t)sub[y] + ax]; return px < py || (px == py && sub[x] < sub[y]); });
        for (int r = 0; r < n; ++r) side[idx[r]] = r < n / 2 ? 0 : 1;
        cut_to_separator(side);
        const double c = evaluate(side);
        if (c < best) { best = c; side_best = side; }
      }
    }
    {
      // pseudo-peripheral start: two sweeps
      std::vector<int> order, lvl;
      bfs(sub[0], mark, order, lvl);
      int far = order.back();
      bfs(far, mark, order, lvl);
      far = order.back();
      bfs(far, mark, order, lvl);
      const int nl = (int)lvl.size() - 1;
      // bfs clobbered level[]: rebuild the local index afterwards; candidates first
      int best_l = -1; double best_c = 1e300;
      for (int l = 1; l + 1 < nl; ++l) {
        const int na = lvl[l], ns = lvl[l + 1] - lvl[l], nb = n - lvl[l + 1];
        if (na == 0 || nb == 0) continue;
        const double bal = (double)std::max(na, nb) / (double)(na + nb);
        if (bal > 0.7) continue;                           // peeling thin layers off one end gives a chain, not a tree
        const double c = (double)ns;
        if (c < best_c) { best_c = c; best_l = l; }
      }
      if (best_l < 0) {                                    // no balanced level: the one where the sweep has seen half of the subset
        for (int l = 1; l + 1 < nl; ++l) if (lvl[l + 1] * 2 >= n && lvl[l] > 0 && n - lvl[l + 1] > 0) { best_l = l; best_c = 4.0 * (lvl[l + 1] - lvl[l]); break; }
      }
      if (best_l >= 0 && best_c < best) {
        std::vector<int> lv_of((size_t)n);
        std::vector<int> ord(order);
        for (int i = 0; i < n; ++i) level[sub[i]] = i;
        for (int l = 0; l < nl; ++l) for (int q = lvl[l]; q < lvl[l + 1]; ++q) lv_of[level[ord[q]]] = l;
        for (int i = 0; i < n; ++i) side[i] = lv_of[i] < best_l ? 0 : (lv_of[i] == best_l ? 2 : 1);
        best = best_c; side_best = side;
      } else {
        for (int i = 0; i < n; ++i) level[sub[i]] = i;
      }
    }
    if (side_best.empty() || !(best < 1e299)) return false;
    S.clear(); A.clear(); B.clear();
    for (int i = 0; i < n; ++i) { if (side_best[i] == 0) A.push_back(sub[i]); else if (side_best[i] == 1) B.push_back(sub[i]); else S.push_back(sub[i]); }
    // a separator that swallows most of the subset is not worth a level of the tree
    if ((int)S.size() * 2 > n) return false;
    return !A.empty() && !B.empty() && !S.empty();
  }
};

}  // namespace tc_detail

// n_off off-diagonal blocks (a[e], b[e]) of the reduced camera system, stored after the n_img diagonal blocks;
// pos: optional camera centres [3 * n_img] (NULL: graph-only dissection); n_cam_border: cameras whose 9 intrinsics are
// unknowns of the system (dense border rows, eliminated last).
inline int build_tilechol_plan(int n_img, int n_off, const int* blk_a, const int* blk_b, const double* pos, int n_cam_border,
                               TileCholPlan& P, int leaf_images = 2 * TC_TI) {
  using namespace tc_detail;
  P = TileCholPlan();
  P.n_img = n_img; P.n_cam_border = n_cam_border;
  if (n_img <= 0) return -1;
  Graph g = build_graph(n_img, n_off, blk_a, blk_b);
  if (const char* e = getenv("MM_TC_LEAF")) leaf_images = std::max(1, atoi(e));
  Dissector D(g, pos, leaf_images);
  { std::vector<int> all((size_t)n_img); std::iota(all.begin(), all.end(), 0); D.run(all); }
  std::vector<NDNode>& nodes = D.nodes;
  P.n_nodes = (int)nodes.size();
  // ---- tiles: nodes bottom-up by height; every node padded to whole tiles
  std::vector<int> node_order((size_t)nodes.size());
  std::iota(node_order.begin(), node_order.end(), 0);
  std::stable_sort(node_order.begin(), node_order.end(), [&](int x, int y) { return nodes[x].height < nodes[y].height; });
  P.img_tile.assign((size_t)n_img, -1); P.img_slot.assign((size_t)n_img, -1);
  int nt = 0;
  for (int id : node_order) {
    NDNode& nd = nodes[id];
    if (nd.own.empty()) continue;
    P.node_first.push_back(nt); P.node_nt.push_back((int)((nd.own.size() + TC_TI - 1) / TC_TI));
    P.max_height = std::max(P.max_height, nd.height);
    std::sort(nd.own.begin(), nd.own.end());          // keeps images that follow each other in the sequence in one tile
    for (size_t q = 0; q < nd.own.size(); ++q) {
      const int t = nt + (int)(q / TC_TI);
      P.img_tile[nd.own[q]] = t; P.img_slot[nd.own[q]] = (int)(q % TC_TI);
      if ((int)P.tile_nunk.size() <= t) { P.tile_nunk.resize((size_t)t + 1, 0); P.tile_height.resize((size_t)t + 1, 0); }
      P.tile_nunk[t] += 6; P.tile_height[t] = nd.height;
    }
    nt += (int)((nd.own.size() + TC_TI - 1) / TC_TI);
  }
  for (int i = 0; i < n_img; ++i) if (P.img_tile[i] < 0) return -2;
  P.nt_pose = nt;
  const int nt_border = (n_cam_border + TC_BORDER_PER_TILE - 1) / TC_BORDER_PER_TILE;
  for (int b = 0; b < nt_border; ++b) {
    const int cams = std::min(TC_BORDER_PER_TILE, n_cam_border - b * TC_BORDER_PER_TILE);
    P.tile_nunk.push_back(9 * cams); P.tile_height.push_back(P.max_height + 1);
  }
  if (nt_border) { P.node_first.push_back(nt); P.node_nt.push_back(nt_border); }
  nt += nt_border; P.nt = nt;
  P.n_tnodes = (int)P.node_first.size();
  P.tile_node.assign((size_t)nt, 0);
  for (int nd = 0; nd < P.n_tnodes; ++nd) for (int q = 0; q < P.node_nt[nd]; ++q) P.tile_node[P.node_first[nd] + q] = nd;
  // ---- unknown map
  P.unk_of.assign((size_t)nt * TC_T, -1);
  for (int i = 0; i < n_img; ++i) for (int r = 0; r < 6; ++r) P.unk_of[(size_t)P.img_tile[i] * TC_T + 6 * P.img_slot[i] + r] = 6 * i + r;
  for (int c = 0; c < n_cam_border; ++c) for (int r = 0; r < 9; ++r)
    P.unk_of[(size_t)(P.nt_pose + c / TC_BORDER_PER_TILE) * TC_T + 9 * (c % TC_BORDER_PER_TILE) + r] = 6 * n_img + 9 * c + r;
  // ---- lower tile pattern of A
  std::vector<std::vector<int>> acol((size_t)nt);      // rows > j with entries of A in column j
  for (int e = 0; e < n_off; ++e) {
    const int ta = P.img_tile[blk_a[e]], tb = P.img_tile[blk_b[e]];
    if (ta != tb) acol[std::min(ta, tb)].push_back(std::max(ta, tb));
  }
  for (int j = 0; j < P.nt_pose; ++j) for (int b = 0; b < nt_border; ++b) acol[j].push_back(P.nt_pose + b);
  for (int b = 0; b < nt_border; ++b) for (int b2 = b + 1; b2 < nt_border; ++b2) acol[P.nt_pose + b].push_back(P.nt_pose + b2);
  for (auto& v : acol) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
  // ---- symbolic factorisation: struct(j) = A(j) U  union over children c of struct(c) \ {j};  parent(j) = min struct(j)
  std::vector<std::vector<int>> st((size_t)nt), children((size_t)nt);
  std::vector<int> mark((size_t)nt, -1);
  for (int j = 0; j < nt; ++j) {
    std::vector<int>& s = st[j];
    mark[j] = j;
    for (int r : acol[j]) if (mark[r] != j) { mark[r] = j; s.push_back(r); }
    for (int c : children[j]) for (int r : st[c]) if (r != j && mark[r] != j) { mark[r] = j; s.push_back(r); }
    std::sort(s.begin(), s.end());
    if (!s.empty()) children[s[0]].push_back(j);
  }
  // ---- CSC of L tiles, diagonal first
  P.col_ptr.assign((size_t)nt + 1, 0);
  for (int j = 0; j < nt; ++j) P.col_ptr[j + 1] = P.col_ptr[j] + 1 + (int64_t)st[j].size();
  P.n_l = P.col_ptr[nt];
  if (P.n_l >= ((int64_t)1 << 30)) return -3;                    // tile ids stay in int range
  P.row_idx.resize((size_t)P.n_l); P.col_idx.resize((size_t)P.n_l); P.has_a.assign((size_t)P.n_l, 0);
  for (int j = 0; j < nt; ++j) {
    int64_t q = P.col_ptr[j];
    P.row_idx[q] = j; P.col_idx[q] = j; P.has_a[q] = 1; ++q;
    for (int r : st[j]) { P.row_idx[q] = r; P.col_idx[q] = j; ++q; }
  }
  auto tile_id = [&](int i, int j) -> int64_t {       // L tile (i, j), i >= j; -1 if not in the structure
    if (i == j) return P.col_ptr[j];
    const std::vector<int>& s = st[j];
    auto it = std::lower_bound(s.begin(), s.end(), i);
    if (it == s.end() || *it != i) return -1;
    return P.col_ptr[j] + 1 + (it - s.begin());
  };
  for (int j = 0; j < nt; ++j) for (int r : acol[j]) { const int64_t t = tile_id(r, j); if (t < 0) return -4; P.has_a[t] = 1; }
  for (int64_t t = 0; t < P.n_l; ++t) if (P.has_a[t]) P.a_tiles.push_back((int)t);
  // ---- row patterns
  P.rowp_ptr.assign((size_t)nt + 1, 0);
  for (int k = 0; k < nt; ++k) for (int r : st[k]) P.rowp_ptr[r + 1]++;
  for (int j = 0; j < nt; ++j) P.rowp_ptr[j + 1] += P.rowp_ptr[j];
  P.rowp_tile.resize((size_t)P.rowp_ptr[nt]); P.rowp_col.resize((size_t)P.rowp_ptr[nt]);
  { std::vector<int64_t> fill(P.rowp_ptr.begin(), P.rowp_ptr.end() - 1);
    for (int k = 0; k < nt; ++k) for (size_t q = 0; q < st[k].size(); ++q) { const int r = st[k][q]; P.rowp_tile[fill[r]] = (int)(P.col_ptr[k] + 1 + (int64_t)q); P.rowp_col[fill[r]] = k; fill[r]++; } }
  P.w_row_ptr.assign((size_t)nt + 1, 0);
  for (int i = 0; i < nt; ++i) P.w_row_ptr[i + 1] = P.w_row_ptr[i] + (i - P.node_first[P.tile_node[i]] + 1);
  P.n_w = P.w_row_ptr[nt];
  auto w_index = [&](int i, int j) -> int { return (int)(P.w_row_ptr[i] + (j - P.node_first[P.tile_node[i]])); };
  // The scatter map and the substitution task list depend only on the symbolic structure: they are built on a second host
  // thread beside the update lists and the schedule of the factorisation.
  auto side_work = [&]() -> int {
    // ---- scatter map of the stored S blocks
    P.sc_tile.resize((size_t)n_img + n_off); P.sc_off.resize((size_t)n_img + n_off);
    for (int i = 0; i < n_img; ++i) { P.sc_tile[i] = (int)P.col_ptr[P.img_tile[i]]; P.sc_off[i] = tile_elem(6 * P.img_slot[i], 6 * P.img_slot[i]); }
    for (int e = 0; e < n_off; ++e) {
      const int a = blk_a[e], b = blk_b[e];          // stored block = S(a, b), a < b in the caller's numbering
      const int ta = P.img_tile[a], tb = P.img_tile[b], sa = P.img_slot[a], sb = P.img_slot[b];
      // lower triangle in the new numbering: row = the later of the two
      const bool a_is_row = ta > tb || (ta == tb && sa > sb);
      const int tr = a_is_row ? ta : tb, tcn = a_is_row ? tb : ta, sr = a_is_row ? sa : sb, scn = a_is_row ? sb : sa;
      const int64_t t = tile_id(tr, tcn);
      if (t < 0) return -6;
      P.sc_tile[n_img + e] = (int)t;
      // a_is_row: the tile entry (row of a, column of b) is S(a,b) as stored; otherwise the entry (row of b, column of a) is S(a,b)'
      P.sc_off[n_img + e] = tile_elem(6 * sr, 6 * scn) | (a_is_row ? 0 : (1 << 30));
    }
    // ---- substitution tasks
    {
      // slots: [0, nt) forward right-hand sides t, [nt, 2nt) y, [2nt, 3nt) backward right-hand sides s, [3nt, 4nt) x, then partial sums
      int n_slots = 4 * nt;
      auto push_task = [&](int kind, int out, int base, int tile) {
        P.st_kind.push_back(kind); P.st_out.push_back(out); P.st_base.push_back(base); P.st_tile.push_back(tile);
        P.st_item_ptr.push_back((int64_t)P.it_mat.size());
      };
      auto push_item = [&](int sel, int64_t idx, int src) { P.it_mat.push_back((sel << 28) | (int)idx); P.it_src.push_back(src); };
      if (P.n_l >= (1 << 28) || P.n_w >= (1 << 28)) return -7;
      const int H = P.tile_height.empty() ? 0 : *std::max_element(P.tile_height.begin(), P.tile_height.end());
      std::vector<std::vector<int>> by_h((size_t)H + 1);
      for (int i = 0; i < nt; ++i) by_h[P.tile_height[i]].push_back(i);
      std::vector<std::vector<int>> partials((size_t)nt);
      // forward: level by level, bottom-up
      for (int h = 0; h <= H; ++h) {
        for (int i : by_h[h]) {                                  // accumulation over the columns of descendant nodes
          const int f = P.node_first[P.tile_node[i]];
          int in_chunk = 0;
          for (int64_t q = P.rowp_ptr[i]; q < P.rowp_ptr[i + 1]; ++q) {
            const int k = P.rowp_col[q];
            if (k >= f) continue;                                 // same node: covered by W
            if (in_chunk == 0) { push_task(TC_ST_MV, n_slots, -1, i); partials[i].push_back(n_slots); ++n_slots; }
            push_item(TC_MAT_L, P.rowp_tile[q], nt + k);
            if (++in_chunk == TC_SOLVE_CHUNK) in_chunk = 0;
          }
        }
        for (int i : by_h[h]) {                                  // t_i = b_i - sum of the partial sums
          push_task(TC_ST_SUM, i, -1, i);
          for (int sl : partials[i]) push_item(0, 0, sl);
        }
        for (int i : by_h[h]) {                                  // y_i = sum_j W(i,j) t_j over the node
          const int f = P.node_first[P.tile_node[i]];
          push_task(TC_ST_MV, nt + i, -1, i);
          for (int j = f; j <= i; ++j) push_item(TC_MAT_WC, w_index(i, j), j);
        }
      }
      // backward: top-down
      for (auto& v : partials) v.clear();
      for (int h = H; h >= 0; --h) {
        for (int j : by_h[h]) {                                  // sum over the rows of ancestor nodes of L(i,j)' x_i
          const int last = P.node_first[P.tile_node[j]] + P.node_nt[P.tile_node[j]] - 1;
          int in_chunk = 0;
          for (int64_t id = P.col_ptr[j] + 1; id < P.col_ptr[j + 1]; ++id) {
            const int i = P.row_idx[id];
            if (i <= last) continue;
            if (in_chunk == 0) { push_task(TC_ST_MVT, n_slots, -1, j); partials[j].push_back(n_slots); ++n_slots; }
            push_item(TC_MAT_L, id, 3 * nt + i);
            if (++in_chunk == TC_SOLVE_CHUNK) in_chunk = 0;
          }
        }
        for (int j : by_h[h]) {                                  // s_j = y_j - partial sums
          push_task(TC_ST_SUM, 2 * nt + j, nt + j, j);
          for (int sl : partials[j]) push_item(0, 0, sl);
        }
        for (int j : by_h[h]) {                                  // x_j = sum_i W(i,j)' s_i over the node  (row-major W(i,j) = column-major W(i,j)')
          const int last = P.node_first[P.tile_node[j]] + P.node_nt[P.tile_node[j]] - 1;
          push_task(TC_ST_MV_OUT, 3 * nt + j, -1, j);
          for (int i = j; i <= last; ++i) push_item(TC_MAT_WR, w_index(i, j), 2 * nt + i);
        }
      }
      P.st_item_ptr.push_back((int64_t)P.it_mat.size());
      P.n_slots = n_slots; P.n_stasks = (int)P.st_kind.size();
    }
    return 0;
  };
  int side_rc = 0;
  std::thread side_thread;
  const bool side_par = std::thread::hardware_concurrency() >= 4 && !getenv("MM_TC_PLAN_SERIAL");
  bool side_spawned = false;
  if (side_par) { try { side_thread = std::thread([&] { side_rc = side_work(); }); side_spawned = true; } catch (...) {} }
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } side_joiner{ side_thread };
  // ---- update lists: task (i, j) subtracts L(i,k) L(j,k)' for every k < j with i, j in struct(k).  One enumeration (j, then k
  // ascending, then i ascending: struct(k) and struct(j) are both sorted, so the tile of (i, j) is found by a merge walk), then a
  // stable counting sort by task keeps that order inside every task.
  {
    // (the columns j are independent: ranges of them are enumerated on host threads and concatenated in order)
    const int n_part = (std::thread::hardware_concurrency() >= 4 && !getenv("MM_TC_PLAN_SERIAL") && nt >= 64) ? 4 : 1;
    std::vector<int> put[4], pua[4], pub[4]; int perr[4] = {0, 0, 0, 0};
    auto enumerate = [&](int part) {
      // balanced by the number of row-pattern entries
      const int64_t tot = P.rowp_ptr[nt];
      const int j0 = (int)(std::lower_bound(P.rowp_ptr.begin(), P.rowp_ptr.end(), tot * part / n_part) - P.rowp_ptr.begin());
      const int j1 = part + 1 == n_part ? nt : (int)(std::lower_bound(P.rowp_ptr.begin(), P.rowp_ptr.end(), tot * (part + 1) / n_part) - P.rowp_ptr.begin());
      std::vector<int> ut, ua, ub;                  // (locals: the headers of put[] share cache lines)
      const size_t guess = (size_t)(P.n_l * 20 / n_part) + 1024; ut.reserve(guess); ua.reserve(guess); ub.reserve(guess);
      for (int j = std::min(j0, nt); j < std::min(j1, nt); ++j) {
        const std::vector<int>& sj = st[j];
        for (int64_t q = P.rowp_ptr[j]; q < P.rowp_ptr[j + 1]; ++q) {
          const int k = P.rowp_col[q]; const int tjk = P.rowp_tile[q];
          const std::vector<int>& s = st[k];
          const size_t first = std::lower_bound(s.begin(), s.end(), j) - s.begin();      // s[first] == j
          size_t w = 0;
          for (size_t u = first; u < s.size(); ++u) {
            const int i = s[u];
            int64_t t;
            if (i == j) t = P.col_ptr[j];
            else {
              while (w < sj.size() && sj[w] < i) ++w;
              if (w == sj.size() || sj[w] != i) { perr[part] = -5; return; }                  // fill must contain it
              t = P.col_ptr[j] + 1 + (int64_t)w;
            }
            ut.push_back((int)t); ua.push_back((int)(P.col_ptr[k] + 1 + (int64_t)u)); ub.push_back(tjk);
          }
        }
      }
      put[part].swap(ut); pua[part].swap(ua); pub[part].swap(ub);
    };
    { std::thread th[4]; bool spawned[4] = {false, false, false, false};
      for (int part = 1; part < n_part; ++part) { try { th[part] = std::thread(enumerate, part); spawned[part] = true; } catch (...) {} }
      enumerate(0);
      for (int part = 1; part < n_part; ++part) { if (spawned[part]) th[part].join(); else enumerate(part); } }
    for (int part = 0; part < n_part; ++part) if (perr[part]) return perr[part];
    std::vector<int> ut, ua, ub;
    { size_t tot = 0; for (int part = 0; part < n_part; ++part) tot += put[part].size();
      ut.reserve(tot); ua.reserve(tot); ub.reserve(tot);
      for (int part = 0; part < n_part; ++part) { ut.insert(ut.end(), put[part].begin(), put[part].end()); ua.insert(ua.end(), pua[part].begin(), pua[part].end()); ub.insert(ub.end(), pub[part].begin(), pub[part].end()); } }
    P.n_upd = (int64_t)ut.size();
    P.upd_ptr.assign((size_t)P.n_l + 1, 0);
    for (int t : ut) P.upd_ptr[(size_t)t + 1]++;
    for (int64_t t = 0; t < P.n_l; ++t) P.upd_ptr[t + 1] += P.upd_ptr[t];
    P.upd_a.resize((size_t)P.n_upd); P.upd_b.resize((size_t)P.n_upd);
    std::vector<int64_t> fill(P.upd_ptr.begin(), P.upd_ptr.end() - 1);
    for (size_t e = 0; e < ut.size(); ++e) { const int64_t d = fill[ut[e]]++; P.upd_a[d] = ua[e]; P.upd_b[d] = ub[e]; }
  }
  P.flops = 2.0 * TC_T * TC_T * TC_T * ((double)P.n_upd + (double)(P.n_l - nt)) + (double)nt * TC_T * TC_T * TC_T * (2.0 / 3.0);
  // ---- W: inverse of every node's diagonal block
  std::vector<int> w_task_of((size_t)P.n_w, -1);          // W storage index -> task id whose flag covers it
  for (int j = 0; j < nt; ++j) w_task_of[w_index(j, j)] = (int)P.col_ptr[j];
  P.wupd_ptr.assign(1, 0);
  for (int nd = 0; nd < P.n_tnodes; ++nd) {
    const int f = P.node_first[nd], m = P.node_nt[nd];
    for (int i = f + 1; i < f + m; ++i) for (int j = f; j < i; ++j) {
      const int task = (int)(P.n_l + (int64_t)P.wt_row.size());
      P.wt_row.push_back(i); P.wt_col.push_back(j); P.wt_store.push_back(w_index(i, j));
      w_task_of[w_index(i, j)] = task;
      for (int k = j; k < i; ++k) {
        const int64_t t = tile_id(i, k);
        if (t < 0) continue;
        P.wupd_l.push_back((int)t); P.wupd_w.push_back(w_index(k, j)); P.wupd_flag.push_back(w_task_of[w_index(k, j)]);
      }
      P.wupd_ptr.push_back((int64_t)P.wupd_l.size());
    }
  }
  P.n_wtask = (int64_t)P.wt_row.size(); P.n_wupd = (int64_t)P.wupd_l.size();
  P.flops += 2.0 * TC_T * TC_T * TC_T * ((double)P.n_wupd + (double)P.n_wtask);
  // ---- static list schedule.  Tasks are handed to the CTAs of the persistent kernel in ONE fixed order and a CTA waits
  // in place for the operands of its task, so the order decides how much of the machine idles.  A cost model (time per
  // tile product, per Cholesky, per hop of the ready flags) gives every task its earliest finish on an unbounded machine
  // and from it the latest start that does not delay that finish; tasks are issued by that start time (made monotone along
  // dependencies so that the order stays topological: a task never waits for one that is issued after it).
  {
    const int64_t nT = P.n_l + P.n_wtask;
    // measured on B200 with two CTAs per SM (tools/tc_trace.py): ~3.5 us per tile product, ~30 us for a diagonal tile
    // (Cholesky + inverse), ~8 us for the finishing product of an off-diagonal tile, ~3 us from a flag to the operand on chip
    double g = 3.5, lat = 3.0, chol = 30.0, fin = 8.0;
    if (const char* e = getenv("MM_TC_COST")) sscanf(e, "%lf,%lf,%lf,%lf", &g, &lat, &chol, &fin);
    std::vector<double> finish((size_t)nT, 0.0), key((size_t)nT, 0.0);
    for (int64_t t = 0; t < P.n_l; ++t) {
      double cur = 0.0, kmax = 0.0;
      for (int64_t u = P.upd_ptr[t]; u < P.upd_ptr[t + 1]; ++u) {
        const int a = P.upd_a[u], b = P.upd_b[u];
        cur = std::max(cur, std::max(finish[a], finish[b]) + lat) + g;
        kmax = std::max(kmax, std::max(key[a], key[b]));
      }
      const bool diag = P.row_idx[t] == P.col_idx[t];
      if (diag) cur += chol;
      else { const int64_t d = P.col_ptr[P.col_idx[t]]; cur = std::max(cur, finish[d] + lat) + fin; kmax = std::max(kmax, key[d]); }
      finish[t] = cur;
      const double work = (double)(P.upd_ptr[t + 1] - P.upd_ptr[t]) * g + (diag ? chol : fin);
      key[t] = std::max(cur - work, kmax);
    }
    for (int64_t w = 0; w < P.n_wtask; ++w) {
      const int64_t t = P.n_l + w;
      double cur = 0.0, kmax = 0.0;
      for (int64_t u = P.wupd_ptr[w]; u < P.wupd_ptr[w + 1]; ++u) {
        const int a = P.wupd_l[u], b = P.wupd_flag[u];
        cur = std::max(cur, std::max(finish[a], finish[b]) + lat) + g;
        kmax = std::max(kmax, std::max(key[a], key[b]));
      }
      const int64_t d = P.col_ptr[P.wt_row[w]];
      cur = std::max(cur, finish[d] + lat) + fin; kmax = std::max(kmax, key[d]);
      finish[t] = cur;
      key[t] = std::max(cur - ((double)(P.wupd_ptr[w + 1] - P.wupd_ptr[w]) * g + fin), kmax);
    }
    P.task_order.resize((size_t)nT);
    std::iota(P.task_order.begin(), P.task_order.end(), 0);
    if (!getenv("MM_TC_NO_SCHEDULE"))
      std::stable_sort(P.task_order.begin(), P.task_order.end(), [&](int x, int y) { return key[x] < key[y]; });
  }
  if (side_spawned) side_thread.join(); else side_rc = side_work();
  if (side_rc) return side_rc;
  return 0;
}

}  // namespace mm
