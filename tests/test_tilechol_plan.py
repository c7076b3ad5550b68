"""Host-side symbolic analysis of the sparse tile Cholesky (csrc/tilechol_plan.h), checked without a GPU:
the plan's task lists are executed with numpy on a random block-sparse SPD matrix and the result must equal a
dense solve.  This pins the nested-dissection ordering, the tile structure (fill), the left-looking update lists,
the scatter maps and the substitution schedules; the CUDA kernels that walk the same lists are checked against
the same dense solve in tests/test_gpu_tilechol.py.
"""
import ctypes as C

import numpy as np
import pytest

from mavmap_b200 import _lib
from mavmap_b200._abi import as_ptr, p_f64, p_i32, p_i64


def plan_arrays(n_img, blk_a, blk_b, pos, ncb):
    L = _lib.lib()
    h = C.c_void_p()
    a = np.ascontiguousarray(blk_a, dtype=np.int32); b = np.ascontiguousarray(blk_b, dtype=np.int32)
    p = None if pos is None else np.ascontiguousarray(pos, dtype=np.float64)
    _lib.check(L.mm_debug_tilechol_plan_create(n_img, len(a), as_ptr(a, p_i32), as_ptr(b, p_i32), as_ptr(p, p_f64), ncb, C.byref(h)))
    names = ["scalars", "img_tile", "img_slot", "tile_nunk", "col_ptr", "row_idx", "col_idx", "has_a", "upd_ptr", "upd_a", "upd_b",
             "rowp_ptr", "rowp_tile", "rowp_col", "unk_of", "sc_tile", "sc_off", "tile_height", "tile_node", "node_first", "node_nt",
             "w_row_ptr", "wt_row", "wt_col", "wt_store", "wupd_ptr", "wupd_l", "wupd_w", "wupd_flag", "task_order",
             "st_kind", "st_out", "st_base", "st_tile", "st_item_ptr", "it_mat", "it_src"]
    out = {}
    for which, name in enumerate(names):
        n = L.mm_debug_tilechol_plan_array(h, which, None, 0)
        buf = np.zeros(max(n, 1), dtype=np.int64)
        L.mm_debug_tilechol_plan_array(h, which, as_ptr(buf, p_i64), n)
        out[name] = buf[:n]
    L.mm_debug_tilechol_plan_destroy(h)
    sc = out["scalars"]
    out.update(nt_pose=int(sc[0]), nt=int(sc[1]), n_l=int(sc[2]), n_upd=int(sc[3]), n_nodes=int(sc[4]), max_height=int(sc[5]), T=int(sc[7]), TI=int(sc[8]),
               n_w=int(sc[9]), n_wtask=int(sc[10]), n_slots=int(sc[11]), n_stasks=int(sc[12]), n_tnodes=int(sc[13]))
    return out


def grid_graph(strips, per, wk, ws, rng, drop=0.0):
    """image graph of a serpentine survey: images within +-wk along and +-ws across see common points"""
    n = strips * per
    idx = np.arange(n).reshape(strips, per)
    idx[1::2] = idx[1::2, ::-1]                         # serpentine numbering
    pos = np.zeros((n, 3))
    a_l, b_l = [], []
    for s in range(strips):
        for k in range(per):
            i = idx[s, k]; pos[i] = (k * 25.6, s * 38.4, 100.0)
            for ds in range(0, ws + 1):
                for dk in range(-wk, wk + 1):
                    if ds == 0 and dk <= 0:
                        continue
                    s2, k2 = s + ds, k + dk
                    if 0 <= s2 < strips and 0 <= k2 < per and rng.random() >= drop:
                        j = idx[s2, k2]
                        a_l.append(min(i, j)); b_l.append(max(i, j))
    return n, np.array(a_l, dtype=np.int32), np.array(b_l, dtype=np.int32), pos


def random_system(n, a, b, ncb, rng):
    """dense SPD matrix with the given 6x6 block pattern (+ dense 9*ncb border) and its block storage"""
    N = 6 * n + 9 * ncb
    S = np.zeros((N, N))
    blocks = np.zeros((n + len(a), 6, 6))
    for e, (i, j) in enumerate(zip(a, b)):
        B = rng.normal(0, 0.1, (6, 6))
        blocks[n + e] = B
        S[6 * i:6 * i + 6, 6 * j:6 * j + 6] = B; S[6 * j:6 * j + 6, 6 * i:6 * i + 6] = B.T
    Bm = rng.normal(0, 0.05, (n, ncb, 6, 9)) if ncb else None
    Cm = None
    if ncb:
        for i in range(n):
            for c in range(ncb):
                S[6 * i:6 * i + 6, 6 * n + 9 * c:6 * n + 9 * c + 9] = Bm[i, c]; S[6 * n + 9 * c:6 * n + 9 * c + 9, 6 * i:6 * i + 6] = Bm[i, c].T
        G = rng.normal(0, 0.1, (9 * ncb, 9 * ncb)); S[6 * n:, 6 * n:] = G + G.T
    # diagonally dominant -> SPD
    off = np.abs(S).sum(axis=1)
    for i in range(n):
        D = rng.normal(0, 0.1, (6, 6)); D = D + D.T + np.diag(off[6 * i:6 * i + 6] + 1.0)
        S[6 * i:6 * i + 6, 6 * i:6 * i + 6] = D; blocks[i] = D
    if ncb:
        S[6 * n:, 6 * n:] += np.diag(off[6 * n:] + 1.0)
        Cm = S[6 * n:, 6 * n:].copy()
    return S, blocks, Bm, Cm


def emulate(plan, n, blocks, Bm, Cm, rhs, ncb):
    """numpy execution of the plan: assembly -> left-looking factorisation -> forward / backward substitution"""
    T = plan["T"]; nt, nt_pose, n_l = plan["nt"], plan["nt_pose"], plan["n_l"]
    col_ptr, row_idx, col_idx = plan["col_ptr"], plan["row_idx"], plan["col_idx"]
    tiles = [None] * n_l
    for t in range(n_l):
        if plan["has_a"][t]:
            tiles[t] = np.zeros((T, T))
    for b in range(len(blocks)):
        off = int(plan["sc_off"][b]); tr = (off >> 30) & 1; base = off & ((1 << 30) - 1)
        r0, c0 = base % T, base // T
        tl = tiles[int(plan["sc_tile"][b])]
        assert tl is not None, "scatter into a tile that is not flagged has_a"
        tl[r0:r0 + 6, c0:c0 + 6] = blocks[b].T if tr else blocks[b]
    nb_t = nt - nt_pose
    per = T // 9
    for i in range(n):
        for c in range(ncb):
            j = int(plan["img_tile"][i]); t = int(col_ptr[j + 1]) - nb_t + c // per
            assert row_idx[t] == nt_pose + c // per
            tiles[t][9 * (c % per):9 * (c % per) + 9, 6 * int(plan["img_slot"][i]):6 * int(plan["img_slot"][i]) + 6] = Bm[i, c].T
    for br in range(nb_t):
        for bc in range(br + 1):
            t = int(col_ptr[nt_pose + bc]) + (br - bc)
            assert row_idx[t] == nt_pose + br and col_idx[t] == nt_pose + bc
            blk = Cm[9 * per * br:9 * per * (br + 1), 9 * per * bc:9 * per * (bc + 1)]
            tiles[t][:blk.shape[0], :blk.shape[1]] = blk
    # factorisation + node-block inverses, executed in the plan's static schedule (must be topological)
    n_w, n_wtask = plan["n_w"], plan["n_wtask"]
    Ltile = [None] * n_l; W = [None] * n_w
    w_row_ptr, tile_node, node_first = plan["w_row_ptr"], plan["tile_node"], plan["node_first"]
    widx = lambda i, j: int(w_row_ptr[i]) + (j - int(node_first[tile_node[i]]))
    done = np.zeros(n_l + n_wtask, dtype=bool)
    order = plan["task_order"]
    assert sorted(order.tolist()) == list(range(n_l + n_wtask))
    for t in order:
        t = int(t)
        if t < n_l:
            i, j = int(row_idx[t]), int(col_idx[t])
            Cc = tiles[t].copy() if plan["has_a"][t] else np.zeros((T, T))
            for u in range(int(plan["upd_ptr"][t]), int(plan["upd_ptr"][t + 1])):
                ta, tb = int(plan["upd_a"][u]), int(plan["upd_b"][u])
                assert done[ta] and done[tb], "schedule is not topological"
                assert row_idx[ta] == i and row_idx[tb] == j and col_idx[ta] == col_idx[tb]
                Cc -= Ltile[ta] @ Ltile[tb].T
            if i == j:
                nu = int(plan["tile_nunk"][j])
                Cc = np.tril(Cc) + np.tril(Cc, -1).T                       # the device reads the lower triangle only
                Cc[nu:, :] = 0; Cc[:, nu:] = 0; Cc[np.arange(nu, T), np.arange(nu, T)] = 1.0
                Ltile[t] = np.linalg.cholesky(Cc); W[widx(j, j)] = np.linalg.inv(Ltile[t])
            else:
                assert done[int(col_ptr[j])]
                Ltile[t] = Cc @ W[widx(j, j)].T
        else:
            w = t - n_l
            i, j = int(plan["wt_row"][w]), int(plan["wt_col"][w])
            assert tile_node[i] == tile_node[j] and i > j and int(plan["wt_store"][w]) == widx(i, j)
            acc = np.zeros((T, T))
            for u in range(int(plan["wupd_ptr"][w]), int(plan["wupd_ptr"][w + 1])):
                tl, wi, fl = int(plan["wupd_l"][u]), int(plan["wupd_w"][u]), int(plan["wupd_flag"][u])
                assert done[tl] and done[fl] and W[wi] is not None and row_idx[tl] == i
                acc += Ltile[tl] @ W[wi]
            assert done[int(col_ptr[i])]
            W[widx(i, j)] = -W[widx(i, i)] @ acc
        done[t] = True
    # substitutions through the task list (every task writes one slot; reads only slots written before)
    slots = [None] * plan["n_slots"]
    unk = plan["unk_of"].reshape(nt, T)
    out = np.zeros_like(rhs)
    ip = plan["st_item_ptr"]
    for k in range(plan["n_stasks"]):
        kind, o, base, tl = int(plan["st_kind"][k]), int(plan["st_out"][k]), int(plan["st_base"][k]), int(plan["st_tile"][k])
        items = range(int(ip[k]), int(ip[k + 1]))
        if kind == 2:
            v = np.where(unk[tl] >= 0, rhs[np.maximum(unk[tl], 0)], 0.0) if base < 0 else slots[base].copy()
            for q in items:
                v = v - slots[int(plan["it_src"][q])]
        else:
            v = np.zeros(T)
            for q in items:
                mat = int(plan["it_mat"][q]); sel, idx = mat >> 28, mat & ((1 << 28) - 1)
                src = slots[int(plan["it_src"][q])]
                assert src is not None, "substitution task reads a slot that is not written yet"
                M = Ltile[idx] if sel == 0 else W[idx]
                if kind == 1 or sel == 2:
                    v = v + M.T @ src
                else:
                    v = v + M @ src
            if kind == 3:
                m = unk[tl] >= 0
                out[unk[tl][m]] = v[m]
        assert slots[o] is None
        slots[o] = v
    return out


@pytest.mark.parametrize("strips,per,ncb,geom", [(4, 30, 0, True), (5, 24, 0, False), (3, 20, 1, True), (4, 16, 7, True), (1, 9, 2, True)])
def test_plan_solves_like_dense(strips, per, ncb, geom):
    rng = np.random.default_rng(100 * strips + per + ncb)
    n, a, b, pos = grid_graph(strips, per, 3, 1, rng, drop=0.15)
    S, blocks, Bm, Cm = random_system(n, a, b, ncb, rng)
    plan = plan_arrays(n, a, b, pos if geom else None, ncb)
    assert sorted(zip(plan["img_tile"], plan["img_slot"])) == sorted(set(zip(plan["img_tile"], plan["img_slot"])))     # a permutation
    assert plan["nt"] - plan["nt_pose"] == (ncb + 4) // 5
    rhs = rng.normal(size=S.shape[0])
    x = emulate(plan, n, blocks, Bm, Cm, rhs, ncb)
    ref = np.linalg.solve(S, rhs)
    assert np.abs(x - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())


def test_plan_disconnected_and_dense_graphs():
    rng = np.random.default_rng(3)
    # two disconnected strips + an isolated image
    n1, a1, b1, p1 = grid_graph(1, 40, 2, 0, rng)
    n2, a2, b2, p2 = grid_graph(2, 15, 2, 1, rng)
    n = n1 + n2 + 1
    a = np.concatenate([a1, a2 + n1]); b = np.concatenate([b1, b2 + n1])
    pos = np.concatenate([p1, p2 + [0, 500, 0], [[1e3, 1e3, 0]]])
    S, blocks, _, _ = random_system(n, a, b, 0, rng)
    plan = plan_arrays(n, a, b, pos, 0)
    rhs = rng.normal(size=6 * n)
    assert np.abs(emulate(plan, n, blocks, None, None, rhs, 0) - np.linalg.solve(S, rhs)).max() < 1e-10
    # complete graph (every image sees every other one): no separator exists, one dense node
    n = 30
    a, b = np.triu_indices(n, 1)
    S, blocks, _, _ = random_system(n, a.astype(np.int32), b.astype(np.int32), 0, rng)
    plan = plan_arrays(n, a, b, None, 0)
    rhs = rng.normal(size=6 * n)
    assert np.abs(emulate(plan, n, blocks, None, None, rhs, 0) - np.linalg.solve(S, rhs)).max() < 1e-10
    assert plan["n_l"] == plan["nt"] * (plan["nt"] + 1) // 2


def test_plan_fill_of_a_survey_block():
    """the dissection keeps the work far below a banded factorisation on a survey-shaped graph (20 strips x 50, +-5 x +-2)"""
    rng = np.random.default_rng(7)
    n, a, b, pos = grid_graph(20, 50, 5, 2, rng)
    plan = plan_arrays(n, a, b, pos, 0)
    T = plan["T"]
    flops = 2.0 * T ** 3 * (plan["n_upd"] + plan["n_l"])
    banded = (6.0 * n) * (6.0 * 2 * 50 * 2) ** 2          # n * bandwidth^2, bandwidth = two strips of images
    assert flops < 0.5 * banded, (flops, banded)
    # heights are a topological order: every tile's column structure lies in tiles of greater or equal height
    h = plan["tile_height"]
    assert np.all(h[plan["row_idx"]] >= h[plan["col_idx"]])
