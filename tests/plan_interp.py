"""Test infrastructure: a numpy interpreter of the group records the host planner produces.

`debug_plan(ff)` runs the library's real planner on the host (bartint_debug_plan_soa, no device) and returns
what forest creation would upload.  `interpret(plan, codes, sums_only)` then does, value by value, what the device
does with it -- build_tables_kernel (kernels_misc.cu), dual_offsets / gather_offset / code_of
(kernels_gather.cu), pred_mask4 / table_index (kernels_eval.cu) and the record-flag logic of the evaluation loops --
so the planner's bin packing, byte selectors, biases, slot weights and flags are checked against the oracle
without a GPU.  It restates the kernels' integer arithmetic; it does not replace the GPU parity tests.
"""
from __future__ import annotations

import numpy as np

import hosthooks

FLAG_END_DRAW, FLAG_WALK, FLAG_JOINT, FLAG_COUNTABLE, FLAG_END_EVAL = 1, 2, 4, 8, 16
LEAF_VAR = 0xFFFF


def debug_plan(ff) -> dict:
    """The library's own planner run on the host (bartint_debug_plan_soa through bartint_b200.hosthooks)."""
    return hosthooks.plan_soa(ff)


# ---- the PTX pieces -------------------------------------------------------------------------------------------

def byte_perm(a, b, sel):
    """__byte_perm(a, b, sel): output byte i = byte (nibble i of sel) of the 8 bytes {a: 0..3, b: 4..7}."""
    a = np.asarray(a, np.uint64)
    b = np.asarray(b, np.uint64)
    pool = a | (b << np.uint64(32))
    out = np.zeros_like(pool)
    for i in range(4):
        nib = (int(sel) >> (4 * i)) & 0xF
        assert nib < 8, "planner emitted a selector nibble with the sign-replicate bit set"
        out |= ((pool >> np.uint64(8 * nib)) & np.uint64(0xFF)) << np.uint64(8 * i)
    return out.astype(np.uint32)


def dp4a_sign(v, w, acc):
    """__dp4a(sign_mask4(v), w, acc): every byte of v with bit 7 set contributes -(signed byte of w)."""
    v = np.asarray(v, np.uint32)
    o = np.full(v.shape, acc, np.int64)
    for i in range(4):
        wb = (int(w) >> (8 * i)) & 0xFF
        wb = wb - 256 if wb >= 128 else wb
        o += np.where((v >> np.uint32(8 * i + 7)) & np.uint32(1), -wb, 0)
    return o


def add32(a, b):
    return ((np.asarray(a, np.uint64) + np.uint64(int(b))) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def pack_words(codes, regs=4):
    """codes [N, d] -> point-major packed words [N, 4] (byte j of the 16 = code of variable j), as K0 writes them."""
    N, d = codes.shape
    w = np.zeros((N, 4), np.uint32)
    for j in range(d):
        w[:, j >> 2] |= codes[:, j].astype(np.uint32) << np.uint32(8 * (j & 3))
    return w


# ---- build_tables_kernel --------------------------------------------------------------------------------------

def _tree_value_for_idx(plan, k, idx):
    """Leaf value of tree k selected by predicate bits idx (right at a node iff bit node_bit[node] of idx)."""
    node = int(plan["tree_off"][k])
    while (int(plan["nodes_x"][node]) & 0xFFFF) != LEAF_VAR:
        bit = int(plan["node_bit"][node])
        node += int(plan["nodes_y"][node]) if (idx >> bit) & 1 else 1
    return plan["leaf_mu"][int(plan["leaf_off"][k]) + int(plan["nodes_y"][node])]


def build_table(plan, k0, k1, n_entries):
    tbl = np.zeros(n_entries)
    for idx in range(n_entries):
        s = 0.0
        for kp in range(k0, k1):
            s += _tree_value_for_idx(plan, int(plan["tree_perm"][kp]), idx)
        tbl[idx] = s
    return tbl


def walk_tree(plan, k, codes):
    """WALK record / walk kernel: per point node walk on codes (right iff code > cut index)."""
    N = codes.shape[0]
    base = int(plan["tree_off"][k])
    node = np.full(N, base, np.int64)
    for _ in range(int(plan["tree_off"][k + 1]) - base):
        x = plan["nodes_x"][node].astype(np.int64)
        internal = (x & 0xFFFF) != LEAF_VAR
        if not internal.any():
            break
        v = np.where(internal, x & 0xFFFF, 0)
        go_right = codes[np.arange(N), v] > (x >> 16)
        node = np.where(internal, node + np.where(go_right, plan["nodes_y"][node].astype(np.int64), 1), node)
    return plan["leaf_mu"][int(plan["leaf_off"][k]) + plan["nodes_y"][node].astype(np.int64)]


# ---- the evaluation loop --------------------------------------------------------------------------------------

def record_values(plan, g, codes, words):
    """Sum over the record's trees of the leaf value of every point, formed the way the kernels form it."""
    flags, tb, te = int(plan["hdr"][g]), int(plan["group_tree_off"][g]), int(plan["group_tree_off"][g + 1])
    if flags & FLAG_WALK:
        assert te - tb == 1
        return walk_tree(plan, int(plan["tree_perm"][tb]), codes)
    dsc = [int(v) for v in plan["desc"][g]]
    R, nb = plan["regs"], plan["nb"]
    if not plan["gather"]:
        idx = np.zeros(codes.shape[0], np.int64)
        for j in range(nb):
            var, bias = dsc[2 * j], dsc[2 * j + 1] & 0xFF
            assert dsc[2 * j + 1] == bias * 0x01010101
            idx |= ((codes[:, var].astype(np.int64) + bias) >= 128).astype(np.int64) << j
        return build_table(plan, tb, te, 1 << nb)[idx]
    c = [words[:, r] for r in range(4)]
    d0, d1, d2, d3, d4, kind, sel_t = dsc[:7]
    if flags & FLAG_JOINT:
        gA = byte_perm(c[0], c[1], d0 & 0xFFFF)
        gB = byte_perm(c[R - 2], c[R - 1], d0 >> 16)
        o = dp4a_sign(add32(gB, d2), d4, dp4a_sign(add32(gA, d1), d3, 0))
        assert (o % 8 == 0).all() and (o >= 0).all() and (o < (8 << nb)).all()
        return build_table(plan, tb, te, 1 << nb)[o // 8]
    sp = int(plan["split"][g])
    assert 0 <= sp <= te - tb
    assert kind in (0, 1, 2)
    lo = c[0] if (R == 2 or kind == 1) else c[R - 2] if kind == 0 else byte_perm(c[0], c[1], sel_t & 0xFFFF)
    gA = byte_perm(c[0], c[1], d0 & 0xFFFF)
    gB = byte_perm(lo, c[R - 1], d0 >> 16)
    o1 = dp4a_sign(add32(gA, d1), d3, 0)
    o2 = dp4a_sign(add32(gB, d2), d4, 128)
    assert (o1 % 8 == 0).all() and (o1 >= 0).all() and (o1 < 128).all()
    assert (o2 % 8 == 0).all() and (o2 >= 128).all() and (o2 < 256).all()
    t1 = build_table(plan, tb, tb + sp, 16)
    t2 = build_table(plan, tb + sp, te, 16)
    return t1[o1 // 8] + t2[(o2 - 128) // 8]


def interpret(plan, codes, sums_only: bool):
    """(S, N) per-draw tree sums as the evaluation kernel accumulates them.  sums_only on a counting-mode plan skips
    the COUNTABLE records and closes the draw at END_EVAL (the skipped trees are returned as a list per draw)."""
    codes = np.asarray(codes, np.int64)
    words = pack_words(codes) if plan["gather"] else None      # point-major words exist for d <= 16 only
    S, N = plan["S"], codes.shape[0]
    out = np.zeros((S, N))
    skipped = [[] for _ in range(S)]
    counting = sums_only and plan["count_level"] > 0
    for s in range(S):
        g0, g1 = int(plan["draw_group_off"][s]), int(plan["draw_group_off"][s + 1])
        assert g1 > g0, "every draw needs at least one record"
        closed = 0
        for g in range(g0, g1):
            flags = int(plan["hdr"][g])
            if counting and flags & FLAG_COUNTABLE:
                assert closed == 1, "COUNTABLE records must follow the END_EVAL record"
                tb, te = int(plan["group_tree_off"][g]), int(plan["group_tree_off"][g + 1])
                skipped[s] += [int(plan["tree_perm"][kp]) for kp in range(tb, te)]
            else:
                assert closed == 0, "an evaluated record after the draw was closed"
                out[s] += record_values(plan, g, codes, words)
            if flags & (FLAG_END_EVAL if counting else FLAG_END_DRAW):
                closed += 1
        assert closed == 1, "a draw must be closed exactly once"
        assert int(plan["hdr"][g1 - 1]) & FLAG_END_DRAW
    return out, skipped


def check_structure(plan):
    """Invariants every plan must satisfy whatever the format."""
    S, m = plan["S"], plan["m"]
    perm = plan["tree_perm"][:S * m]
    for s in range(S):
        assert sorted(perm[s * m:(s + 1) * m]) == list(range(s * m, (s + 1) * m)), "tree_perm must permute within a draw"
    gto, dgo = plan["group_tree_off"], plan["draw_group_off"]
    assert (np.diff(gto) >= 0).all() and int(gto[-1]) == S * m
    assert int(dgo[0]) == 0 and int(dgo[-1]) == plan["n_groups"]
    for s in range(S):
        assert int(gto[dgo[s]]) == s * m, "groups never span draws"


# ---- counting mode: hist_codes / suffix_counts / hist2 / suffix2 / count_draw_sums_kernel ------------------------

def dual_able(words, R):
    """tg_dual_able (common.cuh)."""
    last = 1 << (R - 1)
    return (words & ~3) == 0 or (words & ~((1 << (R - 2)) | last)) == 0 or (words & ~(3 | last)) == 0


def count_sums(plan, codes):
    """Per-draw sums over the points of the trees the counting kernel handles, from cumulative code histograms only
    (never touching a point again), with the kernel's own rule for which trees those are.  Returns (sums [S], trees
    counted per draw)."""
    codes = np.asarray(codes, np.int64)
    N, d = codes.shape
    level, R = plan["count_level"], plan["regs"]
    hist = np.zeros((d, 128), np.int64)
    for v in range(d):
        hist[v] = np.bincount(codes[:, v] & 127, minlength=128)
    G = np.zeros((d, 128), np.int64)                      # G[v][c] = #{code[v] > c}
    for v in range(d):
        G[v] = hist[v][::-1].cumsum()[::-1] - hist[v]
    J2 = {}
    if level >= 2:
        for a in range(d):
            for b in range(a + 1, d):
                h2 = np.zeros((128, 128), np.int64)
                np.add.at(h2, (codes[:, a] & 127, codes[:, b] & 127), 1)
                e = h2[:, ::-1].cumsum(axis=1)[:, ::-1] - h2                   # along y: sum over y' > y
                J2[(a, b)] = e[::-1].cumsum(axis=0)[::-1] - e                  # along x: sum over x' > x
    nx, ny, toff, loff, mu = plan["nodes_x"], plan["nodes_y"], plan["tree_off"], plan["leaf_off"], plan["leaf_mu"]
    S, m = plan["S"], plan["m"]
    sums, counted = np.zeros(S), [[] for _ in range(S)]
    for s in range(S):
        for t in range(m):
            k = s * m + t
            base, cnt, lb = int(toff[k]), int(toff[k + 1] - toff[k]), int(loff[k])
            if cnt == 3:
                v, c = int(nx[base]) & 0xFFFF, int(nx[base]) >> 16
                g = int(G[v][c])
                sums[s] += mu[lb] * float(N - g) + mu[lb + 1] * float(g)
                counted[s].append(k)
            elif cnt == 5 and level >= 2:
                child_left = (int(nx[base + 1]) & 0xFFFF) != LEAF_VAR
                ch = base + 1 if child_left else base + int(ny[base])
                v1, c1, v2, c2 = int(nx[base]) & 0xFFFF, int(nx[base]) >> 16, int(nx[ch]) & 0xFFFF, int(nx[ch]) >> 16
                if not dual_able((1 << (v1 >> 2)) | (1 << (v2 >> 2)), R):
                    continue
                g1, g2 = int(G[v1][c1]), int(G[v2][c2])
                if v1 == v2:
                    j = int(G[v1][max(c1, c2)])
                elif v1 < v2:
                    j = int(J2[(v1, v2)][c1][c2])
                else:
                    j = int(J2[(v2, v1)][c2][c1])
                if child_left:
                    n_lr = g2 - j
                    sums[s] += mu[lb] * float(N - g1 - n_lr) + mu[lb + 1] * float(n_lr) + mu[lb + 2] * float(g1)
                else:
                    sums[s] += mu[lb] * float(N - g1) + mu[lb + 1] * float(g1 - j) + mu[lb + 2] * float(j)
                counted[s].append(k)
    return sums, counted
