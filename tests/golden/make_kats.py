"""Generates tests/golden/kats.json: hand-derived known-answer tests for the hot path.

The reference ships no golden vectors for this path, and it cannot be run here
(no R).  These KATs are derived BY HAND from the reference's formulas (src/BARTInt.R:32-72, 10-30,
165-173, 260) with closed-form arithmetic written out below -- independently of oracle/ -- so they pin
the oracle, which in turn checks the CUDA path.  Run:  python tests/golden/make_kats.py
"""
import json
import math
import os

import numpy as np


def Phi(x):
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def leaf():
    return dict(var=-1, cut=0, left=-1, right=-1)


def forest(trees, cuts, y_min=-0.5, y_max=0.5, S=1):
    """trees: list (len S*m) of pre-order node lists [(var, cut, left, right, mu)]"""
    off, var, cut, left, right, mu = [0], [], [], [], [], []
    for t in trees:
        for (v, c, l, r, u) in t:
            var.append(v); cut.append(c); left.append(l); right.append(r); mu.append(u)
        off.append(len(var))
    cut_offsets = [0]
    cut_values = []
    for c in cuts:
        cut_values += list(c)
        cut_offsets.append(len(cut_values))
    return dict(S=S, m=len(trees) // S, d=len(cuts), tree_offset=off, var=var, cut=cut, left=left, right=right, mu=mu,
                cut_offsets=cut_offsets, cut_values=cut_values, y_min=y_min, y_max=y_max)


L = lambda u: (-1, 0, -1, -1, u)
kats = []

# 1. stump: root is a leaf -> probability 1 (BARTInt.R:67-69), integral = mu under every measure
kats.append(dict(
    name="stump", forest=forest([[L(0.25)]], [[0.5]], y_min=2.0, y_max=6.0),
    integrals=dict(uniform=[0.25], gaussian=[0.25], exponential=[0.25]),
    X=[[0.1], [0.9]], leaf=[[0, 0]], tree_sum=[[0.25, 0.25]],
    predict=[[(6.0 - 2.0) * (0.25 + 0.5) + 2.0] * 2]))

# 2. one split at 0.5, mu = (-0.5, +0.5): uniform integral 0; rescaled mean (ymin+ymax)/2
c = [0.25, 0.5, 0.75]
kats.append(dict(
    name="single_split_half", forest=forest([[(0, 1, 1, 2, 0.0), L(-0.5), L(0.5)]], [c], y_min=1.0, y_max=3.0),
    integrals=dict(uniform=[0.5 * -0.5 + 0.5 * 0.5],
                   # gaussian as written (:49-52): left = (Phi(.5)-Phi(0)) / (Phi(.5)-Phi(-.5)), right = 1-left
                   gaussian=[((Phi(0.5) - Phi(0.0)) / (Phi(0.5) - Phi(-0.5))) * -0.5
                             + (1 - (Phi(0.5) - Phi(0.0)) / (Phi(0.5) - Phi(-0.5))) * 0.5],
                   # exponential (:53-56) on [0, Inf): left = pexp(.5), right = 1 - left
                   exponential=[(1 - math.exp(-0.5)) * -0.5 + math.exp(-0.5) * 0.5]),
    # routing rule: x <= cut -> left; x == cut goes LEFT, nextafter goes RIGHT, NaN compares false -> LEFT
    X=[[0.5], [float(np.nextafter(0.5, 1.0))], [-3.0], [7.0], [float("nan")], [0.49999]],
    leaf=[[0, 1, 0, 1, 0, 0]], tree_sum=[[-0.5, 0.5, -0.5, 0.5, -0.5, -0.5]],
    predict=[[1.0, 3.0, 1.0, 3.0, 1.0, 1.0]]))

# 3. nested splits on the same variable: root x<=0.5; left child x<=0.25.  P = .5*.5, .5*.5, .5
kats.append(dict(
    name="nested_same_var",
    forest=forest([[(0, 1, 1, 4, 0.0), (0, 0, 1, 2, 0.0)[:2] + (2, 3, 0.0), L(1.0), L(2.0), L(3.0)]], [c]),
    integrals=dict(uniform=[0.25 * 1.0 + 0.25 * 2.0 + 0.5 * 3.0]),
    X=[[0.1], [0.25], [0.3], [0.5], [0.6]], leaf=[[0, 0, 1, 1, 2]], tree_sum=[[1.0, 1.0, 2.0, 2.0, 3.0]],
    predict=[[1.0, 1.0, 2.0, 2.0, 3.0]]))   # y range [-0.5, 0.5]: (1)*(sum+0.5)-0.5 = sum

# 4. d = 2: root x0<=0.4 -> leaf 1; right: x1<=0.7 -> 2 | 3.  P = .4, .6*.7, .6*.3
kats.append(dict(
    name="two_dims",
    forest=forest([[(0, 0, 1, 2, 0.0), L(1.0), (1, 1, 3, 4, 0.0), L(2.0), L(3.0)]], [[0.4, 0.6], [0.2, 0.7]]),
    integrals=dict(uniform=[0.4 * 1.0 + (0.6 * 0.7) * 2.0 + (0.6 * 0.3) * 3.0]),
    X=[[0.4, 0.9], [0.41, 0.7], [0.41, 0.71], [0.0, 0.0]], leaf=[[0, 1, 2, 0]], tree_sum=[[1.0, 2.0, 3.0, 1.0]],
    predict=[[1.0, 2.0, 3.0, 1.0]]))

# 5. exponential measure, split at 1.0 then (right) at 2.0 on the same variable:
#    left = pexp(1); right child box [1, Inf): its left = (pexp(2)-pexp(1))/(1-pexp(1)), right = 1 - that
p1, p2 = 1 - math.exp(-1.0), 1 - math.exp(-2.0)
rl = (p2 - p1) / (1.0 - p1)
kats.append(dict(
    name="exponential_nested",
    forest=forest([[(0, 0, 1, 2, 0.0), L(1.0), (0, 1, 3, 4, 0.0), L(2.0), L(4.0)]], [[1.0, 2.0]]),
    integrals=dict(exponential=[p1 * 1.0 + ((1 - p1) * rl) * 2.0 + ((1 - p1) * (1 - rl)) * 4.0]),
    X=[[0.5], [1.0], [1.5], [2.0], [2.5]], leaf=[[0, 0, 1, 1, 2]], tree_sum=[[1.0, 1.0, 2.0, 2.0, 4.0]],
    predict=[[1.0, 1.0, 2.0, 2.0, 4.0]]))

# 6. two draws x two trees: sums over trees, per-draw integrals, variance over draws
#    draw 0: stump(0.1) + split@0.5(-0.2, 0.2); draw 1: stump(-0.1) + split@0.25(0.3, 0.1)
t_a = [(0, 1, 1, 2, 0.0), L(-0.2), L(0.2)]
t_b = [(0, 0, 1, 2, 0.0), L(0.3), L(0.1)]
f6 = forest([[L(0.1)], t_a, [L(-0.1)], t_b], [c], y_min=0.0, y_max=2.0, S=2)
ts = [[0.1 - 0.2, 0.1 - 0.2, 0.1 + 0.2], [-0.1 + 0.3, -0.1 + 0.1, -0.1 + 0.1]]   # X = .2, .4, .8
pred = [[2.0 * (v + 0.5) for v in row] for row in ts]
var = [((pred[0][i] - pred[1][i]) ** 2) / 2.0 for i in range(3)]   # n-1 = 1: var = (a-b)^2/2
kats.append(dict(
    name="two_draws_two_trees", forest=f6,
    integrals=dict(uniform=[0.1 + (0.5 * -0.2 + 0.5 * 0.2), -0.1 + (0.25 * 0.3 + 0.75 * 0.1)]),
    X=[[0.2], [0.4], [0.8]], leaf=[[0, 0, 0], [0, 0, 1], [0, 0, 0], [0, 1, 1]], tree_sum=ts, predict=pred,
    col_vars=var, row_means=[sum(r) / 3.0 for r in pred]))

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kats.json")
with open(out, "w") as fh:
    json.dump(kats, fh, indent=1)
print("wrote", out, len(kats), "KATs")
