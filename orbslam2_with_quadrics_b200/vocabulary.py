"""Synthetic ORB vocabularies of DBoW2's shape (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) for Frame::ComputeBoW parity
tests and timing.  The reference checkout does not ship Vocabulary/ORBvoc.txt, so trees with its parameters (branching
k = 10, depth L = 6, TF-IDF weights, L1 scoring) are generated from a seed: random 256-bit node descriptors, idf-like
positive leaf weights.  `irregular=True` adds what trained vocabularies also contain: nodes with fewer than k children,
leaves above depth L and stopped words (weight 0).

Flat layout (also what orbx_vocabulary_create takes): node i has children child_items[child_start[i] : child_start[i+1]]
in DBoW2's order, a 32-byte descriptor, a weight, and -- for leaves -- a word id (word ids ascend with node ids, as
TemplatedVocabulary::createWords assigns them)."""
from __future__ import annotations

import numpy as np


def random_vocabulary(k: int = 10, L: int = 6, seed: int = 0, irregular: bool = False) -> dict:
    rng = np.random.default_rng(seed)
    if not irregular:
        level_start = [0]
        for d in range(L + 1):
            level_start.append(level_start[-1] + k ** d)
        n = level_start[-1]
        child_count = np.zeros(n, np.int64)
        child_count[:level_start[L]] = k
        child_start = np.zeros(n + 1, np.int64)
        np.cumsum(child_count, out=child_start[1:])
        child_items = np.arange(1, n, dtype=np.int32)              # BFS ids: the children of node i are 1 + k*i .. k + k*i
    else:
        children, depth = [[]], [0]
        frontier = [0]
        while frontier:
            nxt = []
            for i in frontier:
                if depth[i] >= L or (depth[i] >= 2 and rng.random() < 0.08):
                    continue
                for _ in range(int(rng.integers(2, k + 1))):
                    children.append([])
                    depth.append(depth[i] + 1)
                    children[i].append(len(children) - 1)
                    nxt.append(len(children) - 1)
            frontier = nxt
        n = len(children)
        child_start = np.zeros(n + 1, np.int64)
        np.cumsum([len(c) for c in children], out=child_start[1:])
        child_items = np.array([c for cs in children for c in cs], np.int32)
    is_leaf = (child_start[1:] - child_start[:-1]) == 0
    node_word = np.full(n, -1, np.int32)
    node_word[is_leaf] = np.arange(int(is_leaf.sum()), dtype=np.int32)
    node_weight = np.zeros(n, np.float64)
    node_weight[is_leaf] = np.log(rng.uniform(1.5, 4000.0, int(is_leaf.sum())))       # idf = log(N / Ni)
    if irregular:
        stop = is_leaf & (rng.random(n) < 0.05)
        node_weight[stop] = 0.0
    node_desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    return dict(k=k, L=L, n_nodes=n, child_start=child_start.astype(np.int32), child_items=child_items, node_desc=node_desc,
                node_weight=node_weight, node_word=node_word)
