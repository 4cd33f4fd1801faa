"""Synthetic doc-word graphs in exactly the layout `Text2GraphTransformer.fit_transform` emits
(textgcn/lib/text2graph.py:162-193): word nodes [0, V), document nodes [V, V+D); edge order =
(a) word-word PMI edges as interleaved (i,j),(j,i) pairs in upper-triangle row-major order
(graphbuilder.pyx:181-192), (b) (doc+V, word) for every doc-word non-zero in doc-major order,
(c) (word, doc+V) in the same order; `edge_index` is the non-contiguous `.T` view of an (E, 2)
int64 tensor; edge_attr fp32; x = sparse identity; y = 0 on word rows; masks over doc rows.

The datasets of the reference are stripped from the mount and there is no network, so every
benchmark shape comes from here, seeded.  Word popularity is Zipf-Mandelbrot
and popularity rank is decoupled from node id by a random permutation (CountVectorizer orders
the vocabulary alphabetically, so hubs are scattered over the id range in real graphs too).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import warnings

import numpy as np
import torch

from .data import Data


@dataclass(frozen=True)
class GraphShape:
    name: str
    n_words: int
    n_docs: int
    ww_pairs: int           # undirected PMI pairs (directed edges = 2x)
    words_per_doc: float
    n_classes: int
    hidden: int
    dropout: float = 0.5
    amsgrad: bool = True
    lr: float = 0.05


# named shapes of SURVEY.md 8d / BASELINE.json configs
SHAPES = {
    # tiny shapes for CPU-side tests and smoke
    "tiny": GraphShape("tiny", 120, 80, 600, 12, 4, 16),
    "small": GraphShape("small", 1500, 1200, 30000, 30, 6, 64),
    # R8: 7,688 words + 7,674 docs, E ~ 3e6 directed
    "r8": GraphShape("r8", 7688, 7674, 1_250_000, 42, 8, 200),
    # 20NG: 42,757 words + 18,846 docs, E ~ 2.2e7 directed (the headline shape)
    "20ng": GraphShape("20ng", 42757, 18846, 9_100_000, 100, 20, 200),
    # flat_amazon.py: ~50k docs, V ~ 2e4, C = 64, hidden 100 (script) , dropout 0.7, AMSGrad
    "amazon": GraphShape("amazon", 20000, 50000, 3_000_000, 40, 64, 100, dropout=0.7),
    # perlevel_dbpedia.py: 337,739 docs, V ~ 1e4, max_length 15, hidden 32, plain Adam
    "dbpedia": GraphShape("dbpedia", 10000, 337739, 400_000, 12, 219, 32, amsgrad=False),
    # scaling sweep: 1M docs + 200k words, hidden 256
    "scale": GraphShape("scale", 200_000, 1_000_000, 20_000_000, 50, 20, 256),
}


def _zipf_probs(n: int, exponent: float, offset: float) -> np.ndarray:
    r = np.arange(1, n + 1, dtype=np.float64)
    p = (r + offset) ** (-exponent)
    return p / p.sum()


def _sample_ranks(rng: np.random.Generator, cdf: np.ndarray, size: int) -> np.ndarray:
    # inverse-CDF sampling; torch.searchsorted is multi-threaded (np.searchsorted is ~10x slower here)
    u = torch.from_numpy(rng.random(size))
    r = torch.searchsorted(torch.from_numpy(cdf), u, right=True).clamp_(0, cdf.size - 1)
    return r.numpy()


def _unique_sorted(keys: np.ndarray) -> np.ndarray:
    """Sorted unique values (sort + adjacent-difference; numpy 2.3's np.unique is ~40x slower)."""
    if keys.size == 0:
        return keys
    keys = np.sort(keys)
    keep = np.empty(keys.size, dtype=bool)
    keep[0] = True
    np.not_equal(keys[1:], keys[:-1], out=keep[1:])
    return keys[keep]


def make_edges(shape: GraphShape, seed: int = 0):
    """Returns (coo int64 [E,2], weights fp32 [E], n_ww_directed) in reference edge order."""
    rng = np.random.default_rng(seed)
    V, D = shape.n_words, shape.n_docs
    rank_to_id = rng.permutation(V).astype(np.int64)

    # ---- (a) word-word PMI pairs ----
    target = min(shape.ww_pairs, V * (V - 1) // 2)
    cdf = np.cumsum(_zipf_probs(V, 0.8, 10.0))
    keys = np.empty(0, dtype=np.int64)
    while keys.size < target:
        need = int((target - keys.size) * 1.35) + 1024
        a = rank_to_id[_sample_ranks(rng, cdf, need)]
        b = rank_to_id[_sample_ranks(rng, cdf, need)]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        ok = lo != hi
        keys = _unique_sorted(np.concatenate([keys, lo[ok] * V + hi[ok]]))
    if keys.size > target:
        keys = np.sort(rng.choice(keys, size=target, replace=False))
    wi, wj = keys // V, keys % V                       # sorted: upper-triangle row-major
    pmi = rng.uniform(1e-3, 6.0, size=keys.size).astype(np.float32)
    ww = np.empty((2 * keys.size, 2), dtype=np.int64)
    ww[0::2, 0], ww[0::2, 1] = wi, wj
    ww[1::2, 0], ww[1::2, 1] = wj, wi
    ww_w = np.repeat(pmi, 2)

    # ---- (b)/(c) doc-word TF-IDF-like edges ----
    k = np.clip(np.rint(rng.lognormal(np.log(shape.words_per_doc) - 0.125, 0.5, size=D)), 1, V).astype(np.int64)
    doc_of = np.repeat(np.arange(D, dtype=np.int64), k)
    cdf_dw = np.cumsum(_zipf_probs(V, 1.0, 2.0))
    words = rank_to_id[_sample_ranks(rng, cdf_dw, doc_of.size)]
    dk = _unique_sorted(doc_of * V + words)                 # doc-major, ascending word id (th.nonzero order)
    dd, dwrd = dk // V, dk % V
    raw = rng.uniform(0.1, 1.0, size=dk.size)
    norm = np.sqrt(np.bincount(dd, weights=raw * raw, minlength=D))
    tfidf = (raw / norm[dd]).astype(np.float32)        # L2-normalised rows, like TfidfTransformer()
    dw = np.stack([dd + V, dwrd], axis=1)
    wd = np.stack([dwrd, dd + V], axis=1)

    coo = np.concatenate([ww, dw, wd], axis=0)
    w = np.concatenate([ww_w, tfidf, tfidf]).astype(np.float32)
    return coo, w, int(ww.shape[0])


def make_graph(shape, seed: int = 0, hierarchy_classes: Optional[int] = None, sparse_x: bool = True) -> Data:
    """Build the synthetic `Data` for a named shape (or a GraphShape).  CPU tensors, like the
    reference's fit_transform; move with `.to('cuda')`."""
    if isinstance(shape, str):
        shape = SHAPES[shape]
    rng = np.random.default_rng(seed + 7919)
    V, D = shape.n_words, shape.n_docs
    N = V + D
    coo, w, n_ww = make_edges(shape, seed)
    coo_t = torch.from_numpy(coo)                      # (E, 2) int64
    y_docs = rng.integers(0, shape.n_classes, size=D)
    y = torch.zeros(N, dtype=torch.int64)
    y[V:] = torch.from_numpy(y_docs)
    u = rng.random(D)
    test_mask = torch.zeros(N, dtype=torch.bool)
    val_mask = torch.zeros(N, dtype=torch.bool)
    test_mask[V:] = torch.from_numpy(u >= 0.7)
    val_mask[V:] = torch.from_numpy((u >= 0.6) & (u < 0.7))
    train_mask = torch.logical_not(torch.logical_or(test_mask, val_mask))
    train_mask[:V] = False
    idx = torch.arange(N, dtype=torch.int64)
    inds = torch.stack([idx, idx])
    vals = torch.ones(N, dtype=torch.float32)
    n_cols = N
    if hierarchy_classes:
        parent = rng.integers(0, hierarchy_classes, size=D)
        inds = torch.cat([inds, torch.stack([torch.arange(D, dtype=torch.int64) + V,
                                             torch.from_numpy(parent).to(torch.int64) + N])], dim=1)
        vals = torch.cat([vals, torch.ones(D, dtype=torch.float32)])
        n_cols = N + hierarchy_classes
    warnings.filterwarnings("ignore", message="Sparse invariant checks")
    if sparse_x:
        x = torch.sparse_coo_tensor(inds, vals, size=(N, n_cols), dtype=torch.float32,
                                    check_invariants=False).coalesce()
    else:
        x = torch.sparse_coo_tensor(inds, vals, size=(N, n_cols), dtype=torch.float32,
                                    check_invariants=False).to_dense()
    g = Data(x=x, edge_index=coo_t.T, edge_attr=torch.from_numpy(w), y=y,
             test_mask=test_mask, train_mask=train_mask, val_mask=val_mask, n_vocab=V)
    return g
