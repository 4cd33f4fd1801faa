"""Shared graph fixtures for the parity tests (CPU tensors; the GPU tests move them)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pytextgcn_b200.data import Data  # noqa: E402


def karate_graph():
    """KarateClub with x = I_34, the reference's own model fixture (textgcn/test/test_model.py:10-40),
    rebuilt from networkx because torch_geometric.datasets is not installable here."""
    import networkx as nx
    G = nx.karate_club_graph()
    n = G.number_of_nodes()
    src, dst = [], []
    for u, v in G.edges():
        src += [u, v]
        dst += [v, u]
    ei = torch.tensor([src, dst], dtype=torch.int64)
    clubs = sorted({G.nodes[i]["club"] for i in range(n)})
    y = torch.tensor([clubs.index(G.nodes[i]["club"]) for i in range(n)], dtype=torch.int64)
    idx = torch.arange(n)
    x = torch.sparse_coo_tensor(torch.stack([idx, idx]), torch.ones(n), size=(n, n)).coalesce()
    train_mask = torch.zeros(n, dtype=torch.bool)
    train_mask[::3] = True
    return Data(x=x, edge_index=ei, edge_attr=torch.ones(ei.shape[1]), y=y, train_mask=train_mask,
                val_mask=~train_mask, test_mask=~train_mask, n_vocab=0)


def random_graph(n, e, seed=0, symmetric=True, self_loops=0, duplicates=0, transposed_view=False,
                 isolated=0, weight_range=(0.05, 4.0)):
    """Random weighted graph in COO form.  `isolated` nodes get no edges at all (their CSR row is
    just the self loop); `self_loops` pre-existing (i,i) edges exercise add_remaining_self_loops;
    `duplicates` repeats some edges (summed by scatter_add, kept as separate CSR entries)."""
    rng = np.random.default_rng(seed)
    m = n - isolated
    s = rng.integers(0, m, size=e)
    d = rng.integers(0, m, size=e)
    keep = s != d
    s, d = s[keep], d[keep]
    w = rng.uniform(*weight_range, size=s.size).astype(np.float32)
    if symmetric:
        key = np.minimum(s, d) * n + np.maximum(s, d)
        _, first = np.unique(key, return_index=True)
        s, d, w = s[first], d[first], w[first]
        s, d, w = np.concatenate([s, d]), np.concatenate([d, s]), np.concatenate([w, w])
    if duplicates:
        pick = rng.integers(0, s.size, size=duplicates)
        s, d, w = np.concatenate([s, s[pick]]), np.concatenate([d, d[pick]]), np.concatenate([w, w[pick]])
    if self_loops:
        l = rng.integers(0, m, size=self_loops)
        lw = rng.uniform(*weight_range, size=self_loops).astype(np.float32)
        pos = rng.integers(0, s.size + 1, size=self_loops)
        s, d, w = np.insert(s, pos, l), np.insert(d, pos, l), np.insert(w, pos, lw)
    coo = torch.from_numpy(np.stack([s, d], axis=1).astype(np.int64))
    ei = coo.T if transposed_view else coo.T.contiguous()
    return ei, torch.from_numpy(w.astype(np.float32))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  -- the relative error the 1e-5 (fp32) / 1e-2 (bf16) tolerances refer to."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)
