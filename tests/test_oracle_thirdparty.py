"""Third-party anchor for the oracle's gcn_norm / GCNConv restatement (oracle/gcn_oracle.py).

torch_geometric 1.6.3 cannot be installed here, so the oracle cannot be pinned on PyG itself
("parity unpinned").  What can be pinned is the published arithmetic: networkx's
`normalized_laplacian_matrix` is an independent implementation of I - D^-1/2 A D^-1/2, so on a graph
with the self loops added, I - L is the A_hat of Kipf & Welling that GCNConv(add_self_loops=True)
(textgcn/lib/models.py:11-15) computes.  Fixtures: KarateClub with x = I (the reference's own model
fixture, textgcn/test/test_model.py:10-40) and a random weighted graph."""
import networkx as nx
import numpy as np
import torch

from helpers import karate_graph, random_graph
from oracle import gcn_oracle as O


def _ahat_networkx(ei, w, n):
    G = nx.Graph()
    G.add_nodes_from(range(n))
    for s, d, ww in zip(ei[0].tolist(), ei[1].tolist(), w.tolist()):
        G.add_edge(s, d, weight=ww)
    for i in range(n):
        G.add_edge(i, i, weight=1.0)              # add_remaining_self_loops(fill_value=1)
    L = nx.normalized_laplacian_matrix(G, nodelist=range(n), weight="weight").toarray()
    return np.eye(n) - L


def _ahat_oracle(ei, w, n):
    ei2, w_hat = O.gcn_norm(ei, w, n)
    M = np.zeros((n, n))
    np.add.at(M, (ei2[1].numpy(), ei2[0].numpy()), w_hat.double().numpy())    # row = target, col = source
    return M


def test_gcn_norm_matches_networkx_on_karate_and_weighted_graphs():
    g = karate_graph()
    n = g.x.shape[0]
    assert np.abs(_ahat_oracle(g.edge_index, g.edge_attr, n) - _ahat_networkx(g.edge_index, g.edge_attr, n)).max() < 1e-6
    ei, w = random_graph(80, 700, seed=4)
    assert np.abs(_ahat_oracle(ei, w, 80) - _ahat_networkx(ei, w, 80)).max() < 1e-6
    # the oracle's own independent fp64 formulation sits on the same matrix
    assert np.abs(O.dense_ahat_fp64(ei, w, 80).numpy() - _ahat_networkx(ei, w, 80)).max() < 1e-12


def test_two_layer_forward_matches_kipf_formula_with_networkx_ahat():
    """logits = A_hat (A_hat X W1 + b1) W2 + b2 -- the reference GCN in eval mode (no activation, models.py:22)."""
    g = karate_graph()
    n = g.x.shape[0]
    torch.manual_seed(0)
    gcn = O.OracleGCN(n, 4, n_hidden_gcn=16, dropout=0.5).eval()
    with torch.no_grad():
        for p in gcn.parameters():
            p.uniform_(-0.5, 0.5)
        z = gcn(g).double().numpy()
    A = _ahat_networkx(g.edge_index, g.edge_attr, n)
    W1, b1, W2, b2 = (p.detach().double().numpy() for p in gcn.parameters())
    ref = A @ ((A @ W1 + b1) @ W2) + b2
    assert np.abs(z - ref).max() / np.abs(ref).max() < 1e-5
