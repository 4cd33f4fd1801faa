"""Generates the committed golden vectors under tests/golden/ -- run in the build container, where
/root/reference is mounted (`make -C oracle` first builds the reference's own Cython graph builder
into oracle/_ref/).  Nothing on the GPU box reads /root/reference: the tests read these files.

  graphbuilder_*.npz   inputs + outputs of the REFERENCE builder itself
                       (textgcn/lib/clib/graphbuilder.pyx compute_word_word_edges / sliding_window_tester)
  gcn_karate.npz       logits / loss / gradients of the plain-torch restatement (oracle/gcn_oracle.py) on the
                       KarateClub fixture of textgcn/test/test_model.py -- pins the ORACLE against drift
                       (the reference's GCNConv cannot be imported here: torch_geometric is not installable)

NOTE (reference bug): graphbuilder.pyx:240-244 indexes its no-diagonal packed array from 1, so
edges_from_counts writes one float past its malloc (graphbuilder.pyx:134,158-166).  Vocabulary
sizes with V(V-1)/2 % 4 == 2 have no allocator slack there and abort with heap corruption; the
fixtures use other sizes.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden")


def tokens(rng, D, L, V, full_rows=1):
    lens = rng.integers(0, L + 1, size=D)
    X = np.full((D, L), -1, dtype=np.int32)
    z = rng.zipf(1.3, size=(D, L)) % V
    for d in range(D):
        X[d, :lens[d]] = z[d, :lens[d]]
    X[:full_rows, :] = rng.integers(0, V, size=(full_rows, L))
    return X


def main():
    from oracle import graphbuilder_oracle as GO
    from oracle import gcn_oracle as O
    os.makedirs(OUT, exist_ok=True)
    ref = GO.reference_module()
    if ref is None:
        raise SystemExit("oracle/_ref is not built (needs /root/reference): run `make -C oracle`")
    rng = np.random.default_rng(20261018)
    # the reference's own KAT input (textgcn/test/test_cfunc.py:83-86) + random corpora
    kat = np.array([[0, 1, 2, 3, 4, -1, -1, -1], [5, 3, 4, 1, 2, 0, 5, 1]], dtype=np.int32)
    cases = {"kat": (kat, 6, 3), "a": (tokens(rng, 40, 24, 30, 2), 30, 5), "b": (tokens(rng, 120, 48, 302, 2), 302, 20),
             "c": (tokens(rng, 60, 10, 50, 1), 50, 10), "d": (tokens(rng, 200, 16, 1000, 1), 1000, 4)}
    for name, (X, V, w) in cases.items():
        assert (V * (V - 1) // 2) % 4 != 2
        D, L = X.shape
        cij = np.asarray(ref.sliding_window_tester(X, V, D, L, window_size=w)).copy()
        coo, wt = ref.compute_word_word_edges(X, V, D, L, w)
        np.savez_compressed(os.path.join(OUT, f"graphbuilder_{name}.npz"), X=X, n_vocab=V, window=w, c_ij=cij,
                            coo=np.asarray(coo).copy(), weights=np.asarray(wt).copy())
        print(name, X.shape, "edges", np.asarray(coo).shape[0])

    # oracle self-pin on the KarateClub fixture
    from helpers import karate_graph
    g = karate_graph()
    torch.manual_seed(1234)
    gcn = O.OracleGCN(34, 4, n_hidden_gcn=64, dropout=0.5)
    with torch.no_grad():
        for l in gcn.layers:
            l.bias.uniform_(-0.2, 0.2)
    keep = torch.rand(34, 64) > 0.5
    gcn.train()
    z = gcn(g, drop_masks=[keep])
    loss = O.masked_cross_entropy(z, g.y, g.train_mask)
    loss.backward()
    rowptr, col, val, dis, _ = O.csr_from_gcn_norm(g.edge_index, g.edge_attr, 34)
    np.savez_compressed(os.path.join(OUT, "gcn_karate.npz"),
                        edge_index=g.edge_index.numpy(), edge_attr=g.edge_attr.numpy(), y=g.y.numpy(),
                        train_mask=g.train_mask.numpy(), keep=keep.numpy(),
                        W1=gcn.layers[0].weight.detach().numpy(), b1=gcn.layers[0].bias.detach().numpy(),
                        W2=gcn.layers[1].weight.detach().numpy(), b2=gcn.layers[1].bias.detach().numpy(),
                        logits=z.detach().numpy(), loss=loss.item(),
                        gW1=gcn.layers[0].weight.grad.numpy(), gb1=gcn.layers[0].bias.grad.numpy(),
                        gW2=gcn.layers[1].weight.grad.numpy(), gb2=gcn.layers[1].bias.grad.numpy(),
                        rowptr=rowptr.numpy(), colidx=col.numpy(), val=val.numpy(), dis=dis.numpy())
    print("gcn_karate loss", loss.item())


if __name__ == "__main__":
    main()
