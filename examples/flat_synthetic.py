"""The reference's flat training script (flat_amazon.py:60-142) on a synthetic labelled corpus, using
the drop-in import surface: Text2GraphTransformer -> Data -> GCN -> the reference's epoch loop.

    python examples/flat_synthetic.py [--docs 3000] [--epochs 60] [--fast]

`--fast` swaps the loop body for TextGCNTrainer (fused loss / Adam / CUDA graphs); without it the loop is
the reference's, line for line, on torch.optim.Adam.  The corpus: every class has its own topic words
mixed with common words, so a working pipeline reaches high validation accuracy within a few epochs.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from textgcn import Text2GraphTransformer          # noqa: E402  (reference import path)
from textgcn.lib.models import GCN                  # noqa: E402


def make_corpus(n_docs, n_classes, seed=0, vocab_common=400, vocab_topic=60, doc_len=40):
    rng = np.random.default_rng(seed)
    common = [f"w{i}" for i in range(vocab_common)]
    topic = [[f"t{c}x{i}" for i in range(vocab_topic)] for c in range(n_classes)]
    y = rng.integers(0, n_classes, size=n_docs)
    docs = []
    for c in y:
        k = int(rng.integers(doc_len // 2, doc_len * 2))
        words = [topic[c][int(i)] if rng.random() < 0.35 else common[int(min(rng.zipf(1.3), vocab_common) - 1)]
                 for i in rng.integers(0, vocab_topic, size=k)]
        docs.append(" ".join(words))
    return docs, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=3000)
    ap.add_argument("--classes", type=int, default=8)
    ap.add_argument("--epochs", type=int, default=60)
    ap.add_argument("--hidden", type=int, default=100)
    ap.add_argument("--dropout", type=float, default=0.5)
    ap.add_argument("--lr", type=float, default=0.02)
    ap.add_argument("--fast", action="store_true")
    args = ap.parse_args()

    x, y = make_corpus(args.docs, args.classes)
    idx = np.random.default_rng(1).permutation(args.docs)
    test_idx, val_idx = idx[: args.docs // 5], idx[args.docs // 5: args.docs * 3 // 10]
    t0 = time.time()
    t2g = Text2GraphTransformer(n_jobs=8, min_df=2, save_path=None, verbose=0, max_df=0.9, window_size=10,
                                rm_stopwords=False)
    g = t2g.fit_transform(x, y, test_idx=test_idx, val_idx=val_idx)
    print(f"graph built in {time.time() - t0:.2f}s: {t2g.n_vocabs_} words + {t2g.n_docs_} docs, {g.edge_index.shape[1]} edges")

    gcn = GCN(g.x.shape[1], len(np.unique(y)), n_hidden_gcn=args.hidden, dropout=args.dropout)
    criterion = th.nn.CrossEntropyLoss(reduction="mean")
    device = th.device("cuda")
    gcn = gcn.to(device).float()
    g = g.to(device)
    t0 = time.time()
    if args.fast:
        from pytextgcn_b200.trainer import TextGCNTrainer
        tr = TextGCNTrainer(gcn, g, lr=args.lr, amsgrad=True)
        for epoch in range(args.epochs):
            st = tr.epoch()
            if epoch % 10 == 9 or epoch == args.epochs - 1:
                print(f"[{epoch + 1:3}] loss: {st['loss']: .3f}, training accuracy: {st['acc_train']: .3f}, val_acc: {st['acc_val']: .3f}")
        acc_val = st["acc_val"]
    else:
        optimizer = th.optim.Adam(gcn.parameters(), lr=args.lr, amsgrad=True)
        for epoch in range(args.epochs):                       # flat_amazon.py:99-117
            gcn.train()
            outputs = gcn(g)[g.train_mask]
            loss = criterion(outputs, g.y[g.train_mask])
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            optimizer.step()
            gcn.eval()
            with th.no_grad():
                logits = gcn(g)
                pred_val = np.argmax(logits[g.val_mask].cpu().numpy(), axis=1)
                pred_train = np.argmax(logits[g.train_mask].cpu().numpy(), axis=1)
                acc_val = float((pred_val == g.y.cpu()[g.val_mask.cpu()].numpy()).mean())
                acc_train = float((pred_train == g.y.cpu()[g.train_mask.cpu()].numpy()).mean())
            if epoch % 10 == 9 or epoch == args.epochs - 1:
                print(f"[{epoch + 1:3}] loss: {loss.item(): .3f}, training accuracy: {acc_train: .3f}, val_acc: {acc_val: .3f}")
    th.cuda.synchronize()
    print(f"Training took {time.time() - t0:.2f}s for {args.epochs} epochs; final val accuracy {acc_val:.3f}")
    with th.no_grad():
        gcn.eval()
        pred_test = np.argmax(gcn(g)[g.test_mask].cpu().numpy(), axis=1)
        acc_test = float((pred_test == g.y.cpu()[g.test_mask.cpu()].numpy()).mean())
    print(f"Test Accuracy: {acc_test: .3f}")
    return acc_val, acc_test


if __name__ == "__main__":
    main()
