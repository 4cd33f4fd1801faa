"""Times the product's native word-word edge builder (pytextgcn_b200/csrc_host) beside the REFERENCE's own
builder -- graphbuilder.pyx compiled into oracle/_ref by oracle/Makefile -- on the same synthetic corpus, and checks that
the two outputs are bit-identical.  CPU only; lives under tests/ because it executes the oracle side.

    python tests/bench_graphbuilder.py [n_docs=10000] [seq_len=200] [n_vocab=20000] [window=20] > profiles/r02_graphbuilder_cpu.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import graphbuilder_oracle as GO  # noqa: E402
from pytextgcn_b200.graphbuilder import compute_word_word_edges  # noqa: E402


def main():
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    V = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
    w = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    if (V * (V - 1) // 2) % 4 == 2:
        raise SystemExit("pick another n_vocab: the reference builder writes one float past its allocation for this size "
                         "(graphbuilder.pyx:240-244, see DESIGN.md section 4)")
    rng = np.random.default_rng(0)
    lens = rng.integers(L // 4, L + 1, size=D)
    X = (rng.zipf(1.2, size=(D, L)) % V).astype(np.int32)
    X[np.arange(L)[None, :] >= lens[:, None]] = -1          # ragged documents, -1 padding (text2graph.py:115-126)
    cores = os.cpu_count() or 1
    out = {"corpus": {"n_docs": D, "seq_len": L, "n_vocab": V, "window": w, "tokens": int((X >= 0).sum())}, "cores": cores}

    def timed(fn, reps):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            r = fn()
            ts.append(time.perf_counter() - t0)
        return r, sorted(ts)[len(ts) // 2]
    got1, t1 = timed(lambda: compute_word_word_edges(X, V, D, L, w, n_jobs=1), 3)
    gotn, tn = timed(lambda: compute_word_word_edges(X, V, D, L, w, n_jobs=cores), 3)
    out["product"] = {"s_1_thread": t1, "s_all_cores": tn, "edges": int(got1[0].shape[0])}
    same_threads = np.array_equal(got1[0], gotn[0]) and np.array_equal(got1[1].view(np.int32), gotn[1].view(np.int32))
    ref = GO.reference_module()
    if ref is None:
        out["reference"] = {"unavailable": "oracle/_ref not built (needs /root/reference, see oracle/Makefile)"}
    else:
        r, tr = timed(lambda: ref.compute_word_word_edges(X, V, D, L, w), 3)
        rc, rw = np.asarray(r[0]), np.asarray(r[1])
        out["reference"] = {"s": tr, "edges": int(rc.shape[0]), "what": "graphbuilder.pyx of the reference, compiled unchanged into oracle/_ref"}
        out["bit_identical_to_reference"] = bool(rc.shape == got1[0].shape and np.array_equal(rc, got1[0]) and
                                                 np.array_equal(rw.view(np.int32), got1[1].view(np.int32)))
        out["speedup_1_thread"] = tr / t1
        out["speedup_all_cores"] = tr / tn
    out["bit_identical_across_thread_counts"] = bool(same_threads)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
