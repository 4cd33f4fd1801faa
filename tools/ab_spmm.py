"""A/B of the wide propagation: tgcn_spmm (per-non-zero L2 gathers) vs tgcn_spmm_staged (operand rows of a
panel staged in shared memory) on a named synthetic shape.  Checks the two agree, then times each
configuration with CUDA events (L2 flushed between launches) and prints one JSON line per configuration.

    python tools/ab_spmm.py [shape=20ng] [reps=20]            # on the B200 box (gpurun)
    ncu --set full -k regex:k_spmm -c 4 python tools/ab_spmm.py 20ng 1 --only staged:28,1,64,4,0

Staged configurations are `W,RPW,KC,NP,MODE` = consumer warps per panel, chunks per consumer warp, operand
rows per stage, producer warps, producer mode (0 = cp.async.bulk per row, 1 = 16-byte cp.async)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytextgcn_b200 import make_graph, ops  # noqa: E402
from pytextgcn_b200.graph import upload_graph  # noqa: E402
from pytextgcn_b200.synthetic import SHAPES  # noqa: E402

DEFAULT_CONFIGS = ["28,1,64,4,0", "28,1,64,4,1", "28,2,64,4,0", "28,2,64,4,1", "30,1,64,2,0", "30,1,96,2,0",
                   "28,1,32,4,0", "28,1,128,4,0", "24,1,64,4,0", "16,1,64,4,0", "28,2,32,4,0"]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    shape = args[0] if args else "20ng"
    reps = int(args[1]) if len(args) > 1 else 20
    only = None
    if "--only" in sys.argv:
        only = sys.argv[sys.argv.index("--only") + 1]
    dev = torch.device("cuda")
    g = make_graph(shape)
    n = g.x.shape[0]
    gr = upload_graph(g.edge_index.T.contiguous().to(dev).T, g.edge_attr.to(dev), n)
    F = SHAPES[shape].hidden
    B = torch.randn(n, F, device=dev)
    bias = torch.randn(F, device=dev)
    out = torch.empty(n, F, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    alg_bytes = gr.nnz * 8 + (n + 1) * 4 + 2 * n * F * 4 + F * 4

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    ref = None
    if only is None or only == "default":
        med, best = timed(lambda: ops.spmm(gr, B, out=out, bias=bias, staged=False))
        ref = out.clone()
        print(json.dumps(dict(kernel="tgcn_spmm", shape=shape, F=F, nnz=gr.nnz, ms_median=med, ms_best=best,
                              algorithmic_gbs=alg_bytes / med / 1e6, gathered_rows=gr.nnz)), flush=True)
    cfgs = DEFAULT_CONFIGS if only is None else ([] if only == "default" else [only.split(":", 1)[1]])
    for c in cfgs:
        W, RPW, KC, NP, MODE = (int(x) for x in c.split(","))
        ops.STAGED_CFG.update(warps_per_panel=W, rows_per_warp=RPW, tile_cols=KC, n_producers=NP, producer_mode=MODE)
        try:
            sp = gr.staged_plan(gr.plan(), W, RPW, KC)
            med, best = timed(lambda: ops.spmm(gr, B, out=out, bias=bias, staged=True))
        except RuntimeError as e:
            print(json.dumps(dict(kernel="tgcn_spmm_staged", config=c, error=str(e)[:200])), flush=True)
            continue
        err = None
        if ref is not None:
            err = float((out - ref).abs().max() / ref.abs().max())
        print(json.dumps(dict(kernel="tgcn_spmm_staged", config=c, shape=shape, F=F, ms_median=med, ms_best=best,
                              algorithmic_gbs=alg_bytes / med / 1e6, gathered_rows=sp.gathered_rows(),
                              gather_ratio=sp.gathered_rows() / gr.nnz, plan_mb=sp.bytes() / 1e6,
                              max_rel_diff_vs_default=err)), flush=True)


if __name__ == "__main__":
    main()
