"""Hybrid (tensor-core + gather) propagation against the gather kernel and an fp64 host reference, with timings.
    python tools/tc_check.py [shape=small] [F=200] [min_density=0.03] [reps=10]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pytextgcn_b200 import make_graph, ops  # noqa: E402
from pytextgcn_b200.graph import upload_graph  # noqa: E402
from pytextgcn_b200.tc_plan import build_tc_plan  # noqa: E402


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "small"
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    dens = float(sys.argv[3]) if len(sys.argv) > 3 else 0.03
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    dev = torch.device("cuda")
    g = make_graph(shape)
    n = int(g.x.shape[0])
    gr = upload_graph(g.edge_index.T.contiguous().to(dev).T, g.edge_attr.to(dev), n)
    t0 = time.perf_counter()
    tc = build_tc_plan(gr, min_density=dens, n_sms=torch.cuda.get_device_properties(dev).multi_processor_count, width=F)
    torch.cuda.synchronize()
    info = {"shape": shape, "F": F, "min_density": dens, "plan_s": time.perf_counter() - t0, "nnz": gr.nnz}
    if tc is None:
        print(json.dumps(dict(info, note="no dense tiles")))
        return
    info.update(n_tiles=tc.n_tiles, n_slots=tc.n_slots, n_units=tc.n_units, nnz_dense=tc.nnz_dense, nnz_remainder=tc.remainder.nnz,
                a_tiles_mb=tc.A_tiles.numel() * 4 / 1e6)
    print(json.dumps(info), flush=True)
    torch.manual_seed(0)
    B = torch.randn(n, F, device=dev)
    bias = torch.randn(F, device=dev)
    ref, _ = ops.spmm(gr, B, bias=bias)
    torch.cuda.synchronize()
    out, _ = ops.spmm_hybrid(tc, B, bias=bias, plan=tc.remainder.plan())
    torch.cuda.synchronize()
    # fp64 on the device (torch sparse), the yardstick
    rows = gr.row_ids()
    A64 = torch.sparse_coo_tensor(torch.stack([rows, gr.colidx.long()]), gr.val.double(), size=(n, n)).coalesce()
    z64 = torch.sparse.mm(A64, B.double()) + bias.double()
    den = float(z64.abs().max())
    res = {"hybrid_vs_fp64": float((out.double() - z64).abs().max()) / den, "gather_vs_fp64": float((ref.double() - z64).abs().max()) / den,
           "hybrid_vs_gather": float((out - ref).abs().max()) / den}
    out2, _ = ops.spmm_hybrid(tc, B, bias=bias, plan=tc.remainder.plan())
    res["deterministic"] = bool(torch.equal(out, out2))
    print(json.dumps(res), flush=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib_plan = tc.remainder.plan()

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]
    from pytextgcn_b200 import _native
    import ctypes as C
    lib = _native.load()
    bt, part = tc.buffers(F)
    cp = tc.c_struct()

    def dense_only():
        _native.check(lib.tgcn_spmm_tc(C.byref(cp), B.data_ptr(), B.stride(0), F, bt.data_ptr(), part.data_ptr(), part.stride(0),
                                       torch.cuda.current_stream().cuda_stream))
    t = {"gather_ms": timed(lambda: ops.spmm(gr, B, bias=bias, out=ref)),
         "hybrid_ms": timed(lambda: ops.spmm_hybrid(tc, B, bias=bias, plan=lib_plan, out=out)),
         "dense_part_ms": timed(dense_only),
         "remainder_part_ms": timed(lambda: ops.spmm(tc.remainder, B, bias=bias, plan=lib_plan, out=out, tc=tc))}
    print(json.dumps(t), flush=True)


if __name__ == "__main__":
    main()
