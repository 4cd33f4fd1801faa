import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from pytextgcn_b200 import make_graph, ops
from pytextgcn_b200.graph import upload_graph
from pytextgcn_b200.dist import RowPartition, shard_graph
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2]
g = make_graph("20ng"); N = g.x.shape[0]; dev = torch.device("cuda")
full = upload_graph(g.edge_index.T.contiguous().to(dev).T, g.edge_attr.to(dev), N)
B = torch.randn(N, 200, device=dev); out = torch.empty(N, 200, device=dev)
W2 = torch.randn(200, 20, device=dev); b1 = torch.randn(200, device=dev); P = torch.zeros(N, 20, device=dev)
Pn = torch.randn(N, 20, device=dev); outn = torch.empty(N, 20, device=dev)
plan = full.plan()
t1 = timeit(lambda: ops.spmm(full, B, plan=plan, out=out))
t2 = timeit(lambda: ops.spmm(full, B, plan=plan, out=out, bias=b1, drop_mode=ops.DROP_PHILOX, drop_p=0.5, philox_seed=1, W_proj=W2, P=P))
t3 = timeit(lambda: ops.spmm(full, Pn, plan=plan, out=outn))
row_nnz = (full.rowptr[1:] - full.rowptr[:-1]).long()
part = RowPartition(row_nnz, 8); sh = shard_graph(full, part, 0)
Bp = torch.randn(part.n_pad, 200, device=dev); o8 = torch.empty(part.n_loc, 200, device=dev)
t4 = timeit(lambda: ops.spmm(sh, Bp, plan=sh.plan(), out=o8))
print(f"RESULT plain {t1*1e3:.1f} fused {t2*1e3:.1f} narrow {t3*1e3:.1f} shard8 {t4*1e3:.1f} us", flush=True)
