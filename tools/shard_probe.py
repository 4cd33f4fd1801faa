"""1-GPU emulation of a rank's shard SpMM (rank 0 of `world`), chunk-size sweep."""
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
row_nnz = (full.rowptr[1:] - full.rowptr[:-1]).long()
for world in (1, 8):
    part = RowPartition(row_nnz, world)
    sh = shard_graph(full, part, 0)
    B = torch.randn(part.n_pad, 200, device=dev)
    Bz = B * (torch.rand_like(B) > 0.5)
    Pn = torch.randn(part.n_pad, 20, device=dev)
    out = torch.empty(part.n_loc, 200, device=dev); outn = torch.empty(part.n_loc, 20, device=dev)
    for chunk in (64, 128, 256, 512, 1024, 2048):
        plan = sh.plan(chunk_nnz=chunk)
        t1 = timeit(lambda: ops.spmm(sh, B, plan=plan, out=out))
        t2 = timeit(lambda: ops.spmm(sh, Bz, plan=plan, out=out))
        t3 = timeit(lambda: ops.spmm(sh, Pn, plan=plan, out=outn))
        print(f"world {world} chunk {chunk:5d} chunks {plan.n_chunks:7d} split {plan.n_split_rows:6d}: wide {t1*1e3:7.1f} us  wide(50% zeros) {t2*1e3:7.1f} us  narrow {t3*1e3:6.1f} us", flush=True)
