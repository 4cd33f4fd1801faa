"""One wide + one narrow SpMM at a named shape, for ncu captures."""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from pytextgcn_b200 import make_graph, ops
from pytextgcn_b200.graph import upload_graph
from pytextgcn_b200.synthetic import SHAPES
shape = sys.argv[1] if len(sys.argv) > 1 else "20ng"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = make_graph(shape)
N = g.x.shape[0]
dev = torch.device("cuda")
gr = upload_graph(g.edge_index.T.contiguous().to(dev).T, g.edge_attr.to(dev), N)
H, C = SHAPES[shape].hidden, SHAPES[shape].n_classes
B = torch.randn(N, H, device=dev); out = torch.empty(N, H, device=dev)
P = torch.randn(N, ops.pad4(C), device=dev); outp = torch.empty(N, ops.pad4(C), device=dev)
for _ in range(reps):
    ops.spmm(gr, B, out=out)
    ops.spmm(gr, P, out=outp)
torch.cuda.synchronize()
print("ok", gr.nnz)
