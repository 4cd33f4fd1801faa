"""Developer timing script (not the contract bench): per-kernel CUDA-event timings."""
import sys, time, torch
sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from pytextgcn_b200 import make_graph, ops, GCN
from pytextgcn_b200.graph import upload_graph

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts)//2]

shape = sys.argv[1] if len(sys.argv) > 1 else "20ng"
t = time.time(); g = make_graph(shape); print("gen", time.time() - t)
N = g.x.shape[0]
dev = torch.device("cuda")
t = time.time(); ei = g.edge_index.T.contiguous().to(dev).T; ea = g.edge_attr.to(dev); torch.cuda.synchronize(); print("h2d", time.time() - t)
t = time.time(); gr = upload_graph(ei, ea, N); torch.cuda.synchronize(); print("upload", time.time() - t)
t = time.time(); gr2 = upload_graph(ei, ea, N); torch.cuda.synchronize(); print("upload2", time.time() - t)
print("nnz", gr.nnz)
from pytextgcn_b200.synthetic import SHAPES
H, C = SHAPES[shape].hidden, SHAPES[shape].n_classes
import itertools
for chunk, srt in itertools.product((256, 512, 1024, 2048), (False, True)):
    plan = gr.plan(chunk_nnz=chunk, sort_chunks=srt)
    print("sorted" if srt else "unsorted", end=" ")
    print("chunk", chunk, "n_chunks", plan.n_chunks, "split rows", plan.n_split_rows, "slots", plan.n_slots, "max", plan.max_row_nnz)
    for F in (H, ops.pad4(C)):
        B = torch.randn(N, F, device=dev)
        out = torch.empty(N, F, device=dev)
        ms = timeit(lambda: ops.spmm(gr, B, plan=plan, out=out))
        bytes_ = gr.nnz * 8 + (N + 1) * 4 + 2 * N * F * 4
        print(f"  spmm F={F}: {ms*1e3:.1f} us  alg {bytes_/1e6:.0f} MB -> {bytes_/ms/1e6:.0f} GB/s ; gather {gr.nnz*F*4/ms/1e6:.0f} GB/s")
Bb = torch.randn(N, ops.pad8(H), device=dev).to(torch.bfloat16)
outb = torch.empty(N, ops.pad8(H), device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: ops.spmm(gr, Bb, out=outb))
print(f"  spmm bf16 F={H}: {ms*1e3:.1f} us")
