"""Runs the fused trainer on every named benchmark shape and prints ms/epoch."""
import sys, time, json, torch
sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from pytextgcn_b200 import make_graph, GCN, SHAPES
from pytextgcn_b200.trainer import TextGCNTrainer
from pytextgcn_b200.graph import upload_graph
names = sys.argv[1:] or ["r8", "20ng", "amazon", "dbpedia"]
dev = torch.device("cuda")
for name in names:
    shape = SHAPES[name]
    hier = {"dbpedia": 70}.get(name)
    t0 = time.time(); g = make_graph(shape, seed=0, hierarchy_classes=hier); t_gen = time.time() - t0
    n, in_ch = int(g.x.shape[0]), int(g.x.shape[1])
    t0 = time.time()
    ei = g.edge_index.T.contiguous().to(dev).T; ea = g.edge_attr.to(dev)
    graph = upload_graph(ei, ea, n); torch.cuda.synchronize(); t_up = time.time() - t0
    gd = g.clone(); gd.edge_index, gd.edge_attr = ei, ea; gd = gd.to(dev)
    torch.manual_seed(0)
    gcn = GCN(in_ch, shape.n_classes, n_hidden_gcn=shape.hidden, dropout=shape.dropout).to(dev)
    tr = TextGCNTrainer(gcn, gd, lr=shape.lr, amsgrad=shape.amsgrad, graph=graph)
    first = tr.epoch()
    for _ in range(4): tr.epoch()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    a.record()
    for _ in range(K): tr.train_step(); tr.eval_step()
    b.record(); torch.cuda.synchronize()
    last = tr.epoch()
    print(json.dumps({"shape": name, "n_nodes": n, "in_channels": in_ch, "nnz": graph.nnz, "hidden": shape.hidden, "classes": shape.n_classes,
                      "ms_per_epoch": a.elapsed_time(b) / K, "epochs_per_s": 1e3 * K / a.elapsed_time(b),
                      "gen_s": round(t_gen, 1), "upload_s": round(t_up, 2), "chunk_nnz": tr.plan.chunk_nnz,
                      "loss_first": first["loss"], "loss_last": last["loss"], "acc_train_last": last["acc_train"],
                      "mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)
    del tr, gcn, gd, graph, g, ei, ea
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
