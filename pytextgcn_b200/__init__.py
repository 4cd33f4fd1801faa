"""pytextgcn_b200 -- B200-native (sm_100a) implementation of PyTextGCN's training hot path:
the 2-layer GCN forward/backward over the Text2GraphTransformer doc-word graph.

Public surface (mirrors the reference, BeFranke/PyTextGCN):
    Data                    graph object (torch_geometric.data.Data when importable)
    GCN, GCNConv            textgcn/lib/models.py:6-25 drop-ins
    upload_graph/get_graph  one-off COO -> CSR(A_hat) conversion, bit-exact vs gcn_norm
    make_graph, SHAPES      synthetic doc-word graphs of the benchmark shapes
The compute lives in pytextgcn_b200/lib/libtextgcn_b200.so (C ABI: include/textgcn_b200.h).
"""
from .data import Data
from .graph import GraphCSR, upload_graph, upload_graph_cached, get_graph, clear_cache, save_csr, load_csr
from .models import GCN, GCNConv
from .synthetic import make_graph, SHAPES, GraphShape
from . import ops

__all__ = ["Data", "GraphCSR", "upload_graph", "upload_graph_cached", "save_csr", "load_csr", "get_graph", "clear_cache", "GCN", "GCNConv",
           "make_graph", "SHAPES", "GraphShape", "ops"]
