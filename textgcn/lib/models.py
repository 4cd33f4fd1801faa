"""`textgcn.lib.models.GCN` of the reference (textgcn/lib/models.py:6-25), running on the sm_100a kernels.
EGCN / JumpingKnowledgeNetwork / MLP are outside the hot path this repository implements."""
from pytextgcn_b200.models import GCN, GCNConv

__all__ = ["GCN", "GCNConv"]
