"""Mirrors textgcn/lib/__init__.py of the reference."""
from pytextgcn_b200.text2graph import Text2GraphTransformer
from pytextgcn_b200.graphbuilder import compute_word_word_edges, sliding_window_tester
from . import models

__all__ = ["Text2GraphTransformer", "compute_word_word_edges", "sliding_window_tester", "models"]
