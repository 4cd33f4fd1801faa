from pytextgcn_b200.graphbuilder import compute_word_word_edges, sliding_window_tester

__all__ = ["compute_word_word_edges", "sliding_window_tester"]
