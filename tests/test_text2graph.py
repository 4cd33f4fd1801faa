"""Text2GraphTransformer: the emitted Data must have the reference's layout (text2graph.py:162-193).
The expected graph is re-derived here by replaying the reference's dense recipe
(text2graph.py:130-171: CountVectorizer(...).toarray(), TfidfTransformer().todense(), th.nonzero,
flip, cat) with plain sklearn/torch on the 4-sentence corpus of textgcn/test/test_text2graph.py:15-18."""
import pickle

import numpy as np
import pytest
import torch as th
from sklearn.feature_extraction.text import CountVectorizer, TfidfTransformer

from textgcn import Text2GraphTransformer
from textgcn.lib import compute_word_word_edges
from oracle import graphbuilder_oracle as GO

CORPUS = ["Time is an illusion. Lunchtime doubly so.",
          "The ships hung in the sky in much the same way that bricks don't.",
          "If there's anything more important than my ego around, I want it caught and shot now.",
          "Would it save you a lot of time if I just gave up and went mad now?"]


def _reference_recipe(corpus, min_df, window, stop_words, max_length=None):
    import re
    cv = CountVectorizer(stop_words=stop_words, min_df=min_df, max_df=1.0)
    occ = cv.fit_transform(corpus).toarray()                                     # text2graph.py:130-131
    n_docs, V = occ.shape
    toks = [[cv.vocabulary_[t.lower()] for t in re.findall(r"\w+", d) if t.lower() in cv.vocabulary_][:max_length]
            for d in corpus]
    L = max(map(len, toks))
    X = np.array([t + [-1] * (L - len(t)) for t in toks], dtype=np.int32)        # text2graph.py:40-44
    tfidf = th.from_numpy(np.asarray(TfidfTransformer().fit_transform(occ).todense()))   # :145
    docu = th.nonzero(th.from_numpy(occ))                                        # :148
    sym = th.flip(docu, dims=[1])                                                # :150
    ww, www = GO.compute_word_word_edges(X, V, window)                           # :156-160 (dense C restatement)
    weights = th.cat([th.from_numpy(www).double(), tfidf[tuple(docu.T)], tfidf[tuple(docu.T)]])     # :162-166
    coo = th.vstack([th.from_numpy(ww).long(), docu + th.tensor([V, 0]), sym + th.tensor([0, V])]).long()   # :167-171
    return coo.T, weights.float(), V, n_docs


@pytest.mark.parametrize("window,min_df", [(3, 1), (20, 1), (5, 2)])
def test_layout_matches_reference_recipe(window, min_df):
    t2g = Text2GraphTransformer(min_df=min_df, window_size=window, rm_stopwords=False)
    g = t2g.fit_transform(CORPUS, y=[1, 0, 1, 0], test_idx=[2], val_idx=[1])
    ei, ew, V, D = _reference_recipe(CORPUS, min_df, window, None)
    assert t2g.n_vocabs_ == V and t2g.n_docs_ == D and t2g.n_nodes_ == V + D and g.n_vocab == V
    assert g.edge_index.dtype == th.int64 and tuple(g.edge_index.shape) == tuple(ei.shape)
    assert not g.edge_index.is_contiguous()                                      # coo.T view, text2graph.py:171,192
    assert th.equal(g.edge_index, ei)
    assert g.edge_attr.dtype == th.float32 and th.equal(g.edge_attr, ew)
    assert g.x.is_sparse and tuple(g.x.shape) == (V + D, V + D) and th.equal(g.x.to_dense(), th.eye(V + D))
    assert g.y.dtype == th.int64 and g.y[:V].sum() == 0 and g.y[V:].tolist() == [1, 0, 1, 0]
    assert g.test_mask[V:].tolist() == [False, False, True, False]
    assert g.val_mask[V:].tolist() == [False, True, False, False]
    assert g.train_mask[V:].tolist() == [True, False, False, True] and not g.train_mask[:V].any()


def test_survey_replay_counts():
    # the 4-sentence corpus with window 3 gives 43 words + 4 docs, 172 + 2*48 = 268 edges,
    # symmetric, no loops, no duplicates
    g = Text2GraphTransformer(min_df=1, window_size=3, rm_stopwords=False).fit_transform(CORPUS, y=[0, 1, 0, 1], test_idx=[3])
    assert g.n_vocab == 43 and g.edge_index.shape[1] == 268
    n = 47
    A = th.zeros(n, n)
    A[g.edge_index[0], g.edge_index[1]] = g.edge_attr
    assert th.equal(A, A.T) and A.diag().abs().sum() == 0
    assert th.unique(g.edge_index[0] * n + g.edge_index[1]).numel() == 268


def test_hierarchy_features_and_max_length():
    hf = th.nn.functional.one_hot(th.tensor([0, 2, 1, 2]), 3).float()
    t2g = Text2GraphTransformer(min_df=1, window_size=3, rm_stopwords=False, max_length=5)
    g = t2g.fit_transform(CORPUS, y=[0, 1, 2, 1], test_idx=[0], hierarchy_feats=hf)
    V, n = g.n_vocab, int(g.x.shape[0])
    assert tuple(g.x.shape) == (n, n + 3)                                        # perlevel_dbpedia.py:140-141
    xd = g.x.to_dense()
    assert th.equal(xd[:, :n], th.eye(n)) and th.equal(xd[V:, n:], hf) and xd[:V, n:].abs().sum() == 0
    assert t2g.max_sent_len_ == 5
    # doc-word edges still come from the untruncated text (text2graph.py:131 vs :139)
    ei, ew, _, _ = _reference_recipe(CORPUS, 1, 3, None, max_length=5)
    assert th.equal(g.edge_index, ei) and th.equal(g.edge_attr, ew)


def test_stopwords_vocabulary_property_and_pickle(tmp_path):
    t2g = Text2GraphTransformer(min_df=1, window_size=3, rm_stopwords=True, save_path=str(tmp_path))
    g = t2g.fit_transform(CORPUS, y=[0, 1, 0, 1], test_idx=[1])
    assert "the" not in t2g.vocabulary and "ships" in t2g.vocabulary
    files = list(tmp_path.glob("TGData_*.p"))
    assert len(files) == 1
    g2 = Text2GraphTransformer.load_graph(str(files[0]))                        # text2graph.py:206-217
    assert th.equal(g2.edge_index, g.edge_index) and th.equal(g2.edge_attr, g.edge_attr) and g2.n_vocab == g.n_vocab
    with pytest.raises(FileNotFoundError):
        Text2GraphTransformer.load_graph(str(tmp_path / "nope.p"))
    assert th.is_grad_enabled()                                                  # restored (text2graph.py:114,203)
    pickle.dumps(g)


def test_reference_import_surface():
    import textgcn
    from textgcn.lib.models import GCN
    from textgcn.lib.clib import compute_word_word_edges as c2
    assert textgcn.Text2GraphTransformer is Text2GraphTransformer and c2 is compute_word_word_edges
    m = GCN(10, 3, n_hidden_gcn=100, dropout=0.7)
    assert [k for k, _ in m.named_parameters()] == ["layers.0.weight", "layers.0.bias", "layers.1.weight", "layers.1.bias"]


def test_token_encoding_in_worker_processes_matches_single_process(monkeypatch):
    """Large corpora are encoded by n_jobs worker processes, one contiguous chunk each: same token matrix."""
    import numpy as np
    from pytextgcn_b200 import text2graph as T
    rng = np.random.default_rng(3)
    vocab = {f"w{i}": i for i in range(50)}
    docs = [" ".join(f"W{j}" if j % 7 == 0 else f"w{j}" for j in rng.integers(0, 60, size=rng.integers(0, 30))) for _ in range(101)]
    one, l1 = T._encode_input(docs, 1, vocab, 0, len(docs), 12)
    monkeypatch.setattr(T, "_PARALLEL_MIN_DOCS", 10)
    many, l2 = T._encode_input(docs, 3, vocab, 0, len(docs), 12)
    assert l1 == l2 and np.array_equal(one, many)
    assert one.max() < 50 and (one >= -1).all()
