"""`Text2GraphTransformer`: corpus -> TextGCN doc-word graph, same constructor, `fit_transform`
signature and emitted `Data` as the reference (textgcn/lib/text2graph.py:49-246).

The emitted object is identical in layout to the reference's (text2graph.py:162-193):
  * word nodes [0, V), document nodes [V, V+D);
  * edge order: word-word PMI pairs (i,j),(j,i) in upper-triangle row-major order, then
    (doc+V, word) for every doc-word non-zero in doc-major order, then (word, doc+V);
  * `edge_index` is the non-contiguous `.T` view of an (E, 2) int64 tensor, `edge_attr` fp32
    (PMI, then sklearn TfidfTransformer() weights twice), `x` sparse COO identity (optionally
    `[I | hierarchy_feats]` on the document rows), `y` with 0 on word rows, three bool masks,
    `n_vocab`.
What changed underneath (one-off CPU preprocessing, not part of epochs/sec):
  * the count / tf-idf matrices stay sparse (the reference densifies both: text2graph.py:131,145);
  * word-word edges come from the threaded native builder (graphbuilder.py), bit-identical to
    the reference's Cython one;
  * tokenisation is `re.findall(r"\\w+")` -- what nltk.RegexpTokenizer(r"\\w+") does -- so nltk
    is optional; stop words come from nltk when its corpus is installed, else sklearn's built-in
    English list (no network here: `nltk.download` at text2graph.py:85 cannot run).
"""
from __future__ import annotations

import glob
import os
import pickle
import re
import time
from typing import Dict, List, Optional, Union

import numpy as np
import torch as th
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.feature_extraction.text import CountVectorizer, TfidfTransformer

from .data import Data
from .graphbuilder import compute_word_word_edges

_TOKEN = re.compile(r"\w+", re.UNICODE | re.MULTILINE | re.DOTALL)


def _stop_words() -> Optional[List[str]]:
    try:  # the reference's source of stop words (text2graph.py:84-86)
        import nltk
        return sorted(set(nltk.corpus.stopwords.words("english")))
    except Exception:
        from sklearn.feature_extraction.text import ENGLISH_STOP_WORDS
        return sorted(ENGLISH_STOP_WORDS)


_PARALLEL_MIN_DOCS = 50000      # below this a single process is faster than starting workers


def _encode_chunk(docs, vocabulary, max_len):
    """Token ids of every document of the chunk: tokenise, lower-case each token, keep in-vocabulary tokens, truncate."""
    sl = slice(None) if max_len is None else slice(max_len)
    get = vocabulary.get
    return [[i for i in map(get, [m.lower() for m in _TOKEN.findall(doc)]) if i is not None][sl] for doc in docs]


def _encode_input(X, n_jobs, vocabulary, verbose, n_docs, max_len):
    """text2graph.py:20-46: tokenise, lower-case, keep in-vocabulary tokens, truncate to max_len,
    pad with -1 to the longest document.  The reference hands every document to joblib as its own task, which pickles
    the vocabulary with each batch; here one process encodes ~10 k documents per second, so worker processes are used
    only for large corpora, one contiguous chunk per worker (the vocabulary crosses the process boundary n_jobs times)."""
    if n_jobs and n_jobs > 1 and len(X) > _PARALLEL_MIN_DOCS:
        import joblib as jl
        k = min(int(n_jobs), 64)
        bounds = [len(X) * i // k for i in range(k + 1)]
        parts = jl.Parallel(n_jobs=k)(jl.delayed(_encode_chunk)(X[bounds[i]:bounds[i + 1]], vocabulary, max_len) for i in range(k))
        docs = [d for part in parts for d in part]
    else:
        docs = _encode_chunk(X, vocabulary, max_len)
    max_sent_len = max(map(len, docs)) if docs else 0
    max_sent_len = max(max_sent_len, 1)
    out = np.full((n_docs, max_sent_len), -1, dtype=np.int32)
    for i, d in enumerate(docs):
        out[i, :len(d)] = d
    if verbose > 1:
        print(f"Sequence length is {max_sent_len}")
    return out, max_sent_len


class Text2GraphTransformer(BaseEstimator, TransformerMixin):
    def __init__(self, min_df: Union[int, float] = 5, window_size: int = 20, save_path: str = None,
                 n_jobs: int = 1, max_df=1.0, verbose=0, rm_stopwords=True, sparse_features=True,
                 max_length: Optional[int] = None):
        self.max_length = max_length
        self.sparse_features = sparse_features
        self.rm_stopwords = rm_stopwords
        self.verbose = verbose
        self.max_df = max_df
        self.n_jobs = n_jobs
        assert min_df > 0
        self.min_df = min_df
        self.save_path = save_path
        self.input = None
        self.cv = None
        self.window_size = window_size
        self.stop_words = _stop_words() if self.rm_stopwords else None

    def fit_transform(self, X: Union[List[str], str],
                      y: Union[th.Tensor, np.ndarray, List[int], None] = None,
                      test_idx: Union[th.Tensor, np.ndarray, List[int], None] = None,
                      val_idx: Union[th.Tensor, np.ndarray, List[int], None] = None,
                      hierarchy_feats: Union[th.Tensor, None] = None) -> Data:
        """Corpus -> Data (see module docstring).  Arguments as the reference (text2graph.py:88-113)."""
        prev_grad = th.is_grad_enabled()
        th.set_grad_enabled(False)
        try:
            test_idx = th.as_tensor(test_idx if test_idx is not None else [], dtype=th.long).view(-1)
            if y is not None:
                y = th.as_tensor(np.asarray(y), dtype=th.long)
            if isinstance(X, list):
                self.input = X
            else:
                self.input = []
                for f in sorted(glob.glob(os.path.join(X, "*.txt"))):
                    with open(f, "r") as fp:
                        self.input.append(fp.read())
            self.cv = CountVectorizer(stop_words=self.stop_words, min_df=self.min_df, max_df=self.max_df)
            occ = self.cv.fit_transform(self.input).tocsr()
            occ.sort_indices()
            n_docs, n_vocabs = occ.shape
            self.n_docs_, self.n_vocabs_, self.n_nodes_ = n_docs, n_vocabs, n_docs + n_vocabs
            if self.verbose > 1:
                print(f"Number of documents in input: {n_docs}\nVocabulary size: {n_vocabs}")
            Xtok, self.max_sent_len_ = _encode_input(self.input, self.n_jobs, self.cv.vocabulary_, self.verbose,
                                                     n_docs, self.max_length)
            # doc-word edges: non-zeros of the count matrix in doc-major order, weight = tf-idf (text2graph.py:145-150)
            tfidf = TfidfTransformer().fit_transform(occ).tocsr()
            tfidf.sort_indices()
            occ_coo = occ.tocoo()
            d_idx = th.from_numpy(occ_coo.row.astype(np.int64))
            w_idx = th.from_numpy(occ_coo.col.astype(np.int64))
            dw_w = th.from_numpy(np.asarray(tfidf[occ_coo.row, occ_coo.col]).reshape(-1))      # float64, like the reference
            # word-word edges (text2graph.py:156-160)
            ww_coo, ww_w = compute_word_word_edges(Xtok, n_vocabs, n_docs, self.max_sent_len_, self.window_size,
                                                   self.n_jobs, self.verbose)
            edge_weights = th.cat([th.from_numpy(ww_w).double(), dw_w, dw_w])
            coo = th.cat([th.from_numpy(ww_coo).long(),
                          th.stack([d_idx + n_vocabs, w_idx], dim=1),
                          th.stack([w_idx, d_idx + n_vocabs], dim=1)], dim=0)
            if self.verbose > 0:
                print(f"total edge shape is {coo.shape}")
            node_feats = self.node_feats(hierarchy_feats) if self.sparse_features else th.eye(self.n_nodes_)
            test_mask = th.zeros(self.n_nodes_, dtype=th.bool)
            val_mask = th.zeros(self.n_nodes_, dtype=th.bool)
            test_mask[test_idx + n_vocabs] = True
            if val_idx is not None:
                val_mask[th.as_tensor(val_idx, dtype=th.long).view(-1) + n_vocabs] = True
            train_mask = th.logical_not(th.logical_or(test_mask, val_mask))
            train_mask[:n_vocabs] = False
            y_nodes = th.zeros(self.n_nodes_, dtype=th.long)        # pseudo-labels 0 on word rows (text2graph.py:190-191)
            if y is not None:
                y_nodes[n_vocabs:] = y
            g = Data(x=node_feats.float(), edge_index=coo.T, edge_attr=edge_weights.float(), y=y_nodes,
                     test_mask=test_mask, train_mask=train_mask, val_mask=val_mask, n_vocab=n_vocabs)
            if self.save_path is not None:
                os.makedirs(self.save_path, exist_ok=True)
                savefile = os.path.join(self.save_path, f"TGData_{time.time()}.p")
                with open(savefile, "wb") as fp:
                    pickle.dump(g, fp)
            return g
        finally:
            th.set_grad_enabled(prev_grad)

    @staticmethod
    def load_graph(save_path):
        """Loads a pickled graph written by fit_transform (text2graph.py:206-217)."""
        if not os.path.exists(save_path):
            raise FileNotFoundError("Given file does not exist!")
        with open(save_path, "rb") as fp:
            return pickle.load(fp)

    @property
    def vocabulary(self) -> Dict[str, int]:
        return self.cv.vocabulary_

    def node_feats(self, hierarchy_feats):
        """Sparse feature matrix I_N or [I_N | hierarchy_feats on the doc rows] (text2graph.py:226-246)."""
        n = self.n_nodes_
        idx = th.arange(n, dtype=th.long)
        inds, vals, n_cols = th.stack([idx, idx]), th.ones(n, dtype=th.float32), n
        if hierarchy_feats is not None:
            hf = th.as_tensor(np.asarray(hierarchy_feats), dtype=th.float32)
            r, c = th.nonzero(hf, as_tuple=True)
            inds = th.cat([inds, th.stack([r + self.n_vocabs_, c + n])], dim=1)
            vals = th.cat([vals, hf[r, c]])
            n_cols = n + hf.shape[1]
        import warnings
        warnings.filterwarnings("ignore", message="Sparse invariant checks")
        return th.sparse_coo_tensor(inds, vals, size=(n, n_cols), dtype=th.float32, check_invariants=False).coalesce()
