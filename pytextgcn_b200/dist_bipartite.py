"""Multi-GPU TextGCN for graphs with many more documents than words: the word-block exchange of SURVEY.md 8e.

A Text2GraphTransformer graph is bipartite apart from its word-word block (text2graph.py:148-170, words first):

    A_hat = [[ WW , WD ],          WW: PMI word-word edges,  WD = DW^T: TF-IDF word-document edges,
             [ DW , I' ]]          documents never neighbour documents (only their self-loop).

The row partition of dist.py exchanges all N rows of every operand.  Here every rank owns a block of the WORDS and a
block of the DOCUMENTS (both dealt in snake order of decreasing row length), and only word rows ever cross ranks:

    own rows of A_hat X  =  main_r @ [ X_words (all-gathered) ; X_own_rows ]          (word rows: WW part only)
    word rows, WD part   =  sum over ranks s of  Q_s @ X_docs_of_s                     (Q_s = the columns of WD rank s owns)

so per propagation a rank all-gathers the [V, F] word block and exchanges (all-to-all) the [V, F] partial word rows
the ranks computed from their own documents -- 2 V F values instead of N F.  The partial rows are added by the SpMM
epilogue in rank order (fixed order: deterministic).  At the scale configuration (200 k words, 1 M documents, F = 256)
that is 1/3 of the row partition's traffic and the document rows need nothing remote beyond the word block.

The class-wide propagations need even less: logits are read on document rows only (no partial word rows), and the
loss gradient is zero on word rows (no all-gather, only the partial word rows of the training documents).

With symmetric memory (NVLink peer mappings + NVSwitch multicast) the exchange is fused into the kernels: producers repeat
their word rows with multimem.st into every rank's operand buffer (all-gather), the Q SpMM's epilogue stores every partial
word row straight into its owner's slot buffer (all-to-all), one device barrier per propagation.  Fallback: the NCCL
collectives of torch.distributed (all_gather_into_tensor, all_to_all_single; the all-gather runs concurrently with the Q
SpMM).  Either way the whole epoch is one CUDA graph.
Parity: dist.parity_against_single_gpu(..., partition="words") -- the single-GPU trainer on the
renumbered graph; only the order of the partial sums differs (fp32 rounding).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .dist import DistTextGCNTrainer, RowPartition
from .graph import GraphCSR


class BipartitePartition(RowPartition):
    """Words and documents dealt separately (snake order of decreasing row length).  Rank r owns the contiguous new ids
    [r * n_loc, (r+1) * n_loc): first its v_loc words, then its d_loc documents.  Same interface as RowPartition
    (new_id / old_id / rows_of / to_new / to_old), so renumbered_data() and the parity helper work unchanged."""

    def __init__(self, row_nnz: torch.Tensor, n_vocab: int, world: int):
        n = int(row_nnz.numel())
        if not 0 < n_vocab < n:
            raise ValueError("the bipartite partition needs 0 < n_vocab < n_nodes")
        self.n, self.world, self.n_vocab = n, world, int(n_vocab)
        self.v_loc = (n_vocab + world - 1) // world
        self.d_loc = (n - n_vocab + world - 1) // world
        self.n_loc = self.v_loc + self.d_loc
        self.n_pad = self.n_loc * world
        self.v_pad = self.v_loc * world
        dev = row_nnz.device
        w = row_nnz.to(torch.int64)
        self.new_id = torch.empty(n, dtype=torch.int64, device=dev)
        for first, last, offset in ((0, n_vocab, 0), (n_vocab, n, self.v_loc)):
            order = torch.sort(w[first:last], descending=True, stable=True).indices
            k = torch.arange(last - first, device=dev)
            blk, pos = k // world, k % world
            rank = torch.where(blk % 2 == 0, pos, world - 1 - pos)
            self.new_id[first + order] = rank * self.n_loc + offset + blk
        self.old_id = torch.full((self.n_pad,), -1, dtype=torch.int64, device=dev)
        self.old_id[self.new_id] = torch.arange(n, device=dev)
        self.row_nnz = row_nnz


def _csr(n_rows: int, rows: torch.Tensor, cols: torch.Tensor, val: torch.Tensor):
    """(rowptr int32, colidx int32, val) of the COO entries, rows ascending, entry order inside a row kept."""
    order = torch.sort(rows, stable=True).indices
    rp = torch.zeros(n_rows + 1, dtype=torch.int64, device=rows.device)
    rp[1:] = torch.cumsum(torch.bincount(rows, minlength=n_rows), 0)
    return rp.to(torch.int32), cols[order].to(torch.int32).contiguous(), val[order].contiguous()


def shard_bipartite(rowptr: torch.Tensor, colidx: torch.Tensor, val: torch.Tensor, part: BipartitePartition, rank: int):
    """The two CSR pieces of rank `rank` (device-agnostic torch index ops; covered by the CPU tests):

    main: n_loc rows (own words, own documents) x (v_pad + n_loc) columns = [all words in new-id order of their
          owners ; the rank's own rows (words, documents)]; holds every entry of the own rows EXCEPT word-row x
          document-column ones.  Word columns point into the first v_pad positions, document columns into the tail
          (v_pad + local id), so an operand buffer is [gathered word block ; X_loc] and the kernels that produce X_loc
          write it in place.
    q:    v_pad rows (all words, ordered by owner rank) x d_loc columns (own documents): the entries A_hat[w, d] of the
          documents this rank owns; q @ X_own_docs is this rank's contribution to every word row.
    Returns ((rowptr, colidx, val) of main, (rowptr, colidx, val) of q)."""
    dev = rowptr.device
    n, nl, vl, vp = part.n, part.n_loc, part.v_loc, part.v_pad
    counts = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(n, device=dev), counts)
    new_id = part.new_id.to(dev)
    rn, cn = new_id[rows], new_id[colidx.to(torch.int64)]
    del rows
    r_rank, r_loc, c_rank, c_loc = rn // nl, rn % nl, cn // nl, cn % nl
    r_word, c_word = r_loc < vl, c_loc < vl
    if bool((~r_word & ~c_word & (rn != cn)).any()):
        raise NotImplementedError("document-document edges: the word-block exchange needs a Text2GraphTransformer graph "
                                  "(text2graph.py:148-170); use the row partition (DistTextGCNTrainer)")
    del rn, cn
    m = (r_rank == rank) & ~(r_word & ~c_word)
    m_cols = torch.where(c_word[m], c_rank[m] * vl + c_loc[m], vp + c_loc[m])
    main = _csr(nl, r_loc[m], m_cols, val[m])
    q = r_word & ~c_word & (c_rank == rank)
    qq = _csr(vp, r_rank[q] * vl + r_loc[q], c_loc[q] - vl, val[q])
    return main, qq


class BipartiteTextGCNTrainer(DistTextGCNTrainer):
    """Same epoch as DistTextGCNTrainer (train step + eval forward + masked val loss / accuracy), same public surface;
    the partition and the exchange are the word-block scheme described in the module docstring."""

    def __init__(self, g, n_classes: int, hidden: int, dropout: float, lr: float, amsgrad: bool,
                 rank: int, world: int, dev: torch.device, seed: int = 0, betas=(0.9, 0.999), eps: float = 1e-8,
                 graph: Optional[GraphCSR] = None, init_weights: Optional[Dict[str, torch.Tensor]] = None,
                 use_cuda_graph: bool = False, keep_w1_grad: bool = True, share_h1: bool = True,
                 tensor_cores: Optional[bool] = None, tc_min_density: float = 0.05, overlap: bool = True,
                 exchange: str = "peer", fused_stores: bool = True, fuse_adam: bool = True):
        import torch.distributed as dist
        from . import ops
        from .graph import upload_graph
        from .models import decode_features
        self.dist, self.ops = dist, ops
        self.rank, self.world, self.dev = rank, world, dev
        if not fuse_adam:
            raise NotImplementedError("the word-block trainer always runs Adam on W1 in the SpMM epilogue")
        n = int(g.x.shape[0])
        feat = decode_features(g.x, getattr(g, "n_vocab", None))
        if feat is None or feat.Fdoc is not None:
            raise NotImplementedError("the word-block trainer expects featureless input x = I (text2graph.py:226-246)")
        n_vocab = int(getattr(g, "n_vocab", 0) or 0)
        if hidden % 4 != 0:
            raise NotImplementedError("hidden width must be a multiple of 4")
        self.hier, self.c_prev = False, 0
        self.n, self.H, self.C, self.Cp = n, hidden, n_classes, ops.pad4(n_classes)
        self.p, self.lr, self.amsgrad, self.betas, self.eps, self.seed = dropout, lr, amsgrad, betas, eps, seed
        masks = g.train_mask.cpu() | g.val_mask.cpu()
        if getattr(g, "test_mask", None) is not None:
            masks = masks | g.test_mask.cpu()
        if bool(masks[:n_vocab].any()):
            raise NotImplementedError("masked word nodes: the word-block trainer computes logits on document rows only")
        full = graph if graph is not None else upload_graph(g.edge_index.to(dev), g.edge_attr.to(dev), n)
        if not full.is_symmetric():
            raise NotImplementedError("the word-block trainer needs a symmetric A_hat (Text2GraphTransformer graphs are)")
        row_nnz = (full.rowptr[1:] - full.rowptr[:-1]).to(torch.int64)
        self.part = part = BipartitePartition(row_nnz, n_vocab, world)
        (mrp, mci, mv), (qrp, qci, qv) = shard_bipartite(full.rowptr, full.colidx, full.val, part, rank)
        self.nnz_global = full.nnz
        del full
        nl, vl, dl, vp, H, Cp = part.n_loc, part.v_loc, part.d_loc, part.v_pad, hidden, self.Cp
        self.shard = GraphCSR(nl, mrp, mci, mv, None, None, n_cols=vp + nl)
        self.shard._symmetric = True          # only ever used as rows of the symmetric global matrix
        self.qshard = GraphCSR(vp, qrp, qci, qv, None, None, n_cols=dl)
        self.qshard._symmetric = True
        self.plan, self.plan_q = self.shard.plan(), self.qshard.plan()
        self.tc = None
        auto = tensor_cores is None and self.shard.nnz >= 200_000 and (vp + nl) * ((H + 15) // 16 * 16) * 8 <= (104 << 20)
        if (tensor_cores or auto) and 64 <= H <= 256:
            from .tc_plan import build_tc_plan
            tc = build_tc_plan(self.shard, min_density=tc_min_density, width=H,
                               n_sms=torch.cuda.get_device_properties(dev).multi_processor_count)
            if tc is not None and (tensor_cores or tc.nnz_dense >= 0.15 * self.shard.nnz):
                self.tc = tc
        f32 = dict(dtype=torch.float32, device=dev)
        self.overlap = bool(overlap) and world > 1
        n_small = H + H * n_classes + n_classes
        self.n_small, self.n_small_pad = n_small, (n_small + 3) // 4 * 4
        # Exchanged buffers.  OP*: operands [all words (v_pad rows, rank-major) ; this rank's own rows X_loc (n_loc)]: W1,
        # dZ1, P of the train and of the eval forward.  The tail IS the local operand (W1_loc, dZ1_loc, P_loc are views
        # of it): its producer writes it in place and the main SpMM reads the document columns from there, no copies.
        # SL*: partial word rows, slot s = rank s's contribution to MY words (two hidden-wide ones: the forward and the
        # backward propagation alternate, so a peer never overwrites the slots a slower rank is still adding).
        # With symmetric memory (NVLink peer mappings + NVSwitch multicast) the exchange is fused into the kernels: the
        # producers of an operand repeat their WORD rows with multimem.st into every rank's OP buffer, the Q SpMM stores
        # each partial word row straight into its owner's slot (peer stores), and one device barrier separates
        # producers from consumers.  Otherwise: NCCL all_gather_into_tensor / all_to_all_single.
        # The peer / NCCL decision is collective (see DistTextGCNTrainer).
        self.exchange, self.exchange_error, self.px = "nccl word-block", None, None
        xshapes = {"OPW": (vp + nl, H), "OPD": (vp + nl, H), "OPCt": (vp + nl, Cp), "OPCe": (vp + nl, Cp),
                   "SLf": (world, vl, H), "SLb": (world, vl, H), "SLc": (world, vl, Cp), "small": (world, self.n_small_pad)}
        xb: Dict[str, torch.Tensor] = {}
        if world > 1 and exchange == "peer":
            from .dist import PeerExchange
            ok = 1
            try:
                self.px = PeerExchange(dist.group.WORLD, rank, world, dev)
                for name, shp in xshapes.items():
                    xb[name] = self.px.alloc(name, shp)
            except Exception as e:
                self.exchange_error, ok = repr(e), 0
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                self.exchange = "peer word-block"
            else:
                self.px, xb = None, {}
                self.exchange_error = self.exchange_error or "symmetric allocation failed on another rank"
        for name, shp in xshapes.items():
            if name not in xb:
                xb[name] = torch.zeros(shp, **f32)
        self.X = xb
        self.bases = {}
        if self.px is not None:
            self.bases = {k: torch.tensor([int(q) for q in self.px.handles[k].buffer_ptrs], dtype=torch.int64, device=dev)
                          for k in ("SLf", "SLb", "SLc")}
        self.fused_stores = bool(fused_stores and self.px is not None and
                                 all(self.px.multicast.get(k, 0) for k in ("OPW", "OPD", "OPCt", "OPCe")))
        self.Q = {} if self.px is not None else {F: torch.zeros((vp, F), **f32) for F in {H, Cp}}   # NCCL mode: a2a source
        self._pending_reads = set()
        self._w1_words_ready = False
        # parameters: same init on every rank (same draws as DistTextGCNTrainer), W1 rows in the new order
        gen = torch.Generator().manual_seed(seed)
        if init_weights is None:
            a1, a2 = (6.0 / (n + H)) ** 0.5, (6.0 / (H + n_classes)) ** 0.5
            W1 = (torch.rand(n, H, generator=gen) * 2 - 1) * a1
            W2 = (torch.rand(H, n_classes, generator=gen) * 2 - 1) * a2
            b1, b2 = torch.zeros(H), torch.zeros(n_classes)
        else:
            W1, b1, W2, b2 = (init_weights[k].detach().cpu().float() for k in
                              ("layers.0.weight", "layers.0.bias", "layers.1.weight", "layers.1.bias"))
        lo = rank * nl
        self.W1_loc = self.X["OPW"][vp:]                        # authoritative copy of this rank's W1 rows
        self.W1_loc.copy_(part.to_new(W1[:n])[lo:lo + nl].to(dev))
        self.W1_cat = None
        self.b1, self.W2, self.b2 = b1.to(dev), W2.to(dev).contiguous(), b2.to(dev)
        self.small_slots = self.X["small"]                      # slot r = rank r's partial sums (peer mode)
        self.small_local = self.small_slots[rank]
        self.small = torch.zeros(self.n_small_pad, **f32)

        def views(buf):
            return buf[:H], buf[H:H + H * n_classes].view(H, n_classes), buf[H + H * n_classes:n_small]
        self.l_b1, self.l_W2, self.l_b2 = views(self.small_local)
        self.g_b1, self.g_W2, self.g_b2 = views(self.small)
        self.keep_w1_grad = bool(keep_w1_grad)
        self.g_W1 = torch.zeros((nl, H), **f32) if self.keep_w1_grad else None

        def state(t):
            return [torch.zeros_like(t), torch.zeros_like(t), torch.zeros_like(t) if amsgrad else None]
        self.st = [state(self.W1_loc), state(self.b1), state(self.W2), state(self.b2)]
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.adam_hyper = torch.zeros(2, dtype=torch.float32, device=dev)
        self.H1d = torch.empty((nl, H), **f32)
        self.share_h1 = bool(share_h1)
        self.H1 = torch.empty((nl, H), **f32) if (self.share_h1 and dropout > 0) else self.H1d
        self._h1_valid = False
        self.Pt_loc, self.Pe_loc = self.X["OPCt"][vp:], self.X["OPCe"][vp:]
        self.Z2 = torch.zeros((nl, Cp), **f32)
        self.OPZ = torch.zeros((vp + nl, Cp), **f32)             # operand of the class-wide backward (word part never read)
        self.dZ2_loc = self.OPZ[vp:]
        self.G2 = torch.zeros((nl, Cp), **f32)
        self.dZ1_loc = self.X["OPD"][vp:]
        self.loss_part = torch.zeros(2, dtype=torch.float64, device=dev)
        self.loss_part_val = torch.zeros(2, dtype=torch.float64, device=dev)
        self.loss_buf = torch.zeros(2, **f32)
        self.pred = torch.zeros(nl, dtype=torch.int32, device=dev)
        self.correct = torch.zeros(1, dtype=torch.int32, device=dev)
        self._nll_ws = torch.empty(2 * ((nl * 4 + 255) // 256 * 256) + 4096, dtype=torch.uint8, device=dev)
        self._db_ws = None
        y_new = part.to_new(g.y.cpu(), 0)
        tm_new = part.to_new(g.train_mask.cpu(), False)
        vm_new = part.to_new(g.val_mask.cpu(), False)
        self.y = y_new[lo:lo + nl].to(dev).contiguous()
        self.train_mask = tm_new[lo:lo + nl].to(dev).contiguous()
        self.val_mask = vm_new[lo:lo + nl].to(dev).contiguous()
        self.n_train, self.n_val = int(g.train_mask.sum()), int(g.val_mask.sum())
        yc = g.y.cpu()
        if bool(((g.train_mask.cpu() | g.val_mask.cpu()) & ((yc < 0) | (yc >= n_classes))).any()):
            raise RuntimeError(f"labels of masked rows must lie in [0, {n_classes})")
        # class-wide propagations: logits only on the masked (document) rows; the loss gradient only reaches the
        # training documents' own rows (self-loop) and, through q, the word rows
        self.restrict_rows = True
        rows_loc = part.to_new(masks, False)[lo:lo + nl].to(dev)
        self.plan_z2 = self.shard.plan_for_rows(rows_loc, self.plan)
        col_train = torch.zeros(vp + nl, dtype=torch.bool, device=dev)
        col_train[vp:] = self.train_mask
        self.shard_g2 = self.shard.select_columns(col_train)
        is_word_row = torch.arange(nl, device=dev) < vl
        nonempty = (self.shard_g2.rowptr[1:] - self.shard_g2.rowptr[:-1]) > 0
        self.plan_g2 = self.shard_g2.plan_for_rows(is_word_row | nonempty, self.shard_g2.plan())
        self.qshard_g2 = self.qshard.select_columns(self.train_mask[vl:])
        self.plan_q_g2 = self.qshard_g2.plan()
        self.use_cuda_graph = use_cuda_graph
        self._graph, self._eager_epochs, self.graph_error = None, 0, None
        self.profile = None
        self.launches_per_epoch = 0
        self.w1_stale = False
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # ---- one propagation over this rank's rows ----
    def _mirror_of(self, op: str):
        """(multicast address of this rank's word rows inside operand buffer `op`, number of word rows) or (None, 0)."""
        if not self.fused_stores:
            return None, 0
        OP = self.X[op]
        return self.px.multicast[op] + self.rank * self.part.v_loc * OP.stride(0) * 4, self.part.v_loc

    def _propagate(self, OP: torch.Tensor, F: int, op: Optional[str], sl: Optional[str], graph: GraphCSR, plan,
                   qgraph: Optional[GraphCSR], qplan, gather_words: bool, words_mirrored: bool = False, tc=None, **kw):
        """epi(rows of A_hat X owned by this rank).  OP = operand buffer [word block ; X_loc] whose tail already holds
        this rank's rows of X; `op` names it when it is an exchanged buffer, `sl` the slot buffer of the partial word
        rows (None: word rows not needed).
        words_mirrored: the kernel that produced X_loc already stored its word rows into every rank's word block."""
        ops, part, r = self.ops, self.part, self.rank
        vl, vp = part.v_loc, part.v_pad
        X_loc = OP[vp:]
        peer = self.px is not None
        work = None
        if gather_words and not words_mirrored:
            if peer:
                from . import _native
                self._before_write(op)
                mine = OP[r * vl:(r + 1) * vl]
                mine.copy_(X_loc[:vl])
                with torch.cuda.device(self.dev):
                    _native.check(_native.load().tgcn_peer_push(
                        mine.data_ptr(), self.px.peer_arrays[op], self.world, r, mine.numel() * 4,
                        mine.data_ptr() - OP.data_ptr(), self.px.multicast[op] or None, torch.cuda.current_stream().cuda_stream))
            elif self.world > 1:
                work = self.dist.all_gather_into_tensor(OP[:vp], X_loc[:vl], async_op=self.overlap)
            else:
                OP[:vl].copy_(X_loc[:vl])
        raw = None
        if qgraph is not None:
            # this rank's contribution to EVERY word row, from its own documents
            if peer:
                self._before_write(sl)
                ops.spmm(qgraph, X_loc[vl:], F=F, plan=qplan, out=self.X[sl].view(vp, F),
                         scatter=dict(bases=self.bases[sl], rows=vl, row0=r * vl))       # all-to-all in the epilogue's stores
                raw = self.X[sl]
            else:
                ops.spmm(qgraph, X_loc[vl:], F=F, plan=qplan, out=self.Q[F])             # runs while the word block is gathered
                if work is not None and self.overlap:
                    work.wait()
                    work = None
                if self.world > 1:
                    self.dist.all_to_all_single(self.X[sl].view(vp, F), self.Q[F])
                    raw = self.X[sl]
                else:
                    raw = self.Q[F].view(1, vp, F)
        if work is not None and self.overlap:
            work.wait()
        if peer and (gather_words or qgraph is not None):
            self._barrier()                   # word rows and slots of every rank are in place
        self._mark("exchange_F%d" % F)
        if tc is not None:
            out = ops.spmm_hybrid(tc, OP, F=F, plan=tc.remainder.plan(), raw_slots=raw, **kw)
        else:
            out = ops.spmm(graph, OP, F=F, plan=plan, raw_slots=raw, **kw)
        if gather_words and op is not None:
            self._note_read(op)
        if qgraph is not None and sl is not None:
            self._note_read(sl)
        return out

    def _gather_w1(self) -> None:      # nothing is kept gathered between steps
        return

    def _forward(self, training: bool) -> None:
        ops = self.ops
        self._mark("begin")
        drop = training and self.p > 0
        dkw = dict(drop_mode=ops.DROP_PHILOX if drop else ops.DROP_NONE, drop_p=self.p, philox_seed=self.seed,
                   philox_offset_dev=self.step_dev if drop else None, row_id_offset=self.rank * self.part.n_loc)
        fused_drop = False
        if training and self.share_h1 and self._h1_valid:
            h, fused_drop = (self.H1, True) if drop else (self.H1d, False)
        else:
            h = self.H1d if training else self.H1
            self._propagate(self.X["OPW"], self.H, "OPW", "SLf", self.shard, self.plan, self.qshard, self.plan_q, True,
                            words_mirrored=self._w1_words_ready, tc=self.tc, out=h, bias=self.b1, **dkw)
            self._mark("spmm_wide_fwd")
        opc, P_loc = ("OPCt", self.Pt_loc) if training else ("OPCe", self.Pe_loc)
        mir, mrows = self._mirror_of(opc)
        if mir is not None:
            self._before_write(opc)
        if fused_drop:
            ops.project(h, self.W2, K=self.H, out=P_loc, dropped_out=self.H1d, mirror=mir, mirror_rows=mrows, **dkw)
        else:
            ops.project(h, self.W2, K=self.H, out=P_loc, mirror=mir, mirror_rows=mrows)
        self._mark("project")
        self._propagate(self.X[opc], self.Cp, opc, None, self.shard, self.plan_z2, None, None, True,
                        words_mirrored=mir is not None, out=self.Z2, bias=self.b2)
        self._mark("spmm_narrow_fwd")

    def train_step(self) -> None:
        ops = self.ops
        self._forward(True)
        ops.masked_nll(self.Z2, self.C, self.y, self.train_mask, self.n_train, want_grad=True, dZ=self.dZ2_loc,
                       loss_out=self.loss_buf, workspace=self._nll_ws, partial=self.loss_part)
        self._mark("masked_nll")
        self._propagate(self.OPZ, self.Cp, None, "SLc", self.shard_g2, self.plan_g2, self.qshard_g2, self.plan_q_g2,
                        False, out=self.G2)
        self._mark("spmm_narrow_bwd")
        drop = self.p > 0
        mir, mrows = self._mirror_of("OPD")
        if mir is not None:
            self._before_write("OPD")
        self._before_write("small")
        r = ops.dense_bwd(self.G2, self.H1d, self.W2, self.dZ2_loc, H=self.H, n_classes=self.C,
                          drop_mode=ops.DROP_PHILOX if drop else ops.DROP_NONE, drop_p=self.p, philox_seed=self.seed,
                          philox_offset_dev=self.step_dev if drop else None, row_offset=self.rank * self.part.n_loc,
                          dZ1=self.dZ1_loc, workspace=self._db_ws, dW2=self.l_W2, db_hidden=self.l_b1, db_out=self.l_b2,
                          dZ1_mirror=mir, dZ1_mirror_rows=mrows)
        self._db_ws = r["workspace"]
        self._mark("dense_bwd")
        self._all_reduce_small()
        self._mark("allreduce_small_grads")
        kw = dict(lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, amsgrad=self.amsgrad, step_dev=self.step_dev)
        # dW1 rows of this rank never leave registers: Adam on W1[own rows] runs in the SpMM epilogue, and the updated
        # word rows go straight into every rank's W1 operand buffer (multimem.st) for the next forward
        ops.adam_prepare(self.step_dev, self.adam_hyper, self.lr, self.betas[0], self.betas[1])
        wmir, wrows = self._mirror_of("OPW")
        if wmir is not None:
            self._before_write("OPW")
        self._propagate(self.X["OPD"], self.H, "OPD", "SLb", self.shard, self.plan, self.qshard, self.plan_q, True,
                        words_mirrored=mir is not None, tc=self.tc, out=self.g_W1, want_out=self.keep_w1_grad,
                        adam=dict(param=self.W1_loc, exp_avg=self.st[0][0], exp_avg_sq=self.st[0][1], max_exp_avg_sq=self.st[0][2],
                                  hyper=self.adam_hyper, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                                  mirror=wmir, mirror_rows=wrows))
        self._w1_words_ready = wmir is not None
        self._mark("spmm_wide_bwd")
        ops.adam_step_small([self.b1, self.W2, self.b2], [self.g_b1, self.g_W2, self.g_b2], [s_[0] for s_ in self.st[1:]],
                            [s_[1] for s_ in self.st[1:]], [s_[2] for s_ in self.st[1:]], **kw)
        self._mark("adam")
        self._h1_valid = False

    def gathered_parameters(self) -> Dict[str, torch.Tensor]:
        full = torch.zeros((self.part.n_pad, self.H), dtype=torch.float32, device=self.dev)
        if self.world > 1:
            self.dist.all_gather_into_tensor(full, self.W1_loc)
        else:
            full.copy_(self.W1_loc)
        return {"layers.0.weight": self.part.to_old(full), "layers.0.bias": self.b1, "layers.1.weight": self.W2,
                "layers.1.bias": self.b2}

    def bytes_per_train_step(self) -> int:
        """Bytes each rank RECEIVES per epoch's train step: per hidden-wide propagation the word block and the partial
        word rows, per class-wide propagation one of the two."""
        P, vl = self.world, self.part.v_loc
        wide = 2 * (P - 1) * vl * self.H * 4
        narrow = (P - 1) * vl * self.Cp * 4
        return 2 * wide + 2 * narrow + self.n_small_pad * 4
