"""Multi-GPU TextGCN: 1D row partition of A_hat with NCCL all-gathers between layers
(BASELINE.json north_star "row-partitioned across the 8 GPUs of one box").

Partition.  Rows (= nodes) are dealt to the P ranks in snake order of decreasing nnz, so every
rank owns ceil(N/P) rows AND (almost exactly) nnz/P non-zeros -- word rows are ~5x heavier than
document rows, so contiguous ranges cannot balance both.  Nodes are renumbered so that rank r
owns the contiguous id range [r*n_loc, (r+1)*n_loc) of a padded id space of N_pad = P*n_loc
ids: every all-gather is then a plain equal-count ncclAllGather straight into the operand
buffer the next SpMM reads, no packing kernels.  Because X = I, W1's rows are nodes: W1, its
gradient and its Adam state are sharded by the same partition with no extra traffic.

Per train step (rank r owns rows R_r):
    all_gather(W1[R_r])  -> SpMM(F=H) + bias/dropout + fused projection   (big: N*H*4 bytes)
    all_gather(P[R_r])   -> SpMM(F=C) + bias -> masked NLL (global divisor) (small)
    all_gather(dZ2[R_r]) -> SpMM(F=C)                                     (small)
    dense backward on R_r -> all_reduce(dW2, db1, db2)                     (tiny)
    all_gather(dZ1[R_r]) -> SpMM(F=H) = dW1[R_r]                           (big)
    Adam on the W1 shard (local) and on the replicated small parameters (identical on all ranks)
A_hat is symmetric for TextGCN graphs, so the backward needs no transpose / reduce-scatter.
"""
from __future__ import annotations

import json
import os
import time
from typing import Dict, Optional

import torch

from .graph import GraphCSR


# --------------------------------------------------------------------------------------
# host-side partition logic (device-agnostic torch ops; covered by the gloo CPU tests)
# --------------------------------------------------------------------------------------
class RowPartition:
    """Snake-order row partition + node renumbering."""

    def __init__(self, row_nnz: torch.Tensor, world: int):
        n = int(row_nnz.numel())
        self.n, self.world = n, world
        self.n_loc = (n + world - 1) // world
        self.n_pad = self.n_loc * world
        dev = row_nnz.device
        order = torch.sort(row_nnz.to(torch.int64), descending=True, stable=True).indices   # heavy rows first
        k = torch.arange(n, device=dev)
        blk, pos = k // world, k % world
        rank = torch.where(blk % 2 == 0, pos, world - 1 - pos)          # snake: 0..P-1, P-1..0, ...
        new_of_sorted = rank * self.n_loc + blk
        self.new_id = torch.empty(n, dtype=torch.int64, device=dev)
        self.new_id[order] = new_of_sorted                                # old id -> new (padded) id
        self.old_id = torch.full((self.n_pad,), -1, dtype=torch.int64, device=dev)
        self.old_id[self.new_id] = torch.arange(n, device=dev)           # new id -> old id, -1 = padding
        self.row_nnz = row_nnz

    def rows_of(self, rank: int) -> torch.Tensor:
        """Old ids of the rows rank owns, in local order (-1 for padding rows)."""
        return self.old_id[rank * self.n_loc:(rank + 1) * self.n_loc]

    def nnz_of(self, rank: int) -> int:
        r = self.rows_of(rank)
        return int(self.row_nnz[r[r >= 0]].sum().item())

    def to_new(self, x: torch.Tensor, fill=0) -> torch.Tensor:
        """Permute a per-node tensor [N, ...] into the padded new numbering [N_pad, ...]."""
        out = torch.full((self.n_pad,) + tuple(x.shape[1:]), fill, dtype=x.dtype, device=x.device)
        out[self.new_id.to(x.device)] = x
        return out

    def to_old(self, x: torch.Tensor) -> torch.Tensor:
        """Inverse of to_new for a [N_pad, ...] tensor."""
        return x[self.new_id.to(x.device)]


def shard_csr(rowptr: torch.Tensor, colidx: torch.Tensor, val: torch.Tensor, part: RowPartition, rank: int):
    """Rows of rank `rank` out of the global CSR, columns renumbered into the padded id space.
    Returns (rowptr_loc int32[n_loc+1], colidx_loc int32, val_loc fp32).  Entry order inside a row
    is kept."""
    dev = rowptr.device
    rows = part.rows_of(rank).to(dev)
    valid = rows >= 0
    safe = rows.clamp_min(0)
    rp = rowptr.to(torch.int64)
    counts = torch.where(valid, rp[safe + 1] - rp[safe], torch.zeros_like(safe))
    loc_rowptr = torch.zeros(part.n_loc + 1, dtype=torch.int64, device=dev)
    loc_rowptr[1:] = torch.cumsum(counts, 0)
    total = int(loc_rowptr[-1].item())
    starts_loc = torch.repeat_interleave(loc_rowptr[:-1], counts)
    starts_glob = torch.repeat_interleave(rp[safe], counts)
    idx = torch.arange(total, device=dev) - starts_loc + starts_glob
    new_id = part.new_id.to(dev)
    col_loc = new_id[colidx[idx].to(torch.int64)].to(torch.int32)
    return loc_rowptr.to(torch.int32), col_loc.contiguous(), val[idx].contiguous()


def shard_graph(graph: GraphCSR, part: RowPartition, rank: int) -> GraphCSR:
    rp, ci, v = shard_csr(graph.rowptr, graph.colidx, graph.val, part, rank)
    g = GraphCSR(part.n_loc, rp, ci, v, None, None, n_cols=part.n_pad)
    g._symmetric = True      # the shard is only ever used as rows of the (symmetric) global matrix
    return g


# --------------------------------------------------------------------------------------
# exchange over NVLink peer memory (symmetric allocations + multicast / peer stores + barrier)
# --------------------------------------------------------------------------------------
class PeerExchange:
    """Symmetric buffers whose peer (and, with NVSwitch multicast, multicast) mappings are handed to
    `tgcn_peer_push`.  An exchange = push my slice into every rank's buffer + device-side barrier;
    it replaces ncclAllGather between layers and runs on the compute stream (graph-capturable)."""

    def __init__(self, group, rank: int, world: int, dev: torch.device):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem, self.C = symm_mem, C
        self.group, self.rank, self.world, self.dev = group, rank, world, dev
        self.handles = {}
        self.peer_arrays = {}
        self.multicast = {}
        self._barrier_hdl = None

    def alloc(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        t = self.symm_mem.empty(tuple(shape), dtype=dtype, device=self.dev)
        hdl = self.symm_mem.rendezvous(t, self.group)
        t.zero_()
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        self.handles[name] = hdl
        self.peer_arrays[name] = (self.C.c_void_p * self.world)(*ptrs)
        mc = 0
        try:
            mc = int(hdl.multicast_ptr)
        except Exception:
            mc = 0
        self.multicast[name] = mc
        if self._barrier_hdl is None:
            self._barrier_hdl = hdl
        return t

    def barrier(self) -> None:
        self._barrier_hdl.barrier(channel=0)

    def push(self, name: str, src: torch.Tensor, dst_offset_bytes: int) -> None:
        """Store `src` (a contiguous slice living in MY copy of buffer `name`) into every peer's copy at
        the same offset, then barrier.  After it returns (in stream order) all ranks see all slices."""
        from . import _native
        lib = _native.load()
        with torch.cuda.device(self.dev):
            stream = torch.cuda.current_stream().cuda_stream
            _native.check(lib.tgcn_peer_push(src.data_ptr(), self.peer_arrays[name], self.world, self.rank,
                                             src.numel() * src.element_size(), dst_offset_bytes,
                                             self.multicast[name] or None, stream))
        self.barrier()


# --------------------------------------------------------------------------------------
# distributed trainer (one process per GPU, torch.distributed over NCCL)
# --------------------------------------------------------------------------------------
class DistTextGCNTrainer:
    """Row-partitioned counterpart of TextGCNTrainer.  Every rank builds the same global graph
    (same synthetic generator / same Data), keeps its row shard, and owns W1[R_r] + Adam state."""

    def __init__(self, g, n_classes: int, hidden: int, dropout: float, lr: float, amsgrad: bool,
                 rank: int, world: int, dev: torch.device, seed: int = 0, betas=(0.9, 0.999), eps: float = 1e-8,
                 graph: Optional[GraphCSR] = None, init_weights: Optional[Dict[str, torch.Tensor]] = None,
                 use_cuda_graph: bool = False, exchange: str = "peer", fused_stores: bool = True,
                 fuse_adam: bool = True, keep_w1_grad: bool = True, share_h1: bool = True, restrict_rows: bool = True,
                 tensor_cores: Optional[bool] = None, tc_min_density: float = 0.05):
        import torch.distributed as dist
        from . import ops
        from .graph import upload_graph
        self.dist, self.ops = dist, ops
        self.rank, self.world, self.dev = rank, world, dev
        from .models import decode_features
        n = int(g.x.shape[0])
        feat = decode_features(g.x, getattr(g, "n_vocab", None))
        if feat is None:
            raise RuntimeError("the row-partitioned trainer expects featureless input x = I or [I | F] (text2graph.py:226-246)")
        # x = [I | F] (perlevel_dbpedia.py:140-141): W1 has c_prev extra rows.  They are replicated on every rank; the
        # exchanged layer-1 operand is then X W1 = W1[:N] + F W1[N:] (own rows computed locally) instead of W1 itself.
        self.c_prev = 0 if feat.Fdoc is None else int(feat.Fdoc.shape[1])
        self.hier = self.c_prev > 0
        if hidden % 4 != 0:
            raise NotImplementedError("hidden width must be a multiple of 4")
        self.n, self.H, self.C, self.Cp = n, hidden, n_classes, ops.pad4(n_classes)
        self.p, self.lr, self.amsgrad, self.betas, self.eps, self.seed = dropout, lr, amsgrad, betas, eps, seed
        full = graph if graph is not None else upload_graph(g.edge_index.to(dev), g.edge_attr.to(dev), n)
        if not full.is_symmetric():
            raise NotImplementedError("the row-partitioned trainer needs a symmetric A_hat (Text2GraphTransformer graphs are)")
        row_nnz = (full.rowptr[1:] - full.rowptr[:-1]).to(torch.int64)
        self.part = RowPartition(row_nnz, world)
        self.shard = shard_graph(full, self.part, rank)
        self.nnz_global = full.nnz
        del full
        self.plan = self.shard.plan()
        nl, npad, H, Cp = self.part.n_loc, self.part.n_pad, hidden, self.Cp
        # hybrid hidden-wide propagation on the shard (csrc/spmm_tc.cu; same rule as TextGCNTrainer: L2-resident operand)
        self.tc = None
        auto = tensor_cores is None and self.shard.nnz >= 200_000 and npad * ((H + 15) // 16 * 16) * 8 <= (104 << 20)
        if (tensor_cores or auto) and H % 4 == 0 and 64 <= H <= 256:
            from .tc_plan import build_tc_plan
            tc = build_tc_plan(self.shard, min_density=tc_min_density, width=H,
                               n_sms=torch.cuda.get_device_properties(dev).multi_processor_count)
            if tc is not None and (tensor_cores or tc.nnz_dense >= 0.15 * self.shard.nnz):
                self.tc = tc
        f32 = dict(dtype=torch.float32, device=dev)
        # exchange buffers: symmetric (peer-mapped) allocations when the NVLink path is available.  The decision is
        # COLLECTIVE: every rank tries to allocate all of them, the outcomes are all-reduced (MIN), and either every
        # rank takes the peer path or every rank takes the NCCL path -- a rank falling back alone would deadlock the
        # others in the device barrier.
        self.exchange, self.exchange_error, self.px = "nccl", None, None
        n_small = H + H * n_classes + n_classes + self.c_prev * H            # packed grads of b1, W2, b2 (+ the [I|F] tail of W1)
        self.n_small = n_small
        self.n_small_pad = (n_small + 3) // 4 * 4
        # Propagate-first order of layer 2 when there are more classes than hidden units (perlevel_dbpedia.py: 70 / 219
        # classes, hidden 32; see TextGCNTrainer.propagate_first): Z2 = (A_hat H1d) W2 + b2, dH1d = A_hat (dZ2 W2^T),
        # dW2 = (A_hat H1d)^T dZ2 -- every exchanged operand is then hidden-wide (N x 32 instead of N x 220).
        self.propagate_first = self.Cp > H
        if self.propagate_first:
            xshapes = {"W1": (npad, H), "small": (world, self.n_small_pad), "Ht": (npad, H), "He": (npad, H),
                       "T": (npad, H), "dZ1": (npad, H)}
        else:
            xshapes = {"W1": (npad, H), "small": (world, self.n_small_pad), "Pt": (npad, Cp), "Pe": (npad, Cp),
                       "dZ2": (npad, Cp), "dZ1": (npad, H)}
        xbufs: Dict[str, torch.Tensor] = {}
        if world > 1 and exchange == "peer":
            ok = 1
            try:
                self.px = PeerExchange(dist.group.WORLD, rank, world, dev)
                for name, shp in xshapes.items():
                    xbufs[name] = self.px.alloc(name, shp)
            except Exception as e:          # symmetric memory not available in this stack: NCCL collectives
                self.exchange_error = repr(e)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                self.exchange = "peer"
            else:
                self.px, xbufs = None, {}
                if self.exchange_error is None:
                    self.exchange_error = "symmetric allocation failed on another rank"

        def xbuf(name, shape):
            assert tuple(shape) == tuple(xshapes[name])
            return xbufs[name] if name in xbufs else torch.zeros(tuple(shape), **f32)
        # parameters: same init on every rank (same seed), W1 kept in the NEW row order
        gen = torch.Generator().manual_seed(seed)
        if init_weights is None:
            a1, a2 = (6.0 / (n + self.c_prev + H)) ** 0.5, (6.0 / (H + n_classes)) ** 0.5
            W1 = (torch.rand(n + self.c_prev, H, generator=gen) * 2 - 1) * a1
            W2 = (torch.rand(H, n_classes, generator=gen) * 2 - 1) * a2
            b1, b2 = torch.zeros(H), torch.zeros(n_classes)
        else:
            W1, b1, W2, b2 = (init_weights[k].detach().cpu().float() for k in
                              ("layers.0.weight", "layers.0.bias", "layers.1.weight", "layers.1.bias"))
        self.W1_full = xbuf("W1", (npad, H))                                # [N_pad, H]; rows of other ranks are gathered
        lo = rank * nl
        if not self.hier:
            self.W1_full.copy_(self.part.to_new(W1[:n]).to(dev))
            self.W1_loc = self.W1_full[lo:lo + nl]                           # view: this rank's shard (authoritative)
            self.W1_cat = None
        else:
            # own rows of W1 followed by the replicated tail; W1_full holds the operand X W1 (see _gather_w1)
            self.W1_cat = torch.cat([self.part.to_new(W1[:n])[lo:lo + nl], W1[n:]]).to(dev).contiguous()
            self.W1_loc = self.W1_cat
            F_full = torch.zeros((n, self.c_prev), dtype=torch.float32)
            F_full[feat.n_vocab:] = feat.Fdoc.detach().cpu()
            self.F_loc = self.part.to_new(F_full)[lo:lo + nl].to(dev).contiguous()
            self.XW_loc = self.W1_full[lo:lo + nl]
        self.b1, self.W2, self.b2 = b1.to(dev), W2.to(dev).contiguous(), b2.to(dev)
        self.small_slots = xbuf("small", (world, self.n_small_pad))          # slot r = rank r's partial sums
        self.small_local = self.small_slots[rank]                             # dense_bwd writes my slot in place
        self.small = torch.zeros(self.n_small_pad, **f32)                     # summed over ranks
        def views(buf):
            o = H + H * n_classes + n_classes
            return buf[:H], buf[H:H + H * n_classes].view(H, n_classes), buf[H + H * n_classes:o], buf[o:n_small].view(self.c_prev, H)
        self.l_b1, self.l_W2, self.l_b2, self.l_tail = views(self.small_local)   # local partials (kernel outputs)
        self.g_b1, self.g_W2, self.g_b2, self.g_tail = views(self.small)         # global gradients (Adam inputs)
        self.g_W1 = torch.zeros((nl + self.c_prev, H), **f32)
        def state(t):
            return [torch.zeros_like(t), torch.zeros_like(t), torch.zeros_like(t) if amsgrad else None]
        self.st = [state(self.W1_loc), state(self.b1), state(self.W2), state(self.b2)]
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.adam_hyper = torch.zeros(2, dtype=torch.float32, device=dev)
        self.fuse_adam, self.keep_w1_grad = bool(fuse_adam) and not self.hier, bool(keep_w1_grad) or self.hier
        # activations
        # pre-dropout hidden rows of the last eval forward, reused by the next train step (same W1/b1,
        # flat_amazon.py:100-110): one hidden-wide SpMM less per epoch and per rank, bit-identical (trainer.py)
        self.share_h1 = bool(share_h1)
        self._h1_valid = False
        self.Z2 = torch.zeros((nl, Cp), **f32)
        if self.propagate_first:
            # the hidden rows themselves are exchanged: they live in this rank's slice of the exchange buffers
            self.Ht_full, self.He_full = xbuf("Ht", (npad, H)), xbuf("He", (npad, H))
            self.H1d = self.Ht_full[lo:lo + nl]
            self.H1 = self.He_full[lo:lo + nl] if (self.share_h1 and dropout > 0) else self.H1d
            self.T_full = xbuf("T", (npad, H))           # dZ2 W2^T
            self.T_loc = self.T_full[lo:lo + nl]
            self.U = torch.zeros((nl, H), **f32)         # A_hat H1d on the own rows
            self.W2t = torch.zeros((n_classes, H), **f32)
            self._cs_ws = torch.empty(4096 * H * 4, dtype=torch.uint8, device=dev)
            self.dZ2_loc = torch.zeros((nl, Cp), **f32)
            self._h_exchanged = None
        else:
            self.H1d = torch.empty((nl, H), **f32)
            self.H1 = torch.empty((nl, H), **f32) if (self.share_h1 and dropout > 0) else self.H1d
            self.Pt_full = xbuf("Pt", (npad, Cp))        # projected rows, train forward
            self.Pt_loc = self.Pt_full[lo:lo + nl]
            self.Pe_full = xbuf("Pe", (npad, Cp))        # projected rows, eval forward (double buffer)
            self.Pe_loc = self.Pe_full[lo:lo + nl]
            self.dZ2_full = xbuf("dZ2", (npad, Cp))
            self.dZ2_loc = self.dZ2_full[lo:lo + nl]
            self.G2 = torch.zeros((nl, Cp), **f32)
        self.dZ1_full = xbuf("dZ1", (npad, H))
        self.dZ1_loc = self.dZ1_full[lo:lo + nl]
        self.loss_part = torch.zeros(2, dtype=torch.float64, device=dev)        # train step: sum nll, count (local rows)
        self.loss_part_val = torch.zeros(2, dtype=torch.float64, device=dev)    # eval step
        self.loss_buf = torch.zeros(2, **f32)
        self.stats = torch.zeros(6, dtype=torch.float64, device=dev)         # packed scalars for one all-reduce
        self.pred = torch.zeros(nl, dtype=torch.int32, device=dev)
        self.correct = torch.zeros(1, dtype=torch.int32, device=dev)
        self._nll_ws = torch.empty(2 * ((nl * 4 + 255) // 256 * 256) + 4096, dtype=torch.uint8, device=dev)
        self._db_ws = None
        # labels / masks in the new order, local slices
        y_new = self.part.to_new(g.y.cpu(), 0)
        tm_new = self.part.to_new(g.train_mask.cpu(), False)
        vm_new = self.part.to_new(g.val_mask.cpu(), False)
        self.y = y_new[lo:lo + nl].to(dev).contiguous()
        self.train_mask = tm_new[lo:lo + nl].to(dev).contiguous()
        self.val_mask = vm_new[lo:lo + nl].to(dev).contiguous()
        self.n_train, self.n_val = int(g.train_mask.sum()), int(g.val_mask.sum())
        used = g.train_mask.cpu() | g.val_mask.cpu()
        if bool((used & ((g.y.cpu() < 0) | (g.y.cpu() >= n_classes))).any()):
            raise RuntimeError(f"labels of masked rows must lie in [0, {n_classes})")
        # class-wide propagations restricted to what the masked loss reads (see TextGCNTrainer.restrict_rows)
        self.restrict_rows = bool(restrict_rows)
        self.plan_z2, self.shard_g2, self.plan_g2 = self.plan, self.shard, self.plan
        if self.restrict_rows:
            test_new = self.part.to_new(g.test_mask.cpu(), False) if getattr(g, "test_mask", None) is not None else torch.zeros_like(tm_new)
            rows_loc = (tm_new | vm_new | test_new)[lo:lo + nl].to(dev)
            self.plan_z2 = self.shard.plan_for_rows(rows_loc, self.plan)
            self.shard_g2 = self.shard.select_columns(tm_new.to(dev))
            self.plan_g2 = self.shard_g2.plan_nonempty()
        self.w1_stale = self.hier     # x = I: every rank initialised the full W1 identically; [I|F]: X W1 is built by the first gather
        self._w1_mirrored = False
        self._pending_reads = set()
        self.fused_stores = bool(fused_stores and self.px is not None and
                                 all(self.px.multicast.get(k, 0) for k in xshapes if k != "small"))
        self._w1_mirror_ok = not self.hier
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._eager_epochs = 0
        self.graph_error = None
        self.profile = None           # list of (name, event) when per-phase timing is on
        self.launches_per_epoch = 0
        # every rank has finished zeroing / filling its symmetric buffers before any peer may store into them
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # ---- optional per-phase timing (eager mode only; used by the scaling analysis in DESIGN.md) ----
    def _mark(self, name: str) -> None:
        if self.profile is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.profile.append((name, ev))

    def phase_times_ms(self) -> Dict[str, float]:
        """Sum of the time between consecutive marks, keyed by the phase that ENDS at the mark."""
        torch.cuda.synchronize(self.dev)
        out: Dict[str, float] = {}
        for (n0, e0), (n1, e1) in zip(self.profile[:-1], self.profile[1:]):
            if n1 != "begin":
                out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out

    def _wide_spmm(self, B: torch.Tensor, **kw):
        """Hidden-wide propagation over this rank's rows: hybrid (tensor-core tiles + gathered remainder) when planned."""
        if self.tc is not None:
            return self.ops.spmm_hybrid(self.tc, B, F=self.H, plan=self.tc.remainder.plan(), **kw)
        return self.ops.spmm(self.shard, B, F=self.H, plan=self.plan, **kw)

    # ---- collectives ----
    # One-sided stores need two guarantees the two-sided ncclAllGather gave for free:
    #  (1) visibility: a barrier after the producers, before any rank reads the exchanged buffer;
    #  (2) no overwrite of a buffer a slower rank may still be reading: at least one barrier must lie
    #      between a buffer's last consumer and its next producer.  `_pending_reads` tracks (2) on the
    #      host -- identically on every rank, the control flow is the same -- and inserts an extra
    #      barrier only when none happened in between (never in the steady-state epoch: P is
    #      double-buffered between the train and the eval forward for exactly that reason).
    def _barrier(self) -> None:
        self.px.barrier()
        self._pending_reads.clear()

    def _before_write(self, name: str) -> None:
        if self.px is not None and name in self._pending_reads:
            self._barrier()

    def _note_read(self, name: str) -> None:
        if self.px is not None:
            self._pending_reads.add(name)

    def _mirror(self, name: str, loc: torch.Tensor, full: torch.Tensor) -> Optional[int]:
        """Multicast address of `loc` inside symmetric buffer `name` (None when the producer cannot fuse)."""
        if self.px is None or not self.fused_stores:
            return None
        return self.px.multicast[name] + (loc.data_ptr() - full.data_ptr())

    def _exchange(self, full: torch.Tensor, loc: torch.Tensor, name: str, produced_with_mirror: bool) -> None:
        """Make every rank's slice of `name` visible everywhere."""
        if self.world == 1:
            if full.data_ptr() != loc.data_ptr():
                full[:loc.shape[0]].copy_(loc)
            return
        if self.px is None:
            self.dist.all_gather_into_tensor(full, loc)          # in place: loc is the rank-th slice of full
            return
        if not produced_with_mirror:                              # separate push kernel (peer or multicast stores)
            from . import _native
            lib = _native.load()
            with torch.cuda.device(self.dev):
                _native.check(lib.tgcn_peer_push(loc.data_ptr(), self.px.peer_arrays[name], self.world, self.rank,
                                                 loc.numel() * loc.element_size(), loc.data_ptr() - full.data_ptr(),
                                                 self.px.multicast[name] or None, torch.cuda.current_stream().cuda_stream))
        self._barrier()

    def _all_reduce_small(self) -> None:
        """self.small = sum over ranks of the packed (db1, dW2, db2) partials."""
        if self.world == 1:
            self.small.copy_(self.small_local)
        elif self.px is not None:
            from . import _native
            lib = _native.load()
            self._exchange(self.small_slots, self.small_local, "small", False)
            with torch.cuda.device(self.dev):
                _native.check(lib.tgcn_sum_slots(self.small_slots.data_ptr(), self.world, self.n_small_pad, self.n_small_pad,
                                                 self.small.data_ptr(), torch.cuda.current_stream().cuda_stream))
            self._note_read("small")
        else:
            self.small.copy_(self.small_local)
            self.dist.all_reduce(self.small)

    def _gather_w1(self) -> None:
        if self.w1_stale:
            if self.hier:      # own rows of X W1 = W1[own] + F[own] W1[N:], written into this rank's slice of the exchange buffer
                self._before_write("W1")
                self.ops.hier_forward(self.W1_cat, self.part.n_loc, 0, self.F_loc, out=self.XW_loc)
                self._exchange(self.W1_full, self.XW_loc, "W1", False)
            else:
                self._exchange(self.W1_full, self.W1_loc, "W1", self._w1_mirrored)
            self.w1_stale = False

    def _forward_propagate_first(self, training: bool) -> None:
        """Layer 2 as (A_hat H1d) W2 + b2: the hidden rows are exchanged (N x H), the class-wide product stays local."""
        ops = self.ops
        self._mark("begin")
        drop = training and self.p > 0
        dkw = dict(drop_mode=ops.DROP_PHILOX if drop else ops.DROP_NONE, drop_p=self.p, philox_seed=self.seed,
                   philox_offset_dev=self.step_dev if drop else None, row_id_offset=self.rank * self.part.n_loc)
        h = self.H1d if training else self.H1
        hname, hfull = ("Ht", self.Ht_full) if h is self.H1d else ("He", self.He_full)
        need_exchange = True
        if training and self.share_h1 and self._h1_valid:
            if drop:
                self._before_write(hname)
                ops.dropout_apply(self.H1, F=self.H, out=self.H1d, **dkw)
            else:
                need_exchange = False          # H1 is H1d and every rank's slice is already in place from the eval forward
            self._mark("dropout_apply")
        else:
            self._gather_w1()
            self._mark("allgather_W1")
            self._before_write(hname)
            self._wide_spmm(self.W1_full, out=h, bias=self.b1, **dkw)
            self._note_read("W1")
            self._mark("spmm_wide_fwd")
        if need_exchange:
            self._exchange(hfull, h, hname, False)
        self._mark("allgather_P")
        ops.spmm(self.shard, hfull, F=self.H, plan=self.plan_z2, out=self.U)
        self._note_read(hname)
        ops.project(self.U, self.W2, K=self.H, out=self.Z2, bias=self.b2)
        self._mark("spmm_narrow_fwd")

    def _train_step_propagate_first(self) -> None:
        ops = self.ops
        self._forward_propagate_first(True)
        ops.masked_nll(self.Z2, self.C, self.y, self.train_mask, self.n_train, want_grad=True, dZ=self.dZ2_loc,
                       loss_out=self.loss_buf, workspace=self._nll_ws, partial=self.loss_part)
        self._mark("masked_nll")
        drop = self.p > 0
        dkw = dict(drop_mode=ops.DROP_PHILOX if drop else ops.DROP_NONE, drop_p=self.p, philox_seed=self.seed,
                   philox_offset_dev=self.step_dev if drop else None, row_id_offset=self.rank * self.part.n_loc)
        self.W2t.copy_(self.W2.t())
        self._before_write("T")
        ops.project(self.dZ2_loc, self.W2t, K=self.C, out=self.T_loc)                 # T = dZ2 W2^T (zero off the train rows)
        self._exchange(self.T_full, self.T_loc, "T", False)
        self._mark("allgather_dZ2")
        self._before_write("dZ1")
        ops.spmm(self.shard_g2, self.T_full, F=self.H, plan=self.plan_g2, out=self.dZ1_loc, **dkw)   # dZ1 = dropout'(A_hat T)
        self._note_read("T")
        self._mark("spmm_narrow_bwd")
        self._before_write("small")
        r = ops.dense_bwd(self.dZ2_loc, self.U, self.W2, self.dZ2_loc, H=self.H, n_classes=self.C, want_dz1=False,
                          workspace=self._db_ws, dW2=self.l_W2, db_out=self.l_b2)      # dW2 = U^T dZ2, db2 = colsum(dZ2)
        self._db_ws = r["workspace"]
        ops.colsum(self.dZ1_loc, F=self.H, out=self.l_b1, workspace=self._cs_ws)
        self._mark("dense_bwd")
        if not self.hier:
            self._all_reduce_small()
            self._mark("allreduce_small_grads")
        self._exchange(self.dZ1_full, self.dZ1_loc, "dZ1", False)
        self._mark("allgather_dZ1")
        self._wide_backward_and_adam()

    def _forward(self, training: bool) -> None:
        if self.propagate_first:
            return self._forward_propagate_first(training)
        ops = self.ops
        self._mark("begin")
        drop = training and self.p > 0
        dkw = dict(drop_mode=ops.DROP_PHILOX if drop else ops.DROP_NONE, drop_p=self.p, philox_seed=self.seed,
                   philox_offset_dev=self.step_dev if drop else None, row_id_offset=self.rank * self.part.n_loc)
        fused_drop = False
        if training and self.share_h1 and self._h1_valid:
            # the eval forward that preceded this step saw the same W1/b1: its rows only need this step's dropout mask,
            # applied inside the projection kernel below (which also writes the dropped rows for the backward pass)
            h, fused_drop = (self.H1, True) if drop else (self.H1d, False)
            self._mark("dropout_apply")
        else:
            self._gather_w1()
            self._mark("allgather_W1")
            h = self.H1d if training else self.H1
            self._wide_spmm(self.W1_full, out=h, bias=self.b1, **dkw)
            self._note_read("W1")
            self._mark("spmm_wide_fwd")
        pname = "Pt" if training else "Pe"                       # double-buffered, see _pending_reads
        P_full, P_loc = (self.Pt_full, self.Pt_loc) if training else (self.Pe_full, self.Pe_loc)
        self._before_write(pname)
        mir = self._mirror(pname, P_loc, P_full)
        if fused_drop:
            ops.project(h, self.W2, K=self.H, out=P_loc, mirror=mir, dropped_out=self.H1d, **dkw)
        else:
            ops.project(h, self.W2, K=self.H, out=P_loc, mirror=mir)     # layer 2's thin X W, stored to all ranks
        self._exchange(P_full, P_loc, pname, mir is not None)
        self._mark("allgather_P")
        ops.spmm(self.shard, P_full, F=self.Cp, plan=self.plan_z2, out=self.Z2, bias=self.b2)
        self._note_read(pname)
        self._mark("spmm_narrow_fwd")

    def train_step(self) -> None:
        if self.propagate_first:
            return self._train_step_propagate_first()
        ops = self.ops
        self._forward(True)
        self._before_write("dZ2")
        mir = self._mirror("dZ2", self.dZ2_loc, self.dZ2_full)
        ops.masked_nll(self.Z2, self.C, self.y, self.train_mask, self.n_train, want_grad=True, dZ=self.dZ2_loc,
                       loss_out=self.loss_buf, workspace=self._nll_ws, partial=self.loss_part, dZ_mirror=mir)
        self._mark("masked_nll")
        self._exchange(self.dZ2_full, self.dZ2_loc, "dZ2", mir is not None)
        self._mark("allgather_dZ2")
        ops.spmm(self.shard_g2, self.dZ2_full, F=self.Cp, plan=self.plan_g2, out=self.G2)
        self._note_read("dZ2")
        self._mark("spmm_narrow_bwd")
        drop = self.p > 0
        self._before_write("dZ1")
        self._before_write("small")
        mir = self._mirror("dZ1", self.dZ1_loc, self.dZ1_full)
        r = ops.dense_bwd(self.G2, self.H1d, self.W2, self.dZ2_loc, H=self.H, n_classes=self.C,
                          drop_mode=ops.DROP_PHILOX if drop else ops.DROP_NONE, drop_p=self.p, philox_seed=self.seed,
                          philox_offset_dev=self.step_dev if drop else None, row_offset=self.rank * self.part.n_loc,
                          dZ1=self.dZ1_loc, workspace=self._db_ws, dW2=self.l_W2, db_hidden=self.l_b1, db_out=self.l_b2,
                          dZ1_mirror=mir)
        self._db_ws = r["workspace"]
        self._mark("dense_bwd")
        if not self.hier:
            self._all_reduce_small()          # its barrier also publishes the mirrored dZ1 stores
            self._mark("allreduce_small_grads")
        dz1_mirrored = mir is not None and self.px is not None
        if not dz1_mirrored:
            self._exchange(self.dZ1_full, self.dZ1_loc, "dZ1", False)
        elif self.hier:
            self._barrier()                   # publishes the mirrored dZ1 stores (the small all-reduce comes later here)
        self._mark("allgather_dZ1")
        self._wide_backward_and_adam()

    def _wide_backward_and_adam(self) -> None:
        """dW1[own rows] = A_hat dZ1 on the shard + Adam on W1 (fused into the epilogue when x = I), Adam on b1 / W2 / b2."""
        ops = self.ops
        kw = dict(lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, amsgrad=self.amsgrad,
                  step_dev=self.step_dev)
        self._before_write("W1")
        mir = self._mirror("W1", self.W1_loc, self.W1_full)
        if self.fuse_adam:
            # dW1 rows of this shard never leave registers: Adam on W1[R_r] runs in the SpMM epilogue and the
            # updated rows go straight to every rank's copy of W1 (multimem.st) -- compute, optimiser and
            # exchange in one kernel
            ops.adam_prepare(self.step_dev, self.adam_hyper, self.lr, self.betas[0], self.betas[1])
            self._wide_spmm(self.dZ1_full, out=self.g_W1 if self.keep_w1_grad else None,
                            want_out=self.keep_w1_grad,
                            adam=dict(param=self.W1_loc, exp_avg=self.st[0][0], exp_avg_sq=self.st[0][1], max_exp_avg_sq=self.st[0][2],
                               hyper=self.adam_hyper, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, mirror=mir))
            self._note_read("dZ1")
            self._mark("spmm_wide_bwd")
        elif self.hier:
            # [I | F]: dW1[own rows] = A_hat dZ1 on the shard; dW1[N:] = F^T (A_hat dZ1)[docs], summed over the ranks together
            # with the other small gradients; Adam on (own rows ; replicated tail) in one launch, identical tail on all ranks
            nl_ = self.part.n_loc
            self._wide_spmm(self.dZ1_full, out=self.g_W1[:nl_])
            self._note_read("dZ1")
            self._mark("spmm_wide_bwd")
            ops.hier_backward(self.g_W1[:nl_], nl_, 0, self.F_loc, self.H, self.l_tail)
            self._all_reduce_small()
            self._mark("allreduce_small_grads")
            self.g_W1[nl_:].copy_(self.g_tail)
            ops.increment_step(self.step_dev)
            ops.adam_step(self.W1_cat, self.g_W1, *self.st[0], **kw)
            mir = None
        else:
            self._wide_spmm(self.dZ1_full, out=self.g_W1)
            self._note_read("dZ1")
            self._mark("spmm_wide_bwd")
            ops.increment_step(self.step_dev)
            ops.adam_step(self.W1_loc, self.g_W1, *self.st[0], param_mirror=mir, **kw)    # updated rows go to every rank
        self._w1_mirrored = mir is not None
        ops.adam_step_small([self.b1, self.W2, self.b2], [self.g_b1, self.g_W2, self.g_b2], [s_[0] for s_ in self.st[1:]],
                            [s_[1] for s_ in self.st[1:]], [s_[2] for s_ in self.st[1:]], **kw)
        self._mark("adam")
        self.w1_stale = True
        self._h1_valid = False

    def eval_step(self) -> None:
        ops = self.ops
        self._forward(False)
        ops.masked_nll(self.Z2, self.C, self.y, self.val_mask, max(self.n_val, 1), want_grad=False,
                       loss_out=self.loss_buf, workspace=self._nll_ws, pred=self.pred, correct=self.correct,
                       partial=self.loss_part_val)
        self._mark("masked_nll")
        self._h1_valid = self.share_h1

    def epoch(self) -> None:
        """train_step + eval_step; after two eager epochs the pair is captured (kernels AND the NCCL
        collectives) in one CUDA graph and replayed -- at 8 ranks the ~60 launches of an epoch would
        otherwise cost more host time than the GPUs need to run them."""
        if not self.use_cuda_graph:
            self.train_step()
            self.eval_step()
            return
        if self._graph is not None:
            self._graph.replay()
            return
        if self._eager_epochs < 2:
            from . import _native
            lib = _native.load()
            c0 = lib.tgcn_launch_count()
            self.train_step()
            self.eval_step()
            self.launches_per_epoch = int(lib.tgcn_launch_count() - c0)
            self._eager_epochs += 1
            return
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        try:
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                self.train_step()
                self.eval_step()
            self._graph = gph
        except Exception as e:      # capture of the collectives not supported in this stack: stay eager
            self.use_cuda_graph = False
            self.graph_error = repr(e)
            torch.cuda.synchronize(self.dev)
            self.train_step()
            self.eval_step()
            return
        self._graph.replay()

    def epoch_stats(self) -> Dict[str, float]:
        """Global (all-reduced) train loss of the last train step is not kept here; this returns the
        val loss / accuracy of the last eval_step."""
        s = torch.stack([self.loss_part_val[0], self.correct[0].double()])
        if self.world > 1:
            self.dist.all_reduce(s)
        v = s.cpu().tolist()
        return dict(val_loss=v[0] / max(self.n_val, 1), acc_val=v[1] / max(self.n_val, 1))

    def train_loss(self) -> float:
        s = self.loss_part[:1].clone()
        if self.world > 1:
            self.dist.all_reduce(s)
        return float(s.item()) / self.n_train

    def gathered_parameters(self) -> Dict[str, torch.Tensor]:
        """Parameters in the ORIGINAL node order (reference layout: layers.{i}.weight (in,out), bias)."""
        if self.hier:
            nl_ = self.part.n_loc
            full = torch.zeros((self.part.n_pad, self.H), dtype=torch.float32, device=self.dev)
            if self.world > 1:
                self.dist.all_gather_into_tensor(full, self.W1_cat[:nl_].contiguous())
            else:
                full[:nl_].copy_(self.W1_cat[:nl_])
            W1 = torch.cat([self.part.to_old(full), self.W1_cat[nl_:]])
        else:
            self._gather_w1()
            W1 = self.part.to_old(self.W1_full)
        return {"layers.0.weight": W1, "layers.0.bias": self.b1, "layers.1.weight": self.W2, "layers.1.bias": self.b2}

    def logits_old_order(self) -> torch.Tensor:
        """All-gathered logits of the last forward, original node order (host-facing helper)."""
        full = torch.zeros((self.part.n_pad, self.Cp), dtype=torch.float32, device=self.dev)
        if self.world > 1:
            self.dist.all_gather_into_tensor(full, self.Z2.contiguous())
        else:
            full[:self.Z2.shape[0]].copy_(self.Z2)
        return self.part.to_old(full)[:, :self.C]

    def bytes_per_train_step(self) -> int:
        """Bytes each rank RECEIVES per train step (+ the W1 gather that precedes the forward)."""
        P, nl = self.world, self.part.n_loc
        big = (P - 1) * nl * self.H * 4
        small = big if self.propagate_first else (P - 1) * nl * self.Cp * 4
        return 2 * big + 2 * small + 2 * self.n_small_pad * 4


def shutdown(trainer: Optional["DistTextGCNTrainer"] = None) -> None:
    """Orderly exit of a rank.  Graphs that captured NCCL collectives must be released before the
    communicator is torn down (destroy_process_group with such a graph alive hangs with NCCL 2.28 /
    torch 2.11); if teardown still stalls, leave without it -- the work is done and flushed."""
    import gc
    import sys
    import threading
    import torch.distributed as dist
    if trainer is not None:
        trainer._graph = None
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    done = threading.Event()

    def _destroy():
        try:
            dist.destroy_process_group()
        finally:
            done.set()
    t = threading.Thread(target=_destroy, daemon=True)
    t.start()
    if not done.wait(timeout=20):
        os._exit(0)


# --------------------------------------------------------------------------------------
# the same graph in the partition's node numbering (parity reference for the N-rank path)
# --------------------------------------------------------------------------------------
def renumbered_data(g, part: RowPartition):
    """`g` with its nodes renumbered into the partition's padded id space (padding ids = isolated nodes outside every
    mask), so a single-GPU TextGCNTrainer on it sees exactly the rows, the within-row entry order and the global
    Philox element indices the N ranks see: the single-GPU run is then the parity reference of the partitioned run."""
    from .data import Data
    new_id = part.new_id.cpu()
    npad = part.n_pad
    idx = torch.arange(npad, dtype=torch.int64)
    x = torch.sparse_coo_tensor(torch.stack([idx, idx]), torch.ones(npad), size=(npad, npad), check_invariants=False).coalesce()
    ei = new_id[g.edge_index.cpu().contiguous()]
    return Data(x=x, edge_index=ei, edge_attr=g.edge_attr.cpu(), y=part.to_new(g.y.cpu(), 0),
                train_mask=part.to_new(g.train_mask.cpu(), False), val_mask=part.to_new(g.val_mask.cpu(), False),
                test_mask=part.to_new(g.test_mask.cpu(), False), n_vocab=getattr(g, "n_vocab", 0))


def choose_partition(g, requested: str = "auto") -> str:
    """"row" (DistTextGCNTrainer: all N rows of an operand cross ranks) or "words" (dist_bipartite: the word block and
    the partial word rows, 2 V rows).  auto: the word-block scheme when x = I and it moves at most 3/4 of the rows."""
    if requested in ("row", "words"):
        return requested
    n, v = int(g.x.shape[0]), int(getattr(g, "n_vocab", 0) or 0)
    return "words" if (int(g.x.shape[1]) == n and 0 < v and 2 * v <= 0.75 * n) else "row"


def make_dist_trainer(g, shape, rank: int, world: int, dev: torch.device, partition: str = "auto", **kw):
    if choose_partition(g, partition) == "words":
        from .dist_bipartite import BipartiteTextGCNTrainer
        if not kw.get("fuse_adam", True):          # a row-partition switch (bench --no-fuse-adam); the word-block trainer has no unfused form
            kw = {k: v for k, v in kw.items() if k != "fuse_adam"}
        return BipartiteTextGCNTrainer(g, shape.n_classes, shape.hidden, shape.dropout, shape.lr, shape.amsgrad, rank, world, dev, **kw)
    return DistTextGCNTrainer(g, shape.n_classes, shape.hidden, shape.dropout, shape.lr, shape.amsgrad, rank, world, dev, **kw)


def parity_against_single_gpu(g, shape, rank: int, world: int, dev: torch.device, seed: int, epochs: int = 5,
                              partition: str = "row", **trainer_kw):
    """Runs `epochs` epochs of a FRESH N-rank trainer in its shipped configuration (CUDA graph from the third epoch on,
    multicast stores fused into the producers, dropout on) and, on rank 0, the same epochs of the single-GPU
    TextGCNTrainer on the renumbered graph with the same seed and initial weights.  Returns on rank 0
    {max_rel_err_loss, max_rel_err_W2, max_rel_err_W1, ...}.  Collective: every rank must call it."""
    import torch.distributed as dist
    from .models import GCN
    from .trainer import TextGCNTrainer
    n = int(g.x.shape[0])
    tr = make_dist_trainer(g, shape, rank, world, dev, partition, seed=seed, **trainer_kw)
    losses = []
    for _ in range(epochs):
        tr.epoch()
        losses.append(tr.train_loss())                   # all-reduce of the per-rank partial sums
    val = tr.epoch_stats()
    params = tr.gathered_parameters()
    graphed = tr._graph is not None
    out = None
    if rank == 0:
        gen = torch.Generator().manual_seed(seed)        # the draw order of DistTextGCNTrainer.__init__
        a1, a2 = (6.0 / (n + shape.hidden)) ** 0.5, (6.0 / (shape.hidden + shape.n_classes)) ** 0.5
        W1 = (torch.rand(n, shape.hidden, generator=gen) * 2 - 1) * a1
        W2 = (torch.rand(shape.hidden, shape.n_classes, generator=gen) * 2 - 1) * a2
        gp = renumbered_data(g, tr.part).to(dev)
        gcn = GCN(tr.part.n_pad, shape.n_classes, n_hidden_gcn=shape.hidden, dropout=shape.dropout).to(dev)
        with torch.no_grad():
            gcn.layers[0].weight.copy_(tr.part.to_new(W1).to(dev))
            gcn.layers[1].weight.copy_(W2.to(dev))
            gcn.layers[0].bias.zero_(); gcn.layers[1].bias.zero_()
        one = TextGCNTrainer(gcn, gp, lr=shape.lr, amsgrad=shape.amsgrad, seed=seed, use_cuda_graph=False, assume_symmetric=True)
        ref_losses = []
        for _ in range(epochs):
            r = one.epoch()
            ref_losses.append(r["loss"])

        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        out = {"epochs": epochs, "cuda_graph": graphed, "fused_stores": tr.fused_stores, "exchange": tr.exchange,
               "dropout": shape.dropout, "partition": type(tr.part).__name__,
               "max_rel_err_loss": max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(losses, ref_losses)),
               "max_rel_err_W2": rel(params["layers.1.weight"], gcn.layers[1].weight.data),
               "max_rel_err_W1": rel(params["layers.0.weight"], tr.part.to_old(gcn.layers[0].weight.data)),
               "val_loss": [val["val_loss"], r["val_loss"]],
               "reference": "single-GPU TextGCNTrainer (eager) on the renumbered graph, same seed / weights / Philox indices"}
        del one, gcn, gp
    tr._graph = None
    del tr
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------
# bench entry for N > 1 (called by bench.py under torchrun)
# --------------------------------------------------------------------------------------
def _timed_dist_workload(levels, args, rank, local_rank, world, dev, K, W, sample_clocks: bool):
    """Warm-up, K device-timed steps (max over ranks), then the same steps end to end with host buffers.
    levels = [(graph, shape), ...]: one row-partitioned trainer each; a step = one epoch of every level in turn."""
    import torch.distributed as dist
    from . import _native
    lib = _native.load()
    trs = [make_dist_trainer(g, shape, rank, world, dev, getattr(args, "partition", "auto"), seed=args.seed,
                             use_cuda_graph=not getattr(args, "no_cuda_graph", False),
                             exchange=getattr(args, "exchange", "peer"), fused_stores=not getattr(args, "no_fused_stores", False),
                             fuse_adam=not getattr(args, "no_fuse_adam", False), keep_w1_grad=False) for g, shape in levels]
    tr = trs[-1]

    def epoch():
        for t in trs:
            t.epoch()
    for _ in range(W):
        epoch()
    torch.cuda.synchronize()
    dist.barrier()
    # stretch the timed region to >= 1 s in whole multiples of K (same count on every rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        epoch()
    ev1.record()
    torch.cuda.synchronize()
    est = torch.tensor([ev0.elapsed_time(ev1) / 3], device=dev, dtype=torch.float64)
    dist.all_reduce(est, op=dist.ReduceOp.MAX)
    rounds = max(1, int(-(-1000.0 // max(K * float(est.item()), 1e-6)))) if sample_clocks else 1
    sampler = None
    if rank == 0 and sample_clocks:
        from bench import ClockSampler
        sampler = ClockSampler(local_rank)
        sampler.start()
    l0 = lib.tgcn_launch_count()
    torch.cuda.synchronize()
    dist.barrier()
    ev0.record()
    for _ in range(rounds * K):
        epoch()
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = int(lib.tgcn_launch_count() - l0)
    if all(t._graph is not None for t in trs):
        launches = sum(t.launches_per_epoch for t in trs) * rounds * K
    ms_per_step = float(ms.item()) / (rounds * K)

    # e2e: labels/masks from pinned host memory each epoch, losses + local argmax read back
    nl = tr.part.n_loc
    pins = [(t, t.y.cpu().pin_memory(), t.train_mask.cpu().pin_memory(), t.val_mask.cpu().pin_memory(),
             torch.empty(t.part.n_loc, dtype=torch.int32).pin_memory()) for t in trs]

    def epoch_e2e():
        out = None
        for t, y_pin, tm_pin, vm_pin, pred_pin in pins:
            t.y.copy_(y_pin, non_blocking=True)
            t.train_mask.copy_(tm_pin, non_blocking=True)
            t.val_mask.copy_(vm_pin, non_blocking=True)
            t.epoch()
            pred_pin.copy_(t.pred, non_blocking=True)
            out = t.epoch_stats()                    # all-reduce + D2H of the global val loss / accuracy
        return out
    for _ in range(2):
        epoch_e2e()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(rounds * K):
        last = epoch_e2e()
    torch.cuda.synchronize()
    dist.barrier()
    e2e = torch.tensor([(time.perf_counter() - t0) * 1e3 / (rounds * K)], device=dev, dtype=torch.float64)
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler is not None else None
    nnz_loc = torch.tensor([tr.shard.nnz + (tr.qshard.nnz if hasattr(tr, "qshard") else 0)], device=dev, dtype=torch.float64)
    nnz_all = [torch.zeros_like(nnz_loc) for _ in range(world)]
    dist.all_gather(nnz_all, nnz_loc)
    rec = dict(ms_per_step=ms_per_step, e2e_ms=float(e2e.item()), launches=launches, timed_steps=rounds * K, clocks=clocks,
               nl=nl, nnz_per_rank=[int(t.item()) for t in nnz_all], bytes_per_train_step=sum(t.bytes_per_train_step() for t in trs),
               last=last, cuda_graph=all(t._graph is not None for t in trs), graph_error=tr.graph_error, exchange=tr.exchange,
               exchange_error=tr.exchange_error, fused_stores=tr.fused_stores, share_h1=tr.share_h1,
               restrict_rows=tr.restrict_rows, partition=[type(t.part).__name__ for t in trs],
               multicast=bool(tr.px is not None and any(tr.px.multicast.values())),
               launches_per_epoch=sum(t.launches_per_epoch for t in trs))
    for t in trs:
        t._graph = None
    del tr, trs, pins
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.empty_cache()
    return rec


def run_distributed_bench(args, rank: int, local_rank: int, world: int, dev: torch.device):
    import torch.distributed as dist
    from .synthetic import SHAPES, make_graph
    from bench import METRIC, UNIT, resolve_workload, workload_config
    specs = resolve_workload(args.workload)
    label, shape, hier0 = specs[0]
    K, W = args.steps, max(args.warmup, 5)
    levels = [(make_graph(sh, seed=args.seed, hierarchy_classes=hier), sh) for _, sh, hier in specs]   # same graphs on every rank
    g = levels[0][0]
    main = _timed_dist_workload(levels, args, rank, local_rank, world, dev, K, W, sample_clocks=True)
    extra = {"nnz_per_rank": main["nnz_per_rank"], "rows_per_rank": main["nl"],
             "collective_bytes_received_per_rank_per_train_step": main["bytes_per_train_step"],
             "last_epoch": main["last"], "cuda_graph": main["cuda_graph"], "cuda_graph_error": main["graph_error"],
             "exchange": main["exchange"], "exchange_error": main["exchange_error"], "fused_stores": main["fused_stores"],
             "multicast": main["multicast"], "share_h1": main["share_h1"], "restrict_rows": main["restrict_rows"],
             "timed_steps": main["timed_steps"], "partition": main["partition"],
             "kernels_per_epoch": main["launches_per_epoch"]}
    if not getattr(args, "no_extras", False):
        # (1) the shipped N-rank configuration (CUDA graph + multimem stores + dropout) against the single-GPU trainer
        try:
            if hier0 is not None:
                raise NotImplementedError("parity helper covers x = I")
            extra["parity"] = parity_against_single_gpu(
                g, shape, rank, world, dev, args.seed, epochs=5, partition=choose_partition(g, getattr(args, "partition", "auto")),
                use_cuda_graph=not getattr(args, "no_cuda_graph", False),
                exchange=getattr(args, "exchange", "peer"), fused_stores=not getattr(args, "no_fused_stores", False),
                fuse_adam=not getattr(args, "no_fuse_adam", False), keep_w1_grad=False)
        except Exception as e:
            extra["parity"] = {"error": repr(e)}
        # (2) the >= 1 M-node configuration north_star names for scaling, at this N
        if args.workload != "scale":
            try:
                del g, levels
                sc = SHAPES["scale"]
                gs = make_graph(sc, seed=args.seed)
                r = _timed_dist_workload([(gs, sc)], args, rank, local_rank, world, dev, 10, 5, sample_clocks=False)
                extra["scale_config"] = {"epochs_per_s": 1e3 / r["ms_per_step"], "ms_per_step": r["ms_per_step"], "steps": 10,
                                         "n_nodes": int(gs.x.shape[0]), "n_edges": int(gs.edge_index.shape[1]),
                                         "hidden": sc.hidden, "nnz_per_rank": r["nnz_per_rank"],
                                         "collective_bytes_received_per_rank_per_train_step": r["bytes_per_train_step"],
                                         "e2e_epochs_per_s": 1e3 / r["e2e_ms"], "cuda_graph": r["cuda_graph"],
                                         "partition": r["partition"][0], "exchange": r["exchange"]}
                g = gs
            except Exception as e:
                extra["scale_config"] = {"error": repr(e)}
    if rank == 0:
        sh_last, hier_last = specs[-1][1], specs[-1][2]
        g_cfg = make_graph(sh_last, seed=args.seed, hierarchy_classes=hier_last)
        cfg = workload_config(args.workload, specs, g_cfg)
        cfg["parallelism"] = (f"word-block partition x{world} (words and documents dealt in snake order; exchange = all-gather of the "
                              "word block + all-to-all of the partial word rows: " +
                              ("multimem stores in the producers / peer stores in the Q SpMM epilogue + device barrier)"
                               if main["exchange"].startswith("peer") else "NCCL)") if main["partition"][0] == "BipartitePartition" else
                              f"1D row partition x{world} (snake order by nnz); exchange between layers: " +
                              (("multimem stores fused into the producer kernels + device barrier" if main["fused_stores"] else
                                "peer-store push kernel into symmetric buffers + device barrier") if main["exchange"] == "peer"
                               else "NCCL all_gather_into_tensor"))
        nl = main["nl"]
        line = {
            "metric": METRIC, "value": 1e3 / main["ms_per_step"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": main["clocks"],
            "e2e": {"value": 1e3 / main["e2e_ms"], "unit": UNIT, "h2d_bytes_per_step": int(nl * 10) * world,
                    "d2h_bytes_per_step": int(nl * 4 + 16) * world, "ms_per_step": main["e2e_ms"]},
            "gpu_launches": main["launches"], "timed_steps": main["timed_steps"],
            "extra": extra,
        }
        print(json.dumps(line), flush=True)
    shutdown(None)
