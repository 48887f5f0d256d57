"""1-D node-partitioned message passing: halo plan + exchange (one process per GPU).

The reference is single-process (no distributed code at all); this module is the
multi-GPU extension named by the north star.  Rank r owns a contiguous range of target nodes
with ALL their in-edges, their feature rows and their output rows.  Before each aggregation
the feature rows of remote sources ("halo") are fetched:

    send_buf = kgb_gather_rows(x_local, send_idx)              (pack, K7)
    all_to_all_single(x_ext[n_local:], send_buf)               (NCCL over NVLink / NVSwitch)
    out      = kgb_gather_reduce(x_ext, local CSR)             (columns remapped to [local | halo])

and in the backward the halo gradients travel the reverse way and are summed into the owners'
rows by a deterministic segmented sum (kgb_gather_reduce over the CSR of send_idx), not by
atomics.  Degrees (mean / GCN normalisation) need no communication: every in-edge of an owned
target is local.  The conv layers accept a ``PartitionedGraph`` in place of ``edge_index``.

Works with any torch.distributed backend (tests use gloo on the CPU for the plan logic; the
exchange of CUDA tensors needs NCCL).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch.autograd.function import once_differentiable


def partition_bounds(n_global: int, world: int) -> list:
    """Contiguous, balanced node ranges: rank r owns [b[r], b[r+1])."""
    base, extra = divmod(int(n_global), int(world))
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < extra else 0))
    return b


class HaloPlan:
    """Pure index logic of the exchange (device-agnostic; unit-tested with gloo on the CPU).

    Built from the edges whose TARGET this rank owns (global ids).  Attributes:
      n_local, n_halo       owned rows / distinct remote source rows
      col_local  [E_loc]    source ids remapped to [0, n_local) U [n_local, n_local+n_halo)
      dst_local  [E_loc]    target ids minus the range start
      halo_global [n_halo]  sorted global ids of the halo rows (grouped by owner because ranges are contiguous)
      recv_counts [world]   rows this rank receives from each peer   (sum = n_halo)
      send_counts [world]   rows this rank sends to each peer
      send_idx [n_send]     local row ids to pack, grouped by destination peer
    """

    def __init__(self, src_global: torch.Tensor, dst_global: torch.Tensor, n_global: int, rank: int, world: int,
                 group=None):
        self.rank, self.world, self.group = rank, world, group
        self.bounds = partition_bounds(n_global, world)
        lo, hi = self.bounds[rank], self.bounds[rank + 1]
        self.lo, self.hi = lo, hi
        self.n_local = hi - lo
        dev = src_global.device
        src = src_global.long()
        dst = dst_global.long()
        if dst.numel() and (int(dst.min()) < lo or int(dst.max()) >= hi):
            raise ValueError(f"rank {rank} was given edges whose target is outside its range [{lo}, {hi})")
        remote = (src < lo) | (src >= hi)
        halo_global = torch.unique(src[remote])  # sorted
        self.halo_global = halo_global
        self.n_halo = int(halo_global.numel())
        col = src - lo
        if self.n_halo:
            pos = torch.searchsorted(halo_global, src[remote])
            col[remote] = self.n_local + pos
        self.col_local = col.to(torch.int32)
        self.dst_local = (dst - lo).to(torch.int32)
        bnd = torch.tensor(self.bounds, device=dev, dtype=torch.long)
        owner = torch.bucketize(halo_global, bnd[1:], right=True) if self.n_halo else halo_global
        self.recv_counts = torch.bincount(owner, minlength=world).tolist() if self.n_halo else [0] * world
        # tell every owner which of its rows I need
        recv_c = torch.tensor(self.recv_counts, device=dev, dtype=torch.long)
        send_c = torch.empty_like(recv_c)
        if world > 1:
            dist.all_to_all_single(send_c, recv_c, group=group)
        else:
            send_c.copy_(recv_c)
        self.send_counts = send_c.tolist()
        want = torch.empty(int(sum(self.send_counts)), device=dev, dtype=torch.long)
        if world > 1:
            dist.all_to_all_single(want, halo_global.contiguous(), output_split_sizes=self.send_counts,
                                   input_split_sizes=self.recv_counts, group=group)
        self.send_idx = (want - lo).to(torch.int32)
        if want.numel() and (int(self.send_idx.min()) < 0 or int(self.send_idx.max()) >= self.n_local):
            raise RuntimeError("halo plan: a peer requested a row this rank does not own")
        self.n_send = int(self.send_idx.numel())

    def edge_index_local(self) -> torch.Tensor:
        return torch.stack([self.col_local, self.dst_local]).contiguous()

    def split_edges(self):
        """The rank's edges as two COO lists that keep the original edge order: sources this rank owns (source ids
        in [0, n_local)) and halo sources (ids in [0, n_halo): rows of the exchanged buffer)."""
        ei = self.edge_index_local()
        is_halo = ei[0] >= self.n_local
        ei_l = ei[:, ~is_halo].contiguous()
        ei_h = ei[:, is_halo].clone()
        ei_h[0] -= self.n_local
        return ei_l, ei_h.contiguous()


class PartitionedGraph:
    """Halo plan + device structures of one rank.  Pass it to SAGEConv / GCNConv instead of
    ``edge_index``; ``x`` is then this rank's [n_local, F] slice of the node features."""

    def __init__(self, src_global, dst_global, n_global: int, rank: int | None = None, world: int | None = None,
                 group=None, n_loops_local: bool = False):
        from .graph import GraphStructure, build_csr
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.plan = HaloPlan(src_global, dst_global, n_global, rank, world, group)
        p = self.plan
        self.n_local, self.n_halo, self.n_ext = p.n_local, p.n_halo, p.n_local + p.n_halo
        self.group, self.world, self.rank = group, world, rank
        self.n_global = int(n_global)
        self.device = src_global.device
        self._n_loops = self.n_local if n_loops_local else 0
        self._graph = None
        self._split = None
        # deterministic backward of the pack: CSR of send_idx (which packed rows came from local row r)
        if p.n_send:
            s_ei = torch.stack([torch.zeros_like(p.send_idx), p.send_idx]).contiguous()
            self.send_csr = build_csr(s_ei, self.n_local, 1, 0, by_source=False)
        else:
            self.send_csr = None
        self._dis_ext = None

    @property
    def graph(self):
        """All in-edges of the owned rows over the [local | halo] source space (built on first use)."""
        if self._graph is None:
            from .graph import GraphStructure
            # self-loops i->i are local edges on the owned rows (ids [0, n_local) in both spaces)
            self._graph = GraphStructure(self.plan.edge_index_local(), self.n_local, self.n_ext, self._n_loops)
        return self._graph

    @property
    def split(self):
        """(graph_local, graph_halo, inv_deg): the same edges as two structures - sources this rank owns
        ([0, n_local)) and halo sources ([0, n_halo), rows of the exchanged buffer) - so that a linear aggregator
        can reduce the local part while the halo rows are still in flight.  inv_deg = 1 / max(total in-degree, 1e-8)."""
        if self._split is None:
            from .graph import GraphStructure
            ei_l, ei_h = self.plan.split_edges()
            g_l = GraphStructure(ei_l, self.n_local, self.n_local, 0)
            g_h = GraphStructure(ei_h, self.n_local, max(self.n_halo, 1), 0)
            deg = (g_l.csr.deg + g_h.csr.deg).to(torch.float32)
            self._split = (g_l, g_h, 1.0 / torch.clamp(deg, min=1e-8))
        return self._split

    # ---- forward/backward exchange ------------------------------------------------------------
    # The exchange (pack kernel, all-to-all, and in the backward the reverse all-to-all + segmented sum) runs on a
    # dedicated CUDA stream.  autograd replays each node's backward on the stream its forward ran on and inserts the
    # cross-stream dependencies itself, so the halo traffic of both directions overlaps the work that does not need
    # it (root-weight GEMM forward; weight-gradient and root GEMMs backward).
    def _comm_stream(self) -> torch.cuda.Stream:
        cs = getattr(self, "_cs", None)
        if cs is None:
            cs = self._cs = torch.cuda.Stream(device=self.device)
        return cs

    def exchange_start(self, x_local: torch.Tensor) -> torch.Tensor:
        """[n_local, F] -> [n_local + n_halo, F] (autograd-aware), left in flight on the communication stream.
        Call ``exchange_finish`` before the first use of the result on the current stream."""
        self.exchange_finish()
        cur = torch.cuda.current_stream(self.device)
        cs = self._comm_stream()
        cs.wait_stream(cur)
        with torch.cuda.stream(cs):
            x_ext = _HaloExchange.apply(x_local, self)
        x_local.record_stream(cs)
        self._pending = x_ext
        return x_ext

    def exchange_finish(self) -> None:
        x_ext, self._pending = getattr(self, "_pending", None), None
        if x_ext is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_stream(self._comm_stream())
            x_ext.record_stream(cur)

    def exchange(self, x_local: torch.Tensor) -> torch.Tensor:
        x_ext = self.exchange_start(x_local)
        self.exchange_finish()
        return x_ext

    def halo_start(self, x_local: torch.Tensor) -> torch.Tensor:
        """[n_local, F] -> the [n_halo, F] halo rows only (autograd-aware), in flight on the communication stream;
        pair with the ``split`` structures and call ``halo_finish`` before the first use on the current stream."""
        self.exchange_finish()
        cur = torch.cuda.current_stream(self.device)
        cs = self._comm_stream()
        cs.wait_stream(cur)
        with torch.cuda.stream(cs):
            halo = _HaloRows.apply(x_local, self)
        x_local.record_stream(cs)
        self._pending = halo
        return halo

    halo_finish = exchange_finish

    # ---- raw (no autograd) halves of the exchange, for nodes that schedule the overlap themselves ----
    def halo_rows_raw(self, x_local: torch.Tensor) -> torch.Tensor:
        """pack + all-to-all on the CURRENT stream: [n_local, F] -> [n_halo, F]."""
        from . import ops
        p = self.plan
        F = int(x_local.shape[1])
        halo = torch.empty((max(self.n_halo, 1), F), dtype=x_local.dtype, device=x_local.device)
        send = ops.gather_rows(x_local, p.send_idx) if p.n_send else x_local.new_empty((0, F))
        dist.all_to_all_single(halo[:self.n_halo], send, output_split_sizes=p.recv_counts,
                               input_split_sizes=p.send_counts, group=self.group)
        return halo

    def halo_grad_raw(self, g_halo: torch.Tensor) -> torch.Tensor:
        """reverse all-to-all on the CURRENT stream: [n_halo, F] gradient rows -> [n_send, F] rows at their owners
        (to be summed per owner row with ``send_csr``)."""
        p = self.plan
        F = int(g_halo.shape[1])
        back = torch.empty((max(p.n_send, 1), F), dtype=g_halo.dtype, device=g_halo.device)
        dist.all_to_all_single(back[:p.n_send], g_halo[:self.n_halo], output_split_sizes=p.send_counts,
                               input_split_sizes=p.recv_counts, group=self.group)
        return back

    def exchange_vector(self, v_local: torch.Tensor) -> torch.Tensor:
        """Per-node scalar (e.g. GCN dis) -> [n_ext]; no autograd."""
        return _exchange_fwd(v_local.reshape(-1, 1).contiguous(), self).reshape(-1)

    def gcn_dis_ext(self) -> torch.Tensor:
        """deg^-1/2 for local rows followed by the halo rows' values (fetched once per graph)."""
        if self._dis_ext is None:
            from . import _lib
            from .graph import _stream
            lib = _lib.load()
            dis = torch.empty(self.n_local, dtype=torch.float32, device=self.graph.device)
            _lib.check(lib.kgb_gcn_norm(self.graph.device.index, self.graph.csr.deg.data_ptr(), self.n_local, None,
                                        0, 0, dis.data_ptr(), None, _stream(self.graph.device)), "kgb_gcn_norm")
            self._dis_ext = (dis, self.exchange_vector(dis))
        return self._dis_ext


def _exchange_fwd(x_local: torch.Tensor, pg: PartitionedGraph) -> torch.Tensor:
    from . import ops
    p = pg.plan
    F = int(x_local.shape[1])
    x_ext = torch.empty((pg.n_ext, F), dtype=x_local.dtype, device=x_local.device)
    if pg.world > 1:
        send = ops.gather_rows(x_local, p.send_idx) if p.n_send else x_local.new_empty((0, F))
        dist.all_to_all_single(x_ext[pg.n_local:], send, output_split_sizes=p.recv_counts,
                               input_split_sizes=p.send_counts, group=pg.group)
    x_ext[:pg.n_local].copy_(x_local)
    return x_ext


class _HaloExchange(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, pg: PartitionedGraph):
        ctx.pg = pg
        return _exchange_fwd(x_local.contiguous(), pg)

    @staticmethod
    @once_differentiable
    def backward(ctx, g_ext):
        from . import _lib, ops
        pg = ctx.pg
        p = pg.plan
        g_ext = g_ext.contiguous()
        g_ext.record_stream(torch.cuda.current_stream(g_ext.device))  # produced on the compute stream
        g_local = g_ext[:pg.n_local].clone()
        if pg.world > 1:
            F = int(g_ext.shape[1])
            back = torch.empty((p.n_send, F), dtype=g_ext.dtype, device=g_ext.device)
            dist.all_to_all_single(back, g_ext[pg.n_local:].contiguous(), output_split_sizes=p.send_counts,
                                   input_split_sizes=p.recv_counts, group=pg.group)
            if p.n_send:
                add, _ = ops.gather_reduce_raw(back, pg.send_csr, _lib.OP_SUM, col=pg.send_csr.perm)
                g_local += add
        return g_local, None


class _HaloRows(torch.autograd.Function):
    """x_local [n_local, F] -> halo [n_halo, F]: pack + all-to-all; backward = reverse all-to-all + the deterministic
    segmented sum of the returned gradient rows into their owners (zero for rows nobody asked for)."""

    @staticmethod
    def forward(ctx, x_local, pg: PartitionedGraph):
        from . import ops
        ctx.pg = pg
        p = pg.plan
        x_local = x_local.contiguous()
        F = int(x_local.shape[1])
        halo = torch.empty((max(pg.n_halo, 1), F), dtype=x_local.dtype, device=x_local.device)
        if pg.n_halo == 0:
            halo.zero_()
        if pg.world > 1:
            send = ops.gather_rows(x_local, p.send_idx) if p.n_send else x_local.new_empty((0, F))
            dist.all_to_all_single(halo[:pg.n_halo], send, output_split_sizes=p.recv_counts,
                                   input_split_sizes=p.send_counts, group=pg.group)
        return halo

    @staticmethod
    @once_differentiable
    def backward(ctx, g_halo):
        from . import _lib, ops
        pg = ctx.pg
        p = pg.plan
        g_halo = g_halo.contiguous()
        g_halo.record_stream(torch.cuda.current_stream(g_halo.device))  # produced on the compute stream
        F = int(g_halo.shape[1])
        if pg.world > 1 and p.n_send:
            back = torch.empty((p.n_send, F), dtype=g_halo.dtype, device=g_halo.device)
            dist.all_to_all_single(back, g_halo[:pg.n_halo], output_split_sizes=p.send_counts,
                                   input_split_sizes=p.recv_counts, group=pg.group)
            g_local, _ = ops.gather_reduce_raw(back, pg.send_csr, _lib.OP_SUM, col=pg.send_csr.perm)
        else:
            if pg.world > 1:
                dist.all_to_all_single(g_halo.new_empty((0, F)), g_halo[:pg.n_halo], output_split_sizes=p.send_counts,
                                       input_split_sizes=p.recv_counts, group=pg.group)
            g_local = torch.zeros((pg.n_local, F), dtype=g_halo.dtype, device=g_halo.device)
        return g_local, None
