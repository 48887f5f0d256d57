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

Transport.  With CUDA tensors and more than one rank the halo rows do not go through NCCL send/recv: every rank
owns a WINDOW of device memory that its peers map with CUDA IPC (``PeerWindow``), and the pack kernel
(kgb_halo_push) stores each requested row straight into the receiver's window over NVLink.  One tiny all-reduce on
the communication stream orders the ranks (all rows landed); the window is double-buffered so that no "window free"
barrier is needed before the push.  ``KGB200_HALO=nccl`` (or a failed IPC
mapping) selects ``all_to_all_single`` instead; the CPU tests use gloo for the plan logic.
"""
from __future__ import annotations

import ctypes
import os
import warnings

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


def check_bounds(bounds, n_global: int, world: int) -> list:
    """Validate user-supplied contiguous ranges: world + 1 non-decreasing cuts from 0 to n_global."""
    b = [int(v) for v in bounds]
    if len(b) != world + 1 or b[0] != 0 or b[-1] != int(n_global) or any(b[i] > b[i + 1] for i in range(world)):
        raise ValueError(f"bounds must be {world + 1} non-decreasing cuts from 0 to {n_global}, got {b}")
    return b


def cost_balanced_bounds(dst_global: torch.Tensor, n_global: int, world: int, node_weight: float = 28.0) -> list:
    """Contiguous node ranges of ~equal cost = in-edges + node_weight * nodes.  Power-law graphs with skewed ids
    (RMAT) give one rank most of the edges under equal node counts and most of the dense-transform rows under equal
    edge counts; ``node_weight`` is the cost of one node's dense transforms in units of one gathered edge
    (28 for the 256-wide SAGE step on a B200).  Every rank must pass the SAME ``dst_global`` (all edges)."""
    deg = torch.bincount(dst_global.long(), minlength=n_global).to(torch.float64) + float(node_weight)
    cum = torch.cumsum(deg, 0)
    total = float(cum[-1]) if n_global else 0.0
    targets = torch.tensor([total * r / world for r in range(1, world)], device=dst_global.device,
                           dtype=torch.float64)
    cuts = (torch.searchsorted(cum, targets) + 1).tolist() if world > 1 and n_global else [0] * (world - 1)
    b = [0] + [min(int(c), n_global) for c in cuts] + [n_global]
    for i in range(1, len(b)):
        b[i] = max(b[i], b[i - 1])
    return b


def scramble_ids(edge_index: torch.Tensor, n_global: int, seed: int = 0x5eed, group=None):
    """Hash partitioning for the contiguous 1-D split: relabel the nodes with a seeded random bijection (the same on
    every rank: generated on rank 0, broadcast) and return ``(edge_index_new, perm)`` with ``perm[old id] = new id``.

    Generators such as RMAT correlate the degree with the id (the hubs are the low ids), so contiguous ranges are
    balanced either in nodes or in edges, never both - and the layers synchronise at every exchange, so the forward
    gathers wait for the rank that owns the hubs and the dense transforms for the rank that owns most rows.  After
    scrambling every range holds ~n/world nodes AND ~E/world edges and needs about the same number of halo rows
    (Graph500 scrambles its RMAT ids for the same reason).  Features / labels must be indexed with the new ids:
    ``x_new[perm] = x_old``."""
    dev = edge_index.device
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if rank == 0:
        gen = torch.Generator(device=dev).manual_seed(int(seed))
        perm = torch.randperm(int(n_global), device=dev, generator=gen).to(torch.int32)
    else:
        perm = torch.empty(int(n_global), dtype=torch.int32, device=dev)
    if world > 1:
        dist.broadcast(perm, src=0, group=group)
    out = torch.empty_like(edge_index)
    for r in range(2):   # one row at a time: the int64 index temporary is 8 B per edge
        out[r] = perm[edge_index[r].long()]
    return out, perm


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
                 group=None, bounds=None):
        self.rank, self.world, self.group = rank, world, group
        self.bounds = partition_bounds(n_global, world) if bounds is None else check_bounds(bounds, n_global, world)
        if bounds is not None and world > 1:   # every rank must cut the node range at the same places
            theirs = [None] * world
            dist.all_gather_object(theirs, self.bounds, group=group)
            if any(t != self.bounds for t in theirs):
                raise ValueError(f"rank {rank}: partition bounds differ between ranks: {theirs}")
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
        # recv_matrix[p][q] = rows rank p receives from rank q (== rows q sends to p): every rank knows where its
        # block starts inside each peer's halo buffer (forward push) and inside each owner's send list (backward push)
        if world > 1:
            rm = [None] * world
            dist.all_gather_object(rm, [int(c) for c in self.recv_counts], group=group)
            self.recv_matrix = rm
        else:
            self.recv_matrix = [[0]]
        self.any_halo = any(sum(row) > 0 for row in self.recv_matrix)   # rank-uniform: some rank needs remote rows

    def push_tables(self):
        """Index arithmetic of the peer-memory exchange (pure host logic, tested on the CPU).
        forward  (rows of mine -> peers' halo buffers):  slots grouped by destination peer p,
                 slot_begin = prefix(send_counts), my block starts at row sum_{q < me} recv_matrix[p][q] of p's halo;
        backward (gradients of my halo rows -> their owners' send lists): slots grouped by owner q,
                 slot_begin = prefix(recv_counts), my block starts at row sum_{p < me} recv_matrix[p][q] of q's list."""
        me, W, rm = self.rank, self.world, self.recv_matrix
        fwd_begin, bwd_begin = [0], [0]
        for p in range(W):
            fwd_begin.append(fwd_begin[-1] + int(self.send_counts[p]))
            bwd_begin.append(bwd_begin[-1] + int(self.recv_counts[p]))
        fwd_row0 = [sum(rm[p][:me]) for p in range(W)]
        bwd_row0 = [sum(rm[p][q] for p in range(me)) for q in range(W)]
        n_halo_all = [sum(rm[p]) for p in range(W)]
        n_send_all = [sum(rm[p][q] for p in range(W)) for q in range(W)]
        return {"fwd_begin": fwd_begin, "fwd_row0": fwd_row0, "bwd_begin": bwd_begin, "bwd_row0": bwd_row0,
                "n_halo_all": n_halo_all, "n_send_all": n_send_all}

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


class _DeviceBytes:
    """``__cuda_array_interface__`` view of raw device memory (the window is not torch-allocated)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


class PeerWindow:
    """One symmetric window per rank: two copies ("phases") of ``[halo region | back region]``, each region
    ``rows x f_cap`` floats, allocated with cudaMalloc by libkgb200 and mapped into every peer with CUDA IPC.
    ``ptr[r]`` is rank r's window as seen from THIS device (NVLink peer memory for r != rank).

    Consecutive exchanges of the same direction alternate between the two phases, which is what lets an exchange
    start pushing WITHOUT first waiting for "every rank has consumed its previous rows": a rank writes phase p of
    direction d at exchange k only after it has left the landed-barrier of exchange k-1, which every peer entered
    after its communication stream waited for its compute stream at the start of exchange k-1 - i.e. after the peer
    had finished reading the rows of the last exchange that used (d, p), two same-direction exchanges ago.  One
    barrier per exchange (all rows landed) instead of two; an early rank's rows travel while the late ranks are
    still computing."""

    def __init__(self, plan: HaloPlan, f_cap: int, device, group=None):
        """Collective.  Every stage that can fail locally is followed by a collective that all ranks reach, so a
        failure on one rank turns into ``self.ok == False`` on all of them instead of a hang."""
        from . import _lib
        lib = _lib.load()
        self.lib, self.device, self.group, self.f_cap = lib, device, group, int(f_cap)
        self.rank, self.world = plan.rank, plan.world
        t = plan.push_tables()
        self.tables = t
        row_bytes = 4 * self.f_cap
        self.back_off = [((t["n_halo_all"][r] * row_bytes + 255) // 256) * 256 for r in range(self.world)]
        # bytes of one phase of rank r's window (every rank knows every rank's layout)
        self.phase_stride = [((self.back_off[r] + max(t["n_send_all"][r], 1) * row_bytes + 255) // 256) * 256
                             for r in range(self.world)]
        self._phase = {True: 0, False: 0}     # next phase of the forward / backward direction
        nbytes = 2 * self.phase_stride[self.rank] + 256
        self.nbytes = nbytes
        self.base, self.ptr, self._opened, self.error = None, [], [], None
        raw = None
        try:   # stage 1 (local): allocate + export
            base = ctypes.c_void_p()
            _lib.check(lib.kgb_window_alloc(device.index, nbytes, ctypes.byref(base)), "kgb_window_alloc")
            self.base = base.value
            handle = ctypes.create_string_buffer(64)
            _lib.check(lib.kgb_ipc_export(device.index, self.base, handle), "kgb_ipc_export")
            raw = bytes(handle.raw)
        except Exception as e:  # noqa: BLE001
            self.error = e
        handles = [None] * self.world
        dist.all_gather_object(handles, raw, group=group)
        if self.error is None and all(h is not None for h in handles):
            try:   # stage 2 (local): map the peers
                for r, h in enumerate(handles):
                    if r == self.rank:
                        self.ptr.append(self.base)
                        continue
                    p = ctypes.c_void_p()
                    _lib.check(lib.kgb_ipc_open(device.index, h, ctypes.byref(p)), "kgb_ipc_open")
                    self.ptr.append(p.value)
                    self._opened.append(p.value)
                self.local = torch.as_tensor(_DeviceBytes(self.base, nbytes), device=device)
            except Exception as e:  # noqa: BLE001
                self.error = e
        elif self.error is None:
            self.error = RuntimeError("a peer could not allocate / export its window")
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)
        ok = torch.tensor([1 if self.error is None else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        self.ok = bool(int(ok))

    def barrier(self) -> None:
        """Orders the ranks on the CURRENT stream (a one-element all-reduce; stream-ordered like every collective)."""
        dist.all_reduce(self._flag, group=self.group)

    def next_phase(self, forward: bool) -> int:
        """Phase of the exchange that starts now in this direction (same sequence on every rank); toggles it."""
        ph = self._phase[forward]
        self._phase[forward] = ph ^ 1
        return ph

    def halo_view(self, n_rows: int, F: int, phase: int = 0) -> torch.Tensor:
        o = phase * self.phase_stride[self.rank]
        return self.local[o:o + max(n_rows, 1) * F * 4].view(torch.float32).view(max(n_rows, 1), F)

    def back_view(self, n_rows: int, F: int, phase: int = 0) -> torch.Tensor:
        o = phase * self.phase_stride[self.rank] + self.back_off[self.rank]
        return self.local[o:o + max(n_rows, 1) * F * 4].view(torch.float32).view(max(n_rows, 1), F)

    def push_args(self, F: int, forward: bool, phase: int = 0):
        """Destination table (``kgb_halo_push_args`` without a source): forward = my rows -> peers' halo regions,
        else my halo-row gradients -> their owners' back regions, of the given phase.  Rows are dense (ld = F)."""
        from . import _lib
        t = self.tables
        a = _lib.HaloPushArgs()
        a.F, a.n_peers = F, self.world
        begin = t["fwd_begin"] if forward else t["bwd_begin"]
        row0 = t["fwd_row0"] if forward else t["bwd_row0"]
        for p in range(self.world + 1):
            a.slot_begin[p] = begin[p]
        for p in range(self.world):
            a.dst[p] = self.ptr[p] + phase * self.phase_stride[p] + (0 if forward else self.back_off[p])
            a.dst_row0[p] = row0[p]
        a.ldd = F
        a.slot_rot = begin[(self.rank + 1) % self.world]   # stagger the receivers across the ranks
        return a

    def push(self, src: torch.Tensor, idx, F: int, forward: bool, phase: int = 0) -> None:
        """kgb_halo_push on the CURRENT stream (pack + store into the peers' windows)."""
        from . import _lib
        from .graph import _stream
        a = self.push_args(F, forward, phase)
        a.src, a.lds, a.idx = src.data_ptr(), src.stride(0), (idx.data_ptr() if idx is not None else None)
        _lib.check(self.lib.kgb_halo_push(self.device.index, ctypes.byref(a), _stream(self.device)), "kgb_halo_push")

    def close(self) -> None:
        from . import _lib
        torch.cuda.synchronize(self.device)
        for p in self._opened:
            self.lib.kgb_ipc_close(self.device.index, p)
        self._opened = []
        if dist.is_initialized():
            try:
                dist.barrier(group=self.group)   # nobody unmaps-after-free: peers closed their mappings first
            except Exception:  # noqa: BLE001
                pass
        if self.base is not None:
            _lib.check(self.lib.kgb_window_free(self.device.index, self.base), "kgb_window_free")
        self.base = None


class PartitionedGraph:
    """Halo plan + device structures of one rank.  Pass it to SAGEConv / GCNConv instead of
    ``edge_index``; ``x`` is then this rank's [n_local, F] slice of the node features."""

    def __init__(self, src_global, dst_global, n_global: int, rank: int | None = None, world: int | None = None,
                 group=None, n_loops_local: bool = False, bounds=None):
        """``bounds``: optional world + 1 cuts of the node range (e.g. ``cost_balanced_bounds``), identical on every
        rank; default: equal node counts (``partition_bounds``)."""
        from .graph import GraphStructure, build_csr
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.plan = HaloPlan(src_global, dst_global, n_global, rank, world, group, bounds=bounds)
        self.any_halo = self.plan.any_halo
        self._window = None
        self._p2p_failed = False
        p = self.plan
        self.n_local, self.n_halo, self.n_ext = p.n_local, p.n_halo, p.n_local + p.n_halo
        self.group, self.world, self.rank = group, world, rank
        self.n_global = int(n_global)
        self.device = src_global.device
        self._n_loops = self.n_local if n_loops_local else 0
        self._graph = None
        self._split = None
        # deterministic backward of the pack: CSR over the DISTINCT rows of send_idx (which returned gradient rows
        # belong to owned row r).  Only rows somebody asked for are touched when the gradients land.
        if p.n_send:
            uniq, inv = torch.unique(p.send_idx.long(), return_inverse=True)
            s_ei = torch.stack([torch.zeros_like(inv), inv]).to(torch.int32).contiguous()
            self.send_csr = build_csr(s_ei, int(uniq.numel()), 1, 0, by_source=False)
            self.send_rows = uniq.to(torch.int32)
        else:
            self.send_csr = None
            self.send_rows = None
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
            g_l = GraphStructure(ei_l, self.n_local, self.n_local, self._n_loops)   # self-loops are local edges
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
    # ---- transport -----------------------------------------------------------------------------------------
    def _p2p_window(self, F: int):
        """The peer-memory window (created collectively on first use, re-created when a wider row arrives), or None
        when the transport is NCCL (KGB200_HALO=nccl, CPU tensors, or the IPC mapping failed)."""
        if (self.world <= 1 or self.device.type != "cuda" or self._p2p_failed
                or os.environ.get("KGB200_HALO", "p2p").lower() == "nccl" or dist.get_backend(self.group) != "nccl"):
            return None
        w = self._window
        if w is not None and F <= w.f_cap:
            return w
        if w is not None:
            w.close()
            self._window = None
        f_cap = max(64, (int(F) + 63) // 64 * 64)
        w = PeerWindow(self.plan, f_cap, self.device, self.group)
        if not w.ok:   # decided collectively: all ranks fall back to the same transport
            warnings.warn(f"keras_geometric_b200.dist: peer-memory halo window unavailable ({w.error}); "
                          "using NCCL all_to_all")
            w.close()
            self._p2p_failed = True
            return None
        self._window = w
        return w

    def halo_rows_raw(self, x_local: torch.Tensor) -> torch.Tensor:
        """pack + exchange on the CURRENT stream: [n_local, F] -> [n_halo, F].  With the peer-memory transport the
        result is a view of this rank's window: valid until the second-next forward exchange on this graph."""
        from . import ops
        p = self.plan
        F = int(x_local.shape[1])
        win = self._p2p_window(F)
        if win is not None:
            ph = win.next_phase(True)                       # no "window free" barrier: see PeerWindow
            with ops._prof(f"halo_push_fwd_F{F}", p.n_send * F * 4, self.device):
                win.push(x_local, p.send_idx if p.n_send else None, F, forward=True, phase=ph)
            with ops._prof("halo_wait_post", 0, self.device):
                win.barrier()                               # every rank's rows have landed
            return win.halo_view(self.n_halo, F, ph)
        halo = torch.empty((max(self.n_halo, 1), F), dtype=x_local.dtype, device=x_local.device)
        send = ops.gather_rows(x_local, p.send_idx) if p.n_send else x_local.new_empty((0, F))
        dist.all_to_all_single(halo[:self.n_halo], send, output_split_sizes=p.recv_counts,
                               input_split_sizes=p.send_counts, group=self.group)
        return halo

    def halo_grad_raw(self, g_halo: torch.Tensor) -> torch.Tensor:
        """reverse exchange on the CURRENT stream: [n_halo, F] gradient rows -> [n_send, F] rows at their owners
        (to be summed per owner row with ``send_csr``).  Peer-memory transport: a view of this rank's window, valid
        until the second-next backward exchange on this graph."""
        p = self.plan
        F = int(g_halo.shape[1])
        win = self._p2p_window(F)
        if win is not None:
            from . import ops
            ph = win.next_phase(False)
            with ops._prof(f"halo_push_bwd_F{F}", self.n_halo * F * 4, self.device):
                win.push(g_halo, None, F, forward=False, phase=ph)
            with ops._prof("halo_wait_post", 0, self.device):
                win.barrier()
            return win.back_view(p.n_send, F, ph)
        back = torch.empty((max(p.n_send, 1), F), dtype=g_halo.dtype, device=g_halo.device)
        dist.all_to_all_single(back[:p.n_send], g_halo[:self.n_halo], output_split_sizes=p.send_counts,
                               input_split_sizes=p.recv_counts, group=self.group)
        return back

    def halo_grad_fused(self, g: torch.Tensor, tgt_scale=None, halo_scale=None) -> torch.Tensor:
        """Backward of the halo part on the CURRENT stream, exchange included: the transposed gather over the
        halo-source structure (one output row per halo row: sum of ``tgt_scale_i * g_i`` over the owned targets i it
        feeds, times ``halo_scale``) stores every finished row STRAIGHT INTO ITS OWNER'S WINDOW over NVLink
        (kgb_gather_reduce with a push table) - the rows travel while the rest of the kernel is still reducing, and
        no [n_halo, F] buffer is written or re-read.  Returns the [n_send, F] rows this rank received (to be summed
        per owned row with ``send_csr``)."""
        from . import _lib, ops
        g_h = self.split[1]
        F = int(g.shape[1])
        win = self._p2p_window(F)
        if win is None:
            g_halo, _ = ops.gather_reduce_raw(g, g_h.csc, _lib.OP_SUM, src_scale=tgt_scale, out_scale=halo_scale)
            return self.halo_grad_raw(g_halo)
        ph = win.next_phase(False)
        if self.n_halo:
            args = win.push_args(F, forward=False, phase=ph)
            dummy = self._dummy_row(F)
            ops.gather_reduce_raw(g, g_h.csc, _lib.OP_SUM, src_scale=tgt_scale, out_scale=halo_scale, out=dummy,
                                  out2_push=args, n_split_out=0, label="halo_grad_push")
        with ops._prof("halo_wait_post", 0, self.device):
            win.barrier()                                   # every rank's rows have landed
        return win.back_view(self.plan.n_send, F, ph)

    def land_into(self, back: torch.Tensor, acc: torch.Tensor) -> torch.Tensor:
        """acc[r] += sum of the returned gradient rows of owned row r, IN PLACE, for the rows some peer asked for
        (deterministic segmented sum in slot order; every other row of ``acc`` is left untouched)."""
        from . import _lib, ops
        if self.send_csr is not None:
            ops.gather_reduce_raw(back, self.send_csr, _lib.OP_SUM, col=self.send_csr.perm, addend=acc,
                                  row_ids=self.send_rows, out=acc, label="halo_grad_land")
        return acc

    def _dummy_row(self, F: int) -> torch.Tensor:
        d = getattr(self, "_dummy", None)
        if d is None or d.shape[1] < F:
            d = self._dummy = torch.empty((1, max(F, 256)), dtype=torch.float32, device=self.device)
        return d[:, :F]

    def close(self) -> None:
        """Release the peer-memory window (collective: every rank must call it)."""
        if self._window is not None:
            self._window.close()
            self._window = None

    def exchange_vector(self, v_local: torch.Tensor) -> torch.Tensor:
        """Per-node scalar (e.g. GCN dis) -> [n_ext]; no autograd."""
        return _exchange_fwd(v_local.reshape(-1, 1).contiguous(), self).reshape(-1)

    def gcn_dis_split(self):
        """(dis [n_local], dis_ext [n_local + n_halo]): deg^-1/2 of the TOTAL in-degree (local + halo sources + self-loops) for
        the owned rows, and the same values of the halo rows (fetched once per graph) - for the split structures."""
        if getattr(self, "_dis_split", None) is None:
            from . import _lib
            from .graph import _stream
            lib = _lib.load()
            g_l, g_h, _ = self.split
            deg = (g_l.csr.deg + g_h.csr.deg).contiguous()
            dis = torch.empty(self.n_local, dtype=torch.float32, device=self.device)
            _lib.check(lib.kgb_gcn_norm(self.device.index, deg.data_ptr(), self.n_local, None, 0, 0, dis.data_ptr(),
                                        None, _stream(self.device)), "kgb_gcn_norm")
            halo = self.halo_rows_raw(dis.reshape(-1, 1).contiguous())[:self.n_halo].reshape(-1)
            self._dis_split = (dis, torch.cat([dis, halo]))     # (owned rows, [owned | halo] rows)
        return self._dis_split

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
    F = int(x_local.shape[1])
    x_ext = torch.empty((pg.n_ext, F), dtype=x_local.dtype, device=x_local.device)
    if pg.world > 1:
        halo = pg.halo_rows_raw(x_local)
        if pg.n_halo:
            x_ext[pg.n_local:].copy_(halo[:pg.n_halo])   # out of the window: x_ext may be saved for a backward
    x_ext[:pg.n_local].copy_(x_local)
    return x_ext


def _land(back: torch.Tensor, pg: PartitionedGraph, F: int, device, dtype) -> torch.Tensor:
    """Deterministic per-owner sum of the returned gradient rows (CSR of the send list)."""
    return pg.land_into(back, torch.zeros((pg.n_local, F), dtype=dtype, device=device))


class _HaloExchange(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, pg: PartitionedGraph):
        ctx.pg = pg
        return _exchange_fwd(x_local.contiguous(), pg)

    @staticmethod
    @once_differentiable
    def backward(ctx, g_ext):
        pg = ctx.pg
        g_ext = g_ext.contiguous()
        g_ext.record_stream(torch.cuda.current_stream(g_ext.device))  # produced on the compute stream
        g_local = g_ext[:pg.n_local].clone()
        if pg.world > 1:
            F = int(g_ext.shape[1])
            back = pg.halo_grad_raw(g_ext[pg.n_local:])
            if pg.plan.n_send:
                g_local += _land(back, pg, F, g_ext.device, g_ext.dtype)
        return g_local, None


class _HaloRows(torch.autograd.Function):
    """x_local [n_local, F] -> halo [n_halo, F]: pack + exchange; backward = reverse exchange + the deterministic
    segmented sum of the returned gradient rows into their owners (zero for rows nobody asked for)."""

    @staticmethod
    def forward(ctx, x_local, pg: PartitionedGraph):
        ctx.pg = pg
        x_local = x_local.contiguous()
        F = int(x_local.shape[1])
        if pg.world > 1:
            halo = pg.halo_rows_raw(x_local)
            if pg._window is not None:
                halo = halo.clone()    # autograd-visible result must not alias the (reused) window
        else:
            halo = torch.zeros((max(pg.n_halo, 1), F), dtype=x_local.dtype, device=x_local.device)
        if pg.n_halo == 0:
            halo.zero_()
        return halo

    @staticmethod
    @once_differentiable
    def backward(ctx, g_halo):
        pg = ctx.pg
        g_halo = g_halo.contiguous()
        g_halo.record_stream(torch.cuda.current_stream(g_halo.device))  # produced on the compute stream
        F = int(g_halo.shape[1])
        if pg.world > 1:
            back = pg.halo_grad_raw(g_halo)
            return _land(back, pg, F, g_halo.device, g_halo.dtype), None
        return torch.zeros((pg.n_local, F), dtype=g_halo.dtype, device=g_halo.device), None
