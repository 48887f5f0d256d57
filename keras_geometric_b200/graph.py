"""Device-resident graph structure (CSR by target + CSC by source) and its cache.

Replaces the reference's int32-cast cache keyed by ``id(edge_index)``
(/root/reference/src/keras_geometric/layers/message_passing.py:257-266) and the implicit
per-call grouping done inside ``keras.ops.segment_*``.  A structure is built once per
``edge_index`` tensor by ``kgb_csr_build`` (include/kgb200.h) and reused by every layer,
forward and backward.
"""
from __future__ import annotations

import os
import weakref
from collections import OrderedDict

import torch

from . import _lib

HUB_THRESHOLD = 512   # rows with more edges than this are cut into chunks ...
HUB_CHUNK = 256       # ... of this many edges (see csrc/gather_reduce.cu)


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"keras_geometric_b200: {what} must live on a CUDA device "
                           "(there is no CPU path; see keras_geometric_b200._lib)")


class Csr:
    """One orientation: ``rowptr`` int64 [n+1], ``col`` int32 [nnz] (the other endpoint),
    ``perm`` int32 [nnz] (original edge id of each slot, stable), ``deg`` int32 [n]."""

    __slots__ = ("n_rows", "n_cols", "nnz", "rowptr", "col", "perm", "deg", "hub_row", "hub_chunk_base",
                 "hub_nchunks", "chunk_hub", "n_hubs", "n_chunks", "_inv_deg", "_partials", "_deg_f", "_work", "_unit_order",
                 "_hot")

    def __init__(self):
        self._inv_deg = None
        self._deg_f = None
        self._partials = {}
        self.n_hubs = 0
        self.n_chunks = 0
        self.hub_row = self.hub_chunk_base = self.hub_nchunks = self.chunk_hub = None
        self._work = {}
        self._unit_order = {}
        self._hot = {}

    def work(self, stream_id: int) -> torch.Tensor:
        """Task-queue counters of kgb_gather_reduce (zero between launches), one pair per stream."""
        w = self._work.get(stream_id)
        if w is None:
            w = torch.zeros(64, dtype=torch.int32, device=self.rowptr.device)
            self._work[stream_id] = w
        return w

    def unit_order(self, unit_rows: int | None = None):
        """Row units of a kernel's task queue, heaviest first (longest-processing-time order): on a power-law
        graph the natural order leaves a tail of ~25 % of the kernel in which most warps have run dry."""
        ur = int(_lib.load().kgb_gather_unit_rows()) if unit_rows is None else int(unit_rows)
        if self.n_rows <= 4 * ur:
            return None
        order = self._unit_order.get(ur)
        if order is None:
            d = self.rowptr[1:] - self.rowptr[:-1]
            if self.n_hubs > 0:
                d = torch.where(d > HUB_THRESHOLD, torch.zeros_like(d), d)   # hub rows go through the chunk tasks
            d = torch.nn.functional.pad(d, (0, (-self.n_rows) % ur)).view(-1, ur).sum(1)
            order = self._unit_order[ur] = torch.sort(d, descending=True, stable=True)[1].to(torch.int32)
        return order

    def col_hot(self, F: int):
        """Copy of ``col`` with bit 31 set on the K most frequently referenced rows, K = budget / row bytes, for
        kgb_gather_reduce's L2 eviction hints (hot rows evict_last, one-touch rows evict_first).  On a power-law
        graph 3 % of the rows carry > 60 % of the references.  Measured on C4 (mean gather forward / transposed
        backward, tools/exp_kernels.py mean): F = 256 6.07 / 6.37 -> 5.71 / 5.92 ms with a 32 MB set; F = 100 2.75 ->
        2.77 ms (64 MB set, no gain), F = 48 a loss.  Tagging costs ~1 ms per (structure, set), so in the default mode
        (KGB200_HOT_MB unset or "auto") a set is only built for rows of >= 800 bytes and only at the THIRD use of the
        structure at that width - a static training graph pays it once, a structure that lives for one step (fresh
        edge list every step) never does.  KGB200_HOT_MB=<n> forces an n MB set from the first use at every width,
        0 turns the hints off.  None when off, not yet due, or when the whole matrix fits in the L2 anyway."""
        row_bytes = 4 * int(F)
        mode = os.environ.get("KGB200_HOT_MB", "auto")
        if mode == "auto":
            budget = (32 << 20) if row_bytes >= 800 else 0
        else:
            budget = int(mode) << 20
        if budget <= 0 or self.nnz < (1 << 20) or self.n_cols * row_bytes <= (96 << 20):
            return None
        K = max(1, budget // row_bytes)
        K = 1 << (K.bit_length() - 1)          # few distinct sets per structure
        if K >= self.n_cols:
            return None
        tagged = self._hot.get(K)
        if tagged is None:
            if mode == "auto":
                uses = self._hot.get(("uses", K), 0) + 1
                self._hot[("uses", K)] = uses
                if uses < 3:
                    return None
            cnt = torch.bincount(self.col.long(), minlength=self.n_cols)
            thr = torch.topk(cnt, K).values[-1].clamp(min=2)        # rows gathered once gain nothing from residency
            hot = (cnt >= thr)[self.col.long()]
            tagged = torch.where(hot, self.col | torch.tensor(-2 ** 31, dtype=torch.int32, device=self.col.device),
                                 self.col).contiguous()
            self._hot[K] = tagged
        return tagged

    def hub_table(self, partial_bytes: int, stream_id: int, gat: bool = False):
        """ctypes ``kgb_hub_table`` for the GATv2 kernels (keeps the partial buffer alive on self)."""
        t = _lib.HubTable()
        if self.n_chunks > 0:
            key = ("gat", partial_bytes)
            buf = self._partials.get(key)
            if buf is None:
                buf = torch.empty(max(partial_bytes // 4, 1), dtype=torch.float32, device=self.rowptr.device)
                self._partials[key] = buf
            t.hub_row, t.hub_chunk_base = self.hub_row.data_ptr(), self.hub_chunk_base.data_ptr()
            t.hub_nchunks, t.chunk_hub = self.hub_nchunks.data_ptr(), self.chunk_hub.data_ptr()
            t.n_hubs, t.n_chunks, t.threshold, t.chunk = self.n_hubs, self.n_chunks, HUB_THRESHOLD, HUB_CHUNK
            t.partial = buf.data_ptr()
        t.work = self.work(stream_id).data_ptr()
        if gat:
            order = self.unit_order(int(_lib.load().kgb_gatv2_unit_rows()))
            t.unit_order = order.data_ptr() if order is not None else None
        return t

    @property
    def inv_deg(self) -> torch.Tensor:
        """1 / max(deg, 1e-8) as float32 (the mean aggregator's divisor, aggregators.py:77-81)."""
        if self._inv_deg is None:
            self._inv_deg = 1.0 / torch.clamp(self.deg.to(torch.float32), min=1e-8)
        return self._inv_deg

    def partial(self, F: int, op: int) -> torch.Tensor | None:
        if self.n_chunks == 0:
            return None
        nbytes = _lib.load().kgb_gather_reduce_partial_bytes(self.n_chunks, F, op)
        key = ("gr", nbytes)
        buf = self._partials.get(key)
        if buf is None:
            buf = torch.empty(nbytes // 4, dtype=torch.float32, device=self.rowptr.device)
            self._partials = {k: v for k, v in self._partials.items() if k[0] != "gr"}
            self._partials[key] = buf
        return buf


class _PendingCsr:
    """A structure whose kernels are enqueued but whose hub counts / status the host has not read yet."""
    __slots__ = ("c", "meta", "keep", "event", "limits", "oob_msg")


_SIDE_STREAMS: dict = {}


def _side_stream(dev) -> "torch.cuda.Stream":
    st = _SIDE_STREAMS.get(dev.index)
    if st is None:
        st = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return st


def _launch_csr(edge_index: torch.Tensor, n_seg: int, n_val: int, n_loops: int, by_source: bool,
                side: "torch.cuda.Stream | None" = None) -> _PendingCsr:
    """Allocate the outputs on the current stream and enqueue kgb_csr_build + kgb_csr_hubs - on ``side`` when given
    (it first waits for the current stream, so reused allocator blocks are safe to overwrite)."""
    require_cuda(edge_index, "edge_index")
    assert edge_index.dtype == torch.int32 and edge_index.is_contiguous() and edge_index.dim() == 2
    lib = _lib.load()
    dev = edge_index.device
    E = int(edge_index.shape[1])
    M = E + n_loops
    c = Csr()
    c.n_rows, c.n_cols, c.nnz = n_seg, n_val, M
    c.rowptr = torch.empty(n_seg + 1, dtype=torch.int64, device=dev)
    c.col = torch.empty(max(M, 1), dtype=torch.int32, device=dev)[:M]
    c.perm = torch.empty(max(M, 1), dtype=torch.int32, device=dev)[:M]
    c.deg = torch.empty(max(n_seg, 1), dtype=torch.int32, device=dev)[:n_seg]
    meta = torch.zeros(4, dtype=torch.int32, device=dev)  # [status, n_hubs, n_chunks, -]
    ws_bytes = lib.kgb_csr_build_workspace_bytes(M, n_seg)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    max_hubs = M // HUB_THRESHOLD + 1
    max_chunks = M // HUB_CHUNK + max_hubs + 1
    hub = torch.empty(3 * max_hubs + max_chunks, dtype=torch.int32, device=dev)
    c.hub_row, c.hub_chunk_base, c.hub_nchunks = hub[:max_hubs], hub[max_hubs:2 * max_hubs], hub[2 * max_hubs:3 * max_hubs]
    c.chunk_hub = hub[3 * max_hubs:]
    pend = _PendingCsr()
    pend.c, pend.meta, pend.keep, pend.event, pend.limits = c, meta, [ws, edge_index], None, (max_hubs, max_chunks)
    pend.oob_msg = (f"edge_index contains node ids outside [0, {n_seg if not by_source else n_val}) / "
                    f"[0, {n_val if not by_source else n_seg})")
    if side is not None:
        side.wait_stream(torch.cuda.current_stream(dev))
        for t in (c.rowptr, c.col, c.perm, c.deg, meta, ws, hub):
            t.record_stream(side)   # freed blocks are not handed out again before the side stream is done with them
    st = side.cuda_stream if side is not None else _stream(dev)
    _lib.check(lib.kgb_csr_build(dev.index, edge_index.data_ptr(), E, int(by_source), n_seg, n_val, n_loops,
                                 c.rowptr.data_ptr(), c.col.data_ptr(), c.perm.data_ptr(), c.deg.data_ptr(),
                                 meta.data_ptr(), ws.data_ptr(), ws_bytes, st), "kgb_csr_build")
    _lib.check(lib.kgb_csr_hubs(dev.index, c.rowptr.data_ptr(), n_seg, HUB_THRESHOLD, HUB_CHUNK,
                                c.hub_row.data_ptr(), c.hub_chunk_base.data_ptr(), c.hub_nchunks.data_ptr(),
                                c.chunk_hub.data_ptr(), max_hubs, max_chunks, meta[1:].data_ptr(), st),
               "kgb_csr_hubs")
    if side is not None:
        pend.event = torch.cuda.Event()
        pend.event.record(side)
    return pend


def _finish_csr(pend: _PendingCsr) -> Csr:
    c = pend.c
    if pend.event is not None:   # built on the side stream: order the consumer's stream behind it
        torch.cuda.current_stream(c.rowptr.device).wait_event(pend.event)
    status, n_hubs, n_chunks, _ = pend.meta.tolist()  # the one host sync of a structure build
    pend.keep = None
    if status & _lib.STATUS_OOB_INDEX:
        raise IndexError(pend.oob_msg)
    assert n_hubs <= pend.limits[0] and n_chunks <= pend.limits[1]
    c.n_hubs, c.n_chunks = int(n_hubs), int(n_chunks)
    return c


def build_csr(edge_index: torch.Tensor, n_seg: int, n_val: int, n_loops: int, by_source: bool) -> Csr:
    """Run kgb_csr_build (+ hub table).  ``edge_index`` int32 [2,E] contiguous on CUDA."""
    return _finish_csr(_launch_csr(edge_index, n_seg, n_val, n_loops, by_source))


class GraphStructure:
    """CSR (by target, forward) and lazily CSC (by source, backward) of one edge list.

    ``n_dst`` target rows, ``n_src`` source rows (equal unless bipartite); ``n_loops`` self-loops
    are appended after the real edges like ``add_self_loops`` does (utils/main.py:8-16)."""

    def __init__(self, edge_index: torch.Tensor, n_dst: int, n_src: int | None = None, n_loops: int = 0):
        n_src = n_dst if n_src is None else n_src
        if edge_index.dtype != torch.int32 or not edge_index.is_contiguous():
            edge_index = edge_index.to(torch.int32).contiguous()
        self.edge_index = edge_index
        self.n_dst, self.n_src, self.n_loops = int(n_dst), int(n_src), int(n_loops)
        self.E = int(edge_index.shape[1])
        self.nnz = self.E + self.n_loops
        self.device = edge_index.device
        self.csr = build_csr(edge_index, self.n_dst, self.n_src, self.n_loops, by_source=False)
        self._csc = None
        self._csc_pending = None
        self._gcn = None

    @property
    def csc(self) -> Csr:
        if self._csc is None:
            if self._csc_pending is not None:
                pend, self._csc_pending = self._csc_pending, None
                self._csc = _finish_csr(pend)
            else:
                self._csc = build_csr(self.edge_index, self.n_src, self.n_dst, self.n_loops, by_source=True)
        return self._csc

    def prefetch_csc(self) -> None:
        """Enqueue the build of the source-major orientation on a side stream without waiting for it: called by the
        forward of every op whose backward walks the CSC, so that on a fresh edge list the second radix sort runs
        beside the forward pass instead of in front of the backward one.  The host reads the hub counts when ``csc``
        is first touched.  No-op when the CSC exists, is already on its way, or the graph is small.
        Opt-in (KGB200_PREFETCH_CSC=1): on the C4 end-to-end step it changed nothing (59.65 vs 59.62 ms) - the
        persistent gather CTAs fill the register files and the GEMM CTAs the shared memory, so the sort kernels of the
        side stream only run in the gaps the main stream leaves anyway."""
        if self._csc is not None or self._csc_pending is not None or self.nnz < (1 << 20):
            return
        if os.environ.get("KGB200_PREFETCH_CSC", "0") != "1":
            return
        self._csc_pending = _launch_csr(self.edge_index, self.n_src, self.n_dst, self.n_loops, by_source=True,
                                        side=_side_stream(self.device))

    def csc_to_csr(self) -> torch.Tensor:
        """slot_map[k'] = CSR slot of the edge stored in slot k' of the CSC (int32 [nnz]); built once per graph."""
        if getattr(self, "_c2r", None) is None:
            inv = torch.empty(self.nnz, dtype=torch.int32, device=self.device)
            inv[self.csr.perm.long()] = torch.arange(self.nnz, dtype=torch.int32, device=self.device)
            self._c2r = inv[self.csc.perm.long()].contiguous()
        return self._c2r

    def gcn_norm(self):
        """(dis [n_dst], w_coo [nnz]) of utils/main.py:20-33 via kgb_gcn_norm."""
        if self._gcn is None:
            if self.n_dst != self.n_src:
                raise ValueError("GCN normalisation needs a square graph")
            lib = _lib.load()
            dis = torch.empty(self.n_dst, dtype=torch.float32, device=self.device)
            w = torch.empty(max(self.nnz, 1), dtype=torch.float32, device=self.device)[:self.nnz]
            _lib.check(lib.kgb_gcn_norm(self.device.index, self.csr.deg.data_ptr(), self.n_dst,
                                        self.edge_index.data_ptr(), self.E, self.n_loops, dis.data_ptr(),
                                        w.data_ptr(), _stream(self.device)), "kgb_gcn_norm")
            self._gcn = (dis, w)
        return self._gcn

    def full_edge_index(self) -> torch.Tensor:
        """edge_index with the self-loops materialised ([2, E + n_loops], int32)."""
        if self.n_loops == 0:
            return self.edge_index
        loop = torch.arange(self.n_loops, dtype=torch.int32, device=self.device)
        return torch.cat([self.edge_index, torch.stack([loop, loop])], dim=1)


# ---- cache -------------------------------------------------------------------------------
_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_CACHE_SIZE = 8


def clear_cache() -> None:
    _CACHE.clear()


def get_graph(edge_index: torch.Tensor, n_dst: int, n_src: int | None = None, n_loops: int = 0) -> GraphStructure:
    """Structure of ``edge_index``, cached on the identity of the tensor object.

    The key is ``(id(tensor), _version, shape, sizes)`` and the entry holds a weak reference to
    the tensor, so neither an in-place update nor a recycled address can return a stale
    structure (the reference's ``id()`` cache has both problems)."""
    n_src = n_dst if n_src is None else n_src
    key = (id(edge_index), edge_index._version, tuple(edge_index.shape), edge_index.dtype, int(n_dst), int(n_src),
           int(n_loops))
    hit = _CACHE.get(key)
    if hit is not None:
        ref, g = hit
        if ref() is edge_index:
            _CACHE.move_to_end(key)
            return g
        del _CACHE[key]
    g = GraphStructure(edge_index, n_dst, n_src, n_loops)
    try:
        _CACHE[key] = (weakref.ref(edge_index), g)
    except TypeError:
        return g
    while len(_CACHE) > _CACHE_SIZE:
        _CACHE.popitem(last=False)
    return g
