"""torch.autograd wrappers around the C ABI (include/kgb200.h).

Every function here launches hand-written sm_100a kernels from libkgb200.so on the current
CUDA stream with borrowed ``data_ptr()``s.  No function has a CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from .graph import HUB_CHUNK, HUB_THRESHOLD, Csr, GraphStructure, _stream, require_cuda

_ACT = {None: _lib.ACT_NONE, "linear": _lib.ACT_NONE, "relu": _lib.ACT_RELU}

# When set to a list, every gather-reduce launch appends {label, start, end, bytes} with CUDA events
# recorded on the launching stream (bench.py reads it for the per-kernel roofline).
PROFILE = None
_OP_NAMES = {0: "sum", 1: "mean", 2: "max", 3: "min", 4: "max", 5: "sqdev"}


def _algorithmic_bytes(nnz, n_rows, F, per_edge_extra, per_row_extra):
    return nnz * (4 * F + 4 + per_edge_extra) + n_rows * (4 * F + per_row_extra) + (n_rows + 1) * 8


class _prof:
    """``with _prof(label, algorithmic_bytes, device):`` records the launch(es) inside it in PROFILE (CUDA events on
    the launching stream) when profiling is on; free otherwise."""

    __slots__ = ("rec", "dev")

    def __init__(self, label: str, nbytes: int, dev):
        self.rec = None
        if PROFILE is not None:
            self.rec = {"label": label, "bytes": int(nbytes), "start": torch.cuda.Event(enable_timing=True),
                        "end": torch.cuda.Event(enable_timing=True)}
            self.dev = dev

    def __enter__(self):
        if self.rec is not None:
            self.rec["start"].record(torch.cuda.current_stream(self.dev))
        return self

    def __exit__(self, *exc):
        if self.rec is not None:
            self.rec["end"].record(torch.cuda.current_stream(self.dev))
            if exc[0] is None and PROFILE is not None:
                PROFILE.append(self.rec)
        return False


def next_dropout_seed() -> int:
    """A fresh 63-bit seed for an in-kernel dropout mask, drawn from torch's default CPU generator (so
    ``torch.manual_seed`` makes training runs reproducible)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def dropout_mask(edge_ids, n_edges: int, width: int, p: float, seed: int, per_head: bool = False, device=None):
    """The mask the kernels apply ([n_edges, width]: 1 / (1 - p) or 0), written out by kgb_dropout_mask - for tests."""
    lib = _lib.load()
    dev = edge_ids.device if edge_ids is not None else torch.device(device or "cuda")
    out = torch.empty((n_edges, width), dtype=torch.float32, device=dev)
    _lib.check(lib.kgb_dropout_mask(dev.index if dev.index is not None else torch.cuda.current_device(), _ptr(edge_ids),
                                    n_edges, width, int(per_head), float(p), int(seed), out.data_ptr(), _stream(dev)),
               "kgb_dropout_mask")
    return out


def gat_bytes(nnz: int, n: int, H: int, C: int) -> int:
    """B_fwd(gatv2) = E'(4HC + 4) + N*4HC (h_i) + N*4HC (out) + N*H*8 (max, denom) + (N+1)*8."""
    return nnz * (4 * H * C + 4) + 2 * n * 4 * H * C + n * H * 8 + (n + 1) * 8


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    require_cuda(t, what)
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if t.dim() == 2 and t.stride(1) != 1:
        t = t.contiguous()
    elif t.dim() != 2 and not t.is_contiguous():
        t = t.contiguous()
    return t


def _ptr(t):
    return None if t is None else t.data_ptr()


def gather_reduce_raw(x: torch.Tensor, csr: Csr, op: int, *, col: torch.Tensor | None = None,
                      edge_w=None, src_scale=None, out_scale=None, addend=None, addend_scale: float = 1.0,
                      bias=None, act: int = 0, want_arg: bool = False, row_ids=None, out=None,
                      n_out_rows: int | None = None, x2=None, n_split_src: int | None = None, out2=None,
                      out2_push=None, n_split_out: int | None = None, label: str | None = None,
                      drop_p: float = 0.0, drop_seed: int = 0):
    """One kgb_gather_reduce launch (no autograd).  Returns ``(out, arg_or_None)``.
    ``x2`` / ``n_split_src``: column ids >= n_split_src read ``x2`` (partitioned graphs: [owned | halo] sources without
    a concatenated copy).  ``out2`` or ``out2_push`` (a ``HaloPushArgs`` destination table) / ``n_split_out``: output
    rows >= n_split_out go to a second buffer or straight into the peers' windows."""
    lib = _lib.load()
    require_cuda(x, "x")
    F = int(x.shape[1])
    dev = x.device
    n_rows = csr.n_rows
    n_out = n_rows if n_out_rows is None else n_out_rows
    if out is None:
        out = torch.empty((n_out, F), dtype=torch.float32, device=dev)
    arg = torch.empty((out.shape[0], out.stride(0)), dtype=torch.int32, device=dev) if want_arg else None
    if n_rows == 0 or F == 0:
        return out, (arg[:, :F] if arg is not None else None)
    a = _lib.GatherReduceArgs()
    a.x, a.ldx, a.n_src_rows, a.F, a.op = x.data_ptr(), x.stride(0), x.shape[0], F, op
    a.rowptr, a.col, a.n_rows = csr.rowptr.data_ptr(), (csr.col if col is None else col).data_ptr(), n_rows
    a.row_ids = _ptr(row_ids)
    a.edge_w, a.src_scale, a.out_scale = _ptr(edge_w), _ptr(src_scale), _ptr(out_scale)
    a.addend = _ptr(addend)
    a.ld_addend = addend.stride(0) if addend is not None else 0
    a.addend_scale = float(addend_scale)
    a.bias, a.act = _ptr(bias), act
    a.out, a.ldo, a.arg = out.data_ptr(), out.stride(0), _ptr(arg)
    partial = csr.partial(F, op)
    if partial is not None:
        a.hub_row, a.hub_chunk_base = csr.hub_row.data_ptr(), csr.hub_chunk_base.data_ptr()
        a.hub_nchunks, a.chunk_hub = csr.hub_nchunks.data_ptr(), csr.chunk_hub.data_ptr()
        a.n_hubs, a.n_chunks = csr.n_hubs, csr.n_chunks
        a.hub_threshold, a.hub_chunk = HUB_THRESHOLD, HUB_CHUNK
        a.partial = partial.data_ptr()
    a.work = csr.work(_stream(dev)).data_ptr()
    a.unit_order = _ptr(csr.unit_order())
    if drop_p > 0.0:   # element-wise dropout of the gathered rows, regenerated from the original edge ids
        a.drop_p, a.drop_seed, a.edge_id = float(drop_p), int(drop_seed), csr.perm.data_ptr()
    elif (col is None and x2 is None and out2 is None and out2_push is None and op != _lib.OP_SQDEV
          and op not in _lib.MAX_OPS):
        # L2 eviction hints for the hub rows (None: off / not yet due / everything fits).  Not for max / min: those
        # kernels are issue-bound (ncu: 6.0 G instructions at 78 % issue utilisation for F = 256), and the hint
        # bookkeeping adds 8 % more instructions
        a.col_hot = _ptr(csr.col_hot(F))
    if x2 is not None:
        a.x2, a.ldx2, a.n_split_src = x2.data_ptr(), x2.stride(0), int(n_split_src)
    if out2 is not None or out2_push is not None:
        a.n_split_out = int(n_split_out)
        if out2_push is not None:
            a.out2_push = ctypes.addressof(out2_push)   # the caller's struct outlives this call
        else:
            a.out2, a.ldo2 = out2.data_ptr(), out2.stride(0)
    rec = None
    if PROFILE is not None:
        rec = {"start": torch.cuda.Event(enable_timing=True), "end": torch.cuda.Event(enable_timing=True)}
        rec["start"].record(torch.cuda.current_stream(dev))
    _lib.check(lib.kgb_gather_reduce(dev.index, ctypes.byref(a), _stream(dev)), "kgb_gather_reduce")
    if rec is not None:
        rec["end"].record(torch.cuda.current_stream(dev))
        per_edge = (4 if edge_w is not None else 0) + (4 if src_scale is not None else 0)
        per_row = (4 if out_scale is not None else 0) + (4 * F if addend is not None else 0) + (4 * F if want_arg else 0)
        rec["bytes"] = _algorithmic_bytes(csr.nnz, n_rows, F, per_edge, per_row)
        rec["label"] = label or (f"gather_reduce_{_OP_NAMES[op]}_F{F}" + ("_w" if per_edge else "")
                                 + ("_epi" if (addend is not None or bias is not None or act) else ""))
        if label:
            rec["label"] += f"_F{F}"
        PROFILE.append(rec)
    return out, (arg[:, :F] if arg is not None else None)


def _max_bwd(g, arg, out, x, csr: Csr, col, op: int, n_src_rows: int):
    lib = _lib.load()
    F = int(g.shape[1])
    gx = torch.zeros((n_src_rows, F), dtype=torch.float32, device=g.device)
    if csr.n_rows and F:
        st = _stream(g.device)
        hubs = csr.hub_table(lib.kgb_gather_max_bwd_workspace_bytes(csr.n_hubs, csr.n_chunks, F), st)
        # fixed-point accumulators of the deterministic scatter (freed right after the call; stream-ordered)
        acc = torch.empty(lib.kgb_gather_max_bwd_acc_bytes(n_src_rows, F), dtype=torch.uint8, device=g.device)
        # B_bwd(max) = N*F*(4 g + 4 arg) + N*F*4 (zero-init) + N*F*4 (scattered writes)
        with _prof(f"gather_max_bwd_F{F}", csr.n_rows * F * 16, g.device):
            _lib.check(lib.kgb_gather_max_bwd(g.device.index, g.data_ptr(), g.stride(0), arg.data_ptr(),
                                              out.data_ptr(), out.stride(0), x.data_ptr(), x.stride(0),
                                              csr.rowptr.data_ptr(), col.data_ptr(), None, csr.n_rows, F, op,
                                              gx.data_ptr(), gx.stride(0), n_src_rows, acc.data_ptr(),
                                              ctypes.byref(hubs), st), "kgb_gather_max_bwd")
    return gx


def gather_rows(src: torch.Tensor, idx: torch.Tensor | None, scale: float = 1.0, out=None) -> torch.Tensor:
    """out[r,:] = scale * src[idx[r],:] via kgb_gather_rows."""
    lib = _lib.load()
    require_cuda(src, "src")
    n_out = int(idx.shape[0]) if idx is not None else int(src.shape[0])
    F = int(src.shape[1])
    if out is None:
        out = torch.empty((n_out, F), dtype=torch.float32, device=src.device)
    if n_out and F:
        _lib.check(lib.kgb_gather_rows(src.device.index, src.data_ptr(), src.stride(0), _ptr(idx), n_out, F,
                                       float(scale), out.data_ptr(), out.stride(0), _stream(src.device)),
                   "kgb_gather_rows")
    return out


def relu_bwd(g: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """g * (y > 0) in one pass (kgb_relu_bwd); g and y are [rows, F] with unit column stride."""
    lib = _lib.load()
    rows, F = int(g.shape[0]), int(g.shape[1])
    out = torch.empty((rows, F), dtype=torch.float32, device=g.device)
    if rows and F:
        _lib.check(lib.kgb_relu_bwd(g.device.index, g.data_ptr(), g.stride(0), y.data_ptr(), y.stride(0), rows, F,
                                    out.data_ptr(), _stream(g.device)), "kgb_relu_bwd")
    return out


def _rows16(t: torch.Tensor) -> bool:
    return (t.dim() == 2 and t.dtype == torch.float32 and t.stride(1) == 1 and t.stride(0) % 4 == 0
            and t.data_ptr() % 16 == 0 and t.shape[1] % 4 == 0 and 0 < t.shape[1] <= 1024)


def relu_bwd_colsum(g: torch.Tensor, y: torch.Tensor | None, want_masked: bool = True):
    """(g * (y > 0), column sums of that) in ONE pass over g (kgb_relu_bwd_colsum): the ReLU backward and the bias
    gradient of a layer.  ``y`` None: no mask, returns ``(g, g.sum(0))``."""
    rows = int(g.shape[0])
    if rows == 0 or not _rows16(g) or (y is not None and not _rows16(y)):
        gp = relu_bwd(g, y) if y is not None else g
        return gp, gp.sum(dim=0)
    lib = _lib.load()
    F = int(g.shape[1])
    dev = g.device
    out = torch.empty((rows, F), dtype=torch.float32, device=dev) if y is not None else None
    n_parts = int(lib.kgb_colsum_parts(dev.index, rows))
    parts = torch.empty((n_parts, F), dtype=torch.float32, device=dev)
    _lib.check(lib.kgb_relu_bwd_colsum(dev.index, g.data_ptr(), g.stride(0), _ptr(y), y.stride(0) if y is not None else 0,
                                       rows, F, _ptr(out), F, parts.data_ptr(), n_parts, _stream(dev)),
               "kgb_relu_bwd_colsum")
    if n_parts == 1:
        col = parts[0]
    else:
        col = torch.empty(F, dtype=torch.float32, device=dev)
        _lib.check(lib.kgb_reduce_parts(dev.index, parts.data_ptr(), n_parts, F, col.data_ptr(), _stream(dev)),
                   "kgb_reduce_parts")
    return (out if y is not None else g), col


class _SoftmaxXent(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        lib = _lib.load()
        require_cuda(logits, "logits")
        if logits.dtype != torch.float32 or logits.dim() != 2 or logits.stride(1) != 1:
            logits = logits.to(torch.float32).contiguous()
        labels = labels.to(torch.int64).contiguous()
        rows, C = int(logits.shape[0]), int(logits.shape[1])
        if labels.shape[0] != rows:
            raise ValueError(f"softmax_cross_entropy: {rows} rows of logits but {labels.shape[0]} labels")
        row_loss = torch.empty(rows, dtype=torch.float32, device=logits.device)
        _lib.check(lib.kgb_softmax_xent_fwd(logits.device.index, logits.data_ptr(), logits.stride(0), labels.data_ptr(),
                                            rows, C, row_loss.data_ptr(), _stream(logits.device)), "kgb_softmax_xent_fwd")
        ctx.save_for_backward(logits, labels)
        return row_loss.mean()

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        logits, labels = ctx.saved_tensors
        lib = _lib.load()
        rows, C = int(logits.shape[0]), int(logits.shape[1])
        ldd = (C + 3) // 4 * 4   # padded rows: the gradient of a [:, :C] slice of a padded buffer is then a view
        buf = torch.empty((rows, ldd), dtype=torch.float32, device=logits.device)
        if ldd != C:
            buf[:, C:].zero_()
        g = g.to(torch.float32).reshape(1).contiguous()
        _lib.check(lib.kgb_softmax_xent_bwd(logits.device.index, logits.data_ptr(), logits.stride(0), labels.data_ptr(),
                                            rows, C, g.data_ptr(), 1.0 / max(rows, 1), buf.data_ptr(), ldd,
                                            _stream(logits.device)), "kgb_softmax_xent_bwd")
        return buf[:, :C], None


def softmax_cross_entropy(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """Mean softmax cross-entropy over integer labels (== torch.nn.functional.cross_entropy) in one pass forward
    and one pass backward; C <= 1024."""
    if int(logits.shape[1]) > 1024:
        raise ValueError("softmax_cross_entropy supports up to 1024 classes")
    return _SoftmaxXent.apply(logits, labels)


class _L2Normalize(torch.autograd.Function):
    """y = x / max(||x||_2, eps) per row (keras.ops.normalize(axis=-1, order=2), layers/sage_conv.py:432-433) in one pass
    forward and one pass backward."""

    @staticmethod
    def forward(ctx, x, eps):
        x = _f32c(x, "x")
        if x.stride(1) != 1:
            x = x.contiguous()
        lib = _lib.load()
        dev = x.device
        rows, F = int(x.shape[0]), int(x.shape[1])
        y = torch.empty((rows, F), dtype=torch.float32, device=dev)
        norm = torch.empty(max(rows, 1), dtype=torch.float32, device=dev)[:rows]
        _lib.check(lib.kgb_l2_normalize(dev.index, x.data_ptr(), x.stride(0), rows, F, float(eps), y.data_ptr(), y.stride(0),
                                        norm.data_ptr(), _stream(dev)), "kgb_l2_normalize")
        ctx.eps = float(eps)
        ctx.save_for_backward(y, norm)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        y, norm = ctx.saved_tensors
        g = _f32c(g, "grad")
        if g.stride(1) != 1:
            g = g.contiguous()
        lib = _lib.load()
        dev = g.device
        rows, F = int(y.shape[0]), int(y.shape[1])
        gx = torch.empty((rows, F), dtype=torch.float32, device=dev)
        _lib.check(lib.kgb_l2_normalize_bwd(dev.index, g.data_ptr(), g.stride(0), y.data_ptr(), y.stride(0), norm.data_ptr(),
                                            rows, F, ctx.eps, gx.data_ptr(), gx.stride(0), _stream(dev)),
                   "kgb_l2_normalize_bwd")
        return gx, None


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """Row-wise ``x / max(||x||_2, eps)`` on the row kernel (K10)."""
    require_cuda(x, "x")
    if x.dim() != 2 or int(x.shape[1]) == 0:
        raise ValueError("l2_normalize expects a [rows, F] matrix with F > 0")
    return _L2Normalize.apply(x, eps)


def permute_f32(w: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(perm.shape[0], dtype=torch.float32, device=w.device)
    if perm.shape[0]:
        _lib.check(lib.kgb_permute_f32(w.device.index, w.data_ptr(), perm.data_ptr(), perm.shape[0], out.data_ptr(),
                                       _stream(w.device)), "kgb_permute_f32")
    return out


class _GatherReduce(torch.autograd.Function):
    """out = act(scale_out * OP_k(w_k * x[col_k]) + addend_scale * addend + bias)."""

    @staticmethod
    def forward(ctx, x, addend, bias, graph: GraphStructure, op_name: str, weight, addend_scale, act, drop_p=0.0,
                drop_seed=0):
        op = _lib.OPS[op_name]
        ctx.drop = (float(drop_p), int(drop_seed))
        if drop_p > 0.0:
            kw_drop = {"drop_p": float(drop_p), "drop_seed": int(drop_seed)}
            if op in _lib.MAX_OPS or not (weight is None or isinstance(weight, (str, tuple))):
                raise ValueError("fused dropout supports sum / mean without per-edge weight tensors")
        else:
            kw_drop = {}
        x = _f32c(x, "x")
        csr = graph.csr
        if ctx.needs_input_grad[0] and op not in _lib.MAX_OPS:
            graph.prefetch_csc()    # the backward walks the other orientation: build it beside the forward pass
        kw = {}
        ctx.weight_kind = None
        if weight is not None:
            if op in _lib.MAX_OPS:
                raise ValueError("max/min aggregation does not take edge weights")
            if isinstance(weight, str) and weight == "gcn":
                dis, _ = graph.gcn_norm()
                kw = {"src_scale": dis, "out_scale": dis}
                ctx.weight_kind = "scales"
                ctx.scales = (dis, dis)
            elif isinstance(weight, tuple):  # (per-source-row scale [n_src], per-target-row scale [n_dst])
                kw = {"src_scale": weight[0], "out_scale": weight[1]}
                ctx.weight_kind = "scales"
                ctx.scales = (weight[0], weight[1])
            else:
                w = _f32c(weight, "edge weight").reshape(-1)
                if w.shape[0] != graph.nnz:
                    raise ValueError(f"edge weight has {w.shape[0]} entries, graph has {graph.nnz} edges")
                kw = {"edge_w": permute_f32(w, csr.perm)}
                ctx.weight_kind = "edge"
                ctx.w_coo = w
        addend_c = _f32c(addend, "addend") if addend is not None else None
        bias_c = _f32c(bias, "bias") if bias is not None else None
        is_max = op in _lib.MAX_OPS
        if is_max and (addend is not None or bias is not None or act not in (None, "linear")):
            raise ValueError("max/min aggregation cannot be fused with an epilogue (ties are detected on the raw result)")
        out, arg = gather_reduce_raw(x, csr, op, addend=addend_c, addend_scale=addend_scale, bias=bias_c,
                                     act=_ACT[act], want_arg=is_max, **kw, **kw_drop)
        ctx.graph, ctx.op, ctx.act, ctx.addend_scale = graph, op, act, float(addend_scale)
        ctx.has_addend, ctx.has_bias = addend is not None, bias is not None
        ctx.n_src = int(x.shape[0])
        saved = []
        if is_max:
            saved = [x, arg]
        if is_max or act == "relu":
            saved.append(out)
        ctx.save_for_backward(*saved)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        graph, op = ctx.graph, ctx.op
        g = _f32c(g, "grad")
        saved = list(ctx.saved_tensors)
        is_max = op in _lib.MAX_OPS
        out = saved[-1] if (is_max or ctx.act == "relu") else None
        g_bias = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            g, g_bias = relu_bwd_colsum(g, out if ctx.act == "relu" else None)
        elif ctx.act == "relu":
            g = relu_bwd(g, out)
        g_addend = (g * ctx.addend_scale if ctx.addend_scale != 1.0 else g) if ctx.has_addend else None
        gx = None
        if ctx.needs_input_grad[0]:
            if is_max:
                x, arg = saved[0], saved[1]
                gx = _max_bwd(g, arg, out, x, graph.csr, graph.csr.col, op, ctx.n_src)
            else:
                csc = graph.csc
                kw = {}
                if ctx.weight_kind == "scales":  # transposed: targets are gathered, sources are written
                    kw = {"src_scale": ctx.scales[1], "out_scale": ctx.scales[0]}
                elif ctx.weight_kind == "edge":
                    kw = {"edge_w": permute_f32(ctx.w_coo, csc.perm)}
                if op == _lib.OP_MEAN:
                    kw["src_scale"] = graph.csr.inv_deg
                if ctx.drop[0] > 0.0:   # the transposed pass regenerates the forward's mask from the edge ids
                    kw.update(drop_p=ctx.drop[0], drop_seed=ctx.drop[1])
                gx, _ = gather_reduce_raw(g, csc, _lib.OP_SUM, **kw)
        return gx, g_addend, g_bias, None, None, None, None, None, None, None


def gather_reduce(x, graph: GraphStructure, op: str = "sum", *, weight=None, addend=None,
                  addend_scale: float = 1.0, bias=None, act=None, dropout: float = 0.0,
                  dropout_seed: int | None = None) -> torch.Tensor:
    """Fused gather + segmented reduction over ``graph`` (K3/K4, backward K5).

    Equivalent to the reference's ``take(x, src)`` -> message -> ``Aggregator.aggregate`` chain
    (layers/message_passing.py:195-212) without materialising any [E, F] tensor.
    ``weight``: None, ``"gcn"`` (symmetric normalisation, utils/main.py:20-33), a
    ``(src_scale [n_src], dst_scale [n_dst])`` pair (w_e = dst_scale[i] * src_scale[j]) or a COO-ordered
    [nnz] tensor.  ``dropout`` > 0: fused element-wise dropout of the gathered rows (sum / mean)."""
    if op not in _lib.OPS:
        raise ValueError(f"Invalid aggregator: {op}. Available aggregators: {list(_lib.OPS)}")
    if dropout > 0.0:
        # element-wise dropout of the gathered rows x[src] BEFORE weighting / reduction (the reference's per-edge
        # message dropout), generated in the kernel: no [E, F] tensor, the backward regenerates the mask
        seed = next_dropout_seed() if dropout_seed is None else int(dropout_seed)
        return _GatherReduce.apply(x, addend, bias, graph, op, weight, addend_scale, act, float(dropout), seed)
    return _GatherReduce.apply(x, addend, bias, graph, op, weight, addend_scale, act)


class _GatherStd(torch.autograd.Function):
    """Population standard deviation of the gathered rows per target (``StdAggregator``, layers/aggregators.py:174-232)
    in two fused passes over the structure - the mean, then sqrt(mean squared deviation) with the row's mean held in
    registers - instead of the reference's five [E, F] temporaries.  ``messages`` False: rows are x[col] (fused with
    the gather); True: ``x`` holds one message per edge in COO order (the generic ``Aggregator.aggregate``).
    Backward: d std_i / d m_e = (m_e - mean_i) / (n_i std_i); like the reference's autograd it is NaN for every edge
    of a row whose variance is exactly 0 (sqrt'(0) = inf times a zero deviation), including the rows with a single
    message whose output is forced to 0."""

    @staticmethod
    def forward(ctx, x, graph: GraphStructure, messages: bool):
        x = _f32c(x, "x")
        csr = graph.csr
        col = csr.perm if messages else None
        mean, _ = gather_reduce_raw(x, csr, _lib.OP_MEAN, col=col)
        std, _ = gather_reduce_raw(x, csr, _lib.OP_SQDEV, col=col, addend=mean)
        ctx.graph, ctx.messages = graph, messages
        ctx.save_for_backward(x, mean, std)
        return std

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, mean, std = ctx.saved_tensors
        graph = ctx.graph
        g = _f32c(g, "grad")
        cnt = graph.csr.deg.to(torch.float32).unsqueeze(1)
        c = torch.where(cnt <= 1, torch.zeros_like(g), g) / (torch.clamp(cnt, min=1e-8) * std)   # [n_dst, F]
        b = c * mean
        if ctx.messages:
            dst = graph.full_edge_index()[1]
            return x * gather_rows(c, dst) - gather_rows(b, dst), None, None
        sc, _ = gather_reduce_raw(c, graph.csc, _lib.OP_SUM)      # sum over the targets each source feeds
        sb, _ = gather_reduce_raw(b, graph.csc, _lib.OP_SUM)
        return x * sc - sb, None, None


def gather_std(x, graph: GraphStructure) -> torch.Tensor:
    """std over the in-neighbour rows x[src] of every target, fused with the gather (no [E, F] tensor)."""
    return _GatherStd.apply(x, graph, False)


def segment_std(messages, graph: GraphStructure) -> torch.Tensor:
    """``StdAggregator.aggregate(messages, target_idx, dim_size)`` over a prebuilt structure (COO-ordered messages)."""
    return _GatherStd.apply(messages, graph, True)


class _SegmentReduce(torch.autograd.Function):
    """Generic Aggregator.aggregate over materialised messages [E, F'] (K9)."""

    @staticmethod
    def forward(ctx, messages, graph: GraphStructure, op_name: str):
        op = _lib.OPS[op_name]
        m = _f32c(messages, "messages")
        csr = graph.csr
        is_max = op in _lib.MAX_OPS
        out, arg = gather_reduce_raw(m, csr, op, col=csr.perm, want_arg=is_max)
        ctx.graph, ctx.op, ctx.n_msg = graph, op, int(m.shape[0])
        ctx.save_for_backward(*([m, arg, out] if is_max else []))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        graph, op = ctx.graph, ctx.op
        g = _f32c(g, "grad")
        if op in _lib.MAX_OPS:
            m, arg, out = ctx.saved_tensors
            return _max_bwd(g, arg, out, m, graph.csr, graph.csr.perm, op, ctx.n_msg), None, None
        if op == _lib.OP_MEAN:
            g = g * graph.csr.inv_deg.unsqueeze(1)
        dst = graph.full_edge_index()[1]
        return gather_rows(g, dst), None, None


def segment_reduce(messages, graph: GraphStructure, op: str = "sum") -> torch.Tensor:
    """``Aggregator.aggregate(messages, target_idx, dim_size)`` (layers/aggregators.py:24-39) for
    sum/mean/max/min over a prebuilt structure; messages are in the graph's COO edge order."""
    if op not in _lib.OPS:
        raise ValueError(f"Invalid aggregator: {op}. Available aggregators: {list(_lib.OPS)}")
    return _SegmentReduce.apply(messages, graph, op)


class _TakeRows(torch.autograd.Function):
    """x[idx] with idx one row of a graph's edge list (``ops.take`` of the reference,
    layers/message_passing.py:195-196).  Backward is the segmented sum over the matching
    structure (deterministic), not an atomic scatter."""

    @staticmethod
    def forward(ctx, x, graph: GraphStructure, which: str):
        x = _f32c(x, "x")
        idx = graph.full_edge_index()[0 if which == "src" else 1]
        ctx.graph, ctx.which, ctx.n = graph, which, int(x.shape[0])
        return gather_rows(x, idx)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = _f32c(g, "grad")
        s = ctx.graph.csc if ctx.which == "src" else ctx.graph.csr
        gx, _ = gather_reduce_raw(g, s, _lib.OP_SUM, col=s.perm)
        return gx, None, None


def take_rows(x, graph: GraphStructure, which: str = "src") -> torch.Tensor:
    """Materialise the per-edge rows x[src] / x[dst] ([nnz, F]) for the generic message path."""
    return _TakeRows.apply(x, graph, which)


def _gat_drop(drop, structure: Csr):
    """ctypes ``kgb_gat_dropout`` for a (p, seed) pair over ``structure`` (its ``perm`` = original edge ids), or NULL."""
    if drop[0] <= 0.0:
        return None
    d = _lib.GatDropout()
    d.p, d.seed, d.edge_id = drop[0], drop[1], structure.perm.data_ptr()
    return ctypes.byref(d)


class _GatV2(torch.autograd.Function):
    """Fused GATv2 attention + aggregation over a graph structure (K6)."""

    @staticmethod
    def forward(ctx, h_src, h_dst, att, bias, graph: GraphStructure, H: int, C: int, slope: float, drop_p=0.0,
                drop_seed=0):
        lib = _lib.load()
        ctx.drop = (float(drop_p), int(drop_seed))
        same = h_src is h_dst
        h_src = _f32c(h_src, "h_src").contiguous()
        h_dst = h_src if same else _f32c(h_dst, "h_dst").contiguous()
        att_c = _f32c(att, "att").contiguous()
        bias_c = _f32c(bias, "bias").contiguous() if bias is not None else None
        dev = h_src.device
        n_src, n_dst = int(h_src.shape[0]), int(h_dst.shape[0])
        csr = graph.csr
        if any(ctx.needs_input_grad[:3]):
            graph.prefetch_csc()    # the per-source backward pass walks the other orientation
        out = torch.empty((n_dst, H * C), dtype=torch.float32, device=dev)
        rowmax = torch.empty((n_dst, H), dtype=torch.float32, device=dev)
        rowden = torch.empty((n_dst, H), dtype=torch.float32, device=dev)
        hubs = csr.hub_table(lib.kgb_gatv2_partial_bytes(csr.n_chunks, H, C), _stream(dev), gat=True)
        with _prof(f"gatv2_fwd_H{H}_C{C}", gat_bytes(csr.nnz, n_dst, H, C), dev):
            _lib.check(lib.kgb_gatv2_fwd(dev.index, h_src.data_ptr(), h_dst.data_ptr(), n_src, n_dst, H, C,
                                         att_c.data_ptr(), float(slope), csr.rowptr.data_ptr(), csr.col.data_ptr(),
                                         _ptr(bias_c), out.data_ptr(), rowmax.data_ptr(), rowden.data_ptr(),
                                         _gat_drop(ctx.drop, csr), ctypes.byref(hubs), _stream(dev)), "kgb_gatv2_fwd")
        ctx.graph, ctx.H, ctx.C, ctx.slope, ctx.same = graph, H, C, float(slope), same
        ctx.has_bias = bias is not None
        ctx.save_for_backward(h_src, h_dst, att_c, out, rowmax, rowden, *([bias_c] if bias is not None else []))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        saved = ctx.saved_tensors
        h_src, h_dst, att_c, out, rowmax, rowden = saved[:6]
        g = _f32c(g, "grad").contiguous()
        H, C, graph = ctx.H, ctx.C, ctx.graph
        dev = g.device
        n_src, n_dst = int(h_src.shape[0]), int(h_dst.shape[0])
        bias_c = saved[6] if ctx.has_bias else None
        csr, csc = graph.csr, graph.csc
        st = _stream(dev)
        n_parts = lib.kgb_gatv2_bwd_parts(dev.index, n_dst, H, C)
        if n_parts <= 0:
            raise _lib.KgbError("kgb_gatv2_bwd_parts: unsupported shape")
        g_hdst = torch.empty_like(h_dst)
        r = torch.empty((n_dst, H, 4), dtype=torch.float32, device=dev)   # (max, 1/den, r, -) per (target, head)
        part = torch.empty((n_parts, H * C), dtype=torch.float32, device=dev)
        # per-edge records (alpha * dropout, d logit, sign bits of z): the per-source pass then gathers one row per
        # edge instead of two and recomputes nothing.  Measured on C4 (+ self-loops, tools/exp_gat.py): the per-source
        # pass gets faster (H8C8 8.84 -> 7.32 ms) but writing the records costs the per-target pass more (4.93 ->
        # 9.98 ms; H1C64 5.76 -> 10.4), so the recomputing path stays the default and this one is opt-in
        # (KGB200_GAT_REC=1; covered by test_gatv2_record_path).
        rec_ld = int(lib.kgb_gatv2_rec_floats(H, C)) if (os.environ.get("KGB200_GAT_REC", "0") == "1") else 0
        rec = torch.empty((csr.nnz, rec_ld), dtype=torch.float32, device=dev) if (rec_ld > 0 and csr.nnz > 0) else None
        with _prof(f"gatv2_bwd_dst_H{H}_C{C}", gat_bytes(csr.nnz, n_dst, H, C), dev):
            _lib.check(lib.kgb_gatv2_bwd_dst(dev.index, g.data_ptr(), out.data_ptr(), h_src.data_ptr(), h_dst.data_ptr(),
                                             n_src, n_dst, H, C, att_c.data_ptr(), ctx.slope, csr.rowptr.data_ptr(),
                                             csr.col.data_ptr(), rowmax.data_ptr(), rowden.data_ptr(), _ptr(bias_c),
                                             g_hdst.data_ptr(), r.data_ptr(), part.data_ptr(), n_parts, _ptr(rec),
                                             _gat_drop(ctx.drop, csr),
                                             ctypes.byref(csr.hub_table(lib.kgb_gatv2_partial_bytes(csr.n_chunks, H, C), st, gat=True)),
                                             st), "kgb_gatv2_bwd_dst")
        g_att = torch.empty(H * C, dtype=torch.float32, device=dev)
        _lib.check(lib.kgb_reduce_parts(dev.index, part.data_ptr(), n_parts, H * C, g_att.data_ptr(), st),
                   "kgb_reduce_parts")
        g_hsrc = torch.empty_like(h_src)
        if rec is not None:
            with _prof(f"gatv2_bwd_src_H{H}_C{C}", gat_bytes(csc.nnz, n_src, H, C), dev):
                _lib.check(lib.kgb_gatv2_bwd_src_rec(dev.index, g.data_ptr(), n_src, n_dst, H, C, att_c.data_ptr(),
                                                     ctx.slope, csc.rowptr.data_ptr(), csc.col.data_ptr(),
                                                     graph.csc_to_csr().data_ptr(), rec.data_ptr(),
                                                     g_hdst.data_ptr() if ctx.same else None, g_hsrc.data_ptr(),
                                                     ctypes.byref(csc.hub_table(lib.kgb_gatv2_partial_bytes(csc.n_chunks, H, C), st, gat=True)),
                                                     st), "kgb_gatv2_bwd_src_rec")
            g_bias = relu_bwd_colsum(g, None)[1] if ctx.has_bias else None
            if ctx.same:
                return g_hsrc, None, g_att, g_bias, None, None, None, None, None, None
            return g_hsrc, g_hdst, g_att, g_bias, None, None, None, None, None, None
        with _prof(f"gatv2_bwd_src_H{H}_C{C}", gat_bytes(csc.nnz, n_src, H, C), dev):
            _lib.check(lib.kgb_gatv2_bwd_src(dev.index, g.data_ptr(), h_src.data_ptr(), h_dst.data_ptr(), n_src, n_dst,
                                             H, C, att_c.data_ptr(), ctx.slope, csc.rowptr.data_ptr(),
                                             csc.col.data_ptr(), r.data_ptr(),
                                             g_hdst.data_ptr() if ctx.same else None, g_hsrc.data_ptr(),
                                             _gat_drop(ctx.drop, csc),
                                             ctypes.byref(csc.hub_table(lib.kgb_gatv2_partial_bytes(csc.n_chunks, H, C), st, gat=True)),
                                             st), "kgb_gatv2_bwd_src")
        g_bias = relu_bwd_colsum(g, None)[1] if ctx.has_bias else None   # column sums in one pass (kgb_relu_bwd_colsum)
        if ctx.same:   # the per-target part was added in the per-source kernel's epilogue
            return g_hsrc, None, g_att, g_bias, None, None, None, None, None, None
        return g_hsrc, g_hdst, g_att, g_bias, None, None, None, None, None, None


def gatv2_aggregate(h_src, h_dst, att, graph: GraphStructure, heads: int, channels: int,
                    negative_slope: float = 0.2, bias=None, dropout: float = 0.0,
                    dropout_seed: int | None = None) -> torch.Tensor:
    """Attention logits, per-target softmax and weighted aggregation of GATv2 in one kernel
    (reference: layers/gatv2_conv.py:241-335).  ``h_*`` are [N, H*C], ``att`` has H*C entries;
    returns [n_dst, H*C] (+ bias when given)."""
    if dropout > 0.0:   # attention dropout (training): one decision per (edge, head), generated inside the kernels
        seed = next_dropout_seed() if dropout_seed is None else int(dropout_seed)
        return _GatV2.apply(h_src, h_dst, att.reshape(-1), bias, graph, int(heads), int(channels),
                            float(negative_slope), float(dropout), seed)
    return _GatV2.apply(h_src, h_dst, att.reshape(-1), bias, graph, int(heads), int(channels), float(negative_slope))


# ------------------------------------------------------------------------------------------ K8
_TC_SLAB = 256  # widest output tile of the tcgen05 kernels; wider layers loop column slabs of the weights


def _align4(v: int) -> int:
    return (int(v) + 3) // 4 * 4


def _gemm_ok(*tensors) -> bool:
    """Operands the TMA descriptors can address in place (no padded copy needed)."""
    for t in tensors:
        if t is None:
            continue
        if t.dtype != torch.float32 or not t.is_cuda or t.dim() != 2 or t.stride(1) != 1:
            return False
        if t.data_ptr() % 16 or t.stride(0) % 4 or t.shape[1] % 4:
            return False
    return True


def _tma_rows(t: torch.Tensor, pad_cols: bool = True) -> torch.Tensor:
    """[rows, cols] fp32 operand a TMA descriptor can address: unit column stride, 16-byte aligned base, row stride a
    multiple of 4 floats.  Anything else is copied (kgb_gather_rows) into a buffer whose rows are padded to a multiple
    of 4 floats, pad columns zero; the returned view keeps the logical width.  ``pad_cols``: also guarantee that the
    columns up to the next multiple of 4 are addressable (they hold zeros or the parent buffer's own data)."""
    t = _f32c(t, "gemm operand")
    rows, cols = int(t.shape[0]), int(t.shape[1])
    if t.stride(1) == 1 and t.data_ptr() % 16 == 0 and t.stride(0) % 4 == 0 and (rows > 1 or cols % 4 == 0):
        return t  # a row stride that is a multiple of 4 floats also covers the pad columns
    ld = _align4(cols)
    buf = torch.empty((rows, ld), dtype=torch.float32, device=t.device)
    if ld != cols:
        buf[:, cols:].zero_()
    return gather_rows(t, None, out=buf[:, :cols])


def _split_weight(w: torch.Tensor, transpose: bool):
    """tf32 hi/lo parts of a weight matrix laid out [n_out, k] (zero-padded rows and columns) for kgb_linear_tc."""
    lib = _lib.load()
    rows, cols = int(w.shape[0]), int(w.shape[1])
    n_out, k = (cols, rows) if transpose else (rows, cols)
    bn = lib.kgb_linear_tc_rows(n_out)
    kp = int(lib.kgb_linear_tc2_k(k, 0))
    buf = torch.zeros((2, bn, kp), dtype=torch.float32, device=w.device)
    _lib.check(lib.kgb_split_tf32_ld(w.device.index, w.data_ptr(), rows, cols, w.stride(0), int(transpose),
                                     buf[0].data_ptr(), buf[1].data_ptr(), kp, _stream(w.device)), "kgb_split_tf32_ld")
    return buf[0], buf[1]


def _split_weight_pair(w1: torch.Tensor, w2: torch.Tensor, transpose: bool):
    """tf32 hi/lo parts of [W1 ; W2] stacked along the reduction dimension for kgb_linear_tc2: [n_out, K1p + K2]
    with W1's block zero-padded to a multiple of 32 columns."""
    lib = _lib.load()
    outs = []
    for w in (w1, w2):
        rows, cols = int(w.shape[0]), int(w.shape[1])
        outs.append((cols, rows) if transpose else (rows, cols))
    (n_out, k1), (n_out2, k2) = outs
    if n_out != n_out2:
        raise ValueError("the two weights must produce the same output width")
    kcat = int(lib.kgb_linear_tc2_k(k1, k2))
    bn = lib.kgb_linear_tc_rows(n_out)
    buf = torch.zeros((2, bn, kcat), dtype=torch.float32, device=w1.device)
    for w, off in ((w1, 0), (w2, (k1 + 31) // 32 * 32)):
        _lib.check(lib.kgb_split_tf32_ld(w.device.index, w.data_ptr(), int(w.shape[0]), int(w.shape[1]), w.stride(0),
                                         int(transpose), buf[0, :, off:].data_ptr(), buf[1, :, off:].data_ptr(), kcat,
                                         _stream(w.device)), "kgb_split_tf32_ld")
    return buf[0], buf[1]


def linear_tc2(a1: torch.Tensor, a2: torch.Tensor, w_hi: torch.Tensor, w_lo: torch.Tensor, n_out: int, *, c=None,
               bias=None, relu: bool = False) -> torch.Tensor:
    """One kgb_linear_tc2 launch: [a1 | a2] @ [W1 ; W2] (+ c) (+ bias) (ReLU) on tcgen05 (3xTF32)."""
    lib = _lib.load()
    M = int(a1.shape[0])
    out = torch.empty((M, n_out), dtype=torch.float32, device=a1.device)
    _lib.check(lib.kgb_linear_tc2(a1.device.index, a1.data_ptr(), a1.stride(0), int(a1.shape[1]), a2.data_ptr(),
                                  a2.stride(0), int(a2.shape[1]), M, w_hi.data_ptr(), w_lo.data_ptr(), n_out, _ptr(c),
                                  c.stride(0) if c is not None else 0, _ptr(bias),
                                  _lib.ACT_RELU if relu else _lib.ACT_NONE, out.data_ptr(), out.stride(0),
                                  _stream(a1.device)), "kgb_linear_tc2")
    return out


def linear_tc(a: torch.Tensor, w_hi: torch.Tensor, w_lo: torch.Tensor, n_out: int, *, c=None, bias=None,
              relu: bool = False, out=None) -> torch.Tensor:
    """One kgb_linear_tc launch: a [M,K] @ Wt[n_out,K]^T (+ c) (+ bias) (ReLU) on tcgen05 (3xTF32); ``n_out`` is a
    multiple of 4 and at most 256, ``out`` may be a column block of a wider buffer."""
    lib = _lib.load()
    M, K = int(a.shape[0]), int(a.shape[1])
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.float32, device=a.device)
    _lib.check(lib.kgb_linear_tc(a.device.index, a.data_ptr(), a.stride(0), M, K, w_hi.data_ptr(), w_lo.data_ptr(),
                                 n_out, _ptr(c), c.stride(0) if c is not None else 0, _ptr(bias),
                                 _lib.ACT_RELU if relu else _lib.ACT_NONE, out.data_ptr(), out.stride(0),
                                 _stream(a.device)), "kgb_linear_tc")
    return out


def _matmul_tc(a: torch.Tensor, w: torch.Tensor, transpose_w: bool, *, c=None, bias=None, relu: bool = False):
    """a [M,K] @ W (+ c) (+ bias) (ReLU) for ANY shape on the hand-written tcgen05 kernel.  ``w`` is [K,N]
    (``transpose_w`` False) or [N,K] (True: the dX = G W^T case).  The output is allocated with its width padded to a
    multiple of 4 floats (pad columns are exact zeros) and the logical [:, :N] view is returned; widths above 256 loop
    256-wide column slabs of the weights (the node-feature operand is re-read from L2/HBM once per slab)."""
    M = int(a.shape[0])
    N = int(w.shape[0] if transpose_w else w.shape[1])
    n4 = _align4(N)
    out = torch.empty((M, n4), dtype=torch.float32, device=a.device)
    if M == 0 or N == 0:
        return out[:, :N]
    a = _tma_rows(a)
    if bias is not None and n4 != N:
        bias = torch.nn.functional.pad(bias, (0, n4 - N))
    if c is not None:
        c = _tma_rows(c, pad_cols=True)
    for n0 in range(0, n4, _TC_SLAB):
        ns = min(_TC_SLAB, n4 - n0)
        nr = min(ns, N - n0)  # real weight columns of this slab
        ws = w[n0:n0 + nr, :] if transpose_w else w[:, n0:n0 + nr]
        hi, lo = _split_weight(ws, transpose=not transpose_w)
        linear_tc(a, hi, lo, ns, c=c[:, n0:] if c is not None else None, bias=bias[n0:] if bias is not None else None,
                  relu=relu, out=out[:, n0:])
    return out[:, :N]


def _dw_tc(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """dW[K,N] = X^T G on the hand-written tcgen05 kernel: one node slice per CTA -> per-CTA partials (promoted to
    fp32 every 256 nodes), added in CTA order (deterministic, no float atomics across CTAs).  Widths above 256 / not
    multiples of 4 loop 256-wide slabs of zero-padded operands."""
    lib = _lib.load()
    M, K, N = int(x.shape[0]), int(x.shape[1]), int(g.shape[1])
    dev = x.device
    if M == 0 or K == 0 or N == 0:
        return torch.zeros((K, N), dtype=torch.float32, device=dev)
    x, g = _tma_rows(x, pad_cols=True), _tma_rows(g, pad_cols=True)
    n_parts = lib.kgb_linear_tc_dw_parts(dev.index, M)
    single = K <= _TC_SLAB and N <= _TC_SLAB and K % 4 == 0 and N % 4 == 0
    gw = None if single else torch.empty((K, N), dtype=torch.float32, device=dev)
    for k0 in range(0, K, _TC_SLAB):
        ks = min(_TC_SLAB, K - k0)
        ks4 = _align4(ks)
        for n0 in range(0, N, _TC_SLAB):
            ns = min(_TC_SLAB, N - n0)
            ns4 = _align4(ns)
            parts = torch.empty((n_parts, ks4, ns4), dtype=torch.float32, device=dev)
            _lib.check(lib.kgb_linear_tc_dw(dev.index, x[:, k0:].data_ptr(), x.stride(0), g[:, n0:].data_ptr(),
                                            g.stride(0), M, ks4, ns4, parts.data_ptr(), n_parts, _stream(dev)),
                       "kgb_linear_tc_dw")
            if n_parts == 1:
                slab = parts[0]
            else:
                slab = torch.empty((ks4, ns4), dtype=torch.float32, device=dev)
                _lib.check(lib.kgb_reduce_parts(dev.index, parts.data_ptr(), n_parts, ks4 * ns4, slab.data_ptr(),
                                                _stream(dev)), "kgb_reduce_parts")
            if single:
                return slab
            gather_rows(slab[:ks, :ns], None, out=gw[k0:k0 + ks, n0:n0 + ns])
    return gw


def _dw_tc2(x: torch.Tensor, ga: torch.Tensor, gb: torch.Tensor):
    """(X^T Ga, X^T Gb) in ONE pass over X (kgb_linear_tc_dw2): the two gradient operands sit side by side in the
    kernel's G tile, so X is loaded and split once.  Shapes outside the single-launch range use two _dw_tc calls."""
    lib = _lib.load()
    M, K = int(x.shape[0]), int(x.shape[1])
    Na, Nb = int(ga.shape[1]), int(gb.shape[1])
    ok = (M > 0 and 0 < K <= _TC_SLAB and K % 4 == 0 and Na % 4 == 0 and Nb % 4 == 0 and Na > 0 and Nb > 0
          and int(lib.kgb_linear_tc_dw2_cols(Na, Nb)) <= _TC_SLAB and _gemm_ok(x, ga, gb))
    if not ok:
        return _dw_tc(x, ga), _dw_tc(x, gb)
    dev = x.device
    ncols = int(lib.kgb_linear_tc_dw2_cols(Na, Nb))
    n_parts = lib.kgb_linear_tc_dw_parts(dev.index, M)
    parts = torch.empty((n_parts, K, ncols), dtype=torch.float32, device=dev)
    _lib.check(lib.kgb_linear_tc_dw2(dev.index, x.data_ptr(), x.stride(0), ga.data_ptr(), ga.stride(0), Na,
                                     gb.data_ptr(), gb.stride(0), Nb, M, K, parts.data_ptr(), n_parts, _stream(dev)),
               "kgb_linear_tc_dw2")
    if n_parts == 1:
        both = parts[0]
    else:
        both = torch.empty((K, ncols), dtype=torch.float32, device=dev)
        _lib.check(lib.kgb_reduce_parts(dev.index, parts.data_ptr(), n_parts, K * ncols, both.data_ptr(), _stream(dev)),
                   "kgb_reduce_parts")
    return both[:, :Na].contiguous(), both[:, ncols - Nb:].contiguous()


def _dw_tc_x2(xa: torch.Tensor, xb: torch.Tensor, g: torch.Tensor):
    """(Xa^T G, Xb^T G) in ONE pass over G (kgb_linear_tc_dw_x2): the two feature operands sit one after the other in
    the kernel's X tile, so G is loaded and split once.  Needs ceil32(Ka) + Kb <= 256; otherwise two _dw_tc calls."""
    lib = _lib.load()
    M, N = int(g.shape[0]), int(g.shape[1])
    Ka, Kb = int(xa.shape[1]), int(xb.shape[1])
    ok = (M > 0 and 0 < N <= _TC_SLAB and N % 4 == 0 and Ka % 4 == 0 and Kb % 4 == 0 and Ka > 0 and Kb > 0
          and int(lib.kgb_linear_tc_dw2_cols(Ka, Kb)) <= _TC_SLAB and _gemm_ok(xa, xb, g))
    if not ok:
        return _dw_tc(xa, g), _dw_tc(xb, g)
    dev = g.device
    nrows = int(lib.kgb_linear_tc_dw2_cols(Ka, Kb))
    n_parts = lib.kgb_linear_tc_dw_parts(dev.index, M)
    parts = torch.empty((n_parts, nrows, N), dtype=torch.float32, device=dev)
    _lib.check(lib.kgb_linear_tc_dw_x2(dev.index, xa.data_ptr(), xa.stride(0), Ka, xb.data_ptr(), xb.stride(0), Kb,
                                       g.data_ptr(), g.stride(0), N, M, parts.data_ptr(), n_parts, _stream(dev)),
               "kgb_linear_tc_dw_x2")
    if n_parts == 1:
        both = parts[0]
    else:
        both = torch.empty((nrows, N), dtype=torch.float32, device=dev)
        _lib.check(lib.kgb_reduce_parts(dev.index, parts.data_ptr(), n_parts, nrows * N, both.data_ptr(), _stream(dev)),
                   "kgb_reduce_parts")
    return both[:Ka], both[nrows - Kb:]


class _Linear(torch.autograd.Function):
    """out = act(x @ w + addend + bias) on the tensor cores (K8).  Forward, dX = G W^T and dW = X^T G all run the
    hand-written tcgen05 3xTF32 kernels (csrc/tc_gemm.cu) for every shape: ragged widths are zero-padded to multiples
    of 4 floats, widths above 256 loop column slabs, short row counts rely on the TMA's out-of-bounds zero fill."""

    @staticmethod
    def forward(ctx, x, w, addend, bias, act):
        x = _f32c(x, "x")
        w = _f32c(w, "w")
        relu = act == "relu"
        if act not in (None, "linear", "relu"):
            raise ValueError(f"linear: unsupported fused activation {act!r}")
        if x.dim() != 2 or w.dim() != 2 or int(x.shape[1]) != int(w.shape[0]):
            raise ValueError(f"linear: cannot multiply {tuple(x.shape)} by {tuple(w.shape)}")
        bias_c = _f32c(bias, "bias").contiguous() if bias is not None else None
        x = _tma_rows(x, pad_cols=True)   # the padded copy (if one was needed) is what the backward's dW reads
        out = _matmul_tc(x, w, False, c=_f32c(addend, "addend") if addend is not None else None, bias=bias_c, relu=relu)
        ctx.save_for_backward(x, w, *([out] if relu else []))
        ctx.has_addend, ctx.has_bias, ctx.relu = addend is not None, bias is not None, relu
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors[:2]
        g = _f32c(g, "grad")
        g_bias = None
        if ctx.has_bias and ctx.needs_input_grad[3]:
            g, g_bias = relu_bwd_colsum(g, ctx.saved_tensors[2] if ctx.relu else None)
        elif ctx.relu:
            g = relu_bwd(g, ctx.saved_tensors[2])
        g = _tma_rows(g, pad_cols=True)
        gx = _matmul_tc(g, w, True) if ctx.needs_input_grad[0] else None
        gw = _dw_tc(x, g) if ctx.needs_input_grad[1] else None
        return gx, gw, (g if ctx.has_addend else None), g_bias, None


def linear(x, w, addend=None, bias=None, act=None) -> torch.Tensor:
    """act(x @ w + addend + bias) - the dense node-feature transform of every conv layer (K8); ``act`` in
    (None, "relu").  Addend, bias and ReLU run in the GEMM epilogue."""
    return _Linear.apply(x, w, addend, bias, act)


def sage_layer_ok(x: torch.Tensor, w_neigh: torch.Tensor, w_self: torch.Tensor) -> bool:
    """Shapes the one-node SAGE layer (``sage_layer``) covers: everything on the hand-written tcgen05 kernels."""
    M, K, N = int(x.shape[0]), int(x.shape[1]), int(w_neigh.shape[1])
    return (_gemm_ok(x, w_neigh, w_self) and M >= 0 and K <= _TC_SLAB and N <= _TC_SLAB
            and tuple(w_self.shape) == tuple(w_neigh.shape))


class _SageLayer(torch.autograd.Function):
    """act(OP_j(x_j) @ w_neigh + x @ w_self + bias) as ONE autograd node.  x feeds both the aggregation and the root
    transform; as separate nodes autograd adds their two [N, F] gradients in an extra pass, here the root-weight
    dX GEMM accumulates onto the transposed gather's output in its epilogue, and the bias gradient comes out of
    the ReLU-backward pass."""

    @staticmethod
    def forward(ctx, x, w_neigh, w_self, bias, graph: GraphStructure, op_name: str, relu: bool):
        op = _lib.OPS[op_name]
        x = _f32c(x, "x")
        w_neigh, w_self = _f32c(w_neigh, "w_neigh"), _f32c(w_self, "w_self")
        bias_c = _f32c(bias, "bias").contiguous() if bias is not None else None
        N = int(w_neigh.shape[1])
        if any(ctx.needs_input_grad):
            # training: this layer's or a later layer's backward walks the source-major orientation; on a fresh edge
            # list its build runs on a side stream beside the forward pass (no-op once the structure has it)
            graph.prefetch_csc()
        agg, _ = gather_reduce_raw(x, graph.csr, op)
        hi, lo = _split_weight_pair(w_neigh, w_self, transpose=True)   # [agg | x] @ [w_neigh ; w_self] in one pass
        out = linear_tc2(agg, x, hi, lo, N, bias=bias_c, relu=relu)
        # backward formulation (see there): with an input gradient and N <= K the incoming gradient itself is gathered
        # and `agg` is not needed again
        ctx.gather_grad = bool(ctx.needs_input_grad[0]) and N <= int(x.shape[1])
        ctx.save_for_backward(x, agg if not ctx.gather_grad else x.new_empty(0), w_neigh, w_self, *([out] if relu else []))
        ctx.graph, ctx.op, ctx.relu, ctx.has_bias = graph, op, relu, bias is not None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, agg, w_neigh, w_self = ctx.saved_tensors[:4]
        out = ctx.saved_tensors[4] if ctx.relu else None
        g = _f32c(g, "grad")
        g_bias = None
        if ctx.has_bias and ctx.needs_input_grad[3]:
            g, g_bias = relu_bwd_colsum(g, out)
        elif ctx.relu:
            g = relu_bwd(g, out)
        if g.stride(1) != 1 or g.stride(0) % 4 or g.data_ptr() % 16:
            g = g.contiguous()
        K, N = int(x.shape[1]), int(g.shape[1])
        kw = {"src_scale": ctx.graph.csr.inv_deg} if ctx.op == _lib.OP_MEAN else {}
        if ctx.gather_grad:
            # The transposed aggregation commutes with lin_neigh: A^T (g Wn^T) = (A^T g) Wn^T.  Gathering g itself
            # (N <= K columns) gives q = A^T g, and then dx = [q | g] @ [Wn^T ; Ws^T] is ONE K-concatenated GEMM instead
            # of two GEMMs over g, and dWn = (A x)^T g = x^T q shares x with dWs = x^T g.  C4 step, same box:
            # 41.2 -> 40.2 ms.
            q, _ = gather_reduce_raw(g, ctx.graph.csc, _lib.OP_SUM, **kw)
            hi, lo = _split_weight_pair(w_neigh, w_self, transpose=False)
            gx = linear_tc2(q, g, hi, lo, K)
            if ctx.needs_input_grad[1] and ctx.needs_input_grad[2]:
                g_wn, g_ws = _dw_tc2(x, q, g)
            else:
                g_wn = _dw_tc(x, q) if ctx.needs_input_grad[1] else None
                g_ws = _dw_tc(x, g) if ctx.needs_input_grad[2] else None
            return gx, g_wn, g_ws, g_bias, None, None, None
        if ctx.needs_input_grad[1] and ctx.needs_input_grad[2]:
            g_wn, g_ws = _dw_tc_x2(agg, x, g)      # [agg | x]^T g: one pass over g when both inputs are narrow
        else:
            g_wn = _dw_tc(agg, g) if ctx.needs_input_grad[1] else None
            g_ws = _dw_tc(x, g) if ctx.needs_input_grad[2] else None
        gx = None
        if ctx.needs_input_grad[0]:
            hi, lo = _split_weight(w_neigh, transpose=False)
            d_agg = linear_tc(g, hi, lo, K)
            gx, _ = gather_reduce_raw(d_agg, ctx.graph.csc, _lib.OP_SUM, **kw)
            hi, lo = _split_weight(w_self, transpose=False)
            gx = linear_tc(g, hi, lo, K, c=gx)
        return gx, g_wn, g_ws, g_bias, None, None, None


def sage_layer(x, w_neigh, w_self, bias, graph: GraphStructure, op: str, relu: bool) -> torch.Tensor:
    if op not in ("mean", "sum"):
        raise ValueError("sage_layer fuses the linear aggregators (mean, sum) only")
    return _SageLayer.apply(x, w_neigh, w_self, bias, graph, op, relu)


class _LinearPair(torch.autograd.Function):
    """(x @ w_a, x @ w_b) as one autograd node: the second dX GEMM accumulates onto the first in its epilogue
    instead of autograd adding two [N, F] gradients in a separate pass."""

    @staticmethod
    def forward(ctx, x, w_a, w_b):
        x, w_a, w_b = _f32c(x, "x"), _f32c(w_a, "w_a"), _f32c(w_b, "w_b")
        N = int(w_a.shape[1])
        if 2 * N <= _TC_SLAB and N % 4 == 0:
            # x @ [w_a | w_b] as ONE GEMM: x (the large operand) is loaded and split once; the two results are column
            # blocks of one [M, 2N] buffer (the gather and the epilogues downstream take any leading dimension)
            hi, lo = _split_weight(torch.cat([w_a, w_b], dim=1), transpose=True)
            both = linear_tc(x, hi, lo, 2 * N)
            a, b = both[:, :N], both[:, N:]
        else:
            hi, lo = _split_weight(w_a, transpose=True)
            a = linear_tc(x, hi, lo, N)
            hi, lo = _split_weight(w_b, transpose=True)
            b = linear_tc(x, hi, lo, N)
        ctx.save_for_backward(x, w_a, w_b)
        return a, b

    @staticmethod
    @once_differentiable
    def backward(ctx, ga, gb):
        x, w_a, w_b = ctx.saved_tensors
        K = int(x.shape[1])

        def prep(g):
            g = _f32c(g, "grad")
            return g.contiguous() if (g.stride(1) != 1 or g.stride(0) % 4 or g.data_ptr() % 16) else g

        ga, gb = prep(ga), prep(gb)
        if ctx.needs_input_grad[1] and ctx.needs_input_grad[2]:
            g_wa, g_wb = _dw_tc2(x, ga, gb)                              # X^T [ga | gb]: X is loaded and split once
        else:
            g_wa = _dw_tc(x, ga) if ctx.needs_input_grad[1] else None
            g_wb = _dw_tc(x, gb) if ctx.needs_input_grad[2] else None
        gx = None
        if ctx.needs_input_grad[0]:
            hi, lo = _split_weight_pair(w_a, w_b, transpose=False)      # [ga | gb] @ [w_a^T ; w_b^T] in one pass
            gx = linear_tc2(ga, gb, hi, lo, K)
        return gx, g_wa, g_wb


def linear_pair(x, w_a, w_b):
    """(x @ w_a, x @ w_b) for two weights of the same shape; falls back to two ``linear`` calls off the fast path."""
    if sage_layer_ok(x, w_a, w_b):
        return _LinearPair.apply(x, w_a, w_b)
    return linear(x, w_a), linear(x, w_b)


class _SagePartitioned(torch.autograd.Function):
    """SAGEConv (mean / sum) on a 1-D node partition as ONE autograd node that schedules its own overlap.

    forward:  halo rows travel on the communication stream while the local-source edges are reduced (and, with
              ``reorder``, the root transform runs); the halo part is added on top and the two GEMMs run as one.
    backward: the gradient of the halo rows is produced FIRST and sent back; the weight-gradient GEMMs, the
              local-source transposed gather and the root dX GEMM run while it travels; the returned rows are summed
              into their owners by the deterministic segmented sum, accumulating onto the local result.
    ``reorder``: aggregate after lin_neigh (narrower rows travel): out = act(A (x Wn) + x Ws + b)."""

    @staticmethod
    def forward(ctx, x, w_neigh, w_self, bias, pg, op_name: str, relu: bool, reorder: bool):
        x = _f32c(x, "x")
        w_neigh, w_self = _f32c(w_neigh, "w_neigh"), _f32c(w_self, "w_self")
        bias_c = _f32c(bias, "bias").contiguous() if bias is not None else None
        g_l, g_h, inv = pg.split
        scale = inv if op_name == "mean" else None
        N = int(w_neigh.shape[1])
        dev = x.device
        cur, cs = torch.cuda.current_stream(dev), pg._comm_stream()
        act = _lib.ACT_RELU if relu else _lib.ACT_NONE
        if reorder:
            hi, lo = _split_weight(w_neigh, transpose=True)
            z = linear_tc(x, hi, lo, N)
            cs.wait_stream(cur)
            with torch.cuda.stream(cs):
                halo = pg.halo_rows_raw(z)
            z.record_stream(cs)
            hi, lo = _split_weight(w_self, transpose=True)
            root = linear_tc(x, hi, lo, N)
            part, _ = gather_reduce_raw(z, g_l.csr, _lib.OP_SUM, out_scale=scale, addend=root)
            cur.wait_stream(cs)
            halo.record_stream(cur)
            out, _ = gather_reduce_raw(halo, g_h.csr, _lib.OP_SUM, out_scale=scale, addend=part, bias=bias_c, act=act)
            ctx.save_for_backward(x, w_neigh, w_self, *([out] if relu else []))
        else:
            cs.wait_stream(cur)
            with torch.cuda.stream(cs):
                halo = pg.halo_rows_raw(x)
            x.record_stream(cs)
            part, _ = gather_reduce_raw(x, g_l.csr, _lib.OP_SUM, out_scale=scale)
            # the root transform does not need the halo: it runs while the rows are still travelling, and the
            # neighbour transform adds it in its epilogue (one extra [n, N] round trip buys ~the whole GEMM of overlap)
            hi, lo = _split_weight(w_self, transpose=True)
            root = linear_tc(x, hi, lo, N)
            cur.wait_stream(cs)
            halo.record_stream(cur)
            agg, _ = gather_reduce_raw(halo, g_h.csr, _lib.OP_SUM, out_scale=scale, addend=part)
            hi, lo = _split_weight(w_neigh, transpose=True)
            out = linear_tc(agg, hi, lo, N, c=root, bias=bias_c, relu=relu)
            ctx.save_for_backward(x, w_neigh, w_self, agg, *([out] if relu else []))
        ctx.pg, ctx.scale, ctx.relu, ctx.reorder, ctx.has_bias = pg, scale, relu, reorder, bias is not None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        pg, scale = ctx.pg, ctx.scale
        g_l, g_h, _ = pg.split
        saved = ctx.saved_tensors
        x, w_neigh, w_self = saved[:3]
        out = saved[-1] if ctx.relu else None
        g = _f32c(g, "grad")
        g_bias = None
        if ctx.has_bias and ctx.needs_input_grad[3]:
            g, g_bias = relu_bwd_colsum(g, out)
        elif ctx.relu:
            g = relu_bwd(g, out)
        if g.stride(1) != 1 or g.stride(0) % 4 or g.data_ptr() % 16:
            g = g.contiguous()
        dev = x.device
        cur, cs = torch.cuda.current_stream(dev), pg._comm_stream()
        K = int(x.shape[1])
        need_x = ctx.needs_input_grad[0]

        def send_back(grad):   # transposed gather over the halo-source edges FUSED with the exchange: every finished
            cs.wait_stream(cur)   # halo-row gradient is stored straight into its owner's window (communication stream)
            with torch.cuda.stream(cs):
                back = pg.halo_grad_fused(grad, tgt_scale=scale)
            grad.record_stream(cs)
            return back

        def land(back, acc):   # + per-owner segmented sum of the returned rows
            cur.wait_stream(cs)
            back.record_stream(cur)
            return pg.land_into(back, acc)

        if ctx.reorder or (need_x and int(g.shape[1]) <= K):
            # out = act(S A_l z + S A_h halo(z) + x Ws + b), z = x Wn:  dz = A^T (S g) needs the exchange even when x
            # itself needs no gradient (it feeds dWn).
            # The aggregate-first layer out = act((S A x) Wn + x Ws + b) has the SAME backward: A^T S (g Wn^T) =
            # (A^T S g) Wn^T and dWn = (S A x)^T g = x^T (A^T S g).  Taking it when the gradient is not wider than x means
            # the halo gradients leave at once (no g Wn^T GEMM in front of the push) and dx is one K-concatenated GEMM.
            back = send_back(g)
            g_ws = _dw_tc(x, g) if ctx.needs_input_grad[2] else None
            dz, _ = gather_reduce_raw(g, g_l.csc, _lib.OP_SUM, src_scale=scale)
            dz = land(back, dz)
            g_wn = _dw_tc(x, dz) if ctx.needs_input_grad[1] else None
            gx = None
            if need_x:
                hi, lo = _split_weight_pair(w_neigh, w_self, transpose=False)
                gx = linear_tc2(dz, g, hi, lo, K)
            return gx, g_wn, g_ws, g_bias, None, None, None, None
        agg = saved[3]
        gx = None
        back = None
        if need_x:
            hi, lo = _split_weight(w_neigh, transpose=False)
            d_agg = linear_tc(g, hi, lo, K)
            back = send_back(d_agg)
        if ctx.needs_input_grad[1] and ctx.needs_input_grad[2]:
            g_wn, g_ws = _dw_tc_x2(agg, x, g)      # [agg | x]^T g: one pass over g when both inputs are narrow
        else:
            g_wn = _dw_tc(agg, g) if ctx.needs_input_grad[1] else None
            g_ws = _dw_tc(x, g) if ctx.needs_input_grad[2] else None
        if need_x:
            gx, _ = gather_reduce_raw(d_agg, g_l.csc, _lib.OP_SUM, src_scale=scale)
            hi, lo = _split_weight(w_self, transpose=False)
            gx = linear_tc(g, hi, lo, K, c=gx)
            gx = land(back, gx)
        return gx, g_wn, g_ws, g_bias, None, None, None, None


class _AggregatePartitioned(torch.autograd.Function):
    """out = act(out_scale * sum_j src_scale_j * x_j + bias) over a 1-D node partition as ONE autograd node (sum /
    mean / GCN-normalised sum).
    forward:  halo rows are pushed into this rank's window; ONE gather pass over the rank's edges reads owned
              sources from ``x`` and halo sources from the window (split-source kernel: no [local | halo] copy, no
              second pass over the output).
    backward: the transposed gather over the halo-source edges runs on the communication stream and stores every
              finished gradient row straight into its owner's window (fused gather + exchange) while the compute
              stream reduces the local-source edges; the received rows are summed into their owners' rows by the
              deterministic segmented sum.
    ``scales`` = (src_ext [n_local + n_halo] | None, dst [n_local] | None)."""

    @staticmethod
    def forward(ctx, x, bias, pg, scales, relu: bool):
        x = _f32c(x, "x")
        s_ext, s_dst = scales
        bias_c = _f32c(bias, "bias").contiguous() if bias is not None else None
        dev = x.device
        cur, cs = torch.cuda.current_stream(dev), pg._comm_stream()
        cs.wait_stream(cur)
        with torch.cuda.stream(cs):
            halo = pg.halo_rows_raw(x)
        x.record_stream(cs)
        cur.wait_stream(cs)
        halo.record_stream(cur)
        out, _ = gather_reduce_raw(x, pg.graph.csr, _lib.OP_SUM, x2=halo, n_split_src=pg.n_local, src_scale=s_ext,
                                   out_scale=s_dst, bias=bias_c, act=_lib.ACT_RELU if relu else _lib.ACT_NONE)
        ctx.pg, ctx.scales, ctx.relu, ctx.has_bias = pg, scales, relu, bias is not None
        ctx.save_for_backward(*([out] if relu else []))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        pg = ctx.pg
        g_l = pg.split[0]
        s_ext, s_dst = ctx.scales
        s_loc = s_ext[:pg.n_local] if s_ext is not None else None
        s_halo = s_ext[pg.n_local:] if s_ext is not None else None
        g = _f32c(g, "grad")
        out = ctx.saved_tensors[0] if ctx.relu else None
        g_bias = None
        if ctx.has_bias and ctx.needs_input_grad[1]:
            g, g_bias = relu_bwd_colsum(g, out)
        elif ctx.relu:
            g = relu_bwd(g, out)
        gx = None
        if ctx.needs_input_grad[0]:
            dev = g.device
            cur, cs = torch.cuda.current_stream(dev), pg._comm_stream()
            cs.wait_stream(cur)
            with torch.cuda.stream(cs):   # transposed halo part + exchange in one kernel, on the communication stream
                back = pg.halo_grad_fused(g, tgt_scale=s_dst, halo_scale=s_halo)
            g.record_stream(cs)
            gx, _ = gather_reduce_raw(g, g_l.csc, _lib.OP_SUM, src_scale=s_dst, out_scale=s_loc)
            cur.wait_stream(cs)
            back.record_stream(cur)
            gx = pg.land_into(back, gx)
        return gx, g_bias, None, None, None


def aggregate_partitioned(x, pg, op: str = "sum", bias=None, relu: bool = False) -> torch.Tensor:
    """Linear neighbourhood aggregation of this rank's rows on a partitioned graph (``keras_geometric_b200.dist``):
    ``op`` in "sum", "mean", "gcn" (symmetric normalisation with the total in-degree, utils/main.py:20-33)."""
    if op == "sum":
        scales = (None, None)
    elif op == "mean":
        scales = (None, pg.split[2])
    elif op == "gcn":
        dis, dis_ext = pg.gcn_dis_split()
        scales = (dis_ext, dis)
    else:
        raise ValueError(f"aggregate_partitioned: unsupported aggregation {op!r}")
    return _AggregatePartitioned.apply(x, bias, pg, scales, relu)


def sage_partitioned(x, w_neigh, w_self, bias, pg, op: str, relu: bool, reorder: bool) -> torch.Tensor:
    return _SagePartitioned.apply(x, w_neigh, w_self, bias, pg, op, relu, reorder)
