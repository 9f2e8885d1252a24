"""Host-side operators of the D-MPNN hot path: thin tensor -> C-ABI wrappers and the hand-written
autograd around them. PyTorch is plumbing here (device memory, streams, autograd graph); all
arithmetic happens in ``libnotorch_b200.so``. CPU tensors raise — there is no fallback.

Kernel legend: K0 edge_init, K1 edge->atom, K2 fused layer forward, K3 read-out,
K4 dgrad/wgrad, K5 gather backward (segmented sum by src), K6 layer backward epilogue, K-l CSR.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import torch
from torch import Tensor

from . import _lib
from ._lib import GEMM_BF16, GEMM_FP32, GEMM_TF32, GEMM_TF32X3, NT_BF16, NT_F32

__all__ = [
    "GraphCSR", "SegmentCSR", "build_segment_csr", "graph_csr", "segment_csr_for", "seg_reduce", "gather_add",
    "edge_init", "edge_to_atom", "readout", "layer", "set_gemm_mode", "get_gemm_mode", "collate_packed",
    "dropout_mask", "set_index_validation", "KernelTimer", "embed_edge_init", "embed_edge_init_supported", "carry_graph_caches",
    "set_save_messages", "set_dispatch",
]

ACT_CODES = {
    torch.nn.Identity: (_lib.ACT_IDENTITY, lambda m: 0.0),
    torch.nn.ReLU: (_lib.ACT_RELU, lambda m: 0.0),
    torch.nn.LeakyReLU: (_lib.ACT_LEAKY_RELU, lambda m: float(m.negative_slope)),
    torch.nn.ELU: (_lib.ACT_ELU, lambda m: float(m.alpha)),
    torch.nn.SiLU: (_lib.ACT_SILU, lambda m: 0.0),
    torch.nn.GELU: (_lib.ACT_GELU, lambda m: 0.0),
    torch.nn.Tanh: (_lib.ACT_TANH, lambda m: 0.0),
}

_GEMM_MODES = {"tf32x3": GEMM_TF32X3, "fp32": GEMM_FP32, "tf32": GEMM_TF32, "bf16": GEMM_BF16}
_gemm_mode = _GEMM_MODES[os.environ.get("NOTORCH_B200_GEMM", "tf32x3").lower()]
_validate_mode = os.environ.get("NOTORCH_B200_VALIDATE", "sync").lower()  # "sync" | "deferred" | "off"
_side_wprep = os.environ.get("NOTORCH_B200_SIDE_WPREP", "1") != "0"  # weight images (forward + transposed) filled on a side stream under K1
_parallel_csr = os.environ.get("NOTORCH_B200_PARALLEL_CSR", "1") != "0"  # by_src / by_dst / by_rev built on three streams
_fuse_k5_k6 = os.environ.get("NOTORCH_B200_FUSE_K5K6", "1") != "0"  # backward epilogue sums the outgoing-edge gradients itself (no K5 launch)


_save_messages = os.environ.get("NOTORCH_B200_SAVE_M", "1") != "0"


def set_save_messages(flag: bool) -> None:
    """``True`` (default): K2 also writes the message tensor ``m_l`` and the weight gradient streams it as dense tiles (fastest).
    ``False``: nothing but ``h_l`` and the small ``n_l`` [V, d] is kept per depth — the weight gradient re-gathers
    ``m = n[src] - act(h)[rev]`` itself (``wgrad_tc.cu``), halving the activation memory saved for backward (SURVEY.md §7:
    17 GB at BASELINE configs[2]) for a slower K4b."""
    global _save_messages
    _save_messages = bool(flag)


def set_gemm_mode(mode: str) -> None:
    """``"tf32x3"`` (tcgen05, error-compensated; default), ``"fp32"`` (FFMA), ``"tf32"`` (single pass) or ``"bf16"`` (bf16 operands
    for W_h and the message tensor, one tcgen05 kind::f16 pass with fp32 accumulation - BASELINE configs[4]; NOT within the fp32 bound)."""
    global _gemm_mode
    _gemm_mode = _GEMM_MODES[mode.lower()]


def get_gemm_mode() -> str:
    return {v: k for k, v in _GEMM_MODES.items()}[_gemm_mode]


_dispatch = os.environ.get("NOTORCH_B200_DISPATCH", "function").lower()  # "function" (autograd.Function) | "ops" (torch.ops.notorch_b200.*)


def set_dispatch(mode: str) -> None:
    """``"function"`` (default): the modules call the kernels through ``torch.autograd.Function`` (cheapest eager call).
    ``"ops"``: through the registered dispatcher ops ``torch.ops.notorch_b200.*`` (``torch_ops.py``) — what tracing needs; the
    modules switch to it by themselves while ``torch.compile`` traces them or when they see fake tensors."""
    global _dispatch
    assert mode in ("function", "ops")
    _dispatch = mode


def _via_ops(*tensors) -> bool:
    if _dispatch == "ops" or torch.compiler.is_compiling():
        return True
    from torch._subclasses.fake_tensor import FakeTensor

    return any(isinstance(t, FakeTensor) for t in tensors)


def _torch_ops():
    from . import torch_ops  # registers torch.ops.notorch_b200.* on first use

    return torch_ops


def set_index_validation(mode: str) -> None:
    """When the out-of-range verdict of the CSR build is read back: ``"sync"`` (immediately; one
    device sync per new batch), ``"deferred"`` (checked on the next batch) or ``"off"``."""
    global _validate_mode
    assert mode in ("sync", "deferred", "off")
    _validate_mode = mode


def act_code(act: torch.nn.Module | None) -> tuple[int, float]:
    """Map an activation *module instance* to the closed set compiled into the kernels."""
    if act is None:
        return _lib.ACT_IDENTITY, 0.0
    for cls, (code, param) in ACT_CODES.items():
        if type(act) is cls:
            if cls is torch.nn.GELU and getattr(act, "approximate", "none") != "none":
                break
            return code, param(act)
    raise NotImplementedError(
        f"notorch_b200: activation {type(act).__name__} is not compiled into the kernels "
        f"(supported: {', '.join(c.__name__ for c in ACT_CODES)}); there is no fallback path")


# ------------------------------------------------------------------------------------------------
# plumbing
# ------------------------------------------------------------------------------------------------

def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _require(t: Tensor, name: str, dtype: torch.dtype | None = None, dim: int | None = None) -> Tensor:
    if not isinstance(t, Tensor):
        raise TypeError(f"notorch_b200: {name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"notorch_b200: {name} is on {t.device}; the hot path runs on CUDA only (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"notorch_b200: {name} must be {dtype}, got {t.dtype}")
    if dim is not None and t.dim() != dim:
        raise RuntimeError(f"notorch_b200: {name} must be {dim}-D, got shape {tuple(t.shape)}")
    return t.contiguous()


def _require_float(t: Tensor, name: str) -> Tensor:
    t = _require(t, name, dim=2)
    if t.dtype != torch.float32:
        raise RuntimeError(f"notorch_b200: {name} has dtype {t.dtype}; this build implements float32 only")
    return t


class KernelTimer:
    """Optional per-kernel CUDA-event timing (used by bench.py for the roofline numbers): while
    active, every C-ABI call is bracketed by two events on the launching stream."""

    def __init__(self):
        self.records: list[tuple[str, torch.cuda.Event, torch.cuda.Event]] = []

    def __enter__(self):
        global _timer
        _timer = self
        return self

    def __exit__(self, *exc):
        global _timer
        _timer = None

    def summary(self) -> dict[str, dict[str, float]]:
        torch.cuda.synchronize()
        out: dict[str, dict[str, float]] = {}
        for tag, s, e in self.records:
            rec = out.setdefault(tag, {"launches": 0, "total_ms": 0.0})
            rec["launches"] += 1
            rec["total_ms"] += s.elapsed_time(e)
        for rec in out.values():
            rec["avg_ms"] = rec["total_ms"] / rec["launches"]
        return out


_timer: KernelTimer | None = None


def _run(tag: str, fn, *args) -> None:
    """Call one C-ABI entry point; raise on a non-zero status."""
    t = _timer
    if t is None:
        _lib.check(fn(*args), tag)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = fn(*args)
    e.record()
    t.records.append((tag, s, e))
    _lib.check(rc, tag)


def _workspace(device: torch.device, nbytes: int, slot: int = 0) -> Tensor:
    """Scratch bytes for ONE kernel call, drawn from the caching allocator on the calling stream: two streams or threads running
    backward on one device never share a buffer (a process-wide scratch tensor would let their split-K partials overwrite each
    other), and after warm-up the allocation is a free-list hit. ``slot`` is kept for call-site readability only."""
    del slot
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# CSR bundles (K-l)
# ------------------------------------------------------------------------------------------------

@dataclass
class SegmentCSR:
    """Stable CSR of items grouped by key: ``perm[rowptr[s]:rowptr[s+1]]`` = item ids with key ``s``
    in ascending order; ``keys32`` = the int32 copy of the key vector."""

    rowptr: Tensor  # [S+1] int32
    perm: Tensor | None  # [n] int32 (None = identity / contiguous segments)
    keys32: Tensor  # [n] int32
    num_segments: int
    status: Tensor | None = None  # [1] int32 device flag, bit 0 = key out of range
    ell: Tensor | None = None  # [S, 4] int32: first four item ids of every segment (-1 padded), see nt_csr_to_ell

    def to(self, device, non_blocking: bool = False) -> "SegmentCSR":
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return SegmentCSR(mv(self.rowptr), mv(self.perm), mv(self.keys32), self.num_segments, mv(self.status), mv(self.ell))


def _ell_of(csr: "SegmentCSR") -> Tensor | None:
    """ELL copy of a permuted CSR (built once per batch, cached on the CSR object); contiguous segments keep the plain kernel."""
    if csr.perm is None or csr.num_segments == 0:
        return None
    if csr.ell is None:
        ell = torch.empty((csr.num_segments, 4), dtype=torch.int32, device=csr.rowptr.device)
        _run("ell:nt_csr_to_ell", _lib.lib().nt_csr_to_ell, _p(csr.rowptr), _p(csr.perm), csr.num_segments, _p(ell), _stream())
        csr.ell = ell
    return csr.ell


_STATUS_SLOTS = 256
_status_ring: Tensor | None = None  # pinned int32 ring: no per-batch pinned allocation (cudaHostAlloc synchronises)
_status_events: list = [None] * _STATUS_SLOTS
_status_what: list = [""] * _STATUS_SLOTS
_status_head = 0  # next slot to write
_status_tail = 0  # oldest slot not yet examined


def _drain_status(block_slot: int | None = None) -> None:
    global _status_tail
    while _status_tail < _status_head:
        slot = _status_tail % _STATUS_SLOTS
        ev = _status_events[slot]
        if block_slot is not None and slot == block_slot:
            ev.synchronize()
        elif not ev.query():
            return
        _status_tail += 1
        if int(_status_ring[slot]) != 0:
            raise IndexError(f"notorch_b200: {_status_what[slot]}: index out of range (detected on a later batch)")


def _check_status(status: Tensor, what: str) -> None:
    global _status_ring, _status_head
    if _validate_mode == "off":
        return
    if _validate_mode == "sync":
        if int(status.item()) != 0:
            raise IndexError(f"notorch_b200: {what}: index out of range")
        return
    # deferred: copy the verdict into a pinned ring slot and look at the verdicts that have landed by now
    if _status_ring is None:
        _status_ring = torch.zeros(_STATUS_SLOTS, dtype=torch.int32).pin_memory()
    slot = _status_head % _STATUS_SLOTS
    if _status_head - _status_tail >= _STATUS_SLOTS:
        _drain_status(block_slot=slot)  # ring full: wait for the oldest verdict
    _status_ring[slot:slot + 1].copy_(status, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _status_events[slot], _status_what[slot] = ev, what
    _status_head += 1
    _drain_status()


def _csr_outputs(n: int, num_segments: int, dev: torch.device) -> tuple[Tensor, Tensor, Tensor]:
    return (torch.empty(num_segments + 1, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
            torch.empty(n, dtype=torch.int32, device=dev))


def _launch_build_csr(keys: Tensor, num_segments: int, outs: tuple[Tensor, Tensor, Tensor], status: Tensor, slot: int) -> None:
    rowptr, perm, keys32 = outs
    n = keys.numel()
    L = _lib.lib()
    ws = _workspace(keys.device, L.nt_build_csr_workspace_bytes(n, num_segments), slot=slot)
    _run("csr:nt_build_csr", L.nt_build_csr, _p(keys), n, num_segments, _p(keys32), _p(rowptr), _p(perm), _p(status), _p(ws), ws.numel(), _stream())


def build_segment_csr(keys: Tensor, num_segments: int, what: str = "index", status: Tensor | None = None,
                      validate: bool = True) -> SegmentCSR:
    if _via_ops(keys):
        _torch_ops()
        rowptr, perm, keys32, st = torch.ops.notorch_b200.build_csr(keys, num_segments)
        if validate and not torch.compiler.is_compiling() and type(st) is Tensor:  # a real verdict exists only outside tracing
            _check_status(st, what)
        return SegmentCSR(rowptr, perm, keys32, num_segments, st)
    keys = _require(keys, what, torch.int64, 1)
    dev = keys.device
    with torch.cuda.device(dev):
        outs = _csr_outputs(keys.numel(), num_segments, dev)
        if status is None:
            status = torch.zeros(1, dtype=torch.int32, device=dev)
        _launch_build_csr(keys, num_segments, outs, status, slot=0)
    if validate:
        _check_status(status, what)
    return SegmentCSR(outs[0], outs[1], outs[2], num_segments, status)


_side_streams: dict[tuple[int, int], torch.cuda.Stream] = {}


def _side_stream(device: torch.device, which: int) -> torch.cuda.Stream:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _side_streams.get((idx, which))
    if st is None:
        st = _side_streams[(idx, which)] = torch.cuda.Stream(device=idx)
    return st


@dataclass
class GraphCSR:
    """Everything the kernels need about one (batched) graph's topology, int32, built once per batch."""

    V: int
    E: int
    by_dst: SegmentCSR  # K1: incoming edges of each atom (keys32 = dst)
    by_src: SegmentCSR  # K5: outgoing edges of each atom (keys32 = src)
    by_rev: SegmentCSR  # K6: inverse map of the (arbitrary) gather index rev (keys32 = rev)
    key: tuple = field(default_factory=tuple)
    # the index tensors the bundle was built from: holding them keeps their storage alive, so the (address, shape, version) key
    # above can never match a NEW tensor that the caching allocator placed at a recycled address
    source: tuple = field(default_factory=tuple, repr=False)

    def to(self, device, non_blocking: bool = False) -> "GraphCSR":
        """The bundle on another device (int32 copies; nothing is rebuilt). ``key`` / ``source`` are set by the caller."""
        return GraphCSR(self.V, self.E, self.by_dst.to(device, non_blocking), self.by_src.to(device, non_blocking),
                        self.by_rev.to(device, non_blocking))

    @property
    def src(self) -> Tensor:
        return self.by_src.keys32

    @property
    def dst(self) -> Tensor:
        return self.by_dst.keys32

    @property
    def rev(self) -> Tensor:
        return self.by_rev.keys32


def _tensor_key(t: Tensor) -> tuple:
    if type(t) is not Tensor and type(t) is not torch.nn.Parameter:  # fake / traced tensors have no address: identity of the object
        return (id(t), tuple(t.shape), 0, t.device.index)
    return (t.data_ptr(), tuple(t.shape), t._version, t.device.index)


def build_graph_csr(edge_index: Tensor, rev_index: Tensor, num_nodes: int) -> GraphCSR:
    if _via_ops(edge_index, rev_index):
        E = edge_index.shape[1]
        segs = [build_segment_csr(k, S, "edge_index / rev_index") for k, S in ((edge_index[0], num_nodes), (edge_index[1], num_nodes), (rev_index, E))]
        return GraphCSR(num_nodes, E, segs[1], segs[0], segs[2])
    edge_index = _require(edge_index, "edge_index", torch.int64, 2)
    rev_index = _require(rev_index, "rev_index", torch.int64, 1)
    if edge_index.shape[0] != 2:
        raise RuntimeError(f"notorch_b200: edge_index must have shape [2, E], got {tuple(edge_index.shape)}")
    E = edge_index.shape[1]
    if rev_index.shape[0] != E:
        raise RuntimeError(f"notorch_b200: rev_index has {rev_index.shape[0]} entries for {E} edges")
    dev = edge_index.device
    keys = (edge_index[0], edge_index[1], rev_index)
    sizes = (num_nodes, num_nodes, E)
    with torch.cuda.device(dev):
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        outs = [_csr_outputs(E, S, dev) for S in sizes]  # allocated on the caller's stream, whichever stream fills them
        if _parallel_csr and _timer is None and E > 0:
            # the three CSRs are independent chains of small, latency-bound kernels: build two of them on side streams
            main = torch.cuda.current_stream()
            fork = torch.cuda.Event()
            fork.record(main)
            _launch_build_csr(keys[0], sizes[0], outs[0], status, slot=0)
            for which in (1, 2):
                side = _side_stream(dev, which)
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    _launch_build_csr(keys[which], sizes[which], outs[which], status, slot=4 + which)
                    done = torch.cuda.Event()
                    done.record(side)
                main.wait_event(done)
        else:
            for which in range(3):
                _launch_build_csr(keys[which], sizes[which], outs[which], status, slot=0)
    by_src, by_dst, by_rev = (SegmentCSR(o[0], o[1], o[2], S, status) for o, S in zip(outs, sizes))
    _check_status(status, "edge_index / rev_index")
    return GraphCSR(num_nodes, E, by_dst, by_src, by_rev)


def mol_edge_csr(G) -> SegmentCSR:
    """The molecules' contiguous edge ranges of a device-collated batch (``BatchedGraph.from_packed``) as a segment CSR without a
    permutation; ``keys32`` = the molecule of every edge (``batch_edge_index`` as int32). Cached on the graph object."""
    csr = getattr(G, "_nt_mol_edge_csr", None)
    if csr is None:
        csr = SegmentCSR(G._nt_mol_edge_ptr, None, G.batch_edge_index.to(torch.int32), len(G))
        G._nt_mol_edge_csr = csr
    return csr


def peek_feats(G, name: str):
    """``G.node_feats`` / ``G.edge_feats`` without computing a pending embedding (``Graph.peek``); plain attribute on foreign graphs."""
    peek = getattr(G, "peek", None)
    return peek(name) if peek is not None else getattr(G, name)


def graph_csr(G, num_nodes: int | None = None) -> GraphCSR:
    """CSR bundle of a ``Graph``/``BatchedGraph``, cached on the object (``Graph.update`` makes
    shallow copies, so the cache rides along GraphEmbedding -> ChempropBlock -> Aggregation)."""
    V = len(peek_feats(G, "node_feats")) if num_nodes is None else num_nodes
    if torch.compiler.is_compiling():  # under tracing the graph object is an input of the trace: no address-keyed cache, three build_csr ops
        return build_graph_csr(G.edge_index, G.rev_index, V)
    key = (_tensor_key(G.edge_index), _tensor_key(G.rev_index), V)
    cached = getattr(G, "_nt_csr", None)
    if cached is not None and cached.key == key:
        return cached
    csr = build_graph_csr(G.edge_index, G.rev_index, V)
    csr.key, csr.source = key, (G.edge_index, G.rev_index)
    try:
        G._nt_csr = csr
    except AttributeError:  # pragma: no cover - slotted foreign object
        pass
    return csr


def carry_graph_caches(G, caches: dict, before: dict) -> None:
    """``Graph.to(cuda device)``: re-attach the per-batch preprocessing that was cached on ``G`` (CSR bundle, molecule row
    pointers, read-out CSR) after its index tensors moved — copied to the new device when the device changed, re-keyed to the
    new tensors either way. ``before`` maps field name -> the tensor it held before the move."""
    dev = G.edge_index.device
    csr = caches.get("_nt_csr")
    if csr is not None and csr.key == (_tensor_key(before["edge_index"]), _tensor_key(before["rev_index"]), csr.V):
        moved = csr if csr.by_dst.rowptr.device == dev else csr.to(dev, non_blocking=True)
        if moved is not csr and getattr(csr, "_atom_csr", None) is not None:
            pass  # the atom neighbour lists are derived lazily from the moved bundle
        moved.key = (_tensor_key(G.edge_index), _tensor_key(G.rev_index), csr.V)
        moved.source = (G.edge_index, G.rev_index)
        G._nt_csr = moved
    for attr in ("_nt_mol_ptr", "_nt_mol_edge_ptr"):
        ptr = caches.get(attr)
        if ptr is not None:
            setattr(G, attr, ptr.to(dev, non_blocking=True))
    seg = caches.get("_nt_seg_csr")
    if seg:
        kept = {}
        for attr, (key, c, t_old) in seg.items():
            if attr in before and before[attr] is t_old:
                t_new = getattr(G, attr)
                kept[attr] = ((_tensor_key(t_new), key[1]), c if c.rowptr.device == dev else c.to(dev, non_blocking=True), t_new)
        if kept:
            G._nt_seg_csr = kept


_layer_csr_cache: list[tuple[tuple, GraphCSR]] = []


def graph_csr_from_tensors(edge_index: Tensor, rev_index: Tensor, num_nodes: int) -> GraphCSR:
    """Small LRU for the stand-alone ``ChempropLayer.forward(edge_feats, node_feats, edge_index, rev_index)``."""
    key = (_tensor_key(edge_index), _tensor_key(rev_index), num_nodes)
    for k, c in _layer_csr_cache:
        if k == key:
            return c
    csr = build_graph_csr(edge_index, rev_index, num_nodes)
    csr.key, csr.source = key, (edge_index, rev_index)  # the entry owns its key tensors: their addresses cannot be recycled while it lives
    _layer_csr_cache.insert(0, (key, csr))
    del _layer_csr_cache[4:]
    return csr


def segment_csr_for(G, attr: str, num_segments: int) -> SegmentCSR:
    """CSR of ``G.<attr>`` (e.g. ``batch_node_index``), cached on the graph object."""
    t = getattr(G, attr)
    if torch.compiler.is_compiling():
        return build_segment_csr(t, num_segments, attr)
    key = (_tensor_key(t), num_segments)
    cache = getattr(G, "_nt_seg_csr", None)
    if cache is None:
        cache = {}
        try:
            G._nt_seg_csr = cache
        except AttributeError:  # pragma: no cover
            pass
    hit = cache.get(attr)
    if hit is not None and hit[0] == key:
        return hit[1]
    csr = build_segment_csr(t, num_segments, attr)
    cache[attr] = (key, csr, t)  # `t` is kept so that its address cannot be handed to another index tensor while the entry lives
    return csr


# ------------------------------------------------------------------------------------------------
# raw kernels
# ------------------------------------------------------------------------------------------------

def _seg_reduce_raw(x: Tensor, csr: SegmentCSR, act: int = 0, act_param: float = 0.0, mean: bool = False, scale: float = 1.0,
                    tag: str = "K1") -> Tensor:
    d = x.shape[1]
    out = torch.empty((csr.num_segments, d), dtype=x.dtype, device=x.device)
    ell = _ell_of(csr)
    if ell is not None:
        _run(f"{tag}:nt_seg_reduce", _lib.lib().nt_seg_reduce_ell, _p(x), d, _p(csr.rowptr), _p(csr.perm), _p(ell), csr.num_segments, act, act_param,
             int(mean), scale, None, None, _p(out), NT_F32, _stream())
        return out
    _run(f"{tag}:nt_seg_reduce", _lib.lib().nt_seg_reduce, _p(x), d, _p(csr.rowptr), _p(csr.perm), csr.num_segments, act, act_param, int(mean),
         scale, _p(out), NT_F32, _stream())
    return out


def _seg_extreme_raw(x: Tensor, csr: SegmentCSR, act: int, act_param: float, is_min: bool, tag: str = "K1x") -> tuple[Tensor, Tensor]:
    """max / min of act(x) over every segment + the int32 argument rows (first extreme wins; empty segment: 0 / -1)."""
    d = x.shape[1]
    out = torch.empty((csr.num_segments, d), dtype=x.dtype, device=x.device)
    arg = torch.empty((csr.num_segments, d), dtype=torch.int32, device=x.device)
    _run(f"{tag}:nt_seg_extreme", _lib.lib().nt_seg_extreme, _p(x), d, _p(csr.rowptr), _p(csr.perm), csr.num_segments, act, act_param, int(is_min),
         _p(out), _p(arg), NT_F32, _stream())
    return out, arg


def _gather_add_raw(base: Tensor | None, x: Tensor, idx32: Tensor, mean_rowptr: Tensor | None, scale: float = 1.0,
                    tag: str = "K0") -> Tensor:
    n, d = idx32.numel(), x.shape[1]
    out = torch.empty((n, d), dtype=x.dtype, device=x.device)
    _run(f"{tag}:nt_gather_add", _lib.lib().nt_gather_add, _p(base), _p(x), _p(idx32), _p(mean_rowptr), n, d, scale, _p(out), NT_F32, _stream())
    return out


def _weight_image_alloc(W: Tensor) -> Tensor | None:
    if _gemm_mode == GEMM_FP32 or W.shape[0] % 4 != 0:
        return None
    return torch.empty(_lib.lib().nt_weight_image_bytes(W.shape[0]), dtype=torch.uint8, device=W.device)


def _weight_image_fill(W: Tensor, transpose: bool, img: Tensor | None) -> None:
    if img is not None:
        _run("Wprep:nt_weight_prepare", _lib.lib().nt_weight_prepare, _p(W), W.shape[0], int(transpose), _p(img),
             NT_BF16 if _gemm_mode == GEMM_BF16 else NT_F32, _stream())


def _weight_image(W: Tensor, transpose: bool) -> Tensor | None:
    img = _weight_image_alloc(W)
    _weight_image_fill(W, transpose, img)
    return img


# ------------------------------------------------------------------------------------------------
# autograd
# ------------------------------------------------------------------------------------------------

class _SegReduce(torch.autograd.Function):
    """K1 (no activation) / K3: out[s] = sum|mean of x over segment s; backward is a row gather."""

    @staticmethod
    def forward(ctx, x: Tensor, csr: SegmentCSR, mean: bool, scale: float, tag: str):
        x = _require_float(x, "x")
        if x.shape[0] != csr.keys32.numel():
            raise RuntimeError(f"notorch_b200: {x.shape[0]} rows but the index has {csr.keys32.numel()} entries")
        with torch.cuda.device(x.device):
            out = _seg_reduce_raw(x, csr, 0, 0.0, mean, scale, tag=tag)
        ctx.csr, ctx.mean, ctx.scale, ctx.tag = csr, mean, scale, tag
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        csr = ctx.csr
        g = g.contiguous()
        with torch.cuda.device(g.device):
            gx = _gather_add_raw(None, g, csr.keys32, csr.rowptr if ctx.mean else None, ctx.scale, tag=ctx.tag + "bwd")
        return gx, None, None, None, None


class _SegExtreme(torch.autograd.Function):
    """``scatter(x, index, reduce="max" | "min")`` (torch_scatter arg-reductions; chemprop.py:86 with reduce in {max, min}):
    the gradient flows to the first row attaining the extreme, empty segments give 0."""

    @staticmethod
    def forward(ctx, x: Tensor, csr: SegmentCSR, is_min: bool, tag: str):
        x = _require_float(x, "x")
        if x.shape[0] != csr.keys32.numel():
            raise RuntimeError(f"notorch_b200: {x.shape[0]} rows but the index has {csr.keys32.numel()} entries")
        with torch.cuda.device(x.device):
            out, arg = _seg_extreme_raw(x, csr, _lib.ACT_IDENTITY, 0.0, is_min, tag=tag)
        ctx.save_for_backward(arg)
        ctx.csr, ctx.n, ctx.tag = csr, x.shape[0], tag
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        (arg,) = ctx.saved_tensors
        return _seg_extreme_backward_raw(g.contiguous(), arg, ctx.csr.keys32, ctx.n, ctx.tag), None, None, None


def _seg_extreme_backward_raw(g: Tensor, arg: Tensor, keys32: Tensor, n: int, tag: str = "K1x") -> Tensor:
    d = g.shape[1]
    with torch.cuda.device(g.device):
        gx = torch.empty((n, d), dtype=g.dtype, device=g.device)
        _run(f"{tag}bwd:nt_seg_max_backward", _lib.lib().nt_seg_max_backward, _p(g), _p(arg), _p(keys32), n, d, _p(gx), NT_F32, _stream())
    return gx


class _GatherAdd(torch.autograd.Function):
    """K0: out[i] = base[i] + x[idx[i]]  (chemprop.py:83); backward of the gather is K5."""

    @staticmethod
    def forward(ctx, base: Tensor, x: Tensor, csr: SegmentCSR):
        base, x = _require_float(base, "edge_feats"), _require_float(x, "node_feats")
        if base.shape[0] != csr.keys32.numel() or x.shape[0] != csr.num_segments or base.shape[1] != x.shape[1]:
            raise RuntimeError(f"notorch_b200: edge_init shape mismatch: node_feats {tuple(x.shape)}, edge_feats {tuple(base.shape)}, "
                               f"E={csr.keys32.numel()}, V={csr.num_segments}")
        with torch.cuda.device(x.device):
            out = _gather_add_raw(base, x, csr.keys32, None)
        ctx.csr = csr
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        g = g.contiguous()
        gx = None
        if ctx.needs_input_grad[1]:
            with torch.cuda.device(g.device):
                gx = _seg_reduce_raw(g, ctx.csr, tag="K5")
        return (g if ctx.needs_input_grad[0] else None), gx, None


def _embed_edge_init_raw(table_v: Tensor, table_e: Tensor, node_types: Tensor, edge_types: Tensor, src32: Tensor, V: int) -> Tensor:
    E, d = edge_types.shape[0], table_v.shape[1]
    with torch.cuda.device(table_v.device):
        h0 = torch.empty((E, d), dtype=table_v.dtype, device=table_v.device)
        status = torch.zeros(1, dtype=torch.int32, device=table_v.device)
        _run("K0e:nt_embed_edge_init", _lib.lib().nt_embed_edge_init, _p(table_v), table_v.shape[0], _p(table_e), table_e.shape[0],
             _p(node_types), node_types.shape[1], _p(edge_types), edge_types.shape[1], _p(src32), E, V, d, _p(h0), _p(status), NT_F32, _stream())
    _check_status(status, "GraphEmbedding type indices")
    return h0


def _embed_edge_init_backward_raw(g: Tensor, node_types: Tensor, edge_types: Tensor, src32: Tensor, V: int, Tv: int, Te: int) -> tuple[Tensor, Tensor]:
    E, d = g.shape
    L = _lib.lib()
    with torch.cuda.device(g.device):
        gv = torch.empty((Tv, d), dtype=g.dtype, device=g.device)
        ge = torch.empty((Te, d), dtype=g.dtype, device=g.device)
        ws = _workspace(g.device, L.nt_embed_edge_init_backward_workspace_bytes(E, Tv, Te, d))
        _run("K0ebwd:nt_embed_edge_init_backward", L.nt_embed_edge_init_backward, _p(g), _p(node_types), node_types.shape[1], _p(edge_types),
             edge_types.shape[1], _p(src32), E, V, Tv, Te, d, _p(gv), _p(ge), _p(ws), ws.numel(), NT_F32, _stream())
    return gv, ge


class _EmbedEdgeInit(torch.autograd.Function):
    """GraphEmbedding fused into K0: h0 = bag(Tv, node_types)[src] + bag(Te, edge_types)
    (embed.py:20-24 + chemprop.py:83) in one kernel, both tables in shared memory; backward = one pass over g_{h0}."""

    @staticmethod
    def forward(ctx, table_v: Tensor, table_e: Tensor, node_types: Tensor, edge_types: Tensor, csr: "GraphCSR"):
        table_v, table_e = _require_float(table_v, "node embedding table"), _require_float(table_e, "edge embedding table")
        node_types = _require(node_types, "node type indices", torch.int64, 2)
        edge_types = _require(edge_types, "edge type indices", torch.int64, 2)
        V, E, d = node_types.shape[0], edge_types.shape[0], table_v.shape[1]
        if table_e.shape[1] != d or E != csr.E or V != csr.V:
            raise RuntimeError(f"notorch_b200: fused embedding shape mismatch: tables {tuple(table_v.shape)} / {tuple(table_e.shape)}, "
                               f"ids {tuple(node_types.shape)} / {tuple(edge_types.shape)}, graph V={csr.V} E={csr.E}")
        h0 = _embed_edge_init_raw(table_v, table_e, node_types, edge_types, csr.src, V)
        ctx.save_for_backward(node_types, edge_types)
        ctx.csr, ctx.shapes = csr, (table_v.shape, table_e.shape)
        return h0

    @staticmethod
    def backward(ctx, g: Tensor):
        node_types, edge_types = ctx.saved_tensors
        (Tv, d), (Te, _) = ctx.shapes
        csr = ctx.csr
        gv, ge = _embed_edge_init_backward_raw(g.contiguous(), node_types, edge_types, csr.src, csr.V, Tv, Te)
        return gv, ge, None, None, None


_fuse_embedding = os.environ.get("NOTORCH_B200_FUSE_EMBED", "1") != "0"
# a Sum / Mean / Norm read-out of a sum-reduced block is a sum over each molecule's EDGES: run it over h_L directly and leave the
# block's node_feats a placeholder that is computed only if something reads it (agg.py)
_fuse_readout = os.environ.get("NOTORCH_B200_FUSE_READOUT", "1") != "0"
_pooled_backward = os.environ.get("NOTORCH_B200_POOLED_LAST", "1") != "0"  # last depth collapsed onto the molecules under a sum read-out (§5.10)


def embed_edge_init_supported(table_v: Tensor, table_e: Tensor, node_types: Tensor, edge_types: Tensor) -> bool:
    """Whether ``nt_embed_edge_init`` (+ backward) takes this problem; otherwise the caller materialises x_v / x_e and runs K0."""
    if not _fuse_embedding or not (table_v.is_cuda and table_v.dtype == torch.float32 and table_e.dtype == torch.float32):
        return False
    d, T = table_v.shape[1], table_v.shape[0] + table_e.shape[0]
    if d % 4 != 0 or table_e.shape[1] != d or node_types.dim() != 2 or edge_types.dim() != 2:
        return False
    if node_types.shape[1] + edge_types.shape[1] > 32 or edge_types.shape[0] == 0:
        return False
    return T * 16 + 128 * 32 * 4 <= 200 * 1024 and _lib.lib().nt_embed_edge_init_backward_workspace_bytes(max(edge_types.shape[0], 1), table_v.shape[0],
                                                                                                         table_e.shape[0], d) > 0


def embed_edge_init(table_v: Tensor, table_e: Tensor, node_types: Tensor, edge_types: Tensor, csr: "GraphCSR") -> Tensor:
    if _via_ops(table_v, node_types):
        _torch_ops()
        return torch.ops.notorch_b200.embed_edge_init(table_v, table_e, node_types, edge_types, csr.src, csr.V)
    return _EmbedEdgeInit.apply(table_v, table_e, node_types, edge_types, csr)


_dropout_calls = 0


def _weight_images_begin(W: Tensor, want_transposed: bool, E: int) -> tuple[Tensor | None, Tensor | None, "torch.cuda.Event | None"]:
    """The forward weight image and (``want_transposed``) the image of W^T that the backward reads, filled on a side stream so that
    the two small launches run under whatever the caller launches next; the caller waits on the returned event before it reads
    them (``None``: they were filled on the calling stream). Buffers come from the caller's stream - it is the one that reads and
    frees them."""
    img = _weight_image_alloc(W)
    img_t = _weight_image_alloc(W) if want_transposed else None
    if not (_side_wprep and _timer is None and E > 0):
        _weight_image_fill(W, False, img)
        _weight_image_fill(W, True, img_t)
        return img, img_t, None
    main = torch.cuda.current_stream()
    fork = torch.cuda.Event()
    fork.record(main)
    side = _side_stream(W.device, 3)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        _weight_image_fill(W, False, img)
        _weight_image_fill(W, True, img_t)
        join = torch.cuda.Event()
        join.record(side)
    return img, img_t, join


def _layer_forward_raw(h: Tensor, W: Tensor, b: Tensor | None, csr: GraphCSR, act: int, act_param: float, mean: bool, residual: bool,
                       p: float, seed: int, offset: int, mode: int, save_m: bool,
                       extreme: int = 0, img_t_out: list | None = None) -> tuple[Tensor, Tensor | None, Tensor, Tensor | None]:
    """K1 + K2 of one depth on raw tensors (no autograd): returns (h', m or None, n, arg or None).
    ``extreme``: 0 = sum / mean (``mean``), 1 = max, 2 = min — K1 is then the arg-reduction and ``arg`` its [V, d] argument rows.
    ``img_t_out``: a list that receives the weight image of the TRANSPOSED weight (what K4a reads), prepared here, beside the forward
    image, on a side stream under K1 — the two small launches are then off the critical path of both passes."""
    E, d = h.shape
    L = _lib.lib()
    with torch.cuda.device(h.device):
        arg = None
        use_img = mode != GEMM_FP32
        img, img_t, join = _weight_images_begin(W, img_t_out is not None, E) if use_img else (None, None, None)
        if extreme:
            n, arg = _seg_extreme_raw(h, csr.by_dst, act, act_param, extreme == 2)
        else:
            n = _seg_reduce_raw(h, csr.by_dst, act, act_param, mean, tag="K1")
        if join is not None:
            torch.cuda.current_stream().wait_event(join)
        if img_t_out is not None and img_t is not None:
            img_t_out.append(img_t)
        out = torch.empty_like(h)
        m = torch.empty_like(h) if save_m else None
        _run("K2:nt_layer_forward", L.nt_layer_forward, _p(h), _p(n), _p(csr.src), _p(csr.rev), _p(W), _p(img), _p(b), E, csr.V, d, act,
             act_param, int(residual), p, seed, offset, _p(out), _p(m), NT_F32, mode, _stream())
    return out, m, n, arg


def _layer_backward_raw(g: Tensor, h: Tensor, m: Tensor | None, n: Tensor | None, W: Tensor, has_bias: bool, csr: GraphCSR, act: int,
                        act_param: float, mean: bool, residual: bool, p: float, seed: int, offset: int, mode: int, need_w: bool,
                        need_h: bool, arg: Tensor | None = None,
                        img_t: Tensor | None = None) -> tuple[Tensor | None, Tensor | None, Tensor | None]:
    """K4b, K4a, K5 + K6 of one depth on raw tensors: returns (g_h, g_W, g_b). ``arg``: argument rows of a max / min forward;
    ``img_t``: the transposed weight image if the forward pass already prepared it."""
    E, d = h.shape
    g = g.contiguous()
    L = _lib.lib()
    gW = gb = gh = None
    with torch.cuda.device(g.device):
        if need_w:
            gW = torch.empty_like(W)
            gb = torch.empty(d, dtype=W.dtype, device=W.device) if has_bias else None
            nbytes = L.nt_layer_backward_wgrad_workspace_bytes(E, d)
            ws = _workspace(g.device, nbytes, slot=1)
            _run("K4b:nt_layer_backward_wgrad", L.nt_layer_backward_wgrad, _p(g), _p(m), _p(h), _p(n), _p(csr.src), _p(csr.rev), E, csr.V, d,
                 act, act_param, p, seed, offset, _p(gW), _p(gb), _p(ws), ws.numel(), NT_F32, mode, _stream())
        if need_h:
            if img_t is None and mode != GEMM_FP32:
                img_t = _weight_image(W, True)
            g_m = torch.empty_like(h)
            _run("K4a:nt_layer_backward_dgrad", L.nt_layer_backward_dgrad, _p(g), _p(W), _p(img_t), E, d, p, seed, offset, _p(g_m), NT_F32,
                 mode, _stream())
            gh = torch.empty_like(h)
            ell = _ell_of(csr.by_src) if (_fuse_k5_k6 and d % 4 == 0 and arg is None) else None
            if arg is not None:  # max / min: the atom gradient goes to the argument edge of every (atom, channel) only
                g_n = _seg_reduce_raw(g_m, csr.by_src, tag="K5")
                _run("K6:nt_layer_backward_epilogue_arg", L.nt_layer_backward_epilogue_arg, _p(g), _p(h), _p(g_n), _p(g_m), _p(csr.dst), _p(arg),
                     _p(csr.by_rev.rowptr), _p(csr.by_rev.perm), E, d, act, act_param, int(residual), _p(gh), NT_F32, _stream())
            elif ell is not None:  # K5 + K6 in one kernel: g_n is never materialised
                _run("K6:nt_layer_backward_epilogue", L.nt_layer_backward_epilogue_fused, _p(g), _p(h), _p(g_m), _p(csr.dst),
                     _p(csr.by_src.rowptr), _p(csr.by_src.perm), _p(ell), _p(csr.by_rev.rowptr), _p(csr.by_rev.perm), _p(csr.by_dst.rowptr),
                     E, d, act, act_param, int(residual), int(mean), _p(gh), NT_F32, _stream())
            else:
                g_n = _seg_reduce_raw(g_m, csr.by_src, tag="K5")
                _run("K6:nt_layer_backward_epilogue", L.nt_layer_backward_epilogue, _p(g), _p(h), _p(g_n), _p(g_m), _p(csr.dst),
                     _p(csr.by_rev.rowptr), _p(csr.by_rev.perm), _p(csr.by_dst.rowptr), E, d, act, act_param, int(residual), int(mean),
                     _p(gh), NT_F32, _stream())
    return gh, gW, gb


class _Layer(torch.autograd.Function):
    """One message-passing depth: K1 + K2 forward, K4a + K4b + K5 + K6 backward.

    h' = [h +] Dropout(Linear(n[src] - act(h)[rev])),  n = reduce_dst(act(h))   (chemprop.py:36-41, residual.py:28)
    """

    @staticmethod
    def forward(ctx, h: Tensor, W: Tensor, b: Tensor | None, csr: GraphCSR, act: int, act_param: float, mean: bool,
                residual: bool, p: float, seed: int, offset: int, mode: int, extreme: int = 0):
        h = _require_float(h, "edge_feats")
        W = _require_float(W, "weight")
        E, d = h.shape
        if E != csr.E:
            raise RuntimeError(f"notorch_b200: edge_feats has {E} rows but edge_index has {csr.E} edges")
        if W.shape != (d, d):
            raise RuntimeError(f"notorch_b200: weight {tuple(W.shape)} does not match hidden size {d}")
        if b is not None:
            b = _require(b, "bias", torch.float32, 1)
        # tensor-core path: K2 also writes the message tensor m, which K4b then streams as dense tiles
        save_m = _save_messages and mode != GEMM_FP32 and d % 4 == 0 and (ctx.needs_input_grad[1] or (b is not None and ctx.needs_input_grad[2]))
        img_t_out: list | None = [] if (ctx.needs_input_grad[0] and mode != GEMM_FP32) else None
        out, m, n, arg = _layer_forward_raw(h, W, b, csr, act, act_param, mean, residual, p, seed, offset, mode, save_m, extreme, img_t_out)
        ctx.save_for_backward(h, m if save_m else n, W, arg)
        ctx.img_t = img_t_out[0] if img_t_out else None  # W is a saved tensor: autograd refuses to run backward if it was changed in place
        ctx.csr, ctx.cfg, ctx.has_bias, ctx.has_m = csr, (act, act_param, mean, residual, p, seed, offset, mode), b is not None, save_m
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        h, n_or_m, W, arg = ctx.saved_tensors
        m, n = (n_or_m, None) if ctx.has_m else (None, n_or_m)
        need_w = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        gh, gW, gb = _layer_backward_raw(g, h, m, n, W, ctx.has_bias, ctx.csr, *ctx.cfg, need_w, ctx.needs_input_grad[0], arg, ctx.img_t)
        return gh, gW, gb, None, None, None, None, None, None, None, None, None, None


class _LastDepthPooled(torch.autograd.Function):
    """The LAST message-passing depth as seen through a sum read-out over each molecule's edges (chemprop.py:37-41,
    residual.py:28 + agg.py:27,36 + the identity of §5.9): ``H_sum[b] = sum_{e in b} h_L[e]`` from ``h = h_{L-1}`` WITHOUT computing
    h_L. Forward: one pass over h gives ``M = sum_{e in b} m[e]`` and ``S = sum_{e in b} h[e] + |b| bias`` ([B, d]), the Linear runs on
    B rows. Backward: the gradient of h_L is the broadcast ``g[e] = G[mol(e)]``, so every contraction over the E edges collapses to
    one over the B molecules (``pooled_backward.cu``). No [E, d] tensor besides h itself is read or written, none is saved."""

    @staticmethod
    def forward(ctx, h: Tensor, W: Tensor, b: Tensor | None, csr: GraphCSR, pool: SegmentCSR, act: int, act_param: float, mean: bool,
                residual: bool, mode: int):
        h = _require_float(h, "edge_feats")
        W = _require_float(W, "weight")
        E, d = h.shape
        if E != csr.E or pool.keys32.numel() != E:
            raise RuntimeError(f"notorch_b200: edge_feats has {E} rows but the graph has {csr.E} edges")
        if W.shape != (d, d):
            raise RuntimeError(f"notorch_b200: weight {tuple(W.shape)} does not match hidden size {d}")
        if b is not None:
            b = _require(b, "bias", torch.float32, 1)
        if mode == GEMM_FP32 or d % 4 != 0:
            raise RuntimeError("notorch_b200: the pooled last depth needs a tensor-core GEMM mode and d % 4 == 0")
        B = pool.num_segments
        L = _lib.lib()
        with torch.cuda.device(h.device):
            img, img_t, join = _weight_images_begin(W, ctx.needs_input_grad[0], E)
            M = torch.empty((B, d), dtype=h.dtype, device=h.device)
            S = torch.empty_like(M)
            ws = _workspace(h.device, L.nt_pooled_message_sum_workspace_bytes(E))
            _run("Kpm:nt_pooled_message_sum", L.nt_pooled_message_sum, _p(h), _p(csr.rev), _p(csr.dst), _p(csr.by_src.rowptr), _p(csr.by_dst.rowptr),
                 _p(pool.rowptr), _p(b), E, B, d, act, act_param, int(residual), int(mean), _p(M), _p(S), _p(ws), ws.numel(), NT_F32, _stream())
            if join is not None:
                torch.cuda.current_stream().wait_event(join)
            H = torch.empty_like(M)
            _run("K2p:nt_dense_forward", L.nt_dense_forward, _p(M), _p(img), None, _p(S), B, d, 0.0, 0, 0, _p(H), NT_F32, mode, _stream())
        ctx.save_for_backward(h, M, W)
        ctx.img_t = img_t
        ctx.csr, ctx.pool, ctx.cfg, ctx.has_bias = csr, pool, (act, act_param, mean, residual, mode), b is not None
        return H

    @staticmethod
    def backward(ctx, G: Tensor):
        h, M, W = ctx.saved_tensors
        act, act_param, mean, residual, mode = ctx.cfg
        csr, pool = ctx.csr, ctx.pool
        need_w = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        need_h = ctx.needs_input_grad[0]
        E, d = h.shape
        G = G.contiguous()
        B = G.shape[0]
        L = _lib.lib()
        gh = gW = gb = None
        with torch.cuda.device(h.device):
            if need_w:
                gW = torch.empty_like(W)
                ws = _workspace(h.device, L.nt_layer_backward_wgrad_workspace_bytes(B, d), slot=1)
                _run("K4bp:nt_layer_backward_wgrad", L.nt_layer_backward_wgrad, _p(G), _p(M), None, None, None, None, B, 1, d, act, act_param, 0.0, 0, 0,
                     _p(gW), None, _p(ws), ws.numel(), NT_F32, mode, _stream())
                if ctx.has_bias:
                    gb = torch.empty(d, dtype=W.dtype, device=W.device)
                    _run("gbp:nt_weighted_colsum", L.nt_weighted_colsum, _p(G), _p(pool.rowptr), B, d, _p(gb), NT_F32, _stream())
            if need_h:
                img_t = ctx.img_t if ctx.img_t is not None else _weight_image(W, True)
                GW = torch.empty_like(G)
                _run("K4ap:nt_layer_backward_dgrad", L.nt_layer_backward_dgrad, _p(G), _p(W), _p(img_t), B, d, 0.0, 0, 0, _p(GW), NT_F32, mode, _stream())
                gh = torch.empty_like(h)
                ws6 = _workspace(h.device, L.nt_layer_backward_epilogue_pooled_workspace_bytes(E))
                _run("K6p:nt_layer_backward_epilogue_pooled", L.nt_layer_backward_epilogue_pooled, _p(G), _p(GW), _p(h), _p(pool.keys32), _p(csr.dst),
                     _p(csr.by_src.rowptr), _p(csr.by_rev.rowptr), _p(csr.by_rev.perm), _p(csr.by_dst.rowptr), E, B, d, act, act_param, int(residual),
                     int(mean), _p(gh), _p(ws6), ws6.numel(), NT_F32, _stream())
        return (gh, gW, gb) + (None,) * 7


def last_depth_pooled_supported(h: Tensor, dropout: float, training: bool, reduce: str) -> bool:
    """Whether the collapsed form applies: sum / mean reduction inside the depth, no active dropout (a per-edge mask does not commute
    with the sum over a molecule's edges), a tensor-core GEMM mode, 16-byte rows, real CUDA tensors."""
    return (_pooled_backward and reduce in ("sum", "mean") and not (training and dropout > 0.0) and _gemm_mode != GEMM_FP32 and h.shape[1] % 4 == 0
            and h.is_cuda and h.dtype == torch.float32 and not _via_ops(h))


def last_depth_pooled(h: Tensor, weight: Tensor, bias: Tensor | None, csr: GraphCSR, pool: SegmentCSR, *,
                      act: tuple[int, float] = (_lib.ACT_RELU, 0.0), reduce: str = "sum", residual: bool = True) -> Tensor:
    """``sum_{e in b} h'[e]`` for ``h' = [h +] Linear(n[src] - act(h)[rev])`` without computing h' (``_LastDepthPooled``)."""
    if reduce not in ("sum", "mean"):
        raise ValueError("notorch_b200: the pooled last depth needs a sum / mean reduction")
    return _LastDepthPooled.apply(h, weight, bias, csr, pool, act[0], act[1], reduce == "mean", residual, _gemm_mode)


# ------------------------------------------------------------------------------------------------
# public functional API
# ------------------------------------------------------------------------------------------------

_REDUCTIONS = ("sum", "mean", "max", "min")  # notorch/types.py: Reduction


def seg_reduce(x: Tensor, csr: SegmentCSR, reduce: str = "sum", scale: float = 1.0, tag: str = "K1") -> Tensor:
    if reduce not in _REDUCTIONS:
        raise ValueError(f"notorch_b200: unknown reduce '{reduce}' (one of {_REDUCTIONS})")
    if reduce in ("max", "min"):
        if scale != 1.0:
            raise ValueError("notorch_b200: a scale factor only applies to sum / mean reductions")
        if _via_ops(x):
            _torch_ops()
            return torch.ops.notorch_b200.seg_extreme(x, csr.rowptr, csr.perm, csr.keys32, csr.num_segments, reduce == "min")[0]
        return _SegExtreme.apply(x, csr, reduce == "min", tag + "x")
    if _via_ops(x):
        _torch_ops()
        return torch.ops.notorch_b200.seg_reduce(x, csr.rowptr, csr.perm, csr.keys32, csr.num_segments, reduce == "mean", float(scale))
    return _SegReduce.apply(x, csr, reduce == "mean", float(scale), tag)


def gather_add(base: Tensor, x: Tensor, csr: SegmentCSR) -> Tensor:
    return _GatherAdd.apply(base, x, csr)


def edge_init(node_feats: Tensor, edge_feats: Tensor, csr: GraphCSR) -> Tensor:
    """K0: ``h0 = node_feats[src] + edge_feats`` (chemprop.py:83)."""
    if _via_ops(node_feats, edge_feats):
        return _torch_ops().edge_init_op(node_feats, edge_feats, csr)
    return _GatherAdd.apply(edge_feats, node_feats, csr.by_src)


def edge_to_atom(edge_feats: Tensor, csr: GraphCSR, reduce: str = "sum") -> Tensor:
    """K1 without activation: ``scatter(edge_feats, dst, dim_size=V, reduce)`` (chemprop.py:86)."""
    return seg_reduce(edge_feats, csr.by_dst, reduce)


def readout(node_feats: Tensor, mol_csr: SegmentCSR, kind: str = "sum", norm: float = 100.0) -> Tensor:
    """K3: ``scatter_sum`` / ``scatter_mean`` over ``batch_node_index`` (agg.py:27,36); ``norm`` = sum / constant."""
    if kind == "norm":
        return seg_reduce(node_feats, mol_csr, "sum", 1.0 / norm, tag="K3")
    return seg_reduce(node_feats, mol_csr, kind, tag="K3")


def layer(h: Tensor, weight: Tensor, bias: Tensor | None, csr: GraphCSR, *, act: tuple[int, float] = (_lib.ACT_RELU, 0.0),
          reduce: str = "sum", residual: bool = True, dropout: float = 0.0, training: bool = False) -> Tensor:
    """One fused message-passing depth (K1+K2; hand-written backward K4-K6)."""
    global _dropout_calls
    if reduce not in _REDUCTIONS:
        raise ValueError(f"notorch_b200: unknown reduce '{reduce}' (one of {_REDUCTIONS})")
    p = float(dropout) if training else 0.0
    if not 0.0 <= p < 1.0:
        if p == 1.0:
            raise NotImplementedError("notorch_b200: dropout p=1.0 is not supported")
        raise ValueError(f"dropout probability has to be between 0 and 1, but got {p}")
    seed = offset = 0
    if p > 0.0:
        seed = int(torch.empty((), dtype=torch.int64).random_().item())  # CPU generator: follows torch.manual_seed, no device sync
        _dropout_calls += 1
        offset = _dropout_calls
    extreme = {"max": 1, "min": 2}.get(reduce, 0)
    if extreme == 0 and _via_ops(h, weight):
        return _torch_ops().layer_from_csr(h, weight, bias, csr, act[0], act[1], reduce == "mean", residual, p, seed, offset, _gemm_mode)
    return _Layer.apply(h, weight, bias, csr, act[0], act[1], reduce == "mean", residual, p, seed, offset, _gemm_mode, extreme)


# ------------------------------------------------------------------------------------------------
# atom-state message passing (extension named by BASELINE configs[3]; not in the reference tree)
# ------------------------------------------------------------------------------------------------

@dataclass
class AtomCSR:
    """Atom-to-atom neighbour lists derived from a GraphCSR: for every atom the SOURCE atoms of its incoming edges
    (``nbr_in``, ascending edge order) and the DESTINATION atoms of its outgoing edges (``nbr_out``)."""

    V: int
    E: int
    nbr_in: SegmentCSR   # rowptr = by_dst.rowptr, perm[j] = src[by_dst.perm[j]]
    nbr_out: SegmentCSR  # rowptr = by_src.rowptr, perm[j] = dst[by_src.perm[j]]
    ident: Tensor        # arange(V) int32 (identity gather index)


def atom_csr(csr: GraphCSR) -> AtomCSR:
    cached = getattr(csr, "_atom_csr", None)
    if cached is not None:
        return cached
    src, dst = csr.src, csr.dst
    perm_in = src[csr.by_dst.perm.long()].contiguous()
    perm_out = dst[csr.by_src.perm.long()].contiguous()
    acsr = AtomCSR(csr.V, csr.E, SegmentCSR(csr.by_dst.rowptr, perm_in, perm_in, csr.V), SegmentCSR(csr.by_src.rowptr, perm_out, perm_out, csr.V),
                   torch.arange(csr.V, dtype=torch.int32, device=src.device))
    csr._atom_csr = acsr
    return acsr


def _seg_reduce_ex_raw(x: Tensor, seg: SegmentCSR, act: int, act_param: float, mean: bool, base: Tensor | None, dact_of: Tensor | None,
                       tag: str) -> Tensor:
    d = x.shape[1]
    out = torch.empty((seg.num_segments, d), dtype=x.dtype, device=x.device)
    ell = _ell_of(seg)
    if ell is not None:
        _run(f"{tag}:nt_seg_reduce_ex", _lib.lib().nt_seg_reduce_ell, _p(x), d, _p(seg.rowptr), _p(seg.perm), _p(ell), seg.num_segments, act,
             act_param, int(mean), 1.0, _p(base), _p(dact_of), _p(out), NT_F32, _stream())
        return out
    _run(f"{tag}:nt_seg_reduce_ex", _lib.lib().nt_seg_reduce_ex, _p(x), d, _p(seg.rowptr), _p(seg.perm), seg.num_segments, act, act_param,
         int(mean), 1.0, _p(base), _p(dact_of), _p(out), NT_F32, _stream())
    return out


def _atom_layer_forward_raw(h: Tensor, s_e: Tensor, W: Tensor, b: Tensor | None, acsr: AtomCSR, act: int, act_param: float, mean: bool,
                            residual: bool, p: float, seed: int, offset: int, mode: int) -> tuple[Tensor, Tensor]:
    h, s_e, W = _require_float(h, "node_feats"), _require_float(s_e, "edge aggregate"), _require_float(W, "weight")
    V, d = h.shape
    if V != acsr.V or s_e.shape != h.shape or W.shape != (d, d):
        raise RuntimeError(f"notorch_b200: atom layer shape mismatch: h {tuple(h.shape)}, s_e {tuple(s_e.shape)}, W {tuple(W.shape)}, V={acsr.V}")
    if mode == GEMM_FP32 or d % 4 != 0:
        raise NotImplementedError("notorch_b200: atom message passing runs on the tensor-core path only (gemm_mode tf32x3 / tf32, d % 4 == 0)")
    if b is not None:
        b = _require(b, "bias", torch.float32, 1)
    L = _lib.lib()
    with torch.cuda.device(h.device):
        n = _seg_reduce_ex_raw(h, acsr.nbr_in, act, act_param, mean, s_e, None, tag="A1")
        img = _weight_image(W, False)
        out = torch.empty_like(h)
        _run("A2:nt_dense_forward", L.nt_dense_forward, _p(n), _p(img), _p(b), _p(h) if residual else None, V, d, p, seed, offset, _p(out),
             NT_F32, mode, _stream())
    return out, n


def _atom_layer_backward_raw(g: Tensor, h: Tensor, n: Tensor, W: Tensor, has_bias: bool, acsr: AtomCSR, act: int, act_param: float, mean: bool,
                             residual: bool, p: float, seed: int, offset: int, mode: int, need_w: bool, need_h: bool):
    V, d = h.shape
    L = _lib.lib()
    gW = gb = gh = gs = None
    with torch.cuda.device(g.device):
        if need_w:
            gW = torch.empty_like(W)
            gb = torch.empty(d, dtype=W.dtype, device=W.device) if has_bias else None
            ws = _workspace(g.device, L.nt_layer_backward_wgrad_workspace_bytes(V, d), slot=1)
            _run("A4b:nt_layer_backward_wgrad", L.nt_layer_backward_wgrad, _p(g), _p(n), None, None, None, None, V, V, d, act, act_param, p, seed,
                 offset, _p(gW), _p(gb), _p(ws), ws.numel(), NT_F32, mode, _stream())
        if need_h:
            g_n = torch.empty_like(h)
            _run("A4a:nt_layer_backward_dgrad", L.nt_layer_backward_dgrad, _p(g), _p(W), _p(_weight_image(W, True)), V, d, p, seed, offset,
                 _p(g_n), NT_F32, mode, _stream())
            gs = g_n  # n = s_e + reduce(...): the edge aggregate receives g_n as is
            # mean: every incoming message of atom v carries 1 / indeg(v)
            g_msg = _gather_add_raw(None, g_n, acsr.ident, acsr.nbr_in.rowptr, tag="A5") if mean else g_n
            gh = _seg_reduce_ex_raw(g_msg, acsr.nbr_out, act, act_param, False, g if residual else None, h, tag="A6")
    return gh, gs, gW, gb


class _AtomLayer(torch.autograd.Function):
    """One depth of atom-state message passing:

        n[v]  = reduce_{e: dst[e]=v} (act(h)[src[e]] + x_e[e]) = reduce_in(act(h)[src]) + s_e[v]      (s_e = reduce_dst(x_e), once per batch)
        h'    = [h +] Dropout(Linear(n))

    forward: nt_seg_reduce_ex (neighbour gather-reduce with activation prologue, no [E,d] intermediate) + nt_dense_forward (tcgen05);
    backward: K4b/K4a on [V,d] operands, then nt_seg_reduce_ex in its backward form over the outgoing-neighbour list.
    """

    @staticmethod
    def forward(ctx, h: Tensor, s_e: Tensor, W: Tensor, b: Tensor | None, acsr: AtomCSR, act: int, act_param: float, mean: bool,
                residual: bool, p: float, seed: int, offset: int, mode: int):
        out, n = _atom_layer_forward_raw(h, s_e, W, b, acsr, act, act_param, mean, residual, p, seed, offset, mode)
        ctx.save_for_backward(h.contiguous(), n, W.contiguous())
        ctx.acsr, ctx.cfg, ctx.has_bias = acsr, (act, act_param, mean, residual, p, seed, offset, mode), b is not None
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        h, n, W = ctx.saved_tensors
        need_w = ctx.needs_input_grad[2] or (ctx.has_bias and ctx.needs_input_grad[3])
        need_h = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        gh, gs, gW, gb = _atom_layer_backward_raw(g.contiguous(), h, n, W, ctx.has_bias, ctx.acsr, *ctx.cfg, need_w, need_h)
        return gh, gs, gW, gb, None, None, None, None, None, None, None, None, None


def atom_layer(h: Tensor, s_e: Tensor, weight: Tensor, bias: Tensor | None, acsr: AtomCSR, *, act: tuple[int, float] = (_lib.ACT_RELU, 0.0),
               reduce: str = "sum", residual: bool = True, dropout: float = 0.0, training: bool = False) -> Tensor:
    global _dropout_calls
    if reduce not in ("sum", "mean"):
        raise NotImplementedError(f"notorch_b200: reduce='{reduce}' is not implemented (sum and mean are); no fallback")
    p = float(dropout) if training else 0.0
    if not 0.0 <= p < 1.0:
        raise ValueError(f"dropout probability has to be in [0, 1), but got {p}")
    seed = offset = 0
    if p > 0.0:
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        _dropout_calls += 1
        offset = _dropout_calls
    if _via_ops(h, weight):
        _torch_ops()
        return torch.ops.notorch_b200.atom_layer(h, s_e, weight, bias, acsr.nbr_in.rowptr, acsr.nbr_in.perm, acsr.nbr_out.rowptr, acsr.nbr_out.perm,
                                                 acsr.ident, act[0], act[1], reduce == "mean", residual, p, seed, offset, _gemm_mode)[0]
    return _AtomLayer.apply(h, s_e, weight, bias, acsr, act[0], act[1], reduce == "mean", residual, p, seed, offset, _gemm_mode)


def dropout_mask(n_rows: int, d: int, p: float, seed: int, offset: int, device) -> Tensor:
    """The keep-mask K2/K4 derive from (seed, offset) — exposed for tests."""
    mask = torch.empty((n_rows, d), dtype=torch.float32, device=device)
    with torch.cuda.device(mask.device):
        _run("nt_dropout_mask", _lib.lib().nt_dropout_mask, n_rows, d, p, seed, offset, _p(mask), _stream())
    return mask


def collate_packed(num_atoms: Tensor, num_edges: Tensor, local_edge_index: Tensor, local_rev_index: Tensor,
                   V: int, E: int, fixed_rev: bool = False) -> dict[str, Tensor]:
    """K-l: device-side ``BatchedGraph.from_graphs`` on packed int32 inputs already on the GPU
    (graph.py:186-223). Returns the reference's int64 index tensors plus int32 molecule row pointers."""
    if _via_ops(num_atoms, local_edge_index):
        _torch_ops()
        outs = torch.ops.notorch_b200.collate(num_atoms, num_edges, local_edge_index, local_rev_index, V, E, fixed_rev)
        return dict(zip(("edge_index", "rev_index", "batch_node_index", "batch_edge_index", "mol_atom_ptr", "mol_edge_ptr"), outs))
    return _collate_packed_raw(num_atoms, num_edges, local_edge_index, local_rev_index, V, E, fixed_rev)


def _collate_packed_raw(num_atoms: Tensor, num_edges: Tensor, local_edge_index: Tensor, local_rev_index: Tensor,
                        V: int, E: int, fixed_rev: bool = False) -> dict[str, Tensor]:
    num_atoms = _require(num_atoms, "num_atoms", torch.int32, 1)
    num_edges = _require(num_edges, "num_edges", torch.int32, 1)
    lei = _require(local_edge_index, "local_edge_index", torch.int32, 2)
    lrev = _require(local_rev_index, "local_rev_index", torch.int32, 1)
    B, dev = num_atoms.numel(), num_atoms.device
    if lei.shape != (2, E) or lrev.shape != (E,) or num_edges.numel() != B:
        raise RuntimeError("notorch_b200: packed molecule arrays have inconsistent shapes")
    L = _lib.lib()
    with torch.cuda.device(dev):
        out = {
            "edge_index": torch.empty((2, E), dtype=torch.int64, device=dev),
            "rev_index": torch.empty(E, dtype=torch.int64, device=dev),
            "batch_node_index": torch.empty(V, dtype=torch.int64, device=dev),
            "batch_edge_index": torch.empty(E, dtype=torch.int64, device=dev),
            "mol_atom_ptr": torch.empty(B + 1, dtype=torch.int32, device=dev),
            "mol_edge_ptr": torch.empty(B + 1, dtype=torch.int32, device=dev),
        }
        ws = _workspace(dev, L.nt_collate_workspace_bytes(B))
        _run("collate:nt_collate", L.nt_collate, _p(num_atoms), _p(num_edges), B, _p(lei), _p(lrev), V, E, int(fixed_rev), _p(out["edge_index"]),
             _p(out["rev_index"]), _p(out["batch_node_index"]), _p(out["batch_edge_index"]), _p(out["mol_atom_ptr"]), _p(out["mol_edge_ptr"]),
             _p(ws), ws.numel(), _stream())
    return out
