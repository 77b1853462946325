"""
Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over
NVLink / NVSwitch) for the only exchange the path has.

What shards and how:

* fit       -- data rows are split into contiguous ranges, one per rank; each rank
               accumulates the partial normal equations of ITS rows (G, J^T W d and
               the column statistics are all plain sums over rows); ONE all-reduce
               (sum, float64) of the packed lower triangle + the 4*M statistics
               vector follows.  The m x m Cholesky is then shared too: outer panels
               (8 block columns) are owned cyclically, the owner factors its panel and
               broadcasts it (one contiguous range of the tile-packed triangle), every
               rank applies it to the block columns it owns; the substitution runs
               replicated on the complete factor (``set_cholesky``).
* predict   -- the raveled prediction points are split into contiguous ranges; no
               communication during compute, one all-gather of the results.
* SplineCV  -- independent (fold, mindist) Gram jobs are dealt round-robin to ranks;
               one all-reduce of the (zero-filled) score table at the end.

With a single process (or ``torch.distributed`` not initialised) every helper
degenerates to the local computation.  The helpers that only do index
arithmetic or collectives on small tensors also run on the ``gloo`` backend,
which is how tests/test_dist_cpu.py covers the N > 1 logic without a GPU.
"""
import numpy as np

_SHARDING = "auto"


def set_sharding(mode):
    """``"auto"`` (shard when torch.distributed has world_size > 1), ``True`` or ``False``."""
    global _SHARDING
    if mode not in ("auto", True, False):
        raise ValueError("sharding mode must be 'auto', True or False")
    _SHARDING = mode


def world():
    "(rank, world_size) of the default process group, or (0, 1)."
    try:
        import torch.distributed as td
    except ImportError:  # pragma: no cover
        return 0, 1
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def backend():
    "Backend name of the default process group ('nccl' / 'gloo'), or None without one."
    try:
        import torch.distributed as td
    except ImportError:  # pragma: no cover
        return None
    if td.is_available() and td.is_initialized():
        return td.get_backend()
    return None


def sharding_active():
    if _SHARDING is False:
        return False
    return world()[1] > 1


def shard_bounds(total, rank=None, size=None):
    "Contiguous, balanced [start, stop) of *total* items for *rank* of *size*."
    if rank is None or size is None:
        rank, size = world()
    base, extra = divmod(int(total), int(size))
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return start, stop


def round_robin(njobs, rank=None, size=None):
    "Indices of the jobs this rank owns when jobs are dealt round-robin."
    if rank is None or size is None:
        rank, size = world()
    return list(range(rank, int(njobs), size))


_CHOLESKY = "auto"


def set_cholesky(mode):
    """
    How the m x m Cholesky of a row-sharded fit runs on N > 1 ranks: ``"replicated"`` (every rank
    factors the all-reduced matrix: the Amdahl term), ``"distributed"`` (outer panels owned
    cyclically, panel broadcast over NCCL, each rank updates its own block columns) or ``"auto"``
    (distributed when there are at least two outer panels per rank).
    """
    global _CHOLESKY
    if mode not in ("auto", "replicated", "distributed"):
        raise ValueError("cholesky mode must be 'auto', 'replicated' or 'distributed'")
    _CHOLESKY = mode


def panel_cyclic_schedule(block_rows, panel_tiles, size):
    """
    Outer panels of the tiled Cholesky and their owners: a list of ``(k0, width, owner)`` covering
    block columns ``0..block_rows`` in panels of *panel_tiles*, panel p owned by rank ``p % size``.
    """
    block_rows, panel_tiles, size = int(block_rows), int(panel_tiles), int(size)
    if block_rows < 1 or panel_tiles < 1 or size < 1:
        raise ValueError("block_rows, panel_tiles and size must be positive")
    return [(k0, min(panel_tiles, block_rows - k0), p % size)
            for p, k0 in enumerate(range(0, block_rows, panel_tiles))]


def panel_tiles(size=None):
    """
    Outer-panel width (block columns of 128) of the panel-cyclic Cholesky.  The chain *panel arrives -> update the
    next panel -> factor it -> push it* is serial over the panels and its per-panel cost grows with the width, while
    per-rank update work shrinks with N: measured at m = 50k (profiles/r02m_dist_pipeline_8gpu_panel_widths.jsonl,
    r02n_dist_pipeline_4gpu_panel_widths.jsonl) N = 8: 234 / 200 / 189 / 182 ms for widths 12 / 8 / 6 / 4;
    N = 4: 327 / 325 / 325 / 339 ms for 8 / 6 / 4 / 2.  Identical on every rank.
    """
    if size is None:
        size = world()[1]
    return 4 if size >= 4 else 8


def owned_ranges(panels, rank):
    "Block-column ranges [lo, hi) of the panels in *panels* (schedule entries) that *rank* owns, ascending."
    return [(q0, q0 + qwidth) for q0, qwidth, qowner in panels if qowner == rank]


def use_distributed_cholesky(block_rows, panel_tiles, size=None):
    "Decision rule of ``set_cholesky``; identical on every rank (it only depends on sizes)."
    if size is None:
        size = world()[1]
    if size < 2 or _CHOLESKY == "replicated":
        return False
    if _CHOLESKY == "distributed":
        return True
    return len(panel_cyclic_schedule(block_rows, panel_tiles, size)) >= 2 * size


_PANEL_EXCHANGE = "broadcast"
_PANEL_PIPELINE = "peer"
_COMM_STREAM = {}
_EPOCH = [0]


def set_panel_pipeline(mode):
    """
    Schedule of the panel-cyclic Cholesky:

    * ``"serial"``    factor, ``ncclBroadcast``, update -- one after the other on one stream (round 1);
    * ``"lookahead"`` the owner of the next panel updates and factors it first; ``ncclBroadcast`` on a high-priority
      side stream beside the updates (the receivers' NCCL kernels hold SMs while they wait for the sender);
    * ``"peer"``      (default) the same look-ahead order, but the owner PUSHES each factored panel into every peer's
      matrix with copy engines over NVLink peer memory (CUDA IPC; next owner first) and raises a per-panel flag word;
      receivers park their stream on the flag.  No SM is spent on communication and nobody waits in a collective.

    Same arithmetic per tile in all three -> bit-identical factors.
    """
    global _PANEL_PIPELINE
    if mode not in ("serial", "lookahead", "peer"):
        raise ValueError("panel pipeline must be 'serial', 'lookahead' or 'peer'")
    _PANEL_PIPELINE = mode


def panel_pipeline():
    return _PANEL_PIPELINE


_MERGED_UPDATES = True
_READY_HANDSHAKE = True  # peer exchange: wait for every peer's "matrix in place" word before the first push (tools/
                         # run_dist_pipeline.py switches it off once to show that its skew test catches the race)


def set_merged_updates(flag):
    """
    Look-ahead / peer schedules: apply a factored panel to ALL the later panels a rank owns with one launch of the
    DMMA kernel (``vb200_chol_update_ranges``; default) instead of one launch per owned panel.  Same tiles, same
    operands, same order per tile -> bit-identical factors; what changes is one tail per outer panel instead of one
    per (outer panel, owned panel) pair.
    """
    global _MERGED_UPDATES
    _MERGED_UPDATES = bool(flag)


def merged_updates():
    return _MERGED_UPDATES


def next_epoch():
    "A number that grows with every collective call of the panel-cyclic factorisation (identical on every rank)."
    _EPOCH[0] += 1
    return _EPOCH[0]


class SharedBuffer:
    """
    A device buffer of this rank that the other ranks of the node can map (``vb200_shared_alloc``: cudaMalloc +
    CUDA-IPC handle), wrapped as a torch tensor without copying.  CUDA forbids freeing exported memory that is still
    mapped elsewhere, so ``free`` is only called (``GramSystem.release``) after every peer has closed its mapping
    (``PeerWindow.close``, then a barrier) or when no peer ever mapped it.
    """

    def __init__(self, count, dtype):
        import ctypes

        import torch

        from . import _cabi

        lib = _cabi.require_device()
        itemsize = torch.empty(0, dtype=dtype).element_size()
        self.nbytes = int(count) * itemsize
        ptr = ctypes.c_void_p()
        self.handle = ctypes.create_string_buffer(64)
        _cabi.check(lib.vb200_shared_alloc(self.nbytes, ctypes.byref(ptr), self.handle), "shared_alloc")
        self.ptr = int(ptr.value)
        typestr = {torch.float64: "<f8", torch.int64: "<i8"}[dtype]
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (self.ptr, False), "version": 2}
        self.tensor = torch.as_tensor(self, device=torch.device("cuda", torch.cuda.current_device()))

    def free(self):
        "cudaFree (idempotent).  The caller guarantees that no tensor view and no peer mapping is used afterwards."
        from . import _cabi

        if self.ptr:
            import ctypes

            self.tensor = None
            _cabi.load().vb200_shared_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0


class PeerWindow:
    """
    ``window.ptr(r, name)``: device address, valid on THIS rank's GPU, of rank r's ``SharedBuffer`` *name* -- the
    peers' buffers mapped into this process in this rank's own device context (``vb200_shared_open``), so that copies
    into them are NVLink peer writes done by this GPU's copy engines.  Construction is collective (one
    ``all_gather_object``); a mapping that fails (no peer access between two GPUs, CUDA IPC unavailable in the
    container) is remembered in ``self.error`` instead of raised, so that the caller can make the decision to fall
    back collective (``agree_on_status``) -- a rank that raised here would strand the others.
    """

    def __init__(self, buffers):
        import ctypes

        import torch.distributed as td

        from . import _cabi

        lib = _cabi.require_device()
        rank, size = world()
        self.local = dict(buffers)
        mine = {name: bytes(b.handle.raw) for name, b in self.local.items()}
        everyone = [None] * size
        td.all_gather_object(everyone, mine)
        self._ptr = {}
        self.error = None
        for r, entry in enumerate(everyone):
            for name, handle in (entry or {}).items():
                if r == rank:
                    self._ptr[(r, name)] = self.local[name].ptr
                elif self.error is None:
                    ptr = ctypes.c_void_p()
                    rc = lib.vb200_shared_open(ctypes.create_string_buffer(handle, 64), ctypes.byref(ptr))
                    if rc != 0:
                        self.error = _cabi.load().vb200_last_error().decode("utf-8", "replace")
                    else:
                        self._ptr[(r, name)] = int(ptr.value)
        if any(not entry for entry in everyone):
            self.error = self.error or "a peer has no shareable buffers"

    def ptr(self, r, name):
        return self._ptr[(r, name)]

    def close(self):
        "Unmap the peers' buffers from this process (local; the owners free them after a barrier)."
        import ctypes

        from . import _cabi

        rank, _ = world()
        lib = _cabi.load()
        for (r, _name), address in list(self._ptr.items()):
            if r != rank:
                lib.vb200_shared_close(ctypes.c_void_p(address))
        self._ptr = {}


def comm_stream():
    "The (cached, per device) high-priority CUDA stream the panel broadcasts are issued from."
    import torch

    dev = torch.cuda.current_device()
    if dev not in _COMM_STREAM:
        _COMM_STREAM[dev] = torch.cuda.Stream(device=dev, priority=-1)
    return _COMM_STREAM[dev]


def set_panel_exchange(mode):
    """
    How a factored Cholesky panel reaches the other ranks: ``"broadcast"`` (``ncclBroadcast``, a ring) or
    ``"allreduce"`` (every other rank zeroes its copy of the panel and the ranks all-reduce it: x + 0 = x
    exactly, and NCCL's all-reduce runs the NVLink-switch reduction (NVLS) and all channels, which its
    broadcast does not).
    """
    global _PANEL_EXCHANGE
    if mode not in ("broadcast", "allreduce"):
        raise ValueError("panel exchange must be 'broadcast' or 'allreduce'")
    _PANEL_EXCHANGE = mode


def broadcast_(tensor, src):
    "In-place broadcast of a torch tensor from rank *src* (NCCL for CUDA tensors, gloo for CPU)."
    import torch.distributed as td

    rank, size = world()
    if size > 1:
        if _PANEL_EXCHANGE == "allreduce":
            if rank != src:
                tensor.zero_()
            td.all_reduce(tensor, op=td.ReduceOp.SUM)
        else:
            td.broadcast(tensor, src=src)
    return tensor


def agree_on_status(code):
    """
    Make a per-rank status code collective: returns ``(lowest, highest)`` over all ranks, so that every
    rank raises (or not) together.  A failed pivot is only seen by the rank that owns the panel.
    """
    import torch

    rank, size = world()
    if size == 1:
        return int(code), int(code)
    import torch.distributed as td

    t = torch.tensor([-int(code), int(code)], dtype=torch.int64)
    if td.get_backend() == "nccl":
        t = t.cuda()
    td.all_reduce(t, op=td.ReduceOp.MAX)
    lo, hi = t.cpu().tolist()
    return -int(lo), int(hi)


def allreduce_sum_(tensor):
    "In-place sum over ranks of a torch tensor (NCCL for CUDA tensors, gloo for CPU)."
    import torch.distributed as td

    if world()[1] > 1:
        td.all_reduce(tensor, op=td.ReduceOp.SUM)
    return tensor


def allreduce_sum_numpy(array):
    "Sum a small host array over ranks (used for score tables)."
    import torch

    rank, size = world()
    if size == 1:
        return array
    import torch.distributed as td

    t = torch.from_numpy(np.ascontiguousarray(array, dtype=np.float64).copy())
    if td.get_backend() == "nccl":
        t = t.cuda()
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return t.cpu().numpy().reshape(np.shape(array))


def allgather_concat(local, total):
    """
    Concatenate each rank's contiguous slice (``shard_bounds`` order) of a 1-D
    float64 result of global length *total*; every rank gets the full array.
    """
    import torch

    rank, size = world()
    local = np.ascontiguousarray(local, dtype=np.float64)
    if size == 1:
        return local
    import torch.distributed as td

    nccl = td.get_backend() == "nccl"
    longest = max(shard_bounds(total, r, size)[1] - shard_bounds(total, r, size)[0] for r in range(size))
    buf = torch.zeros(longest, dtype=torch.float64)
    buf[: local.size] = torch.from_numpy(local)
    if nccl:
        buf = buf.cuda()
    parts = [torch.empty_like(buf) for _ in range(size)]
    td.all_gather(parts, buf)
    out = np.empty(int(total), dtype=np.float64)
    for r, part in enumerate(parts):
        a, b = shard_bounds(total, r, size)
        out[a:b] = part[: b - a].cpu().numpy()
    return out


def gather_device_shards(compute_local, total, components=1):
    """
    Point-sharded compute + gather WITHOUT host round trips (NCCL backend): allocates one CUDA tensor of shape
    ``(components, world, longest_shard)``, lets ``compute_local(views, start, stop)`` write this rank's results
    into its own slices (``views[c]`` is the contiguous device slice of component c), all-gathers IN PLACE over
    NVLink (``all_gather_into_tensor`` with the input aliasing the rank's slot) and copies every component to the
    host once.  Returns a list of *components* full-length float64 numpy arrays, identical on every rank.
    """
    import torch
    import torch.distributed as td

    rank, size = world()
    bounds = [shard_bounds(total, r, size) for r in range(size)]
    longest = max(b - a for a, b in bounds)
    dev = torch.device("cuda", torch.cuda.current_device())
    buf = torch.empty((components, size, longest), dtype=torch.float64, device=dev)
    a, b = bounds[rank]
    compute_local([buf[c, rank, : b - a] for c in range(components)], a, b)
    outs = []
    for c in range(components):
        td.all_gather_into_tensor(buf[c].view(-1), buf[c, rank])
        # Page-locked destination from torch's caching host allocator: every rank copies the WHOLE result (256 MB at
        # config 3's bench size), and eight pageable copies at once are bound by the driver's bounce buffers.  The
        # numpy array returned is a view that keeps the pinned block alive; it goes back to the cache when dropped.
        host = torch.empty(int(total), dtype=torch.float64, pin_memory=True)
        if all(bb - aa == longest for aa, bb in bounds):
            host.copy_(buf[c].view(-1))
        else:
            host.copy_(torch.cat([buf[c, r, : bb - aa] for r, (aa, bb) in enumerate(bounds)]))
        outs.append(host.numpy())
    return outs
