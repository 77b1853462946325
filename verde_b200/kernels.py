"""
Host-side mirror of the reference's inner numeric functions.

Same argument order and caller-allocates-output convention as the numba
functions they replace, so the estimator classes read like the reference's:

===========================  ==================================================
here                         reference
===========================  ==================================================
jacobian_b200                verde/spline.py:642-650  jacobian_numba
predict_b200                 verde/spline.py:629-639  predict_numba
predict_multi_b200           the same for several coefficient vectors on shared forces
                             (verde/vector.py:134-141 Vector.predict)
jacobian_2d_b200             verde/vector.py:461-475  jacobian_2d_numba
predict_2d_b200              verde/vector.py:443-458  predict_2d_numba
least_squares                verde/base/least_squares.py:17-71
fit_spline / fit_spline_2d   verde/spline.py:462-463 / verde/vector.py:287-288
                             (Jacobian build + least_squares fused on the GPU;
                             the Jacobian never exists in memory)
===========================  ==================================================

Everything executes in ``libverde_b200.so`` on the current CUDA device.
"""
import ctypes
import warnings
from warnings import warn

import numpy as np

from . import _cabi

LAST_FIT_INFO = {}


def _remember(info):
    LAST_FIT_INFO.clear()
    LAST_FIT_INFO.update(info.as_dict())


def jacobian_b200(east, north, force_east, force_north, mindist, jac):
    "Fill the (n, m) row-major float64 array *jac* with the spline Green's functions."
    lib = _cabi.require_device()
    e, n, fe, fn = (_cabi.f64(a) for a in (east, north, force_east, force_north))
    if jac.dtype != np.float64 or not jac.flags.c_contiguous:
        out = np.empty((e.size, fe.size), dtype=np.float64)
    else:
        out = jac
    _cabi.check(lib.vb200_jacobian(_cabi.ptr(e), _cabi.ptr(n), e.size, _cabi.ptr(fe), _cabi.ptr(fn),
                                   fe.size, float(mindist), _cabi.ptr(out)), "jacobian")
    if out is not jac:
        jac[:] = out
    return jac


def predict_b200(east, north, force_east, force_north, mindist, forces, result):
    "result[i] = sum_j g(east[i]-force_east[j], north[i]-force_north[j]) * forces[j]"
    lib = _cabi.require_device()
    e, n, fe, fn, f = (_cabi.f64(a) for a in (east, north, force_east, force_north, forces))
    direct = result.dtype == np.float64 and result.flags.c_contiguous
    out = result if direct else np.empty(e.size, dtype=np.float64)
    _cabi.check(lib.vb200_predict(_cabi.ptr(e), _cabi.ptr(n), e.size, _cabi.ptr(fe), _cabi.ptr(fn),
                                  _cabi.ptr(f), f.size, float(mindist), _cabi.ptr(out)), "predict")
    if not direct:
        result[:] = out
    return result


def predict_multi_b200(east, north, force_east, force_north, mindist, forces, result):
    """
    ``result[k]`` = predict with coefficient vector ``forces[k]`` on the shared forces, for all k at once
    (one Green's evaluation per pair instead of one per component).  forces: (k, m); result: (k, n) float64.
    """
    lib = _cabi.require_device()
    e, n, fe, fn = (_cabi.f64(a) for a in (east, north, force_east, force_north))
    f = np.ascontiguousarray(forces, dtype=np.float64)
    if f.ndim != 2 or f.shape[1] != fe.size:
        raise ValueError("forces must have shape (components, n_forces)")
    direct = result.dtype == np.float64 and result.flags.c_contiguous and result.shape == (f.shape[0], e.size)
    out = result if direct else np.empty((f.shape[0], e.size), dtype=np.float64)
    _cabi.check(lib.vb200_predict_multi(_cabi.ptr(e), _cabi.ptr(n), e.size, _cabi.ptr(fe), _cabi.ptr(fn), _cabi.ptr(f),
                                        fe.size, f.shape[0], float(mindist), _cabi.ptr(out)), "predict_multi")
    if not direct:
        result[...] = out
    return result


def jacobian_2d_b200(east, north, force_east, force_north, mindist, poisson, jac):
    "Fill the (2n, 2m) block Jacobian [[ee, ne], [ne, nn]] of the elastic spline."
    lib = _cabi.require_device()
    e, n, fe, fn = (_cabi.f64(a) for a in (east, north, force_east, force_north))
    direct = jac.dtype == np.float64 and jac.flags.c_contiguous
    out = jac if direct else np.empty((2 * e.size, 2 * fe.size), dtype=np.float64)
    _cabi.check(lib.vb200_jacobian_2d(_cabi.ptr(e), _cabi.ptr(n), e.size, _cabi.ptr(fe), _cabi.ptr(fn),
                                      fe.size, float(mindist), float(poisson), _cabi.ptr(out)), "jacobian_2d")
    if not direct:
        jac[:] = out
    return jac


def predict_2d_b200(east, north, force_east, force_north, mindist, poisson, forces, vec_east, vec_north):
    "Coupled 2-component prediction; *forces* = east forces then north forces."
    lib = _cabi.require_device()
    e, n, fe, fn, f = (_cabi.f64(a) for a in (east, north, force_east, force_north, forces))
    if f.size != 2 * fe.size:
        raise ValueError("forces must have 2 * n_forces elements ({} given for {} forces)".format(f.size, fe.size))
    direct = all(a.dtype == np.float64 and a.flags.c_contiguous for a in (vec_east, vec_north))
    oe = vec_east if direct else np.empty(e.size, dtype=np.float64)
    on = vec_north if direct else np.empty(e.size, dtype=np.float64)
    _cabi.check(lib.vb200_predict_2d(_cabi.ptr(e), _cabi.ptr(n), e.size, _cabi.ptr(fe), _cabi.ptr(fn),
                                     _cabi.ptr(f), fe.size, float(mindist), float(poisson),
                                     _cabi.ptr(oe), _cabi.ptr(on)), "predict_2d")
    if not direct:
        vec_east[:] = oe
        vec_north[:] = on
    return vec_east, vec_north


# scipy.linalg.lstsq's ``cond`` in the damping=None branch: scikit-learn >= 1.9 passes ``LinearRegression.tol``,
# whose default is 1e-6 (sklearn/linear_model/_base.py; the goldens in tests/golden/damping_none.npz were made
# with it).  Older scikit-learn passed None (machine precision); set this to reproduce those versions.
LSTSQ_COND = 1e-6


def _require_damping(damping):
    "The Cholesky (normal-equation) entry points need a positive damping; damping=None goes through ``*_lstsq``."
    if damping is None:
        raise ValueError("this entry point solves the damped normal equations: damping must be a positive float")
    if not damping > 0:
        raise ValueError("damping must be positive, got {!r}".format(damping))


def _warn_underdetermined(ndata, nparams):
    # verde/base/least_squares.py:57-62 (same text)
    if ndata < nparams:
        warn("Under-determined problem detected (ndata, nparams)={}.".format((ndata, nparams)))


def least_squares(jacobian, data, weights, damping=None, copy_jacobian=False):  # noqa: U100
    """
    Weighted least squares on an explicit Jacobian, on the GPU (verde/base/least_squares.py:17-71).

    Same algebra as the reference: column scaling by the population standard deviation, then
    ``(Js^T W Js + damping I) c = Js^T W d`` by Cholesky for a positive *damping*, or -- ``damping=None`` -- the
    minimum-norm solution of ``sqrt(W) Js c = sqrt(W) d`` with singular values below ``LSTSQ_COND * sigma_max``
    dropped (what scikit-learn's ``LinearRegression`` computes); parameters unscaled on return.  The Jacobian
    is never modified, so *copy_jacobian* is accepted for signature compatibility only.
    """
    lib = _cabi.require_device()
    jacobian = np.ascontiguousarray(jacobian, dtype=np.float64)
    if jacobian.ndim != 2:
        raise ValueError("jacobian must be 2-D")
    _warn_underdetermined(*jacobian.shape)
    if damping is not None:
        _require_damping(damping)
    d = _cabi.f64(data)
    if d.size != jacobian.shape[0]:
        raise ValueError("data size {} does not match the Jacobian rows {}".format(d.size, jacobian.shape[0]))
    w = None if weights is None else _cabi.f64(weights)
    params = np.empty(jacobian.shape[1], dtype=np.float64)
    info = _cabi.FitInfo()
    if damping is None:
        rc = lib.vb200_least_squares_lstsq(_cabi.ptr(jacobian), jacobian.shape[0], jacobian.shape[1], _cabi.ptr(d),
                                           _cabi.ptr(w), float(LSTSQ_COND), _cabi.ptr(params), ctypes.byref(info))
    else:
        rc = lib.vb200_least_squares(_cabi.ptr(jacobian), jacobian.shape[0], jacobian.shape[1], _cabi.ptr(d),
                                     _cabi.ptr(w), float(damping), _cabi.ptr(params), ctypes.byref(info))
    _remember(info)
    _cabi.check(rc, "least_squares")
    return params


def fit_spline(east, north, data, weights, force_east, force_north, mindist, damping):
    "Forces of the scalar spline; Jacobian build, Gram, Cholesky and solve all on the GPU."
    lib = _cabi.require_device()
    e, n, d, fe, fn = (_cabi.f64(a) for a in (east, north, data, force_east, force_north))
    _warn_underdetermined(e.size, fe.size)
    if damping is not None:
        _require_damping(damping)
    w = None if weights is None else _cabi.f64(weights)
    force = np.empty(fe.size, dtype=np.float64)
    info = _cabi.FitInfo()
    if damping is None:
        rc = lib.vb200_spline_fit_lstsq(_cabi.ptr(e), _cabi.ptr(n), _cabi.ptr(d), _cabi.ptr(w), e.size, _cabi.ptr(fe),
                                        _cabi.ptr(fn), fe.size, float(mindist), float(LSTSQ_COND), _cabi.ptr(force),
                                        ctypes.byref(info))
    else:
        rc = lib.vb200_spline_fit(_cabi.ptr(e), _cabi.ptr(n), _cabi.ptr(d), _cabi.ptr(w), e.size, _cabi.ptr(fe),
                                  _cabi.ptr(fn), fe.size, float(mindist), float(damping), _cabi.ptr(force),
                                  ctypes.byref(info))
    _remember(info)
    _cabi.check(rc, "Spline.fit")
    return force


def fit_spline_2d(east, north, data, weights, force_east, force_north, mindist, poisson, damping):
    "Forces (east then north) of the 2-D elastic spline; *data*/*weights* stacked east-then-north."
    lib = _cabi.require_device()
    e, n, d, fe, fn = (_cabi.f64(a) for a in (east, north, data, force_east, force_north))
    _warn_underdetermined(2 * e.size, 2 * fe.size)
    if damping is not None:
        _require_damping(damping)
    w = None if weights is None else _cabi.f64(weights)
    force = np.empty(2 * fe.size, dtype=np.float64)
    info = _cabi.FitInfo()
    if damping is None:
        rc = lib.vb200_spline_fit_2d_lstsq(_cabi.ptr(e), _cabi.ptr(n), _cabi.ptr(d), _cabi.ptr(w), e.size, _cabi.ptr(fe),
                                           _cabi.ptr(fn), fe.size, float(mindist), float(poisson), float(LSTSQ_COND),
                                           _cabi.ptr(force), ctypes.byref(info))
    else:
        rc = lib.vb200_spline_fit_2d(_cabi.ptr(e), _cabi.ptr(n), _cabi.ptr(d), _cabi.ptr(w), e.size, _cabi.ptr(fe),
                                     _cabi.ptr(fn), fe.size, float(mindist), float(poisson), float(damping),
                                     _cabi.ptr(force), ctypes.byref(info))
    _remember(info)
    _cabi.check(rc, "VectorSpline2D.fit")
    return force


class GramSystem:
    """
    Staged fit on device-resident normal equations (C ABI ``vb200_gram_*``).

    Holds G = J^T W J (tile-packed lower triangle) and the statistics vector in
    HBM as torch CUDA tensors so that (a) ranks can all-reduce them over NCCL
    between ``accumulate`` and ``solve`` and (b) SplineCV can re-solve one Gram
    matrix for many dampings (``keep=True``).
    """

    def __init__(self, cols, peer_exchange=None):
        import torch

        self._lib = _cabi.require_device()
        self.cols = int(cols)
        from . import dist

        dev = torch.device("cuda", torch.cuda.current_device())
        count = self._lib.vb200_gram_doubles(self.cols)
        self._shared = None
        if peer_exchange is None:
            # only systems that will be factored panel-cyclically need exportable memory (identical on every rank:
            # the rule depends on sizes alone)
            peer_exchange = (dist.world()[1] > 1 and dist.backend() == "nccl" and dist.panel_pipeline() == "peer"
                             and dist.use_distributed_cholesky((self.cols + 127) // 128, dist.panel_tiles()))
        if peer_exchange:
            # the matrix and the per-panel flag words live in CUDA-IPC-exportable allocations: the other ranks map
            # them (PeerWindow, at the first panel-cyclic factorisation) and push factored panels straight into them
            try:
                self._shared = {"gram": dist.SharedBuffer(count, torch.float64), "flags": dist.SharedBuffer(4096, torch.int64)}
            except _cabi.EngineError as exc:
                # CUDA IPC is not available here: plain memory; the first factorisation makes the fall-back to the
                # NCCL exchange collective (this rank then reports "no shareable buffers" to its peers)
                warnings.warn("peer-memory panel exchange unavailable ({}): using ncclBroadcast".format(exc))
                self._shared = None
        if self._shared is not None:
            self.gram = self._shared["gram"].tensor
            self._flags = self._shared["flags"].tensor
            self._flags.zero_()
            self._epoch_src = torch.zeros(1, dtype=torch.int64, device=dev)
        else:
            self.gram = torch.empty(count, dtype=torch.float64, device=dev)
        self._wants_peer = bool(peer_exchange)  # was built for the peer exchange (even if the allocation failed)
        self.stats = torch.empty(self._lib.vb200_stats_doubles(self.cols), dtype=torch.float64, device=dev)
        self.info = _cabi.FitInfo()
        self._torch = torch
        self._window = None  # dist.PeerWindow onto the other ranks' (gram, flags), made by the first peer-mode factorisation

    def release(self):
        """
        Give the CUDA-IPC allocations back.  COLLECTIVE when peers have mapped them (every rank calls it at the same
        point: unmap the peers' buffers, barrier, free the own ones); local otherwise.  The object is unusable after.
        """
        from . import dist

        if self._shared is None:
            return
        if self._window is not None:
            self._torch.cuda.synchronize()
            self._window.close()
            self._window = None
            import torch.distributed as td

            td.barrier()
        self.gram = self._flags = None
        for buf in self._shared.values():
            buf.free()
        self._shared = None

    def __del__(self):
        # never mapped by a peer (no panel-cyclic factorisation ran on it): the exported memory can go now; a mapped
        # one is only freed by the collective release() -- dropping it silently would strand the peers' mappings
        try:
            if getattr(self, "_shared", None) is not None and getattr(self, "_window", None) is None:
                self.gram = self._flags = None
                for buf in self._shared.values():
                    buf.free()
                self._shared = None
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def accumulate(self, kind, east, north, data, weights, force_east, force_north, mindist,
                   poisson=0.0, accumulate=False):
        e, n, d, fe, fn = (_cabi.f64(a) for a in (east, north, data, force_east, force_north))
        w = None if weights is None else _cabi.f64(weights)
        self._torch.cuda.synchronize()
        rc = self._lib.vb200_gram_accumulate(
            int(kind), _cabi.ptr(e), _cabi.ptr(n), _cabi.ptr(d), _cabi.ptr(w), e.size, _cabi.ptr(fe),
            _cabi.ptr(fn), fe.size, float(mindist), float(poisson), _cabi.ptr(self.gram),
            _cabi.ptr(self.stats), int(bool(accumulate)), ctypes.byref(self.info))
        _cabi.check(rc, "gram_accumulate")
        return self

    def allreduce(self):
        "Sum the partial normal equations over ranks (NCCL all-reduce over NVLink)."
        from . import dist

        dist.allreduce_sum_(self.gram)
        dist.allreduce_sum_(self.stats)
        self._torch.cuda.synchronize()
        return self

    def solve(self, total_rows, damping, keep=False):
        _require_damping(damping)
        params = np.empty(self.cols, dtype=np.float64)
        self._torch.cuda.synchronize()
        rc = self._lib.vb200_gram_solve(_cabi.ptr(self.gram), _cabi.ptr(self.stats), self.cols,
                                        int(total_rows), float(damping), int(bool(keep)),
                                        _cabi.ptr(params), ctypes.byref(self.info))
        _remember(self.info)
        _cabi.check(rc, "gram_solve")
        return params

    def solve_multi(self, total_rows, dampings):
        """
        Parameters for every damping of a sweep, shape (len(dampings), cols).  The Gram matrix is
        kept; all damped systems are factored and solved side by side on the device
        (``vb200_gram_solve_multi``).  A non-positive-definite system raises LinAlgError.
        """
        damp = np.ascontiguousarray([float(d) for d in dampings], dtype=np.float64)
        for d in damp:
            _require_damping(d)
        params = np.empty((damp.size, self.cols), dtype=np.float64)
        pivots = np.zeros(damp.size, dtype=np.int32)
        self._torch.cuda.synchronize()
        rc = self._lib.vb200_gram_solve_multi(_cabi.ptr(self.gram), _cabi.ptr(self.stats), self.cols,
                                              int(total_rows), _cabi.ptr(damp), int(damp.size),
                                              _cabi.ptr(params), _cabi.ptr(pivots), ctypes.byref(self.info))
        _remember(self.info)
        _cabi.check(rc, "gram_solve_multi")
        return params

    def solve_distributed(self, total_rows, damping, panel_tiles=8, out=None, phase_times=None):
        """
        Like ``solve`` on N > 1 ranks that all hold the all-reduced Gram matrix, but the m^3/3
        flops of the Cholesky are shared: outer panel p is factored by rank ``p % N`` and broadcast
        (NCCL), every rank applies it to the block columns it owns.  All ranks return the same
        parameters.  Everything between ``vb200_chol_begin`` and ``vb200_chol_finish`` is enqueued
        without host synchronisation; NCCL orders itself against the library stream (the legacy
        default stream, torch's current stream).
        """
        from . import dist

        _require_damping(damping)
        rank, size = dist.world()
        lib, cols = self._lib, self.cols
        block_rows = (cols + 127) // 128
        schedule = dist.panel_cyclic_schedule(block_rows, panel_tiles, size)
        gram_ptr = _cabi.ptr(self.gram)
        _cabi.check(lib.vb200_chol_begin(gram_ptr, _cabi.ptr(self.stats), cols, int(total_rows), float(damping)),
                    "chol_begin")
        offset, count = ctypes.c_int64(), ctypes.c_int64()
        marks = []  # (phase, event) pairs on the library stream when the caller asked for a breakdown

        def mark(phase):
            if phase_times is not None:
                event = self._torch.cuda.Event(enable_timing=True)
                event.record()
                marks.append((phase, event))

        mark("start")
        # A rank-local failure (bad argument, CUDA error) must not leave the other ranks waiting in the next
        # broadcast: the first failing status is remembered, this rank stops issuing its own kernels but keeps
        # taking part in every collective, and all ranks raise together after ``agree_on_status``.
        status = 0
        torch = self._torch

        def update_owned(k0, width, panels):
            "apply panel [k0, k0 + width) to the later panels this rank owns: ONE launch over all their block columns"
            owned = dist.owned_ranges(panels, rank)
            if not owned:
                return 0
            if not dist.merged_updates():
                for lo, hi in owned:
                    rc = lib.vb200_chol_update(gram_ptr, cols, k0, width, lo, hi)
                    if rc != 0:
                        return rc
                return 0
            los = (ctypes.c_int32 * len(owned))(*[lo for lo, _ in owned])
            his = (ctypes.c_int32 * len(owned))(*[hi for _, hi in owned])
            return lib.vb200_chol_update_ranges(gram_ptr, cols, k0, width, len(owned), los, his)

        pipeline = dist.panel_pipeline() if len(schedule) > 1 else "serial"
        if pipeline == "peer" and not self._wants_peer:
            pipeline = "lookahead"  # this matrix was not allocated for the peer exchange (the same on every rank)
        if pipeline == "peer" and self._window is None:
            # first factorisation on this matrix: map the peers' buffers.  Whether that worked is agreed on by all
            # ranks (a failed mapping or allocation on ONE rank switches EVERY rank to the NCCL exchange, for good).
            window = dist.PeerWindow(self._shared or {})
            lowest, _ = dist.agree_on_status(-1 if window.error else 0)
            if lowest < 0:
                if window.error:
                    warnings.warn("peer-memory panel exchange unavailable ({}): using ncclBroadcast".format(window.error))
                window.close()  # whatever this rank did map must not outlive the decision
                self._wants_peer = False
                pipeline = "lookahead"
            else:
                self._window = window
        if pipeline == "peer":
            # Look-ahead order + peer-memory pushes: see dist.set_panel_pipeline.  flags[p] on every rank is written
            # by the owner of panel p (the fit's epoch) after the panel itself has landed in this rank's matrix.
            ready_base = self._flags.numel() - 1024  # flags[ready_base + r]: rank r's matrix is ready to receive
            if len(schedule) > ready_base or size > 1024:
                raise ValueError("too many outer panels for the peer exchange ({})".format(len(schedule)))
            epoch = dist.next_epoch()
            main = torch.cuda.current_stream()
            comm = dist.comm_stream()
            self._epoch_src.fill_(epoch)
            seeded = torch.cuda.Event()
            seeded.record(main)
            comm.wait_event(seeded)
            flag_ptr = self._flags.data_ptr()
            comm_handle = ctypes.c_void_p(comm.cuda_stream)
            my_gram, epoch_ptr = self.gram.data_ptr(), self._epoch_src.data_ptr()
            # Ready handshake.  vb200_chol_begin rewrites the WHOLE matrix in place (scaling + damping); a panel
            # pushed by a faster peer before that pass has finished here would be scaled again or overwritten.  So
            # every rank tells its peers "my matrix is in place" (ordered after chol_begin through `seeded`), and
            # parks its copy stream until every peer has said the same, before its first push.
            peers = [r for r in range(size) if r != rank] if dist._READY_HANDSHAKE else []
            for r in peers:
                lib.vb200_peer_copy_async(ctypes.c_void_p(self._window.ptr(r, "flags") + 8 * (ready_base + rank)),
                                          ctypes.c_void_p(epoch_ptr), 8, comm_handle)
            for r in peers:
                rc = lib.vb200_stream_wait_flag(ctypes.c_void_p(flag_ptr + 8 * (ready_base + r)), epoch, comm_handle)
                status = status or rc

            def push(p, k0, width, ready):
                "owner: copy panel p and then its flag into every peer, the owner of panel p+1 first"
                lib.vb200_chol_panel_range(cols, k0, width, ctypes.byref(offset), ctypes.byref(count))
                lo, nbytes = 8 * offset.value, 8 * count.value
                first = schedule[p + 1][2] if p + 1 < len(schedule) else (rank + 1) % size
                comm.wait_event(ready)
                for r in ((first + i) % size for i in range(size)):
                    if r != rank:
                        lib.vb200_peer_copy_async(ctypes.c_void_p(self._window.ptr(r, "gram") + lo),
                                                  ctypes.c_void_p(my_gram + lo), nbytes, comm_handle)
                        lib.vb200_peer_copy_async(ctypes.c_void_p(self._window.ptr(r, "flags") + 8 * p),
                                                  ctypes.c_void_p(epoch_ptr), 8, comm_handle)

            ready = torch.cuda.Event()
            if rank == schedule[0][2]:
                status = lib.vb200_chol_panel(gram_ptr, cols, schedule[0][0], schedule[0][1])
            ready.record(main)
            mark("panel")
            for p, (k0, width, owner) in enumerate(schedule):
                if rank == owner:
                    push(p, k0, width, ready)
                elif status == 0:
                    status = lib.vb200_chol_wait_panel(ctypes.c_void_p(flag_ptr + 8 * p), epoch)
                mark("broadcast")
                later = schedule[p + 1:]
                if later and later[0][2] == rank:
                    q0, qwidth, _ = later[0]
                    if status == 0:
                        status = lib.vb200_chol_update(gram_ptr, cols, k0, width, q0, q0 + qwidth)
                    if status == 0:
                        status = lib.vb200_chol_panel(gram_ptr, cols, q0, qwidth)
                    ready = torch.cuda.Event()
                    ready.record(main)
                    later = later[1:]
                    mark("panel")
                if status == 0:
                    status = update_owned(k0, width, later)
                mark("update")
            done = torch.cuda.Event()
            done.record(comm)
            main.wait_event(done)  # this rank's pushes have left before anything overwrites the matrix
        elif pipeline == "lookahead":
            # Look-ahead: the panel broadcasts run on a high-priority side stream beside the DMMA updates of the main
            # (library) stream, and the owner of panel p+1 updates and factors THAT panel first, as soon as panel p has
            # arrived, so that its broadcast is on the wire while everybody is still busy with panel p's updates.
            main = torch.cuda.current_stream()
            comm = dist.comm_stream()
            last = None
            ready = torch.cuda.Event()
            if rank == schedule[0][2]:
                status = lib.vb200_chol_panel(gram_ptr, cols, schedule[0][0], schedule[0][1])
            ready.record(main)
            # EVERY rank orders its side stream behind vb200_chol_begin (and, on the owner, the first panel): that call
            # rewrites the whole matrix in place, and a panel received before it has finished here would be scaled
            # again (found by the skew test of tools/run_dist_pipeline.py; invisible between barriers)
            comm.wait_event(ready)
            mark("panel")
            for p, (k0, width, owner) in enumerate(schedule):
                lib.vb200_chol_panel_range(cols, k0, width, ctypes.byref(offset), ctypes.byref(count))
                if rank == owner and p > 0:
                    comm.wait_event(ready)  # the panel is factored
                with torch.cuda.stream(comm):
                    dist.broadcast_(self.gram[offset.value:offset.value + count.value], owner)
                    last = torch.cuda.Event()
                    last.record(comm)
                if rank != owner:
                    main.wait_event(last)  # the panel has arrived
                mark("broadcast")
                later = schedule[p + 1:]
                if later and later[0][2] == rank:
                    q0, qwidth, _ = later[0]
                    if status == 0:
                        status = lib.vb200_chol_update(gram_ptr, cols, k0, width, q0, q0 + qwidth)
                    if status == 0:
                        status = lib.vb200_chol_panel(gram_ptr, cols, q0, qwidth)
                    ready = torch.cuda.Event()
                    ready.record(main)
                    later = later[1:]
                    mark("panel")
                if status == 0:
                    status = update_owned(k0, width, later)
                mark("update")
            main.wait_event(last)
        else:
            for p, (k0, width, owner) in enumerate(schedule):
                if rank == owner and status == 0:
                    status = lib.vb200_chol_panel(gram_ptr, cols, k0, width)
                mark("panel")
                lib.vb200_chol_panel_range(cols, k0, width, ctypes.byref(offset), ctypes.byref(count))
                dist.broadcast_(self.gram[offset.value:offset.value + count.value], owner)
                mark("broadcast")
                for q0, qwidth, qowner in schedule[p + 1:]:
                    if qowner == rank and status == 0:
                        status = lib.vb200_chol_update(gram_ptr, cols, k0, width, q0, q0 + qwidth)
                mark("update")
        params = np.empty(cols, dtype=np.float64) if out is None else out  # host array or CUDA tensor
        message = _cabi.load().vb200_last_error().decode("utf-8", "replace") if status != 0 else ""
        rc = lib.vb200_chol_finish(gram_ptr, cols, _cabi.ptr(params), ctypes.byref(self.info))
        if status != 0:
            rc = status
        _remember(self.info)
        # a non-positive pivot is only recorded on the rank that owns its panel: agree before raising
        if phase_times is not None:  # device time between consecutive marks, summed per phase (ms)
            for (_, before), (phase, after) in zip(marks, marks[1:]):
                phase_times[phase] = phase_times.get(phase, 0.0) + before.elapsed_time(after)
        lowest, highest = dist.agree_on_status(rc)
        if rc == 0 and (lowest < 0 or highest > 0):
            other = lowest if lowest < 0 else highest
            if other > 0:
                raise np.linalg.LinAlgError(
                    "chol_finish: the damped normal matrix is not positive definite (pivot {}, found by another rank)".format(other))
            raise _cabi.EngineError("chol_finish: another rank failed with status {}".format(other))
        if status != 0:
            _cabi.raise_status(rc, "distributed Cholesky", message)
        _cabi.check(rc, "chol_finish")
        return params

    def backward_error(self, kind, east, north, data, weights, force_east, force_north, mindist, poisson,
                       force, total_rows, damping):
        """
        Normwise backward error of *force* for the (row-sharded) system whose statistics this object holds:
        every rank evaluates ``J^T W (d - J force)`` over ITS rows matrix-free (``vb200_normal_residual``), the m
        partial gradients are summed over ranks, ``vb200_backward_error`` forms the ratio.  Stored in
        ``info.backward_err`` (and returned); needs ``info.lambda_max`` from the solve.
        """
        from . import dist

        def as_input(a):  # host arrays are made contiguous float64; CUDA tensors pass through as device pointers
            return a if hasattr(a, "data_ptr") else _cabi.f64(a)

        e, n, d, fe, fn, f = (as_input(a) for a in (east, north, data, force_east, force_north, force))
        w = None if weights is None else as_input(weights)
        grad = self._torch.empty(self.cols, dtype=self._torch.float64, device=self.gram.device)
        rc = self._lib.vb200_normal_residual(int(kind), _cabi.ptr(e), _cabi.ptr(n), _cabi.ptr(d), _cabi.ptr(w),
                                             int(e.numel() if hasattr(e, "numel") else e.size),
                                             _cabi.ptr(fe), _cabi.ptr(fn),
                                             int(fe.numel() if hasattr(fe, "numel") else fe.size), float(mindist),
                                             float(poisson), _cabi.ptr(f), _cabi.ptr(grad))
        lowest, highest = dist.agree_on_status(rc)
        _cabi.check(rc if rc != 0 else (lowest if lowest < 0 else highest), "normal_residual")
        dist.allreduce_sum_(grad)
        self._torch.cuda.synchronize()
        eta = ctypes.c_double(-1.0)
        _cabi.check(self._lib.vb200_backward_error(_cabi.ptr(grad), _cabi.ptr(self.stats), self.cols, int(total_rows),
                                                   float(damping), _cabi.ptr(f), float(self.info.lambda_max),
                                                   ctypes.byref(eta)), "backward_error")
        self.info.backward_err = eta.value
        _remember(self.info)
        return eta.value

    def factor(self, total_rows, damping):
        "Scale, damp and factor in place (the Gram matrix becomes L); keeps 1/s for later solves."
        _require_damping(damping)
        self.inv_scale = self._torch.empty(self._lib.vb200_padded_cols(self.cols), dtype=self._torch.float64,
                                           device=self.gram.device)
        self._torch.cuda.synchronize()
        rc = self._lib.vb200_gram_factor(_cabi.ptr(self.gram), _cabi.ptr(self.stats), self.cols, int(total_rows),
                                         float(damping), _cabi.ptr(self.inv_scale), ctypes.byref(self.info))
        _remember(self.info)
        _cabi.check(rc, "gram_factor")
        return self

    def rhs(self):
        "Device view of J^T W d of the data passed to ``accumulate`` (unscaled)."
        return self.stats[: self.cols]

    def solve_rhs(self, rhs):
        "Parameters for one unscaled right-hand side J^T W d (numpy array or CUDA tensor); needs ``factor``."
        if not hasattr(rhs, "data_ptr"):
            rhs = _cabi.f64(rhs)
        params = np.empty(self.cols, dtype=np.float64)
        self._torch.cuda.synchronize()
        rc = self._lib.vb200_chol_solve(_cabi.ptr(self.gram), _cabi.ptr(self.inv_scale), self.cols, _cabi.ptr(rhs),
                                        _cabi.ptr(params), ctypes.byref(self.info))
        _remember(self.info)
        _cabi.check(rc, "chol_solve")
        return params


def jacobian_transpose_apply(east, north, force_east, force_north, mindist, v):
    "J^T v for the scalar spline without forming J (exact Green's function)."
    lib = _cabi.require_device()
    e, n, fe, fn, vv = (_cabi.f64(a) for a in (east, north, force_east, force_north, v))
    if vv.size != e.size:
        raise ValueError("v must have one entry per data point")
    out = np.empty(fe.size, dtype=np.float64)
    _cabi.check(lib.vb200_jacobian_transpose_apply(_cabi.ptr(e), _cabi.ptr(n), e.size, _cabi.ptr(fe), _cabi.ptr(fn),
                                                   fe.size, float(mindist), _cabi.ptr(vv), _cabi.ptr(out)),
                "jacobian_transpose_apply")
    return out


_SHARED_SYSTEM = {}


def _shared_system(cols):
    """
    The GramSystem of the row-sharded fits, kept between fits of the same size (one at a time): its matrix is the
    10 GB allocation of the fit, and in peer mode it owns the CUDA-IPC window onto the other ranks' matrices, whose
    set-up (handle exchange + mapping) should not be paid by every fit.
    """
    import torch

    key = (torch.cuda.current_device(), int(cols))
    system = _SHARED_SYSTEM.get(key)
    if system is None:
        # a different size: let the old matrix go first (collective when its peers mapped it -- every rank is here
        # with the same sizes in the same order)
        for old in _SHARED_SYSTEM.values():
            old.release()
        _SHARED_SYSTEM.clear()
        system = _SHARED_SYSTEM[key] = GramSystem(cols)
    system.info = _cabi.FitInfo()
    return system


def fit_sharded(kind, east, north, data, weights, force_east, force_north, mindist, poisson, damping):
    """
    Row-sharded fit: every rank holds the full inputs, builds the normal
    equations of its contiguous row range, one all-reduce, replicated solve.
    *data*/*weights* are per-component tuples for kind 1, plain arrays for kind 0.
    """
    from . import dist

    if damping is None:
        # the truncated-SVD branch has no row-additive form: every rank solves the whole (modest) system itself,
        # deterministically, so all ranks hold bit-identical forces without a collective
        if kind == 1:
            stacked_w = None if weights is None else np.concatenate([_cabi.f64(c) for c in weights])
            return fit_spline_2d(east, north, np.concatenate([_cabi.f64(c) for c in data]), stacked_w, force_east,
                                 force_north, mindist, poisson, None)
        return fit_spline(east, north, data, weights, force_east, force_north, mindist, None)
    e, n, fe, fn = (_cabi.f64(a) for a in (east, north, force_east, force_north))
    a, b = dist.shard_bounds(e.size)
    if kind == 1:
        d = np.concatenate([_cabi.f64(c)[a:b] for c in data])
        w = None if weights is None else np.concatenate([_cabi.f64(c)[a:b] for c in weights])
        rows, cols = 2 * e.size, 2 * fe.size
    else:
        d = _cabi.f64(data)[a:b]
        w = None if weights is None else _cabi.f64(weights)[a:b]
        rows, cols = e.size, fe.size
    _warn_underdetermined(rows, cols)
    _require_damping(damping)
    # a rank whose accumulation fails (out of memory, bad input) must not strand the others in the all-reduce
    system, failure = None, None
    try:
        system = _shared_system(cols)
        system.accumulate(kind, e[a:b], n[a:b], d, w, fe, fn, mindist, poisson)
    except (MemoryError, ValueError, _cabi.EngineError) as exc:  # noqa: PERF203
        failure = exc
    lowest, _ = dist.agree_on_status(-1 if failure is not None else 0)
    if failure is not None:
        raise failure
    if lowest < 0:
        raise _cabi.EngineError("row-sharded fit: the Gram accumulation failed on another rank")
    system.allreduce()
    if dist.use_distributed_cholesky((cols + 127) // 128, dist.panel_tiles()):
        force = system.solve_distributed(rows, damping, panel_tiles=dist.panel_tiles())
    else:
        force = system.solve(rows, damping)
    if system.info.lambda_max > 0:  # diagnostics on: backward error of the sharded system
        system.backward_error(kind, e[a:b], n[a:b], d, w, fe, fn, mindist, poisson, force, rows, damping)
    return force


def predict_multi_sharded(east, north, force_east, force_north, mindist, forces):
    """
    ``predict_multi_b200`` with the points sharded over the ranks: each rank evaluates its contiguous slice for all
    components at once (one Green's evaluation per pair), one gather per component; returns (components, n) float64,
    identical on every rank.
    """
    from . import dist

    e, n = _cabi.f64(east), _cabi.f64(north)
    f = np.ascontiguousarray(forces, dtype=np.float64)
    comps = f.shape[0]
    if dist.backend() == "nccl":
        import torch

        lib = _cabi.require_device()
        fe, fn = _cabi.f64(force_east), _cabi.f64(force_north)

        def compute(views, a, b):
            if b <= a:
                return
            # the kernel writes a contiguous (components, b - a) matrix; the gather buffer keeps one slot per rank
            local = torch.empty((comps, b - a), dtype=torch.float64, device=views[0].device)
            _cabi.check(lib.vb200_predict_multi(_cabi.ptr(e[a:b]), _cabi.ptr(n[a:b]), b - a, _cabi.ptr(fe), _cabi.ptr(fn),
                                                _cabi.ptr(f), fe.size, comps, float(mindist), _cabi.ptr(local)),
                        "sharded predict_multi")
            for c in range(comps):
                views[c].copy_(local[c])

        return np.stack(dist.gather_device_shards(compute, e.size, components=comps))
    a, b = dist.shard_bounds(e.size)
    local = np.empty((comps, b - a), dtype=np.float64)
    if b > a:
        predict_multi_b200(e[a:b], n[a:b], force_east, force_north, mindist, f, local)
    return np.stack([dist.allgather_concat(local[c], e.size) for c in range(comps)])


def predict_sharded(east, north, force_east, force_north, mindist, forces, poisson=None):
    "Point-sharded predict + all-gather; returns full-length 1-D result(s)."
    from . import dist

    e, n = _cabi.f64(east), _cabi.f64(north)
    if dist.backend() == "nccl":
        # results stay on the device until the gather is done: predict writes into this rank's slot of the
        # gathered tensor, one in-place all-gather over NVLink, one device-to-host copy per component
        lib = _cabi.require_device()
        fe, fn, f = (_cabi.f64(x) for x in (force_east, force_north, forces))

        def compute(views, a, b):
            if b <= a:
                return
            if poisson is None:
                rc = lib.vb200_predict(_cabi.ptr(e[a:b]), _cabi.ptr(n[a:b]), b - a, _cabi.ptr(fe), _cabi.ptr(fn),
                                       _cabi.ptr(f), f.size, float(mindist), _cabi.ptr(views[0]))
            else:
                rc = lib.vb200_predict_2d(_cabi.ptr(e[a:b]), _cabi.ptr(n[a:b]), b - a, _cabi.ptr(fe), _cabi.ptr(fn),
                                          _cabi.ptr(f), fe.size, float(mindist), float(poisson),
                                          _cabi.ptr(views[0]), _cabi.ptr(views[1]))
            _cabi.check(rc, "sharded predict")

        outs = dist.gather_device_shards(compute, e.size, components=1 if poisson is None else 2)
        return outs[0] if poisson is None else (outs[0], outs[1])
    a, b = dist.shard_bounds(e.size)
    if poisson is None:
        local = np.empty(b - a, dtype=np.float64)
        predict_b200(e[a:b], n[a:b], force_east, force_north, mindist, forces, local)
        return dist.allgather_concat(local, e.size)
    le = np.empty(b - a, dtype=np.float64)
    ln = np.empty(b - a, dtype=np.float64)
    predict_2d_b200(e[a:b], n[a:b], force_east, force_north, mindist, poisson, forces, le, ln)
    return dist.allgather_concat(le, e.size), dist.allgather_concat(ln, e.size)
