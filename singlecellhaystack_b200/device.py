"""Thin object wrapper over the C ABI: one `DeviceContext` per process and GPU.

Host arrays go in as numpy (R layout: column-major double); the `*_dev` methods take torch CUDA
tensors and pass their device pointers straight through (torch is plumbing for device memory and
streams only -- every kernel is in libhaystack_b200.so).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import HaystackLibraryError

__all__ = ["DeviceContext", "HaystackLibraryError"]


def _f64_colmajor(a, name):
    a = np.asarray(a)
    if a.dtype != np.float64 or not a.flags.f_contiguous:
        a = np.asfortranarray(a, dtype=np.float64)
    if a.ndim not in (1, 2):
        raise ValueError(f"{name} must be 1-D or 2-D")
    return a


def _ptr(a) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)


def _dptr(t) -> C.c_void_p:
    """Device pointer of a torch CUDA tensor (or None)."""
    if t is None:
        return C.c_void_p(None)
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return C.c_void_p(t.data_ptr())


def _wrap_device_ptr(ptr: int, nbytes: int, dev):
    """torch uint8 view of device memory owned by the library (so collectives can address it)."""
    import torch

    class _Cai:
        __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}

    return torch.as_tensor(_Cai(), device=dev)


def perm_sample(seed: int, stream: int, n: int, size: int, index_base: int = 1) -> np.ndarray:
    """hs_perm_sample: the `size` cell ids (index_base-based, in draw order) that the seeded randomization
    entry points pair with a gene's values for (seed, stream).  Host-only; stands in for
    sample(x = n, size = size) (R/haystack_continuous.R:275)."""
    lib = _lib.load()
    out = np.empty(int(size), dtype=np.int32)
    rc = lib.hs_perm_sample(int(seed) % (1 << 64), int(stream) % (1 << 64), int(n), int(size), int(index_base), _ptr(out))
    if rc != _lib.HS_OK:
        raise ValueError(f"hs_perm_sample failed ({rc}): {_lib.last_error(None)}")
    return out


class DeviceContext:
    """Owns an `hs_ctx*`.  Not thread-safe (neither is R, the intended caller)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p(None)
        rc = self._lib.hs_create(int(device), C.byref(h))
        if rc != _lib.HS_OK:
            raise HaystackLibraryError(f"hs_create failed ({rc}): {_lib.last_error(None)}")
        self._h = h
        self.device = int(device)

    # -- plumbing -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.hs_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int, what: str):
        if rc != _lib.HS_OK:
            raise HaystackLibraryError(f"{what} failed ({rc}): {_lib.last_error(self._h)}")

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self._lib.hs_set_stream(self._h, C.c_void_p(cuda_stream_ptr or None)), "hs_set_stream")

    def use_torch_stream(self, stream=None):
        """Run the library's kernels on a torch stream (default: torch's current stream), so that
        torch.cuda.Event timings bracket them.  torch's default stream is the legacy NULL stream,
        which the C ABI spells cudaStreamLegacy (NULL means "the context's own stream")."""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        self.set_stream(s.cuda_stream or 0x1)

    def synchronize(self):
        self._check(self._lib.hs_synchronize(self._h), "hs_synchronize")

    @property
    def launch_count(self) -> int:
        return int(self._lib.hs_launch_count(self._h))

    def enable_timers(self, on: bool = True):
        self._check(self._lib.hs_enable_timers(self._h, int(bool(on))), "hs_enable_timers")

    def last_timers(self) -> dict:
        buf = (C.c_float * 5)()
        n = self._lib.hs_last_timers(self._h, buf, 5)
        names = ["density", "csr", "dense", "perm", "copy"]
        return {names[i]: float(buf[i]) for i in range(n)}

    @property
    def shape(self):
        n, g = C.c_int64(0), C.c_int(0)
        self._check(self._lib.hs_density_shape(self._h, C.byref(n), C.byref(g)), "hs_density_shape")
        return int(n.value), int(g.value)

    # -- CSC -> CSR (R/haystack_continuous.R:113-116) -------------------------------------------
    def csc_to_csr(self, cp, ri, x, n_rows: int, n_cols: int):
        """hs_csc_to_csr: dgCMatrix slots (p, i, x) -> dgRMatrix slots (p int64, j int32, x fp64), column ids
        ascending inside every row."""
        cp = np.ascontiguousarray(cp)
        if cp.dtype not in (np.int32, np.int64):
            cp = cp.astype(np.int64)
        ri = np.ascontiguousarray(ri, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        nnz = int(cp[-1]) if cp.size else 0
        p_out = np.empty(int(n_rows) + 1, dtype=np.int64)
        j_out = np.empty(nnz, dtype=np.int32)
        x_out = np.empty(nnz, dtype=np.float64)
        self._check(self._lib.hs_csc_to_csr(self._h, _ptr(cp), int(cp.dtype == np.int64), _ptr(ri), _ptr(x), int(n_rows),
                                            int(n_cols), _ptr(p_out), _ptr(j_out), _ptr(x_out)), "hs_csc_to_csr")
        return p_out, j_out, x_out

    def csc_to_csr_dev(self, cp_dev, ri_dev, x_dev, n_rows: int, n_cols: int, p_out_dev, j_out_dev, x_out_dev):
        self._check(self._lib.hs_csc_to_csr_dev(self._h, _dptr(cp_dev), _dptr(ri_dev), _dptr(x_dev), int(n_rows), int(n_cols),
                                                _dptr(p_out_dev), _dptr(j_out_dev), _dptr(x_out_dev)), "hs_csc_to_csr_dev")

    # -- row statistics (R/haystack_continuous.R:57-74) ---------------------------------------
    def row_stats_csr(self, p, x, n_cells: int):
        """hs_row_stats_csr: (mean, sd) per CSR row, sd with the n-1 denominator and the implicit zeros
        counted; two-pass fp64 on the device."""
        p = np.ascontiguousarray(p)
        if p.dtype not in (np.int32, np.int64):
            p = p.astype(np.int64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        m = p.size - 1
        mean, sd = np.empty(m), np.empty(m)
        self._check(self._lib.hs_row_stats_csr(self._h, _ptr(p), int(p.dtype == np.int64), _ptr(x), m, int(n_cells),
                                               _ptr(mean), _ptr(sd)), "hs_row_stats_csr")
        return mean, sd

    def row_stats_csr_dev(self, p_dev, x_dev, n_cells: int, mean_dev, sd_dev):
        self._check(self._lib.hs_row_stats_csr_dev(self._h, _dptr(p_dev), _dptr(x_dev), p_dev.numel() - 1, int(n_cells),
                                                   _dptr(mean_dev), _dptr(sd_dev)), "hs_row_stats_csr_dev")

    # -- grid selection, "seeding" method (R/haystack_highD.R:87-129) ---------------------------
    def grid_seed_begin(self, x) -> None:
        """hs_grid_seed_begin: coordinates (N, d) resident for the distance passes of the seeding method."""
        x = _f64_colmajor(x, "x")
        if x.ndim == 1:
            x = x.reshape(-1, 1, order="F")
        self._seed_n = int(x.shape[0])
        self._check(self._lib.hs_grid_seed_begin(self._h, _ptr(x), int(x.shape[0]), int(x.shape[1])), "hs_grid_seed_begin")

    def grid_seed_step(self, index: int, want_probs: bool = True):
        """hs_grid_seed_step: cell `index` (0-based) becomes a grid point.  Returns (probs, nearest3): the weights
        min.dist.to.grid^2 of the next draw (None unless wanted) and the 0-based ids of the three cells nearest to the
        new grid point."""
        probs = np.empty(self._seed_n, dtype=np.float64) if want_probs else None
        near = np.empty(3, dtype=np.int64)
        self._check(self._lib.hs_grid_seed_step(self._h, int(index), _ptr(probs) if probs is not None else None, _ptr(near)),
                    "hs_grid_seed_step")
        return probs, near[near >= 0]

    def select_by_cv(self, mean, sd, ranks, return_cv=False):
        """hs_select_by_cv_dev: rows at the 0-based `ranks` of order(coeffVar), coeffVar = (sd + 1e-300) / (mean + 1e-300)
        (R/haystack_continuous.R:222, :227-236) -- sorted on the device.  mean / sd: the row statistics WITHOUT the
        1e-300, numpy arrays or device tensors (as hs_row_stats_csr_dev leaves them)."""
        import torch
        dev = torch.device("cuda", self.device)
        m = mean if isinstance(mean, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(mean, dtype=np.float64), device=dev)
        s = sd if isinstance(sd, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(sd, dtype=np.float64), device=dev)
        if m.dtype != torch.float64 or s.dtype != torch.float64 or m.numel() != s.numel():
            raise ValueError("mean and sd must be float64 vectors of the same length")
        r = np.ascontiguousarray(ranks, dtype=np.int64)
        rows = np.empty(r.size, dtype=np.int64)
        cv = torch.empty(m.numel(), dtype=torch.float64, device=dev) if return_cv else None
        torch.cuda.current_stream(dev).synchronize()
        self._check(self._lib.hs_select_by_cv_dev(self._h, _dptr(m), _dptr(s), int(m.numel()), _ptr(r), int(r.size),
                                                  _dptr(cv) if cv is not None else None, _ptr(rows)), "hs_select_by_cv_dev")
        return (rows, cv) if return_cv else rows

    # -- K1 -----------------------------------------------------------------------------------
    def density(self, x, grid, w_q=None, want_densities=False, want_nearest=False):
        """hs_density: x (N,d), grid (G,d) in the same (scaled) space.  Returns dict with
        bandwidth, Q and optionally densities (N,G) fp64 and nearest (N,) int32 0-based."""
        x = _f64_colmajor(x, "x")
        grid = _f64_colmajor(grid, "grid")
        if x.ndim == 1:
            x = x.reshape(-1, 1, order="F")
        if grid.ndim == 1:
            grid = grid.reshape(-1, 1, order="F")
        if x.shape[1] != grid.shape[1]:
            raise ValueError("Coordinates and grid points have different number of columns.")
        n, d = x.shape
        g = grid.shape[0]
        w = None
        if w_q is not None:
            w = np.ascontiguousarray(w_q, dtype=np.float64)
            if w.shape != (n,):
                raise ValueError("The length of 'weights.advanced.Q' must be the same as the number of rows in 'x'")
        bw = C.c_double(0.0)
        q = np.empty(g, dtype=np.float64)
        dens = np.empty((n, g), dtype=np.float64, order="F") if want_densities else None
        near = np.empty(n, dtype=np.int32) if want_nearest else None
        rc = self._lib.hs_density(self._h, _ptr(x), n, d, _ptr(grid), g, _ptr(w), C.byref(bw), _ptr(q),
                                  _ptr(dens), _ptr(near))
        self._check(rc, "hs_density")
        return dict(bandwidth=float(bw.value), Q=q, densities=dens, nearest=near)

    def density_dev(self, x_dev, grid_dev, w_q_dev=None):
        """x_dev (d,N)-contiguous torch fp64 tensor == column-major N x d; grid_dev likewise (d,G)."""
        d, n = x_dev.shape
        d2, g = grid_dev.shape
        if d != d2:
            raise ValueError("Coordinates and grid points have different number of columns.")
        bw = C.c_double(0.0)
        q = np.empty(g, dtype=np.float64)
        rc = self._lib.hs_density_dev(self._h, _dptr(x_dev), n, d, _dptr(grid_dev), g, _dptr(w_q_dev),
                                      C.byref(bw), _ptr(q))
        self._check(rc, "hs_density_dev")
        return dict(bandwidth=float(bw.value), Q=q)

    def density_buffers(self, n_cells: int, n_grid: int):
        """hs_density_buffers: (d_ptr, d_bytes, q_ptr, q_bytes) of the resident (or freshly reserved) state."""
        d, q = C.c_void_p(None), C.c_void_p(None)
        db, qb = C.c_int64(0), C.c_int64(0)
        self._check(self._lib.hs_density_buffers(self._h, int(n_cells), int(n_grid), C.byref(d), C.byref(db),
                                                 C.byref(q), C.byref(qb)), "hs_density_buffers")
        return int(d.value), int(db.value), int(q.value), int(qb.value)

    def density_commit(self, n_cells: int, n_grid: int):
        self._check(self._lib.hs_density_commit(self._h, int(n_cells), int(n_grid)), "hs_density_commit")

    def density_broadcast(self, n_cells: int, n_grid: int, src: int = 0, group=None):
        """Multi-GPU exchange step of the path: the density matrix, Q and the bandwidth produced by
        hs_density on rank `src` are broadcast (torch.distributed: NCCL over NVLink on GPUs) straight
        between the contexts' own device buffers -- no staging copy."""
        import torch
        import torch.distributed as dist
        d_ptr, d_bytes, q_ptr, q_bytes = self.density_buffers(n_cells, n_grid)
        dev = torch.device("cuda", self.device)
        # The collective runs on torch's stream, the library on the context's stream (its own
        # non-blocking stream unless use_torch_stream() joined them): order the two explicitly.
        # Before: everything the context has queued (the density stage on `src`, readers of the
        # old state elsewhere) must have finished.  After: the context may not read D / Q /
        # bandwidth until the broadcast has landed (NCCL calls return before they complete).
        self.synchronize()
        for ptr, nbytes in ((d_ptr, d_bytes), (q_ptr, q_bytes)):
            t = _wrap_device_ptr(ptr, nbytes, dev)
            dist.broadcast(t, src=src, group=group)
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()
        if dist.get_rank(group) != src:
            self.density_commit(n_cells, n_grid)

    def density_sharded_dev(self, x_dev, grid_dev, w_q_dev=None, group=None, rank=None, world=None, all_gather=None):
        """The density stage with the cells sharded over the ranks of `group` (hs_density_shard_*): every rank
        computes 1 / world of the distances and densities; three in-place all-gathers (row minima 8 B per cell,
        density rows, per-block column sums) leave every rank with the state hs_density_dev would have produced,
        bit for bit.  `all_gather(buffer_ptr, bytes_per_rank)` may replace the torch.distributed collective
        (tests drive several ranks' parts through one context that way)."""
        import torch
        import torch.distributed as dist
        d, n = x_dev.shape
        d2, g = grid_dev.shape
        if d != d2:
            raise ValueError("Coordinates and grid points have different number of columns.")
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        dev = torch.device("cuda", self.device)

        def gather(ptr, per):
            if all_gather is not None:
                return all_gather(ptr, per)
            if world == 1:
                return
            self.synchronize()
            whole = _wrap_device_ptr(ptr, per * world, dev)
            dist.all_gather_into_tensor(whole, whole[rank * per:(rank + 1) * per], group=group)
            torch.cuda.current_stream(dev).synchronize()

        self._check(self._lib.hs_density_shard_dist_dev(self._h, _dptr(x_dev), n, d, _dptr(grid_dev), g, int(rank), int(world)),
                    "hs_density_shard_dist_dev")
        rm, dd, qp = C.c_void_p(None), C.c_void_p(None), C.c_void_p(None)
        rmb, ddb, qpb = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._check(self._lib.hs_density_shard_buffers(self._h, n, g, int(world), C.byref(rm), C.byref(rmb), C.byref(dd),
                                                       C.byref(ddb), C.byref(qp), C.byref(qpb)), "hs_density_shard_buffers")
        gather(rm.value, rmb.value)
        self._check(self._lib.hs_density_shard_dens_dev(self._h, n, g, _dptr(w_q_dev), int(rank), int(world)),
                    "hs_density_shard_dens_dev")
        gather(dd.value, ddb.value)
        gather(qp.value, qpb.value)
        bw = C.c_double(0.0)
        q = np.empty(g, dtype=np.float64)
        self._check(self._lib.hs_density_shard_finish(self._h, n, g, C.byref(bw), _ptr(q)), "hs_density_shard_finish")
        return dict(bandwidth=float(bw.value), Q=q)

    def dkl_perm_csr_seeded_units_dev(self, gene_ptr_dev, val_dev, n_perm, seed, unit_stream_dev, out_dev):
        """hs_dkl_perm_csr_seeded_units_dev: unit u / randomization r draws from stream unit_stream_dev[u] + r."""
        n_units = gene_ptr_dev.numel() - 1
        self._check(self._lib.hs_dkl_perm_csr_seeded_units_dev(self._h, _dptr(gene_ptr_dev), n_units, _dptr(val_dev),
                                                               int(n_perm), int(seed) % (1 << 64), _dptr(unit_stream_dev),
                                                               _dptr(out_dev)), "hs_dkl_perm_csr_seeded_units_dev")

    # -- K2/K3 --------------------------------------------------------------------------------
    def dkl_dense(self, expression) -> np.ndarray:
        """expression: (M, N) genes x cells; any layout, passed as an R matrix (column-major fp64)."""
        e = _f64_colmajor(expression, "expression")
        if e.ndim != 2:
            raise ValueError("'expression' must be a matrix")
        m, n = e.shape
        if n != self.shape[0]:
            raise ValueError("The number of columns in 'expression' must be the same as the rows in 'x'")
        out = np.empty(m, dtype=np.float64)
        self._check(self._lib.hs_dkl_dense(self._h, _ptr(e), m, m, _ptr(out)), "hs_dkl_dense")
        return out

    def dkl_csr(self, p, j, x) -> np.ndarray:
        """dgRMatrix slots: p (M+1, int32/int64), j (int32, 0-based), x (float64)."""
        p = np.ascontiguousarray(p)
        if p.dtype == np.int64:
            is64 = 1
        else:
            p = np.ascontiguousarray(p, dtype=np.int32)
            is64 = 0
        j = np.ascontiguousarray(j, dtype=np.int32)
        x = np.asarray(x)
        m = p.size - 1
        out = np.empty(max(m, 0), dtype=np.float64)
        if x.dtype == np.float32:      # float32 hosts (AnnData): no widening, a third less to upload
            x = np.ascontiguousarray(x)
            self._check(self._lib.hs_dkl_csr_f32(self._h, _ptr(p), is64, _ptr(j), _ptr(x), m, _ptr(out)),
                        "hs_dkl_csr_f32")
            return out
        x = np.ascontiguousarray(x, dtype=np.float64)
        self._check(self._lib.hs_dkl_csr(self._h, _ptr(p), is64, _ptr(j), _ptr(x), m, _ptr(out)), "hs_dkl_csr")
        return out

    # -- prepare_clustering (R/haystack_continuous.R:604-666) ---------------------------------
    def profile_csr(self, p, j, x) -> np.ndarray:
        """hs_profile_csr: (M, G) matrix of per-gene column sums colSums(D[ind,] * val), rows rescaled to sum to 1."""
        p = np.ascontiguousarray(p)
        if p.dtype not in (np.int32, np.int64):
            p = p.astype(np.int64)
        j = np.ascontiguousarray(j, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        m = p.size - 1
        out = np.empty((m, self.shape[1]), dtype=np.float64, order="F")
        self._check(self._lib.hs_profile_csr(self._h, _ptr(p), int(p.dtype == np.int64), _ptr(j), _ptr(x), m, _ptr(out)),
                    "hs_profile_csr")
        return out

    def profile_dense(self, expression) -> np.ndarray:
        e = _f64_colmajor(expression, "expression")
        if e.ndim != 2 or e.shape[1] != self.shape[0]:
            raise ValueError("The number of columns in 'expression' must be the same as the rows in 'x'")
        m = e.shape[0]
        out = np.empty((m, self.shape[1]), dtype=np.float64, order="F")
        self._check(self._lib.hs_profile_dense(self._h, _ptr(e), m, m, _ptr(out)), "hs_profile_dense")
        return out

    # -- K4 -----------------------------------------------------------------------------------
    def dkl_perm_dense(self, v, perm, index_base=1) -> np.ndarray:
        """v (N,), perm (N, R) column r = sample.int(N).  Returns (R,)."""
        v = np.ascontiguousarray(v, dtype=np.float64)
        perm = np.asfortranarray(perm, dtype=np.int32)
        n, r = perm.shape
        if v.shape != (n,) or n != self.shape[0]:
            raise ValueError("perm must be n_cells x n_perm and v of length n_cells")
        out = np.empty(r, dtype=np.float64)
        self._check(self._lib.hs_dkl_perm_dense(self._h, _ptr(v), _ptr(perm), r, int(index_base), _ptr(out)),
                    "hs_dkl_perm_dense")
        return out

    def dkl_perm_dense_seeded(self, v, n_perm: int, seed: int, stream_base: int = 0) -> np.ndarray:
        """hs_dkl_perm_dense_seeded: randomization r scores v[pi_r], pi_r = perm_sample(seed, stream_base + r,
        N, N) drawn on the device.  Returns (R,)."""
        v = np.ascontiguousarray(v, dtype=np.float64)
        if v.shape != (self.shape[0],):
            raise ValueError("v must have length n_cells")
        out = np.empty(int(n_perm), dtype=np.float64)
        self._check(self._lib.hs_dkl_perm_dense_seeded(self._h, _ptr(v), int(n_perm), int(seed) % (1 << 64),
                                                       int(stream_base) % (1 << 64), _ptr(out)),
                    "hs_dkl_perm_dense_seeded")
        return out

    def dkl_perm_dense_seeded_batch(self, expression, rows, n_perm: int, seed: int, stream_base: int = 0) -> np.ndarray:
        """hs_dkl_perm_dense_seeded_batch: the dense randomization of all `rows` (0-based) of the genes x cells matrix in
        one call; stream of (s, r) = stream_base + s * n_perm + r.  Returns (len(rows), R)."""
        e = _f64_colmajor(expression, "expression")
        if e.ndim != 2 or e.shape[1] != self.shape[0]:
            raise ValueError("expression must be genes x n_cells")
        r = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.empty((r.size, int(n_perm)), dtype=np.float64, order="F")
        self._check(self._lib.hs_dkl_perm_dense_seeded_batch(self._h, _ptr(e), int(e.shape[0]), _ptr(r), int(r.size),
                                                             int(n_perm), int(seed) % (1 << 64),
                                                             int(stream_base) % (1 << 64), _ptr(out)),
                    "hs_dkl_perm_dense_seeded_batch")
        return out

    def dkl_perm_csr(self, vals: list, inds: list, index_base=1) -> np.ndarray:
        """vals[s]: (nnz_s,) values of selected gene s; inds[s]: (nnz_s, R) column r =
        sample(n_cells, nnz_s).  Returns (S, R) (all.D_KL.randomized)."""
        s_ = len(vals)
        if s_ == 0:
            return np.zeros((0, 0))
        r = inds[0].shape[1]
        gp = np.zeros(s_ + 1, dtype=np.int64)
        for i, v in enumerate(vals):
            if inds[i].shape != (len(v), r):
                raise ValueError("inds[s] must be nnz_s x n_perm")
            gp[i + 1] = gp[i] + len(v)
        val = np.concatenate([np.asarray(v, dtype=np.float64) for v in vals]) if gp[-1] else np.zeros(0)
        ind = (np.concatenate([np.asfortranarray(a, dtype=np.int32).ravel(order="F") for a in inds])
               if gp[-1] else np.zeros(0, dtype=np.int32))
        out = np.empty((s_, r), dtype=np.float64, order="F")
        self._check(self._lib.hs_dkl_perm_csr(self._h, _ptr(gp), s_, _ptr(val), _ptr(ind), r, int(index_base),
                                              _ptr(out)), "hs_dkl_perm_csr")
        return out

    def dkl_perm_csr_begin(self, vals: list, inds: list, index_base=1) -> None:
        """hs_dkl_perm_csr_begin: start dkl_perm_csr(vals, inds) on the library's worker thread and return at once; the
        host can draw the next batch meanwhile.  dkl_perm_csr_wait() delivers the result."""
        s_ = len(vals)
        r = inds[0].shape[1] if s_ else 1
        gp = np.zeros(s_ + 1, dtype=np.int64)
        for i, v in enumerate(vals):
            if inds[i].shape != (len(v), r):
                raise ValueError("inds[s] must be nnz_s x n_perm")
            gp[i + 1] = gp[i] + len(v)
        val = np.concatenate([np.asarray(v, dtype=np.float64) for v in vals]) if gp[-1] else np.zeros(0)
        ind = (np.concatenate([np.asfortranarray(a, dtype=np.int32).ravel(order="F") for a in inds])
               if gp[-1] else np.zeros(0, dtype=np.int32))
        self._async_keep = (gp, val, ind, (s_, r))          # the library reads val / ind until the wait
        self._check(self._lib.hs_dkl_perm_csr_begin(self._h, _ptr(gp), s_, _ptr(val), _ptr(ind), r, int(index_base)),
                    "hs_dkl_perm_csr_begin")

    def dkl_perm_csr_wait(self) -> np.ndarray:
        """hs_dkl_perm_csr_wait: block until the call started by dkl_perm_csr_begin is done.  Returns (S, R)."""
        keep = getattr(self, "_async_keep", None)
        if keep is None:
            raise RuntimeError("no asynchronous call in flight")
        out = np.empty(keep[3], dtype=np.float64, order="F")
        try:
            self._check(self._lib.hs_dkl_perm_csr_wait(self._h, _ptr(out)), "hs_dkl_perm_csr_wait")
        finally:
            self._async_keep = None
        return out

    def dkl_perm_csr_seeded(self, vals: list, n_perm: int, seed: int, stream_base: int = 0) -> np.ndarray:
        """hs_dkl_perm_csr_seeded: like dkl_perm_csr, but gene s / randomization r draws its cells on the
        device from the keyed permutation (seed, stream_base + s * n_perm + r); `perm_sample` returns the
        same ids on the host.  Returns (S, R)."""
        s_ = len(vals)
        if s_ == 0:
            return np.zeros((0, 0))
        gp = np.zeros(s_ + 1, dtype=np.int64)
        for i, v in enumerate(vals):
            gp[i + 1] = gp[i] + len(v)
        val = np.concatenate([np.asarray(v, dtype=np.float64) for v in vals]) if gp[-1] else np.zeros(0)
        out = np.empty((s_, int(n_perm)), dtype=np.float64, order="F")
        self._check(self._lib.hs_dkl_perm_csr_seeded(self._h, _ptr(gp), s_, _ptr(val), int(n_perm),
                                                     int(seed) % (1 << 64), int(stream_base) % (1 << 64), _ptr(out)),
                    "hs_dkl_perm_csr_seeded")
        return out

    def dkl_csr_randomized(self, p, j, x, sel_rows, n_perm: int, seed: int, stream_base: int = 0, out_rand=None):
        """hs_dkl_csr_randomized: D_KL of every CSR row plus the seeded randomization batch of rows `sel_rows`
        (0-based) in one call; the batch computes while the matrix uploads.  Returns (dkl (M,), rand (S, R))."""
        p = np.ascontiguousarray(p)
        if p.dtype not in (np.int32, np.int64):
            p = p.astype(np.int64)
        j = np.ascontiguousarray(j, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        sel = np.ascontiguousarray(sel_rows, dtype=np.int64)
        m = p.size - 1
        dkl = np.empty(m, dtype=np.float64)
        if out_rand is None:
            out_rand = np.empty((sel.size, int(n_perm)), dtype=np.float64, order="F")
        self._check(self._lib.hs_dkl_csr_randomized(self._h, _ptr(p), int(p.dtype == np.int64), _ptr(j), _ptr(x), m,
                                                    _ptr(sel), sel.size, int(n_perm), int(seed) % (1 << 64),
                                                    int(stream_base) % (1 << 64), _ptr(dkl), _ptr(out_rand)),
                    "hs_dkl_csr_randomized")
        return dkl, out_rand

    def dkl_perm_csr_seeded_flat(self, gene_ptr, val, n_perm, seed, stream_base=0, out=None) -> np.ndarray:
        """hs_dkl_perm_csr_seeded on flattened host arrays (gene_ptr int64 [S+1], val fp64)."""
        gene_ptr = np.ascontiguousarray(gene_ptr, dtype=np.int64)
        val = np.ascontiguousarray(val, dtype=np.float64)
        s_ = gene_ptr.size - 1
        if out is None:
            out = np.empty((s_, int(n_perm)), dtype=np.float64, order="F")
        self._check(self._lib.hs_dkl_perm_csr_seeded(self._h, _ptr(gene_ptr), s_, _ptr(val), int(n_perm),
                                                     int(seed) % (1 << 64), int(stream_base) % (1 << 64), _ptr(out)),
                    "hs_dkl_perm_csr_seeded")
        return out

    def dkl_perm_csr_flat(self, gene_ptr, val, ind, n_perm, index_base=1, out=None) -> np.ndarray:
        """hs_dkl_perm_csr on already flattened host arrays (gene_ptr int64 [S+1], val fp64, ind int32
        laid out gene by gene as nnz_s x n_perm column-major blocks).  No host-side copies."""
        gene_ptr = np.ascontiguousarray(gene_ptr, dtype=np.int64)
        val = np.asarray(val)
        f32 = val.dtype == np.float32
        val = np.ascontiguousarray(val, dtype=np.float32 if f32 else np.float64)
        ind = np.ascontiguousarray(ind, dtype=np.int32)
        s_ = gene_ptr.size - 1
        if out is None:
            out = np.empty((s_, int(n_perm)), dtype=np.float64, order="F")
        fn = self._lib.hs_dkl_perm_csr_f32 if f32 else self._lib.hs_dkl_perm_csr
        self._check(fn(self._h, _ptr(gene_ptr), s_, _ptr(val), _ptr(ind), int(n_perm), int(index_base), _ptr(out)),
                    "hs_dkl_perm_csr")
        return out

    # -- device-resident variants -------------------------------------------------------------
    def dkl_csr_dev(self, p_dev, j_dev, x_dev, out_dev):
        m = p_dev.numel() - 1
        self._check(self._lib.hs_dkl_csr_dev(self._h, _dptr(p_dev), _dptr(j_dev), _dptr(x_dev), m, _dptr(out_dev)),
                    "hs_dkl_csr_dev")

    def dkl_perm_csr_dev(self, gene_ptr_dev, val_dev, ind_dev, n_perm, out_dev, index_base=0):
        s_ = gene_ptr_dev.numel() - 1
        self._check(self._lib.hs_dkl_perm_csr_dev(self._h, _dptr(gene_ptr_dev), s_, _dptr(val_dev), _dptr(ind_dev),
                                                  int(n_perm), int(index_base), _dptr(out_dev)),
                    "hs_dkl_perm_csr_dev")

    def dkl_perm_csr_seeded_dev(self, gene_ptr_dev, val_dev, n_perm, seed, out_dev, stream_base=0):
        s_ = gene_ptr_dev.numel() - 1
        self._check(self._lib.hs_dkl_perm_csr_seeded_dev(self._h, _dptr(gene_ptr_dev), s_, _dptr(val_dev), int(n_perm),
                                                         int(seed) % (1 << 64), int(stream_base) % (1 << 64),
                                                         _dptr(out_dev)), "hs_dkl_perm_csr_seeded_dev")

    def dkl_dense_dev(self, expr_dev, n_genes, ld, out_dev):
        self._check(self._lib.hs_dkl_dense_dev(self._h, _dptr(expr_dev), int(n_genes), int(ld), _dptr(out_dev)),
                    "hs_dkl_dense_dev")

    def dkl_perm_dense_seeded_dev(self, v_dev, n_perm, seed, out_dev, stream_base=0):
        self._check(self._lib.hs_dkl_perm_dense_seeded_dev(self._h, _dptr(v_dev), int(n_perm), int(seed) % (1 << 64),
                                                           int(stream_base) % (1 << 64), _dptr(out_dev)),
                    "hs_dkl_perm_dense_seeded_dev")

    def dkl_perm_dense_dev(self, v_dev, perm_dev, n_perm, out_dev, index_base=0):
        self._check(self._lib.hs_dkl_perm_dense_dev(self._h, _dptr(v_dev), _dptr(perm_dev), int(n_perm),
                                                    int(index_base), _dptr(out_dev)), "hs_dkl_perm_dense_dev")
