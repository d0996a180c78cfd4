"""Multi-GPU sigma build: the CI vector and sigma are partitioned by alpha-string rows over the ranks
of one torch.distributed group (one process per GPU, NCCL over NVLink on the B200 box).

No reference counterpart exists (forte2 is single-process); correctness is
defined as equality with the single-GPU / reference result.

Rank r owns rows [r*nIa/world, (r+1)*nIa/world) of every determinant block (same rule for columns).
The local vector is the concatenation of the owned row slabs in block order.

Full form (any string space), per sigma build
  1. all-gather of the trial vector (each rank contributes its owned row slabs)            -- NCCL
  2. local work through the C ABI (f2_sigma_apply_part): H0, beta same-spin and the opposite-spin term
     for the owned ROWS (these only write owned rows), and the alpha same-spin term, which couples
     rows, for the owned COLUMNS of every row
  3. exchange of the alpha same-spin partial results to the row owners (all-to-all) and add  -- NCCL

Rows form (spaces served by a gather-form opposite-spin kernel: t3g_kernel for one determinant block,
t3gp_kernel for symmetry / GAS blocks; no rank ever holds the full vector), per sigma build
  1. all-to-all of C: rank r sends C[rows of r, columns of q] of every block to q and assembles the
     column slabs [all rows][columns of r]                                                   -- NCCL, ndet/world doubles
  2. local work (f2_sigma_apply_rows): H0 and beta same-spin term of the owned rows from the local
     slab; alpha same-spin term of the column slabs; opposite-spin term of the owned INPUT rows,
     scattered to every sigma row in a buffer of world equal slots (slot q = the local vector of rank q)
  3. all-to-all of the alpha same-spin result back to the row owners, add                      -- NCCL
  4. reduce-scatter of the opposite-spin buffer, add                                          -- NCCL, ndet doubles
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class Layout:
    """Ownership arithmetic, identical to make_part() in csrc/sigma.cu (checked in the GPU tests)."""

    def __init__(self, blocks: np.ndarray, rank: int, world: int):
        self.rank, self.world = rank, world
        self.blocks = [(int(off), int(nA), int(nB)) for _, _, off, nA, nB in blocks]
        self.rows, self.cols = [], []  # per block: (first, count, offset in local vector / exchange buffer)
        loc = y = 0
        for off, nA, nB in self.blocks:
            lo, hi = nA * rank // world, nA * (rank + 1) // world
            cl, ch = nB * rank // world, nB * (rank + 1) // world
            self.rows.append((lo, hi - lo, loc))
            self.cols.append((cl, ch - cl, y))
            if nA * nB > 0:
                loc += (hi - lo) * nB
                y += nA * (ch - cl)
        self.local_size, self.exchange_size = loc, y
        self.ndet = sum(nA * nB for _, nA, nB in self.blocks)
        # owned pieces: (first local index, first CI-vector address, length), ascending in both
        self.pieces = [(loc_b, off + lo * nB, n * nB) for (off, nA, nB), (lo, n, loc_b) in zip(self.blocks, self.rows)
                       if nA * nB > 0 and n > 0]

    def global_of_local(self, idx: np.ndarray) -> np.ndarray:
        """CI-vector addresses of local indices"""
        idx = np.asarray(idx, dtype=np.int64)
        starts = np.array([p[0] for p in self.pieces], dtype=np.int64)
        k = np.searchsorted(starts, idx, side="right") - 1
        shift = np.array([p[1] - p[0] for p in self.pieces], dtype=np.int64)
        return idx + shift[k]

    def local_of_global(self, addr: np.ndarray) -> np.ndarray:
        """local indices of CI-vector addresses, -1 where this rank does not own the determinant"""
        addr = np.asarray(addr, dtype=np.int64)
        out = np.full(addr.shape, -1, dtype=np.int64)
        for loc_b, g0, n in self.pieces:
            m = (addr >= g0) & (addr < g0 + n)
            out[m] = addr[m] - g0 + loc_b
        return out


class CudaEngine:
    """Local work on one GPU through the C ABI."""

    def __init__(self, builder):
        self.builder = builder

    def apply_part(self, c_full, sigma_loc, yexch, rank, world):
        stream = torch.cuda.current_stream(self.builder.device).cuda_stream
        _lib.check(_lib.lib().f2_sigma_apply_part(self.builder._h, c_full.data_ptr(), sigma_loc.data_ptr(),
                                                  yexch.data_ptr(), rank, world, stream))

    def rows_form(self, rank, world):
        """(available, doubles per reduce-scatter slot) of the rows form (f2_sigma_rows_form)"""
        out = np.zeros(4, dtype=np.int64)
        _lib.check(_lib.lib().f2_sigma_rows_form(self.builder._h, rank, world, out.ctypes.data_as(_lib.i64p)))
        return bool(out[0]), int(out[1])

    def apply_rows(self, c_loc, xcols, sigma_loc, ycols, scatter, rank, world, phases=3):
        stream = torch.cuda.current_stream(self.builder.device).cuda_stream
        _lib.check(_lib.lib().f2_sigma_apply_rows(self.builder._h, c_loc.data_ptr(), xcols.data_ptr(),
                                                  sigma_loc.data_ptr(), ycols.data_ptr(), scatter.data_ptr(),
                                                  rank, world, phases, stream))

    def apply_rows_t3_piece(self, c_loc, scatter, rank, world, row_first, row_count, flags):
        stream = torch.cuda.current_stream(self.builder.device).cuda_stream
        _lib.check(_lib.lib().f2_sigma_apply_rows_t3_piece(self.builder._h, c_loc.data_ptr(), scatter.data_ptr(), rank, world,
                                                           int(row_first), int(row_count), int(flags), stream))

    def layout_from_library(self, nblocks, rank, world):
        sizes = np.zeros(2, dtype=np.int64)
        rows = np.zeros((nblocks, 6), dtype=np.int64)
        _lib.check(_lib.lib().f2_sigma_part_layout(self.builder._h, rank, world, sizes.ctypes.data_as(_lib.i64p),
                                                   rows.ctypes.data_as(_lib.i64p)))
        return sizes, rows


class ShardedSigma:
    """sigma = H C with C and sigma sharded by alpha rows.  `engine` does the local work: CudaEngine on
    the GPU box; the gloo tests pass a CPU stand-in with the same apply_part signature."""

    def __init__(self, builder=None, rank=None, world=None, group=None, engine=None, blocks=None, device=None,
                 form="auto", collectives="auto"):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.engine = engine if engine is not None else CudaEngine(builder)
        if blocks is None:
            blocks = builder.ci_strings.blocks()
        self.layouts = [Layout(blocks, q, self.world) for q in range(self.world)]
        self.layout = self.layouts[self.rank]
        self.local_size = self.layout.local_size
        self.ndet = self.layout.ndet
        self.device = device if device is not None else (
            torch.device("cuda", builder.device) if builder is not None else torch.device("cpu"))
        self._max_local = max(lay.local_size for lay in self.layouts)
        self._kw = kw = dict(dtype=torch.float64, device=self.device)
        self._phase_events = None
        self._full = self._gather = self._pad = None  # full-layout buffers, allocated on first use
        self._yexch = torch.zeros(max(self.layout.exchange_size, 1), **kw)
        # rows form: wanted unless form == "full", used when the engine offers it for this space
        self.form = "full"
        if form not in ("auto", "full", "rows"):
            raise ValueError("form must be 'auto', 'full' or 'rows'")
        if form != "full" and self.world > 1 and hasattr(self.engine, "rows_form"):
            ok, slot = self.engine.rows_form(self.rank, self.world)
            if ok:
                assert slot == self._max_local
                self.form = "rows"
                self._slot = slot
                self._xcols = torch.zeros(max(self.layout.exchange_size, 1), **kw)
                self._scatter = torch.empty(self.world * self._slot, **kw)
                self._rs_out = torch.empty(self._slot, **kw)
        if form == "rows" and self.form != "rows" and self.world > 1:
            raise RuntimeError("the rows form of the sharded build is not available for this string space")
        # who issues the collectives of the rows form: this module (torch.distributed) or the C library (f2_dist_apply:
        # one pack / unpack kernel per exchange instead of one copy per block and peer).  "auto": the library for spaces
        # with many determinant blocks (3 GAS / 20 orbitals, 225 blocks, 8 GPUs: 737 against 784 ms per build), this
        # module otherwise (one block is received in place: CAS(18,18) 397 against 399 ms)
        if collectives not in ("auto", "python", "library"):
            raise ValueError("collectives must be 'auto', 'python' or 'library'")
        self._lib = None
        nonempty = sum(1 for _, nA, nB in self.layout.blocks if nA * nB > 0)
        want_lib = collectives == "library" or (collectives == "auto" and nonempty > 8)
        if (want_lib and self.form == "rows" and isinstance(self.engine, CudaEngine) and dist.is_initialized()
                and dist.get_backend(group) == "nccl"):
            self._lib = LibraryShardedSigma(builder, self.rank, self.world, group)
        elif collectives == "library":
            raise RuntimeError("the library-owned collectives need the rows form on CUDA with an NCCL process group")
        if isinstance(self.engine, CudaEngine):
            sizes, rows = self.engine.layout_from_library(len(self.layout.blocks), self.rank, self.world)
            assert int(sizes[0]) == self.layout.local_size and int(sizes[1]) == self.layout.exchange_size
            for b, (lo, n, off) in enumerate(self.layout.rows):
                assert (int(rows[b, 0]), int(rows[b, 1])) == (lo, n)

    # ---- layout helpers ----
    def scatter_full(self, full: torch.Tensor) -> torch.Tensor:
        """owned row slabs of a full-layout vector (for tests and initial guesses)"""
        out = torch.empty(self.local_size, dtype=full.dtype, device=full.device)
        for (off, nA, nB), (lo, n, loc) in zip(self.layout.blocks, self.layout.rows):
            if nA * nB > 0 and n > 0:
                out[loc:loc + n * nB] = full[off + lo * nB: off + (lo + n) * nB]
        return out

    def all_gather_full(self, c_loc: torch.Tensor) -> torch.Tensor:
        """full-layout copy of a sharded vector on every rank (step 1)"""
        if self._full is None:
            self._full = torch.empty(self.ndet, **self._kw)
        if self.world == 1:
            self._full.copy_(c_loc)
            return self._full
        if self._gather is None:
            self._gather = [torch.empty(self._max_local, **self._kw) for _ in range(self.world)]
            self._pad = torch.zeros(self._max_local, **self._kw)
        self._pad[: self.local_size].copy_(c_loc)
        dist.all_gather(self._gather, self._pad, group=self.group)
        for q, lay in enumerate(self.layouts):
            for (off, nA, nB), (lo, n, loc) in zip(lay.blocks, lay.rows):
                if nA * nB > 0 and n > 0:
                    self._full[off + lo * nB: off + (lo + n) * nB].copy_(self._gather[q][loc:loc + n * nB])
        return self._full

    def _exchange_plan(self):
        """per peer q: the slices of the exchange buffer that go to q, the receive buffer, and where the
        received pieces are added (computed once; the buffers are reused by every build)"""
        if getattr(self, "_xplan", None) is not None:
            return self._xplan
        me = self.layout
        plan = []
        for q, lay in enumerate(self.layouts):
            send_slices, adds, n_recv = [], [], 0
            for b, (off, nA, nB) in enumerate(me.blocks):
                if nA * nB == 0:
                    continue
                cl, nc, yoff = me.cols[b]
                lo_q, n_q, _ = lay.rows[b]
                if nc > 0 and n_q > 0:
                    send_slices.append((yoff + lo_q * nc, yoff + (lo_q + n_q) * nc))
                lo, n, loc = me.rows[b]
                cl_q, nc_q, _ = lay.cols[b]
                if n > 0 and nc_q > 0:
                    adds.append((loc, n, nB, cl_q, nc_q, n_recv))
                n_recv += n * nc_q
            n_send = sum(hi - lo for lo, hi in send_slices)
            # one contiguous slice needs no packing; several (symmetry / GAS blocks) are packed into a buffer
            send_buf = torch.empty(n_send, **self._kw) if len(send_slices) > 1 else None
            recv_buf = torch.empty(n_recv, **self._kw) if q != self.rank else None
            plan.append((send_slices, send_buf, recv_buf, adds))
        self._xplan = plan
        return plan

    def _exchange_add_start(self):
        """send the alpha same-spin partials [rows of q][my columns] to q (step 3); the transfers run beside whatever is
        launched next, _exchange_add_finish waits for them and adds what arrived"""
        plan = self._exchange_plan()
        ops, sends = [], []
        for q, (send_slices, send_buf, recv_buf, adds) in enumerate(plan):
            if len(send_slices) == 1:
                send = self._yexch[send_slices[0][0]:send_slices[0][1]]
            elif send_slices:
                pos = 0
                for lo, hi in send_slices:
                    send_buf[pos:pos + hi - lo].copy_(self._yexch[lo:hi])
                    pos += hi - lo
                send = send_buf
            else:
                send = self._yexch[:0]
            sends.append(send)
            if q == self.rank:
                continue
            if send.numel():
                ops.append(dist.P2POp(dist.isend, send, q, group=self.group))
            if recv_buf.numel():
                ops.append(dist.P2POp(dist.irecv, recv_buf, q, group=self.group))
        works = dist.batch_isend_irecv(ops) if ops else []
        return works, sends

    def _exchange_add_finish(self, pending, sigma_loc: torch.Tensor):
        works, sends = pending
        for w in works:
            w.wait()
        for q, (send_slices, send_buf, recv_buf, adds) in enumerate(self._exchange_plan()):
            src = sends[q] if q == self.rank else recv_buf
            for loc, n, nB, cl_q, nc_q, pos in adds:
                blk = sigma_loc[loc:loc + n * nB].view(n, nB)
                blk[:, cl_q:cl_q + nc_q] += src[pos:pos + n * nc_q].view(n, nc_q)

    def _exchange_add(self, sigma_loc: torch.Tensor):
        self._exchange_add_finish(self._exchange_add_start(), sigma_loc)

    def _cols_plan(self):
        """rows form, step 1, per peer q: the pieces C[my rows, columns of q] of every block that go to q
        (packed into one send buffer) and where the pieces [rows of q][my columns] that arrive belong in the
        column slab (computed once; the buffers are reused by every build)"""
        if getattr(self, "_cplan", None) is not None:
            return self._cplan
        me = self.layout
        plan = []
        for q, lay in enumerate(self.layouts):
            sends, recvs = [], []
            for b, (off, nA, nB) in enumerate(me.blocks):
                if nA * nB == 0:
                    continue
                lo, n, loc = me.rows[b]
                cl, nc, yoff = me.cols[b]
                lo_q, n_q, _ = lay.rows[b]
                cl_q, nc_q, _ = lay.cols[b]
                if n > 0 and nc_q > 0:
                    sends.append((loc, n, nB, cl_q, nc_q))
                if n_q > 0 and nc > 0:
                    recvs.append((yoff + lo_q * nc, n_q * nc))
            n_send = sum(n * nc_q for _, n, _, _, nc_q in sends)
            n_recv = sum(m for _, m in recvs)
            send_buf = torch.empty(n_send, **self._kw) if q != self.rank else None
            # one contiguous piece is received in place; several (symmetry / GAS blocks) through a buffer
            recv_buf = torch.empty(n_recv, **self._kw) if q != self.rank and len(recvs) > 1 else None
            plan.append((sends, recvs, send_buf, recv_buf))
        self._cplan = plan
        return plan

    def _exchange_cols_start(self, c_loc: torch.Tensor):
        """rows form, step 1: column slabs [all rows][my columns] of C (per block, laid out like the
        exchange buffer) from the row slabs; the transfers are left in flight (the opposite-spin kernel, which reads the
        local rows only, runs beside them), _exchange_cols_finish waits and unpacks"""
        ops = []
        plan = self._cols_plan()
        for q, (sends, recvs, send_buf, recv_buf) in enumerate(plan):
            if q == self.rank:
                for (loc, n, nB, cl_q, nc_q), (start, m) in zip(sends, recvs):
                    self._xcols[start:start + m].view(n, nc_q).copy_(c_loc[loc:loc + n * nB].view(n, nB)[:, cl_q:cl_q + nc_q])
                continue
            pos = 0
            for loc, n, nB, cl_q, nc_q in sends:
                send_buf[pos:pos + n * nc_q].view(n, nc_q).copy_(c_loc[loc:loc + n * nB].view(n, nB)[:, cl_q:cl_q + nc_q])
                pos += n * nc_q
            if pos:
                ops.append(dist.P2POp(dist.isend, send_buf, q, group=self.group))
            if len(recvs) == 1:
                ops.append(dist.P2POp(dist.irecv, self._xcols[recvs[0][0]:recvs[0][0] + recvs[0][1]], q, group=self.group))
            elif recvs:
                ops.append(dist.P2POp(dist.irecv, recv_buf, q, group=self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def _exchange_cols_finish(self, works):
        for w in works:
            w.wait()
        for q, (sends, recvs, send_buf, recv_buf) in enumerate(self._cols_plan()):
            if recv_buf is None:
                continue
            pos = 0
            for start, m in recvs:
                self._xcols[start:start + m].copy_(recv_buf[pos:pos + m])
                pos += m

    def _exchange_cols(self, c_loc: torch.Tensor):
        self._exchange_cols_finish(self._exchange_cols_start(c_loc))

    def _reduce_scatter_start(self):
        """rows form, step 4: sum over ranks of slot `rank` of the opposite-spin buffers; asynchronous on
        NCCL, so that it overlaps the same-spin kernels launched next"""
        if dist.get_backend(self.group) == "nccl":
            return dist.reduce_scatter_tensor(self._rs_out, self._scatter, op=dist.ReduceOp.SUM, group=self.group,
                                              async_op=True)
        # gloo (CPU tests): no reduce-scatter, reduce everything and take the own slot
        dist.all_reduce(self._scatter, op=dist.ReduceOp.SUM, group=self.group)
        self._rs_out[: self._slot].copy_(self._scatter[self.rank * self._slot: (self.rank + 1) * self._slot])
        return None

    def _reduce_scatter_finish(self, work, sigma_loc: torch.Tensor):
        if work is not None:
            work.wait()
        sigma_loc += self._rs_out[: self.local_size]

    def _mark(self, name):
        """phase boundary for profile_phases(): a CUDA event on the current stream"""
        if self._phase_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._phase_events.append((name, ev))

    def profile_phases(self, c_loc, sigma_loc):
        """one build with a CUDA event at every phase boundary; returns [(phase, ms)] (device time on
        this rank's stream, NCCL kernels included)"""
        self._phase_events = []
        self._mark("start")
        self.Hamiltonian(c_loc, sigma_loc)
        torch.cuda.synchronize()
        ev, self._phase_events = self._phase_events, None
        return [(b[0], a[1].elapsed_time(b[1])) for a, b in zip(ev[:-1], ev[1:])]

    def Hamiltonian(self, c_loc: torch.Tensor, sigma_loc: torch.Tensor) -> None:
        assert c_loc.numel() == self.local_size and sigma_loc.numel() == self.local_size
        if self._lib is not None:
            self._lib.Hamiltonian(c_loc, sigma_loc)
            self._mark("f2_dist_apply (collectives issued by the library)")
            return
        if self.form == "rows":
            # the two all-to-alls and the reduce-scatter all run beside kernels that do not need their result:
            #   C column slab  ||  opposite-spin kernel (local rows only)
            #   reduce-scatter ||  H0 + alpha same-spin + beta same-spin
            #   alpha same-spin result back to the row owners  ||  beta same-spin kernel
            cols = self._exchange_cols_start(c_loc)
            self._mark("pack + start of the all-to-all of C (column slab)")
            args = (c_loc, self._xcols, sigma_loc, self._yexch, self._scatter, self.rank, self.world)
            self.engine.apply_rows(*args, phases=1)
            self._mark("opposite-spin kernel (all-to-all of C in flight)")
            work = self._reduce_scatter_start()
            self._exchange_cols_finish(cols)
            self._mark("all-to-all of C: tail + unpack")
            self.engine.apply_rows(*args, phases=4)
            self._mark("H0 + alpha same-spin kernel (reduce-scatter in flight)")
            back = self._exchange_add_start()
            self.engine.apply_rows(*args, phases=8)
            self._mark("beta same-spin kernel (all-to-all back in flight)")
            self._reduce_scatter_finish(work, sigma_loc)
            self._mark("reduce-scatter tail + add")
            self._exchange_add_finish(back, sigma_loc)
            self._mark("all-to-all of the alpha same-spin term: tail + add")
            return
        full = self.all_gather_full(c_loc)
        self._mark("all-gather of C")
        self.engine.apply_part(full, sigma_loc, self._yexch, self.rank, self.world)
        self._mark("local kernels")
        if self.world > 1:
            self._exchange_add(sigma_loc)
            self._mark("all-to-all of the alpha same-spin term + add")


    def Hamiltonian_host(self, c_loc_host, sigma_loc_host) -> None:
        """The local shard from and to host memory (numpy arrays or CPU tensors of local_size doubles; pinned memory makes
        both copies asynchronous DMA): H2D of the shard, the sharded build, D2H of the result, all on the current
        stream, which is synchronised before returning - what a host-resident solver (forte2's numpy Davidson-Liu on
        each rank's slice, forte2/helpers/davidsonliu.py) calls."""
        ch = torch.from_numpy(c_loc_host) if isinstance(c_loc_host, np.ndarray) else c_loc_host
        sh = torch.from_numpy(sigma_loc_host) if isinstance(sigma_loc_host, np.ndarray) else sigma_loc_host
        if ch.numel() != self.local_size or sh.numel() != self.local_size:
            raise ValueError(f"Hamiltonian_host: the local vectors must have {self.local_size} elements")
        if ch.dtype != torch.float64 or sh.dtype != torch.float64 or ch.is_cuda or sh.is_cuda:
            raise ValueError("Hamiltonian_host: float64 host vectors expected")
        if getattr(self, "_host_c", None) is None:
            self._host_c = torch.empty(max(self.local_size, 1), **self._kw)[: self.local_size]
            self._host_s = torch.empty(max(self.local_size, 1), **self._kw)[: self.local_size]
            self._host_copy_stream = torch.cuda.Stream(self.device)
        ch, sh = ch.reshape(-1), sh.reshape(-1)
        if self._streamed_host_ok(ch):
            self._hamiltonian_rows_streamed(ch, self._host_s)
        else:
            self._host_c.copy_(ch, non_blocking=True)
            self.Hamiltonian(self._host_c, self._host_s)
        sh.copy_(self._host_s, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()

    def _streamed_host_ok(self, ch) -> bool:
        """one determinant block in the rows form, pinned source, collectives issued here: the opposite-spin term of the
        rows that have arrived runs while the rest of the shard is still crossing PCIe (F2_NO_STREAMED_SHARD: off)"""
        import os
        nonempty = [b for b in self.layout.blocks if b[1] * b[2] > 0]
        if not (self.form == "rows" and self._lib is None and isinstance(self.engine, CudaEngine) and len(nonempty) == 1
                and len(self.layout.blocks) == 1 and ch.is_pinned() and self.local_size > 0
                and os.environ.get("F2_NO_STREAMED_SHARD") is None):
            return False
        if getattr(self, "_pieces_ok", None) is None:
            try:  # an empty piece: validates that the one-block gather kernel serves this space
                self.engine.apply_rows_t3_piece(self._host_c, self._scatter, self.rank, self.world, 0, 0, 0)
                self._pieces_ok = True
            except (RuntimeError, ValueError):
                self._pieces_ok = False
        return self._pieces_ok

    def _hamiltonian_rows_streamed(self, c_host: torch.Tensor, sigma_loc: torch.Tensor, pieces: int = 4) -> None:
        """Hamiltonian() of the rows form with the H2D copy of the shard inside: the shard is copied in `pieces` row
        pieces on a copy stream; the opposite-spin kernel of piece 0 starts when it has arrived, the all-to-all of the
        column slab when all of them have, beside the opposite-spin kernels of the other pieces."""
        (_, nA, nB), (lo, nr, loc) = self.layout.blocks[0], self.layout.rows[0]
        c_loc = self._host_c
        main = torch.cuda.current_stream(self.device)
        copy = self._host_copy_stream
        copy.wait_stream(main)  # the previous build no longer reads c_loc
        bounds = [nr * i // pieces for i in range(pieces + 1)]
        events = []
        with torch.cuda.stream(copy):
            for i in range(pieces):
                a, b = bounds[i] * nB, bounds[i + 1] * nB
                if b > a:
                    c_loc[a:b].copy_(c_host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
                events.append(ev)
        args = (c_loc, self._xcols, sigma_loc, self._yexch, self._scatter, self.rank, self.world)

        def piece(i):
            flags = (1 if i == 0 else 0) | (2 if i == pieces - 1 else 0)
            self.engine.apply_rows_t3_piece(c_loc, self._scatter, self.rank, self.world, bounds[i], bounds[i + 1] - bounds[i], flags)

        main.wait_event(events[0])
        piece(0)
        main.wait_event(events[-1])
        cols = self._exchange_cols_start(c_loc)
        for i in range(1, pieces):
            piece(i)
        work = self._reduce_scatter_start()
        self._exchange_cols_finish(cols)
        self.engine.apply_rows(*args, phases=4)
        back = self._exchange_add_start()
        self.engine.apply_rows(*args, phases=8)
        self._reduce_scatter_finish(work, sigma_loc)
        self._exchange_add_finish(back, sigma_loc)


class LibraryShardedSigma:
    """The rows form with the collectives issued by the C library itself (f2_dist_init / f2_dist_apply: NCCL resolved by
    the library, two communicators, library-owned streams): what a host without torch.distributed calls.  Here the
    process group is only the channel that carries the 128-byte NCCL unique id from rank 0 to the others."""

    def __init__(self, builder, rank=None, world=None, group=None):
        import ctypes
        self.builder = builder
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        L = _lib.lib()
        uid = ctypes.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(L.f2_dist_unique_id(uid))
        box = [uid.raw if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        _lib.check(L.f2_dist_init(builder._h, box[0], self.rank, self.world))
        n = np.zeros(1, dtype=np.int64)
        _lib.check(L.f2_dist_local_size(builder._h, n.ctypes.data_as(_lib.i64p)))
        self.local_size = int(n[0])
        self.layout = Layout(builder.ci_strings.blocks(), self.rank, self.world)
        assert self.layout.local_size == self.local_size

    def Hamiltonian(self, c_loc: torch.Tensor, sigma_loc: torch.Tensor) -> None:
        assert c_loc.numel() == self.local_size and sigma_loc.numel() == self.local_size
        stream = torch.cuda.current_stream(self.builder.device).cuda_stream
        _lib.check(_lib.lib().f2_dist_apply(self.builder._h, c_loc.data_ptr(), sigma_loc.data_ptr(), stream))

    def allreduce_(self, t: torch.Tensor) -> torch.Tensor:
        stream = torch.cuda.current_stream(self.builder.device).cuda_stream
        _lib.check(_lib.lib().f2_dist_allreduce(self.builder._h, t.data_ptr(), t.numel(), stream))
        return t

    def close(self):
        _lib.check(_lib.lib().f2_dist_finalize(self.builder._h))


def sharded_rdms(sharded: ShardedSigma, builder, left_loc: torch.Tensor, right_loc: torch.Tensor | None = None,
                 kinds=("a_1rdm", "b_1rdm", "aa_2rdm", "bb_2rdm", "ab_2rdm")):
    """RDM spin blocks of sharded vectors: every rank all-gathers the vector(s) into full copies (18.9 GB each at
    CAS(18,18)), computes its slice of the hole strings of every block (f2_sigma_rdm_part: the work, not the vectors, is
    what is split) and the blocks are summed with one all-reduce each.  Returns {kind: numpy array}, the same on every
    rank, shaped as CISigmaBuilder.<kind> returns them."""
    full_l = sharded.all_gather_full(left_loc)
    if right_loc is None or right_loc is left_loc:
        full_r = full_l
    else:
        full_l = full_l.clone()  # all_gather_full reuses its buffer
        full_r = sharded.all_gather_full(right_loc)
    out = {}
    for kind in kinds:
        t = getattr(builder, kind)(full_l, full_r, part=(sharded.rank, sharded.world), device_out=True).contiguous()
        if sharded.world > 1 and t.numel() > 0:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=sharded.group)
        out[kind] = t.cpu().numpy()
    return out
