"""Sketch exchange over NVLink peer memory: the allreduce of the lift  Y = avg_g (1/N_loc) X_g^T W_g  fused with the GEMM
that produces it.

Reference: ``MatrixMultCollectiveOperator.matMvMult`` (hippyflow/collectives/collectiveOperator.py:73-80) applies the local
operator and then ``collective.allReduce(y, mpi_op)`` (collective.py:61-71,108-111: one MPI Allreduce per column through
host copies).  Here the lift GEMM's epilogue stores every 128-row tile straight into the exchange buffer of the rank that
owns those rows (reduce-scatter by push over NVLink, ``hfb_dgemm_peer``), the owner sums the P contributions in fixed rank
order, and every rank pulls the reduced blocks of the others (all-gather).  No NCCL call touches the (n x m) block; the
process group is used once, to swap the CUDA IPC handles of the buffers.

The rows are cut into pipeline chunks of whole GEMM waves: while the GEMM of chunk c+1 runs on the main stream, a narrow
reduce + gather of chunk c runs beside it on a side stream (the exchange is NVLink-bound, a few SMs saturate it)."""
import os
import socket
import warnings

import torch

from . import _lib as K

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None

_NT = (4, 9, 10, 14, 16, 17, 18)        # column-tile widths of the DMMA GEMM in 8-column units (csrc/dgemm_api.cu:choose_nt)
_FLAG_BYTES = 256


def _n_tiles(ncols):
    best = None
    for nt in _NT:
        bn = 8 * nt
        tiles = -(-ncols // bn)
        key = (tiles * bn, tiles)
        if best is None or key < best[0]:
            best = (key, tiles)
    return best[1]


def plan_chunks(n, ncols, nchunk, nranks, sms=148):
    """Row chunks [(lo, hi, block_rows)] of an (n x ncols) lift: whole waves of 128-row GEMM tiles per chunk (a chunk that
    ends inside a wave leaves SMs idle at the launch boundary) and, inside a chunk, one block of ``block_rows`` rows
    (a multiple of 128) per owning rank."""
    m_tiles = -(-n // 128)
    nt = _n_tiles(ncols)
    waves = -(-(m_tiles * nt) // sms)
    nchunk = max(1, min(int(nchunk), waves))
    chunks, lo_tile = [], 0
    for c in range(nchunk):
        w = waves // nchunk + (1 if c < waves % nchunk else 0)
        t = (w * sms) // nt
        hi_tile = m_tiles if c == nchunk - 1 else min(m_tiles, lo_tile + t)
        if hi_tile > lo_tile:
            lo, hi = lo_tile * 128, min(n, hi_tile * 128)
            block = -(-(-(-(hi - lo) // nranks)) // 128) * 128
            chunks.append((lo, hi, block))
        lo_tile = hi_tile
    return chunks


class PeerExchange:
    """Exchange buffers of one process group for lifts of up to ``n`` rows with leading dimension ``ld``.

    Construction is COLLECTIVE over the group (IPC handles are all-gathered); ``PeerExchange.create`` returns None on every
    rank when any rank cannot take part (different hosts, no peer access, allocation failure), so the caller falls back to
    the NCCL route on all ranks together."""

    def __init__(self):
        self.base = None
        self.own = None

    # ------------------------------------------------------------------------------------------------ construction
    @classmethod
    def create(cls, group, device, n, ld, ncols, nchunk):
        if dist is None or not dist.is_initialized() or dist.get_backend(group) != "nccl":
            return None
        size, me = dist.get_world_size(group), dist.get_rank(group)
        if size < 2 or size > K.PEER_MAX_RANKS:
            return None
        self = cls()
        self.group, self.size, self.me, self.device = group, size, me, device
        self.n, self.ld, self.ncols = int(n), int(ld), int(ncols)
        self.chunks = plan_chunks(n, ncols, nchunk, size)
        self.block_max = max(b for _, _, b in self.chunks)
        self.region = (size + 1) * self.block_max * ld           # doubles per chunk: P slots + the reduced block
        nbytes = _FLAG_BYTES + len(self.chunks) * self.region * 8
        self.epoch = 0
        self.timeout_s = float(os.environ.get("HFB_PEER_TIMEOUT_S", 600.0))
        ok, handle, err = 1, b"", ""
        try:
            self.own = K.peer_alloc(nbytes)
            handle = K.peer_get_handle(self.own)
        except Exception as e:                                   # noqa: BLE001 -- any failure means "no peer route"
            ok, err = 0, repr(e)
        infos = [None] * size
        dist.all_gather_object(infos, (ok, handle, socket.gethostname(), torch.cuda.current_device()), group=group)
        if ok and not (all(i[0] for i in infos) and len({i[2] for i in infos}) == 1):
            ok = 0
        base = [None] * size
        if ok:
            try:
                for r in range(size):
                    if r == me:
                        base[r] = self.own
                    else:
                        if not torch.cuda.can_device_access_peer(torch.cuda.current_device(), infos[r][3]):
                            raise K.HfbError("no peer access to cuda:%d" % infos[r][3])
                        base[r] = K.peer_open(infos[r][1])
            except Exception as e:                               # noqa: BLE001
                ok, err = 0, repr(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.base = base
        if int(flag.item()) == 0:
            self.close()
            if err:
                warnings.warn("hippyflow_b200: NVLink peer exchange unavailable (%s); using the NCCL route" % err)
            return None
        self.side = torch.cuda.Stream(device=device, priority=-1)
        return self

    @classmethod
    def local_group(cls, nranks, device, n, ld, ncols, nchunk, timeout_s=30.0):
        """``nranks`` exchange objects inside ONE process whose buffers all live on ``device`` -- the same kernels and
        addresses as the multi-process case with every 'peer' pointer local.  Test aid for single-GPU boxes: each emulated
        rank must run its ``lift_allreduce`` on a stream of its own (the barrier kernels of the ranks wait for each other)."""
        chunks = plan_chunks(n, ncols, nchunk, nranks)
        block_max = max(b for _, _, b in chunks)
        region = (nranks + 1) * block_max * ld
        owns = [K.peer_alloc(_FLAG_BYTES + len(chunks) * region * 8) for _ in range(nranks)]
        group = []
        for g in range(nranks):
            self = cls()
            self.group, self.size, self.me, self.device = None, nranks, g, device
            self.n, self.ld, self.ncols = int(n), int(ld), int(ncols)
            self.chunks, self.block_max, self.region = chunks, block_max, region
            self.epoch, self.timeout_s = 0, float(timeout_s)
            self.base, self.own = list(owns), owns[g]
            self.side = torch.cuda.Stream(device=device, priority=-1)
            self._local = True
            group.append(self)
        return group

    def fits(self, n, ld, ncols, nchunk):
        return (self.n, self.ld, self.ncols) == (int(n), int(ld), int(ncols)) and \
            self.chunks == plan_chunks(n, ncols, nchunk, self.size)

    def close(self):
        if self.base is not None:
            for r, p in enumerate(self.base):
                if p is not None and r != self.me and not getattr(self, "_local", False):
                    try:
                        K.peer_close(p)
                    except Exception:                            # noqa: BLE001
                        pass
            self.base = None
        if self.own is not None:
            try:
                K.peer_free(self.own)
            except Exception:                                    # noqa: BLE001
                pass
            self.own = None

    # ------------------------------------------------------------------------------------------------ addresses
    def _slot(self, rank, chunk, slot):
        """Address of slot ``slot`` of chunk ``chunk`` inside rank ``rank``'s buffer (slot == size: the reduced block)."""
        return self.base[rank] + _FLAG_BYTES + 8 * (chunk * self.region + slot * self.block_max * self.ld)

    def _barrier(self):
        self.epoch += 1
        K.peer_barrier(self.base, self.me, self.epoch, self.timeout_s)

    # ------------------------------------------------------------------------------------------------ the exchange
    def lift_allreduce(self, Xt, W, Y, alpha):
        """Y[:, :ncols] = sum over ranks of alpha * Xt^T W  (Xt: (R, n) local rows, W: (R, ncols), Y: (n, >= ncols) view
        with leading dimension self.ld).  COLLECTIVE: every rank of the group calls it with the same shapes."""
        P, me, ld, ncols = self.size, self.me, self.ld, self.ncols
        assert Xt.shape[1] == self.n and W.shape[1] == ncols and Y.shape[0] == self.n and K._ld(Y) >= ncols
        ldy = K._ld(Y)
        main = torch.cuda.current_stream()
        pipelined = len(self.chunks) > 1
        if pipelined:
            self.side.wait_stream(main)                           # Y and the buffers of the previous exchange are free
        for c, (lo, hi, block) in enumerate(self.chunks):
            last = c == len(self.chunks) - 1
            K.dgemm_peer(Xt[:, lo:hi], W, [self._slot(o, c, me) for o in range(P)], block, ld, alpha)
            if pipelined:
                ev = torch.cuda.Event()
                ev.record(main)
                self.side.wait_event(ev)
            with torch.cuda.stream(self.side if pipelined else main):
                self._barrier()                                   # every rank's tiles of this chunk have landed
                own_lo = lo + me * block
                rows = max(0, min(block, hi - own_lo))
                narrow = pipelined and not last
                if rows > 0:
                    K.peer_reduce(self._slot(me, c, 0), self.block_max * ld, P, rows, ncols, ld, self._slot(me, c, P),
                                  Y.data_ptr() + 8 * own_lo * ldy, ldy, max_ctas=8 if narrow else 0)
                self._barrier()                                   # every owner's block is reduced
                K.peer_gather([self._slot(o, c, P) for o in range(P)], me, block, hi - lo, ncols, ld,
                              Y.data_ptr() + 8 * lo * ldy, ldy, ctas_per_peer=max(1, 8 // (P - 1)) if narrow else 0)
        if pipelined:
            main.wait_stream(self.side)
        return Y
