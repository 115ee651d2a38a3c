"""Sketch exchange over NVLink peer memory: the allreduce of the lift  Y = avg_g (1/N_loc) X_g^T W_g  fused with the GEMM
that produces it.

Reference: ``MatrixMultCollectiveOperator.matMvMult`` (hippyflow/collectives/collectiveOperator.py:73-80) applies the local
operator and then ``collective.allReduce(y, mpi_op)`` (collective.py:61-71,108-111: one MPI Allreduce per column through
host copies).  Here the lift GEMM's epilogue stores every 128-row tile straight into the exchange buffer of the rank that
owns those rows (reduce-scatter by push over NVLink, ``hfb_dgemm_peer``); the owner sums the P contributions in fixed rank
order and stores the sums into its rows of EVERY rank's result block (all-gather by push, ``hfb_peer_reduce_bcast``).  The
result block lives inside the exchange buffer and is handed to the solver as the sketch itself -- no copy out.  No NCCL call
touches the (n x m) block; the process group is used once, to swap the CUDA IPC handles of the buffers.

Measured on 2 x B200 at the cfg2 shard (profiles/r02_peer_lift_2gpu.json): GEMM alone 16.36 ms, fused lift + exchange
17.23 ms, lift in four row blocks with NCCL allreduces 17.28 ms.  Cutting the lift into pipeline chunks whose reduce + gather
run on a side stream beside the next chunk's GEMM was measured SLOWER (18.1 ms at four chunks: a GEMM chunk sized to whole
waves of 148 SMs needs one more wave once the side kernels hold a few SMs), so the exchange is one launch sequence."""
import os
import socket
import warnings

import torch

from . import _lib as K

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None

_FLAG_BYTES = 256


def block_rows_for(n, nranks):
    """Rows per owning rank: ceil(n / nranks) rounded up to the 128-row GEMM tile, so a CTA's tile has ONE destination."""
    return -(-(-(-int(n) // int(nranks))) // 128) * 128


def release_all(exchanges, group):
    """COLLECTIVE retirement of exchange objects: every rank unmaps its peers, a barrier makes sure nobody still addresses a
    remote buffer, and the buffers themselves are freed when their last torch view dies."""
    for ex in exchanges:
        ex.unmap()
    if dist is not None and dist.is_initialized():
        dist.barrier(group=group)


class PeerExchange:
    """Exchange buffer of one rank (flags | P slots | two result blocks used alternately, so the sketch of the previous
    exchange stays valid while the next one is formed) plus the peer-mapped addresses of the other ranks' buffers, for lifts
    with ``n`` rows, ``ncols`` columns and leading dimension ``ld``.

    Construction is COLLECTIVE over the group (IPC handles are all-gathered); ``PeerExchange.create`` returns None on every
    rank when any rank cannot take part (different hosts, no peer access, allocation failure), so the caller falls back to
    the NCCL route on all ranks together."""

    def __init__(self, size, me, device, n, ld, ncols, timeout_s):
        self.size, self.me, self.device = int(size), int(me), device
        self.n, self.ld, self.ncols = int(n), int(ld), int(ncols)
        self.block = block_rows_for(n, size)
        self.slot_elems = self.block * self.ld                      # one slot = one owner block
        self.result_elems = self.size * self.block * self.ld        # one result block = all owner blocks (>= n rows)
        self.nbytes = _FLAG_BYTES + 8 * (self.size * self.slot_elems + 2 * self.result_elems)
        self.parity = 0
        self.epoch = 0
        self.timeout_s = float(timeout_s)
        self.base = None          # base address of every rank's buffer as seen from this process
        self.own = None           # this rank's allocation (freed by close())
        self.local = False        # True: all buffers live in this process (single-GPU emulation), nothing to unmap

    # ------------------------------------------------------------------------------------------------ construction
    @classmethod
    def create(cls, group, device, n, ld, ncols):
        if dist is None or not dist.is_initialized() or dist.get_backend(group) != "nccl":
            return None
        size, me = dist.get_world_size(group), dist.get_rank(group)
        if size < 2 or size > K.PEER_MAX_RANKS:
            return None
        self = cls(size, me, device, n, ld, ncols, os.environ.get("HFB_PEER_TIMEOUT_S", 600.0))
        ok, handle, err = 1, b"", ""
        try:
            self.own = K.peer_alloc(self.nbytes)
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
        return self

    @classmethod
    def local_group(cls, nranks, device, n, ld, ncols, timeout_s=20.0):
        """``nranks`` exchange objects inside ONE process whose buffers all live on ``device``: the same kernels and address
        arithmetic as the multi-process case with every 'peer' pointer local.  Test aid for single-GPU boxes; drive the
        phases (``push`` / ``signal`` / ``wait`` / ``reduce_bcast``) rank by rank on one stream."""
        group = [cls(nranks, g, device, n, ld, ncols, timeout_s) for g in range(nranks)]
        owns = [K.peer_alloc(group[0].nbytes) for _ in range(nranks)]
        for g, ex in enumerate(group):
            ex.base, ex.own, ex.local = list(owns), owns[g], True
        return group

    def unmap(self):
        """Close the mappings of the other ranks' buffers (this rank's own buffer stays: torch views may still use it)."""
        if self.base is not None and not self.local:
            for r, p in enumerate(self.base):
                if p is not None and r != self.me:
                    try:
                        K.peer_close(p)
                    except Exception:                            # noqa: BLE001
                        pass
        self.base = None

    def close(self):
        """Unmap the peers and free this rank's buffer.  Views handed out by ``lift_allreduce`` must be dead; when that is not
        known, drop the object instead -- the views keep it alive and ``__del__`` frees the buffer after the last one."""
        self.unmap()
        if self.own is not None:
            try:
                K.peer_free(self.own)
            except Exception:                                    # noqa: BLE001
                pass
            self.own = None

    def __del__(self):
        try:
            self.close()
        except Exception:                                        # noqa: BLE001 -- interpreter shutdown
            pass

    # ------------------------------------------------------------------------------------------------ addresses
    def _slot(self, rank, slot):
        """Address of slot ``slot`` inside rank ``rank``'s buffer."""
        return self.base[rank] + _FLAG_BYTES + 8 * slot * self.slot_elems

    def _result(self, rank, parity):
        """Address of result block ``parity`` inside rank ``rank``'s buffer."""
        return self.base[rank] + _FLAG_BYTES + 8 * (self.size * self.slot_elems + parity * self.result_elems)

    def result_view(self, parity=None):
        """This rank's result block as an (n, ncols) torch view (leading dimension ld) of the exchange buffer."""
        parity = self.parity if parity is None else parity
        return K.tensor_from_ptr(self._result(self.me, parity), self.n, self.ld, self.ld, self.device, owner=self)[:, :self.ncols]

    # ------------------------------------------------------------------------------------------------ phases
    def push(self, Xt, W, alpha):
        """Lift GEMM alpha * Xt^T W whose epilogue stores the rows of block o into slot ``me`` of rank o's buffer."""
        assert Xt.shape[1] == self.n and W.shape[1] == self.ncols
        K.dgemm_peer(Xt, W, [self._slot(o, self.me) for o in range(self.size)], self.block, self.ld, alpha)

    def signal(self):
        self.epoch += 1
        K.peer_barrier(self.base, self.me, self.epoch, self.timeout_s, K.PEER_SIGNAL)

    def wait(self):
        K.peer_barrier(self.base, self.me, self.epoch, self.timeout_s, K.PEER_WAIT)

    def barrier(self):
        self.epoch += 1
        K.peer_barrier(self.base, self.me, self.epoch, self.timeout_s)

    def reduce_bcast(self):
        """Owner's fixed-order sum of its P slots, stored into its rows of every rank's current result block."""
        lo = self.me * self.block
        rows = max(0, min(self.block, self.n - lo))
        if rows > 0:
            K.peer_reduce_bcast(self._slot(self.me, 0), self.slot_elems, self.me, rows, self.ncols, self.ld,
                                [self._result(r, self.parity) + 8 * lo * self.ld for r in range(self.size)], self.ld)

    # ------------------------------------------------------------------------------------------------ the exchange
    def lift_allreduce(self, Xt, W, alpha):
        """sum over ranks of alpha * Xt^T W  (Xt: (R, n) local rows, W: (R, ncols)) as an (n, ncols) view of this rank's
        exchange buffer; it stays valid until the exchange after the next one.  COLLECTIVE: every rank of the group calls it
        with the same shapes; asynchronous on the current stream."""
        self.parity ^= 1
        self.push(Xt, W, alpha)
        self.barrier()            # every rank's tiles have landed in the owners' slots
        self.reduce_bcast()
        self.barrier()            # every owner's rows have landed in every result block
        return self.result_view()
