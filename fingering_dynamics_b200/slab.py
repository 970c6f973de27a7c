"""Slab decomposition along the flow axis (x) across the GPUs of one box: one process per GPU, each
owning global columns [x0, x1) plus two ghost columns per side.  Per step every rank sends its two
edge columns of POST-collision populations (all 18, one contiguous block in the engine's layout) to
each neighbour and receives theirs into its ghost columns, then advances locally -- the only exchange
the path has (the reference is single-process).

Two columns, not one: the collision at an edge cell needs the NEW psi on its one-cell ring, i.e. g
pulled one column further out.  With Zou-He faces no wrap-around message is needed (the wrapped
populations are overwritten on the faces); x-periodic grids (validation.py) close the ring.

Two transports:
  * halo="peer" (default on GPUs): the exchange is FUSED into the step kernel -- it stores its edge columns
    straight into the neighbours' ghost columns through peer-mapped memory (CUDA IPC over NVLink), and the steps
    of neighbouring engines are ordered by stream-side flags; torch.distributed only carries the IPC handles
    once.  No collective library on the data path, no host involvement between steps.
  * halo="nccl": ncclSend/ncclRecv (torch.distributed batch_isend_irecv) on raw device pointers of the engine,
    enqueued on the engine's own CUDA stream; also what the CPU (gloo) protocol tests drive.
"""
import numpy as np


def slab_bounds(W, world, rank):
    """global columns [x0, x1) of `rank`: contiguous, sizes differ by at most one"""
    base, rem = divmod(int(W), int(world))
    x0 = rank * base + min(rank, rem)
    return x0, x0 + base + (1 if rank < rem else 0)


class _DevBuf:
    """raw device memory viewed through __cuda_array_interface__ (uint8)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class SlabRunner:
    """Drives one engine of a slab-decomposed run.  `engine` needs step(n), get_state(...) and either
    halo_regions() (the CUDA Engine) or halo_tensors() (any object handing out torch tensors)."""

    def __init__(self, engine, rank=0, world=1, periodic=False, halo="nccl"):
        self.e, self.rank, self.world, self.periodic = engine, int(rank), int(world), bool(periodic)
        self.left = rank - 1 if rank > 0 else (world - 1 if periodic else None)
        self.right = rank + 1 if rank < world - 1 else (0 if periodic else None)
        self._cache = {}
        self._stream = None
        self.halo = halo if self.world > 1 else "none"
        if self.halo == "peer":
            self._attach_peers()

    def _attach_peers(self):
        """all-gather the engines' peer descriptions (IPC handles) and attach the neighbours"""
        import torch.distributed as dist
        infos = [None] * self.world
        dist.all_gather_object(infos, self.e.peer_export())
        if self.left is not None:
            self.e.peer_attach(0, infos[self.left])
        if self.right is not None:
            self.e.peer_attach(1, infos[self.right])
        dist.barrier()

    def set_state(self, **kw):
        """engine.set_state on every rank, then a barrier: a neighbour's first step writes into this rank's
        ghost columns, which must not happen before this rank has finished loading its state"""
        self._before_load()
        self.e.set_state(**kw)
        self._after_load()

    def _before_load(self):
        """nobody may still be stepping (pushing halo columns) into the lattices a load is about to overwrite"""
        if self.world > 1:
            import torch.distributed as dist
            self.e.sync()
            dist.barrier()

    def _after_load(self):
        if self.world > 1:
            import torch.distributed as dist
            self.e.sync()
            dist.barrier()

    def init_state(self, **kw):
        """engine.init_state (Compute.__init__ on the device) on every rank, then the same barrier as set_state"""
        self._before_load()
        self.e.init_state(**kw)
        self._after_load()

    def checkpoint(self):
        """every rank saves at the same step (the engine waits for its neighbours' halo stores of that step)"""
        return self.e.checkpoint()

    def restore(self, blob):
        """every rank restores a blob of the same step; the barrier keeps a neighbour's first halo stores out of a
        lattice that is still being loaded"""
        self._before_load()
        self.e.restore(blob)
        self._after_load()

    # -- halo buffers ------------------------------------------------------------------------------
    def _tensors(self):
        if hasattr(self.e, "halo_tensors"):
            return self.e.halo_tensors()
        import torch
        h = self.e.halo_regions()
        key = h.recv_lo
        if key not in self._cache:
            dev = torch.device("cuda", torch.cuda.current_device())
            self._cache[key] = tuple(torch.as_tensor(_DevBuf(p, h.bytes), device=dev)
                                     for p in (h.send_lo, h.recv_lo, h.send_hi, h.recv_hi))
        return self._cache[key]

    def _stream_ctx(self):
        import contextlib
        if hasattr(self.e, "halo_tensors"):
            return contextlib.nullcontext()
        import torch
        if self._stream is None:
            self._stream = torch.cuda.ExternalStream(self.e.stream, device=torch.device("cuda", torch.cuda.current_device()))
        return torch.cuda.stream(self._stream)

    def exchange(self):
        """fill the ghost columns of the current state from the neighbours' edge columns"""
        if self.world == 1 or (self.left is None and self.right is None):
            return
        import torch.distributed as dist
        send_lo, recv_lo, send_hi, recv_hi = self._tensors()
        with self._stream_ctx():
            ops = []
            if self.left is not None:
                ops.append(dist.P2POp(dist.isend, send_lo, self.left))
                ops.append(dist.P2POp(dist.irecv, recv_lo, self.left))
            if self.right is not None:
                ops.append(dist.P2POp(dist.isend, send_hi, self.right))
                ops.append(dist.P2POp(dist.irecv, recv_hi, self.right))
            if self.world == 2 and self.periodic:
                # both neighbours are the same peer: order the pairs so that sends and receives match up
                ops = [dist.P2POp(dist.isend, send_lo, self.left), dist.P2POp(dist.irecv, recv_hi, self.right),
                       dist.P2POp(dist.isend, send_hi, self.right), dist.P2POp(dist.irecv, recv_lo, self.left)]
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    # -- run ---------------------------------------------------------------------------------------
    def step(self, n=1):
        if self.world == 1 or self.halo == "peer":
            self.e.step(n)
            return
        for _ in range(int(n)):
            self.exchange()
            self.e.step(1)

    def get_state(self, names=("psi", "rho", "ux", "uy"), **kw):
        if self.halo != "peer":
            self.exchange()
        return self.e.get_state(names, **kw)
