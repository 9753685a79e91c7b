"""Data-parallel gradient exchange -- the B200 replacement for ``xm.optimizer_step``'s all-reduce
(reference stage_1_train_fn.py:149,166-172; stage_2_train_fn.py:155,164,167).

One process per GPU (torchrun).  Semantics are the reference's: gradients are AVERAGED over replicas right before each
optimizer step, BatchNorm statistics stay per replica, parameters are broadcast from rank 0 once at start
(train.py:78-85).

Two transports:

``PeerComm`` (NCCL process group + NVLink peer access; what ``make_comm`` returns on a multi-GPU box)
    Every optimizer's flat parameter and gradient buffers (``layers.FlatParams``) are allocated in SYMMETRIC memory
    (``torch.distributed._symmetric_memory``: cuMem allocations whose handles are exchanged once, so each rank holds a
    device pointer to every peer's copy).  The optimizer step is then ONE kernel of this package, ``sg_dp_adam_step``
    (csrc/dp_adam.cu): reduce-scatter by P2P loads of the rank's shard, Adam on that shard (the optimizer state is sharded
    over ranks), all-gather by P2P stores, cross-GPU ordering by release/acquire flags in symmetric memory.  No NCCL call
    on the step's path, nothing to cut the CUDA graph of the step at, replicas bit-identical by construction.
    ``torch.distributed`` is used for rendezvous, the initial broadcast and checkpoint-time gathers only.

``DistComm`` (any backend; gloo in the CPU tests)
    One ``all_reduce`` of the flat gradient buffer per optimizer step, then the local fused Adam.  With NCCL it is issued
    on a side stream between two segments of the captured step (``engine._SegmentedGraph``).
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist


class DistComm:
    peer = False

    def __init__(self, device=None, group=None):
        assert dist.is_available() and dist.is_initialized(), "init_process_group first"
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nccl = dist.get_backend(group) == "nccl"
        self.device = device
        self.side = torch.cuda.Stream(device=device) if self.nccl else None
        self.pending = []
        self.bytes_reduced = 0

    def broadcast_params(self, flat):
        """pjrt.broadcast_master_param (train.py:78-85): rank 0's parameters everywhere."""
        dist.broadcast(flat, 0, group=self.group)

    def allreduce_async(self, t):
        """Average ``t`` (a contiguous slice of a flat gradient buffer) over replicas.  NCCL: issued on
        the side stream after everything already queued on the current stream; gloo: synchronous."""
        self.bytes_reduced += t.numel() * t.element_size()
        if self.world == 1:
            return
        if not self.nccl:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world)
            return
        if os.environ.get("SG_COMM_NOOP") == "1":          # debugging aid: segmentation cost without the transfers
            return
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
            done = torch.cuda.Event()
            done.record(self.side)
        self.pending.append(done)

    def wait_all(self):
        """Make the current stream wait for every outstanding bucket."""
        if self.nccl:
            cur = torch.cuda.current_stream(self.device)
            for ev in self.pending:
                cur.wait_event(ev)
        self.pending = []


class PeerComm(DistComm):
    """Symmetric-memory transport (see the module docstring).  ``alloc`` hands out zeroed fp32 buffers every peer can
    address; ``step`` launches the fused reduce-scatter / Adam / all-gather kernel for one ``FlatParams``."""

    peer = True

    def __init__(self, ops, device=None, group=None):
        super().__init__(device=device, group=group)
        import torch.distributed._symmetric_memory as symm
        self.ops, self.symm = ops, symm
        self.pg = group if group is not None else dist.group.WORLD
        lib = ops.lib
        assert self.world <= lib.sg_dp_max_world(), f"world size {self.world} > {lib.sg_dp_max_world()}"
        self._handles = {}                  # data_ptr of a symmetric tensor -> rendezvous handle
        self._tables = {}                   # data_ptr -> ctypes array of the peers' base pointers
        self.flags = self.alloc(lib.sg_dp_flag_ints(), torch.int32)
        self.sync = torch.zeros(lib.sg_dp_sync_ints(), dtype=torch.int32, device=device)
        self._slots = {}
        torch.cuda.synchronize(device)
        dist.barrier(group=group)           # every flag block is zero before anyone may signal into it
        self.write_avg = False              # tests: also store the averaged gradient into every replica's .grad

    def alloc(self, numel, dtype=torch.float32):
        """A zeroed symmetric buffer.  COLLECTIVE: every rank must allocate the same sizes in the same order."""
        t = self.symm.empty(int(numel), dtype=dtype, device=self.device)
        hdl = self.symm.rendezvous(t, self.pg)
        t.zero_()
        self._handles[t.data_ptr()] = hdl
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == t.data_ptr(), (ptrs, t.data_ptr(), self.rank)
        self._tables[t.data_ptr()] = (ctypes.c_void_p * self.world)(*ptrs)
        return t

    def table(self, t):
        return self._tables[t.data_ptr()]

    def slot_of(self, fp):
        if id(fp) not in self._slots:
            self._slots[id(fp)] = len(self._slots)
        return self._slots[id(fp)]

    def step(self, fp):
        """xm.optimizer_step for one optimizer: gradient mean over replicas + Adam, in one kernel on the current stream."""
        self.bytes_reduced += fp.grad.numel() * 4
        self.ops.dp_adam_step(self.table(fp.grad), self.table(fp.flat), self.table(self.flags), fp.m, fp.v, fp.hyper, self.sync,
                              fp.flat.numel(), self.rank, self.world, self.slot_of(fp), self.write_avg)

    def check(self):
        """Raises if a peer ever failed to answer a flag poll within the kernel's time-out (synchronises)."""
        if int(self.sync[-1].item()) != 0:
            raise RuntimeError("sg_dp_adam_step: a peer did not reach the optimizer step within the time-out "
                               "(replicas out of step, or a rank died)")

    def gather_state(self, fp):
        """The optimizer state is sharded: rank r holds exp_avg / exp_avg_sq of elements [r*chunk, (r+1)*chunk) only.  Before a
        checkpoint every rank calls this (COLLECTIVE) so that ``fp.m`` / ``fp.v`` are whole on all of them."""
        n = fp.flat.numel()
        chunk = ((n + self.world - 1) // self.world + 3) // 4 * 4
        for buf in (fp.m, fp.v):
            padded = torch.zeros(chunk * self.world, dtype=buf.dtype, device=buf.device)
            lo = self.rank * chunk
            hi = min(n, lo + chunk)
            mine = torch.zeros(chunk, dtype=buf.dtype, device=buf.device)
            if hi > lo:
                mine[:hi - lo] = buf[lo:hi]
            dist.all_gather_into_tensor(padded, mine, group=self.group)
            buf.copy_(padded[:n])


def make_comm(ops, device=None, group=None):
    """The transport for this process group: ``PeerComm`` when the ranks can address each other's memory (NCCL backend on
    one NVLink box, CUDA kernels available), else ``DistComm``.  ``SG_DP_TRANSPORT=nccl`` forces the all-reduce path."""
    if not (dist.is_available() and dist.is_initialized()):
        return None
    if dist.get_world_size(group) == 1:
        return None
    if (dist.get_backend(group) == "nccl" and not getattr(ops, "is_emulator", False)
            and os.environ.get("SG_DP_TRANSPORT", "peer") != "nccl"):
        return PeerComm(ops, device=device, group=group)
    return DistComm(device=device, group=group)
