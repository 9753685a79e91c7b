"""Data-parallel gradient exchange -- the B200 replacement for ``xm.optimizer_step``'s all-reduce
(reference stage_1_train_fn.py:149,166-172).

One process per GPU (torchrun); every optimizer owns ONE flat fp32 gradient buffer
(``layers.FlatParams``), all-reduced in buckets: a bucket is handed to NCCL on a side stream as soon
as the backward pass has finished writing it (the tail of the critic's flat buffer -- ds4 + head, 75 %
of its parameters -- is complete after the first tenth of the backward pass), the rest of the
backward overlaps with the transfer over NVLink/NVSwitch, and the fused Adam kernel waits on the
bucket events.  Semantics are the reference's: gradients are AVERAGED over replicas right before the
step, BatchNorm statistics stay per replica, parameters are broadcast from rank 0 once at start
(train.py:78-85).

With the gloo backend (CPU tests) the same calls run synchronously.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class DistComm:
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
