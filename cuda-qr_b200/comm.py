"""Transport adapter of the multi-GPU paths (dist_tsqr.py, dist_caqr.py).

One process per GPU talks through ``torch.distributed``: NCCL over NVLink on a multi-GPU box (device tensors go
straight to ncclSend / ncclRecv / ncclAllGather).  With the gloo backend -- the CPU tests, and the GPU parity tests that
run several ranks on ONE device (tests/test_dist_gpu.py) -- the same calls stage through host memory, so the CUDA
combine / apply steps are exercised unchanged while only the wire differs.  Nothing here computes.
"""
from __future__ import annotations


class Comm:
    def __init__(self, stage_host: bool | None = None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if stage_host is None:
            stage_host = dist.is_initialized() and dist.get_backend() == "gloo"
        self.stage_host = bool(stage_host)
        self.bytes_sent = 0

    # contiguous device tensors only (callers pass the storage of column-major slabs)
    def send(self, t, peer: int) -> None:
        assert t.is_contiguous()
        self.bytes_sent += t.numel() * t.element_size()
        self.dist.send(t.cpu() if self.stage_host and t.is_cuda else t, peer)

    def recv(self, t, peer: int) -> None:
        assert t.is_contiguous()
        if self.stage_host and t.is_cuda:
            h = self.torch.empty(t.shape, dtype=t.dtype)
            self.dist.recv(h, peer)
            t.copy_(h)
        else:
            self.dist.recv(t, peer)

    def broadcast(self, t, src: int) -> None:
        assert t.is_contiguous()
        if self.world == 1:
            return
        if self.stage_host and t.is_cuda:
            h = t.cpu()
            self.dist.broadcast(h, src)
            t.copy_(h)
        else:
            self.dist.broadcast(t, src)

    def all_gather_into(self, out, chunk) -> None:
        """out: (world * chunk.numel()) elements, rank r's chunk at [r * chunk.numel(), ...)."""
        assert out.is_contiguous() and chunk.is_contiguous() and out.numel() == self.world * chunk.numel()
        self.bytes_sent += chunk.numel() * chunk.element_size() * (self.world - 1)
        if self.world == 1:
            out.view(-1).copy_(chunk.reshape(-1))
        elif self.stage_host and chunk.is_cuda:
            parts = [self.torch.empty(chunk.shape, dtype=chunk.dtype) for _ in range(self.world)]
            self.dist.all_gather(parts, chunk.cpu())
            out.view(self.world, -1).copy_(self.torch.stack([p.reshape(-1) for p in parts]))
        else:
            self.dist.all_gather_into_tensor(out.view(-1), chunk.reshape(-1))

    def all_reduce_sum(self, t):
        if self.world == 1:
            return t
        if self.stage_host and t.is_cuda:
            h = t.cpu()
            self.dist.all_reduce(h)
            t.copy_(h)
        else:
            self.dist.all_reduce(t)
        return t

    def all_reduce_max(self, t):
        if self.world == 1:
            return t
        if self.stage_host and t.is_cuda:
            h = t.cpu()
            self.dist.all_reduce(h, op=self.dist.ReduceOp.MAX)
            t.copy_(h)
        else:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t

    def barrier(self) -> None:
        if self.world > 1:
            self.dist.barrier()
