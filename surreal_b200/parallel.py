"""Data-parallel learner plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink / NVSwitch on the
GPU box, gloo in the CPU tests) for the few exchanges the path needs:

  * ONE flat gradient all-reduce per optimiser step (the flat parameter layout makes it a single collective),
  * the moments of the advantage normalisation (global batch: ppo.py:413-416),
  * the mean-KL scalar that drives the early stop (so every rank takes the same branch),
  * the z-filter sums, and the reporting statistics.

The reference has no collective at all (its learner is single-process, README.md:24); actors / windows / replay
shards are independent, so nothing on the rollout side communicates."""
import ctypes as C
import os

import torch
import torch.distributed as dist


class PeerChannel:
    """One-shot all-reduce over NVLink peer memory (csrc/peer_allreduce.cu) for the ranks of one box: every rank allocates
    a symmetric buffer, the CUDA-IPC handles travel through ``torch.distributed`` once, after that an all-reduce is ONE
    kernel launch (publish slice -> wait for the peers' flags -> sum the slots in rank order).  Messages here are
    <= ~1.3 MB (the flat gradient of head + stem); NCCL remains the transport for broadcasts and anything bigger."""

    def __init__(self, group, max_floats):
        from . import _lib
        self._lib = _lib
        L = _lib.lib()
        _lib.ensure_device()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        assert self.world <= 8
        self.max_floats = (int(max_floats) + 3) // 4 * 4
        own = C.c_void_p()
        handle = (C.c_char * 64)()
        _lib.check(L.sb200_par_alloc(self.max_floats, C.byref(own), handle), 'sb200_par_alloc')
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.ctx = _lib.Par()
        self.ctx.world, self.ctx.rank, self.ctx.max_floats = self.world, self.rank, self.max_floats
        self._opened = []
        for p in range(self.world):
            if p == self.rank:
                self.ctx.peers[p] = own.value
            else:
                ptr = C.c_void_p()
                buf = (C.c_char * 64).from_buffer_copy(handles[p])
                _lib.check(L.sb200_par_open(buf, C.byref(ptr)), 'sb200_par_open')
                self.ctx.peers[p] = ptr.value
                self._opened.append(ptr.value)
        self._own = own.value
        dist.barrier(group=group)                                  # every rank has mapped every buffer before first use

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def allreduce_(self, t, scale=1.0, opt_ws=None, bump_step=0, stop_flag=None, out=None):
        L = self._lib.lib()
        out = t if out is None else out
        n = t.numel()
        if t.dtype == torch.float64:
            self._lib.check(L.sb200_par_allreduce_f64(C.byref(self.ctx), C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()), n,
                                                      float(scale), self._st()), 'sb200_par_allreduce_f64')
        else:
            assert t.dtype == torch.float32 and n <= self.max_floats
            self._lib.check(L.sb200_par_allreduce_f32(
                C.byref(self.ctx), C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()), n, float(scale), int(bump_step),
                C.c_void_p(opt_ws.data_ptr()) if opt_ws is not None else None,
                C.c_void_p(stop_flag.data_ptr()) if stop_flag is not None else None, self._st()), 'sb200_par_allreduce_f32')
        return out


class LearnerDP:
    """Exchanges of the data-parallel learner.  ``peer_floats`` > 0 routes all-reduces through a PeerChannel (one kernel over
    NVLink peer memory per exchange); otherwise (or with SB200_PEER_ALLREDUCE=0, and always under gloo in the CPU tests)
    they are ``torch.distributed`` collectives."""

    def __init__(self, group=None, peer_floats=0):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.peer = None
        if peer_floats > 0 and self.world > 1 and dist.get_backend(group) == 'nccl' and \
                os.environ.get('SB200_PEER_ALLREDUCE', '1') != '0':
            self.peer = PeerChannel(group, peer_floats)

    def broadcast_(self, t, src=0):
        dist.broadcast(t, src=src, group=self.group)
        return t

    def sum_(self, t):
        if self.peer is not None and t.is_cuda and t.dtype in (torch.float32, torch.float64) and t.is_contiguous() \
                and t.numel() * (2 if t.dtype == torch.float64 else 1) <= self.peer.max_floats:
            return self.peer.allreduce_(t)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def mean_(self, t):
        if self.peer is not None and t.is_cuda and t.dtype in (torch.float32, torch.float64) and t.is_contiguous() \
                and t.numel() * (2 if t.dtype == torch.float64 else 1) <= self.peer.max_floats:
            return self.peer.allreduce_(t, scale=1.0 / self.world)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        t.div_(self.world)
        return t

    def reduce_grad_(self, grad, opt_ws, stop_flag=None):
        """grad <- mean over ranks; global norm + step count into the optimiser workspace.  True when done in ONE fused
        peer kernel; False when the caller must run sb200_grad_reduce_norm_f32 on the summed gradient itself."""
        if self.peer is not None and grad.numel() <= self.peer.max_floats:
            self.peer.allreduce_(grad, scale=1.0 / self.world, opt_ws=opt_ws, bump_step=1, stop_flag=stop_flag)
            return True
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group)
        return False


def combine_moments(moments):
    """[sum, sumsq, count] (already summed over ranks) -> (mean, unbiased std): what sb200_normalize_f32 applies."""
    s, q, n = float(moments[0]), float(moments[1]), float(moments[2])
    mean = s / n
    var = (q - s * s / n) / (n - 1.0)
    return mean, max(var, 0.0) ** 0.5
