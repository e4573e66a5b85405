"""Multi-GPU host logic (one process per GPU, torch.distributed): the two ways the path shards.

1. Independent frame pairs (BASELINE config 4): pair i belongs to rank i mod world -- no data-path collective, one
   gather of the per-pair results at the end.
2. One oversized cloud (BASELINE config 5): every rank holds a contiguous slice of the SOURCE and a replica of the target;
   per ICP pass the 29 normal-equation sums (21 JtJ + 6 Jtr + |C| + sum d2) are all-reduced in rank order (an NCCL all-gather
   over NVLink on GPUs, gloo in the CPU tests, then a fixed-order sum) and every rank applies the same 6x6 solve.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N

ICP_SUMS = 29


# ---- partitioning (pure host logic, tested on CPU) ---------------------------------------------------------------------
def pair_indices(n_pairs, rank, world):
    """Round-robin ownership of independent pairs."""
    return list(range(rank, n_pairs, world))


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n source points owned by rank (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_pair_results(local, n_pairs, rank, world, group=None):
    """local: {pair index: 18-vector (16 transform + fitness + rmse)} -> [n_pairs, 18] on every rank."""
    import torch.distributed as dist
    rows = np.zeros((n_pairs, 18), np.float64)
    for i, v in local.items():
        rows[i] = v
    t = torch.from_numpy(rows)
    if world > 1:
        backend = dist.get_backend(group)
        if backend == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)  # ownership is disjoint, so the sum is a gather
        t = t.cpu()
    return t.numpy()


def all_reduce_sums(sums, group=None):
    """In-place sum of the 29-double tensor over the ranks (no-op without an initialised process group), added IN RANK ORDER:
    an all-gather of the ranks' vectors (232 bytes each) and `((r0 + r1) + r2) + ...` on every rank. A library all-reduce adds in
    the order of its ring / tree, which changes with the world size and the algorithm NCCL picks; the fixed order makes the
    result independent of both and bit-identical to the exchange fused into the pass kernel (b3d_icp_pass_peers adds the
    peers' slots in rank order too) -- at 4 ranks a plain all_reduce differed from it in the last bit."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        parts = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(parts, sums, group=group)
        acc = parts[0]
        for r in range(1, world):
            acc = acc + parts[r]
        sums.copy_(acc)
    return sums


def icp_loop(accumulate, update, all_reduce=all_reduce_sums, max_passes=1000, check_every=1):
    """The sharded ICP driver: accumulate() -> tensor of 29 sums (local shard), all-reduce, update() -> done flag.
    Backend-agnostic so that the CPU tests can drive it with a stand-in shard; returns the number of passes.
    check_every > 1: the done flag is read back (a host synchronisation) only every check_every-th pass; the passes in
    between are enqueued blindly, which is safe because a finished state ignores further accumulate / update calls and every
    rank sees the same all-reduced sums, hence the same flag."""
    for k in range(max_passes):
        sums = accumulate()
        all_reduce(sums)
        look = check_every <= 1 or (k % check_every) == check_every - 1
        if update(look) if check_every > 1 else update():
            return k + 1
    raise RuntimeError("sharded ICP did not terminate")


# ---- device side ---------------------------------------------------------------------------------------------------------
class _DevArray:
    """Wraps a raw device pointer for torch.as_tensor via __cuda_array_interface__."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class ShardedICP:
    """Step-wise ICP over one shard of the source (b3d_icp_begin / accumulate / update / finish)."""

    def __init__(self, kind, src_local, ns_total, tgt, max_dist, tgt_normals=None, src_cov=None, tgt_cov=None, init=None, rel_fitness=1e-6,
                 rel_rmse=1e-6, max_iter=30, device=0):
        from .context import get_context, ptr
        from .ops import _T16
        self.ctx = get_context(device)
        c = self.ctx
        self._keep = [c.to_device(src_local, torch.float64), c.to_device(tgt, torch.float64)]
        self._keep += [None if a is None else c.to_device(a, torch.float64) for a in (tgt_normals, src_cov, tgt_cov)]
        s, t, tn, sc, tc = self._keep
        self.n_local = s.shape[0]
        h = C.c_void_p()
        N.check(N.lib().b3d_icp_begin(c.handle, int(kind), ptr(s), s.shape[0], int(ns_total), ptr(sc), ptr(t), t.shape[0], ptr(tn), ptr(tc),
                                      float(max_dist), _T16(init), float(rel_fitness), float(rel_rmse), int(max_iter), C.byref(h)))
        self.handle = h

    def accumulate(self):
        """Correspondences + partial normal equations of this shard. Returns the CUDA tensor [29] to all-reduce in place."""
        p = C.c_void_p()
        N.check(N.lib().b3d_icp_accumulate(self.ctx.handle, self.handle, C.byref(p)))
        return torch.as_tensor(_DevArray(p.value, ICP_SUMS), device=self.ctx.device)

    def update(self, look=True):
        """Applies the update from the (all-reduced) sums. look=False skips reading the done flag back (no host sync)."""
        if not look:
            N.check(N.lib().b3d_icp_update(self.ctx.handle, self.handle, None))
            return False
        done = C.c_int(0)
        N.check(N.lib().b3d_icp_update(self.ctx.handle, self.handle, C.byref(done)))
        return bool(done.value)

    def finish(self):
        from .context import ptr
        from .ops import _result_dict
        corr = self.ctx.empty((max(self.n_local, 1),), torch.int32)
        r = N.IcpResult()
        N.check(N.lib().b3d_icp_finish(self.ctx.handle, self.handle, C.byref(r), ptr(corr)))
        self.handle = None
        return _result_dict(r, corr[:self.n_local].cpu().numpy())

    # ---- fused exchange over peer memory (NVLink): no collective launch between the reduction and the solve ------------
    def enable_peers(self, group=None):
        """Allocates this rank's exchange buffer in symmetric memory and maps every peer's buffer (rendezvous over the
        process group). After this, run_fused() advances the loop with ONE kernel launch per pass on every rank."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > 8:
            raise RuntimeError("the fused exchange supports up to 8 ranks (one NVSwitch domain)")
        n = 2 * world * 32 + 2 * world
        buf = symm_mem.empty(n, dtype=torch.float64, device=self.ctx.device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, group=(group or dist.group.WORLD))
        torch.cuda.synchronize(self.ctx.device)
        dist.barrier(group)
        # every rank's buffer as a tensor view in THIS process (peer mapping over NVLink); touching it checks the mapping
        views = [hdl.get_buffer(r, (n,), torch.float64) for r in range(world)]
        for v in views:
            assert float(v.sum().item()) == 0.0
        # nobody may launch its first fused pass (it writes into every peer's buffer) while a slower rank is still checking
        dist.barrier(group)
        ptrs = (C.c_void_p * world)(*[C.c_void_p(v.data_ptr()) for v in views])
        N.check(N.lib().b3d_icp_set_peers(self.ctx.handle, self.handle, rank, world, ptrs))
        self._peer = (buf, hdl, views)  # keep the mapping alive

    def pass_fused(self, look=True):
        if not look:
            N.check(N.lib().b3d_icp_pass_peers(self.ctx.handle, self.handle, None))
            return False
        done = C.c_int(0)
        N.check(N.lib().b3d_icp_pass_peers(self.ctx.handle, self.handle, C.byref(done)))
        return bool(done.value)

    def run_fused(self, check_every=2, max_passes=1000):
        for k in range(max_passes):
            if self.pass_fused(look=(k % check_every) == check_every - 1):
                return self.finish()
        raise RuntimeError("sharded ICP did not terminate")

    def run(self, group=None, check_every=2):
        icp_loop(self.accumulate, self.update, lambda s: all_reduce_sums(s, group), check_every=check_every)
        return self.finish()


def register_pairs_distributed(depth_src, depth_tgt, params, rank, world, device=0, group=None):
    """Config 4: this rank registers its round-robin share of the pairs in one batched call; results gathered on all ranks."""
    from . import ops
    n_pairs = depth_src.shape[0]
    mine = pair_indices(n_pairs, rank, world)
    local = {}
    if mine:
        res = ops.register_depth_pairs(depth_src[mine], depth_tgt[mine], params, device=device)
        for i, r in zip(mine, res):
            local[i] = np.concatenate([r["transformation"].reshape(16), [r["fitness"], r["inlier_rmse"]]])
    return gather_pair_results(local, n_pairs, rank, world, group)
