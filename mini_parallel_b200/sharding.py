"""Multi-GPU host logic: pairs are independent, so a job is a contiguous block split of the pair index space, one
process per GPU, and the only exchange is the reduction of the timing / checksum scalars.  No data-path
collective exists; torch.distributed (NCCL on the GPU box, gloo in the CPU tests) is plumbing for these scalars."""


def shard_range(rank, world, n_pairs):
    """Pairs [lo, hi) of rank `rank` out of `world`: contiguous, disjoint, covering, sizes differ by at most one."""
    if not (0 <= rank < world) or n_pairs < 0:
        raise ValueError("bad shard request")
    return rank * n_pairs // world, (rank + 1) * n_pairs // world


def reduce_scalars(values, op="max"):
    """All-reduce a list of Python floats over the default process group (no-op without one).  Works on whichever
    device the backend needs: CUDA tensors under NCCL, CPU tensors under gloo."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op])
    return [float(x) for x in t.tolist()]


def gather_results(local, n_total, rank, world):
    """Optional result gather (north_star: 'NCCL is used only for the final score and position gather, if at all'):
    every rank contributes its shard of swb_result records (an (n,3) int32 tensor), every rank receives all n_total."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    sizes = [shard_range(r, world, n_total)[1] - shard_range(r, world, n_total)[0] for r in range(world)]
    cap = max(sizes)
    pad = torch.zeros((cap, 3), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
