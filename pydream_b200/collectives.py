"""The two exchange steps of the sharded sampler, kept free of CUDA specifics so the
layout logic is testable with the gloo backend on CPU.  On B200 the process group is NCCL and the
tensors live in HBM: the all-gather runs over NVLink 5 / NVSwitch directly into the archive tail."""
import torch.distributed as dist


def allgather_rows(block, c0, nlocal, group):
    """block: [N, ld] view of the archive tail for one append; this rank has written rows
    [c0, c0+nlocal).  In-place all-gather: afterwards every rank holds all N rows in chain order."""
    if group is None or dist.get_world_size(group) == 1:
        return
    mine = block[c0:c0 + nlocal]
    dist.all_gather_into_tensor(block.view(-1), mine.reshape(-1), group=group)


def allreduce_sum(t, group):
    if group is None or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(t, group=group)
