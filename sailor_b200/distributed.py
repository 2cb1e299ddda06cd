"""Multi-GPU sharding of one frame: one process per GPU, scene and BVH replicated, the frame split by
PRIMARY-SAMPLE RANGE or by IMAGE ROWS, and ONE collective per frame on the fp32 accumulators.

The path shards trivially: pixels x primary samples are independent and the random streams are keyed by
(seed, pixel, primary sample), so every shard computes exactly what the unsharded render computes for its part
(tests: test_render_is_deterministic_and_partition_invariant, tests/test_distributed.py).

  sample split  accumulator = sum over primary samples / msaa (PathTracer.cpp:457-469) is linear, so each rank renders
                its sample range against the full `msaa` and the frame is the SUM of the per-rank accumulators:
                one reduce(SUM) to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).
  row split     each rank renders a band of task rows; the other rows of its accumulator are zero, so the same
                reduce(SUM) assembles the frame (a gather would move 1/N of the bytes; the reduce keeps one code path
                and the payload, <= 133 MB at 4K, is far below NVLink's per-frame budget).

Never split the inner S samples of one first hit: MIS and the clamps act on per-hit averages (PathTracer.cpp:838-852).
The output stage (aberration taps cross shard borders) runs on rank 0 after the reduce.
"""
import copy


def split_range(n, world, rank):
    """Contiguous, balanced [begin, end) of n items for `rank` of `world` (first n % world ranks get one more)."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_params(params, rank, world, mode="samples", height=None):
    """Copy of `params` restricted to this rank's shard. mode: 'samples' (primary-sample range) or 'rows'."""
    p = copy.copy(params)
    if world <= 1:
        return p
    if mode == "samples":
        if params.m_msaa < world:
            raise ValueError("msaa %d < world size %d: use mode='rows'" % (params.m_msaa, world))
        p.msaa_range = split_range(params.m_msaa, world, rank)
    elif mode == "rows":
        if height is None:
            raise ValueError("row sharding needs the image height")
        p.rows = split_range(height, world, rank)
    else:
        raise ValueError(mode)
    return p


def reduce_accumulators(acc, dst=0):
    """The one collective of a frame: SUM of the per-rank fp32 accumulators (a torch tensor, cuda for NCCL / cpu for
    gloo) onto rank `dst`. No-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM)
    return acc
