"""Multi-GPU sharding of the STFT hot path.  One process per GPU; the hot path needs
NO collective:

  * batches of independent signals -> contiguous row ranges per rank (`row_range`);
  * one very long signal, forward  -> contiguous slice (frame) ranges per rank; rank r needs the
    samples [p_a*hop - m_num_mid, (p_b-1)*hop - m_num_mid + m_num) of the signal, i.e. neighbouring
    ranks overlap by an (m_num - hop)-sample halo (`frame_shard`), and calls stft(p0=p_a, p1=p_b)
    on its sub-signal with k_offset shifted so that slice indices stay global;
  * gathering the per-rank results into one array is the only place a collective
    (`torch.distributed.all_gather` / `gather`, NCCL on GPUs, gloo in the CPU tests) is used.

The compute callables are injected (`stft_fn`), so the CPU tests can drive the very same
partitioning / gathering logic with the oracle standing in for the CUDA kernels.
"""
from dataclasses import dataclass


def row_range(batch: int, rank: int, world: int):
    """Contiguous, balanced row range of `rank` (first `batch % world` ranks get one extra row)."""
    base, rem = divmod(batch, world)
    a = rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


@dataclass
class FrameShard:
    p_a: int          # first global slice index of this rank
    p_b: int          # one past the last
    s_a: int          # first signal sample this rank needs (clipped to [0, n))
    s_b: int          # one past the last
    k_offset: int     # k_offset to pass to stft() on x[s_a:s_b] so slice p keeps its global meaning
    halo: int         # samples shared with the next rank (m_num - hop when slices abut)


def frame_shard(n: int, m_num: int, hop: int, p_min: int, p_max: int, rank: int, world: int) -> FrameShard:
    """Slice range and sample range (with halo) of `rank` for a single signal of n samples.

    Slice p covers samples [p*hop - mid, p*hop - mid + m_num) (reference src/lib.rs:617-619).
    Calling stft(x[s_a:s_b], p0=p_a, p1=p_b, k_offset=-s_a) reproduces the global slices exactly:
    zero padding outside [0, n) is preserved because the local signal is only ever clipped at the
    global borders."""
    mid = m_num // 2
    a, b = row_range(p_max - p_min, rank, world)
    p_a, p_b = p_min + a, p_min + b
    s_a = max(0, p_a * hop - mid)
    s_b = min(n, (p_b - 1) * hop - mid + m_num) if p_b > p_a else s_a
    s_b = max(s_b, s_a)
    return FrameShard(p_a, p_b, s_a, s_b, -s_a, max(0, m_num - hop))


def sharded_stft_rows(x_rows, stft_fn):
    """Forward transform of this rank's rows; nothing is exchanged."""
    return stft_fn(x_rows)


def gather_rows(local, batch: int, dist, rank: int, world: int, dst: int = 0):
    """Gather row shards (torch tensors [rows_r, ...]) to `dst` in rank order -> [batch, ...] or None.
    Uses all_gather on equal-size padded shards (works with both NCCL and gloo)."""
    import torch
    rows_max = -(-batch // world)
    pad = torch.zeros((rows_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        a, b = row_range(batch, r, world)
        out.append(parts[r][: b - a])
    return torch.cat(out, dim=0)


def gather_frames(local, p_total: int, dist, rank: int, world: int, dst: int = 0):
    """Gather slice-range shards S_r[F, P_r] of one long signal along the time axis -> S[F, P]."""
    import torch
    p_maxr = -(-p_total // world)
    F = local.shape[0]
    pad = torch.zeros((F, p_maxr), dtype=local.dtype, device=local.device)
    pad[:, : local.shape[1]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        a, b = row_range(p_total, r, world)
        out.append(parts[r][:, : b - a])
    return torch.cat(out, dim=1)
