"""Multi-GPU sharding of the STFT hot path.  One process per GPU; the hot path needs
NO collective:

  * batches of independent signals -> contiguous row ranges per rank (`row_range`);
  * one very long signal, forward  -> contiguous slice (frame) ranges per rank; rank r needs the
    samples [p_a*hop - m_num_mid, (p_b-1)*hop - m_num_mid + m_num) of the signal, i.e. neighbouring
    ranks overlap by an (m_num - hop)-sample halo (`frame_shard`), and calls stft(p0=p_a, p1=p_b)
    on its sub-signal with k_offset shifted so that slice indices stay global;
  * one very long signal, inverse  -> every rank owns the output samples whose LATEST contributing
    slice is one of its own (`inverse_shard`); it needs the last ceil(m_num/hop)-1 columns of S
    from the previous rank -- ONE neighbour exchange (`exchange_halo_columns`, point-to-point, a few
    KB) -- and then runs the ordinary istft on [halo | own] columns with shifted (k0, k1): the
    overlap-add is local and in the same increasing-q order as the unsharded transform
    (src/lib.rs:853), hence bit-identical (exchanging partial sums instead would change the order);
  * gathering the per-rank results into one array is the only place a collective
    (`torch.distributed.gather` to the destination rank only; NCCL on GPUs, gloo in the CPU tests) is used.

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


def _gather_to(pad, dist, rank, world, dst):
    """dist.gather of equal-shape tensors to `dst` only; complex tensors travel as (re, im) pairs
    (the gloo backend has no complex gather)."""
    import torch
    cplx = pad.is_complex()
    t = torch.view_as_real(pad).contiguous() if cplx else pad
    parts = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, parts, dst=dst)
    if rank != dst:
        return None
    return [torch.view_as_complex(p) for p in parts] if cplx else parts


def gather_rows(local, batch: int, dist, rank: int, world: int, dst: int = 0):
    """Gather row shards (torch tensors [rows_r, ...]) to `dst` in rank order -> [batch, ...] or None.
    dist.gather of equal-size padded shards: only `dst` receives (NCCL and gloo)."""
    import torch
    rows_max = -(-batch // world)
    pad = torch.zeros((rows_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = _gather_to(pad, dist, rank, world, dst)
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
    parts = _gather_to(pad, dist, rank, world, dst)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        a, b = row_range(p_total, r, world)
        out.append(parts[r][:, : b - a])
    return torch.cat(out, dim=1)


@dataclass
class InverseShard:
    p_a: int          # first global slice owned by this rank
    p_b: int          # one past the last
    halo_cols: int    # columns of S needed from the previous rank (0 for rank 0)
    k_a: int          # global output samples [k_a, k_b) this rank produces
    k_b: int
    k0_local: int     # (k0, k1) to pass to istft() on the [halo | own] columns
    k1_local: int


def inverse_halo_cols(m_num: int, hop: int, p_min: int) -> int:
    """Columns of S a rank needs from its predecessor: the ceil(m_num/hop)-1 overlapping slices, and
    at least ceil(mid/hop) - p_min so that the local k0 = (halo + p_min)*hop - mid stays >= 0."""
    mid = m_num // 2
    return max((m_num + hop - 1) // hop - 1, (mid + hop - 1) // hop - p_min)


def inverse_shard(p_total: int, m_num: int, hop: int, p_min: int, rank: int, world: int,
                  k0: int = 0, k1=None, halo_align: int = 1) -> InverseShard:
    """Output range and halo of `rank` for the inverse transform of S[F, p_total] sharded by slices.

    Slice q adds to samples [q*hop - mid, q*hop - mid + m_num) (src/lib.rs:876-910).  A sample's
    latest contributing slice is floor((k + mid) / hop): rank r with slices [p_a, p_b) therefore owns
    k in [p_a*hop - mid, p_b*hop - mid) (first rank: from k0; last rank: up to k1, default the
    reference's k_max, src/lib.rs:823-824) and needs the ceil(m_num/hop)-1 slices before p_a.
    On the local column block the slice index restarts at p_min, so (k0, k1) shift by
    (p_a - halo_cols - p_min) * hop.

    The halo is widened where needed so that the local k0 is never negative: for k0 < 0 the
    reference starts at q0 = floor(k0/hop) WITHOUT adding p_min (src/lib.rs:842-846), which on a
    local block would skip the first halo columns (hops that do not divide m_num/2, e.g. 1024/100).
    The extra columns lie below the local q0 and contribute nothing.  `halo_align` rounds the halo up
    to a multiple (2 keeps the [halo | own] block of a pitch-aligned S buffer 16-byte aligned)."""
    mid = m_num // 2
    a, b = row_range(p_total, rank, world)
    p_a, p_b = p_min + a, p_min + b
    k_max = (p_min + p_total - 1) * hop + m_num - mid
    k1 = k_max if k1 is None else k1
    need = -(-inverse_halo_cols(m_num, hop, p_min) // halo_align) * halo_align
    halo = 0 if rank == 0 else min(need, a)
    k_a = k0 if rank == 0 else max(k0, p_a * hop - mid)
    k_b = k1 if rank == world - 1 else min(k1, p_b * hop - mid)
    shift = (p_a - halo - p_min) * hop
    return InverseShard(p_a, p_b, halo, k_a, max(k_a, k_b), k_a - shift, max(k_a, k_b) - shift)


def exchange_halo_columns(local, halo_next: int, halo_mine: int, dist, rank: int, world: int):
    """The one data-path exchange of the sharded inverse: send my last `halo_next` columns of
    S_r[F, P_r] to rank+1, receive `halo_mine` columns from rank-1; returns [halo | own] columns.
    Point-to-point (NCCL send/recv over NVLink on GPUs, gloo in the CPU tests)."""
    import torch
    reqs = []
    recv = None
    if rank + 1 < world and halo_next > 0:
        assert local.shape[1] >= halo_next, "a rank must own at least ceil(m_num/hop)-1 slices"
        tail = local[:, local.shape[1] - halo_next:].contiguous()
        reqs.append(dist.isend(tail, rank + 1))
    if rank > 0 and halo_mine > 0:
        recv = torch.empty((local.shape[0], halo_mine), dtype=local.dtype, device=local.device)
        reqs.append(dist.irecv(recv, rank - 1))
    for r in reqs:
        r.wait()
    return local if recv is None else torch.cat([recv, local], dim=1)


def gather_samples(local, k0: int, k1: int, shards, dist, rank: int, world: int, dst: int = 0):
    """Gather the per-rank output segments y_r[k_a:k_b) into y[k0:k1) on `dst` (rank order)."""
    import torch
    n_max = max(s.k_b - s.k_a for s in shards)
    pad = torch.zeros((n_max,), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = _gather_to(pad, dist, rank, world, dst)
    if rank != dst:
        return None
    return torch.cat([parts[r][: shards[r].k_b - shards[r].k_a] for r in range(world)], dim=0)
