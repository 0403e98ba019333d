"""Multi-GPU host logic: captures are independent, so the batch is partitioned by capture index
with no data-path collective.  One process per GPU; rank r of W owns a
contiguous range of captures.  torch.distributed is used only for the barrier / max-over-ranks
timing in bench.py and for gathering tiny result summaries."""


def shard_range(n_captures, rank, world):
    """Contiguous, balanced partition: the first n % W ranks get one extra capture."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(n_captures, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_shards(n_captures, world):
    return [shard_range(n_captures, r, world) for r in range(world)]


# ---- optional: ONE long capture split in time across ranks ------------
# The only exchange step this path can have: every rank averages the frames of its own time range
# (with the 512-sample FFT overlap as halo), then the 1024 bin sums are added across ranks -- a 4 KiB
# all-reduce over NVLink, once per averaging period.  Demodulators are NOT split this way (carried
# IIR state): they stay one capture per GPU.

def split_frames(n_frames, rank, world):
    """Frames [lo, hi) of a capture owned by `rank`."""
    return shard_range(n_frames, rank, world)


def split_capture_bytes(len_bytes, rank, world, nfft=1024, hop=512):
    """(byte_lo, byte_hi, frame_lo, frame_hi): the slice of the capture rank needs for its frames."""
    n = len_bytes // 2
    frames = 0 if n < nfft else (n - nfft) // hop + 1
    lo, hi = split_frames(frames, rank, world)
    if hi <= lo:
        return 0, 0, lo, hi
    return 2 * hop * lo, 2 * (hop * (hi - 1) + nfft), lo, hi


def allreduce_split_spectrum(local_mean, frames_local, frames_total, dist=None):
    """Combine per-rank mean spectra of a split capture into the capture's mean spectrum.
    `local_mean`: torch tensor (1024,) holding the mean over this rank's frames (what
    b200sdr_batch_spectrum_dev returns for the slice).  With `dist` (torch.distributed, NCCL on the
    GPU box, gloo in the CPU tests) the weighted sums are all-reduced in place."""
    local_mean.mul_(float(frames_local) / float(frames_total) if frames_total else 0.0)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(local_mean)
    return local_mean
