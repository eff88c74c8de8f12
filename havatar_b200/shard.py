"""Sharding of the render path across the GPUs of one box: one process per GPU, no data-path collective.

Rays are independent given (planes, skinning volume, MLP weights), frames are independent given the weights
(SURVEY.md section 8e), so the path partitions two ways:
  * frames_of_rank: a batch of frames is split frame-wise (training / multi-frame inference; planes stay local);
  * ray_band_of_rank: one frame's R rays are split into contiguous bands aligned to the kernel's 128-ray tile, so
    every rank's band is a whole number of tiles except the last.
The only communication is optional result collection (gather) and, for training, the gradient all-reduce."""
TILE = 128


def frames_of_rank(num_frames, world, rank):
    """Contiguous frame range [lo, hi) of `rank`; the first num_frames % world ranks get one extra frame."""
    base, extra = divmod(num_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def ray_band_of_rank(num_rays, world, rank, tile=TILE):
    """Contiguous ray range [lo, hi) of `rank`, boundaries on multiples of `tile` (ragged tail on the last rank)."""
    tiles = (num_rays + tile - 1) // tile
    lo_t, hi_t = frames_of_rank(tiles, world, rank)
    return min(lo_t * tile, num_rays), min(hi_t * tile, num_rays)
