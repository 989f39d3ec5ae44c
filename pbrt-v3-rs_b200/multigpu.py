"""Multi-GPU decomposition of Integrator::render.

The reference renders 16x16 tiles from a work queue on one host
(core/src/integrator/sampler_integrator.rs:252-296).  Here the scene is replicated on every GPU, the
pixel rows are cut into bands dealt round-robin to the ranks (interleaving balances sky and geometry),
each rank renders its bands into a zero-initialised film of the full window, and ONE collective — a sum
all-reduce of the {X, Y, Z, weight} film (33 MB at 1080p) over NCCL / NVLink — assembles the image.  With a
box filter the rows are disjoint, so the sum only fills zeros; with wider filters the aprons add up.
Nothing else crosses GPUs: the path has no data-path collective.
"""
import numpy as np

BAND_ROWS = 8


def shard_rows(height, n_shards, shard, band_rows=BAND_ROWS):
    """Pixel rows of `shard`: bands of `band_rows` rows dealt round-robin (mirrors b200pt_render_shard_device)."""
    rows = []
    for band, r0 in enumerate(range(0, height, band_rows)):
        if band % n_shards == shard:
            rows.extend(range(r0, min(height, r0 + band_rows)))
    return np.asarray(rows, dtype=np.int64)


def reduce_film(film, group=None):
    """Sum all-reduce of a (H, W, 4) film tensor across the process group (NCCL on GPUs, gloo in CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(film, op=dist.ReduceOp.SUM, group=group)
    return film


def render_distributed(integrator, band_rows=BAND_ROWS):
    """Renders this rank's shard on the current CUDA device and all-reduces the film.  Returns the (H, W, 4)
    XYZ+weight film as a CUDA tensor, identical on every rank."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if integrator.handle is None:
        integrator.preprocess()
    h, w = integrator.film_shape()
    film = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    integrator.render_shard_device(rank, world, film.data_ptr(), band_rows, torch.cuda.current_stream().cuda_stream)
    return reduce_film(film)
