"""Multi-GPU decomposition of Integrator::render.

The reference renders 16x16 tiles from a work queue on one host
(core/src/integrator/sampler_integrator.rs:252-296).  Here the scene is replicated on every GPU, the
pixel rows are cut into bands dealt in snake order to the ranks (interleaving balances sky and geometry, the snake cancels a gradient),
each rank renders its bands into a zero-initialised film of the full window, and ONE collective over NCCL / NVLink
assembles the image: with a box-sized filter every rank ships only the bands it owns (plus one spill row per band, see
`spill_rows`) to rank 0 (a gather of ~1 / world of the film per rank); with wider filters the aprons overlap and the films are summed (all-reduce).
Nothing else crosses GPUs: the path has no data-path collective.  This module is the one-process-per-GPU (torchrun) form;
one process driving all GPUs goes through b200pt_multi_render (csrc/multi_gpu.cu, `MultiGPURender`).
"""
import numpy as np

BAND_ROWS = 8


def band_owner(band, n_shards):
    """b200pt_band_owner (include/b200pt.h): bands are dealt in snake order 0 1 .. n-1 n-1 .. 1 0 0 1 .., which cancels a cost
    gradient down the image within every pair of passes."""
    from . import lib
    return int(lib().b200pt_band_owner(int(band), int(n_shards)))


def shard_rows(height, n_shards, shard, band_rows=BAND_ROWS):
    """Pixel rows of `shard`: bands of `band_rows` rows dealt in snake order (mirrors b200pt_render_shard_device)."""
    rows = []
    for band, r0 in enumerate(range(0, height, band_rows)):
        if band_owner(band, n_shards) == shard:
            rows.extend(range(r0, min(height, r0 + band_rows)))
    return np.asarray(rows, dtype=np.int64)


def reduce_film(film, group=None):
    """Sum all-reduce of a (H, W, 4) film tensor across the process group (NCCL on GPUs, gloo in CPU tests): the
    general form, needed when the filter is wider than a pixel and neighbouring bands overlap by its apron."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(film, op=dist.ReduceOp.SUM, group=group)
    return film


def spill_rows(height, n_shards, shard, band_rows=BAND_ROWS):
    """Rows just ABOVE each band of `shard` (band index >= 1): a sample whose film position has a zero fractional part in y
    also lands in the pixel row above its own (film_tile.rs:73-76: p0 = ceil(p - 0.5 - radius) includes it), so the first
    sample row of a band can contribute to the last pixel row of the previous band - which another shard owns."""
    return np.asarray([r0 - 1 for band, r0 in enumerate(range(0, height, band_rows))
                       if band >= 1 and band_owner(band, n_shards) == shard and band_owner(band - 1, n_shards) != shard], dtype=np.int64)


class BandGather:
    """Gather of OWNED bands for box-sized filters (radius <= 0.5 px): a rank's samples reach its own rows and, for the
    rare sample that sits exactly on a pixel boundary, the one row above each of its bands (`spill_rows`).  Each rank
    ships just those rows (1 / world of the film plus one row per band) to rank 0 instead of all-reducing the whole film:
    one `dist.gather` over NCCL / NVLink, then rank 0 drops the bands into place and ADDS the spill rows (in the order
    the single-device film kernel adds them: the row's own samples first).  Index tensors and staging buffers are built
    once (outside any timed region)."""

    def __init__(self, height, width, device, band_rows=BAND_ROWS, group=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rows = [torch.from_numpy(shard_rows(height, self.world, r, band_rows)).to(device) for r in range(self.world)]
        self.spill = [torch.from_numpy(spill_rows(height, self.world, r, band_rows)).to(device) for r in range(self.world)]
        self.max_rows = max(int(r.numel()) for r in self.rows)
        self.max_spill = max(int(r.numel()) for r in self.spill)
        self.mine = torch.zeros((self.max_rows + self.max_spill, width, 4), dtype=torch.float32, device=device)
        self.parts = [torch.zeros_like(self.mine) for _ in range(self.world)] if self.rank == 0 else None

    def __call__(self, film):
        """film: this rank's (H, W, 4) shard film; on rank 0 it holds the whole image afterwards."""
        import torch.distributed as dist
        if self.world == 1:
            return film
        my, sp = self.rows[self.rank], self.spill[self.rank]
        self.mine[:my.numel()].copy_(film.index_select(0, my))
        if sp.numel():
            self.mine[self.max_rows:self.max_rows + sp.numel()].copy_(film.index_select(0, sp))
        dist.gather(self.mine, self.parts, dst=0, group=self.group)
        if self.rank == 0:
            for r in range(1, self.world):
                film.index_copy_(0, self.rows[r], self.parts[r][:self.rows[r].numel()])
            for r in range(self.world):  # rank 0's own spill rows were staged in self.mine before the bands overwrote them
                if self.spill[r].numel():
                    src = self.mine if r == 0 else self.parts[r]
                    film.index_add_(0, self.spill[r], src[self.max_rows:self.max_rows + self.spill[r].numel()])
        return film


def render_distributed(integrator, band_rows=BAND_ROWS, gather=None):
    """Renders this rank's shard on the library's device and assembles the film: with a box-sized filter the owned bands
    are gathered on rank 0 (BandGather; other ranks keep their shard), otherwise the films are all-reduced.  Returns the
    (H, W, 4) XYZ+weight film as a CUDA tensor."""
    import torch
    import torch.distributed as dist
    from . import current_device, init
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if current_device() < 0:
        init(torch.cuda.current_device())
    if current_device() != torch.cuda.current_device():
        raise RuntimeError("render_distributed: torch's current device (%d) is not the library's device (%d); call pkg.init(local_rank) after torch.cuda.set_device(local_rank)"
                           % (torch.cuda.current_device(), current_device()))
    if integrator.handle is None:
        integrator.preprocess()
    h, w = integrator.film_shape()
    film = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    if world == 1:
        integrator.render_shard_device(rank, world, film.data_ptr(), band_rows, sp)
        return film
    # shard films are combined as running sums (RGB + weight) and converted to XYZ once, after the collective: the
    # assembled film is then the single-device film bit for bit (see b200pt_render_shard_device_raw)
    integrator.render_shard_device_raw(rank, world, film.data_ptr(), band_rows, sp)
    if integrator.filter_radius()[1] <= 0.5:
        film = (gather or BandGather(h, w, film.device, band_rows))(film)
        if rank == 0:
            integrator.film_finish_device(film.data_ptr(), h * w, sp)
        return film
    film = reduce_film(film)
    integrator.film_finish_device(film.data_ptr(), h * w, sp)
    return film
