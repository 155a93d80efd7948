"""Multi-GPU sharding of the render path: one process per GPU (torch.distributed, NCCL over NVLink).

The path shards by independent units: rays of a frame by contiguous pixel rows,
pi-GAN latents by batch index, density-grid points by linear index.  No collective sits on the data
path; the only exchanges are the final image gather and, in training, ONE all-reduce of the flat
fp32 gradient bucket.  The reference's incumbent is single-process nn.DataParallel
(pi_GAN/train.py:50-52); nerf/ has no multi-GPU path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int, align: int = 1) -> tuple[int, int]:
    """Contiguous block [begin, begin+count) of n units for `rank`; blocks differ by at most
    `align` units and the remainder goes to the last ranks' predecessors evenly."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    blocks = (n + align - 1) // align
    base, rem = divmod(blocks, world)
    b0 = rank * base + min(rank, rem)
    cnt = base + (1 if rank < rem else 0)
    begin = min(b0 * align, n)
    end = min((b0 + cnt) * align, n)
    return begin, end - begin


def gather_image(rgb: torch.Tensor, depth: torch.Tensor, acc: torch.Tensor, out: torch.Tensor, n_rays: int,
                 rank: int, world: int, group=None) -> torch.Tensor:
    """All-gather each rank's [count,5] = (rgb, depth, acc) rows into `out`[n_rays,5] (every rank
    receives the frame).  When the three maps already ARE this rank's rows of `out` (the composite kernel wrote them there:
    render_image_device(fine_out=frame_slice(out, ...))) and the blocks are equal, the collective runs IN PLACE on `out` with no
    pack or staging copy at all; otherwise the rows are packed first, and ragged blocks are padded to the largest."""
    begin, count = shard_range(n_rays, rank, world)
    mine = out[begin:begin + count]
    in_place = (rgb.data_ptr() == mine.data_ptr() and depth.data_ptr() == mine.data_ptr() + 12 and acc.data_ptr() == mine.data_ptr() + 16
                and rgb.stride(0) == 5 and depth.stride(0) == 5 and acc.stride(0) == 5) if count > 0 else True
    if not in_place:
        mine = torch.cat([rgb, depth[:, None], acc[:, None]], -1).contiguous()
    if world == 1:
        if not in_place:
            out[: mine.shape[0]].copy_(mine)
        return out
    counts = [shard_range(n_rays, r, world)[1] for r in range(world)]
    if min(counts) == max(counts):
        # equal blocks: the collective writes every rank's rows straight into the final buffer (in place when `mine` is a slice of it)
        dist.all_gather_into_tensor(out, mine, group=group)
        return out
    # ragged blocks: pad to the largest, gather, unpack
    cmax = max(counts)
    padded = mine.new_zeros((cmax, 5))
    padded[: mine.shape[0]] = mine
    stage = mine.new_empty((world * cmax, 5))
    dist.all_gather_into_tensor(stage, padded, group=group)
    for r in range(world):
        b, c = shard_range(n_rays, r, world)
        out[b:b + c] = stage[r * cmax:r * cmax + c]
    return out


def frame_slice(out: torch.Tensor, n_rays: int, rank: int, world: int) -> torch.Tensor:
    """This rank's rows of the gathered [n_rays,5] frame buffer: pass it as ``fine_out`` so that the composite kernel writes
    (rgb, depth, acc) where the gather expects them (nerf/render.py:161-166 concatenates on the host instead)."""
    begin, count = shard_range(n_rays, rank, world)
    return out[begin:begin + count]


def render_image_sharded(width, height, focal, pose, near, far, coarse_model, fine_model, coarse_sample_num,
                         fine_sample_num, *, t_rand_full: torch.Tensor | None = None, precision=None, group=None,
                         exact_last_sample=None):
    """render_image (nerf/render.py:150-167) with the frame's rays sharded over the ranks of
    `group`; returns (rgb[H,W,3], depth[H,W,1], acc[H,W,1]) CUDA tensors on every rank.
    Every rank draws the SAME global jitter (same seed) and slices its rows, so the result does not
    depend on the number of ranks."""
    from . import nerf_render
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = int(width) * int(height)
    begin, count = shard_range(n, rank, world)
    dev = next(coarse_model.parameters()).device
    if t_rand_full is None:
        t_rand_full = nerf_render._draw_t_rand(n, int(coarse_sample_num), nerf_render.REFERENCE_RAY_CHUNK, dev)
    with torch.no_grad():
        out = torch.empty((n, 5), dtype=torch.float32, device=dev)
        o = nerf_render.render_image_device(width, height, focal, pose, near, far, coarse_model, fine_model,
                                            coarse_sample_num, fine_sample_num, ray_begin=begin, ray_count=count,
                                            t_rand=t_rand_full[begin:begin + count], precision=precision,
                                            exact_last_sample=exact_last_sample, fine_out=frame_slice(out, n, rank, world),
                                            coarse_outputs_unused=True)
        gather_image(o[3], o[4], o[5], out, n, rank, world, group)
    h, w = int(height), int(width)
    return out[:, :3].reshape(h, w, 3), out[:, 3].reshape(h, w, 1), out[:, 4].reshape(h, w, 1)


def flat_grad_bucket(models_: list) -> torch.Tensor:
    """ONE flat fp32 tensor holding the gradients of all given models' parameters (2 x 593,924
    floats = 4.75 MB for the NeRF pair), in canonical parameter order."""
    gs = []
    for m in models_:
        for p in m.parameters():
            gs.append((p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1))
    return torch.cat(gs)


def allreduce_gradients(models_: list, group=None, average: bool = True) -> None:
    """Gradient exchange of the training configuration (C3): one NCCL all-reduce of the flat bucket,
    then scatter back into the .grad tensors."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    bucket = flat_grad_bucket(models_)
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    if average:
        bucket /= dist.get_world_size(group)
    off = 0
    for m in models_:
        for p in m.parameters():
            n = p.numel()
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(bucket[off:off + n].view_as(p))
            off += n
