"""Sharding of a batch of counterfactual prompts over the GPUs of one box.

The path is embarrassingly parallel over the sample axis: every rank holds a full replica of the predictor, takes a
contiguous slice of the (already rectangularised) ``(x, mask)`` batch, runs the forward with no data-path
collective, and one final gather brings the predicted frames (or any per-sample statistic) together.  NCCL over
NVLink on the GPU box; the same code runs on ``gloo`` for the CPU tests of the host logic.
"""
import torch
import torch.distributed as dist


def shard_bounds(num_samples, rank, world_size):
    """Contiguous, balanced slice [lo, hi) of ``num_samples`` for ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(num_samples, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_samples(x, mask, rank=None, world_size=None):
    """Slice of the sample axis owned by this rank.  Rectangularise the masks BEFORE sharding (the reference's
    ``RectangularizeMasks`` draws from a global RNG, masking.py:119-128, so it must see the whole batch)."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(x.shape[0], rank, world_size)
    return x[lo:hi], mask[lo:hi], (lo, hi)


def gather_samples(y_local, num_samples, dst=None):
    """Inverse of ``shard_samples`` along dim 0.  ``dst=None`` -> all ranks get the full tensor (all_gather),
    otherwise only rank ``dst`` does (others get None).  Equal shards land directly in their slice of the result
    (no staging copy); shards that differ in size by one sample are padded to the largest before the collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return y_local
    world = dist.get_world_size()
    rank = dist.get_rank()
    bounds = [shard_bounds(num_samples, r, world) for r in range(world)]
    sizes = [hi - lo for lo, hi in bounds]
    max_n = max(sizes)
    receives = dst is None or rank == dst
    if min(sizes) == max_n:
        y_local = y_local.contiguous()
        full = y_local.new_empty((num_samples,) + tuple(y_local.shape[1:])) if receives else None
        out = [full[lo:hi] for lo, hi in bounds] if receives else None
        if dst is None:
            dist.all_gather(out, y_local)
        else:
            dist.gather(y_local, out, dst=dst)
        return full
    pad = y_local
    if y_local.shape[0] < max_n:
        pad = torch.cat([y_local, y_local.new_zeros((max_n - y_local.shape[0],) + tuple(y_local.shape[1:]))], 0)
    pad = pad.contiguous()
    out = [torch.empty_like(pad) for _ in range(world)] if receives else None
    if dst is None:
        dist.all_gather(out, pad)
    else:
        dist.gather(pad, out, dst=dst)
        if rank != dst:
            return None
    return torch.cat([o[:n] for o, n in zip(out, sizes)], 0)


def sharded_predict(generator, x, mask, frame=-1, batch_size=None, gather=True, **kwargs):
    """``PredictorBasedGenerator.batch_predict_per_sample(sample_dim=0)`` with the sample axis sharded over the
    ranks of the default process group; returns the full [S, ...] prediction on every rank when ``gather``."""
    S = x.shape[0]
    xs, ms, _ = shard_samples(x, mask)
    y = generator.batch_predict_per_sample(xs, ms, frame=frame, batch_size=batch_size, sample_dim=0, **kwargs) \
        if xs.shape[0] > 0 else None
    if y is None:  # a rank with no samples still has to join the collective with the right trailing shape
        T = 1 if frame is not None else x.shape[1]
        y = x.new_zeros((0, T) + tuple(x.shape[2:]), dtype=torch.float32)
    return gather_samples(y, S) if gather else y


def sharded_counterfactual_videos(generator, x, active_patches, passive_patches=None, shifts=None, num_samples=8,
                                  sample_batch_size=8, fix_passive=True, frame=1, gather=True, dst=None,
                                  predict_frame=None, **kwargs):
    """``FlowGenerator.predict_counterfactual_videos`` (cwm/models/segmentation.py:345-430) with the S samples of the
    sweep sharded over the ranks (SURVEY.md section 8e + 8f rank 1).

    Every rank holds the image and the (tiny) patch descriptors.  Rank 0 builds and rectangularises the masks of the
    whole sweep -- the rectangulariser draws from a global RNG (masking.py:119-128), so it must see all rows once --
    and broadcasts them (S*N bytes); each rank then predicts its contiguous slice from the *virtual* counterfactual
    video (nothing but descriptors is ever sent) and the predicted movies are gathered once.
    Returns ``[S, T, C, H, W]`` on every rank (``dst=None``), on rank ``dst`` only, or the local slice (``gather=False``).
    ``predict_frame`` (extension) keeps only that frame of every predicted movie (``-1``: the counterfactual frame,
    ``[S, 1, C, H, W]``) so that the gather moves the predicted frames and not the prompts' frame 0 with them."""
    G = generator
    multi = dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if multi else 0
    world = dist.get_world_size() if multi else 1
    if len(x.shape) == 3:
        x = x.unsqueeze(0).unsqueeze(1).expand(-1, 2, -1, -1, -1)
        fix_passive = True
    elif len(x.shape) == 4:
        x = x.unsqueeze(1).expand(-1, 2, -1, -1, -1)
        fix_passive = True
    elif len(x.shape) == 5 and x.size(1) == 1:
        x = x.expand(-1, 2, -1, -1, -1)
    x = x[:, 0:2]
    G.set_input(x)
    G.reset_shifts()
    if passive_patches is None:
        passive_patches = G.get_zeros_mask().unsqueeze(-1)
    elif len(passive_patches.shape) == 2:
        passive_patches = passive_patches.unsqueeze(-1)
    if len(active_patches.shape) == 2:
        active_patches = active_patches.unsqueeze(-1)
    S = max(active_patches.size(-1), passive_patches.size(-1))
    if S == 1 and num_samples > 1:
        S = num_samples
    G.shifter.set_shapes(x, mask=active_patches[..., 0])
    G.shifter.set_num_shifts(S if shifts is None else (len(shifts) if not hasattr(shifts, 'shape') else shifts.shape[-1]))
    drawn = shifts is None
    shifts = G.shifter._preprocess_shifts_sequence(shifts, is_mask_shift=True)
    if multi and drawn:  # randomly drawn shifts must be the same sweep on every rank (caller-given ones already are)
        obj = [[list(map(int, s)) for s in shifts]] if rank == 0 else [None]
        dist.broadcast_object_list(obj, src=0)
        shifts = obj[0]
    S = len(shifts)
    if active_patches.size(-1) == 1 and S > 1:
        active_patches = active_patches.expand(-1, -1, S)
    if passive_patches.size(-1) == 1 and S > 1:
        passive_patches = passive_patches.expand(-1, -1, S)
    video, masks = G.create_motion_counterfactuals(x, masks=passive_patches, active_patches=active_patches,
                                                   shifts=shifts, num_samples=S, fix_passive=fix_passive,
                                                   reset_shifts=False, frame=frame, virtual=True)
    if multi and not getattr(G, 'device_masks', False):
        # host rectangulariser: global RNG, so rank 0's result is THE result.  With device_masks=True the rectangulariser
        # is a pure function of (seed, row): every rank already holds the same masks and nothing is exchanged.
        m8 = masks.contiguous().view(torch.uint8)
        dist.broadcast(m8, src=0)
        masks = m8.view(torch.bool)
    n_total = masks.shape[0]
    lo, hi = shard_bounds(n_total, rank, world)
    if hi > lo:
        y = G.batch_predict_per_sample(video[lo:hi], masks=masks[lo:hi], frame=predict_frame,
                                       batch_size=(sample_batch_size or (hi - lo)), sample_dim=0, **kwargs)
    else:
        T = x.shape[1] if predict_frame is None else 1
        y = x.new_zeros((0, T) + tuple(x.shape[2:]), dtype=torch.float32)
    if not gather:
        return y
    return gather_samples(y, n_total, dst=dst)


def sharded_counterfactual_motion_map(generator, x, active_patches, passive_patches=None, shifts=None, num_samples=8,
                                      sample_batch_size=8, fix_passive=True, frame=1, raft_iters=None, backward=False,
                                      do_filter=True, normalize_per_sample=False, return_local=False, **kwargs):
    """One iteration of ``sample_counterfactual_motion_map`` (cwm/models/segmentation.py:434-476) with the S samples of
    the sweep sharded over the ranks: every rank predicts its slice of the counterfactual videos, runs the flow network
    and the flow-sample filter on it, sums its flow magnitudes, and the ranks exchange ONE ``[B, H, W]`` all-reduce
    ("only a final NCCL gather of the predicted frames and flow-derived statistics").
    Returns the normalised mean motion map ``[B, H, W]`` on every rank (plus the local videos / flows with
    ``return_local``).  Needs B == 1 like the reference's sweep (segmentation.py:329)."""
    G = generator
    multi = dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if multi else 0
    world = dist.get_world_size() if multi else 1
    y_local = sharded_counterfactual_videos(G, x, active_patches, passive_patches, shifts=shifts, num_samples=num_samples,
                                            sample_batch_size=sample_batch_size, fix_passive=fix_passive, frame=frame,
                                            gather=False, **kwargs)
    from . import sampling
    if len(active_patches.shape) == 2:
        active_patches = active_patches.unsqueeze(-1)
    S = len(G.shifts)  # create_motion_counterfactuals records one shift per sample of the WHOLE sweep on every rank
    lo, hi = shard_bounds(S, rank, world)
    assert y_local.shape[0] == hi - lo, (y_local.shape, lo, hi)
    H, W = y_local.shape[-2:]
    if hi > lo:
        step = sample_batch_size or (hi - lo)
        flows = torch.cat([G.predict_flow(y_local[i:i + step], backward=backward, iters=raft_iters)
                           for i in range(0, hi - lo, step)], 0)
        active_local = active_patches[..., lo:hi] if active_patches.size(-1) > 1 else active_patches.expand(-1, -1, hi - lo)
        view = G.filter_flow_samples(flows, active_local, do_filter=do_filter)
        sums = sampling.flow_magnitude_sum(view, normalize_per_sample=normalize_per_sample)
    else:  # a rank without samples still joins the all-reduce
        flows = y_local.new_zeros((0, 1, 2, H, W))
        sums = torch.zeros(G.x.size(0), H, W, dtype=torch.float32, device=y_local.device)
    if multi:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    mm = sampling.motion_map_finalize(sums, S)
    return (mm, y_local, flows) if return_local else mm
