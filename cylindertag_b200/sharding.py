"""Frame sharding across GPUs: frames are independent units, so a batch or video is cut into contiguous
blocks, one block per rank, with no data-path collective.  The only exchange is the optional final gather of the
fixed-size detection records to rank 0 (a few KB per frame), done with torch.distributed (NCCL over NVLink on the GPU
box, gloo in the CPU tests)."""
import numpy as np


def frame_shard(n_frames: int, rank: int, world: int):
    """Contiguous block of frames for `rank`: frame f belongs to rank floor(f * world / n_frames) (video locality)."""
    start = (n_frames * rank) // world
    end = (n_frames * (rank + 1)) // world
    return start, end


def owner_of(frame: int, n_frames: int, world: int) -> int:
    # inverse of frame_shard
    r = (frame * world) // n_frames
    while frame_shard(n_frames, r, world)[1] <= frame:
        r += 1
    while frame_shard(n_frames, r, world)[0] > frame:
        r -= 1
    return r


def gather_detections(markers: np.ndarray, counts: np.ndarray, n_frames: int, dist=None, device="cpu"):
    """Gathers per-rank results (markers [n_local, cap], counts [n_local]) to rank 0 in frame order.
    Returns (markers [n_frames, cap], counts [n_frames]) on rank 0 and (None, None) elsewhere; the `frame` field of the
    gathered records is rewritten to the global frame index.  With dist=None (single process) it is the identity."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return markers, counts
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    cap = markers.shape[1]
    rec = markers.dtype.itemsize
    sizes = [frame_shard(n_frames, r, world) for r in range(world)]
    max_local = max(e - s for s, e in sizes)
    # fixed-size padded byte buffers so that one gather moves everything
    buf = np.zeros((max_local, cap * rec + 4), np.uint8)
    n_local = markers.shape[0]
    if n_local:
        buf[:n_local, :cap * rec] = np.ascontiguousarray(markers).view(np.uint8).reshape(n_local, cap * rec)
        buf[:n_local, cap * rec:] = np.ascontiguousarray(counts.astype(np.int32)).view(np.uint8).reshape(n_local, 4)
    t = torch.from_numpy(buf).to(device)
    out = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, out, dst=0)
    if rank != 0:
        return None, None
    all_m = np.zeros((n_frames, cap), markers.dtype)
    all_c = np.zeros(n_frames, np.int32)
    for r, (s, e) in enumerate(sizes):
        b = out[r].cpu().numpy()
        k = e - s
        if k:
            all_m[s:e] = b[:k, :cap * rec].copy().view(markers.dtype).reshape(k, cap)
            all_c[s:e] = b[:k, cap * rec:].copy().view(np.int32).reshape(k)
    # ctag_marker::frame is the index inside the batch a rank ran; in the gathered list it is the global frame index
    if "frame" in (markers.dtype.names or ()):
        for f in range(n_frames):
            all_m["frame"][f, :min(int(all_c[f]), cap)] = f
    return all_m, all_c
