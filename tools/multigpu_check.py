"""SURVEY 8e on real GPUs: frames of the config-2 sequence sharded over the ranks of a torchrun job (one rank per GPU),
detections gathered to rank 0 over NCCL, compared with the list rank 0 computes alone -- they must be identical.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multigpu_check.py [n_frames]"""
import os
import sys

import cv2
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cylindertag_b200 import Detector, synth  # noqa: E402
from cylindertag_b200.sharding import frame_shard, gather_detections  # noqa: E402


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    data = os.path.join(ROOT, "tests", "golden", "data")
    gray = cv2.imread(os.path.join(data, "test_gray.png"), cv2.IMREAD_UNCHANGED)
    det = Detector(marker_path=os.path.join(data, "CTag_2f12c.marker"), device=local)
    s, e = frame_shard(n_frames, rank, world)
    mine = synth.video_sequence(gray, 120, 2024, first=s, count=e - s)
    markers, counts, _ = det.detect_batch(mine, 5, True, 5, cap_per_frame=8)
    markers["frame"] = 0  # batch-local index: not part of the result
    allm, allc = gather_detections(markers, counts, n_frames, dist if world > 1 else None, device=f"cuda:{local}")
    if rank == 0:
        whole = synth.video_sequence(gray, 120, 2024, first=0, count=n_frames)
        m1, c1, _ = det.detect_batch(whole, 5, True, 5, cap_per_frame=8)
        m1["frame"] = 0
        same = bool(np.array_equal(allc, c1) and allm.tobytes() == m1.tobytes())
        print(f"multi-GPU check: {world} rank(s), {n_frames} frames, {int(c1.sum())} markers; gathered list equals the "
              f"single-GPU list: {same}")
    det.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
