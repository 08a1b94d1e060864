"""Multi-GPU layout: sensor streams are independent (the reference runs one process per sensor,
launch/demo_touareg.launch:20-31), so they shard one-stream-per-GPU with NO collective on the data path.
torch.distributed is used only to agree on the assignment and to gather per-stream results / timings."""
from __future__ import annotations

import hashlib

import numpy as np


def streams_of_rank(n_streams: int, world_size: int, rank: int) -> list[int]:
    """Stream i runs on rank i mod world_size (one handle, one CUDA stream, one GPU per sensor stream)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return [i for i in range(n_streams) if i % world_size == rank]


def result_digest(events: np.ndarray, cluster_keys) -> str:
    """Order-independent digest of one stream's outputs (finished-column events + finished clusters as point sets)."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(events[["from_gcol", "to_gcol", "ground_points_only"]]).tobytes())
    for key in sorted(cluster_keys):
        h.update(repr(key).encode())
    return h.hexdigest()


def gather_digests(local: dict[int, str]) -> dict[int, str]:
    """All ranks' {stream id: digest} (torch.distributed all_gather_object; identity when not initialised)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    merged: dict[int, str] = {}
    for d in out:
        merged.update(d)
    return merged
