"""Window schedules of the parallel-denoise loop (host-side integer lists).

Mirrors musev/pipelines/context.py:21-66,105-149 (`uniform`, `uniform_v2`, `drop_last_repeat_context`,
`prepare_global_context`) and MMCM/mmcm/utils/itertools_util.py:6-46 (`generate_sample_idxs`).
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import numpy as np


def ordered_halving(val: int) -> float:
    bin_flip = f"{val:064b}"[::-1]
    return int(bin_flip, 2) / (1 << 64)


def uniform(step: int = ..., num_steps: Optional[int] = None, num_frames: int = ..., context_size: Optional[int] = None,
            context_stride: int = 3, context_overlap: int = 4, closed_loop: bool = True):
    if num_frames <= context_size:
        yield list(range(num_frames))
        return
    context_stride = min(context_stride, int(np.ceil(np.log2(num_frames / context_size))) + 1)
    for context_step in 1 << np.arange(context_stride):
        pad = int(round(num_frames * ordered_halving(step)))
        for j in range(int(ordered_halving(step) * context_step) + pad,
                       num_frames + pad + (0 if closed_loop else -context_overlap),
                       (context_size * context_step - context_overlap)):
            yield [int(e % num_frames) for e in range(j, j + context_size * context_step, context_step)]


def generate_sample_idxs(total: int, window_size: int, step: int, sample_rate: int = 1, drop_last: bool = False):
    idxs = [idx for i, idx in enumerate(range(total)) if i % sample_rate == 0]
    sample_idxs, window_start = [], 0
    while window_start < len(idxs):
        window_end = window_start + window_size
        if window_end > len(idxs) and drop_last:
            break
        sample_idxs.append(idxs[window_start:window_end])
        window_start += step
    return sample_idxs


def uniform_v2(step: int = ..., num_steps: Optional[int] = None, num_frames: int = ..., context_size: Optional[int] = None,
               context_stride: int = 3, context_overlap: int = 4, closed_loop: bool = True):
    return generate_sample_idxs(total=num_frames, window_size=context_size, step=context_size - context_overlap,
                                sample_rate=1, drop_last=False)


def get_context_scheduler(name: str) -> Callable:
    if name == "uniform":
        return uniform
    elif name == "uniform_v2":
        return uniform_v2
    raise ValueError(f"Unknown context_overlap policy {name}")


def drop_last_repeat_context(contexts: List[List[int]]) -> List[List[int]]:
    if len(contexts) >= 2 and contexts[-1][-1] == contexts[-2][-1]:
        return contexts[:-1]
    return contexts


def prepare_global_context(context_schedule: str, num_inference_steps: int, time_size: int, context_frames: int,
                           context_stride: int, context_overlap: int, context_batch_size: int):
    context_queue = list(get_context_scheduler(context_schedule)(
        step=0, num_steps=num_inference_steps, num_frames=time_size, context_size=context_frames,
        context_stride=context_stride, context_overlap=context_overlap))
    context_queue = drop_last_repeat_context(context_queue)
    n = math.ceil(len(context_queue) / context_batch_size)
    return [context_queue[i * context_batch_size:(i + 1) * context_batch_size] for i in range(n)]


def assign_windows(window_lengths: List[int], world_size: int) -> List[List[int]]:
    """Contiguous window ranges per rank, balanced by window length (+1 vision-condition frame each): the number
    of frames a window pushes through the UNet is what its step costs. Returns window indices per rank."""
    n = len(window_lengths)
    cost = [l + 1 for l in window_lengths]
    total = sum(cost)
    out: List[List[int]] = [[] for _ in range(world_size)]
    acc, r = 0.0, 0
    for i in range(n):
        # move to the next rank when this window's midpoint passes the rank's share boundary
        while r < world_size - 1 and acc + cost[i] / 2 > total * (r + 1) / world_size:
            r += 1
        out[r].append(i)
        acc += cost[i]
    return out
