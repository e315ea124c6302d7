"""Miscellaneous helpers (reference volprim/utils.py)."""
from __future__ import annotations

import time
from contextlib import contextmanager

import torch


def concatenate_tensors(images):
    '''Concatenate a list of [H, W(, C)] image tensors on the X axis (reference utils.py:15-32).'''
    if images[0].dim() == 2:
        return torch.cat(images, dim=1)[:, :, None]
    return torch.cat(images, dim=1)


@contextmanager
def time_operation(label):
    '''Wall-clock a block of GPU work (reference utils.py:34-47; dr.sync_thread -> cuda synchronize).'''
    print(f'{label} ...')
    start = time.time()
    yield
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    print(f'{label} → done in {(time.time() - start)} sec')
