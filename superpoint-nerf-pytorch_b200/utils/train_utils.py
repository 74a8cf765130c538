"""Host-side helpers shared by the task classes."""
from collections.abc import Mapping, Sequence

import torch


def move_to_device(batch, device, non_blocking=False):
    """Return ``batch`` with every tensor placed on ``device``; containers are rebuilt, other leaves pass through.

    Drop-in for the reference helper of the same name (utils/train_utils.py:4-18), which the export loops call on each
    loader batch; here mappings and sequences of any depth are handled by one generic walk, tuples keep their type, and
    ``non_blocking=True`` lets pinned host tensors be copied asynchronously."""
    def walk(node):
        if isinstance(node, torch.Tensor):
            return node.to(device, non_blocking=non_blocking)
        if isinstance(node, Mapping):
            return type(node)((key, walk(value)) for key, value in node.items())
        if isinstance(node, (str, bytes)):
            return node
        if isinstance(node, Sequence):
            items = [walk(value) for value in node]
            return items if isinstance(node, list) else type(node)(items)
        return node

    return walk(batch)
