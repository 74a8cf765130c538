"""Host-side helpers shared by the task classes."""
from collections.abc import Mapping

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
            pairs = [(key, walk(value)) for key, value in node.items()]
            try:
                return type(node)(pairs)
            except TypeError:                       # e.g. defaultdict(factory): constructor does not take pairs
                return dict(pairs)
        if isinstance(node, (str, bytes)):
            return node
        if isinstance(node, tuple) and hasattr(node, "_fields"):     # namedtuple: positional constructor
            return type(node)(*(walk(value) for value in node))
        if isinstance(node, (list, tuple)):
            items = [walk(value) for value in node]
            return items if isinstance(node, list) else type(node)(items)
        return node                                 # range, arrays, scalars, anything else: unchanged

    return walk(batch)
