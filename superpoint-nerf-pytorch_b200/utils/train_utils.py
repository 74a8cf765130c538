import torch


def move_to_device(obj, device):
    """Recursively move tensors in dict/list containers (reference utils/train_utils.py:4-18)."""
    if torch.is_tensor(obj):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: move_to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, list):
        return [move_to_device(v, device) for v in obj]
    return obj
