"""Loader factory with the reference's signature (utils/data_loaders.py:4-101), export tasks only.

``get_loader(config, task, device, export_split=...)`` imports ``data/<config.data.name>.py`` and instantiates
``config.data.class_name`` like the reference; instead of ``torch.utils.data.DataLoader(num_workers=0)`` (one image
decoded at a time on the loop's thread) it returns a ``PrefetchLoader``: the host part of ``__getitem__`` (file read +
JPEG/PPM decode, which release the GIL) runs on a thread pool a bounded number of items ahead, the device part (one
resize/crop/normalise kernel per image on the uint8 pixels) runs on the consumer's thread and stream.  It yields the
same collated batch dictionaries in the same order; rank sharding (``rank``/``world``) takes items with
index = rank (mod world) BEFORE anything is decoded.
"""
import concurrent.futures as cf
import importlib
import os

_DATA_PACKAGE = __name__.rsplit(".utils.", 1)[0] + ".data"


class PrefetchLoader:
    def __init__(self, dataset, batch_size=1, workers=None, depth=None, rank=0, world=1):
        self.dataset = dataset
        self.batch_size = max(1, int(batch_size))
        self.workers = workers or max(1, min(32, (os.cpu_count() or 2) - 1))
        self.depth = depth or 4 * self.workers
        self.indices = list(range(rank, len(dataset), world))

    def __len__(self):
        return (len(self.indices) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        ds = self.dataset
        split = hasattr(ds, "decode") and hasattr(ds, "finish")
        host_part = ds.decode if split else ds.__getitem__
        with cf.ThreadPoolExecutor(max_workers=self.workers) as pool:
            pending, nxt, batch = [], 0, []
            while nxt < len(self.indices) or pending:
                while nxt < len(self.indices) and len(pending) < self.depth:
                    pending.append(pool.submit(host_part, self.indices[nxt]))
                    nxt += 1
                item = pending.pop(0).result()              # in order; re-raises a worker's exception here
                batch.append(ds.finish(item) if split else item)
                if len(batch) == self.batch_size:
                    yield ds.batch_collator(batch)
                    batch = []
            if batch:
                yield ds.batch_collator(batch)


def get_loader(config, task, device="cuda", validate_training=False, export_split=None, nerf_train=False, rank=0, world=1):
    name, class_name = config["data"]["name"], config["data"]["class_name"]
    batch_size = config["data"]["batch_size"]
    cls = getattr(importlib.import_module(f"{_DATA_PACKAGE}.{name}"), class_name)
    if task in ("train", "test", "export_NeRF_labels") or validate_training or nerf_train:
        raise NotImplementedError(f"task {task!r} is outside the B200 hot path (SURVEY.md section 8): use the reference's loader")
    if task == "export_pseudo_labels":
        dataset = cls(config["data"], task=export_split, device=device)
    elif task in ("export_HPatches_Repeatability", "export_HPatches_Descriptors"):
        dataset = cls(config["data"], device=device)
    else:
        raise ValueError(f"unknown task {task!r}")
    return PrefetchLoader(dataset, batch_size, workers=config["data"].get("loader_workers"), rank=rank, world=world)
