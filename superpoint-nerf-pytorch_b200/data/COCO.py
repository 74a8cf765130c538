"""COCO image dataset for the export tasks, with the reference's class interface (data/COCO.py:14-139).

Export configuration only (``has_labels: False``, no augmentation, no warped pair - what
``configs/magicpoint_coco_export.yaml`` selects); the training-time branches of the reference class (labels,
photometric / homographic augmentation, warped pairs) are outside this package and rejected.  What differs underneath:
``read_image`` decodes to uint8 on a host thread (torchvision's decoder, as the reference) and everything after the
decode - bilinear ratio-preserving resize, centre crop, /255 - is ONE kernel (spn_resize_crop) on the uint8 image.
``decode`` / ``finish`` split ``__getitem__`` so that a prefetching loader can run the decodes on worker threads.
"""
from pathlib import Path

import torch
from torch.utils.data import Dataset

from .. import settings
from .preprocessing import pin_host as _pin, ratio_preserving_resize


class COCO(Dataset):
    def __init__(self, data_config, task="training", device="cuda") -> None:
        super().__init__()
        self.config = data_config
        self.device = device
        if torch.cuda.is_available() and torch.device(device).type == "cuda" and torch.device(device).index is None:
            self.device = f"cuda:{torch.cuda.current_device()}"   # resolved here: decode threads start with device 0 current
        self.action = "training" if task == "training" else "validation" if task == "validation" else "test"
        aug = data_config.get("augmentation", {})
        if data_config.get("has_labels") or data_config.get("warped_pair") or any(aug.get(k, {}).get("enable") for k in aug):
            raise NotImplementedError("the B200 COCO loader serves the export tasks (has_labels / warped_pair / augmentation "
                                      "are training-time features of the reference)")
        self.samples = self._init_dataset()

    def _init_dataset(self):
        """List of image paths and names (COCO.py:34-56)."""
        data_dir = Path(settings.DATA_PATH, self.config["name"], "images", self.action)
        image_paths = list(data_dir.iterdir())
        if self.config.get("truncate"):
            image_paths = image_paths[:int(self.config["truncate"] * len(image_paths))]
        return {"image_paths": [str(p) for p in image_paths], "names": [p.stem for p in image_paths]}

    def __len__(self):
        return len(self.samples["image_paths"])

    def read_image(self, image):
        """COCO.py:61-64, stopping at the decoded uint8 (H0,W0) image (the float conversion is fused into the resize)."""
        import torchvision
        data = torchvision.io.read_file(image)
        return _pin(torchvision.io.decode_image(data, torchvision.io.ImageReadMode.GRAY).squeeze(0), self.device)

    def ratio_preserving_resize(self, image, normalize=False):
        """COCO.py:66-76 on the device: (H0,W0) uint8 / fp32 -> (H,W) fp32."""
        # host images are uploaded by the module function on its side stream (never on the consumer's stream)
        return ratio_preserving_resize(image, self.config["preprocessing"]["resize"], normalize=normalize, device=self.device)

    def decode(self, index):
        """Host part of ``__getitem__`` (thread-safe): -> (uint8 image, name)."""
        return self.read_image(self.samples["image_paths"][index]), self.samples["names"][index]

    def finish(self, decoded):
        """Device part: resize + crop + /255 (COCO.py:89-90,135) -> the reference's item dictionary."""
        image, name = decoded
        return {"raw": {"image": self.ratio_preserving_resize(image, normalize=True)}, "name": name}

    def __getitem__(self, index):
        return self.finish(self.decode(index))

    def batch_collator(self, batch):
        """COCO.py:140-147 for label-free items."""
        assert len(batch) > 0 and isinstance(batch[0], dict)
        return {"raw": {"image": torch.stack([item["raw"]["image"].unsqueeze(0) for item in batch])},
                "name": [item["name"] for item in batch]}
