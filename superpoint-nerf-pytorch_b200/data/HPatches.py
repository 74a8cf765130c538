"""HPatches pair dataset with the reference's class interface (data/HPatches.py:12-150): cv2 decode on the host (as the
reference), resize + centre crop + /255 in one kernel per image, ground-truth homography adapted to the resize on the
host in the reference's fp32 operation order."""
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import Dataset

from .. import settings
from .preprocessing import pin_host as _pin, adapt_homography_to_resize, ratio_preserving_resize


class HPatches(Dataset):
    def __init__(self, data_config, device="cuda") -> None:
        super().__init__()
        self.config = data_config
        self.device = device
        if torch.cuda.is_available() and torch.device(device).type == "cuda" and torch.device(device).index is None:
            self.device = f"cuda:{torch.cuda.current_device()}"   # resolved here: decode threads start with device 0 current
        self.samples = self._init_dataset()

    def _init_dataset(self):
        """HPatches.py:21-52."""
        data_dir = Path(settings.DATA_PATH, self.config["name"])
        image_paths, warped_image_paths, homographies, names = [], [], [], []
        for folder_dir in sorted(x for x in data_dir.iterdir() if x.is_dir()):
            if self.config["alteration"] == "i" != folder_dir.stem[0] != "i":
                continue
            if self.config["alteration"] == "v" != folder_dir.stem[0] != "v":
                continue
            num_images = 1 if self.config["name"] == "COCO" else 5
            file_ext = ".ppm" if self.config["name"] == "HPatches" else ".jpg"
            for i in range(2, 2 + num_images):
                image_paths.append(str(Path(folder_dir, "1" + file_ext)))
                warped_image_paths.append(str(Path(folder_dir, str(i) + file_ext)))
                homographies.append(np.loadtxt(str(Path(folder_dir, "H_1_" + str(i)))))
                names.append(f"{folder_dir.stem}_{1}_{i}")
        return {"image_paths": image_paths, "warped_image_paths": warped_image_paths, "homography": homographies, "names": names}

    def __len__(self):
        return len(self.samples["image_paths"])

    def read_image(self, image):
        """HPatches.py:58-60, stopping at the decoded uint8 image."""
        import cv2
        return _pin(torch.from_numpy(cv2.imread(image, cv2.IMREAD_GRAYSCALE)), self.device)

    def ratio_preserving_resize(self, image, normalize=False):
        # host images are uploaded by the module function on its side stream (never on the consumer's stream)
        return ratio_preserving_resize(image, self.config["preprocessing"]["resize"], normalize=normalize, device=self.device)

    def adapt_homography_to_resize(self, homographies):
        """HPatches.py:74-100."""
        return adapt_homography_to_resize(homographies["homography"], homographies["image_shape"], homographies["warped_image_shape"],
                                          self.config["preprocessing"]["resize"])

    def decode(self, index):
        return (self.read_image(self.samples["image_paths"][index]), self.read_image(self.samples["warped_image_paths"][index]),
                torch.as_tensor(self.samples["homography"][index], dtype=torch.float32), self.samples["names"][index])

    def finish(self, decoded):
        image, warped_image, homography, name = decoded
        if self.config["preprocessing"]["resize"]:
            homography = self.adapt_homography_to_resize({"homography": homography, "image_shape": torch._shape_as_tensor(image),
                                                          "warped_image_shape": torch._shape_as_tensor(warped_image)})
        return {"image": self.ratio_preserving_resize(image, normalize=True),
                "warped_image": self.ratio_preserving_resize(warped_image, normalize=True),
                "homography": homography.to(self.device), "name": name}

    def __getitem__(self, index):
        return self.finish(self.decode(index))

    def batch_collator(self, batch):
        """HPatches.py:137-150."""
        return {"image": torch.stack([item["image"].unsqueeze(0) for item in batch]),
                "warped_image": torch.stack([item["warped_image"].unsqueeze(0) for item in batch]),
                "homography": torch.stack([item["homography"] for item in batch]),
                "name": [item["name"] for item in batch]}
