"""superpoint-nerf-pytorch_b200: B200-native (sm_100a) implementation of one hot path of
AliYoussef97/SuperPoint-NeRF-Pytorch - the MagicPoint/SuperPoint detector/descriptor forward and the
homography-adaptation pseudo-label export - behind the reference's own Python API.

Import name: ``superpoint_nerf_pytorch_b200`` (see the alias module at the repo root; the directory name
contains hyphens).  Sub-packages mirror the reference's module tree (``superpoint/superpoint/...``):

    models.SuperPoint.SuperPoint                        <- models/SuperPoint.py:5-30
    models.model_utils.sp_utils.box_nms                 <- models/model_utils/sp_utils.py:4-28
    data.data_utils.homographic_augmentation.Homographic_aug  <- data/data_utils/homographic_augmentation.py:14-106
    engine_solvers.export.{ExportDetections, Export_Hpatches_Repeatability, Export_Hpatches_Descriptors}
                                                        <- engine_solvers/export.py:17-222
    utils.get_model.get_model / utils.train_utils.move_to_device
    engine.main                                         <- engine.py:43-208 (export tasks only)

All compute goes through the C-ABI library ``libspn_b200.so`` (include/spn_b200.h, csrc/*.cu).  There is no CPU
fallback: importing works anywhere, but every op raises if the library or a B200 is missing.
"""
from ._native import Context, NativeError, get_context, library_path  # noqa: F401

__all__ = ["Context", "NativeError", "get_context", "library_path"]
