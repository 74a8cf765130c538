"""On-GPU counterparts of the reference's export evaluations (evaluations/detector_evaluation.py,
evaluations/descriptor_evaluation.py): same function names, arguments and return values; the keypoint selection,
distance counting and descriptor matching run as sm_100a kernels behind include/spn_b200.h (csrc/eval.cu)."""
