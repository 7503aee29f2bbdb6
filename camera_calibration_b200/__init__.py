"""B200-native joint-optimisation (bundle-adjustment) path of puzzlepaint/camera_calibration.

Only what the hot path needs lives here:
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/b200ba.h)
  cabi.py      ctypes mirror of the C ABI (fails loudly if the library is not built)
  api.py       host-side mirror of the reference interface: CameraModel / Dataset / BAState /
               OptimizeJointly (same names, argument meaning and error behaviour)
  synthetic.py seeded synthetic star-pattern problems (BASELINE.json configs 1-5)
  distributed.py  imageset sharding + NCCL communicator set-up over torch.distributed
"""
__version__ = "0.1.0"
