"""B200-native dense stereo-to-semantic-map path (drop-in for stereo.h / mapper.h of
MuMuJun97/semantic_slam_mapping).  All arithmetic lives in libssm.so (hand-written sm_100a CUDA behind the
C ABI of include/ssm.h); this package is the thin host side: ctypes binding, the reference-shaped
`calDisparity_SGBM` / `Mapper` mirrors, parameter structs and synthetic inputs."""
from .params import Params, cityscapes_params, SEGNET12_BGR, CITYSCAPES19_BGR  # noqa: F401
from .lib import Context, SsmError, load, voxel_owner, LIB_PATH  # noqa: F401
