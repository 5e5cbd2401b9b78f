"""qecnetworks_b200 -- B200-native (sm_100a) implementation of the QEC-Net quaternion
dynamic-routing hot path, behind the reference's own QecModule / symeig interfaces.

The kernels live in ``libqec_b200.so`` (C ABI: include/qec_b200.h), built in-tree by
``python -m qecnetworks_b200.build``.  There is no CPU or eager-PyTorch fallback."""
from . import _lib
from .qec_module import QecModule
from .qec_net import QecNet, QecSiaNet
from .symeig import symeig
from . import functional
from . import eval_ops

__all__ = ["QecModule", "QecNet", "QecSiaNet", "symeig", "functional", "eval_ops", "kernel_launches", "library_path"]


def kernel_launches():
    """kernels of libqec_b200.so launched by this process so far"""
    return _lib.launches()


def library_path():
    return _lib.LIB_PATH
