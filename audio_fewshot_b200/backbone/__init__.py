from .conv_four import Conv64F
from .resnet_12 import resnet12
from .resnet_bdc import BdcPool, resnet12Bdc

__all__ = ["Conv64F", "resnet12", "resnet12Bdc", "BdcPool"]
