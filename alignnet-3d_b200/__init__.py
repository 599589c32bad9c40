"""alignnet-3d_b200: B200-native engine for the AlignNet-3D tp8 hot path.

Host code is Python; all device work goes through the C-ABI library
``csrc/libalignnet_b200.so`` (hand-written sm_100a CUDA) loaded with ctypes.  There is no CPU
fallback: importing the engine without the built library, or running it on a non-sm_100
device, raises.
"""
__version__ = "0.1.0"
