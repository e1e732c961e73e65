"""enzo-e_b200: B200-native VL+CT hydro/MHD block update (EnzoMethodMHDVlct).

Submodules
  abi      ctypes mirror of include/vlct.h (no GPU, no library load)
  lib      loader for csrc/libvlct_b200.so -- fails loudly if it is missing
  method   host-side mirror of the reference's Method plugin interface
  domain   unigrid driver: periodic refresh / NCCL ghost exchange between GPUs
"""
from . import abi  # noqa: F401

__all__ = ["abi", "lib", "method"]
