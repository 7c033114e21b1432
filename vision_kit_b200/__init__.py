"""vision_kit_b200 -- B200-native YOLO detection data path for Vision-Kit.

Letterbox -> Detect decode -> confidence filter -> class-aware NMS as
hand-written sm_100a CUDA kernels behind a C-ABI shared library
(``csrc/`` -> ``libvk_b200.so``, declared in ``include/vk_b200.h``), with host
shims that keep the reference's Python call surface.  There is no CPU
fallback: every entry point raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
