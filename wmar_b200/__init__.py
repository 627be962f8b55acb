"""wmar_b200 -- B200-native hot path of facebookresearch/wmar (watermarked autoregressive image generation).

Host side mirrors the reference's operator surface:
  wmar_b200.watermarking.gentime_watermark.GentimeWatermark   <-> wmar/watermarking/gentime_watermark.py
  wmar_b200.models.{Taming,Rar}ARMMWrapper                    <-> wmar/models/*_wrapper.py
Everything below it is CUDA behind the C-ABI in include/wmar_b200.h (wmar_b200/libwmar_b200.so).
"""
__version__ = "0.1.0"
