"""sunet_tf_b200 - B200-native forward path of SUNet (Swin-Transformer UNet denoiser).

Public surface = the reference's module names (see modules.py) + ``SUNet_model`` + the any-resolution tile driver
(tiles.py) + batch/tile sharding helpers (shard.py).  All device work happens in libsunet_b200.so (csrc/).
"""
from .model.SUNet import SUNet_model  # noqa: F401
from .modules import (BasicLayer, BasicLayer_up, Mlp, PatchEmbed, PatchMerging, SUNet, SwinTransformerBlock, UpSample,  # noqa: F401
                      WindowAttention, window_partition, window_reverse)

__all__ = ["SUNet_model", "SUNet", "SwinTransformerBlock", "WindowAttention", "Mlp", "PatchEmbed", "PatchMerging", "UpSample",
           "BasicLayer", "BasicLayer_up", "window_partition", "window_reverse"]
