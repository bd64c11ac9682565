"""Names of the reference's model/SUNet_detail.py, served by the B200 implementation (sunet_tf_b200/modules.py)."""
from ..modules import (BasicLayer, BasicLayer_up, Mlp, PatchEmbed, PatchMerging, SUNet, SwinTransformerBlock, UpSample,  # noqa: F401
                       WindowAttention, window_partition, window_reverse)
