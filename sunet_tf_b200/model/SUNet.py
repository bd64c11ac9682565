"""Config -> model wrapper with the contract of the reference's model/SUNet.py:5-30.

``SUNet_model(opt)`` takes the parsed training.yaml dict, exposes the network as ``.swin_unet`` (so checkpoint keys
are ``swin_unet.*``), always builds a 3-channel-in / 1-channel-out net (SUNet.py:11-12) and repeats a grey input to
three channels (SUNet.py:27-28; done inside the patch-embed kernel here, no materialised repeat).
"""
import torch.nn as nn

from ..modules import SUNet

# training.yaml key -> SUNet keyword (ATTN_DROP_RATE / FINAL_UPSAMPLE are in the YAML but never forwarded by the
# reference, and USE_CHECKPOINTS lands in **kwargs because of the `u1se_checkpoint` typo, SUNet_detail.py:597)
_YAML_TO_KWARG = {
    "IMG_SIZE": "img_size", "PATCH_SIZE": "patch_size", "EMB_DIM": "embed_dim", "DEPTH_EN": "depths", "HEAD_NUM": "num_heads",
    "WIN_SIZE": "window_size", "MLP_RATIO": "mlp_ratio", "QKV_BIAS": "qkv_bias", "QK_SCALE": "qk_scale", "DROP_RATE": "drop_rate",
    "DROP_PATH_RATE": "drop_path_rate", "APE": "ape", "PATCH_NORM": "patch_norm", "USE_CHECKPOINTS": "use_checkpoint",
}


class SUNet_model(nn.Module):
    def __init__(self, config, out_chans=1):
        super().__init__()
        self.config = config
        section = config["SWINUNET"]
        kwargs = {kw: section[key] for key, kw in _YAML_TO_KWARG.items()}
        self.swin_unet = SUNet(in_chans=3, out_chans=out_chans, **kwargs)

    def forward(self, x, out=None):
        # 1- or 3-channel input; the grey->RGB repeat is folded into the first kernel.  `out` (optional) receives the result.
        return self.swin_unet(x, out=out)

    def forward_u8(self, x, out=None):
        """uint8 (B, H, W, C) in -> uint8 (B, H, W, out_chans) out: the demo.py:70-79 edge fused into the forward."""
        return self.swin_unet.forward_u8(x, out=out)

    def forward_eval(self, x, target, weight=None, eps=1e-3):
        """Validation forward (train.py:432-448): logits, sigmoid(logits) and the error sums in one pass."""
        return self.swin_unet.forward_eval(x, target, weight=weight, eps=eps)
