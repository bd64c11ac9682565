"""The reference's training.yaml SWINUNET section (training.yaml:6-22) as a dict, for callers without the YAML file."""
DEFAULT_OPT = {
    "SWINUNET": {
        "IMG_SIZE": 256, "PATCH_SIZE": 4, "WIN_SIZE": 8, "EMB_DIM": 96, "DEPTH_EN": [8, 8, 8, 8], "HEAD_NUM": [8, 8, 8, 8],
        "MLP_RATIO": 4.0, "QKV_BIAS": True, "QK_SCALE": 8, "DROP_RATE": 0.0, "ATTN_DROP_RATE": 0.0, "DROP_PATH_RATE": 0.1,
        "APE": False, "PATCH_NORM": True, "USE_CHECKPOINTS": False, "FINAL_UPSAMPLE": "Dual up-sample",
    }
}
