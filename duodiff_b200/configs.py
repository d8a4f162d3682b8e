"""model_params of the reference's configs/*.yaml, for synthetic benchmarks and tests."""

CONFIGS = {
    # configs/*.yaml model_params of the reference
    "cifar10_3": dict(img_size=32, patch_size=2, in_chans=3, embed_dim=512, depth=3, num_heads=8, mlp_ratio=4,
                      qkv_bias=False, mlp_time_embed=False, num_classes=-1, normalize_timesteps=True),
    "cifar10": dict(img_size=32, patch_size=2, in_chans=3, embed_dim=512, depth=13, num_heads=8, mlp_ratio=4,
                    qkv_bias=False, mlp_time_embed=False, num_classes=-1, normalize_timesteps=True),
    "celeba_3": dict(img_size=64, patch_size=4, in_chans=3, embed_dim=512, depth=3, num_heads=8, mlp_ratio=4,
                     qkv_bias=False, mlp_time_embed=False, num_classes=-1, normalize_timesteps=True),
    "celeba": dict(img_size=64, patch_size=4, in_chans=3, embed_dim=512, depth=13, num_heads=8, mlp_ratio=4,
                   qkv_bias=False, mlp_time_embed=False, num_classes=-1, normalize_timesteps=True),
    "imagenet64_3": dict(img_size=64, patch_size=4, in_chans=3, embed_dim=768, depth=3, num_heads=12, mlp_ratio=4,
                         qkv_bias=False, mlp_time_embed=False, num_classes=1000, normalize_timesteps=False),
    "imagenet64": dict(img_size=64, patch_size=4, in_chans=3, embed_dim=768, depth=17, num_heads=12, mlp_ratio=4,
                       qkv_bias=False, mlp_time_embed=False, num_classes=1000, normalize_timesteps=False),
    "imagenet256_3": dict(img_size=32, patch_size=2, in_chans=4, embed_dim=1024, depth=3, num_heads=16, mlp_ratio=4,
                          qkv_bias=False, mlp_time_embed=False, num_classes=1001, normalize_timesteps=False),
    "imagenet256": dict(img_size=32, patch_size=2, in_chans=4, embed_dim=1024, depth=21, num_heads=16, mlp_ratio=4,
                        qkv_bias=False, mlp_time_embed=False, num_classes=1001, normalize_timesteps=False),
}
