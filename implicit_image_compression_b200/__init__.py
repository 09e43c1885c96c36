"""siren-b200: B200-native (sm_100a) implementation of the per-image SIREN fit hot path of
varun19299/implicit-image-compression.  The module layout mirrors the reference package `implicit_image`
for the path it replaces (models / utils.train_helper / pipeline.masking / pipeline.quant / data), so a
caller switches by changing the package name.  All compute goes through libsirenb200.so (C ABI,
include/siren_b200.h); there is no CPU or PyTorch fallback."""
__version__ = "0.1.0"
