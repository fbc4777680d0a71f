"""Import shim so that the reference's entry points keep working unchanged:

    python -m model.optimize ...        (scripts/train/optimize_texture_*.sh)
    from model.model import TextureOptimizationStyleTransferPipeline
    from model.losses.content_and_style_losses import ContentAndStyleLoss, VGG

Everything here re-exports stylemesh_b200.model.* (the B200 implementation); nothing is implemented in this package.
"""
