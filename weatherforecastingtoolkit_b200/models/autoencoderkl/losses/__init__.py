from .model import NLayerDiscriminator, hinge_d_loss, weights_init  # noqa: F401
