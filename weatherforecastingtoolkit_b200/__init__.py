"""B200-native (sm_100a) drop-in for the Path-B latent nowcast rollout and its scoring of
Autobot37/weatherforecastingtoolkit. See DESIGN.md."""

__version__ = "0.1.0"
