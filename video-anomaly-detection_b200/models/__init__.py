"""B200-native drop-in for the reference's `models` package (same export list as models/__init__.py:5-11)."""

from .autoencoder import ConvAutoencoder, Encoder, Decoder

__all__ = [
    "ConvAutoencoder",
    "Encoder",
    "Decoder",
]
