"""piano-a2s hot path, B200-native: VQT front end, ConvStack, BiGRU encoder, hierarchical attention decoder."""
from . import _lib  # noqa: F401  (binding is resolved lazily; importing the package never needs a GPU)

__all__ = ["models", "ops", "vqt", "train", "rng", "kern", "audio", "batching", "metrics"]
