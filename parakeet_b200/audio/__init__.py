"""parakeet.audio of the reference (audio/audio.py, audio/spec_normalizer.py): AudioProcessor and the magnitude normalisers."""
from .audio import AudioProcessor
from .spec_normalizer import LogMagnitude, NormalizerBase, UnitMagnitude

__all__ = ["AudioProcessor", "LogMagnitude", "NormalizerBase", "UnitMagnitude"]
