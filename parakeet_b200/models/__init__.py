from .fastspeech2 import FastSpeech2, FastSpeech2Inference, FastSpeech2Loss  # noqa: F401
from .parallel_wavegan import PWGDiscriminator, PWGGenerator, PWGInference  # noqa: F401
from .waveflow import ConditionalWaveFlow  # noqa: F401
