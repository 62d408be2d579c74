from .fs2_step import FastSpeech2TrainStep  # noqa: F401
