from .fs2_step import FastSpeech2TrainStep  # noqa: F401
from .flat import FlatBuffers  # noqa: F401
from .pwg_step import PWGTrainStep  # noqa: F401
