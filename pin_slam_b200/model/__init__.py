from .decoder import Decoder  # noqa: F401
from .neural_points import NeuralPoints  # noqa: F401
