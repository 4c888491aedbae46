"""Make the reference's own import statements resolve to the B200 drop-ins.

`pin_slam.py` (and utils/tracker.py, utils/mapper.py, utils/mesher.py ...) of the
reference import `from model.neural_points import NeuralPoints` and
`from model.decoder import Decoder`.  Calling `install()` before those imports
registers this package's modules under those names, so the reference driver runs
unchanged on top of the CUDA hot path."""
import importlib
import sys


def install(tracker_and_mapper: bool = True) -> None:
    np_mod = importlib.import_module("pin_slam_b200.model.neural_points")
    dec_mod = importlib.import_module("pin_slam_b200.model.decoder")
    sys.modules["model.neural_points"] = np_mod
    sys.modules["model.decoder"] = dec_mod
    if "model" in sys.modules:
        sys.modules["model"].neural_points = np_mod
        sys.modules["model"].decoder = dec_mod
