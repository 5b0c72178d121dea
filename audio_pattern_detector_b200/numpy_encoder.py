"""JSON encoder that understands numpy scalars/arrays (reference numpy_encoder.py:1-19)."""
import json

import numpy as np


class NumpyEncoder(json.JSONEncoder):
    def default(self, o):  # noqa: D102
        if isinstance(o, np.integer):
            return int(o)
        if isinstance(o, np.floating):
            return float(o)
        if isinstance(o, np.ndarray):
            return o.tolist()
        return super().default(o)
