"""Framework-wide constants (reference: maggy/constants.py:23-27)."""
import numpy as np


class USER_FCT:
    """What a user ``train_fn`` may return and what ``reporter.broadcast`` accepts."""

    RETURN_TYPES = (float, int, np.number, dict)
    NUMERIC_TYPES = (float, int, np.number)
