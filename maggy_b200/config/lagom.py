"""Base of all experiment configs (reference: maggy/config/lagom.py:22-34)."""
from abc import ABC


class LagomConfig(ABC):
    """``name``/``description`` label the experiment; ``hb_interval`` is kept for API parity — the ring
    control plane reports immediately, the value only bounds liveness (a worker is declared lost after
    ``max(10 * hb_interval, 30)`` seconds of silence)."""

    def __init__(self, name: str, description: str, hb_interval: int):
        self.name = name
        self.description = description
        self.hb_interval = hb_interval
