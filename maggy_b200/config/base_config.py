"""Single-run experiment config (reference: maggy/config/base_config.py:23-38)."""
from .lagom import LagomConfig


class BaseConfig(LagomConfig):
    def __init__(self, name: str = "base", hb_interval: int = 1, description: str = ""):
        super().__init__(name, description, hb_interval)
