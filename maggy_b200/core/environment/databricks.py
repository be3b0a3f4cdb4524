"""Databricks flavour: artefacts under /dbfs/maggy_log (reference: maggy/core/environment/databricks.py:23-78)."""
import os

from .base import BaseEnv


class DatabricksEnv(BaseEnv):
    def __init__(self):
        self.log_dir = os.environ.get("MAGGY_LOG_DIR") or "/dbfs/maggy_log/"
        os.makedirs(self.log_dir, exist_ok=True)

    def project_path(self):
        return "/dbfs/"
