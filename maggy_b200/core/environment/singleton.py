"""Environment auto-detection (reference: maggy/core/environment/singleton.py:20-62)."""
import os


class EnvSing:
    obj = None

    def __new__(cls, *args, **kwargs):
        raise NotImplementedError("Use EnvSing.get_instance()")

    @classmethod
    def get_instance(cls):
        if cls.obj is None:
            if "REST_ENDPOINT" in os.environ:
                from .hopsworks import HopsworksEnv

                cls.obj = HopsworksEnv()
            elif "DATABRICKS_ROOT_CONDA_ENV" in os.environ:
                from .databricks import DatabricksEnv

                cls.obj = DatabricksEnv()
            else:
                from .base import BaseEnv

                cls.obj = BaseEnv()
        return cls.obj

    @classmethod
    def set_instance(cls, obj):
        cls.obj = obj

    @classmethod
    def reset(cls):
        cls.obj = None
