"""Minimal TensorFlow stand-in so that the UNMODIFIED reference package can be imported in an image without TF."""
import sys
import types

import tensorboard.compat.tensorflow_stub as tf
from tensorboard.compat.proto import tensor_pb2

tf.__version__ = "2.4.1"
tf.make_tensor_proto = lambda v, *a, **k: tensor_pb2.TensorProto()


class _Callback:
    def __init__(self, *a, **k):
        pass


keras = types.ModuleType("tensorflow.keras")
keras.callbacks = types.ModuleType("tensorflow.keras.callbacks")
keras.callbacks.Callback = _Callback
keras.Model = object
tf.keras = keras


class _Config:
    @staticmethod
    def list_physical_devices(kind=None):
        import torch
        n = torch.cuda.device_count() if torch.cuda.is_available() else 0
        return [f"GPU:{i}" for i in range(max(n, 1))]


tf.config = _Config()
sys.modules["tensorflow"] = tf
sys.modules["tensorflow.keras"] = keras
sys.modules["tensorflow.keras.callbacks"] = keras.callbacks
