"""A stand-in for the three TensorFlow entry points the TF executor touches (this image ships no TensorFlow):
``distribute.MultiWorkerMirroredStrategy().scope()``, ``keras.mixed_precision.set_global_policy`` and
``keras.Model``. It records what happened so the tests can assert on it."""
import contextlib
import json
import os
import types

__version__ = "2.99.fake"
EVENTS = []


class _Strategy:
    def __init__(self):
        self.tf_config = json.loads(os.environ["TF_CONFIG"])
        EVENTS.append(("strategy", self.tf_config["task"]["type"]))
        self.in_scope = 0

    @contextlib.contextmanager
    def scope(self):
        self.in_scope += 1
        try:
            yield self
        finally:
            self.in_scope -= 1

    @property
    def num_replicas_in_sync(self):
        c = self.tf_config["cluster"]
        return len(c.get("chief", [])) + len(c.get("worker", []))


class _Model:
    def __init__(self, *a, **k):
        self.built_in_scope = None


distribute = types.SimpleNamespace(MultiWorkerMirroredStrategy=_Strategy)
keras = types.SimpleNamespace(
    Model=_Model,
    mixed_precision=types.SimpleNamespace(set_global_policy=lambda name: EVENTS.append(("policy", name))))
