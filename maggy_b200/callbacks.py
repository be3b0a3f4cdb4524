"""Training-loop callbacks that feed ``reporter.broadcast`` (reference: maggy/callbacks.py:20-66).

The reference ships Keras callbacks; the PyTorch-only rebuild keeps the same two classes (they are
duck-typed Keras callbacks: ``on_batch_end(batch, logs)`` / ``on_epoch_end(epoch, logs)``) so Keras
users can still pass them to ``model.fit`` and PyTorch users can call them from their own loop.
"""
try:  # inherit from keras' Callback when TensorFlow is around (never the case in this image)
    from tensorflow.keras.callbacks import Callback as _Base  # type: ignore
except Exception:  # pragma: no cover
    class _Base:  # minimal stand-in with the Keras hook names
        def set_model(self, model):
            self.model = model

        def set_params(self, params):
            self.params = params


class KerasBatchEnd(_Base):
    """Broadcast ``logs[metric]`` after every batch."""

    def __init__(self, reporter, metric="loss"):
        super().__init__()
        self.metric_name = metric
        self.reporter = reporter

    def on_batch_end(self, batch, logs=None):
        logs = logs or {}
        self.reporter.broadcast(logs.get(self.metric_name, 0))

    on_train_batch_end = on_batch_end


class KerasEpochEnd(_Base):
    """Broadcast ``logs[metric]`` with ``step=epoch`` after every epoch."""

    def __init__(self, reporter, metric="val_loss"):
        super().__init__()
        self.metric_name = metric
        self.reporter = reporter

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.reporter.broadcast(logs.get(self.metric_name, 0), epoch)


# framework-neutral aliases for PyTorch loops
BatchEnd = KerasBatchEnd
EpochEnd = KerasEpochEnd
