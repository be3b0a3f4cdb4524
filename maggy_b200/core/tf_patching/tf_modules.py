"""Strategy-scope model wrapper (reference: maggy/core/tf_patching/tf_modules.py:18-45).

The user hands a Keras model CLASS to ``TfDistributedConfig(model=...)``; the train_fn instantiates it. Variables of
a ``tf.distribute`` strategy must be created inside ``strategy.scope()``, so the class the train_fn receives is a
subclass whose constructor runs under that scope. Framework-agnostic: anything with a ``scope()`` context manager
works as the strategy (the tests use a stand-in; TensorFlow is optional in this image).
"""
from __future__ import annotations


def get_wrapped_model(model, strategy, is_chief: bool):
    """Return a subclass of ``model`` that builds itself under ``strategy.scope()``.

    ``is_chief`` is recorded on the class (``_maggy_is_chief``): Keras callbacks that write checkpoints / summaries
    conventionally run on the chief only."""
    if model is None:
        return None
    if not isinstance(model, type):
        raise TypeError("TfDistributedConfig.model has to be a model CLASS (the train_fn instantiates it), "
                        f"got an instance of {type(model).__name__}")

    class MaggyStrategyModel(model):  # type: ignore[misc, valid-type]
        _maggy_is_chief = bool(is_chief)
        _maggy_strategy = strategy

        def __init__(self, *args, **kwargs):
            with strategy.scope():
                super().__init__(*args, **kwargs)

    MaggyStrategyModel.__name__ = f"Maggy{model.__name__}"
    MaggyStrategyModel.__qualname__ = MaggyStrategyModel.__name__
    return MaggyStrategyModel
