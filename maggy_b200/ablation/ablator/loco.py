"""Leave-One-Component-Out ablation (reference: maggy/ablation/ablator/loco.py:26-261).

One base trial plus one trial per included feature, layer, layer group / prefix and custom model.
Differences to the reference, which rebuilds Keras models from their JSON config and reads TFRecords
from the Hopsworks feature store:

* models are ``torch.nn.Module``s: a layer (by name), a set of layers, or every layer whose name starts
  with a prefix is *replaced by* ``nn.Identity()`` in a freshly generated base model — names are the
  dotted names of ``model.named_modules()``;
* datasets come from the user's dataset generator; if it accepts an ``ablated_feature`` argument the
  feature to drop is passed in (the reference ignores the ablated feature for custom generators).
"""
from __future__ import annotations

import functools
import os
import inspect

from ...core.exceptions import BadArgumentsError, NotSupportedError
from ...trial import Trial
from .abstractablator import AbstractAblator


def _pass_through_module():
    """``nn.Identity`` for sub-modules that are called with extra arguments (e.g. a decoder block that also receives
    rotary tables): returns its first positional argument and ignores the rest."""
    import torch.nn as nn

    class PassThrough(nn.Identity):
        def forward(self, x, *args, **kwargs):
            return x

    return PassThrough()


def remove_layers(model, layer_identifier):
    """Replace the identified sub-modules of a torch model with ``nn.Identity`` (in place).

    ``layer_identifier``: ``str`` = one module name; ``set`` with >1 names = all of them; ``set`` with one
    name = every module whose (leaf) name starts with that prefix (case-insensitive)."""
    import torch.nn as nn

    names = [n for n, _ in model.named_modules() if n]
    if isinstance(layer_identifier, str):
        targets = [n for n in names if n == layer_identifier]
    elif len(layer_identifier) > 1:
        targets = [n for n in names if n in layer_identifier]
    else:
        prefix = next(iter(layer_identifier)).lower()
        targets = [n for n in names if n.split(".")[-1].lower().startswith(prefix) or n.lower().startswith(prefix)]
    # drop children of already selected modules
    targets = [n for n in targets if not any(n != m and n.startswith(m + ".") for m in targets)]
    if not targets:
        raise ValueError(f"ablation: no module matches {layer_identifier!r}; available: {names[:20]}...")
    for name in targets:
        parent = model
        parts = name.split(".")
        for p in parts[:-1]:
            parent = getattr(parent, p)
        setattr(parent, parts[-1], _pass_through_module())
    return model


class LOCO(AbstractAblator):
    def __init__(self, ablation_study, final_store):
        super().__init__(ablation_study, final_store)
        self.base_dataset_generator = self.get_dataset_generator(ablated_feature=None)

    def get_number_of_trials(self):
        s = self.ablation_study
        return (len(s.features.included_features) + len(s.model.layers.included_layers)
                + len(s.model.layers.included_groups) + len(s.model.custom_model_generators) + 1)

    # ------------------------------------------------------------------ generators
    def get_dataset_generator(self, ablated_feature=None, dataset_type="store", shuffle_buffer_size=10000):
        """Dataset factory of one trial. A user-supplied generator wins (it receives ``ablated_feature`` if it takes
        that argument); otherwise the study's training dataset is read from the feature store with the ablated
        feature's column left out (reference :41-80 builds a TFRecord reader through hsfs; here the result is a torch
        ``DataLoader`` over the materialised training dataset — parquet / csv files, or hsfs when it is importable).
        The returned callable has the reference's signature ``create_dataset(num_epochs, batch_size)``."""
        gen = self.ablation_study.custom_dataset_generator
        if gen:
            try:
                accepts = "ablated_feature" in inspect.signature(gen).parameters
            except (TypeError, ValueError):
                accepts = False
            return functools.partial(gen, ablated_feature=ablated_feature) if accepts else gen
        if dataset_type not in ("store", "tfrecord", "parquet", "csv"):
            raise NotSupportedError("dataset type", dataset_type,
                                    " Use 'store' / 'parquet' / 'csv' or write your own custom dataset generator.")
        name = self.ablation_study.hops_training_dataset_name
        version = self.ablation_study.hops_training_dataset_version
        label = self.ablation_study.label_name
        if not name:
            raise NotSupportedError(
                "dataset type", dataset_type,
                " The study names no training dataset: pass AblationStudy(training_dataset_name, version, label) or a "
                "dataset_generator.")

        def create_dataset(num_epochs=1, batch_size=32, shuffle=True, seed=0):
            return feature_store_loader(name, version, label, ablated_feature, num_epochs, batch_size, shuffle,
                                        shuffle_buffer_size, seed)

        return create_dataset

    def get_model_generator(self, layer_identifier=None, custom_model_generator=None):
        if layer_identifier is not None and custom_model_generator is not None:
            raise BadArgumentsError("get_model_generator",
                                    "At least one of 'layer_identifier' or 'custom_model_generator' should be 'None'.")
        if custom_model_generator:
            return custom_model_generator[0]
        base = self.ablation_study.model.base_model_generator
        if layer_identifier is None:
            return base
        ident = layer_identifier if isinstance(layer_identifier, str) else set(layer_identifier)

        def model_generator():
            return remove_layers(base(), ident)

        return model_generator

    # ------------------------------------------------------------------ trials
    def create_trial_dict(self, ablated_feature=None, layer_identifier=None, custom_model_generator=None):
        d = {}
        if ablated_feature is None:
            d["dataset_function"] = self.base_dataset_generator
            d["ablated_feature"] = "None"
        else:
            d["dataset_function"] = self.get_dataset_generator(ablated_feature)
            d["ablated_feature"] = ablated_feature
        if layer_identifier is None and custom_model_generator is None:
            d["model_function"] = self.ablation_study.model.base_model_generator
            d["ablated_layer"] = "None"
        elif custom_model_generator is None:
            d["model_function"] = self.get_model_generator(layer_identifier=layer_identifier)
            if isinstance(layer_identifier, str):
                d["ablated_layer"] = layer_identifier
            elif len(layer_identifier) > 1:
                d["ablated_layer"] = str(sorted(layer_identifier))
            else:
                d["ablated_layer"] = "Layers prefixed " + str(next(iter(layer_identifier)))
        else:
            d["model_function"] = self.get_model_generator(custom_model_generator=custom_model_generator)
            d["ablated_layer"] = "Custom model: " + custom_model_generator[1]
        return d

    def initialize(self):
        s = self.ablation_study
        mk = lambda **kw: Trial(self.create_trial_dict(**kw), trial_type="ablation")  # noqa: E731
        self.trial_buffer.append(mk())
        for feature in sorted(s.features.included_features):
            self.trial_buffer.append(mk(ablated_feature=feature))
        for layer in sorted(s.model.layers.included_layers):
            self.trial_buffer.append(mk(layer_identifier=layer))
        for group in sorted(s.model.layers.included_groups, key=lambda g: sorted(g)):
            self.trial_buffer.append(mk(layer_identifier=set(group)))
        for custom in s.model.custom_model_generators:
            self.trial_buffer.append(mk(custom_model_generator=custom))

    def get_trial(self, trial=None):
        return self.trial_buffer.pop() if self.trial_buffer else None

    def finalize_experiment(self, trials):
        return


def _read_training_dataset(name, version):
    """The materialised training dataset as a pandas DataFrame: hsfs when importable, else the parquet / csv files
    under the environment's training-dataset path (``EnvSing.get_training_dataset_path``)."""
    import pandas as pd

    from ...core.environment.singleton import EnvSing

    env = EnvSing.get_instance()
    try:
        conn = env.connect_hsfs()
        if conn is not None:
            return conn.get_feature_store().get_training_dataset(name, version).read()
    except Exception:  # noqa: BLE001 - fall through to the files
        pass
    path = env.get_training_dataset_path(name, training_dataset_version=version)
    if not path:
        path = os.path.join(os.environ.get("MAGGY_TRAINING_DATASETS", os.path.join(os.getcwd(), "Training_Datasets")),
                            f"{name}_{version}")
    if os.path.isdir(path):
        files = sorted(os.path.join(dp, f) for dp, _, fs in os.walk(path) for f in fs
                       if f.endswith((".parquet", ".csv")) and not f.startswith(("_", ".")))
    else:
        files = [p for p in (path, path + ".parquet", path + ".csv") if os.path.isfile(p)][:1]
    if not files:
        raise FileNotFoundError(f"training dataset {name}_{version}: no parquet / csv files under {path}")
    frames = [pd.read_parquet(f) if f.endswith(".parquet") else pd.read_csv(f) for f in files]
    return pd.concat(frames, ignore_index=True) if len(frames) > 1 else frames[0]


def feature_store_loader(name, version, label, ablated_feature, num_epochs, batch_size, shuffle=True,
                         shuffle_buffer_size=10000, seed=0):
    """Batches ``(features[float32, B x F], label[B])`` of the training dataset without ``ablated_feature``, for
    ``num_epochs`` epochs; pinned when a GPU is present so that the H2D copy of each batch can be asynchronous."""
    import numpy as np
    import torch

    df = _read_training_dataset(name, version)
    if label not in df.columns:
        raise KeyError(f"label column {label!r} is not in training dataset {name}_{version}: {list(df.columns)}")
    features = [c for c in df.columns if c != label]
    if ablated_feature is not None:
        if ablated_feature not in features:
            raise KeyError(f"feature {ablated_feature!r} is not in training dataset {name}_{version}")
        features.remove(ablated_feature)
    x = torch.from_numpy(np.ascontiguousarray(df[features].to_numpy(dtype=np.float32)))
    y_np = df[label].to_numpy()
    y = torch.from_numpy(np.ascontiguousarray(y_np.astype(np.int64 if np.issubdtype(y_np.dtype, np.integer)
                                                          else np.float32)))
    if torch.cuda.is_available():
        x, y = x.pin_memory(), y.pin_memory()
    n = x.shape[0]
    gen = torch.Generator().manual_seed(seed)

    class _Loader:
        feature_names = features

        def __len__(self):
            return num_epochs * ((n + batch_size - 1) // batch_size)

        def __iter__(self):
            for _ in range(num_epochs):
                order = torch.randperm(n, generator=gen) if shuffle else torch.arange(n)
                for i in range(0, n, batch_size):
                    idx = order[i:i + batch_size]
                    yield x[idx], y[idx]

    return _Loader()
