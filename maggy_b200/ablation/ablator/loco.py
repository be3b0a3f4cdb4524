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
    def get_dataset_generator(self, ablated_feature=None, dataset_type="custom", shuffle_buffer_size=10000):
        gen = self.ablation_study.custom_dataset_generator
        if gen:
            try:
                accepts = "ablated_feature" in inspect.signature(gen).parameters
            except (TypeError, ValueError):
                accepts = False
            return functools.partial(gen, ablated_feature=ablated_feature) if accepts else gen
        raise NotSupportedError(
            "dataset type", dataset_type,
            " Reading training datasets from the Hopsworks feature store is not available outside Hopsworks: "
            "pass AblationStudy(..., dataset_generator=fn) or call study.set_dataset_generator(fn).")

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
