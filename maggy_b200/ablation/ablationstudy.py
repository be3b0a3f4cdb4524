"""Declarative description of an ablation study (reference: maggy/ablation/ablationstudy.py:18-408).

    study = AblationStudy("titanic_train", 1, label_name="survived")
    study.features.include("pclass", "fare")
    study.model.layers.include("block2", "block3")
    study.model.layers.include_groups(["fc1", "fc2"])
    study.model.layers.include_groups(prefix="head")
    study.model.set_base_model_generator(make_model)

``features`` / ``model.layers`` are sets of component names; the ablator (LOCO) turns every included
component into one trial that trains without it.
"""
from __future__ import annotations


class _NameSet:
    """A set of string identifiers with ``include(*names_or_lists)`` / ``exclude(...)``."""

    _what = "items"

    def __init__(self):
        self._items = set()

    def _flatten(self, args, verb):
        for arg in args:
            for item in (arg if type(arg) is list else [arg]):
                if type(item) is not str:
                    raise ValueError(f"{self._what}.{verb}() only accepts strings or lists of strings, but it "
                                     f"received {item} which is of type '{type(item).__name__}'.")
                yield item

    def include(self, *args):
        for name in self._flatten(args, "include"):
            self._items.add(name)

    def exclude(self, *args):
        for name in self._flatten(args, "exclude"):
            self._items.discard(name)


class Features(_NameSet):
    _what = "features"

    @property
    def included_features(self):
        return self._items

    def exclude(self, *args):
        for name in self._flatten(args, "exclude"):
            if name in self._items:
                self._items.remove(name)
                print(f"Feature '{name}' is excluded from the ablation study.")

    def list_all(self):
        for f in self._items:
            print(f)


class Layers(_NameSet):
    _what = "layers"

    def __init__(self):
        super().__init__()
        self.included_groups = set()  # frozensets: >1 names = explicit group, 1 name = prefix group

    @property
    def included_layers(self):
        return self._items

    @staticmethod
    def _check_prefix(prefix, verb):
        if prefix is not None and type(prefix) is not str:
            raise ValueError(f"`prefix` argument of layers.{verb}() should either be a `NoneType` or a `str`, but it "
                             f"received {prefix} (of type '{type(prefix).__name__}'.")

    def include_groups(self, *args, prefix=None):
        """Each list of (>1) layer names is ablated together; ``prefix=`` ablates every layer whose
        name starts with the prefix."""
        self._check_prefix(prefix, "include_groups")
        if prefix is not None:
            self.included_groups.add(frozenset([prefix]))
        for arg in args:
            if type(arg) is list and len(arg) > 1:
                self.included_groups.add(frozenset(arg))
            elif type(arg) is list and len(arg) == 1:
                raise ValueError(f"layers.include_groups() received a list ( {arg} ) with only one element: Did you "
                                 "want to include a single layer in your ablation study? then you should use "
                                 "layers.include().")
            else:
                raise ValueError("layers.include_groups() only accepts a prefix string, or lists (with more than one "
                                 f"element) of strings, but it received {arg} (of type '{type(arg).__name__}').")

    def exclude_groups(self, *args, prefix=None):
        self._check_prefix(prefix, "exclude_groups")
        if prefix is not None:
            self.included_groups.discard(frozenset([prefix]))
        for arg in args:
            if type(arg) is list and len(arg) > 1:
                self.included_groups.discard(frozenset(arg))
            else:
                raise ValueError("layers.exclude_groups() only accepts a prefix string, or lists (with more than one "
                                 f"element) of strings, but it received {arg} (of type '{type(arg).__name__}').")

    def print_all(self):
        if self._items:
            print("Included single layers are: \n")
            for layer in self._items:
                print(layer)
        else:
            print("There are no single layers in this ablation study configuration.")

    def print_all_groups(self):
        if not self.included_groups:
            print("There are no layer groups in this ablation study configuration.")
            return
        print("Included layer groups are: \n")
        for g in self.included_groups:
            if len(g) > 1:
                print("--- Layer group " + str(list(g)))
            else:
                print('---- All layers prefixed "' + str(list(g)[0]) + '"')


class Model:
    def __init__(self):
        self.layers = Layers()
        self.base_model_generator = None
        self.custom_model_generators = []

    def set_base_model_generator(self, base_model_generator):
        """``base_model_generator()`` must return the full (un-ablated) model."""
        self.base_model_generator = base_model_generator

    def add_custom_model_generator(self, custom_model_generator, model_identifier):
        """Register a hand-written variant to be evaluated as its own trial."""
        self.custom_model_generators.append((custom_model_generator, model_identifier))


class AblationStudy:
    def __init__(self, training_dataset_name, training_dataset_version, label_name, **kwargs):
        self.features = Features()
        self.model = Model()
        self.hops_training_dataset_name = training_dataset_name
        self.hops_training_dataset_version = training_dataset_version
        self.label_name = label_name
        self.custom_dataset_generator = kwargs.get("dataset_generator", False)

    def set_dataset_generator(self, dataset_generator):
        self.custom_dataset_generator = dataset_generator

    def to_dict(self):
        return {
            "training_dataset_name": self.hops_training_dataset_name,
            "training_dataset_version": self.hops_training_dataset_version,
            "label_name": self.label_name,
            "included_features": list(self.features.included_features),
            "included_layers": list(self.model.layers.included_layers),
            "custom_dataset_generator": bool(self.custom_dataset_generator),
        }
