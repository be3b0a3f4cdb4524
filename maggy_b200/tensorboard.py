"""TensorBoard bridge (reference: maggy/tensorboard.py:25-107).

``tensorboard.logdir()`` returns the directory the running trial should write summaries to. The
HParams-plugin records (experiment config for the searchspace, per-trial hparams) are written with
``torch.utils.tensorboard`` instead of TensorFlow; if tensorboard is not installed a JSON sidecar is
written instead so no information is lost.
"""
from __future__ import annotations

import json
import os

_tensorboard_dir = None


def _register(trial_dir):
    global _tensorboard_dir
    _tensorboard_dir = trial_dir


def logdir():
    """Path of the TensorBoard log directory of the current trial/experiment."""
    return _tensorboard_dir


_METRICS = [
    ("epoch_accuracy", "validation", "accuracy (val.)"),
    ("epoch_loss", "validation", "loss (val.)"),
    ("epoch_accuracy", "train", "accuracy (train)"),
    ("epoch_loss", "train", "loss (train)"),
]


def _hparam_domains(searchspace):
    doms = {}
    for name, kind in searchspace.names().items():
        region = searchspace.get(name)
        if kind == "DOUBLE":
            doms[name] = ("real", [float(region[0]), float(region[1])])
        elif kind == "INTEGER":
            doms[name] = ("int", [int(region[0]), int(region[1])])
        else:
            doms[name] = ("discrete", list(region))
    return doms


def _summary(tag, plugin_data):
    """A Summary proto carrying HParams-plugin data (what tensorboard.plugins.hparams.summary builds —
    that module imports TensorFlow, so the few lines are repeated here on the TF-free protos)."""
    from tensorboard.compat.proto.summary_pb2 import Summary, SummaryMetadata
    from tensorboard.plugins.hparams import metadata, plugin_data_pb2

    plugin_data.version = metadata.PLUGIN_DATA_VERSION
    assert isinstance(plugin_data, plugin_data_pb2.HParamsPluginData)
    smd = SummaryMetadata(plugin_data=SummaryMetadata.PluginData(plugin_name=metadata.PLUGIN_NAME,
                                                                 content=plugin_data.SerializeToString()))
    return Summary(value=[Summary.Value(tag=tag, metadata=smd)])


def _write_hparams_config(log_dir, searchspace):
    """Experiment-level HParams config: one domain per hyperparameter + the four standard metrics."""
    doms = _hparam_domains(searchspace)
    try:
        from google.protobuf import struct_pb2
        from tensorboard.plugins.hparams import api_pb2, metadata, plugin_data_pb2
        from torch.utils.tensorboard import SummaryWriter

        infos = []
        for name, (kind, region) in doms.items():
            info = api_pb2.HParamInfo(name=name)
            if kind in ("real", "int"):
                info.type = api_pb2.DATA_TYPE_FLOAT64
                info.domain_interval.min_value, info.domain_interval.max_value = float(region[0]), float(region[1])
            elif all(isinstance(v, str) for v in region):
                info.type = api_pb2.DATA_TYPE_STRING
                info.domain_discrete.extend([struct_pb2.Value(string_value=v) for v in region])
            else:
                info.type = api_pb2.DATA_TYPE_FLOAT64
                info.domain_discrete.extend([struct_pb2.Value(number_value=float(v)) for v in region])
            infos.append(info)
        metrics = [api_pb2.MetricInfo(name=api_pb2.MetricName(group=g, tag=t), display_name=d)
                   for t, g, d in _METRICS]
        exp = api_pb2.Experiment(hparam_infos=infos, metric_infos=metrics)
        w = SummaryWriter(log_dir)
        w.file_writer.add_summary(_summary(metadata.EXPERIMENT_TAG, plugin_data_pb2.HParamsPluginData(experiment=exp)))
        w.close()
    except Exception:
        os.makedirs(log_dir, exist_ok=True)
        with open(os.path.join(log_dir, "hparams_config.json"), "w") as f:
            json.dump({"hparams": doms, "metrics": _METRICS}, f)


def _write_hparams(hparams, trial_id):
    """Per-trial HParams session start, written into the current ``logdir()``."""
    clean = {k: (v if isinstance(v, (int, float, str, bool)) else str(v)) for k, v in hparams.items()}
    try:
        from google.protobuf import struct_pb2
        from tensorboard.plugins.hparams import api_pb2, metadata, plugin_data_pb2
        from torch.utils.tensorboard import SummaryWriter

        ssi = plugin_data_pb2.SessionStartInfo(group_name=str(trial_id), start_time_secs=__import__("time").time())
        for k, v in clean.items():
            if isinstance(v, bool):
                ssi.hparams[k].bool_value = v
            elif isinstance(v, (int, float)):
                ssi.hparams[k].number_value = float(v)
            else:
                ssi.hparams[k].string_value = v
        w = SummaryWriter(_tensorboard_dir)
        w.file_writer.add_summary(_summary(metadata.SESSION_START_INFO_TAG,
                                           plugin_data_pb2.HParamsPluginData(session_start_info=ssi)))
        w.close()
    except Exception:
        if _tensorboard_dir:
            os.makedirs(_tensorboard_dir, exist_ok=True)
            with open(os.path.join(_tensorboard_dir, "hparams.json"), "w") as f:
                json.dump({"trial_id": trial_id, "hparams": clean}, f)
