"""Hopsworks flavour (reference: maggy/core/environment/hopsworks.py:33-275).

The reference stores artefacts on HopsFS, registers the driver endpoint with the Hopsworks REST API
so Jupyter can poll logs, and records experiment metadata as xattrs. None of the ``hops``/``hsfs``
client libraries exist outside a Hopsworks cluster, so this class keeps the hook points and the
``.summary.json`` artefact, and degrades to the local filesystem when the clients are missing.
"""
import json
import os

from .base import BaseEnv


class HopsworksEnv(BaseEnv):
    def __init__(self):
        try:
            from hops import hdfs as hopshdfs  # type: ignore

            self._hdfs = hopshdfs
            self.log_dir = os.path.join(hopshdfs.project_path(), "Experiments")
        except Exception:
            self._hdfs = None
            super().__init__()

    def project_name(self):
        return os.environ.get("HOPSWORKS_PROJECT_NAME", "")

    def get_user(self):
        return os.environ.get("HADOOP_USER_NAME", super().get_user())

    def str_or_byte(self, s):
        # HopsFS file objects want bytes; the local fallback wants str
        return s.encode() if self._hdfs is not None else s

    def finalize_experiment(self, experiment_json, metric, app_id, run_id, state, duration, logdir,
                            best_logdir, optimization_key):
        from ... import util

        try:
            self.dump(util.build_summary_json(logdir), os.path.join(logdir, ".summary.json"))
        except Exception:
            pass
        return super().finalize_experiment(experiment_json, metric, app_id, run_id, state, duration, logdir,
                                           best_logdir, optimization_key)

    def attach_experiment_xattr(self, exp_ml_id, experiment_json, command):
        # xattrs are a HopsFS feature; keep a sidecar so the information is not lost locally
        try:
            path = os.path.join(self.log_dir, f".{exp_ml_id}.{command}.json")
            self.dump(json.dumps(experiment_json, default=str), path)
        except Exception:
            pass
