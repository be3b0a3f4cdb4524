"""Databricks flavour (reference: maggy/core/environment/databricks.py:23-78): artefacts under ``/dbfs/maggy_log/``
(the DBFS FUSE mount, so plain file operations work), worker count from the cluster's usage tags."""
from __future__ import annotations

import os

from .base import BaseEnv

_TAG = "spark.databricks.clusterUsageTags."


class DatabricksEnv(BaseEnv):
    def __init__(self):
        self.log_dir = os.environ.get("MAGGY_LOG_DIR") or "/dbfs/maggy_log/"
        os.makedirs(self.log_dir, exist_ok=True)

    def mkdir(self, path):
        return os.makedirs(path, exist_ok=True)

    def project_path(self, project=None, exclude_nn_addr=False):
        return "/dbfs/"

    def get_logdir(self, app_id, run_id):
        # the reference keeps one flat log directory per cluster; runs are still separated here so that reruns do not
        # clobber each other's trial directories
        return os.path.join(self.log_dir, str(app_id), str(run_id))

    def cluster_workers(self, conf) -> int:
        """Workers of the attached cluster from its Spark conf (anything with ``.get(key, default)``): the autoscaling
        maximum when the cluster autoscales, else its fixed size (reference get_executors :35-63)."""
        get = conf.get
        if get(_TAG + "clusterScalingType", None) == "autoscaling":
            n = int(get(_TAG + "clusterMaxWorkers", "-1"))
            if n == -1:
                raise KeyError(f'Failed to find "{_TAG}clusterMaxWorkers" property, but clusterScalingType is set to '
                               "autoscaling.")
            return n
        n = int(get(_TAG + "clusterWorkers", "-1"))
        if n == -1:
            raise KeyError(f'Failed to find "{_TAG}clusterWorkers" property.')
        return n

    def get_executors(self, default=None):
        """GPUs of this box bound the worker pool; on a cluster whose size is known from the environment
        (``DATABRICKS_CLUSTER_WORKERS``, exported by init scripts) the smaller of the two is used."""
        n = super().get_executors(default)
        declared = os.environ.get("DATABRICKS_CLUSTER_WORKERS")
        return min(n, int(declared)) if declared and int(declared) > 0 else n
