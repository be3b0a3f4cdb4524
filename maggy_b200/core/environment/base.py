"""Default (local filesystem) environment: where experiment artefacts go and how big the worker pool is.

Reference: maggy/core/environment/base.py:25-222. The Spark-specific parts (executor count from the
Spark conf, socket binding for the RPC server) are replaced by the per-GPU worker pool: the pool size
comes from the visible GPUs (or MAGGY_NUM_EXECUTORS) and the "server address" is the name of the
shared-memory control segment.
"""
from __future__ import annotations

import json
import os
import shutil
import socket
import warnings


class BaseEnv:
    """Artefacts live under ``./experiment_log/<app_id>/<run_id>/`` (override: MAGGY_LOG_DIR)."""

    def __init__(self):
        self.log_dir = os.environ.get("MAGGY_LOG_DIR") or os.path.join(os.getcwd(), "experiment_log")
        os.makedirs(self.log_dir, exist_ok=True)

    # ------------------------------------------------------------------ experiment bookkeeping
    def set_ml_id(self, app_id=0, run_id=0):
        os.environ["ML_ID"] = f"{app_id}_{run_id}"

    def create_experiment_dir(self, app_id, run_id):
        path = self.get_logdir(app_id, run_id)
        if os.path.exists(path):
            shutil.rmtree(path)
        os.makedirs(path)

    def get_logdir(self, app_id, run_id):
        return os.path.join(self.log_dir, str(app_id), str(run_id))

    def populate_experiment(self, config, app_id, run_id, exp_function):
        """Metadata record of a starting experiment (returned, and later finalised)."""
        return {"name": getattr(config, "name", "base"), "description": getattr(config, "description", ""),
                "function": exp_function, "app_id": str(app_id), "run_id": run_id, "state": "RUNNING"}

    def attach_experiment_xattr(self, exp_ml_id, experiment_json, command):
        pass

    def finalize_experiment(self, experiment_json, metric, app_id, run_id, state, duration, logdir,
                            best_logdir, optimization_key):
        if isinstance(experiment_json, dict):
            experiment_json.update({"state": state, "duration": duration, "metric": metric,
                                    "best_logdir": best_logdir, "optimization_key": optimization_key})
        return experiment_json

    def init_ml_tracking(self, app_id, run_id):
        pass

    def log_searchspace(self, app_id, run_id, searchspace):
        pass

    # ------------------------------------------------------------------ filesystem
    def exists(self, path):
        return os.path.exists(path)

    def mkdir(self, path):
        return os.makedirs(path, exist_ok=True)

    def isdir(self, dir_path, project=None):
        return os.path.isdir(dir_path)

    def ls(self, dir_path):
        """Full paths of the entries of ``dir_path``."""
        return [os.path.join(dir_path, e) for e in sorted(os.listdir(dir_path))]

    def delete(self, path, recursive=False):
        if not self.exists(path):
            warnings.warn(f"Could not delete the file in {path}.\nFile does not exists.")
        elif os.path.isdir(path):
            if recursive:
                shutil.rmtree(path)
            elif not os.listdir(path):
                os.rmdir(path)
            else:
                warnings.warn(f"Could not delete the dir {path}, not empty.\nUse recursive=True when calling this function")
        else:
            os.remove(path)

    def dump(self, data, path):
        parent = os.path.dirname(path)
        if parent:
            os.makedirs(parent, exist_ok=True)
        with self.open_file(path, flags="w") as f:
            f.write(data)

    def load(self, path):
        with self.open_file(path, flags="r") as f:
            return f.read()

    def open_file(self, path, flags="r", buff_size=-1):
        return open(path, mode=flags, buffering=buff_size)

    def str_or_byte(self, s):
        return s

    def project_path(self):
        return os.getcwd()

    def get_user(self):
        return os.environ.get("USER", "")

    def project_name(self):
        return ""

    def upload_file_output(self, retval, exec_logdir):
        pass

    def convert_return_file_to_arr(self, return_file):
        return json.loads(self.load(return_file))

    def build_summary_json(self, logdir):
        from ... import util

        return util.build_summary_json(logdir)

    # ------------------------------------------------------------------ cluster facts
    def get_ip_address(self):
        try:
            return socket.gethostbyname(socket.gethostname())
        except OSError:
            return "127.0.0.1"

    def _upload_file_output(self, retval, hdfs_exec_logdir):
        """Hook for environments with a remote file system (reference: environment/base.py:152-153); local: nothing."""

    def connect_host(self, server_sock, server_host_port, exp_driver):
        """Bind + listen a driver-side socket (reference: environment/base.py:137-150, the TCP RPC server's listener).
        The control plane here is a shared-memory segment, so the only socket a driver opens is the optional external
        LOG endpoint (core/logserver.py), which binds through this hook."""
        if not server_host_port:
            server_sock.bind(("", 0))
            server_host_port = (self.get_ip_address(), server_sock.getsockname()[1])
        else:
            server_sock.bind(server_host_port)
        server_sock.listen(10)
        return server_sock, server_host_port

    def get_client(self, server_addr, partition_id, hb_interval, secret, sock):
        """The worker-side end of the control channel (reference: environment/base.py:217-222 builds the TCP ``Client``).
        ``server_addr`` is the control segment's name here; heartbeat interval and secret have no role on a 0600
        shared-memory segment handed to the worker by name (core/rpc.py module docstring)."""
        from ..rpc import WorkerChannel

        return WorkerChannel(server_addr, partition_id)

    def get_executors(self, default=None):
        from ... import util

        return util.num_executors(default)

    def get_constants(self):
        pass

    # feature-store hooks (Hopsworks only)
    def get_training_dataset_path(self, training_dataset, featurestore=None, training_dataset_version=1):
        pass

    def get_training_dataset_tf_record_schema(self, training_dataset, training_dataset_version=1, featurestore=None):
        pass

    def get_featurestore_metadata(self, featurestore=None, update_cache=False):
        pass

    def connect_hsfs(self, engine="training"):
        return None
