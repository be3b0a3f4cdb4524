"""Hopsworks flavour (reference: maggy/core/environment/hopsworks.py:33-275).

What the reference does through the ``hops`` client library is done here against the Hopsworks REST API directly
(``urllib``; the client libraries only exist inside a Hopsworks cluster):

* the driver's LOG endpoint is announced with ``POST /hopsworks-api/api/maggy/drivers`` ``{hostIp, port, appId,
  secret}`` so the Jupyter front-end can poll progress (reference :136-190); a failed registration only logs
  "No connection to Hopsworks for logging.", exactly like the reference;
* experiment metadata is attached with ``PUT /hopsworks-api/api/project/<id>/experiments/<ml_id>?type=<op>`` (the
  xattr the reference writes through ``experiment_utils._attach_experiment_xattr``, :75-76, 201-226), plus a local
  sidecar so nothing is lost when the endpoint is unreachable;
* ``.summary.json`` (``{"combinations": [{parameters, outputs}]}``) is written at the end (:214-217);
* file operations go to HopsFS through ``hops.hdfs`` when that module is importable, else to the local filesystem
  under ``$HOPSWORKS_PROJECT_PATH/Experiments`` (or the default log directory).

Configuration comes from the environment Hopsworks sets for its jobs: ``REST_ENDPOINT``, ``HOPSWORKS_PROJECT_ID``,
``HOPSWORKS_PROJECT_NAME``, ``HADOOP_USER_NAME`` / ``HOPSWORKS_USER``, and an API key (``HOPSWORKS_API_KEY``) or a JWT
(``HOPSWORKS_JWT`` or the ``token.jwt`` file in the working directory).
"""
from __future__ import annotations

import json
import os
import ssl
import urllib.error
import urllib.request

from .base import BaseEnv

REST_RESOURCE = "hopsworks-api/api"


class HopsworksEnv(BaseEnv):
    APP_ID = None

    def __init__(self):
        try:
            from hops import hdfs as hopshdfs  # type: ignore

            self._hdfs = hopshdfs
            self.log_dir = os.path.join(hopshdfs.project_path(), "Experiments")
        except Exception:
            self._hdfs = None
            root = os.environ.get("HOPSWORKS_PROJECT_PATH")
            if root and not os.environ.get("MAGGY_LOG_DIR"):
                self.log_dir = os.path.join(root, "Experiments")
                os.makedirs(self.log_dir, exist_ok=True)
            else:
                super().__init__()
        self.endpoint = os.environ.get("REST_ENDPOINT", "").rstrip("/")
        self.project_id = os.environ.get("HOPSWORKS_PROJECT_ID", "")

    # ------------------------------------------------------------------ REST
    def _auth_header(self) -> dict:
        if os.environ.get("HOPSWORKS_API_KEY"):
            return {"Authorization": "ApiKey " + os.environ["HOPSWORKS_API_KEY"]}
        token = os.environ.get("HOPSWORKS_JWT")
        if not token and os.path.exists("token.jwt"):
            token = open("token.jwt").read().strip()
        return {"Authorization": "Bearer " + token} if token else {}

    def send_request(self, method: str, resource: str, data=None, headers=None, timeout_s: float = 5.0):
        """HTTP request against ``REST_ENDPOINT``; returns ``(status_code, body_text)``."""
        if not self.endpoint:
            raise ConnectionError("REST_ENDPOINT is not set")
        url = self.endpoint + "/" + resource.lstrip("/")
        body = data.encode() if isinstance(data, str) else data
        req = urllib.request.Request(url, data=body, method=method.upper())
        for k, v in {**self._auth_header(), **(headers or {})}.items():
            req.add_header(k, v)
        ctx = None
        if url.startswith("https") and os.environ.get("HOPSWORKS_VERIFY_TLS", "1") == "0":
            ctx = ssl._create_unverified_context()
        try:
            with urllib.request.urlopen(req, timeout=timeout_s, context=ctx) as resp:
                return resp.status, resp.read().decode("utf-8", "replace")
        except urllib.error.HTTPError as e:
            return e.code, e.read().decode("utf-8", "replace")

    def set_app_id(self, app_id):
        self.APP_ID = app_id

    def announce_driver(self, port: int, secret: str, app_id, exp_driver=None) -> bool:
        """Register the driver's LOG endpoint with Hopsworks (reference connect_host, :136-190)."""
        self.APP_ID = app_id
        payload = {"hostIp": self.get_ip_address(), "port": int(port), "appId": str(app_id), "secret": secret}
        try:
            status, _ = self.send_request("POST", f"/{REST_RESOURCE}/maggy/drivers", data=json.dumps(payload),
                                          headers={"Content-Type": "application/json"})
            if status // 100 == 2:
                return True
            msg = "No connection to Hopsworks for logging."
        except Exception as e:  # noqa: BLE001
            msg = f"Connection failed to Hopsworks. No logging. ({e})"
        print(msg)
        if exp_driver is not None:
            exp_driver.log(msg)
        return False

    def attach_experiment_xattr(self, exp_ml_id, experiment_json, command):
        """Experiment metadata as an xattr of the experiment directory, via REST; always mirrored to a sidecar file."""
        doc = json.dumps(experiment_json, default=str)
        try:
            self.dump(doc, os.path.join(self.log_dir, f".{exp_ml_id}.{command}.json"))
        except Exception:  # noqa: BLE001
            pass
        if not (self.endpoint and self.project_id):
            return None
        try:
            status, body = self.send_request(
                "PUT", f"/{REST_RESOURCE}/project/{self.project_id}/experiments/{exp_ml_id}?type={command}", data=doc,
                headers={"Content-Type": "application/json"})
            return status
        except Exception:  # noqa: BLE001 - metadata must not fail the experiment
            return None

    # ------------------------------------------------------------------ identity
    def project_name(self):
        return os.environ.get("HOPSWORKS_PROJECT_NAME", "")

    def project_path(self, project=None, exclude_nn_addr=False):
        if self._hdfs is not None:
            return self._hdfs.project_path(project=project, exclude_nn_addr=exclude_nn_addr)
        return os.environ.get("HOPSWORKS_PROJECT_PATH", os.getcwd())

    def get_user(self):
        return os.environ.get("HOPSWORKS_USER") or os.environ.get("HADOOP_USER_NAME") or super().get_user()

    def str_or_byte(self, s):
        # HopsFS file objects want bytes; the local fallback wants str
        return s.encode() if self._hdfs is not None else s

    # ------------------------------------------------------------------ HopsFS (when the client library exists)
    def exists(self, path, project=None):
        return self._hdfs.exists(path, project=project) if self._hdfs is not None else super().exists(path)

    def mkdir(self, path, project=None):
        return self._hdfs.mkdir(path, project=project) if self._hdfs is not None else super().mkdir(path)

    def dump(self, data, path):
        return self._hdfs.dump(data, path) if self._hdfs is not None else super().dump(data, path)

    def load(self, path):
        return self._hdfs.load(path) if self._hdfs is not None else super().load(path)

    def open_file(self, path, project=None, flags="r", buff_size=-1):
        if self._hdfs is not None:
            return self._hdfs.open_file(path, project=project, flags=flags, buff_size=max(buff_size, 0))
        return super().open_file(path, flags=flags, buff_size=buff_size)

    def isdir(self, dir_path, project=None):
        return self._hdfs.isdir(dir_path, project=project) if self._hdfs is not None else super().isdir(dir_path)

    def ls(self, dir_path, recursive=False, exclude_nn_addr=None):
        if self._hdfs is not None:
            return self._hdfs.ls(dir_path, recursive=recursive, exclude_nn_addr=exclude_nn_addr)
        return super().ls(dir_path)

    def delete(self, path, recursive=False):
        return self._hdfs.delete(path, recursive=recursive) if self._hdfs is not None else \
            super().delete(path, recursive=recursive)

    # ------------------------------------------------------------------ experiment end
    def finalize_experiment(self, experiment_json, metric, app_id, run_id, state, duration, logdir,
                            best_logdir, optimization_key):
        from ... import util

        try:
            outputs = util.build_summary_json(logdir)
            if outputs:
                self.dump(outputs, os.path.join(logdir, ".summary.json"))
        except Exception:  # noqa: BLE001
            pass
        if isinstance(experiment_json, dict):
            if best_logdir:
                root = self.project_path()
                experiment_json["bestDir"] = best_logdir[len(root):] if best_logdir.startswith(root) else best_logdir
            experiment_json.update({"optimizationKey": optimization_key, "metric": metric, "state": state,
                                    "duration": duration})
            self.attach_experiment_xattr(f"{app_id}_{run_id}", experiment_json, "FULL_UPDATE")
        return super().finalize_experiment(experiment_json, metric, app_id, run_id, state, duration, logdir,
                                           best_logdir, optimization_key)

    # ------------------------------------------------------------------ feature store (training datasets)
    def get_training_dataset_path(self, training_dataset, featurestore=None, training_dataset_version=1):
        """``<project>/<project>_Training_Datasets/<name>_<version>`` — the Hopsworks layout; with the ``hops`` client
        the feature store answers, without it the path is derived."""
        try:
            from hops import featurestore as fs  # type: ignore

            return fs.get_training_dataset_path(training_dataset, featurestore=featurestore,
                                                training_dataset_version=training_dataset_version)
        except Exception:  # noqa: BLE001
            project = self.project_name() or "project"
            return os.path.join(self.project_path(), f"{project}_Training_Datasets",
                                f"{training_dataset}_{training_dataset_version}")

    def connect_hsfs(self, engine="training"):
        """hsfs connection when the client library exists (reference :274-275), else None."""
        try:
            import hsfs  # type: ignore

            return hsfs.connection(engine=engine)
        except Exception:  # noqa: BLE001
            return None
