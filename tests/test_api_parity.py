"""Name-level API parity with the reference tree: every public (non-underscore) top-level function, class and method of
the reference's modules resolves on the module that plays its role here. Skipped where the reference tree is absent."""
import ast
import importlib
import os

import pytest

REF = os.environ.get("MAGGY_REFERENCE_TREE", "/root/reference/maggy")

MAPPING = {
    "core/experiment_driver/optimization_driver.py": "maggy_b200.core.drivers.optimization_driver",
    "core/experiment_driver/ablation_driver.py": "maggy_b200.core.drivers.ablation_driver",
    "core/experiment_driver/base_driver.py": "maggy_b200.core.drivers.base_driver",
    "core/experiment_driver/torch_distributed_training_driver.py": "maggy_b200.core.drivers.torch_distributed_training_driver",
    "core/experiment_driver/tf_distributed_training_driver.py": "maggy_b200.core.drivers.tf_distributed_training_driver",
    "core/experiment_driver/python_driver.py": "maggy_b200.core.drivers.base",
    "core/experiment_driver/spark_driver.py": "maggy_b200.core.drivers.base",
    "core/executors/trial_executor.py": "maggy_b200.core.executors.trial_executor",
    "core/executors/torch_dist_executor.py": "maggy_b200.core.executors.torch_dist_executor",
    "core/executors/tf_dist_executor.py": "maggy_b200.core.executors.tf_dist_executor",
    "core/executors/base_executor.py": "maggy_b200.core.executors.base_executor",
    "core/environment/hopsworks.py": "maggy_b200.core.environment.hopsworks",
    "core/environment/base.py": "maggy_b200.core.environment.base",
    "core/environment/databricks.py": "maggy_b200.core.environment.databricks",
    "core/environment/singleton.py": "maggy_b200.core.environment.singleton",
    "core/reporter.py": "maggy_b200.core.reporter",
    "core/config.py": "maggy_b200.core.config",
    "core/exceptions.py": "maggy_b200.core.exceptions",
    "core/rpc.py": "maggy_b200.core.rpc",
    "core/patching/modules.py": "maggy_b200.core.patching.modules",
    "core/patching/optim.py": "maggy_b200.core.patching.optim",
    "core/patching/dataloader.py": "maggy_b200.core.patching.dataloader",
    "core/tf_patching/tf_modules.py": "maggy_b200.core.tf_patching.tf_modules",
    "experiment/experiment.py": "maggy_b200.experiment.experiment",
    "tensorboard.py": "maggy_b200.tensorboard", "callbacks.py": "maggy_b200.callbacks", "trial.py": "maggy_b200.trial",
    "searchspace.py": "maggy_b200.searchspace", "util.py": "maggy_b200.util", "constants.py": "maggy_b200.constants",
    "config/lagom.py": "maggy_b200.config", "config/base_config.py": "maggy_b200.config",
    "config/hyperparameter_optimization.py": "maggy_b200.config", "config/ablation.py": "maggy_b200.config",
    "config/torch_distributed.py": "maggy_b200.config", "config/tf_distributed.py": "maggy_b200.config",
    "ablation/ablationstudy.py": "maggy_b200.ablation.ablationstudy",
    "ablation/ablator/abstractablator.py": "maggy_b200.ablation.ablator.abstractablator",
    "ablation/ablator/loco.py": "maggy_b200.ablation.ablator.loco",
    "optimizer/abstractoptimizer.py": "maggy_b200.optimizer.abstractoptimizer",
    "optimizer/asha.py": "maggy_b200.optimizer.asha", "optimizer/randomsearch.py": "maggy_b200.optimizer.randomsearch",
    "optimizer/gridsearch.py": "maggy_b200.optimizer.gridsearch", "optimizer/singlerun.py": "maggy_b200.optimizer.singlerun",
    "optimizer/bayes/base.py": "maggy_b200.optimizer.bayes.base", "optimizer/bayes/gp.py": "maggy_b200.optimizer.bayes.gp",
    "optimizer/bayes/tpe.py": "maggy_b200.optimizer.bayes.tpe",
    "optimizer/bayes/acquisitions.py": "maggy_b200.optimizer.bayes.acquisitions",
    "pruner/abstractpruner.py": "maggy_b200.pruner.abstractpruner", "pruner/hyperband.py": "maggy_b200.pruner.hyperband",
    "earlystop/medianrule.py": "maggy_b200.earlystop.medianrule", "earlystop/nostop.py": "maggy_b200.earlystop.nostop",
    "earlystop/abstractearlystop.py": "maggy_b200.earlystop.abstractearlystop",
}


def _public(name: str) -> bool:
    return not name.startswith("_") or name == "__init__"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not available")
def test_every_public_reference_name_resolves_here():
    missing = []
    for rel, mod in MAPPING.items():
        path = os.path.join(REF, rel)
        if not os.path.exists(path):
            continue
        tree = ast.parse(open(path).read())
        m = importlib.import_module(mod)
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and _public(node.name):
                if not hasattr(m, node.name):
                    missing.append(f"{rel}: {node.name}")
            elif isinstance(node, ast.ClassDef) and _public(node.name):
                cls = getattr(m, node.name, None)
                if cls is None:
                    missing.append(f"{rel}: class {node.name}")
                    continue
                for sub in node.body:
                    if isinstance(sub, ast.FunctionDef) and _public(sub.name) and not hasattr(cls, sub.name):
                        missing.append(f"{rel}: {node.name}.{sub.name}")
    assert not missing, "\n".join(missing)
