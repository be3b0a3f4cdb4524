"""Environment shims for running the UNMODIFIED reference (baseline/_ref) in an image without TensorFlow, PySpark,
scikit-optimize, statsmodels, DeepSpeed or fairscale. Nothing here touches the reference's code: `tf_standin` puts a
minimal `tensorflow` module (tensorboard's own stub plus the three attributes the reference touches at import time) into
`sys.modules`, `dep_placeholders` answers imports of the optional third-party packages with placeholder classes that
raise if anything ever instantiates them. With both in place `maggy.experiment.lagom(train_fn, BaseConfig())` runs
through the reference's python (non-Spark) driver."""


def install():
    from . import dep_placeholders, tf_standin  # noqa: F401
