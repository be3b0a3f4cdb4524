"""Exception types raised to / caught from user code (reference: maggy/core/exceptions.py:22-121)."""


class EarlyStopException(Exception):
    """Raised by ``reporter.broadcast`` inside the user's train_fn when the driver / policy engine has
    decided to stop the running trial. Carries the last broadcast metric, which becomes the trial's
    final metric."""

    def __init__(self, metric):
        super().__init__(f"trial stopped early at metric {metric}")
        self.metric = metric


class NotSupportedError(Exception):
    """A hyperparameter type or framework feature is not supported by a component."""

    def __init__(self, category, value, suggestion=""):
        self.message = f"({category}: {value}) is not supported by Maggy.{suggestion}"
        super().__init__(self.message)


class ReturnTypeError(TypeError):
    """train_fn returned a dict without the optimization key, or a non-numeric/non-dict value."""

    def __init__(self, optimization_key, return_type):
        self.message = (
            f"Training function cannot return value of type: {type(return_type).__name__}. "
            f"Return type can only be: float, int or dict containing the optimization key '{optimization_key}'."
        )
        super().__init__(self.message)


class MetricTypeError(TypeError):
    """The value under the optimization key is not numeric."""

    def __init__(self, optimization_key, return_type):
        self.message = (
            f"The optimization metric '{optimization_key}' must be numeric (float, int, np.number) "
            f"but the training function returned {type(return_type).__name__}."
        )
        super().__init__(self.message)


class BroadcastMetricTypeError(TypeError):
    def __init__(self, metric):
        self.message = (
            f"reporter.broadcast(metric, ...) expects a number or a 0-dim tensor, got {type(metric).__name__}."
        )
        super().__init__(self.message)


class BroadcastStepTypeError(TypeError):
    def __init__(self, value, step):
        self.message = f"reporter.broadcast(..., step) expects an integer step, got {type(step).__name__} ({step!r})."
        super().__init__(self.message)


class BroadcastStepValueError(ValueError):
    def __init__(self, value, step, prev_step):
        self.message = (
            f"reporter.broadcast steps must increase monotonically: got step {step} after step {prev_step}."
        )
        super().__init__(self.message)


class BadArgumentsError(Exception):
    """A callable received arguments it cannot work with."""

    def __init__(self, callable_name, suggestion=""):
        self.message = f"{callable_name} was called using bad arguments.{suggestion}"
        super().__init__(self.message)


class WorkerLostError(RuntimeError):
    """A worker process died or stopped heart-beating while it owned a trial."""
