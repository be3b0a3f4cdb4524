"""Median stopping rule (reference: maggy/earlystop/medianrule.py:21-60).

A running trial that has reported ``n`` metrics is stopped when its best metric so far is worse than
the median, over all finished trials with at least ``n`` reports, of their running average over the
first ``n`` reports. The same decision is taken natively by the control plane's policy engine
(csrc/ctrl/ring.cpp ``policy_on_metric``) and by the persistent scheduler kernel (K9); this class is
the Python oracle and the fallback for custom drivers.
"""
import statistics

from .abstractearlystop import AbstractEarlyStop


class MedianStoppingRule(AbstractEarlyStop):
    @staticmethod
    def earlystop_check(to_check, finalized_trials, direction):
        n = len(to_check.metric_history)
        if n == 0:
            return None
        averages = [sum(t.metric_history[:n]) / float(n) for t in finalized_trials if len(t.metric_history) >= n]
        if not averages:
            return None
        median = statistics.median(averages)
        if direction == "max" and max(to_check.metric_history) < median:
            return to_check.trial_id
        if direction == "min" and min(to_check.metric_history) > median:
            return to_check.trial_id
        return None
