"""Offline replay of the control loop that OBSERVES the hot path.

The reference's manifests stay as they are; this module only re-states, in plain Python,
the two pieces of arithmetic they contain, so a run of the load generator
(``vectorAdd --duration S --target-util P --nvml``) can be checked against the HPA's
trigger without a cluster:

* the recording rule  ``cuda_test_gpu_avg``            cuda-test-prometheusrule.yaml:12-16
      avg( max by(node, pod, namespace)(dcgm_gpu_utilization)
           * on(pod) group_left(label_app)
           max by(pod, label_app)(kube_pod_labels{label_app="cuda-test"}) )
* the HPA decision    Object metric, targetValue 5, 1..3 replicas   cuda-test-hpa.yaml:11-21
      desired = ceil(current * metric / target), unchanged inside the 10 % tolerance band
      (the Kubernetes HPA algorithm; the controller itself is not part of the reference)

plus the two sampling stages between the GPU and the rule: dcgm-exporter collects every
10 000 ms (dcgm-exporter.yaml:37), Prometheus scrapes every 1 s
(kube-prometheus-stack-values.yaml:5).  Nothing here touches the GPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

HPA_TARGET = 5.0            # cuda-test-hpa.yaml:21   (README.md:113 says 4 %; the YAML is authoritative)
HPA_MIN, HPA_MAX = 1, 3     # cuda-test-hpa.yaml:11-12
DCGM_INTERVAL_S = 10.0      # dcgm-exporter.yaml:37   (-c 10000)
SCRAPE_INTERVAL_S = 1.0     # kube-prometheus-stack-values.yaml:5
HPA_SYNC_S = 15.0           # kube-controller-manager default --horizontal-pod-autoscaler-sync-period
HPA_TOLERANCE = 0.1         # kube-controller-manager default
HPA_DOWNSCALE_STABILIZATION_S = 300.0   # kube-controller-manager default --horizontal-pod-autoscaler-downscale-stabilization:
                                        # the controller scales down only to the HIGHEST recommendation of the last 5 minutes


@dataclass(frozen=True)
class Sample:
    """One `dcgm_gpu_utilization` series value: labels of the exporter + the value in percent."""
    node: str
    pod: str
    namespace: str
    value: float
    gpu: str = "0"


def cuda_test_gpu_avg(dcgm: list[Sample], pod_labels: dict[str, str], app: str = "cuda-test") -> float | None:
    """The recording rule. ``pod_labels`` maps pod -> label_app (kube_pod_labels, value 1).
    Returns None when the expression has no series (Prometheus records nothing)."""
    by_pod: dict[tuple[str, str, str], float] = {}
    for s in dcgm:                                   # max by(node, pod, namespace)
        k = (s.node, s.pod, s.namespace)
        by_pod[k] = max(by_pod.get(k, -math.inf), s.value)
    joined = [v * 1.0 for (node, pod, ns), v in by_pod.items()     # * on(pod) group_left(label_app) ...{label_app=app}
              if pod_labels.get(pod) == app]
    if not joined:
        return None
    return sum(joined) / len(joined)                 # avg(...)


def hpa_desired_replicas(current: int, metric: float | None, target: float = HPA_TARGET,
                         lo: int = HPA_MIN, hi: int = HPA_MAX, tolerance: float = HPA_TOLERANCE) -> int:
    """Object-metric HPA step: ratio = metric/target; within tolerance -> keep; else ceil(current*ratio), clamped."""
    if metric is None or current <= 0:
        return max(lo, min(hi, current if current > 0 else lo))
    ratio = metric / target
    if abs(ratio - 1.0) <= tolerance:
        return max(lo, min(hi, current))
    return max(lo, min(hi, math.ceil(current * ratio)))


@dataclass
class Replay:
    """Feeds per-pod utilisation traces through exporter sampling, scrape, rule and HPA."""
    dcgm_interval_s: float = DCGM_INTERVAL_S
    scrape_interval_s: float = SCRAPE_INTERVAL_S
    hpa_sync_s: float = HPA_SYNC_S
    target: float = HPA_TARGET
    downscale_stabilization_s: float = HPA_DOWNSCALE_STABILIZATION_S
    events: list[tuple[float, float | None, int]] = field(default_factory=list)

    def run(self, traces: dict[str, list[tuple[float, float]]], duration_s: float, replicas: int = 1) -> list[tuple[float, float | None, int]]:
        """``traces[pod]`` = [(t_seconds, util_percent), ...] as sampled on the GPU (e.g. the
        CLI's NVML samples).  Pods beyond ``replicas`` are ignored until the HPA adds them.
        Returns [(t, cuda_test_gpu_avg, replicas)] at every HPA sync."""
        pods = sorted(traces)
        exported: dict[str, float] = {}              # what the exporter currently serves
        self.events = []
        t, next_dcgm, next_hpa, metric = 0.0, 0.0, self.hpa_sync_s, None
        recommendations: list[tuple[float, int]] = []        # (time, desired replicas) of past syncs
        while t <= duration_s + 1e-9:
            if t + 1e-9 >= next_dcgm:                # exporter refreshes its gauges
                for p in pods[:replicas]:
                    past = [u for (ts, u) in traces[p] if ts <= t]
                    if past:
                        exported[p] = past[-1]
                next_dcgm += self.dcgm_interval_s
            live = [Sample("node0", p, "default", exported[p]) for p in pods[:replicas] if p in exported]
            metric = cuda_test_gpu_avg(live, {p: "cuda-test" for p in pods})   # rule evaluated on each scrape
            if t + 1e-9 >= next_hpa:
                desired = hpa_desired_replicas(replicas, metric, self.target)
                recommendations.append((t, desired))
                # scale-ups apply at once; a scale-down is held to the highest recommendation
                # inside the stabilization window (the controller's default behaviour; the reference only
                # says "if the usage drops low enough, a scaledown will occur", README.md:121)
                recent = [d for (ts, d) in recommendations if ts >= t - self.downscale_stabilization_s - 1e-9]
                replicas = min(replicas, max(recent)) if desired < replicas else desired
                replicas = min(replicas, len(pods))
                self.events.append((t, metric, replicas))
                next_hpa += self.hpa_sync_s
            t += self.scrape_interval_s
        return self.events


def would_scale_up(mean_util_percent: float, replicas: int = 1) -> bool:
    """Does a steady utilisation reading move the HPA off ``replicas``?"""
    return hpa_desired_replicas(replicas, mean_util_percent) > replicas
