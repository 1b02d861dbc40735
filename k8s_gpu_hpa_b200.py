"""Import alias: the package directory is ``k8s-gpu-hpa_b200/`` (not a Python identifier),
so this one-file shim exposes it as the package ``k8s_gpu_hpa_b200``."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "k8s-gpu-hpa_b200")]

from k8s_gpu_hpa_b200.capi import *  # noqa: E402,F401,F403
from k8s_gpu_hpa_b200 import capi, vector_add  # noqa: E402,F401
