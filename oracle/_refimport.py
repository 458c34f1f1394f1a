"""Import the UNMODIFIED reference modules from /root/reference (builder container only).

TEST INFRASTRUCTURE. Nothing under ``oracle/`` is part of the shipped product path.

The reference's ODE head is pure Python/PyTorch, but three of its imports are absent from
this image (``timm``, ``pyquaternion``, ``nuscenes``); none of them is reached by the hot
path (``DropPath`` is never active because ``drop_path=0``; ``warp_features`` is imported by
``streamingflow/layers/temporal_ode_bayes.py:8`` and never called).  We inject inert stubs
for exactly those three names and import the reference files where they lie.

``/root/reference`` does NOT exist on the GPU box: only ``oracle/gen_golden.py`` and the
``not gpu`` cross-checks that are skipped when the tree is absent may call this module.
"""
import os
import sys
import types
from types import SimpleNamespace as NS

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# The files of the reference that the ODE head needs, relative to the reference root.  ``stage_reference()``
# (called by ``__graft_entry__.build()`` in the builder container) copies them UNMODIFIED into the git-ignored
# ``baseline/_ref/`` so that ``bench.py --impl reference`` can time the reference itself on the GPU box's host cores,
# where /root/reference does not exist.  Nothing of it is ever part of the repository history or of the product.
REFERENCE_FILES = ("streamingflow/layers/temporal_ode_bayes.py", "streamingflow/layers/res_models.py",
                   "streamingflow/layers/convolutions.py", "streamingflow/layers/temporal.py",
                   "streamingflow/models/model_utils.py", "streamingflow/models/future_prediction_ode.py",
                   "streamingflow/models/decoder.py", "streamingflow/utils/geometry.py", "LICENSE")
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _has_tree(root) -> bool:
    return os.path.isdir(os.path.join(root, "streamingflow", "layers"))


def _pick_root() -> str:
    env = os.environ.get("SF_REFERENCE_ROOT")
    if env:
        return env
    return "/root/reference" if _has_tree("/root/reference") else STAGED_ROOT


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return _has_tree(REFERENCE_ROOT)


def stage_reference(src: str = "/root/reference") -> bool:
    """Copies REFERENCE_FILES from ``src`` to baseline/_ref/ (byte-identical).  Returns False when ``src`` is absent."""
    import shutil

    if not _has_tree(src):
        return False
    for rel in REFERENCE_FILES:
        a, b = os.path.join(src, rel), os.path.join(STAGED_ROOT, rel)
        if not os.path.exists(a):
            continue
        os.makedirs(os.path.dirname(b), exist_ok=True)
        shutil.copyfile(a, b)
    return True


def _install_stubs():
    import torch.nn as nn

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        timm_layers = types.ModuleType("timm.models.layers")

        class DropPath(nn.Identity):  # never active on this path (drop_path == 0)
            def __init__(self, *a, **k):
                super().__init__()

        timm_layers.DropPath = DropPath
        timm.models = timm_models
        timm_models.layers = timm_layers
        sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers})
    if "pyquaternion" not in sys.modules:
        pq = types.ModuleType("pyquaternion")
        pq.Quaternion = type("Quaternion", (), {})
        sys.modules["pyquaternion"] = pq
    if "nuscenes" not in sys.modules:
        ns = types.ModuleType("nuscenes")
        ns_utils = types.ModuleType("nuscenes.utils")
        ns_geo = types.ModuleType("nuscenes.utils.geometry_utils")
        ns_geo.transform_matrix = lambda *a, **k: None
        ns.utils = ns_utils
        ns_utils.geometry_utils = ns_geo
        sys.modules.update({"nuscenes": ns, "nuscenes.utils": ns_utils, "nuscenes.utils.geometry_utils": ns_geo})


def import_reference():
    """Returns a namespace with the reference classes on the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from streamingflow.models.future_prediction_ode import FuturePredictionODE
    from streamingflow.layers import temporal_ode_bayes as tob
    from streamingflow.layers import res_models, convolutions, temporal
    from streamingflow.models import model_utils

    return NS(FuturePredictionODE=FuturePredictionODE, tob=tob, res_models=res_models,
              convolutions=convolutions, temporal=temporal, model_utils=model_utils)


def make_cfg(channels=64, impute=True, solver="euler", variable=True, filter_size=None, skipco=False):
    """The slice of the fvcore cfg node the hot path reads."""
    return NS(MODEL=NS(IMPUTE=impute, SOLVER=solver,
                       SMALL_ENCODER=NS(FILTER_SIZE=filter_size or channels, SKIPCO=skipco),
                       ENCODER=NS(OUT_CHANNELS=channels),
                       FUTURE_PRED=NS(USE_VARIABLE_ODE_STEP=variable)))


class EpsTape:
    """Record / replay the standard-normal draws of ``Normal.rsample`` (model_utils.py:107-108).

    mode 'record': draws with torch's RNG and keeps every tensor; mode 'replay': returns the
    supplied tensors in order (cast to the requested dtype), so a fp64 run and a fp32 run of the
    reference see identical noise.
    """

    def __init__(self, replay=None):
        self.replay = list(replay) if replay is not None else None
        self.tape = []
        self._orig = None

    def __enter__(self):
        import torch
        import torch.distributions.normal as tdn

        self._orig = tdn._standard_normal

        def patched(shape, dtype, device):
            if self.replay is not None:
                e = self.replay[len(self.tape)].to(dtype=dtype, device=device).reshape(shape)
            else:
                e = self._orig(shape, dtype=dtype, device=device)
            self.tape.append(e.detach().clone())
            return e

        tdn._standard_normal = patched
        return self

    def __exit__(self, *exc):
        import torch.distributions.normal as tdn

        tdn._standard_normal = self._orig
        return False
