"""The slice of the reference's fvcore config node that the ODE head reads (reference: streamingflow/config.py:95-153,
configs/Prediction_LC_ODE_Variable.yml).  fvcore is not required: any object with these attributes works."""
from types import SimpleNamespace as NS


def ode_cfg(channels=64, impute=True, solver="euler", variable_step=True, filter_size=None, skipco=False, precision="bf16"):
    """Defaults = the shipped Prediction_LC_ODE_Variable configuration (IMPUTE True, euler, variable ODE step)."""
    return NS(MODEL=NS(IMPUTE=impute, SOLVER=solver, ODE_PRECISION=precision,
                       SMALL_ENCODER=NS(FILTER_SIZE=filter_size or channels, SKIPCO=skipco),
                       ENCODER=NS(OUT_CHANNELS=channels),
                       FUTURE_PRED=NS(USE_VARIABLE_ODE_STEP=variable_step)))
