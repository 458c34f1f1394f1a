"""streamingflow_b200 -- B200-native GRU-ODE-Bayes BEV integration (drop-in for StreamingFlow's ODE head).

Public surface mirrors the reference's module paths:
    streamingflow_b200.models.future_prediction_ode.FuturePredictionODE
    streamingflow_b200.layers.temporal_ode_bayes.{NNFOwithBayesianJumps, DualGRUODECell, DualGRUCell, GRUObservationCell}
The CUDA library (libsf_b200.so, C ABI in include/sf_b200.h) is loaded lazily on first use and there is no fallback.
"""
__version__ = "0.1.0"
