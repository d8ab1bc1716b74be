"""ctypes signatures of include/elynx_b200_model.h (host-side model/tree helpers)."""
SIGNATURES = {}
