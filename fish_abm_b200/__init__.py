"""fish_abm_b200: B200 (sm_100a) implementation of the per-step agent update of fritzfrancisco/fish_abm,
behind the C ABI in include/fishabm.h.  See DESIGN.md."""
from .school import Ensemble, FishAbmError, School, averageturn, correct_coords, correctangle, create_environment  # noqa: F401
