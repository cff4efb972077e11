"""Drop-in replacement of the reference's `data.se3_diffuser` (PEP-420 overlay, see overlay/model/score_network.py)."""
from se3_diffusion_b200.se3_diffuser import SE3Diffuser  # noqa: F401
