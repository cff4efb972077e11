"""Drop-in replacement of the reference's `model.score_network` (PEP-420 overlay: put se3_diffusion_b200/overlay FIRST on
PYTHONPATH, the reference second; every other `model.*` module still resolves to the reference)."""
from se3_diffusion_b200.score_network import ScoreNetwork  # noqa: F401
