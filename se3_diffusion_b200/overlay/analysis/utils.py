"""Drop-in replacement of the reference's `analysis.utils` (PEP-420 overlay, see overlay/model/score_network.py): everything the
reference module defines is re-exported unchanged from the next `analysis/utils.py` on sys.path; only `write_prot_to_pdb` is
replaced by the native writer (byte-identical files)."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_ref = None
for _p in sys.path:
    _cand = os.path.join(_p or '.', 'analysis', 'utils.py')
    if os.path.isfile(_cand) and os.path.dirname(os.path.abspath(_cand)) != _here:
        _ref = _cand
        break
if _ref is None:
    raise ImportError("se3_diffusion_b200 overlay: the reference's analysis/utils.py is not on sys.path")
_spec = importlib.util.spec_from_file_location('analysis._reference_utils', _ref)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith('__')})

from se3_diffusion_b200.pdb_writer import write_prot_to_pdb  # noqa: E402,F401
