"""Writes the PDB golden files with the UNMODIFIED reference writer (analysis/utils.py write_prot_to_pdb), run in the build
container:  python tests/golden/make_golden_pdb.py .  Inputs are stored next to the expected bytes."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402  (installs the import stubs and puts /root/reference on sys.path)

rh.install_stubs()
from analysis import utils as au  # noqa: E402


def cases():
    rng = np.random.RandomState(7)
    out = {}
    # 1: a backbone-like single frame, float32, default aatype / b_factors (what the sampler writes)
    p = np.zeros((12, 37, 3), np.float32)
    p[:, :5] = (rng.randn(12, 5, 3) * 14).astype(np.float32)
    out["single_backbone"] = dict(pos=p)
    # 2: trajectory, mixed residue types incl. UNK (20), b-factors, masked atoms, rounding ties, negative zero, wide fields
    p = (rng.randn(3, 9, 37, 3) * 30).astype(np.float32)
    p[:, :, 7:] = 0.0                                     # only 7 atom slots present
    p[0, 0, 0] = [0.0625, -0.0625, 0.1875]               # exact binary ties at the third decimal
    p[0, 1, 1] = [-0.0001, 0.0004, -0.0005]              # negative values rounding to zero
    p[1, 2, 2] = [12345.678, -1234.5678, 999.9996]       # wider than 8 columns / carry into a new digit
    p[2, 3, 3] = [0.0, 0.0, 0.0]                         # masked out (all-zero position)
    p[2, 4, 4] = [5e-8, 0.0, 0.0]                        # below the 1e-7 mask threshold
    aat = np.array([0, 1, 7, 20, 19, 4, 13, 2, 0])
    bf = rng.rand(9, 37) * 100
    out["traj_mixed"] = dict(pos=p, aatype=aat, b_factors=bf)
    # 3: float64 positions, one residue
    out["one_residue_f64"] = dict(pos=rng.randn(1, 37, 3) * 3)
    return out


def main():
    for name, c in cases().items():
        with tempfile.TemporaryDirectory() as d:
            path = au.write_prot_to_pdb(c["pos"], os.path.join(d, "x.pdb"), aatype=c.get("aatype"), b_factors=c.get("b_factors"), no_indexing=True)
            data = open(path, "rb").read()
        np.savez_compressed(os.path.join(HERE, f"pdb_{name}.npz"), expected=np.frombuffer(data, dtype=np.uint8),
                            **{k: v for k, v in c.items()})
        print(name, len(data), "bytes")


if __name__ == "__main__":
    main()
