"""Golden vectors for the CA eval metrics (analysis/metrics.py:120-132), produced by the UNMODIFIED reference functions in the build
container:  python tests/golden/make_golden_metrics.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

rh.install_stubs()
from analysis import metrics  # noqa: E402

if __name__ == "__main__":
    rs = np.random.RandomState(4)
    out = {}
    for k, (n, step, squash) in enumerate([(60, 3.8, 1.0), (37, 3.75, 0.15), (128, 3.9, 0.05)]):
        d = rs.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
        ca = (np.cumsum(d * (step + 0.05 * rs.standard_normal((n, 1))), axis=0) * squash).astype(np.float32)   # squashed walks produce clashes
        dev, valid = metrics.ca_ca_distance(ca)
        ncl, pcl = metrics.ca_ca_clashes(ca)
        out[f"ca_{k}"] = ca
        out[f"ref_{k}"] = np.array([dev, valid, ncl, pcl], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    print({k: v for k, v in out.items() if k.startswith("ref")})
