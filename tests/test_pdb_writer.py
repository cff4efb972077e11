"""PDB writer: byte parity with files written by the UNMODIFIED reference (tests/golden/make_golden_pdb.py), the
reference's file-naming rules, and the `analysis.utils` overlay.  Host-only: runs without a GPU."""
import os
import sys

import numpy as np
import pytest

from conftest import golden
from se3_diffusion_b200 import pdb_writer

CASES = ["single_backbone", "traj_mixed", "one_residue_f64"]


@pytest.mark.parametrize("name", CASES)
def test_bytes_equal_reference_writer(name):
    g = golden(f"pdb_{name}")
    got = pdb_writer.format_pdb(g["pos"], aatype=g.get("aatype"), b_factors=g.get("b_factors"))
    exp = g["expected"].tobytes()
    assert len(got) == len(exp), (len(got), len(exp))
    if got != exp:   # point at the first differing line
        gl, el = got.split(b"\n"), exp.split(b"\n")
        k = next(i for i, (a, b) in enumerate(zip(gl, el)) if a != b)
        raise AssertionError(f"line {k}:\n got {gl[k]!r}\n exp {el[k]!r}")


def test_structure_and_edge_cases():
    g = golden("pdb_traj_mixed")
    txt = pdb_writer.format_pdb(g["pos"], aatype=g["aatype"], b_factors=g["b_factors"]).decode()
    lines = txt.split("\n")
    assert lines[-1] == "END" and lines[0].rstrip() == "MODEL     1"
    assert sum(l.startswith("MODEL") for l in lines) == 3 and sum(l.startswith("ENDMDL") for l in lines) == 3
    assert all(len(l) >= 80 for l in lines[:-1])                       # padded, never truncated
    assert any(len(l) > 80 for l in lines)                             # the over-wide coordinates widen their line
    assert " UNK A" in txt and "  -0.000" in txt                       # aatype 20 -> UNK; negative values rounding to zero keep the sign
    with pytest.raises(ValueError):
        pdb_writer.format_pdb(g["pos"], aatype=np.full(9, 21))
    with pytest.raises(ValueError):
        pdb_writer.format_pdb(np.zeros((4, 3)))


def test_file_naming_rules(tmp_path):
    g = golden("pdb_single_backbone")
    d = str(tmp_path)
    p1 = pdb_writer.write_prot_to_pdb(g["pos"], os.path.join(d, "sample.pdb"))
    p2 = pdb_writer.write_prot_to_pdb(g["pos"], os.path.join(d, "sample.pdb"))
    assert os.path.basename(p1) == "sample_1.pdb" and os.path.basename(p2) == "sample_2.pdb"
    p3 = pdb_writer.write_prot_to_pdb(g["pos"], os.path.join(d, "sample.pdb"), overwrite=True)
    assert os.path.basename(p3) == "sample_1.pdb"
    p4 = pdb_writer.write_prot_to_pdb(g["pos"], os.path.join(d, "plain.pdb"), no_indexing=True)
    assert os.path.basename(p4) == "plain.pdb"
    assert open(p4, "rb").read() == g["expected"].tobytes()


def test_overlay_module_next_to_reference(tmp_path):
    """With the overlay ahead of the reference on sys.path, `from analysis import utils` keeps every reference symbol and swaps
    only the writer; both writers produce identical files for a random trajectory."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference tree not present on this machine")
    rh.install_stubs()
    import importlib.util
    import se3_diffusion_b200
    ov = os.path.join(os.path.dirname(se3_diffusion_b200.__file__), "overlay", "analysis", "utils.py")
    spec = importlib.util.spec_from_file_location("overlay_analysis_utils", ov)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from analysis import utils as ref_au
    assert mod.write_prot_to_pdb is pdb_writer.write_prot_to_pdb
    assert hasattr(mod, "create_full_prot") and hasattr(mod, "rigids_to_se3_vec")
    rng = np.random.RandomState(0)
    pos = np.zeros((4, 30, 37, 3), np.float32)
    pos[:, :, :5] = (rng.randn(4, 30, 5, 3) * 20).astype(np.float32)
    a = ref_au.write_prot_to_pdb(pos, str(tmp_path / "ref.pdb"), no_indexing=True)
    b = mod.write_prot_to_pdb(pos, str(tmp_path / "ours.pdb"), no_indexing=True)
    assert open(a, "rb").read() == open(b, "rb").read()
