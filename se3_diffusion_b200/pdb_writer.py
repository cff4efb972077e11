"""PDB writer of sampled backbones — native mirror of the reference's `analysis.utils.write_prot_to_pdb`
(/root/reference/analysis/utils.py:39-77, which formats through data/protein.py:146-219 `to_pdb` in Python loops).

Same signature, same file-naming rules, byte-identical files (tests/test_pdb_writer.py pins this against files written by the
unmodified reference); the text is produced by `fd_format_pdb` in the C-ABI library (host code, no GPU needed), which is what makes
500-frame trajectories cheap to write (the on-disk format immediately downstream of sampling)."""
import ctypes as C
import os
import re

import numpy as np

from ._lib import check, load


def _format(prot_pos, aatype=None, b_factors=None):
    """-> (ctypes buffer, length): the file content write_prot_to_pdb would write for prot_pos [N,37,3] or [T,N,37,3]."""
    prot_pos = np.asarray(prot_pos)
    if prot_pos.ndim == 3:
        pos = prot_pos[None]
    elif prot_pos.ndim == 4:
        pos = prot_pos
    else:
        raise ValueError(f'Invalid positions shape {prot_pos.shape}')
    assert pos.shape[-1] == 3 and pos.shape[-2] == 37
    T, N = pos.shape[:2]
    # float32 stays float32 (the atom mask `sum(|pos|) > 1e-7` is evaluated in the array's own precision, analysis/utils.py:62,68;
    # formatting sees the exact values either way)
    is_f32 = pos.dtype == np.float32
    pos = np.ascontiguousarray(pos, dtype=np.float32 if is_f32 else np.float64)
    if aatype is not None and np.any(np.asarray(aatype) > 20):
        raise ValueError('Invalid aatypes.')
    aa = None if aatype is None else np.ascontiguousarray(np.asarray(aatype).astype(np.int64), dtype=np.int32)
    bf = None if b_factors is None else np.ascontiguousarray(b_factors, dtype=np.float64)
    if aa is not None:
        assert aa.shape == (N,)
    if bf is not None:
        assert bf.shape == (N, 37)
    lib = load()
    n = C.c_size_t(0)
    ptr = lambda a, ty: None if a is None else a.ctypes.data_as(ty)
    args = (pos.ctypes.data_as(C.c_void_p), int(is_f32), None, ptr(aa, C.POINTER(C.c_int)), ptr(bf, C.POINTER(C.c_double)), T, N)
    cap = (3 * T + 5 * T * N) * 81 + 4096          # exact for backbone-only frames (N, CA, C, CB, O), the sampler's output
    buf = (C.c_char * cap)()
    rc = lib.fd_format_pdb(*args, C.cast(buf, C.c_void_p), cap, C.byref(n))
    if rc != 0 and n.value > cap:                  # more atoms / over-wide fields: the call reported the length it needs
        cap = n.value
        buf = (C.c_char * cap)()
        rc = lib.fd_format_pdb(*args, C.cast(buf, C.c_void_p), cap, C.byref(n))
    check(rc)
    return buf, n.value


def format_pdb(prot_pos, aatype=None, b_factors=None) -> bytes:
    """The file content write_prot_to_pdb would write for prot_pos [N,37,3] or [T,N,37,3]."""
    buf, n = _format(prot_pos, aatype, b_factors)
    return bytes(memoryview(buf)[:n])


def write_prot_to_pdb(prot_pos: np.ndarray, file_path: str, aatype: np.ndarray = None, overwrite=False, no_indexing=False, b_factors=None):
    """Drop-in for analysis.utils.write_prot_to_pdb (same arguments, same returned path, same bytes on disk)."""
    if overwrite:
        max_existing_idx = 0
    else:
        file_dir = os.path.dirname(file_path)
        # the reference derives the stem with str.strip('.pdb'), i.e. it strips the CHARACTERS '.', 'p', 'd', 'b' from both ends
        file_name = os.path.basename(file_path).strip('.pdb')
        idx = [0]
        for x in os.listdir(file_dir):
            if file_name in x:
                m = re.findall(r'_(\d+).pdb', x)
                if m:
                    idx.append(int(m[0]))
        max_existing_idx = max(idx)
    save_path = file_path if no_indexing else file_path.replace('.pdb', '') + f'_{max_existing_idx + 1}.pdb'
    buf, n = _format(prot_pos, aatype, b_factors)
    with open(save_path, 'wb') as f:
        f.write(memoryview(buf)[:n])
    return save_path
