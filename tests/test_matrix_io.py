"""CPU: `-problem file` -- the binary triplet reader / writer of the host side (host/amg_host.cpp) against what the
reference's own reader (ReadBinary_fread_HypreParCSR, src/Misc.cpp:800-915) returns for the same file: committed
fixture tests/golden/matrix_file.npz (made by tests/golden/make_golden.py from oracle/_ref) and, where oracle/_ref is
present, the live object code."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from async_multigrid_b200 import hierarchy as H
from oracle import oracle as O


@pytest.fixture()
def fixture_file(tmp_path):
    d = dict(np.load(os.path.join(GOLDEN, "matrix_file.npz")))
    path = tmp_path / "m.bin"
    path.write_bytes(d["file_bytes"].tobytes())
    return str(path), d


@pytest.mark.parametrize("flag", [0, 1])
def test_reader_matches_reference_fixture(fixture_file, flag):
    path, d = fixture_file
    m = H.read_matrix(path, symm_flag=flag)
    assert np.array_equal(m.indptr, d["symm%d_indptr" % flag])
    assert np.array_equal(m.indices, d["symm%d_indices" % flag])
    assert np.array_equal(m.data, d["symm%d_data" % flag])      # bit-exact: values are copied
    if flag:
        assert np.array_equal(m.indices[m.indptr[:-1]], np.arange(m.nrows))   # diag first


def test_reader_matches_live_reference(fixture_file):
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built here")
    path, _ = fixture_file
    for flag in (0, 1):
        ip, ix, dv = O.ref_read_matrix(path, flag)
        m = H.read_matrix(path, symm_flag=flag)
        assert np.array_equal(m.indptr, ip) and np.array_equal(m.indices, ix) and np.array_equal(m.data, dv)


def test_write_read_round_trip(tmp_path):
    A = H.laplacian("27pt", 6, 5, 4)
    p = str(tmp_path / "a.bin")
    H.write_matrix(A, p, lower_only=True)
    assert os.path.getsize(p) == 16 * (1 + (A.nnz + A.nrows) // 2)
    B = H.read_matrix(p, symm_flag=1)
    assert (A.to_scipy() != B.to_scipy()).nnz == 0
    H.write_matrix(A, p, lower_only=False)
    C_ = H.read_matrix(p, symm_flag=0)
    assert np.array_equal(C_.indptr, A.indptr) and np.array_equal(C_.indices, A.indices) and np.array_equal(C_.data, A.data)
    # a file read back feeds the setup like a generated matrix
    h = H.amg_setup(B)
    assert h.num_levels >= 2


def test_reader_errors(tmp_path):
    with pytest.raises(IOError):
        H.read_matrix(str(tmp_path / "missing.bin"))
    p = tmp_path / "bad.bin"
    p.write_bytes(b"\x00" * 17)
    with pytest.raises(IOError):
        H.read_matrix(str(p))
    rec = np.zeros(2, dtype=[("i", "<i4"), ("j", "<i4"), ("v", "<f8")])
    rec[0] = (3, 3, 0.0)
    rec[1] = (4, 1, 1.0)                                        # row outside 1..3
    p.write_bytes(rec.tobytes())
    with pytest.raises(IOError):
        H.read_matrix(str(p))
