"""Rutherford-Boeing reader / writer (spral_b200/rb.py; src/rutherford_boeing.f90)."""
import numpy as np

from spral_b200 import matrices as M
from spral_b200 import rb

SAMPLE = """\
Matrix of the SPRAL examples (examples/C/ssids.c), lower triangle       EX5X5
             3             1             1             3
rsa                        5             5             9             0
(6i3)           (9i2)           (3e24.16)
  1  3  6  8  9 10
 1 2 2 3 5 3 4 4 5
  2.0000000000000000E+00  1.0000000000000000E+00  4.0000000000000000E+00
  1.0000000000000000D+00  1.0000000000000000E+00  3.0000000000000000E+00
  2.0000000000000000E+00 -1.0000000000000000E+00  2.0000000000000000E+00
"""


def test_read_a_hand_written_file(tmp_path):
    p = tmp_path / "ex.rb"
    p.write_text(SAMPLE)
    info = rb.rb_peek(str(p))
    assert (info["type_code"], info["m"], info["n"], info["nnz"], info["id"]) == ("rsa", 5, 5, 9, "EX5X5")
    n, ptr, row, val, _ = rb.rb_read(str(p))
    n0, ptr0, row0, val0 = M.example_5x5()
    assert n == n0 and np.array_equal(ptr, ptr0) and np.array_equal(row, row0) and np.array_equal(val, val0)


def test_round_trip_and_upper_triangle_input(tmp_path):
    n, ptr, row, val = M.stencil_3d_27pt(6, shift=13.0)
    p = tmp_path / "st.rb"
    rb.rb_write(str(p), n, ptr, row, val, title="3-D 27-point 6^3 shifted", ident="ST27")
    n2, ptr2, row2, val2, info = rb.rb_read(str(p))
    assert info["type_code"] == "rsa" and n2 == n
    assert np.array_equal(ptr2, ptr) and np.array_equal(row2, row) and np.array_equal(val2, val)   # 17 digits
    # the same matrix stored by its UPPER triangle comes back as the lower one
    A = M.to_scipy(n, ptr, row, val)
    import scipy.sparse as sp
    U = sp.triu(A).tocsc()
    U.sort_indices()
    rb.rb_write(str(p), n, U.indptr + 1, U.indices + 1, U.data)
    n3, ptr3, row3, val3, _ = rb.rb_read(str(p))
    assert np.array_equal(ptr3, ptr) and np.array_equal(row3, row) and np.allclose(val3, val, rtol=0, atol=0)
    rb.rb_write(str(p), n, ptr, row)                                    # pattern only
    n4, ptr4, row4, val4, info4 = rb.rb_read(str(p))
    assert info4["type_code"] == "psa" and np.array_equal(row4, row) and np.all(val4 == 1.0)
