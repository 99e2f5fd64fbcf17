"""crs_post.py -- TEST INFRASTRUCTURE: CPU restatement of the post-processing that ends the reference's matrix
assembly (gen_sparse_matrix, /root/reference/src/matrix.c:3829-3835), used to check nkp_crs_finalize_device:

  sum_dup_vals        src/matrix.c:3621-3650
  strip_matrix_zeros  src/matrix.c:3657-3688
  sort_cols_all_rows  src/matrix.c:3753-3765 (sort_cols_one_row :3730-3747, insertion sort)

Plain Python loops over the rows, in the reference's order of operations (the additions of sum_dup_vals are not
associative): meant for small cases.  Pinned by tests/test_oracle_and_plan.py::test_crs_post_reproduces_gen_A, which
feeds it the pre-processing form of the golden operand and requires the CRS the reference's unchanged gen_A wrote."""
import numpy as np


def sum_dup_vals(rowptr, colind, val):
    val = val.copy()
    dup_cnt = 0
    for r in range(len(rowptr) - 1):
        p1 = rowptr[r + 1]
        for p in range(rowptr[r], p1):
            for q in range(p + 1, p1):
                if colind[q] == colind[p]:
                    val[p] = val[p] + val[q]
                    val[q] = 0.0
                    dup_cnt += 1
    return val, dup_cnt


def strip_matrix_zeros(rowptr, colind, val):
    keep = val != 0.0
    new_rp = np.zeros(len(rowptr), dtype=rowptr.dtype)
    for r in range(len(rowptr) - 1):
        new_rp[r + 1] = new_rp[r] + int(keep[rowptr[r]:rowptr[r + 1]].sum())
    return new_rp, colind[keep].copy(), val[keep].copy()


def sort_cols_all_rows(rowptr, colind, val):
    colind, val = colind.copy(), val.copy()
    for r in range(len(rowptr) - 1):
        c = colind[rowptr[r]:rowptr[r + 1]]
        v = val[rowptr[r]:rowptr[r + 1]]
        for i in range(1, len(c)):
            key, kv = c[i], v[i]
            j = i - 1
            while j >= 0 and c[j] > key:
                c[j + 1], v[j + 1] = c[j], v[j]
                j -= 1
            c[j + 1], v[j + 1] = key, kv
    return colind, val


def finalize(rowptr, colind, val, strip_zeros=True):
    val, dup_cnt = sum_dup_vals(rowptr, colind, val)
    if strip_zeros:
        rowptr, colind, val = strip_matrix_zeros(rowptr, colind, val)
    colind, val = sort_cols_all_rows(rowptr, colind, val)
    return rowptr, colind, val, dup_cnt
