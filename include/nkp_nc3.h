/*
 * nkp_nc3.h -- the one entry point of libnkp_nc3.so that is not part of the NetCDF C API
 * (include/compat/netcdf.h holds the NetCDF subset the reference uses, src/file_io.c).
 *
 * SURVEY.md 8(f) rank 2, matrix-file ingest: get_sparse_matrix (src/matrix.c:3944-4031) reads
 * nzval_row_wise through nc_get_var_double, i.e. a host loop that byte-swaps every big-endian
 * double of the CDF-2 file.  With the extent of the variable a caller reads the raw bytes itself
 * (read/pread into any buffer) and hands them to nkp_factor_be (include/nkprecond.h), which swaps
 * them on the GPU.
 */
#ifndef NKP_NC3_H
#define NKP_NC3_H

#ifdef __cplusplus
extern "C" {
#endif

/* File offset, byte count and external type (NC_INT, NC_DOUBLE, ...) of the data of a fixed-size
 * variable of an open file (data mode).  Returns NC_NOERR or a NetCDF error code. */
int nkp_nc3_inq_var_extent(int ncid, int varid, long long* offset, long long* nbytes, int* xtype);

#ifdef __cplusplus
}
#endif
#endif
