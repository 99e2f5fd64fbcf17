/*
 * netcdf.h -- minimal NetCDF-3 (classic CDF-1 / 64-bit-offset CDF-2) provider.
 *
 * The reference's file layer (src/file_io.c:1-368, src/grid.c:34-330,
 * src/matrix.c:264-460,3845-4031) is compiled UNCHANGED against this header.
 * It declares exactly the 17 entry points and the constants those files use
 * (enumerated by grep over /root/reference/src/*.c).
 * The implementation is nk_ocn_tracer_jacobian_precond_b200/csrc/nc3.c; it is
 * an independent reader/writer of the published classic file format, not a
 * copy of libnetcdf.  Numeric values of the constants follow the public
 * netcdf.h so that binaries agree with files written by the real library.
 */
#ifndef NKP_COMPAT_NETCDF_H
#define NKP_COMPAT_NETCDF_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int nc_type;

#define NC_NAT 0
#define NC_BYTE 1
#define NC_CHAR 2
#define NC_SHORT 3
#define NC_INT 4
#define NC_LONG NC_INT
#define NC_FLOAT 5
#define NC_DOUBLE 6

#define NC_NOWRITE 0x0000
#define NC_WRITE 0x0001
#define NC_CLOBBER 0x0000
#define NC_64BIT_OFFSET 0x0200

#define NC_GLOBAL (-1)
#define NC_MAX_NAME 256

#define NC_NOERR 0
#define NC_EBADID (-33)
#define NC_EEXIST (-35)
#define NC_EINVAL (-36)
#define NC_EPERM (-37)
#define NC_ENOTINDEFINE (-38)
#define NC_EINDEFINE (-39)
#define NC_ENAMEINUSE (-42)
#define NC_ENOTATT (-43)
#define NC_EBADTYPE (-45)
#define NC_EBADDIM (-46)
#define NC_EUNLIMPOS (-47)
#define NC_ENOTVAR (-49)
#define NC_ENOTNC (-51)
#define NC_ECHAR (-56)
#define NC_ERANGE (-60)
#define NC_ENOMEM (-61)
#define NC_EIO (-68)

const char *nc_strerror (int status);

int nc_create (const char *path, int cmode, int *ncidp);
int nc_open (const char *path, int omode, int *ncidp);
int nc_redef (int ncid);
int nc_enddef (int ncid);
int nc_close (int ncid);

int nc_def_dim (int ncid, const char *name, size_t len, int *dimidp);
int nc_inq_dimid (int ncid, const char *name, int *dimidp);
int nc_inq_dimlen (int ncid, int dimid, size_t *lenp);

int nc_def_var (int ncid, const char *name, nc_type xtype, int ndims, const int *dimids, int *varidp);
int nc_inq_varid (int ncid, const char *name, int *varidp);

int nc_put_att_text (int ncid, int varid, const char *name, size_t len, const char *tp);
int nc_put_att_int (int ncid, int varid, const char *name, nc_type xtype, size_t len, const int *ip);
int nc_put_att_double (int ncid, int varid, const char *name, nc_type xtype, size_t len, const double *dp);
int nc_get_att_double (int ncid, int varid, const char *name, double *dp);

int nc_get_var_int (int ncid, int varid, int *ip);
int nc_get_var_double (int ncid, int varid, double *dp);
int nc_put_var_int (int ncid, int varid, const int *ip);
int nc_put_var_double (int ncid, int varid, const double *dp);

#ifdef __cplusplus
}
#endif
#endif
