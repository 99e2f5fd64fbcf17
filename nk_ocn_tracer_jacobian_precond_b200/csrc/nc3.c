/*
 * nc3.c -- NetCDF-3 classic (CDF-1) / 64-bit-offset (CDF-2) reader-writer.
 *
 * Provides the 17 nc_* entry points the reference file layer calls
 * (src/file_io.c, src/grid.c:235-293, src/matrix.c:283-329,3860-3931) so that
 * those sources compile and run unchanged without libnetcdf.  Written from the
 * published classic-format grammar:
 *
 *   header  = magic numrecs dim_list gatt_list var_list
 *   magic   = 'C' 'D' 'F' (1|2)
 *   *_list  = ABSENT(0,0) | tag nelems [elem ...]     tags: dim 0x0A, var 0x0B, att 0x0C
 *   name    = nelems chars pad4
 *   var     = name ndims dimid* vatt_list nc_type vsize begin(32|64 bit)
 *
 * All header integers and all data are big-endian.  Fixed-size variables are
 * laid out back to back in definition order, each padded to 4 bytes; record
 * variables follow.  nc_redef()+nc_close() on a file that already holds data
 * (src/matrix.c:288, :3865) grows the header and relocates the data section.
 *
 * Limits (sufficient for the reference's use): whole-variable get/put only;
 * record variables are readable and writable when numrecs <= 1 (one contiguous slab each);
 * files with interleaved records are refused by get/put and by nc_redef's relocation; no CDF-5.
 */
#define _FILE_OFFSET_BITS 64
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <sys/types.h>
#include <unistd.h>

#include "netcdf.h"

#define TAG_DIM 0x0A
#define TAG_VAR 0x0B
#define TAG_ATT 0x0C

typedef struct {
   char *name;
   nc_type type;
   size_t nelems;
   unsigned char *data;         /* big-endian external representation, unpadded */
} nc3_att;

typedef struct {
   char *name;
   size_t len;                  /* 0 = record dimension */
} nc3_dim;

typedef struct {
   char *name;
   int ndims;
   int *dimids;
   int natts;
   nc3_att *atts;
   nc_type type;
   uint64_t vsize;
   uint64_t begin;
   uint64_t old_begin;          /* location before a redef, (uint64_t)-1 if new */
   int is_rec;
} nc3_var;

typedef struct {
   int used;
   FILE *fp;
   int writable;
   int version;                 /* 1 or 2 */
   int define_mode;
   int header_dirty;
   uint32_t numrecs;
   int ndims;
   nc3_dim *dims;
   int ngatts;
   nc3_att *gatts;
   int nvars;
   nc3_var *vars;
} nc3_file;

#define NC3_MAX_FILES 64
static nc3_file files[NC3_MAX_FILES];

/* ---- helpers ----------------------------------------------------------- */

static size_t
type_size (nc_type t)
{
   switch (t) {
   case NC_BYTE:
   case NC_CHAR:
      return 1;
   case NC_SHORT:
      return 2;
   case NC_INT:
   case NC_FLOAT:
      return 4;
   case NC_DOUBLE:
      return 8;
   default:
      return 0;
   }
}

static uint64_t
pad4 (uint64_t n)
{
   return (n + 3u) & ~(uint64_t) 3u;
}

static nc3_file *
get_file (int ncid)
{
   if (ncid < 0 || ncid >= NC3_MAX_FILES || !files[ncid].used)
      return NULL;
   return &files[ncid];
}

static void
free_atts (int n, nc3_att * a)
{
   int i;
   for (i = 0; i < n; i++) {
      free (a[i].name);
      free (a[i].data);
   }
   free (a);
}

static void
free_file (nc3_file * f)
{
   int i;
   for (i = 0; i < f->ndims; i++)
      free (f->dims[i].name);
   free (f->dims);
   free_atts (f->ngatts, f->gatts);
   for (i = 0; i < f->nvars; i++) {
      free (f->vars[i].name);
      free (f->vars[i].dimids);
      free_atts (f->vars[i].natts, f->vars[i].atts);
   }
   free (f->vars);
   memset (f, 0, sizeof (*f));
}

/* growable big-endian byte buffer for header serialisation */
typedef struct {
   unsigned char *p;
   size_t len, cap;
} bbuf;

static int
bb_reserve (bbuf * b, size_t extra)
{
   if (b->len + extra > b->cap) {
      size_t ncap = b->cap ? b->cap * 2 : 1024;
      unsigned char *np;
      while (ncap < b->len + extra)
         ncap *= 2;
      if ((np = realloc (b->p, ncap)) == NULL)
         return -1;
      b->p = np;
      b->cap = ncap;
   }
   return 0;
}

static int
bb_u32 (bbuf * b, uint32_t v)
{
   if (bb_reserve (b, 4))
      return -1;
   b->p[b->len++] = (unsigned char) (v >> 24);
   b->p[b->len++] = (unsigned char) (v >> 16);
   b->p[b->len++] = (unsigned char) (v >> 8);
   b->p[b->len++] = (unsigned char) v;
   return 0;
}

static int
bb_u64 (bbuf * b, uint64_t v)
{
   return bb_u32 (b, (uint32_t) (v >> 32)) || bb_u32 (b, (uint32_t) v);
}

static int
bb_bytes_padded (bbuf * b, const void *src, size_t n)
{
   size_t np = (size_t) pad4 (n);
   if (bb_reserve (b, np))
      return -1;
   memcpy (b->p + b->len, src, n);
   memset (b->p + b->len + n, 0, np - n);
   b->len += np;
   return 0;
}

static int
bb_name (bbuf * b, const char *s)
{
   size_t n = strlen (s);
   return bb_u32 (b, (uint32_t) n) || bb_bytes_padded (b, s, n);
}

static int
bb_atts (bbuf * b, int n, const nc3_att * a)
{
   int i;
   if (n == 0)
      return bb_u32 (b, 0) || bb_u32 (b, 0);
   if (bb_u32 (b, TAG_ATT) || bb_u32 (b, (uint32_t) n))
      return -1;
   for (i = 0; i < n; i++) {
      if (bb_name (b, a[i].name) || bb_u32 (b, (uint32_t) a[i].type) || bb_u32 (b, (uint32_t) a[i].nelems)
          || bb_bytes_padded (b, a[i].data, a[i].nelems * type_size (a[i].type)))
         return -1;
   }
   return 0;
}

/* serialise the header; var begins must already be set */
static int
build_header (const nc3_file * f, bbuf * b)
{
   int i, d;
   unsigned char magic[4] = { 'C', 'D', 'F', 0 };
   magic[3] = (unsigned char) f->version;
   b->len = 0;
   if (bb_reserve (b, 4))
      return -1;
   memcpy (b->p, magic, 4);
   b->len = 4;
   if (bb_u32 (b, f->numrecs))
      return -1;
   if (f->ndims == 0) {
      if (bb_u32 (b, 0) || bb_u32 (b, 0))
         return -1;
   } else {
      if (bb_u32 (b, TAG_DIM) || bb_u32 (b, (uint32_t) f->ndims))
         return -1;
      for (i = 0; i < f->ndims; i++)
         if (bb_name (b, f->dims[i].name) || bb_u32 (b, (uint32_t) f->dims[i].len))
            return -1;
   }
   if (bb_atts (b, f->ngatts, f->gatts))
      return -1;
   if (f->nvars == 0) {
      if (bb_u32 (b, 0) || bb_u32 (b, 0))
         return -1;
   } else {
      if (bb_u32 (b, TAG_VAR) || bb_u32 (b, (uint32_t) f->nvars))
         return -1;
      for (i = 0; i < f->nvars; i++) {
         const nc3_var *v = &f->vars[i];
         if (bb_name (b, v->name) || bb_u32 (b, (uint32_t) v->ndims))
            return -1;
         for (d = 0; d < v->ndims; d++)
            if (bb_u32 (b, (uint32_t) v->dimids[d]))
               return -1;
         if (bb_atts (b, v->natts, v->atts) || bb_u32 (b, (uint32_t) v->type))
            return -1;
         /* vsize saturates at 2^32-4 for huge variables, as the format note prescribes */
         if (bb_u32 (b, v->vsize > 0xFFFFFFFCu ? 0xFFFFFFFFu : (uint32_t) v->vsize))
            return -1;
         if (f->version == 1) {
            if (bb_u32 (b, (uint32_t) v->begin))
               return -1;
         } else if (bb_u64 (b, v->begin))
            return -1;
      }
   }
   return 0;
}

static uint64_t
var_nelems (const nc3_file * f, const nc3_var * v)
{
   uint64_t n = 1;
   int d;
   for (d = 0; d < v->ndims; d++) {
      size_t len = f->dims[v->dimids[d]].len;
      if (len == 0)
         len = f->numrecs;      /* record dimension */
      n *= len;
   }
   return n;
}

static uint64_t
var_fixed_bytes (const nc3_file * f, const nc3_var * v)
{
   uint64_t n = 1;
   int d;
   for (d = 0; d < v->ndims; d++) {
      size_t len = f->dims[v->dimids[d]].len;
      if (len != 0)
         n *= len;
   }
   return pad4 (n * type_size (v->type));
}

/* ---- header parsing ------------------------------------------------------ */

typedef struct {
   FILE *fp;
   int err;
} rd;

static uint32_t
rd_u32 (rd * r)
{
   unsigned char c[4];
   if (fread (c, 1, 4, r->fp) != 4) {
      r->err = 1;
      return 0;
   }
   return ((uint32_t) c[0] << 24) | ((uint32_t) c[1] << 16) | ((uint32_t) c[2] << 8) | c[3];
}

static uint64_t
rd_u64 (rd * r)
{
   uint64_t hi = rd_u32 (r);
   return (hi << 32) | rd_u32 (r);
}

static char *
rd_name (rd * r)
{
   uint32_t n = rd_u32 (r);
   size_t np = (size_t) pad4 (n);
   char *s;
   if (r->err || n > (1u << 20))
      return NULL;
   if ((s = malloc (np + 1)) == NULL) {
      r->err = 1;
      return NULL;
   }
   if (np && fread (s, 1, np, r->fp) != np) {
      r->err = 1;
      free (s);
      return NULL;
   }
   s[n] = '\0';
   return s;
}

static int
rd_atts (rd * r, int *np, nc3_att ** ap)
{
   uint32_t tag = rd_u32 (r);
   uint32_t n = rd_u32 (r);
   uint32_t i;
   nc3_att *a;
   *np = 0;
   *ap = NULL;
   if (r->err)
      return -1;
   if (tag == 0 && n == 0)
      return 0;
   if (tag != TAG_ATT)
      return -1;
   if ((a = calloc (n ? n : 1, sizeof (nc3_att))) == NULL)
      return -1;
   *ap = a;
   *np = (int) n;
   for (i = 0; i < n; i++) {
      size_t nb, nbp;
      if ((a[i].name = rd_name (r)) == NULL)
         return -1;
      a[i].type = (nc_type) rd_u32 (r);
      a[i].nelems = rd_u32 (r);
      if (r->err || type_size (a[i].type) == 0)
         return -1;
      nb = a[i].nelems * type_size (a[i].type);
      nbp = (size_t) pad4 (nb);
      if ((a[i].data = malloc (nbp ? nbp : 1)) == NULL)
         return -1;
      if (nbp && fread (a[i].data, 1, nbp, r->fp) != nbp)
         return -1;
   }
   return 0;
}

static int
parse_header (nc3_file * f)
{
   rd r;
   unsigned char magic[4];
   uint32_t tag, n, i;
   int d;

   r.fp = f->fp;
   r.err = 0;
   if (fseeko (f->fp, 0, SEEK_SET) || fread (magic, 1, 4, f->fp) != 4)
      return NC_ENOTNC;
   if (magic[0] != 'C' || magic[1] != 'D' || magic[2] != 'F' || (magic[3] != 1 && magic[3] != 2))
      return NC_ENOTNC;
   f->version = magic[3];
   f->numrecs = rd_u32 (&r);
   if (f->numrecs == 0xFFFFFFFFu)
      f->numrecs = 0;           /* STREAMING marker */

   tag = rd_u32 (&r);
   n = rd_u32 (&r);
   if (r.err || (tag != 0 && tag != TAG_DIM))
      return NC_ENOTNC;
   if (tag == TAG_DIM) {
      if ((f->dims = calloc (n ? n : 1, sizeof (nc3_dim))) == NULL)
         return NC_ENOMEM;
      f->ndims = (int) n;
      for (i = 0; i < n; i++) {
         if ((f->dims[i].name = rd_name (&r)) == NULL)
            return NC_ENOTNC;
         f->dims[i].len = rd_u32 (&r);
      }
   }
   if (rd_atts (&r, &f->ngatts, &f->gatts))
      return NC_ENOTNC;

   tag = rd_u32 (&r);
   n = rd_u32 (&r);
   if (r.err || (tag != 0 && tag != TAG_VAR))
      return NC_ENOTNC;
   if (tag == TAG_VAR) {
      if ((f->vars = calloc (n ? n : 1, sizeof (nc3_var))) == NULL)
         return NC_ENOMEM;
      f->nvars = (int) n;
      for (i = 0; i < n; i++) {
         nc3_var *v = &f->vars[i];
         if ((v->name = rd_name (&r)) == NULL)
            return NC_ENOTNC;
         v->ndims = (int) rd_u32 (&r);
         if (r.err || v->ndims < 0 || v->ndims > 1024)
            return NC_ENOTNC;
         if ((v->dimids = calloc ((size_t) (v->ndims ? v->ndims : 1), sizeof (int))) == NULL)
            return NC_ENOMEM;
         for (d = 0; d < v->ndims; d++) {
            v->dimids[d] = (int) rd_u32 (&r);
            if (v->dimids[d] < 0 || v->dimids[d] >= f->ndims)
               return NC_ENOTNC;
         }
         if (rd_atts (&r, &v->natts, &v->atts))
            return NC_ENOTNC;
         v->type = (nc_type) rd_u32 (&r);
         v->vsize = rd_u32 (&r);
         v->begin = (f->version == 1) ? rd_u32 (&r) : rd_u64 (&r);
         v->old_begin = v->begin;
         v->is_rec = (v->ndims > 0 && f->dims[v->dimids[0]].len == 0);
         if (r.err || type_size (v->type) == 0)
            return NC_ENOTNC;
         /* recompute the true size (vsize saturates for >4 GiB variables) */
         v->vsize = var_fixed_bytes (f, v);
      }
   }
   return r.err ? NC_ENOTNC : NC_NOERR;
}

/* ---- layout + header write (the implicit enddef) -------------------------- */

static int
move_bytes (FILE * fp, uint64_t from, uint64_t to, uint64_t n)
{
   /* regions may overlap with to > from: copy chunks from the tail backwards */
   const size_t chunk = 1u << 22;
   unsigned char *buf;
   if (from == to || n == 0)
      return 0;
   if ((buf = malloc (chunk)) == NULL)
      return -1;
   if (to > from) {
      uint64_t left = n;
      while (left > 0) {
         size_t c = left > chunk ? chunk : (size_t) left;
         left -= c;
         if (fseeko (fp, (off_t) (from + left), SEEK_SET) || fread (buf, 1, c, fp) != c
             || fseeko (fp, (off_t) (to + left), SEEK_SET) || fwrite (buf, 1, c, fp) != c) {
            free (buf);
            return -1;
         }
      }
   } else {
      uint64_t done = 0;
      while (done < n) {
         size_t c = (n - done) > chunk ? chunk : (size_t) (n - done);
         if (fseeko (fp, (off_t) (from + done), SEEK_SET) || fread (buf, 1, c, fp) != c
             || fseeko (fp, (off_t) (to + done), SEEK_SET) || fwrite (buf, 1, c, fp) != c) {
            free (buf);
            return -1;
         }
         done += c;
      }
   }
   free (buf);
   return 0;
}

static int
do_enddef (nc3_file * f)
{
   bbuf b = { NULL, 0, 0 };
   uint64_t off, end;
   int i, k, nmove = 0;
   int *order;

   /* The layout below places ONE record slab per record variable, which is the whole record section
    * only while numrecs <= 1.  A file with interleaved records is refused before anything is touched
    * (the reference never redefines such a file: src/matrix.c:288, :3865 act on its own matrix file). */
   if (f->numrecs > 1)
      for (i = 0; i < f->nvars; i++)
         if (f->vars[i].is_rec)
            return NC_EINVAL;

   /* header length does not depend on the begin values, only on the version */
   if (build_header (f, &b)) {
      free (b.p);
      return NC_ENOMEM;
   }
   off = pad4 (b.len);
   for (i = 0; i < f->nvars; i++) {
      if (f->vars[i].is_rec)
         continue;
      f->vars[i].vsize = var_fixed_bytes (f, &f->vars[i]);
      f->vars[i].begin = off;
      off += f->vars[i].vsize;
   }
   for (i = 0; i < f->nvars; i++) {
      if (!f->vars[i].is_rec)
         continue;
      f->vars[i].vsize = var_fixed_bytes (f, &f->vars[i]);
      f->vars[i].begin = off;
      off += f->vars[i].vsize;  /* one record slab each; valid for numrecs <= 1 */
   }
   end = off;
   if (f->version == 1 && end > 0x7FFFFFFFu) {
      free (b.p);
      return NC_EINVAL;
   }

   /* grow the file first (never shrink it), then relocate the existing variables in DESCENDING order
    * of their old position: data only ever moves towards the end of the file, so a variable moved
    * later (lower old offset) cannot be overwritten by one moved earlier */
   fflush (f->fp);
   {
      off_t cur;
      fseeko (f->fp, 0, SEEK_END);
      cur = ftello (f->fp);
      if (cur < 0 || ((uint64_t) cur < end && ftruncate (fileno (f->fp), (off_t) end))) {
         free (b.p);
         return NC_EIO;
      }
   }
   if ((order = malloc (sizeof (int) * (size_t) (f->nvars ? f->nvars : 1))) == NULL) {
      free (b.p);
      return NC_ENOMEM;
   }
   for (i = 0; i < f->nvars; i++)
      if (f->vars[i].old_begin != (uint64_t) - 1 && f->vars[i].old_begin != f->vars[i].begin) {
         int v = i;
         for (k = nmove++; k > 0 && f->vars[order[k - 1]].old_begin < f->vars[v].old_begin; k--)
            order[k] = order[k - 1];
         order[k] = v;
      }
   for (k = 0; k < nmove; k++) {
      nc3_var *v = &f->vars[order[k]];
      if (move_bytes (f->fp, v->old_begin, v->begin, v->vsize)) {
         free (order);
         free (b.p);
         return NC_EIO;
      }
   }
   free (order);
   for (i = 0; i < f->nvars; i++)
      f->vars[i].old_begin = f->vars[i].begin;

   if (build_header (f, &b)) {
      free (b.p);
      return NC_ENOMEM;
   }
   if (fseeko (f->fp, 0, SEEK_SET) || fwrite (b.p, 1, b.len, f->fp) != b.len) {
      free (b.p);
      return NC_EIO;
   }
   free (b.p);
   fflush (f->fp);
   f->define_mode = 0;
   f->header_dirty = 0;
   return NC_NOERR;
}

/* ---- public API ------------------------------------------------------------ */

const char *
nc_strerror (int status)
{
   switch (status) {
   case NC_NOERR:
      return "No error";
   case NC_EBADID:
      return "NetCDF: Not a valid ID";
   case NC_EEXIST:
      return "NetCDF: File exists && NC_NOCLOBBER";
   case NC_EINVAL:
      return "NetCDF: Invalid argument";
   case NC_EPERM:
      return "NetCDF: Write to read only";
   case NC_ENOTINDEFINE:
      return "NetCDF: Operation not allowed in data mode";
   case NC_EINDEFINE:
      return "NetCDF: Operation not allowed in define mode";
   case NC_ENAMEINUSE:
      return "NetCDF: String match to name in use";
   case NC_ENOTATT:
      return "NetCDF: Attribute not found";
   case NC_EBADTYPE:
      return "NetCDF: Not a valid data type or _FillValue type mismatch";
   case NC_EBADDIM:
      return "NetCDF: Invalid dimension ID or name";
   case NC_EUNLIMPOS:
      return "NetCDF: NC_UNLIMITED in the wrong index";
   case NC_ENOTVAR:
      return "NetCDF: Variable not found";
   case NC_ENOTNC:
      return "NetCDF: Unknown file format";
   case NC_ECHAR:
      return "NetCDF: Attempt to convert between text & numbers";
   case NC_ERANGE:
      return "NetCDF: Numeric conversion not representable";
   case NC_ENOMEM:
      return "NetCDF: Memory allocation (malloc) failure";
   case NC_EIO:
      return "NetCDF: I/O failure";
   case 2:
      return "No such file or directory";
   default:
      return "NetCDF: Unknown error";
   }
}

static int
alloc_slot (void)
{
   int i;
   for (i = 0; i < NC3_MAX_FILES; i++)
      if (!files[i].used)
         return i;
   return -1;
}

int
nc_create (const char *path, int cmode, int *ncidp)
{
   int id = alloc_slot ();
   nc3_file *f;
   if (id < 0)
      return NC_ENOMEM;
   if (path == NULL)
      return NC_EINVAL;
   f = &files[id];
   memset (f, 0, sizeof (*f));
   if ((f->fp = fopen (path, "w+b")) == NULL)
      return NC_EIO;
   f->used = 1;
   f->writable = 1;
   f->version = (cmode & NC_64BIT_OFFSET) ? 2 : 1;
   f->define_mode = 1;
   f->header_dirty = 1;
   *ncidp = id;
   return NC_NOERR;
}

int
nc_open (const char *path, int omode, int *ncidp)
{
   int id = alloc_slot ();
   int status;
   nc3_file *f;
   if (id < 0)
      return NC_ENOMEM;
   if (path == NULL)
      return NC_EINVAL;
   f = &files[id];
   memset (f, 0, sizeof (*f));
   if ((f->fp = fopen (path, (omode & NC_WRITE) ? "r+b" : "rb")) == NULL)
      return 2;                 /* ENOENT, as libnetcdf passes errno through */
   f->used = 1;
   f->writable = (omode & NC_WRITE) != 0;
   if ((status = parse_header (f)) != NC_NOERR) {
      fclose (f->fp);
      free_file (f);
      return status;
   }
   *ncidp = id;
   return NC_NOERR;
}

int
nc_redef (int ncid)
{
   nc3_file *f = get_file (ncid);
   if (f == NULL)
      return NC_EBADID;
   if (!f->writable)
      return NC_EPERM;
   if (f->define_mode)
      return NC_EINDEFINE;
   f->define_mode = 1;
   return NC_NOERR;
}

int
nc_enddef (int ncid)
{
   nc3_file *f = get_file (ncid);
   if (f == NULL)
      return NC_EBADID;
   if (!f->define_mode)
      return NC_ENOTINDEFINE;
   return do_enddef (f);
}

int
nc_close (int ncid)
{
   nc3_file *f = get_file (ncid);
   int status = NC_NOERR;
   if (f == NULL)
      return NC_EBADID;
   if (f->define_mode && f->writable)
      status = do_enddef (f);
   if (fclose (f->fp) && status == NC_NOERR)
      status = NC_EIO;
   free_file (f);
   return status;
}

int
nc_def_dim (int ncid, const char *name, size_t len, int *dimidp)
{
   nc3_file *f = get_file (ncid);
   nc3_dim *nd;
   int i;
   if (f == NULL)
      return NC_EBADID;
   if (!f->define_mode)
      return NC_ENOTINDEFINE;
   for (i = 0; i < f->ndims; i++)
      if (strcmp (f->dims[i].name, name) == 0)
         return NC_ENAMEINUSE;
   if ((nd = realloc (f->dims, (size_t) (f->ndims + 1) * sizeof (nc3_dim))) == NULL)
      return NC_ENOMEM;
   f->dims = nd;
   if ((nd[f->ndims].name = strdup (name)) == NULL)
      return NC_ENOMEM;
   nd[f->ndims].len = len;
   if (dimidp)
      *dimidp = f->ndims;
   f->ndims++;
   f->header_dirty = 1;
   return NC_NOERR;
}

int
nc_inq_dimid (int ncid, const char *name, int *dimidp)
{
   nc3_file *f = get_file (ncid);
   int i;
   if (f == NULL)
      return NC_EBADID;
   for (i = 0; i < f->ndims; i++)
      if (strcmp (f->dims[i].name, name) == 0) {
         if (dimidp)
            *dimidp = i;
         return NC_NOERR;
      }
   return NC_EBADDIM;
}

int
nc_inq_dimlen (int ncid, int dimid, size_t *lenp)
{
   nc3_file *f = get_file (ncid);
   if (f == NULL)
      return NC_EBADID;
   if (dimid < 0 || dimid >= f->ndims)
      return NC_EBADDIM;
   if (lenp)
      *lenp = f->dims[dimid].len ? f->dims[dimid].len : f->numrecs;
   return NC_NOERR;
}

int
nc_def_var (int ncid, const char *name, nc_type xtype, int ndims, const int *dimids, int *varidp)
{
   nc3_file *f = get_file (ncid);
   nc3_var *nv, *v;
   int i;
   if (f == NULL)
      return NC_EBADID;
   if (!f->define_mode)
      return NC_ENOTINDEFINE;
   if (type_size (xtype) == 0)
      return NC_EBADTYPE;
   if (ndims < 0)
      return NC_EINVAL;
   for (i = 0; i < f->nvars; i++)
      if (strcmp (f->vars[i].name, name) == 0)
         return NC_ENAMEINUSE;
   for (i = 0; i < ndims; i++) {
      if (dimids[i] < 0 || dimids[i] >= f->ndims)
         return NC_EBADDIM;
      if (i > 0 && f->dims[dimids[i]].len == 0)
         return NC_EUNLIMPOS;
   }
   if ((nv = realloc (f->vars, (size_t) (f->nvars + 1) * sizeof (nc3_var))) == NULL)
      return NC_ENOMEM;
   f->vars = nv;
   v = &nv[f->nvars];
   memset (v, 0, sizeof (*v));
   if ((v->name = strdup (name)) == NULL)
      return NC_ENOMEM;
   v->ndims = ndims;
   if ((v->dimids = calloc ((size_t) (ndims ? ndims : 1), sizeof (int))) == NULL)
      return NC_ENOMEM;
   for (i = 0; i < ndims; i++)
      v->dimids[i] = dimids[i];
   v->type = xtype;
   v->old_begin = (uint64_t) - 1;
   v->is_rec = (ndims > 0 && f->dims[dimids[0]].len == 0);
   if (varidp)
      *varidp = f->nvars;
   f->nvars++;
   f->header_dirty = 1;
   return NC_NOERR;
}

int
nc_inq_varid (int ncid, const char *name, int *varidp)
{
   nc3_file *f = get_file (ncid);
   int i;
   if (f == NULL)
      return NC_EBADID;
   for (i = 0; i < f->nvars; i++)
      if (strcmp (f->vars[i].name, name) == 0) {
         if (varidp)
            *varidp = i;
         return NC_NOERR;
      }
   return NC_ENOTVAR;
}

/* ---- attributes ------------------------------------------------------------ */

static int
att_list (nc3_file * f, int varid, int **np, nc3_att *** ap)
{
   if (varid == NC_GLOBAL) {
      *np = &f->ngatts;
      *ap = &f->gatts;
      return NC_NOERR;
   }
   if (varid < 0 || varid >= f->nvars)
      return NC_ENOTVAR;
   *np = &f->vars[varid].natts;
   *ap = &f->vars[varid].atts;
   return NC_NOERR;
}

static void
store_be (unsigned char *dst, const void *src, size_t size)
{
   /* host is little-endian x86-64 / aarch64-le: reverse the bytes */
   const unsigned char *s = (const unsigned char *) src;
   size_t i;
   for (i = 0; i < size; i++)
      dst[i] = s[size - 1 - i];
}

static int
put_att_raw (int ncid, int varid, const char *name, nc_type type, size_t nelems, unsigned char *data)
{
   nc3_file *f = get_file (ncid);
   int *np;
   nc3_att **ap, *a;
   int i, status;
   if (f == NULL) {
      free (data);
      return NC_EBADID;
   }
   if (!f->define_mode) {
      free (data);
      return NC_ENOTINDEFINE;
   }
   if ((status = att_list (f, varid, &np, &ap)) != NC_NOERR) {
      free (data);
      return status;
   }
   for (i = 0; i < *np; i++)
      if (strcmp ((*ap)[i].name, name) == 0) {
         free ((*ap)[i].data);
         (*ap)[i].type = type;
         (*ap)[i].nelems = nelems;
         (*ap)[i].data = data;
         f->header_dirty = 1;
         return NC_NOERR;
      }
   if ((a = realloc (*ap, (size_t) (*np + 1) * sizeof (nc3_att))) == NULL) {
      free (data);
      return NC_ENOMEM;
   }
   *ap = a;
   if ((a[*np].name = strdup (name)) == NULL) {
      free (data);
      return NC_ENOMEM;
   }
   a[*np].type = type;
   a[*np].nelems = nelems;
   a[*np].data = data;
   (*np)++;
   f->header_dirty = 1;
   return NC_NOERR;
}

int
nc_put_att_text (int ncid, int varid, const char *name, size_t len, const char *tp)
{
   unsigned char *d = malloc (len ? len : 1);
   if (d == NULL)
      return NC_ENOMEM;
   memcpy (d, tp, len);
   return put_att_raw (ncid, varid, name, NC_CHAR, len, d);
}

int
nc_put_att_int (int ncid, int varid, const char *name, nc_type xtype, size_t len, const int *ip)
{
   size_t ts = type_size (xtype), i;
   unsigned char *d;
   if (xtype == NC_CHAR)
      return NC_ECHAR;
   if (ts == 0)
      return NC_EBADTYPE;
   if ((d = malloc (len * ts ? len * ts : 1)) == NULL)
      return NC_ENOMEM;
   for (i = 0; i < len; i++) {
      switch (xtype) {
      case NC_BYTE:{
            signed char v = (signed char) ip[i];
            store_be (d + i * ts, &v, ts);
            break;
         }
      case NC_SHORT:{
            int16_t v = (int16_t) ip[i];
            store_be (d + i * ts, &v, ts);
            break;
         }
      case NC_INT:{
            int32_t v = (int32_t) ip[i];
            store_be (d + i * ts, &v, ts);
            break;
         }
      case NC_FLOAT:{
            float v = (float) ip[i];
            store_be (d + i * ts, &v, ts);
            break;
         }
      default:{
            double v = (double) ip[i];
            store_be (d + i * ts, &v, ts);
            break;
         }
      }
   }
   return put_att_raw (ncid, varid, name, xtype, len, d);
}

int
nc_put_att_double (int ncid, int varid, const char *name, nc_type xtype, size_t len, const double *dp)
{
   size_t ts = type_size (xtype), i;
   unsigned char *d;
   if (xtype == NC_CHAR)
      return NC_ECHAR;
   if (ts == 0)
      return NC_EBADTYPE;
   if ((d = malloc (len * ts ? len * ts : 1)) == NULL)
      return NC_ENOMEM;
   for (i = 0; i < len; i++) {
      switch (xtype) {
      case NC_BYTE:{
            signed char v = (signed char) dp[i];
            store_be (d + i * ts, &v, ts);
            break;
         }
      case NC_SHORT:{
            int16_t v = (int16_t) dp[i];
            store_be (d + i * ts, &v, ts);
            break;
         }
      case NC_INT:{
            int32_t v = (int32_t) dp[i];
            store_be (d + i * ts, &v, ts);
            break;
         }
      case NC_FLOAT:{
            float v = (float) dp[i];
            store_be (d + i * ts, &v, ts);
            break;
         }
      default:{
            double v = dp[i];
            store_be (d + i * ts, &v, ts);
            break;
         }
      }
   }
   return put_att_raw (ncid, varid, name, xtype, len, d);
}

static double
load_as_double (const unsigned char *p, nc_type t)
{
   unsigned char tmp[8];
   size_t ts = type_size (t), i;
   for (i = 0; i < ts; i++)
      tmp[i] = p[ts - 1 - i];
   switch (t) {
   case NC_BYTE:
      return (double) *(signed char *) tmp;
   case NC_SHORT:{
         int16_t v;
         memcpy (&v, tmp, 2);
         return (double) v;
      }
   case NC_INT:{
         int32_t v;
         memcpy (&v, tmp, 4);
         return (double) v;
      }
   case NC_FLOAT:{
         float v;
         memcpy (&v, tmp, 4);
         return (double) v;
      }
   default:{
         double v;
         memcpy (&v, tmp, 8);
         return v;
      }
   }
}

int
nc_get_att_double (int ncid, int varid, const char *name, double *dp)
{
   nc3_file *f = get_file (ncid);
   int *np;
   nc3_att **ap;
   int i, status;
   size_t e;
   if (f == NULL)
      return NC_EBADID;
   if ((status = att_list (f, varid, &np, &ap)) != NC_NOERR)
      return status;
   for (i = 0; i < *np; i++) {
      nc3_att *a = &(*ap)[i];
      if (strcmp (a->name, name) != 0)
         continue;
      if (a->type == NC_CHAR)
         return NC_ECHAR;
      for (e = 0; e < a->nelems; e++)
         dp[e] = load_as_double (a->data + e * type_size (a->type), a->type);
      return NC_NOERR;
   }
   return NC_ENOTATT;
}

/* ---- whole-variable data access ---------------------------------------------- */

#define IO_CHUNK_ELEMS (1u << 20)

static inline uint32_t
bswap32 (uint32_t v)
{
   return __builtin_bswap32 (v);
}

static inline uint64_t
bswap64 (uint64_t v)
{
   return __builtin_bswap64 (v);
}

/* kind: 0 = caller buffer is int, 1 = caller buffer is double */
static int
get_var (int ncid, int varid, void *out, int kind)
{
   nc3_file *f = get_file (ncid);
   nc3_var *v;
   uint64_t n, done = 0;
   size_t ts;
   unsigned char *buf;
   int range_err = 0;
   if (f == NULL)
      return NC_EBADID;
   if (f->define_mode)
      return NC_EINDEFINE;
   if (varid < 0 || varid >= f->nvars)
      return NC_ENOTVAR;
   v = &f->vars[varid];
   if (v->type == NC_CHAR)
      return NC_ECHAR;
   if (v->is_rec && f->numrecs > 1)
      return NC_EINVAL;         /* interleaved records: not needed by the reference's inputs here */
   n = var_nelems (f, v);
   ts = type_size (v->type);
   if ((buf = malloc ((size_t) IO_CHUNK_ELEMS * ts)) == NULL)
      return NC_ENOMEM;
   if (fseeko (f->fp, (off_t) v->begin, SEEK_SET)) {
      free (buf);
      return NC_EIO;
   }
   while (done < n) {
      size_t c = (n - done) > IO_CHUNK_ELEMS ? IO_CHUNK_ELEMS : (size_t) (n - done);
      size_t i;
      if (fread (buf, ts, c, f->fp) != c) {
         free (buf);
         return NC_EIO;
      }
      if (kind == 1 && v->type == NC_DOUBLE) {
         uint64_t *o = (uint64_t *) out + done;
         const uint64_t *s = (const uint64_t *) buf;
         for (i = 0; i < c; i++)
            o[i] = bswap64 (s[i]);
      } else if (kind == 0 && v->type == NC_INT) {
         uint32_t *o = (uint32_t *) out + done;
         const uint32_t *s = (const uint32_t *) buf;
         for (i = 0; i < c; i++)
            o[i] = bswap32 (s[i]);
      } else {
         for (i = 0; i < c; i++) {
            double d = load_as_double (buf + i * ts, v->type);
            if (kind == 1)
               ((double *) out)[done + i] = d;
            else {
               if (d > 2147483647.0 || d < -2147483648.0)
                  range_err = 1;
               ((int *) out)[done + i] = (int) d;
            }
         }
      }
      done += c;
   }
   free (buf);
   return range_err ? NC_ERANGE : NC_NOERR;
}

static int
put_var (int ncid, int varid, const void *in, int kind)
{
   nc3_file *f = get_file (ncid);
   nc3_var *v;
   uint64_t n, done = 0;
   size_t ts;
   unsigned char *buf;
   if (f == NULL)
      return NC_EBADID;
   if (!f->writable)
      return NC_EPERM;
   if (f->define_mode)
      return NC_EINDEFINE;
   if (varid < 0 || varid >= f->nvars)
      return NC_ENOTVAR;
   v = &f->vars[varid];
   if (v->type == NC_CHAR)
      return NC_ECHAR;
   if (v->is_rec && f->numrecs > 1)
      return NC_EINVAL;         /* interleaved records; one record is one contiguous slab, like get_var */
   n = var_nelems (f, v);
   ts = type_size (v->type);
   if ((buf = malloc ((size_t) IO_CHUNK_ELEMS * ts)) == NULL)
      return NC_ENOMEM;
   if (fseeko (f->fp, (off_t) v->begin, SEEK_SET)) {
      free (buf);
      return NC_EIO;
   }
   while (done < n) {
      size_t c = (n - done) > IO_CHUNK_ELEMS ? IO_CHUNK_ELEMS : (size_t) (n - done);
      size_t i;
      if (kind == 1 && v->type == NC_DOUBLE) {
         const uint64_t *s = (const uint64_t *) in + done;
         uint64_t *o = (uint64_t *) buf;
         for (i = 0; i < c; i++)
            o[i] = bswap64 (s[i]);
      } else if (kind == 0 && v->type == NC_INT) {
         const uint32_t *s = (const uint32_t *) in + done;
         uint32_t *o = (uint32_t *) buf;
         for (i = 0; i < c; i++)
            o[i] = bswap32 (s[i]);
      } else {
         for (i = 0; i < c; i++) {
            double d = kind == 1 ? ((const double *) in)[done + i] : (double) ((const int *) in)[done + i];
            unsigned char *p = buf + i * ts;
            switch (v->type) {
            case NC_BYTE:{
                  signed char x = (signed char) d;
                  store_be (p, &x, ts);
                  break;
               }
            case NC_SHORT:{
                  int16_t x = (int16_t) d;
                  store_be (p, &x, ts);
                  break;
               }
            case NC_INT:{
                  int32_t x = (int32_t) d;
                  store_be (p, &x, ts);
                  break;
               }
            case NC_FLOAT:{
                  float x = (float) d;
                  store_be (p, &x, ts);
                  break;
               }
            default:{
                  double x = d;
                  store_be (p, &x, ts);
                  break;
               }
            }
         }
      }
      if (fwrite (buf, ts, c, f->fp) != c) {
         free (buf);
         return NC_EIO;
      }
      done += c;
   }
   free (buf);
   fflush (f->fp);
   return NC_NOERR;
}

int
nc_get_var_int (int ncid, int varid, int *ip)
{
   return get_var (ncid, varid, ip, 0);
}

int
nc_get_var_double (int ncid, int varid, double *dp)
{
   return get_var (ncid, varid, dp, 1);
}

int
nc_put_var_int (int ncid, int varid, const int *ip)
{
   return put_var (ncid, varid, ip, 0);
}

int
nc_put_var_double (int ncid, int varid, const double *dp)
{
   return put_var (ncid, varid, dp, 1);
}

/* Not part of the NetCDF API (declared in include/nkp_nc3.h): where the data of a fixed-size variable
 * lies in the file and what it is, so that a caller can read the big-endian bytes itself and have
 * them converted on the GPU (nkp_factor_be) instead of by the host loops above. */
int
nkp_nc3_inq_var_extent (int ncid, int varid, long long *offset, long long *nbytes, int *xtype)
{
   nc3_file *f = get_file (ncid);
   const nc3_var *v;
   if (f == NULL)
      return NC_EBADID;
   if (f->define_mode)
      return NC_EINDEFINE;
   if (varid < 0 || varid >= f->nvars)
      return NC_ENOTVAR;
   v = &f->vars[varid];
   if (v->is_rec)
      return NC_EINVAL;
   if (offset)
      *offset = (long long) v->begin;
   if (nbytes)
      *nbytes = (long long) (var_nelems (f, v) * type_size (v->type));
   if (xtype)
      *xtype = (int) v->type;
   return NC_NOERR;
}
