/*
 * oracle/standin/fitsio.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Stand-in for the six cfitsio entry points that the reference's
 * src/merging.c uses (merging.c:93-110 open_fits, :228/:411/:417
 * fits_read_pix, :271/:498/:504 fits_close_file, :100/:275/:509
 * fits_report_error).  cfitsio is not installed in this image, so the
 * UNMODIFIED reference sources are compiled against this header to obtain
 * oracle/_ref/libmpdaf_ref.so (see oracle/Makefile).
 *
 * The implementation (fitsio_standin.c) parses real FITS files: primary HDU
 * plus IMAGE extensions, BITPIX -32 / -64, big-endian, selected by EXTNAME
 * through the cfitsio "path[extname]" syntax.  Only TDOUBLE reads exist.
 */
#ifndef MPDAF_B200_STANDIN_FITSIO_H
#define MPDAF_B200_STANDIN_FITSIO_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define READONLY 0
#define TDOUBLE 82

typedef long long LONGLONG;

typedef struct standin_fitsfile {
    int fd;            /* open file descriptor */
    long long data_off;/* byte offset of the first data element of the HDU */
    int bitpix;        /* -32 or -64 */
    int naxis;
    long naxes[3];
    char name[600];
} fitsfile;

int fits_open_file(fitsfile **fptr, const char *name_ext, int mode, int *status);
int fits_close_file(fitsfile *fptr, int *status);
void fits_report_error(FILE *stream, int status);
int fits_get_img_dim(fitsfile *fptr, int *naxis, int *status);
int fits_get_img_size(fitsfile *fptr, int maxdim, long *naxes, int *status);
int fits_read_pix(fitsfile *fptr, int datatype, long *firstpix, LONGLONG nelem,
                  void *nulval, void *array, int *anynul, int *status);

#ifdef __cplusplus
}
#endif
#endif
