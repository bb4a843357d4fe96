/*
 * oracle/standin/fitsio_standin.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Minimal FITS image access behind the cfitsio names declared in the
 * stand-in fitsio.h.  Semantics restated from the cfitsio user guide as the
 * reference uses them (merging.c:93-110, :228, :411, :417): open
 * "file[EXTNAME]" read-only, report NAXIS / NAXISn, and read `nelem`
 * consecutive pixels starting at the 1-based coordinate `firstpix`, widened
 * to double (f32 -> f64 widening is exact; big-endian on disk).
 */
#define _FILE_OFFSET_BITS 64
#include "fitsio.h"

#include <ctype.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

enum { BLOCK = 2880, CARD = 80, ERR_OPEN = 104, ERR_HDU = 301, ERR_READ = 108,
       ERR_TYPE = 410 };

static int card_is(const char *card, const char *key)
{
    size_t k = strlen(key);
    if (strncmp(card, key, k) != 0) return 0;
    for (size_t i = k; i < 8; i++)
        if (card[i] != ' ') return 0;
    return 1;
}

static long long card_int(const char *card) { return atoll(card + 10); }

static void card_str(const char *card, char *out, size_t cap)
{
    const char *p = strchr(card + 10, '\'');
    size_t n = 0;
    out[0] = 0;
    if (!p || p >= card + CARD) return;
    for (p++; p < card + CARD && *p != '\'' && n + 1 < cap; p++) out[n++] = *p;
    while (n > 0 && out[n - 1] == ' ') n--;
    out[n] = 0;
}

static int same_name(const char *a, const char *b)
{
    for (; *a && *b; a++, b++)
        if (toupper((unsigned char)*a) != toupper((unsigned char)*b)) return 0;
    return *a == 0 && *b == 0;
}

/* Walk the HDU list until EXTNAME matches (or take HDU 0 when ext is empty). */
static int locate_hdu(fitsfile *f, const char *ext)
{
    long long off = 0;
    char blk[BLOCK];
    for (int hdu = 0;; hdu++) {
        int bitpix = 0, naxis = 0, done = 0;
        long long ax[8] = {0}, pcount = 0, gcount = 1;
        char extname[72] = "";
        while (!done) {
            if (pread(f->fd, blk, BLOCK, off) != BLOCK) return ERR_HDU;
            off += BLOCK;
            for (int c = 0; c < BLOCK / CARD; c++) {
                const char *card = blk + c * CARD;
                if (card_is(card, "END")) { done = 1; break; }
                if (card_is(card, "BITPIX")) bitpix = (int)card_int(card);
                else if (card_is(card, "NAXIS")) naxis = (int)card_int(card);
                else if (card_is(card, "PCOUNT")) pcount = card_int(card);
                else if (card_is(card, "GCOUNT")) gcount = card_int(card);
                else if (card_is(card, "EXTNAME")) card_str(card, extname, sizeof extname);
                else if (!strncmp(card, "NAXIS", 5) && isdigit((unsigned char)card[5])) {
                    int k = atoi(card + 5);
                    if (k >= 1 && k <= 8 && card[8] == '=') ax[k - 1] = card_int(card);
                }
            }
        }
        long long nelem = naxis > 0 ? 1 : 0;
        for (int k = 0; k < naxis; k++) nelem *= ax[k];
        long long bytes = (llabs(bitpix) / 8) * gcount * (pcount + nelem);
        if ((ext[0] == 0 && hdu == 0) || (ext[0] && same_name(extname, ext))) {
            f->data_off = off;
            f->bitpix = bitpix;
            f->naxis = naxis;
            for (int k = 0; k < 3; k++) f->naxes[k] = k < naxis ? (long)ax[k] : 1;
            return 0;
        }
        off += (bytes + BLOCK - 1) / BLOCK * BLOCK;
    }
}

int fits_open_file(fitsfile **fptr, const char *name_ext, int mode, int *status)
{
    (void)mode;
    if (*status) return *status;
    fitsfile *f = (fitsfile *)calloc(1, sizeof *f);
    char ext[72] = "";
    snprintf(f->name, sizeof f->name, "%s", name_ext);
    char *br = strrchr(f->name, '[');
    if (br && f->name[strlen(f->name) - 1] == ']') {
        snprintf(ext, sizeof ext, "%.*s", (int)(strlen(br) - 2), br + 1);
        *br = 0;
    }
    f->fd = open(f->name, O_RDONLY);
    if (f->fd < 0) { free(f); return *status = ERR_OPEN; }
    int rc = locate_hdu(f, ext);
    if (rc) { close(f->fd); free(f); return *status = rc; }
    *fptr = f;
    return 0;
}

int fits_close_file(fitsfile *f, int *status)
{
    (void)status;
    if (f) { close(f->fd); free(f); }
    return 0;
}

void fits_report_error(FILE *stream, int status)
{
    fprintf(stream, "stand-in fitsio: error status %d\n", status);
}

int fits_get_img_dim(fitsfile *f, int *naxis, int *status)
{
    if (*status) return *status;
    *naxis = f->naxis;
    return 0;
}

int fits_get_img_size(fitsfile *f, int maxdim, long *naxes, int *status)
{
    if (*status) return *status;
    for (int k = 0; k < maxdim && k < 3; k++) naxes[k] = f->naxes[k];
    return 0;
}

int fits_read_pix(fitsfile *f, int datatype, long *firstpix, LONGLONG nelem,
                  void *nulval, void *array, int *anynul, int *status)
{
    (void)nulval; (void)anynul;
    if (*status) return *status;
    if (datatype != TDOUBLE || (f->bitpix != -32 && f->bitpix != -64))
        return *status = ERR_TYPE;
    long long first = ((long long)(firstpix[2] - 1) * f->naxes[1] + (firstpix[1] - 1))
                      * f->naxes[0] + (firstpix[0] - 1);
    int esz = -f->bitpix / 8;
    double *out = (double *)array;
    unsigned char buf[1 << 16];
    long long done = 0;
    while (done < nelem) {
        long long chunk = nelem - done;
        if (chunk > (long long)(sizeof buf) / esz) chunk = (long long)(sizeof buf) / esz;
        ssize_t want = (ssize_t)(chunk * esz);
        if (pread(f->fd, buf, want, f->data_off + (first + done) * esz) != want)
            return *status = ERR_READ;
        if (esz == 4) {
            for (long long i = 0; i < chunk; i++) {
                uint32_t u; float v;
                memcpy(&u, buf + 4 * i, 4);
                u = __builtin_bswap32(u);
                memcpy(&v, &u, 4);
                out[done + i] = (double)v;
            }
        } else {
            for (long long i = 0; i < chunk; i++) {
                uint64_t u;
                memcpy(&u, buf + 8 * i, 8);
                u = __builtin_bswap64(u);
                memcpy(&out[done + i], &u, 8);
            }
        }
        done += chunk;
    }
    return 0;
}
