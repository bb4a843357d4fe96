/*
 * mpdaf_b200.h -- C ABI of the B200-native replacement for MPDAF's `_ctools`
 * shared library (reference: src/merging.c, src/tools.c, src/tools.h, loaded by
 * lib/mpdaf/tools/ctools.py:52-92).
 *
 * The library file is `_ctools.so`; it can be dropped next to the reference's
 * ctools.py (numpy.ctypeslib.load_library("_ctools", dir)) unchanged.  Every
 * entry point takes plain pointers and sizes.  There is no CPU fallback: the
 * merging entry points return MPDAF_B200_ERR_CUDA when no sm_100 device is
 * usable, and never report success with partially written outputs.
 *
 * Section A: the two symbols ctools.py binds (same names, same signatures).
 * Section B: the src/tools.h helper symbols the reference .so also exports.
 * Section C: array entry points (SURVEY.md 8b "needed extension"): the same
 *            computation on in-memory exposure stacks, host or device resident,
 *            so FITS decoding can stay outside the timed region.
 * Section D: FITS image probing and the synthetic-exposure generator used by
 *            the tests and bench.py.
 */
#ifndef MPDAF_B200_H
#define MPDAF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return codes ------------------------------------------------------ */
#define MPDAF_B200_OK 0
#define MPDAF_B200_ERR_ARG 1     /* bad argument (nfiles, element type, null pointer) */
#define MPDAF_B200_ERR_CUDA 2    /* CUDA runtime error or no usable device */
#define MPDAF_B200_ERR_IO 3      /* file missing / unreadable / not a FITS cube */
#define MPDAF_B200_ERR_SHAPE 4   /* exposures do not share one shape (merging.c:195-199) */

/* element types of an exposure stack */
#define MPDAF_B200_F32 4         /* IEEE float32, host byte order (MUSE on-disk type) */
#define MPDAF_B200_F64 8         /* IEEE float64 */

/* where the arrays of an array entry point live */
#define MPDAF_B200_HOST 0        /* host memory (pageable or pinned); staged in slabs */
#define MPDAF_B200_DEVICE 1      /* device memory of `device`; nothing is copied */

#define MPDAF_B200_MAX_FILES 500 /* merging.c:16 MAX_FILES */

/* ---- A. drop-in symbols -------------------------------------------------
 * Replaces src/merging.c:157 (argtypes: ctools.py:69-74).  `input` is the
 * newline-joined, NUL-terminated list of FITS file names; each is opened as
 * name[DATA].  data/expmap have nz*ny*nx elements and are fully written;
 * valid_pix (one int per file) is ADDED to.  Unlike the reference the buffer
 * `input` is not modified (the reference's strtok writes NULs into it,
 * merging.c:43).  Returns 0 on success; a non-zero MPDAF_B200_ERR_* instead of
 * the reference's exit(EXIT_FAILURE). */
int mpdaf_merging_median(char *input, double *data, int *expmap, int *valid_pix);

/* Replaces src/merging.c:289-304 (argtypes: ctools.py:77-92).  typ_var: 0
 * 'propagate' (reads name[STAT]), 1 'stat_mean', 2 'stat_one'.  selected_pix
 * and valid_pix are ADDED to. */
int mpdaf_merging_sigma_clipping(char *input, double *data, double *var, int *expmap,
                                 double *scale, double *offset, int *selected_pix,
                                 int *valid_pix, int nmax, double nclip_low,
                                 double nclip_up, int nstop, int typ_var, int mad);

/* ---- B. src/tools.h helper symbols (host functions) ---------------------
 * Same names, signatures and results as src/tools.c:9-372; exported because the
 * reference `_ctools` exports them (non-static) and merging.pyx links them. */
void mpdaf_mean(double *data, int n, double x[3], int *indx);                /* tools.c:9 */
double mpdaf_sum(double *data, int n, int *indx);                            /* tools.c:27 */
double mpdaf_median(double *data, int n, int *indx);                         /* tools.c:39 */
void mpdaf_mean_mad(double *data, int n, double x[3], int *indx);            /* tools.c:54 */
int indexx(int n, double *arr, int *indx);                                   /* tools.c:307 */
void mpdaf_mean_sigma_clip(double *data, int n, double x[3], int nmax, double nclip_low,
                           double nclip_up, int nstop, int *indx);           /* tools.c:82 */
void mpdaf_mean_madsigma_clip(double *data, int n, double x[3], int nmax, double nclip_low,
                              double nclip_up, int nstop, int *indx);        /* tools.c:122 */
void mpdaf_median_sigma_clip(double *data, int n, double x[3], int nmax, double nclip_low,
                             double nclip_up, int nstop, int *indx);         /* tools.c:162 */
int mpdaf_locate(double *xx, int n, double x);                               /* tools.c:204 */
double mpdaf_linear_interpolation(double *xx, double *yy, int n, double x);  /* tools.c:221 */
void mpdaf_minmax(double data[], int n, int *indx, double res[]);            /* tools.c:234 */
void mpdaf_minmax_int(int data[], int n, int *indx, int res[]);              /* tools.c:262 */

/* ---- C. array entry points ----------------------------------------------
 * Same computation as section A on exposure stacks already in memory.
 *
 *   data[i] / var[i] : base address of exposure i's DATA / STAT samples, nvox
 *                      elements of `elem_type` each (the tables themselves are
 *                      host arrays of nfiles pointers).  var may be NULL unless
 *                      typ_var == 0.
 *   location         : MPDAF_B200_HOST   -> every sample/output pointer is host
 *                        memory; the library stages wavelength slabs through
 *                        `device` (H2D, kernel, D2H pipelined on its own streams)
 *                        and returns when the outputs are complete.
 *                      MPDAF_B200_DEVICE -> every sample/output pointer is device
 *                        memory of `device`; the kernel is enqueued on `stream`
 *                        (a cudaStream_t, NULL = default stream) and the call
 *                        returns after the per-exposure counters were read back.
 *   out/outvar/expmap: nvox elements, fully written (same location as the data).
 *   scale/offset     : host arrays of nfiles doubles, or NULL for 1 / 0.
 *   selected_pix/valid_pix : host arrays of nfiles ints, ADDED to (64-bit
 *                      accumulation on the device, narrowed at the end).
 */
int mpdaf_b200_sigma_clipping_arrays(int nfiles, const void *const *data,
                                     const void *const *var, int elem_type,
                                     long long nvox, int location, int device,
                                     double *out, double *outvar, int *expmap,
                                     const double *scale, const double *offset,
                                     int *selected_pix, int *valid_pix, int nmax,
                                     double nclip_low, double nclip_up, int nstop,
                                     int typ_var, int mad, void *stream);

int mpdaf_b200_median_arrays(int nfiles, const void *const *data, int elem_type,
                             long long nvox, int location, int device, double *out,
                             int *expmap, int *valid_pix, void *stream);

/* 64-bit counter variant for stacks whose per-exposure counts exceed INT_MAX
 * (multi-slab accumulation); same as above otherwise. */
int mpdaf_b200_sigma_clipping_arrays64(int nfiles, const void *const *data,
                                       const void *const *var, int elem_type,
                                       long long nvox, int location, int device,
                                       double *out, double *outvar, int *expmap,
                                       const double *scale, const double *offset,
                                       long long *selected_pix, long long *valid_pix,
                                       int nmax, double nclip_low, double nclip_up,
                                       int nstop, int typ_var, int mad, void *stream);

/* Device time (ms, CUDA events on the launching stream) of the combination
 * kernel(s) enqueued by the most recent array / drop-in call of this thread's
 * process, and the number of kernel launches that call made. */
double mpdaf_b200_last_kernel_ms(void);
int mpdaf_b200_last_launch_count(void);
/* Name of the kernel family the last call dispatched to ("sigclip_reg<48,unit>" ...). */
const char *mpdaf_b200_last_kernel_name(void);

/* Histogram of the exposure map written by the most recent merging call (hist[k] = number of
 * voxels combined from k exposures, k = 0 .. nbins-1): the mode over k >= 1 is the `expnb` of
 * CubeList._compute_expnb (cubelist.py:69-74), obtained on the device instead of a host pass. */
int mpdaf_b200_last_expmap_histogram(long long *hist, int nbins);

/* Finite mask of a DEVICE-resident output cube: mask[i] = 1 where data[i] is NaN or infinite, else 0 (one byte per
 * voxel, device memory of `device`), and the number of masked voxels.  What the reference computes on the host when
 * it wraps the combined arrays in a Cube (lib/mpdaf/obj/cubelist.py:399-400). */
int mpdaf_b200_finite_mask(const double *data, long long nvox, int device, unsigned char *mask,
                           long long *nmasked, void *stream);

/* Diagnostic: voxels that the certified fast path handed to the exact per-voxel fallback
 * on `device` since the last reset (see DESIGN.md "certified decisions"); -1 if unknown. */
long long mpdaf_b200_slow_voxels(int device, int reset);

/* Devices a HOST-location or file-name call spreads its wavelength slabs over (the reference deals plane ranges to
 * its OpenMP threads inside the call, merging.c:112-137; here chunk c of consecutive voxels goes to device c % n, each
 * device with its own staging slots and streams, results land at the chunk's offset of the caller's arrays and the
 * per-exposure counters are summed at the end).  n = 0 restores the default: MPDAF_B200_DEVICES ("all", a count or a
 * comma-separated list), else every usable device.  The array entry points take `device` = MPDAF_B200_ALL_DEVICES to use
 * this set and a device index >= 0 to use exactly that device; the file-name symbols use the set unless MPDAF_B200_DEVICE
 * names one device. */
#define MPDAF_B200_ALL_DEVICES (-1)
int mpdaf_b200_set_devices(int n, const int *ids);

/* Frees everything the library keeps between calls (per-device staging slots: device buffers, pinned host staging,
 * streams).  A long-lived process calls it when it is done combining; the next call allocates again. */
int mpdaf_b200_release(void);

/* Number of usable sm_100 devices (0 when none / no driver). */
int mpdaf_b200_device_count(void);
const char *mpdaf_b200_last_error(void);

/* ---- D. FITS probing and synthetic exposures ----------------------------
 * Shape and BITPIX of image HDU `extname` ("DATA", "STAT") of a FITS file;
 * naxes is x (fastest), y, z like cfitsio (merging.c:103-108). */
int mpdaf_b200_fits_image_info(const char *path, const char *extname, long naxes[3],
                               int *bitpix, long long *data_offset);

/* Counter-based synthetic MUSE-like exposure samples.  Fills
 * `count` float32 samples of exposure `exposure` starting at global voxel v0
 * (x fastest; nx, ny give the spaxel grid for the NaN border).  kind 0 = DATA,
 * 1 = STAT.  location as above; identical bits on host and device. */
int mpdaf_b200_synth_fill(void *dst, int location, int device, int kind,
                          unsigned long long seed, int exposure, long long v0,
                          long long count, int nx, int ny, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MPDAF_B200_H */
