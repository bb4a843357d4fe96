/*
 * oracle/merge_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement ("port") of the reference's per-voxel combination path:
 *   src/merging.c:157-283  mpdaf_merging_median
 *   src/merging.c:289-517  mpdaf_merging_sigma_clipping
 *   src/tools.c:9-24       mpdaf_mean           (sequential mean / population sigma)
 *   src/tools.c:27-36      mpdaf_sum
 *   src/tools.c:39-51      mpdaf_median         (argsort, middle / mean of two middles)
 *   src/tools.c:54-77      mpdaf_mean_mad
 *   src/tools.c:82-118     mpdaf_mean_sigma_clip
 *   src/tools.c:122-158    mpdaf_mean_madsigma_clip
 *
 * It works on in-memory exposure arrays (f32 or f64) instead of FITS files,
 * so it can be fed exactly the same synthetic slabs as the CUDA library.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load it.  It is pinned against the reference itself: tests/test_oracle.py
 * compares it with oracle/_ref/libmpdaf_ref.so (the UNMODIFIED reference C
 * sources compiled against oracle/standin/fitsio.h), with the known-answer
 * vectors of SURVEY.md section 8(c) and with the arithmetic of the
 * reference's own lib/mpdaf/obj/tests/test_cubelist.py.
 *
 * What is restated, in the order the reference does it:
 *  - gather the non-NaN samples of a voxel in ascending file order, applying
 *    w = (offset + d) * scale and v = s * scale * scale (merging.c:425-436);
 *  - statistics of the current sample set taken in the *incoming* order of
 *    the index list (file order on the first evaluation, ascending value
 *    order afterwards because the list has been argsorted and compacted);
 *  - the argsort, the median-centred strict clip window, the stop rules
 *    (ni < nstop, ni == n, nmax exhausted) and the in-order compaction.
 * The argsort here is a stable insertion sort, not the reference's
 * Numerical-Recipes quicksort (tools.c:307-372): equal keys are equal
 * values, so no statistic of the data depends on their relative order; only
 * the propagated-variance sum can differ in its last bits (1e-6 contract).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_N 512

static void argsort_stable(const double *w, int *idx, int n)
{
    for (int j = 1; j < n; j++) {
        int t = idx[j];
        double key = w[t];
        int i = j - 1;
        while (i >= 0 && w[idx[i]] > key) {
            idx[i + 1] = idx[i];
            i--;
        }
        idx[i + 1] = t;
    }
}

/* tools.c:39-51 on an already argsorted list */
static double middle_value(const double *w, const int *idx, int n)
{
    if (n % 2 == 0) return (w[idx[n / 2]] + w[idx[n / 2 - 1]]) / 2;
    return w[idx[n / 2]];
}

/*
 * One voxel's iterative clip (tools.c:82-118 and :122-158 unrolled into a
 * loop; the reference recurses with nmax-1 on the compacted list).
 * idx[0..n) is the caller's index list (0..n-1 on entry, merging.c:429); on
 * return idx[0..x[2]) are the kept samples in ascending value order.
 * x[0]=mean, x[1]=sigma (population sigma, or 1.4826*MAD), x[2]=count.
 */
void oracle_clip(const double *w, int n, double x[3], int nmax, double nclip_low,
                 double nclip_up, int nstop, int mad, int *idx)
{
    double work[ORACLE_MAX_N];
    int widx[ORACLE_MAX_N];
    for (;;) {
        /* mean in incoming order (tools.c:13-17, :61-65) */
        double acc = 0.0;
        for (int i = 0; i < n; i++) acc += w[idx[i]];
        double mean = acc / n;
        double sigma, med;
        if (!mad) {
            double dev = 0.0;               /* tools.c:18-23, incoming order */
            for (int i = 0; i < n; i++) dev += (w[idx[i]] - mean) * (w[idx[i]] - mean);
            sigma = sqrt(dev / n);
            argsort_stable(w, idx, n);      /* tools.c:88 */
            med = middle_value(w, idx, n);
        } else {
            argsort_stable(w, idx, n);      /* tools.c:67 */
            med = middle_value(w, idx, n);
            for (int i = 0; i < n; i++) {   /* tools.c:68-74 */
                widx[i] = i;
                work[i] = fabs(w[idx[i]] - med);
            }
            argsort_stable(work, widx, n);
            sigma = middle_value(work, widx, n) * 1.4826;
        }
        x[0] = mean;
        x[1] = sigma;
        x[2] = n;
        double clip_lo = med - (nclip_low * sigma);     /* tools.c:89-90 */
        double clip_up = med + (nclip_up * sigma);
        int ni = 0;
        for (int i = 0; i < n; i++)
            if (w[idx[i]] < clip_up && w[idx[i]] > clip_lo) ni++;
        if (ni < nstop || ni == n) return;              /* tools.c:100-103 */
        if (nmax <= 0) return;                          /* tools.c:104 */
        int k = 0;
        for (int i = 0; i < n; i++)                     /* tools.c:106-114 */
            if (w[idx[i]] < clip_up && w[idx[i]] > clip_lo) idx[k++] = idx[i];
        n = k;
        nmax--;
    }
}

static inline double sample(const void *base, int elem_bytes, long long v)
{
    return elem_bytes == 4 ? (double)((const float *)base)[v] : ((const double *)base)[v];
}

/*
 * merging.c:289-517 on arrays.  data[i] / var[i] point at exposure i's DATA /
 * STAT samples (elem_bytes 4 = float32, 8 = float64), nvox voxels each.
 * var may be NULL unless typ_var == 0.  selected_pix / valid_pix are ADDED to
 * (merging.c:486-491), the caller zero-fills.
 */
int oracle_sigma_clipping(int nfiles, const void *const *data, const void *const *var,
                          int elem_bytes, long long nvox, double *out, double *outvar,
                          int *expmap, const double *scale, const double *offset,
                          int *selected_pix, int *valid_pix, int nmax, double nclip_low,
                          double nclip_up, int nstop, int typ_var, int mad)
{
    if (nfiles < 1 || nfiles > ORACLE_MAX_N) return 1;
#pragma omp parallel
    {
        long long valid[ORACLE_MAX_N] = {0}, select[ORACLE_MAX_N] = {0};
        double wdata[ORACLE_MAX_N], wvar[ORACLE_MAX_N], x[3];
        int idx[ORACLE_MAX_N], files_id[ORACLE_MAX_N];
#pragma omp for schedule(static)
        for (long long v = 0; v < nvox; v++) {
            int n = 0;
            for (int i = 0; i < nfiles; i++) {          /* merging.c:425-436 */
                double d = sample(data[i], elem_bytes, v);
                if (!isnan(d)) {
                    wdata[n] = (offset[i] + d) * scale[i];
                    files_id[n] = i;
                    idx[n] = n;
                    if (typ_var == 0)
                        wvar[n] = sample(var[i], elem_bytes, v) * scale[i] * scale[i];
                    n++;
                    valid[i]++;
                }
            }
            if (n == 0) {                               /* merging.c:438-441 */
                out[v] = NAN;
                expmap[v] = 0;
                outvar[v] = NAN;
            } else if (n == 1) {                        /* merging.c:442-449 */
                out[v] = wdata[0];
                expmap[v] = 1;
                outvar[v] = typ_var == 0 ? wvar[0] : NAN;
                select[files_id[0]]++;
            } else {                                    /* merging.c:450-476 */
                oracle_clip(wdata, n, x, nmax, nclip_low, nclip_up, nstop, mad, idx);
                int kept = (int)x[2];
                out[v] = x[0];
                expmap[v] = kept;
                if (typ_var == 0) {
                    double s = 0;
                    for (int k = 0; k < kept; k++) s += wvar[idx[k]];
                    outvar[v] = s / (x[2] * x[2]);
                } else if (x[2] > 1) {
                    outvar[v] = x[1] * x[1];
                    if (typ_var == 1) outvar[v] /= (x[2] - 1);
                } else {
                    outvar[v] = NAN;
                }
                for (int k = 0; k < kept; k++) select[files_id[idx[k]]]++;
            }
        }
#pragma omp critical
        for (int i = 0; i < nfiles; i++) {
            valid_pix[i] += (int)valid[i];
            selected_pix[i] += (int)select[i];
        }
    }
    return 0;
}

/* merging.c:157-283 on arrays (no scale/offset, no variance). */
int oracle_median(int nfiles, const void *const *data, int elem_bytes, long long nvox,
                  double *out, int *expmap, int *valid_pix)
{
    if (nfiles < 1 || nfiles > ORACLE_MAX_N) return 1;
#pragma omp parallel
    {
        long long valid[ORACLE_MAX_N] = {0};
        double wdata[ORACLE_MAX_N];
        int idx[ORACLE_MAX_N];
#pragma omp for schedule(static)
        for (long long v = 0; v < nvox; v++) {
            int n = 0;
            for (int i = 0; i < nfiles; i++) {          /* merging.c:233-241 */
                double d = sample(data[i], elem_bytes, v);
                if (!isnan(d)) {
                    wdata[n] = d;
                    idx[n] = n;
                    n++;
                    valid[i]++;
                }
            }
            if (n == 0) {                               /* merging.c:243-252 */
                out[v] = NAN;
                expmap[v] = 0;
            } else if (n == 1) {
                out[v] = wdata[0];
                expmap[v] = 1;
            } else {
                argsort_stable(wdata, idx, n);
                out[v] = middle_value(wdata, idx, n);
                expmap[v] = n;
            }
        }
#pragma omp critical
        for (int i = 0; i < nfiles; i++) valid_pix[i] += (int)valid[i];
    }
    return 0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
