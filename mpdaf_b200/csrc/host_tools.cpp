// Host implementations of the src/tools.h helper symbols.
//
// The reference `_ctools` exports these (src/tools.c:9-372, declared in
// src/tools.h:12-30); only merging.c and the Cython merging.pyx call them.  They
// are kept as plain host functions so that the replacement library exports the
// same symbol set, with the same results:
//   - sums run sequentially over data[indx[i]] in index-list order;
//   - the median argsorts the index list in place, then takes the middle value
//     (mean of the two middle values for even n);
//   - the clip routines evaluate statistics on the incoming order, argsort,
//     apply the strict window med -/+ nclip*sigma, stop when fewer than nstop
//     samples survive, when nothing is rejected, or when nmax is exhausted,
//     and otherwise compact the survivors to the front of the index list.
// Built with -ffp-contract=off: no fused multiply-add may change a rounding.
#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/mpdaf_b200.h"

namespace {

struct Moments {
    double mean;
    double spread;
};

double ordered_sum(const double* data, int n, const int* indx)
{
    double acc = 0.0;
    for (int i = 0; i < n; i++) acc += data[indx[i]];
    return acc;
}

Moments mean_and_sigma(const double* data, int n, const int* indx)
{
    const double mean = ordered_sum(data, n, indx) / n;
    double dev = 0.0;
    for (int i = 0; i < n; i++) {
        const double t = data[indx[i]] - mean;
        dev += t * t;
    }
    return {mean, std::sqrt(dev / n)};
}

void argsort_in_place(const double* arr, int* indx, int n)
{
    std::stable_sort(indx, indx + n, [arr](int a, int b) { return arr[a] < arr[b]; });
}

double middle(const double* data, const int* indx, int n)
{
    const int h = n / 2;
    return (n % 2 == 0) ? (data[indx[h]] + data[indx[h - 1]]) / 2 : data[indx[h]];
}

Moments mean_and_mad(const double* data, int n, int* indx)
{
    const double mean = ordered_sum(data, n, indx) / n;  // incoming order, before the sort
    argsort_in_place(data, indx, n);
    const double med = middle(data, indx, n);
    std::vector<double> dev(n);
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) {
        order[i] = i;
        dev[i] = std::fabs(data[indx[i]] - med);
    }
    argsort_in_place(dev.data(), order.data(), n);
    return {mean, middle(dev.data(), order.data(), n) * 1.4826};
}

enum class Centre { MeanSigma, MeanMad, MedianSigma };

void iterative_clip(double* data, int n, double x[3], int nmax, double nclip_low, double nclip_up,
                    int nstop, int* indx, Centre mode)
{
    for (;;) {
        Moments m = (mode == Centre::MeanMad) ? mean_and_mad(data, n, indx)
                                              : mean_and_sigma(data, n, indx);
        argsort_in_place(data, indx, n);
        const double med = middle(data, indx, n);
        x[0] = (mode == Centre::MedianSigma) ? med : m.mean;
        x[1] = m.spread;
        x[2] = n;
        const double lo = med - (nclip_low * m.spread);
        const double up = med + (nclip_up * m.spread);
        int inside = 0;
        for (int i = 0; i < n; i++) {
            const double v = data[indx[i]];
            if (v < up && v > lo) inside++;
        }
        if (inside < nstop || inside == n || nmax <= 0) return;
        int k = 0;
        for (int i = 0; i < n; i++) {
            const double v = data[indx[i]];
            if (v < up && v > lo) indx[k++] = indx[i];
        }
        n = k;
        nmax--;
    }
}

template <typename T>
void extremes(const T* data, int n, const int* indx, T* res)
{
    T lo = data[indx[0]], hi = data[indx[0]];
    for (int i = 1; i < n; i++) {
        const T v = data[indx[i]];
        if (v > hi) hi = v;
        else if (v < lo) lo = v;
    }
    res[0] = lo;
    res[1] = hi;
}

}  // namespace

extern "C" {

void mpdaf_mean(double* data, int n, double x[3], int* indx)
{
    const Moments m = mean_and_sigma(data, n, indx);
    x[0] = m.mean;
    x[1] = m.spread;
}

double mpdaf_sum(double* data, int n, int* indx) { return ordered_sum(data, n, indx); }

double mpdaf_median(double* data, int n, int* indx)
{
    argsort_in_place(data, indx, n);
    return middle(data, indx, n);
}

void mpdaf_mean_mad(double* data, int n, double x[3], int* indx)
{
    const Moments m = mean_and_mad(data, n, indx);
    x[0] = m.mean;
    x[1] = m.spread;
}

int indexx(int n, double* arr, int* indx)
{
    argsort_in_place(arr, indx, n);
    return 0;
}

void mpdaf_mean_sigma_clip(double* data, int n, double x[3], int nmax, double nclip_low,
                           double nclip_up, int nstop, int* indx)
{
    iterative_clip(data, n, x, nmax, nclip_low, nclip_up, nstop, indx, Centre::MeanSigma);
}

void mpdaf_mean_madsigma_clip(double* data, int n, double x[3], int nmax, double nclip_low,
                              double nclip_up, int nstop, int* indx)
{
    iterative_clip(data, n, x, nmax, nclip_low, nclip_up, nstop, indx, Centre::MeanMad);
}

void mpdaf_median_sigma_clip(double* data, int n, double x[3], int nmax, double nclip_low,
                             double nclip_up, int nstop, int* indx)
{
    iterative_clip(data, n, x, nmax, nclip_low, nclip_up, nstop, indx, Centre::MedianSigma);
}

int mpdaf_locate(double* xx, int n, double x)
{
    // largest j with xx[j] <= x by bisection, clamped to [0, n-2] (tools.c:204-219)
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (hi + lo) >> 1;
        if (x >= xx[mid]) lo = mid;
        else hi = mid;
    }
    return std::max(0, std::min(n - 2, lo));
}

double mpdaf_linear_interpolation(double* xx, double* yy, int n, double x)
{
    const int j = mpdaf_locate(xx, n, x);
    const double slope = (yy[j + 1] - yy[j]) / (xx[j + 1] - xx[j]);
    const double intercept = -slope * xx[j] + yy[j];
    return slope * x + intercept;
}

void mpdaf_minmax(double data[], int n, int* indx, double res[]) { extremes(data, n, indx, res); }

void mpdaf_minmax_int(int data[], int n, int* indx, int res[]) { extremes(data, n, indx, res); }

}  // extern "C"
