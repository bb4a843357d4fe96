// C ABI of the B200 `_ctools` replacement (see include/mpdaf_b200.h).
//
// Host side of the hot path: argument checking, kernel selection, the slab
// pipeline for host-resident stacks and for FITS files, counter read-back.
// Everything computational happens in the kernels of combine_kernels.cuh; there
// is no CPU fallback -- a missing device is an error.
#include <algorithm>
#include <atomic>
#include <climits>
#include <condition_variable>
#include <functional>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "launch.h"
#include "fits_reader.h"
#include "synth_core.h"

namespace {

using namespace mb;

thread_local std::string g_error;
std::mutex g_call_mutex;              // one merging call at a time (ctools.py contract)
double g_last_kernel_ms = 0.0;
unsigned long long g_last_hist[kGenericMax + 1];
bool g_hist_enabled = true;
int g_last_launches = 0;
std::string g_last_kernel = "none";

}  // namespace

int mb::fail(int code, const std::string& msg)
{
    g_error = msg;
    std::fprintf(stderr, "mpdaf_b200: %s\n", msg.c_str());
    return code;
}

namespace {
using mb::fail;

// ---------------------------------------------------------------------------
// per-device state
// ---------------------------------------------------------------------------
struct DeviceState {
    bool ready = false;
    int sm_count = 0;
    unsigned long long* counters = nullptr;  // [2][kGenericMax]: valid, select
    StackTable* table = nullptr;             // device copy for the generic kernels
    unsigned long long* slow = nullptr;      // voxels recomputed by the exact fallback (diagnostic)
    unsigned* slow_q = nullptr;              // queue of voxels the pair kernel hands to the exact kernel
    unsigned* slow_n = nullptr;
    unsigned long long* hist = nullptr;      // [kGenericMax + 1] exposure-map histogram of the last call
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
constexpr unsigned kSlowQueueCap = 4u << 20;   // voxels; beyond that the pair kernel recomputes in place
DeviceState g_dev[16];

int device_ready(int device)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
        return fail(MPDAF_B200_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= count || device >= 16) return fail(MPDAF_B200_ERR_ARG, "bad device index");
    MB_CUDA(cudaSetDevice(device));
    DeviceState& d = g_dev[device];
    if (d.ready) return MPDAF_B200_OK;
    cudaDeviceProp prop;
    MB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(MPDAF_B200_ERR_CUDA, "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");
    d.sm_count = prop.multiProcessorCount;
    // tuning hook: MPDAF_B200_L2_FETCH = 32 | 64 | 128 sets the L2 fetch granularity hint (default: the driver's, 64)
    if (const char* g = std::getenv("MPDAF_B200_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)std::atoi(g));
    MB_CUDA(cudaMalloc(&d.counters, 2 * kGenericMax * sizeof(unsigned long long)));
    MB_CUDA(cudaMalloc(&d.table, sizeof(StackTable)));
    MB_CUDA(cudaMalloc(&d.slow, sizeof(unsigned long long)));
    MB_CUDA(cudaMemset(d.slow, 0, sizeof(unsigned long long)));
    MB_CUDA(cudaMalloc(&d.slow_q, (size_t)kSlowQueueCap * sizeof(unsigned)));
    MB_CUDA(cudaMalloc(&d.slow_n, sizeof(unsigned)));
    MB_CUDA(cudaMalloc(&d.hist, (kGenericMax + 1) * sizeof(unsigned long long)));
    MB_CUDA(cudaEventCreate(&d.ev0));
    MB_CUDA(cudaEventCreate(&d.ev1));
    d.ready = true;
    return MPDAF_B200_OK;
}

// ---------------------------------------------------------------------------
// kernel selection
// ---------------------------------------------------------------------------
int bucket_of(int nfiles)
{
    const int sizes[] = {4, 8, 12, 16, 24, 32, 48, 64, 96, 128};
    for (int s : sizes)
        if (nfiles <= s) return s;
    return 0;
}

template <typename Kernel, typename... Args>
int launch_persistent(Kernel kernel, int device, long long nvox, cudaStream_t stream, Args... args)
{
    int per_sm = 0;
    MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0));
    if (per_sm < 1) per_sm = 1;
    const long long want = (nvox + kThreads - 1) / kThreads;
    const long long cap = (long long)g_dev[device].sm_count * per_sm;
    const int grid = (int)std::max(1LL, std::min(want, cap));
    kernel<<<grid, kThreads, 0, stream>>>(args...);
    MB_CUDA(cudaGetLastError());
    g_last_launches++;
    return MPDAF_B200_OK;
}

// Uploads the stack description for kernels that read it from device memory.  Ordered on
// `stream`; no pinned bounce buffer is needed because cudaMemcpyAsync from pageable memory
// returns after staging the source.
int upload_table(const Job& j, cudaStream_t stream)
{
    static StackTable host;  // guarded by g_call_mutex
    std::memset(&host, 0, sizeof host);
    for (int i = 0; i < j.nfiles; i++) {
        host.data[i] = j.data[i];
        host.var[i] = j.var ? j.var[i] : nullptr;
        host.scale[i] = j.scale ? j.scale[i] : 1.0;
        host.offset[i] = j.offset ? j.offset[i] : 0.0;
        host.scale2[i] = host.scale[i] * host.scale[i];
    }
    MB_CUDA(cudaMemcpyAsync(j.table_dev, &host, sizeof host, cudaMemcpyHostToDevice, stream));
    return MPDAF_B200_OK;
}

template <int MAXN, typename ElemT>
int launch_generic(const Job& j, int device, cudaStream_t stream)
{
    {
        const int rc = upload_table(j, stream);
        if (rc != MPDAF_B200_OK) return rc;
    }
    char name[64];
    std::snprintf(name, sizeof name, "%s_generic<%d,f%d>", j.median ? "median" : "sigclip", MAXN,
                  (int)sizeof(ElemT) * 8);
    g_last_kernel = name;
    if (j.median)
        return launch_persistent(median_generic_kernel<MAXN, ElemT>, device, j.a.nvox, stream,
                                 (const StackTable*)j.table_dev, j.a);
    return launch_persistent(sigclip_generic_kernel<MAXN, ElemT>, device, j.a.nvox, stream,
                             (const StackTable*)j.table_dev, j.a);
}

int enqueue_combine_only(const Job& j, int device, cudaStream_t stream, bool* hist_fused = nullptr);

// Enqueue the combination of one device-resident stack (or slab) on `stream`, followed by the
// exposure-map histogram of that slab (expnb without a host pass).
int enqueue_combine(const Job& j, int device, cudaStream_t stream)
{
    bool fused = false;      // the pair kernels count the exposure map themselves
    const int rc = enqueue_combine_only(j, device, stream, &fused);
    if (rc != MPDAF_B200_OK || !g_hist_enabled || j.a.nvox <= 0 || fused) return rc;
    const int grid = (int)std::min<long long>((j.a.nvox + 255) / 256, (long long)g_dev[device].sm_count * 8);
    expmap_hist_kernel<<<grid, 256, 0, stream>>>(j.a.expmap, j.a.nvox, g_dev[device].hist);
    MB_CUDA(cudaGetLastError());
    g_last_launches++;
    return MPDAF_B200_OK;
}

int enqueue_combine_only(const Job& j, int device, cudaStream_t stream, bool* hist_fused)
{
    unsigned long long* const hist = (hist_fused != nullptr && g_hist_enabled) ? g_dev[device].hist : nullptr;
    bool unit = true;
    for (int i = 0; i < j.nfiles; i++) {
        if (j.scale && j.scale[i] != 1.0) unit = false;
        if (j.offset && j.offset[i] != 0.0) unit = false;
    }
    const int nb = bucket_of(j.nfiles);
    // float32 stacks up to 128 deep: the pair-packed kernel where the layout allows it (warp_inst.cuh decides), else
    // the warp-tiled kernel; one translation unit per depth bucket
    if (nb != 0 && j.elem == MPDAF_B200_F32) {
        char name[64] = "";
        LaunchCtx ctx;
        ctx.sm_count = g_dev[device].sm_count;
        ctx.slow = g_dev[device].slow;
        ctx.slow_q = j.slow_q ? j.slow_q : g_dev[device].slow_q;     // slab slots bring their own queue (one per stream)
        ctx.slow_n = j.slow_q ? j.slow_n : g_dev[device].slow_n;
        ctx.slow_cap = j.slow_q ? j.slow_cap : kSlowQueueCap;
        // tests shrink the queue to exercise the overflow route (voxels recomputed in place by the main kernel)
        if (const char* cap = std::getenv("MPDAF_B200_SLOW_CAP")) ctx.slow_cap = std::min<unsigned>(ctx.slow_cap, (unsigned)std::atoi(cap));
        ctx.stream = stream;
        ctx.name = name;
        ctx.name_len = sizeof name;
        ctx.hist = hist;
        int rc = MPDAF_B200_ERR_ARG;
        switch (nb) {
#define MB_CASE(NB) case NB: rc = j.median ? mb_launch_median_warp_##NB(j, ctx) : mb_launch_warp_##NB(j, unit, ctx); break;
            MB_FOR_EACH_DEPTH(MB_CASE)
#undef MB_CASE
        }
        g_last_kernel = name;
        g_last_launches += ctx.launches;
        if (hist_fused != nullptr) *hist_fused = ctx.hist_fused && rc == MPDAF_B200_OK;
        return rc;
    }
    // 129..512 float32 exposures, plain sigma clip or median: R sorted runs of 64 (sigclip_runs.cuh)
    if (j.nfiles > 128 && j.nfiles <= kGenericMax && j.elem == MPDAF_B200_F32 && (j.median || !j.a.mad) &&
        std::getenv("MPDAF_B200_NO_RUNS") == nullptr) {
        char name[64] = "";
        LaunchCtx ctx;
        ctx.sm_count = g_dev[device].sm_count;
        ctx.slow = g_dev[device].slow;
        ctx.slow_q = j.slow_q ? j.slow_q : g_dev[device].slow_q;     // slab slots bring their own queue (one per stream)
        ctx.slow_n = j.slow_q ? j.slow_n : g_dev[device].slow_n;
        ctx.slow_cap = j.slow_q ? j.slow_cap : kSlowQueueCap;
        // tests shrink the queue to exercise the overflow route (voxels recomputed in place by the main kernel)
        if (const char* cap = std::getenv("MPDAF_B200_SLOW_CAP")) ctx.slow_cap = std::min<unsigned>(ctx.slow_cap, (unsigned)std::atoi(cap));
        ctx.stream = stream;
        ctx.name = name;
        ctx.name_len = sizeof name;
        int rc = upload_table(j, stream);
        if (rc == MPDAF_B200_OK) rc = mb_launch_runs(j, unit, ctx);
        g_last_kernel = name;
        g_last_launches += ctx.launches;
        return rc;
    }
    // float64 samples, sigma clip, up to 64 exposures on equally pitched rows: the pair kernel with float64 stages
    if (j.elem == MPDAF_B200_F64 && !j.median && j.nfiles <= 64) {
        char name[64] = "";
        LaunchCtx ctx;
        ctx.sm_count = g_dev[device].sm_count;
        ctx.slow = g_dev[device].slow;
        ctx.slow_q = j.slow_q ? j.slow_q : g_dev[device].slow_q;
        ctx.slow_n = j.slow_q ? j.slow_n : g_dev[device].slow_n;
        ctx.slow_cap = j.slow_q ? j.slow_cap : kSlowQueueCap;
        if (const char* cap = std::getenv("MPDAF_B200_SLOW_CAP")) ctx.slow_cap = std::min<unsigned>(ctx.slow_cap, (unsigned)std::atoi(cap));
        ctx.stream = stream;
        ctx.name = name;
        ctx.name_len = sizeof name;
        ctx.hist = hist;
        bool taken = false;
        const int rc = j.nfiles <= 12 ? mb_launch_pair64_12(j, unit, ctx, taken)
                     : j.nfiles <= 24 ? mb_launch_pair64_24(j, unit, ctx, taken)
                     : j.nfiles <= 48 ? mb_launch_pair64_48(j, unit, ctx, taken) : mb_launch_pair64_64(j, unit, ctx, taken);
        if (rc != MPDAF_B200_OK || taken) {
            g_last_kernel = name;
            g_last_launches += ctx.launches;
            if (hist_fused != nullptr) *hist_fused = ctx.hist_fused && rc == MPDAF_B200_OK;
            return rc;
        }
    }
    if (j.nfiles <= 128) {
        if (j.elem == MPDAF_B200_F32) return launch_generic<128, float>(j, device, stream);
        return launch_generic<128, double>(j, device, stream);
    }
    if (j.elem == MPDAF_B200_F32) return launch_generic<kGenericMax, float>(j, device, stream);
    return launch_generic<kGenericMax, double>(j, device, stream);
}

int check_common(int nfiles, const void* const* data, int elem_type, long long nvox, int location)
{
    if (nfiles < 1 || nfiles > MPDAF_B200_MAX_FILES) return fail(MPDAF_B200_ERR_ARG, "nfiles must be in 1..500");
    if (elem_type != MPDAF_B200_F32 && elem_type != MPDAF_B200_F64) return fail(MPDAF_B200_ERR_ARG, "elem_type must be 4 or 8");
    if (nvox < 0 || data == nullptr) return fail(MPDAF_B200_ERR_ARG, "bad stack description");
    if (location != MPDAF_B200_HOST && location != MPDAF_B200_DEVICE) return fail(MPDAF_B200_ERR_ARG, "bad location");
    for (int i = 0; i < nfiles; i++)
        if (data[i] == nullptr && nvox > 0) return fail(MPDAF_B200_ERR_ARG, "null exposure pointer");
    return MPDAF_B200_OK;
}

// ---------------------------------------------------------------------------
// slab pipeline: host-resident stacks and FITS files
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// persistent host workers: file reads / gathers of pageable rows into pinned staging
// ---------------------------------------------------------------------------
// Created once, on first use, and kept for the life of the process (a drop-in library inside a long-lived
// Python process used to create and join up to 32 threads per chunk).
class WorkerPool {
public:
    explicit WorkerPool(int n)
    {
        for (int t = 0; t < n; t++) threads_.emplace_back([this] { loop(); });
    }
    int size() const { return (int)threads_.size(); }
    // runs fn(0) .. fn(njobs - 1) on the workers (at most `width` of them at a time) and waits for all of them
    void run(int njobs, int width, const std::function<void(int)>& fn)
    {
        if (njobs <= 0) return;
        std::unique_lock<std::mutex> lk(m_);
        fn_ = &fn;
        njobs_ = njobs;
        next_ = 0;
        done_ = 0;
        width_ = std::max(1, std::min(width, size()));
        generation_++;
        cv_work_.notify_all();
        cv_done_.wait(lk, [this] { return done_ == njobs_; });
        fn_ = nullptr;
    }

private:
    void loop()
    {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_work_.wait(lk, [&] { return generation_ != seen && fn_ != nullptr && next_ < njobs_ && active_ < width_; });
            const unsigned long long gen = generation_;
            active_++;
            while (fn_ != nullptr && generation_ == gen && next_ < njobs_) {
                const int job = next_++;
                const std::function<void(int)>* fn = fn_;
                lk.unlock();
                (*fn)(job);
                lk.lock();
                if (++done_ == njobs_) cv_done_.notify_all();
            }
            active_--;
            if (generation_ == gen) seen = gen;
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)>* fn_ = nullptr;
    int njobs_ = 0, next_ = 0, done_ = 0, width_ = 1, active_ = 0;
    unsigned long long generation_ = 0;
};

WorkerPool& worker_pool()
{
    static WorkerPool* pool = [] {
        const long long hw = std::max(1u, std::thread::hardware_concurrency());
        const char* v = std::getenv("MPDAF_B200_IO_THREADS");
        const long long want = v ? std::atoll(v) : std::min(32LL, hw);
        return new WorkerPool((int)std::max(1LL, std::min(64LL, want)));   // never destroyed: its threads end with the process
    }();
    return *pool;
}

// devices a HOST-location / file-name call spreads its wavelength slabs over (mpdaf_b200_set_devices, MPDAF_B200_DEVICES)
std::vector<int> g_devices;

std::vector<int> usable_devices()
{
    std::vector<int> ids;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return ids;
    for (int i = 0; i < count && i < 16; i++) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ids.push_back(i);
    }
    return ids;
}

// "all" | a count | a comma-separated list of device indices
std::vector<int> parse_devices(const char* spec)
{
    std::vector<int> ids;
    const std::vector<int> all = usable_devices();
    if (spec == nullptr || *spec == 0) return ids;
    const std::string sp(spec);
    if (sp == "all") return all;
    if (sp.find(',') == std::string::npos) {
        const int n = std::atoi(sp.c_str());
        for (int i = 0; i < n && i < (int)all.size(); i++) ids.push_back(all[i]);
        return ids;
    }
    size_t pos = 0;
    while (pos < sp.size()) {
        const size_t e = sp.find(',', pos);
        const std::string tok = sp.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
        if (!tok.empty()) ids.push_back(std::atoi(tok.c_str()));
        if (e == std::string::npos) break;
        pos = e + 1;
    }
    return ids;
}

struct Slot {
    cudaStream_t stream = nullptr;
    char* d_in = nullptr;      // [planes][chunk] samples (data planes then var planes)
    double* d_out = nullptr;
    double* d_var = nullptr;
    int* d_exp = nullptr;
    char* h_in = nullptr;      // pinned staging for file reads
    char* h_out = nullptr;     // pinned staging for results bound for pageable caller buffers: [data f64][var f64][expmap i32]
    size_t h_out_cap = 0;      // voxels
    long long pend_v0 = 0, pend_cnt = 0;   // chunk waiting in h_out for its copy into the caller's buffers
    bool pending = false;
    StackTable* d_table = nullptr;
    unsigned* slow_q = nullptr;   // this stream's queue of voxels for the exact kernel (sigclip_pair.cuh)
    unsigned* slow_n = nullptr;
    cudaEvent_t k0 = nullptr, k1 = nullptr;
    bool timed = false;
    size_t in_bytes = 0, chunk_cap = 0, host_bytes = 0;   // current capacities
};

// The staging slots are kept per device between calls (allocation, pinning and stream creation
// cost more than a whole small combine); they only ever grow.
constexpr int kSlots = 3;
constexpr unsigned kSlotQueueCap = 1u << 20;
Slot g_slots[16][kSlots];

int ensure_slot(Slot& s, size_t in_bytes, size_t chunk, bool host_staging)
{
    if (!s.stream) {
        MB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        MB_CUDA(cudaMalloc(&s.d_table, sizeof(StackTable)));
        MB_CUDA(cudaMalloc(&s.slow_q, (size_t)kSlotQueueCap * sizeof(unsigned)));
        MB_CUDA(cudaMalloc(&s.slow_n, sizeof(unsigned)));
        MB_CUDA(cudaEventCreate(&s.k0));
        MB_CUDA(cudaEventCreate(&s.k1));
    }
    if (in_bytes > s.in_bytes) {
        cudaFree(s.d_in);
        s.in_bytes = 0;
        MB_CUDA(cudaMalloc(&s.d_in, in_bytes));
        s.in_bytes = in_bytes;
    }
    if (chunk > s.chunk_cap) {
        cudaFree(s.d_out);
        cudaFree(s.d_var);
        cudaFree(s.d_exp);
        s.chunk_cap = 0;
        MB_CUDA(cudaMalloc(&s.d_out, chunk * sizeof(double)));
        MB_CUDA(cudaMalloc(&s.d_var, chunk * sizeof(double)));
        MB_CUDA(cudaMalloc(&s.d_exp, chunk * sizeof(int)));
        s.chunk_cap = chunk;
    }
    if (host_staging && in_bytes > s.host_bytes) {
        if (s.h_in) cudaFreeHost(s.h_in);
        s.host_bytes = 0;
        MB_CUDA(cudaHostAlloc((void**)&s.h_in, in_bytes, cudaHostAllocPortable));
        s.host_bytes = in_bytes;
    }
    s.timed = false;
    s.pending = false;
    return MPDAF_B200_OK;
}

int ensure_out_staging(Slot& s, size_t chunk)
{
    if (chunk > s.h_out_cap) {
        if (s.h_out) cudaFreeHost(s.h_out);
        s.h_out_cap = 0;
        MB_CUDA(cudaHostAlloc((void**)&s.h_out, chunk * (2 * sizeof(double) + sizeof(int)), cudaHostAllocPortable));
        s.h_out_cap = chunk;
    }
    return MPDAF_B200_OK;
}

// true when the pointer is page-locked host memory the copy engines can write asynchronously
bool is_pinned_host(const void* p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

struct SlabSource {
    // exactly one of the two is used
    const void* const* host_data = nullptr;
    const void* const* host_var = nullptr;
    std::vector<mbfits::ImageHdu>* fits_data = nullptr;
    std::vector<mbfits::ImageHdu>* fits_var = nullptr;
};

long long env_ll(const char* name, long long dflt)
{
    const char* v = std::getenv(name);
    return v ? std::atoll(v) : dflt;
}

// Runs the whole stack through the devices in chunks of consecutive voxels (wavelength slabs, merging.c:112-137:
// the reference deals plane ranges to its OpenMP threads the same way).  Chunk c goes to device c % ndev; every
// device has its own slots and streams, results are copied straight to the chunk's offset in the caller's arrays,
// and the per-exposure counters of the devices are summed by run_job.  No data moves between devices.
int run_slabs(const SlabSource& src, Job job, const std::vector<int>& devs, double* out, double* outvar, int* expmap)
{
    const int ndev = (int)devs.size();
    const int nfiles = job.nfiles;
    const bool need_var = !job.median && job.a.typ_var == 0;
    const int planes = nfiles * (need_var ? 2 : 1);
    const int esz = job.elem;
    const long long nvox = job.a.nvox;
    const bool from_files = src.fits_data != nullptr;

    // chunk: bounded staging per slot (default 128 MiB of samples: small enough that the copy
    // of one chunk overlaps the kernel of the previous one even for short stacks, large enough
    // that every per-exposure segment is a ~MB-sized copy), multiple of 1024 voxels
    const long long budget = env_ll("MPDAF_B200_SLAB_BYTES", from_files ? 128LL << 20 : 256LL << 20);
    long long chunk = budget / ((long long)planes * esz);
    chunk = std::max(1024LL, chunk / 1024 * 1024);
    chunk = std::min(chunk, (nvox + 1023) / 1024 * 1024);
    if (chunk < 1) chunk = 1024;
    const long long nchunks = (nvox + chunk - 1) / chunk;
    const int nslots = (int)std::max<long long>(1, std::min<long long>(kSlots, (nchunks + ndev - 1) / ndev));

    // pageable host arrays (plain NumPy) are gathered into the slot's pinned staging by worker threads, like file
    // reads: a copy engine reading pageable memory goes through the driver's own single staging buffer and blocks
    const bool stage_in = from_files || !(is_pinned_host(src.host_data[0]) && (!need_var || is_pinned_host(src.host_var[0])));
    for (int dev : devs) {
        MB_CUDA(cudaSetDevice(dev));
        for (int k = 0; k < nslots; k++) {
            const int rc0 = ensure_slot(g_slots[dev][k], (size_t)planes * chunk * esz, (size_t)chunk, stage_in);
            if (rc0 != MPDAF_B200_OK) return rc0;
        }
    }

    // byte distance between consecutive host rows when it is the same for all of them (else 0): [0] DATA, [1] STAT
    long long host_pitch[2] = {0, 0};
    if (!from_files) {
        auto pitch_of = [&](const void* const* rows) -> long long {
            if (rows == nullptr || nfiles < 2) return 0;
            const long long step = (const char*)rows[1] - (const char*)rows[0];
            if (step < nvox * (long long)esz || step > 0x7FFFFFFFLL) return 0;      // cudaMemcpy2D pitches are below 2 GiB
            for (int i = 2; i < nfiles; i++)
                if ((const char*)rows[i] - (const char*)rows[i - 1] != step) return 0;
            return step;
        };
        host_pitch[0] = pitch_of(src.host_data);
        host_pitch[1] = need_var ? pitch_of(src.host_var) : 0;
    }

    // Results bound for pageable buffers (plain NumPy arrays: what ctools.py passes) go through pinned staging
    // and are copied out by this thread when the slot comes round again: a D2H copy straight into pageable
    // memory blocks the thread until the chunk's kernel has finished, which serialises read / H2D / kernel / D2H
    const bool stage_out = !(is_pinned_host(out) && (job.median || is_pinned_host(outvar)) && is_pinned_host(expmap));
    if (stage_out)
        for (int dev : devs) {
            MB_CUDA(cudaSetDevice(dev));
            for (int k = 0; k < nslots; k++) {
                const int rc0 = ensure_out_staging(g_slots[dev][k], (size_t)chunk);
                if (rc0 != MPDAF_B200_OK) return rc0;
            }
        }
    auto drain = [&](Slot& s) {
        if (!s.pending) return;
        const size_t cap = s.h_out_cap;
        const double* hd = reinterpret_cast<const double*>(s.h_out);
        std::memcpy(out + s.pend_v0, hd, (size_t)s.pend_cnt * sizeof(double));
        if (!job.median) std::memcpy(outvar + s.pend_v0, hd + cap, (size_t)s.pend_cnt * sizeof(double));
        std::memcpy(expmap + s.pend_v0, s.h_out + 2 * cap * sizeof(double), (size_t)s.pend_cnt * sizeof(int));
        s.pending = false;
    };

    int rc = MPDAF_B200_OK;
    double kernel_ms = 0.0;
    std::vector<const void*> dptr(planes);
    long long c = 0;
    for (long long v0 = 0; v0 < nvox && rc == MPDAF_B200_OK; v0 += chunk, c++) {
        const int device = devs[c % ndev];
        Slot& s = g_slots[device][(c / ndev) % nslots];
        const long long cnt = std::min(chunk, nvox - v0);
        if (cudaSetDevice(device) != cudaSuccess) { rc = fail(MPDAF_B200_ERR_CUDA, "cudaSetDevice failed"); break; }
        // the slot's previous chunk must have left the device (and its pinned staging)
        if (cudaStreamSynchronize(s.stream) != cudaSuccess) { rc = fail(MPDAF_B200_ERR_CUDA, "slab stream failed"); break; }
        drain(s);
        if (s.timed) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, s.k0, s.k1);
            kernel_ms += ms;
            s.timed = false;
        }
        for (int p = 0; p < planes; p++) dptr[p] = s.d_in + (size_t)p * chunk * esz;

        if (stage_in) {
            // parallel pread (files) or memcpy (pageable arrays) of every exposure's range into pinned staging
            std::atomic<int> bad{0};
            std::string first_err;
            std::mutex err_mutex;
            const std::function<void(int)> stage_plane = [&](int p) {
                if (!from_files) {
                    const char* base = (const char*)(p < nfiles ? src.host_data[p] : src.host_var[p - nfiles]);
                    std::memcpy(s.h_in + (size_t)p * chunk * esz, base + (size_t)v0 * esz, (size_t)cnt * esz);
                    return;
                }
                const mbfits::ImageHdu& h = p < nfiles ? (*src.fits_data)[p] : (*src.fits_var)[p - nfiles];
                std::string err;
                if (mbfits::read_raw(h, v0, cnt, s.h_in + (size_t)p * chunk * esz, &err) != MPDAF_B200_OK) {
                    std::lock_guard<std::mutex> lk(err_mutex);
                    if (!bad.exchange(1)) first_err = err;
                }
            };
            worker_pool().run(planes, planes, stage_plane);
            if (bad.load()) { rc = fail(MPDAF_B200_ERR_IO, first_err); break; }
            if (cudaMemcpyAsync(s.d_in, s.h_in, (size_t)planes * chunk * esz, cudaMemcpyHostToDevice, s.stream) != cudaSuccess) {
                rc = fail(MPDAF_B200_ERR_CUDA, "H2D of a staged slab failed");
                break;
            }
            if (from_files) {
                const long long words = (long long)planes * chunk * (esz / 4);
                // FITS is big-endian: swap 32-bit words (for f64 the word pair is swapped by the 64-bit variant)
                if (esz == 4) {
                    bswap32_kernel<<<1024, 256, 0, s.stream>>>((uint32_t*)s.d_in, words);
                } else {
                    bswap64_kernel<<<1024, 256, 0, s.stream>>>((unsigned long long*)s.d_in, (long long)planes * chunk);
                }
                g_last_launches++;
            }
        } else {
            // equally spaced host rows (one [N][V] array): ONE 2-D copy per chunk for DATA and one for STAT instead
            // of a copy per exposure (measured 45.6 -> 50+ GB/s at 128 MiB chunks; the box's plain H2D rate is 55.5)
            auto copy_rows = [&](const void* const* rows, int first_plane) -> bool {
                const long long pitch = host_pitch[first_plane ? 1 : 0];
                if (pitch > 0) {
                    return cudaMemcpy2DAsync(s.d_in + (size_t)first_plane * chunk * esz, (size_t)chunk * esz,
                                             (const char*)rows[0] + (size_t)v0 * esz, (size_t)pitch, (size_t)cnt * esz,
                                             (size_t)nfiles, cudaMemcpyHostToDevice, s.stream) == cudaSuccess;
                }
                for (int p = 0; p < nfiles; p++)
                    if (cudaMemcpyAsync(s.d_in + (size_t)(first_plane + p) * chunk * esz, (const char*)rows[p] + (size_t)v0 * esz,
                                        (size_t)cnt * esz, cudaMemcpyHostToDevice, s.stream) != cudaSuccess)
                        return false;
                return true;
            };
            if (!copy_rows(src.host_data, 0) || (need_var && !copy_rows(src.host_var, nfiles)))
                rc = fail(MPDAF_B200_ERR_CUDA, "H2D of a host slab failed");
            if (rc != MPDAF_B200_OK) break;
        }

        Job j = job;
        j.data = dptr.data();
        j.var = need_var ? dptr.data() + nfiles : nullptr;
        j.table_dev = s.d_table;
        j.slow_q = s.slow_q;
        j.slow_n = s.slow_n;
        j.slow_cap = kSlotQueueCap;
        j.a.valid_cnt = g_dev[device].counters;
        j.a.select_cnt = job.median ? nullptr : g_dev[device].counters + kGenericMax;
        j.a.nvox = cnt;
        j.a.out = s.d_out;
        j.a.outvar = s.d_var;
        j.a.expmap = s.d_exp;
        cudaEventRecord(s.k0, s.stream);
        rc = enqueue_combine(j, device, s.stream);
        if (rc != MPDAF_B200_OK) break;
        cudaEventRecord(s.k1, s.stream);
        s.timed = true;
        double* to_data = stage_out ? reinterpret_cast<double*>(s.h_out) : out + v0;
        double* to_var = stage_out ? reinterpret_cast<double*>(s.h_out) + s.h_out_cap : outvar + v0;
        int* to_exp = stage_out ? reinterpret_cast<int*>(s.h_out + 2 * s.h_out_cap * sizeof(double)) : expmap + v0;
        bool ok = cudaMemcpyAsync(to_data, s.d_out, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, s.stream) == cudaSuccess;
        if (!job.median)
            ok = ok && cudaMemcpyAsync(to_var, s.d_var, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, s.stream) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(to_exp, s.d_exp, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, s.stream) == cudaSuccess;
        if (!ok) rc = fail(MPDAF_B200_ERR_CUDA, "D2H of a result slab failed");
        if (stage_out && ok) {
            s.pend_v0 = v0;
            s.pend_cnt = cnt;
            s.pending = true;
        }
    }
    for (int dev : devs) {
        cudaSetDevice(dev);
        for (int k = 0; k < nslots; k++) {
            Slot& s = g_slots[dev][k];
            if (!s.stream) continue;
            if (cudaStreamSynchronize(s.stream) != cudaSuccess && rc == MPDAF_B200_OK)
                rc = fail(MPDAF_B200_ERR_CUDA, std::string("slab pipeline: ") + cudaGetErrorString(cudaGetLastError()));
            if (rc == MPDAF_B200_OK) drain(s);
            s.pending = false;
            if (s.timed) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, s.k0, s.k1);
                kernel_ms += ms;
                s.timed = false;
            }
        }
    }
    g_last_kernel_ms = kernel_ms;
    return rc;
}

// ---------------------------------------------------------------------------
// shared driver for all merging entry points
// ---------------------------------------------------------------------------
struct CounterSink {
    int* valid32 = nullptr;
    int* select32 = nullptr;
    long long* valid64 = nullptr;
    long long* select64 = nullptr;
};

// restores the caller's current device when a call returns (torch and other CUDA users share the thread's context)
struct DeviceGuard {
    int saved = -1;
    DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) { saved = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

// devices of a HOST-location call: `device` >= 0 names one; MPDAF_B200_ALL_DEVICES (-1) means the configured set
// (mpdaf_b200_set_devices, else the environment's MPDAF_B200_DEVICES, else every usable device)
std::vector<int> devices_for(int device)
{
    if (device >= 0) return {device};
    if (!g_devices.empty()) return g_devices;
    std::vector<int> ids = parse_devices(std::getenv("MPDAF_B200_DEVICES"));
    if (ids.empty()) ids = usable_devices();
    return ids;
}

int run_job(Job job, const SlabSource* slab_src, int location, int device, double* out, double* outvar,
            int* expmap, CounterSink sink, cudaStream_t stream)
{
    std::lock_guard<std::mutex> lock(g_call_mutex);
    DeviceGuard guard;
    g_last_launches = 0;
    g_last_kernel_ms = 0.0;
    std::vector<int> devs = location == MPDAF_B200_DEVICE ? std::vector<int>{device} : devices_for(device);
    if (devs.empty()) return fail(MPDAF_B200_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (location == MPDAF_B200_DEVICE && device < 0) return fail(MPDAF_B200_ERR_ARG, "a DEVICE-location call needs a device index");
    // no more devices than chunks of work
    const int nfiles = job.nfiles;
    static const bool hist_off = std::getenv("MPDAF_B200_NO_HISTOGRAM") != nullptr;
    g_hist_enabled = !hist_off;
    cudaStream_t cstream = (location == MPDAF_B200_DEVICE) ? stream : nullptr;
    for (int dev : devs) {
        const int rc = device_ready(dev);
        if (rc != MPDAF_B200_OK) return rc;
        DeviceState& d = g_dev[dev];
        MB_CUDA(cudaMemsetAsync(d.counters, 0, 2 * kGenericMax * sizeof(unsigned long long), cstream));
        MB_CUDA(cudaMemsetAsync(d.hist, 0, (kGenericMax + 1) * sizeof(unsigned long long), cstream));
        if (location != MPDAF_B200_DEVICE) MB_CUDA(cudaStreamSynchronize(cstream));   // counters zeroed before the slab streams start
    }
    int rc = MPDAF_B200_OK;
    if (location == MPDAF_B200_DEVICE) {
        DeviceState& d = g_dev[device];
        job.a.valid_cnt = d.counters;
        job.a.select_cnt = job.median ? nullptr : d.counters + kGenericMax;
        job.table_dev = d.table;
        job.a.out = out;
        job.a.outvar = outvar;
        job.a.expmap = expmap;
        if (job.a.nvox > 0) {
            MB_CUDA(cudaEventRecord(d.ev0, stream));
            rc = enqueue_combine(job, device, stream);
            if (rc != MPDAF_B200_OK) return rc;
            MB_CUDA(cudaEventRecord(d.ev1, stream));
        }
    } else if (job.a.nvox > 0) {
        rc = run_slabs(*slab_src, job, devs, out, outvar, expmap);
        if (rc != MPDAF_B200_OK) return rc;
    }
    // per-exposure counters and the exposure-map histogram: summed over the devices (the path's only reduction)
    std::vector<unsigned long long> total(2 * kGenericMax, 0ull), part(2 * kGenericMax);
    std::vector<unsigned long long> hist_part(kGenericMax + 1);
    for (int i = 0; i <= kGenericMax; i++) g_last_hist[i] = 0ull;
    for (int dev : devs) {
        MB_CUDA(cudaSetDevice(dev));
        DeviceState& d = g_dev[dev];
        MB_CUDA(cudaMemcpyAsync(part.data(), d.counters, part.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, cstream));
        MB_CUDA(cudaMemcpyAsync(hist_part.data(), d.hist, hist_part.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, cstream));
        MB_CUDA(cudaStreamSynchronize(cstream));
        for (size_t i = 0; i < total.size(); i++) total[i] += part[i];
        for (int i = 0; i <= kGenericMax; i++) g_last_hist[i] += hist_part[i];
    }
    if (location == MPDAF_B200_DEVICE && job.a.nvox > 0) {
        float ms = 0.f;
        MB_CUDA(cudaEventElapsedTime(&ms, g_dev[device].ev0, g_dev[device].ev1));
        g_last_kernel_ms = ms;
    }
    for (int i = 0; i < nfiles; i++) {
        const unsigned long long v = total[i], sl = total[kGenericMax + i];
        if (sink.valid32) sink.valid32[i] += (int)v;
        if (sink.valid64) sink.valid64[i] += (long long)v;
        if (!job.median) {
            if (sink.select32) sink.select32[i] += (int)sl;
            if (sink.select64) sink.select64[i] += (long long)sl;
        }
    }
    return MPDAF_B200_OK;
}

// Splits the newline-joined file list (merging.c:39-55) WITHOUT touching the buffer.
std::vector<std::string> split_list(const char* input)
{
    std::vector<std::string> names;
    const char* p = input;
    while (*p) {
        const char* e = std::strchr(p, '\n');
        const size_t len = e ? (size_t)(e - p) : std::strlen(p);
        if (len > 0) names.emplace_back(p, len);
        p += len + (e ? 1 : 0);
    }
    return names;
}

int open_all(const std::vector<std::string>& names, const char* ext, std::vector<mbfits::ImageHdu>& hdus,
             const long long* shape)
{
    std::string err;
    for (const std::string& n : names) {
        mbfits::ImageHdu h;
        const int rc = mbfits::open_image(n, ext, &h, &err);
        if (rc != MPDAF_B200_OK) return fail(rc, err);
        hdus.push_back(h);
        const long long* ref = shape ? shape : hdus[0].naxes;
        if (h.naxes[0] != ref[0] || h.naxes[1] != ref[1] || h.naxes[2] != ref[2])  // merging.c:195-199
            return fail(MPDAF_B200_ERR_SHAPE, n + "[" + ext + "] does not have the same size");
        if (h.bitpix != hdus[0].bitpix) return fail(MPDAF_B200_ERR_SHAPE, n + ": mixed BITPIX in one list");
    }
    return MPDAF_B200_OK;
}

int merge_files_impl(char* input, bool median, double* data, double* var, int* expmap, double* scale, double* offset,
                     int* selected_pix, int* valid_pix, int nmax, double lo, double up, int nstop, int typ_var, int mad);

int merge_files(char* input, bool median, double* data, double* var, int* expmap, double* scale, double* offset,
                int* selected_pix, int* valid_pix, int nmax, double lo, double up, int nstop, int typ_var, int mad)
{
    const int rc = merge_files_impl(input, median, data, var, expmap, scale, offset, selected_pix, valid_pix, nmax, lo, up,
                                    nstop, typ_var, mad);
    // failures before the cube shape is known cannot fill the outputs; MPDAF_B200_EXIT_ON_ERROR restores the
    // reference's behaviour (exit) for callers that never look at the return value
    if (rc != MPDAF_B200_OK && std::getenv("MPDAF_B200_EXIT_ON_ERROR")) std::exit(EXIT_FAILURE);
    return rc;
}

int merge_files_impl(char* input, bool median, double* data, double* var, int* expmap, double* scale, double* offset,
                     int* selected_pix, int* valid_pix, int nmax, double lo, double up, int nstop, int typ_var, int mad)
{
    if (!input || !data || !expmap || !valid_pix) return fail(MPDAF_B200_ERR_ARG, "null argument");
    if (!median && (!var || !selected_pix)) return fail(MPDAF_B200_ERR_ARG, "null argument");
    const std::vector<std::string> names = split_list(input);
    if (names.empty()) return fail(MPDAF_B200_ERR_ARG, "empty file list");
    if ((int)names.size() > MPDAF_B200_MAX_FILES) return fail(MPDAF_B200_ERR_ARG, "too many files, limit is 500");  // merging.c:46
    std::vector<mbfits::ImageHdu> hd, hv;
    struct Closer {
        std::vector<mbfits::ImageHdu>&a, &b;
        ~Closer() { for (auto& h : a) mbfits::close_image(&h); for (auto& h : b) mbfits::close_image(&h); }
    } closer{hd, hv};
    auto fill_failed = [&](int code) {
        // the first cube's shape is what the caller allocated for (cubelist.py:447-452)
        if (!hd.empty()) {
            const long long nv = hd[0].naxes[0] * hd[0].naxes[1] * hd[0].naxes[2];
            const double nan = std::nan("");
            for (long long i = 0; i < nv; i++) data[i] = nan;
            if (var) for (long long i = 0; i < nv; i++) var[i] = nan;
            for (long long i = 0; i < nv; i++) expmap[i] = 0;
        }
        return code;
    };
    int rc = open_all(names, "DATA", hd, nullptr);
    if (rc != MPDAF_B200_OK) return fill_failed(rc);
    if (!median && typ_var == 0) {
        rc = open_all(names, "STAT", hv, hd[0].naxes);
        if (rc != MPDAF_B200_OK) return fill_failed(rc);
        if (hv[0].bitpix != hd[0].bitpix) return fill_failed(fail(MPDAF_B200_ERR_SHAPE, "DATA and STAT differ in BITPIX"));
    }
    Job job;
    job.nfiles = (int)names.size();
    job.elem = hd[0].bitpix == -32 ? MPDAF_B200_F32 : MPDAF_B200_F64;
    job.median = median;
    job.scale = scale;
    job.offset = offset;
    job.a.nfiles = job.nfiles;
    job.a.nvox = hd[0].naxes[0] * hd[0].naxes[1] * hd[0].naxes[2];
    job.a.nmax = nmax;
    job.a.nclip_low = lo;
    job.a.nclip_up = up;
    job.a.nstop = nstop;
    job.a.typ_var = median ? 1 : typ_var;
    job.a.mad = mad;
    SlabSource src;
    src.fits_data = &hd;
    src.fits_var = &hv;
    CounterSink sink;
    sink.valid32 = valid_pix;
    sink.select32 = selected_pix;
    // MPDAF_B200_DEVICE pins the call to one device; otherwise the slabs are spread over every usable device
    // (MPDAF_B200_DEVICES = "all" | count | list narrows that down)
    const int device = std::getenv("MPDAF_B200_DEVICE") ? (int)env_ll("MPDAF_B200_DEVICE", 0) : MPDAF_B200_ALL_DEVICES;
    const long long nvox = job.a.nvox;
    rc = run_job(job, &src, MPDAF_B200_HOST, device, data, var, expmap, sink, nullptr);
    if (rc != MPDAF_B200_OK) {
        // The reference exits the process on any failure (merging.c:99-108,196-199) and its Python caller ignores the
        // return value (cubelist.py:456,512): never leave caller-allocated np.empty buffers half written.
        const double nan = std::nan("");
        for (long long i = 0; i < nvox; i++) data[i] = nan;
        if (var) for (long long i = 0; i < nvox; i++) var[i] = nan;
        for (long long i = 0; i < nvox; i++) expmap[i] = 0;
        if (std::getenv("MPDAF_B200_EXIT_ON_ERROR")) std::exit(EXIT_FAILURE);
    }
    return rc;
}

int merge_arrays(int nfiles, const void* const* data, const void* const* var, int elem_type, long long nvox,
                 int location, int device, bool median, double* out, double* outvar, int* expmap,
                 const double* scale, const double* offset, CounterSink sink, int nmax, double lo, double up,
                 int nstop, int typ_var, int mad, void* stream)
{
    int rc = check_common(nfiles, data, elem_type, nvox, location);
    if (rc != MPDAF_B200_OK) return rc;
    if (nvox > 0 && (!out || !expmap || (!median && !outvar))) return fail(MPDAF_B200_ERR_ARG, "null output pointer");
    if (!median && typ_var == 0) {
        if (!var) return fail(MPDAF_B200_ERR_ARG, "var='propagate' needs the STAT arrays");
        for (int i = 0; i < nfiles; i++)
            if (!var[i] && nvox > 0) return fail(MPDAF_B200_ERR_ARG, "null STAT pointer");
    }
    Job job;
    job.nfiles = nfiles;
    job.elem = elem_type;
    job.median = median;
    job.data = data;
    job.var = (!median && typ_var == 0) ? var : nullptr;
    job.scale = scale;
    job.offset = offset;
    job.a.nfiles = nfiles;
    job.a.nvox = nvox;
    job.a.nmax = nmax;
    job.a.nclip_low = lo;
    job.a.nclip_up = up;
    job.a.nstop = nstop;
    job.a.typ_var = median ? 1 : typ_var;
    job.a.mad = mad;
    SlabSource src;
    src.host_data = data;
    src.host_var = var;
    return run_job(job, &src, location, device, out, outvar, expmap, sink, (cudaStream_t)stream);
}

__global__ void synth_kernel(uint32_t* dst, int kind, unsigned long long seed, int exposure, long long v0,
                             long long count, int nx, int ny)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        dst[i] = kind == 0 ? mbsynth::data_bits(seed, exposure, v0 + i, nx, ny)
                           : mbsynth::stat_bits(seed, exposure, v0 + i);
}

}  // namespace

// ---------------------------------------------------------------------------
// exported symbols
// ---------------------------------------------------------------------------
extern "C" {

int mpdaf_merging_median(char* input, double* data, int* expmap, int* valid_pix)
{
    return merge_files(input, true, data, nullptr, expmap, nullptr, nullptr, nullptr, valid_pix, 0, 0, 0, 1, 1, 0);
}

int mpdaf_merging_sigma_clipping(char* input, double* data, double* var, int* expmap, double* scale,
                                 double* offset, int* selected_pix, int* valid_pix, int nmax,
                                 double nclip_low, double nclip_up, int nstop, int typ_var, int mad)
{
    return merge_files(input, false, data, var, expmap, scale, offset, selected_pix, valid_pix, nmax, nclip_low,
                       nclip_up, nstop, typ_var, mad);
}

int mpdaf_b200_sigma_clipping_arrays(int nfiles, const void* const* data, const void* const* var, int elem_type,
                                     long long nvox, int location, int device, double* out, double* outvar,
                                     int* expmap, const double* scale, const double* offset, int* selected_pix,
                                     int* valid_pix, int nmax, double nclip_low, double nclip_up, int nstop,
                                     int typ_var, int mad, void* stream)
{
    if (!selected_pix || !valid_pix) return fail(MPDAF_B200_ERR_ARG, "null counter pointer");
    CounterSink sink;
    sink.valid32 = valid_pix;
    sink.select32 = selected_pix;
    return merge_arrays(nfiles, data, var, elem_type, nvox, location, device, false, out, outvar, expmap, scale,
                        offset, sink, nmax, nclip_low, nclip_up, nstop, typ_var, mad, stream);
}

int mpdaf_b200_sigma_clipping_arrays64(int nfiles, const void* const* data, const void* const* var, int elem_type,
                                       long long nvox, int location, int device, double* out, double* outvar,
                                       int* expmap, const double* scale, const double* offset,
                                       long long* selected_pix, long long* valid_pix, int nmax, double nclip_low,
                                       double nclip_up, int nstop, int typ_var, int mad, void* stream)
{
    if (!selected_pix || !valid_pix) return fail(MPDAF_B200_ERR_ARG, "null counter pointer");
    CounterSink sink;
    sink.valid64 = valid_pix;
    sink.select64 = selected_pix;
    return merge_arrays(nfiles, data, var, elem_type, nvox, location, device, false, out, outvar, expmap, scale,
                        offset, sink, nmax, nclip_low, nclip_up, nstop, typ_var, mad, stream);
}

int mpdaf_b200_median_arrays(int nfiles, const void* const* data, int elem_type, long long nvox, int location,
                             int device, double* out, int* expmap, int* valid_pix, void* stream)
{
    if (!valid_pix) return fail(MPDAF_B200_ERR_ARG, "null counter pointer");
    CounterSink sink;
    sink.valid32 = valid_pix;
    return merge_arrays(nfiles, data, nullptr, elem_type, nvox, location, device, true, out, nullptr, expmap,
                        nullptr, nullptr, sink, 0, 0, 0, 1, 1, 0, stream);
}

int mpdaf_b200_set_devices(int n, const int* ids)
{
    std::lock_guard<std::mutex> lock(g_call_mutex);
    if (n < 0 || (n > 0 && ids == nullptr)) return fail(MPDAF_B200_ERR_ARG, "bad device list");
    const std::vector<int> usable = usable_devices();
    std::vector<int> want;
    for (int i = 0; i < n; i++) {
        if (std::find(usable.begin(), usable.end(), ids[i]) == usable.end()) return fail(MPDAF_B200_ERR_ARG, "not a usable sm_100 device");
        if (std::find(want.begin(), want.end(), ids[i]) == want.end()) want.push_back(ids[i]);
    }
    g_devices = want;   // empty: back to the default (MPDAF_B200_DEVICES, else every usable device)
    return MPDAF_B200_OK;
}

// Frees what the library keeps between calls: the per-device staging slots (device buffers, pinned host staging,
// streams, events) and the per-device scratch.  The next call allocates again.
int mpdaf_b200_release(void)
{
    std::lock_guard<std::mutex> lock(g_call_mutex);
    DeviceGuard guard;
    for (int dev = 0; dev < 16; dev++) {
        bool any = g_dev[dev].ready;
        for (int k = 0; k < kSlots; k++) any = any || g_slots[dev][k].stream != nullptr;
        if (!any || cudaSetDevice(dev) != cudaSuccess) continue;
        cudaDeviceSynchronize();
        for (int k = 0; k < kSlots; k++) {
            Slot& s = g_slots[dev][k];
            if (!s.stream) continue;
            cudaFree(s.d_in); cudaFree(s.d_out); cudaFree(s.d_var); cudaFree(s.d_exp); cudaFree(s.d_table);
            cudaFree(s.slow_q); cudaFree(s.slow_n);
            if (s.h_in) cudaFreeHost(s.h_in);
            if (s.h_out) cudaFreeHost(s.h_out);
            cudaEventDestroy(s.k0); cudaEventDestroy(s.k1);
            cudaStreamDestroy(s.stream);
            s = Slot();
        }
        DeviceState& d = g_dev[dev];
        if (d.ready) {
            cudaFree(d.counters); cudaFree(d.table); cudaFree(d.slow); cudaFree(d.slow_q); cudaFree(d.slow_n); cudaFree(d.hist);
            cudaEventDestroy(d.ev0); cudaEventDestroy(d.ev1);
            d = DeviceState();
        }
    }
    cudaGetLastError();
    return MPDAF_B200_OK;
}

// Finite mask of a device-resident output cube: mask[i] = 1 where data[i] is NaN or infinite.
int mpdaf_b200_finite_mask(const double* data, long long nvox, int device, unsigned char* mask, long long* nmasked, void* stream)
{
    if ((!data || !mask) && nvox > 0) return fail(MPDAF_B200_ERR_ARG, "null argument");
    if (nvox < 0) return fail(MPDAF_B200_ERR_ARG, "bad voxel count");
    std::lock_guard<std::mutex> lock(g_call_mutex);
    DeviceGuard guard;
    const int rc = device_ready(device);
    if (rc != MPDAF_B200_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceState& d = g_dev[device];
    unsigned long long* counter = d.hist;                                   // first histogram bin doubles as the counter
    MB_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    if (nvox > 0) {
        const int grid = (int)std::min<long long>((nvox / 4 + 255) / 256 + 1, (long long)d.sm_count * 16);
        finite_mask_kernel<<<grid, 256, 0, st>>>(data, nvox, mask, counter);
        MB_CUDA(cudaGetLastError());
    }
    unsigned long long host = 0;
    MB_CUDA(cudaMemcpyAsync(&host, counter, sizeof host, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    if (nmasked) *nmasked = (long long)host;
    return MPDAF_B200_OK;
}

double mpdaf_b200_last_kernel_ms(void) { return g_last_kernel_ms; }

int mpdaf_b200_last_expmap_histogram(long long* hist, int nbins)
{
    if (!hist || nbins < 1) return MPDAF_B200_ERR_ARG;
    for (int i = 0; i < nbins; i++) hist[i] = i <= kGenericMax ? (long long)g_last_hist[i] : 0;
    return MPDAF_B200_OK;
}

long long mpdaf_b200_slow_voxels(int device, int reset)
{
    if (device < 0 || device >= 16 || !g_dev[device].ready) return -1;
    DeviceGuard guard;
    unsigned long long v = 0;
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    if (cudaMemcpy(&v, g_dev[device].slow, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (reset) cudaMemset(g_dev[device].slow, 0, sizeof v);
    return (long long)v;
}
int mpdaf_b200_last_launch_count(void) { return g_last_launches; }
const char* mpdaf_b200_last_kernel_name(void) { return g_last_kernel.c_str(); }
const char* mpdaf_b200_last_error(void) { return g_error.c_str(); }

int mpdaf_b200_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    int usable = 0;
    for (int i = 0; i < count; i++) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) usable++;
    }
    return usable;
}

int mpdaf_b200_fits_image_info(const char* path, const char* extname, long naxes[3], int* bitpix,
                               long long* data_offset)
{
    if (!path || !extname || !naxes) return fail(MPDAF_B200_ERR_ARG, "null argument");
    mbfits::ImageHdu h;
    std::string err;
    const int rc = mbfits::open_image(path, extname, &h, &err);
    if (rc != MPDAF_B200_OK) return fail(rc, err);
    for (int k = 0; k < 3; k++) naxes[k] = (long)h.naxes[k];
    if (bitpix) *bitpix = h.bitpix;
    if (data_offset) *data_offset = h.data_offset;
    mbfits::close_image(&h);
    return MPDAF_B200_OK;
}

int mpdaf_b200_synth_fill(void* dst, int location, int device, int kind, unsigned long long seed, int exposure,
                          long long v0, long long count, int nx, int ny, void* stream)
{
    if (!dst || count < 0 || nx < 1 || ny < 1 || (kind != 0 && kind != 1)) return fail(MPDAF_B200_ERR_ARG, "bad synth argument");
    if (location == MPDAF_B200_HOST) {
        uint32_t* p = static_cast<uint32_t*>(dst);
        const int nthreads = (int)std::max(1LL, std::min<long long>(std::thread::hardware_concurrency(), count / 65536 + 1));
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; t++)
            pool.emplace_back([=]() {
                const long long lo = count * t / nthreads, hi = count * (t + 1) / nthreads;
                for (long long i = lo; i < hi; i++)
                    p[i] = kind == 0 ? mbsynth::data_bits(seed, exposure, v0 + i, nx, ny)
                                     : mbsynth::stat_bits(seed, exposure, v0 + i);
            });
        for (auto& th : pool) th.join();
        return MPDAF_B200_OK;
    }
    DeviceGuard guard;
    const int rc = device_ready(device);
    if (rc != MPDAF_B200_OK) return rc;
    if (count == 0) return MPDAF_B200_OK;
    const int grid = (int)std::min<long long>((count + 255) / 256, (long long)g_dev[device].sm_count * 16);
    synth_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<uint32_t*>(dst), kind, seed, exposure, v0,
                                                         count, nx, ny);
    MB_CUDA(cudaGetLastError());
    return MPDAF_B200_OK;
}

}  // extern "C"
