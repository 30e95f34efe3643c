/*
 * fsb200.cu -- host side of libfsb200.so and its C ABI (include/fsb200.h).
 *
 * One process drives one GPU (fsb_init).  Per-frame tables live in HBM inside
 * an fsb_frame; every calling host thread owns a stream plus scratch buffers,
 * so the reference's thread-per-tile dispatch (mthreading.py:48-68) can call
 * fsb_frame_run concurrently and the kernels overlap on the device.
 *
 * There is no CPU fallback: every compute entry point needs a CUDA device.
 */
#include "../../include/fsb200.h"
#include "fsb_kernels.cuh"

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>
#include <string>
#include <thread>
#include <vector>

using namespace fsb;

namespace {

thread_local std::string g_err;
int g_device = -1;
int g_sm_count = 0;
std::mutex g_mutex;
double *g_flush_buf = nullptr;
const long long FLUSH_DOUBLES = (192LL << 20) / 8; /* 192 MiB > 126 MB L2 */

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                  \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess)                                                    \
            return fail(-1, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), \
                        __FILE__, __LINE__, cudaGetErrorString(e_));              \
    } while (0)

int ensure_init()
{
    if (g_device >= 0) return 0;
    return fsb_init(0);
}

/* ---- device-memory pool ---------------------------------------------------
 * Per-frame tables (orbit, derivative paths, BLA tree: 0.1 - 0.5 GB) are
 * created and destroyed once per frame; cudaMalloc / cudaFree of blocks that
 * size cost ~10 ms each (map / unmap), 0.23 s per frame of a zoom movie.  Freed
 * blocks are kept, keyed by size, and handed to the next frame (whose tables
 * have the same sizes); FSB200_POOL_MB bounds the cached bytes (default 16 GB
 * of the 180 GB of HBM), fsb_shutdown releases them. */
struct Pool {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;
    std::unordered_map<void *, size_t> live;
    size_t cached = 0;
    size_t cap()
    {
        static size_t c = [] {
            const char *e = getenv("FSB200_POOL_MB");
            return (size_t)(e ? atoll(e) : 16384) << 20;
        }();
        return c;
    }
};
Pool g_pool;
size_t pool_round(size_t b)
{
    if (b < 1) b = 1;
    const size_t g = (b < (1u << 20)) ? 512 : (2u << 20);
    return (b + g - 1) / g * g;
}
void pool_trim()
{
    std::lock_guard<std::mutex> lock(g_pool.mu);
    for (auto &kv : g_pool.free_blocks) cudaFree(kv.second);
    g_pool.free_blocks.clear();
    g_pool.cached = 0;
}
cudaError_t pool_alloc(void **p, size_t bytes)
{
    bytes = pool_round(bytes);
    {
        std::lock_guard<std::mutex> lock(g_pool.mu);
        auto it = g_pool.free_blocks.lower_bound(bytes);
        if (it != g_pool.free_blocks.end() && it->first <= bytes + bytes / 4) {
            *p = it->second;
            g_pool.live[*p] = it->first;
            g_pool.cached -= it->first;
            g_pool.free_blocks.erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {            /* give the cached blocks back and retry */
        cudaGetLastError();
        pool_trim();
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lock(g_pool.mu);
        g_pool.live[*p] = bytes;
    }
    return e;
}
template <class T> cudaError_t pool_alloc(T **p, size_t bytes) { return pool_alloc((void **)p, bytes); }
void pool_free(void *p)
{
    if (!p) return;
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lock(g_pool.mu);
        auto it = g_pool.live.find(p);
        if (it != g_pool.live.end()) {
            bytes = it->second;
            g_pool.live.erase(it);
            if (g_pool.cached + bytes <= g_pool.cap()) {
                g_pool.free_blocks.emplace(bytes, p);
                g_pool.cached += bytes;
                return;
            }
        }
    }
    cudaFree(p);
}

/* Per host-thread launch context. */
const int MAX_CHUNKS = 16;
const int CTL_WORDS = 8;                 /* per chunk: [work, c0..c3, -, -, -] */
struct Ctx {
    cudaStream_t stream = nullptr, side = nullptr;      /* compute / abort flag */
    cudaStream_t s_alt = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evc0 = nullptr, evc1 = nullptr;
    cudaEvent_t ev_h[MAX_CHUNKS] = {}, ev_k[MAX_CHUNKS] = {};
    void *d_buf = nullptr;  long long d_cap = 0;      /* scratch for in/out   */
    unsigned long long *d_ctl = nullptr;              /* MAX_CHUNKS control blocks + abort flag */
    unsigned long long *h_ctl = nullptr;              /* pinned mirror        */
    int4 *d_tiles = nullptr;  long long tiles_cap = 0; /* tile descriptors of the running call */
    C *d_proj = nullptr;  long long proj_cap = 0;     /* projected pixels of the running call */
    char *d_axes = nullptr;  long long axes_cap = 0;  /* per-tile pixel axes of a grid call */
    int *d_abort() { return (int *)(d_ctl + MAX_CHUNKS * CTL_WORDS); }
    ~Ctx()
    {
        if (d_buf) cudaFree(d_buf);
        if (d_ctl) cudaFree(d_ctl);
        if (d_tiles) cudaFree(d_tiles);
        if (d_proj) cudaFree(d_proj);
        if (d_axes) cudaFree(d_axes);
        if (h_ctl) cudaFreeHost(h_ctl);
        for (cudaEvent_t e : {ev0, ev1, evc0, evc1}) if (e) cudaEventDestroy(e);
        for (int i = 0; i < MAX_CHUNKS; i++) {
            if (ev_h[i]) cudaEventDestroy(ev_h[i]);
            if (ev_k[i]) cudaEventDestroy(ev_k[i]);
        }
        for (cudaStream_t st : {stream, side, s_alt, s_h2d, s_d2h}) if (st) cudaStreamDestroy(st);
    }
};
thread_local Ctx *t_ctx = nullptr;
/* frees the context of a worker thread when the thread ends (the seam is called
 * from thread-per-tile pools); skipped once the CUDA runtime is unloading */
struct CtxReaper {
    ~CtxReaper()
    {
        if (t_ctx && g_device >= 0 && cudaSetDevice(g_device) == cudaSuccess) delete t_ctx;
        t_ctx = nullptr;
    }
};
thread_local CtxReaper t_ctx_reaper;

static int ctx_build(Ctx *c);
int get_ctx(Ctx **out)
{
    if (ensure_init() != 0) return -1;
    if (!t_ctx) {
        (void)&t_ctx_reaper;                      /* instantiate this thread's reaper */
        Ctx *c = new Ctx();
        if (ctx_build(c) != 0) { delete c; return -1; }   /* no half-built context is kept */
        t_ctx = c;
    } else {
        CK(cudaSetDevice(g_device));
    }
    *out = t_ctx;
    return 0;
}

static int ctx_build(Ctx *c)
{
    {
        CK(cudaSetDevice(g_device));
        for (cudaStream_t *st : {&c->stream, &c->side, &c->s_alt, &c->s_h2d, &c->s_d2h})
            CK(cudaStreamCreateWithFlags(st, cudaStreamNonBlocking));
        for (cudaEvent_t *e : {&c->ev0, &c->ev1, &c->evc0, &c->evc1}) CK(cudaEventCreate(e));
        for (int i = 0; i < MAX_CHUNKS; i++) {
            CK(cudaEventCreateWithFlags(&c->ev_h[i], cudaEventDisableTiming));
            CK(cudaEventCreate(&c->ev_k[i]));
        }
        const size_t ctl_bytes = (MAX_CHUNKS * CTL_WORDS + 2) * sizeof(unsigned long long);
        CK(cudaMalloc(&c->d_ctl, ctl_bytes));
        CK(cudaHostAlloc(&c->h_ctl, ctl_bytes, cudaHostAllocDefault));
    }
    return 0;
}

int ctx_reserve(Ctx *c, long long bytes)
{
    if (bytes <= c->d_cap) return 0;
    if (c->d_buf) { CK(cudaFree(c->d_buf)); c->d_buf = nullptr; c->d_cap = 0; }
    long long cap = bytes + (bytes >> 3);
    CK(cudaMalloc(&c->d_buf, (size_t)cap));
    c->d_cap = cap;
    return 0;
}

inline long long align256(long long x) { return (x + 255) & ~255LL; }

/* grid for the persistent pixel kernels */
template <class K> int persistent_grid(K kernel, int block, long long npts, int *grid)
{
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0));
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)per_sm * g_sm_count;
    long long need = (npts + block - 1) / block;
    if (need < 1) need = 1;
    if (g > need) g = need;
    *grid = (int)g;
    return 0;
}

/* wait for `done`, polling the caller's interruption flag */
int wait_event(Ctx *c, cudaEvent_t done, const volatile uint8_t *interrupted,
               bool *was_interrupted)
{
    *was_interrupted = false;
    for (;;) {
        cudaError_t q = cudaEventQuery(done);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) CK(q);
        if (interrupted && *interrupted && !*was_interrupted) {
            *was_interrupted = true;
            c->h_ctl[MAX_CHUNKS * CTL_WORDS] = 1;      /* pinned source */
            CK(cudaMemcpyAsync(c->d_abort(), c->h_ctl + MAX_CHUNKS * CTL_WORDS, sizeof(int),
                               cudaMemcpyHostToDevice, c->side));
        }
        std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
    if (*was_interrupted) CK(cudaStreamSynchronize(c->side));
    return 0;
}

/* The point list of one call, as work units (see Tiling in fsb_kernels.cuh):
 * a flat list (units of 32 consecutive points) or a concatenation of row-major
 * tiles (units = 8 x 4 patches). */
struct Units {
    long long npts = 0;
    int n_units = 0;
    std::vector<int4> tiles;      /* empty: flat list */
    const int4 *d_tiles = nullptr;
    bool tiled() const { return !tiles.empty(); }
};

int units_flat(long long npts, Units &u)
{
    if (npts >= (1LL << 31) - 64) return fail(-3, "too many points in one call (%lld)", npts);
    u.npts = npts;
    u.n_units = (int)((npts + 31) / 32);
    return 0;
}

int units_tiled(Ctx *c, int n_tiles, const int32_t *tw, const int32_t *th, cudaStream_t st, Units &u)
{
    if (n_tiles <= 0 || !tw || !th) return fail(-3, "empty tile list");
    long long pts = 0, units = 0;
    u.tiles.resize((size_t)n_tiles);
    for (int k = 0; k < n_tiles; k++) {
        if (tw[k] <= 0 || th[k] <= 0) return fail(-3, "tile %d has shape %d x %d", k, tw[k], th[k]);
        u.tiles[(size_t)k] = make_int4((int)units, (int)pts, tw[k], th[k]);
        pts += (long long)tw[k] * th[k];
        units += (long long)((tw[k] + 7) / 8) * ((th[k] + 3) / 4);
        if (pts >= (1LL << 31) - 64 || units >= (1LL << 31) - 64)
            return fail(-3, "too many points in one call");
    }
    u.npts = pts;
    u.n_units = (int)units;
    if (n_tiles > c->tiles_cap) {
        if (c->d_tiles) { CK(cudaFree(c->d_tiles)); c->d_tiles = nullptr; c->tiles_cap = 0; }
        CK(cudaMalloc(&c->d_tiles, (size_t)(2 * n_tiles) * sizeof(int4)));
        c->tiles_cap = 2 * n_tiles;
    }
    /* pageable source: the driver stages it before returning */
    CK(cudaMemcpyAsync(c->d_tiles, u.tiles.data(), (size_t)n_tiles * sizeof(int4),
                       cudaMemcpyHostToDevice, st));
    u.d_tiles = c->d_tiles;
    return 0;
}

Tiling tiling_of(const Units &u, int unit_lo, int unit_hi)
{
    Tiling t;
    t.tiles = u.d_tiles;
    t.n_tiles = (int)u.tiles.size();
    t.unit_lo = unit_lo;
    t.unit_hi = unit_hi;
    return t;
}

/* Chunking of a host-buffer call: the frame is cut into up to MAX_CHUNKS
 * slabs of consecutive points so that the H2D copy of slab k+1, the kernel of
 * slab k and the D2H copies of slab k-1 overlap (three copy/compute streams;
 * kernels alternate between two streams so that the tail of one slab overlaps
 * the head of the next).  Slabs end on unit boundaries: a multiple of 32 points
 * of a flat list, whole tiles of a tile list. */
struct Chunks { int n; long long beg[MAX_CHUNKS + 1]; int ubeg[MAX_CHUNKS + 1]; };
Chunks make_chunks(const Units &u)
{
    Chunks ch;
    const long long npts = u.npts;
    long long want = npts / (1LL << 20);
    if (want < 1) want = 1;
    if (want > MAX_CHUNKS) want = MAX_CHUNKS;
    if (npts < (1LL << 17)) want = 1;
    long long per = ((npts + want - 1) / want + 31) & ~31LL;
    ch.n = 0;
    ch.beg[0] = 0;
    ch.ubeg[0] = 0;
    if (!u.tiled()) {
        for (long long b = 0; b < npts; b += per) {
            ++ch.n;
            ch.beg[ch.n] = (b + per < npts) ? b + per : npts;
            ch.ubeg[ch.n] = (int)((ch.beg[ch.n] + 31) / 32);
        }
        return ch;
    }
    const size_t nt = u.tiles.size();
    for (size_t k = 0; k < nt; k++) {
        const long long end = (k + 1 < nt) ? u.tiles[k + 1].y : npts;
        const int uend = (k + 1 < nt) ? u.tiles[k + 1].x : u.n_units;
        const bool last = (k + 1 == nt);
        if (last || (end - ch.beg[ch.n] >= per && ch.n + 1 < MAX_CHUNKS)) {
            ++ch.n;
            ch.beg[ch.n] = end;
            ch.ubeg[ch.n] = uend;
        }
    }
    return ch;
}

/* ---- projection passes ---------------------------------------------------- */
/* Points covered by the units [unit_lo, unit_hi) of a call: a contiguous range
 * (flat list: 32 points per unit; tile list: launches start and end on tile
 * boundaries, see chunks_of). */
int point_range(const Units &u, int unit_lo, int unit_hi, long long &lo, long long &hi)
{
    if (!u.tiled()) {
        lo = 32LL * unit_lo;
        hi = std::min(32LL * unit_hi, u.npts);
        return 0;
    }
    lo = hi = -1;
    for (const int4 &t : u.tiles) {
        if (t.x == unit_lo) lo = t.y;
        if (t.x == unit_hi) hi = t.y;
    }
    if (unit_hi == u.n_units) hi = u.npts;
    if (lo < 0 || hi < 0) return fail(-1, "internal: launch not aligned on tile boundaries");
    return 0;
}
/* scratch plane for the projected pixels of the running call (whole list:
 * concurrent launches of one call work on disjoint point ranges) */
int proj_scratch(Ctx *c, long long npts)
{
    if (npts <= c->proj_cap) return 0;
    CK(cudaDeviceSynchronize());
    if (c->d_proj) { CK(cudaFree(c->d_proj)); c->d_proj = nullptr; c->proj_cap = 0; }
    CK(cudaMalloc(&c->d_proj, (size_t)npts * sizeof(C)));
    c->proj_cap = npts;
    return 0;
}
/* pixels -> projected pixels on stream st; *d_c_pix is redirected to the scratch */
int enqueue_projection(Ctx *c, const ProjDev &P, const Units &u, int unit_lo, int unit_hi,
                       cudaStream_t st, const C **d_c_pix, long long &lo, long long &hi)
{
    if (point_range(u, unit_lo, unit_hi, lo, hi)) return -1;
    if (P.kind == FSB_PROJ_CARTESIAN || hi <= lo) return 0;
    if (proj_scratch(c, u.npts)) return -1;
    k_proj_range<<<(int)((hi - lo + 255) / 256), 256, 0, st>>>(P, lo, hi, *d_c_pix, c->d_proj);
    CK(cudaGetLastError());
    *d_c_pix = c->d_proj;
    return 0;
}

/* ---- kernel dispatch ------------------------------------------------------ */
typedef void (*perturb_kernel_t)(FrameDev, long long, const C *, double *, int *,
                                 signed char *, int *, unsigned long long *,
                                 unsigned long long *, const volatile int *, Tiling);

template <bool XR, bool DC, bool DZ, bool BLA> perturb_kernel_t pick_m2_extra(bool extra, bool fastxr)
{
    if (XR && !DZ && fastxr)
        return extra ? k_perturb_m2<XR, DC, DZ, BLA, true, (XR && !DZ)>
                     : k_perturb_m2<XR, DC, DZ, BLA, false, (XR && !DZ)>;
    return extra ? k_perturb_m2<XR, DC, DZ, BLA, true, false> : k_perturb_m2<XR, DC, DZ, BLA, false, false>;
}
template <bool XR, bool DC, bool DZ> perturb_kernel_t pick_m2_bla(bool bla, bool extra, bool fastxr)
{
    return bla ? pick_m2_extra<XR, DC, DZ, true>(extra, fastxr) : pick_m2_extra<XR, DC, DZ, false>(extra, fastxr);
}
template <bool XR, bool DC> perturb_kernel_t pick_m2_dz(bool dz, bool bla, bool extra, bool fastxr)
{
    return dz ? pick_m2_bla<XR, DC, true>(bla, extra, fastxr) : pick_m2_bla<XR, DC, false>(bla, extra, fastxr);
}
template <bool XR> perturb_kernel_t pick_m2_dc(bool dc, bool dz, bool bla, bool extra, bool fastxr)
{
    return dc ? pick_m2_dz<XR, true>(dz, bla, extra, fastxr) : pick_m2_dz<XR, false>(dz, bla, extra, fastxr);
}
/* event-driven kernel (k_perturb_m2_v2): the common variants -- no interior
 * detection, no periodic reference / calc_orbit, power 2; Xrange frames with
 * the guarded fp64 lane */
template <bool XR> perturb_kernel_t pick_m2_v2(bool dc, bool bla)
{
    if (dc) return bla ? k_perturb_m2_v2<XR, true, true> : k_perturb_m2_v2<XR, true, false>;
    return bla ? k_perturb_m2_v2<XR, false, true> : k_perturb_m2_v2<XR, false, false>;
}
perturb_kernel_t pick_m2(bool xr, bool dc, bool dz, bool bla, bool extra, bool fastxr, bool v2)
{
    if (v2) return xr ? pick_m2_v2<true>(dc, bla) : pick_m2_v2<false>(dc, bla);
    return xr ? pick_m2_dc<true>(dc, dz, bla, extra, fastxr) : pick_m2_dc<false>(dc, dz, bla, extra, fastxr);
}
/* Perturbation_mandelbrot_N: the same loop with the binomial model formulas
 * (EXTRA variants: runtime ref_order wrap; no fp64 fast lane) */
template <bool XR, bool DC, bool DZ> perturb_kernel_t pick_mn_bla(bool bla)
{
    return bla ? k_perturb_m2<XR, DC, DZ, true, true, false, true>
               : k_perturb_m2<XR, DC, DZ, false, true, false, true>;
}
template <bool XR> perturb_kernel_t pick_mn_dc(bool dc, bool dz, bool bla)
{
    if (dc) return dz ? pick_mn_bla<XR, true, true>(bla) : pick_mn_bla<XR, true, false>(bla);
    return dz ? pick_mn_bla<XR, false, true>(bla) : pick_mn_bla<XR, false, false>(bla);
}
perturb_kernel_t pick_mn(bool xr, bool dc, bool dz, bool bla)
{
    return xr ? pick_mn_dc<true>(dc, dz, bla) : pick_mn_dc<false>(dc, dz, bla);
}
template <bool XR, bool H, bool BLA, bool FX> perturb_kernel_t pick_bs_flavor(int flavor)
{
    if constexpr (!XR) {
        return k_perturb_bs<XR, H, BLA, false, 0>;         /* small kernels: run-time flavour */
    } else {
        switch (flavor) {
        case 1: return k_perturb_bs<XR, H, BLA, FX, 1>;
        case 2: return k_perturb_bs<XR, H, BLA, FX, 2>;
        case 3: return k_perturb_bs<XR, H, BLA, FX, 3>;
        case 4: return k_perturb_bs<XR, H, BLA, false, 4>; /* no fp64 lane for flavours 4-5 */
        default: return k_perturb_bs<XR, H, BLA, false, 5>;
        }
    }
}
template <bool XR, bool H> perturb_kernel_t pick_bs_bla(bool bla, bool fastxr, int flavor)
{
    if (XR && fastxr && flavor <= 3)
        return bla ? pick_bs_flavor<XR, H, true, XR>(flavor) : pick_bs_flavor<XR, H, false, XR>(flavor);
    return bla ? pick_bs_flavor<XR, H, true, false>(flavor) : pick_bs_flavor<XR, H, false, false>(flavor);
}
template <bool XR> perturb_kernel_t pick_bs_h(bool h, bool bla, bool fastxr, int flavor)
{
    return h ? pick_bs_bla<XR, true>(bla, fastxr, flavor) : pick_bs_bla<XR, false>(bla, fastxr, flavor);
}
perturb_kernel_t pick_bs(bool xr, bool h, bool bla, bool fastxr, int flavor)
{
    return xr ? pick_bs_h<true>(h, bla, fastxr, flavor) : pick_bs_h<false>(h, bla, fastxr, flavor);
}

} /* namespace */

/* ======================================================================== */
struct fsb_frame {
    fsb_frame_desc d;
    FrameDev dev;
    ProjDev proj;
    int nz = 0;
    bool bla_on = false;
    bool fast_xr = false;     /* Xrange kernel with the guarded fp64 fast path */
    bool v2 = false;          /* event-driven kernel k_perturb_m2_v2 (interleaved orbit table) */
    bool gpu_scan = false;    /* dZndc path by the GPU affine scan (K6) */
    bool gpu_scan_bs = false; /* same for the four Jacobian paths of the burning-ship family */
    std::vector<void *> owned;
    double ms_upload = 0, ms_dzndc = 0, ms_bla = 0;
    long long dzndc_len = 0;
};

namespace {

/* Device copy of a host table, followed by `pad` zero elements.  The orbit and
 * the dZndc path carry one zero pad element: a pixel that walks the whole
 * orbit by BLA steps without a rebase ends with w_iter == L and the reference
 * reads one element past its arrays there (undefined in the reference; defined
 * as 0 here and in the oracle). */
template <class T> int upload(fsb_frame *f, const T *host, long long n, const T **dev,
                              long long pad = 0)
{
    *dev = nullptr;
    if (!host || n <= 0) return 0;
    void *p = nullptr;
    CK(pool_alloc(&p, (size_t)((n + pad) * (long long)sizeof(T))));
    f->owned.push_back(p);
    CK(cudaMemcpy(p, host, (size_t)(n * (long long)sizeof(T)), cudaMemcpyHostToDevice));
    if (pad > 0)
        CK(cudaMemset((char *)p + n * (long long)sizeof(T), 0, (size_t)(pad * (long long)sizeof(T))));
    *dev = (const T *)p;
    return 0;
}

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

/* Stateless lookup in the sorted Xrange index (perturbation.py:2519-2588). */
long long h_xr_find(const int32_t *index, long long n, long long idx)
{
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (index[mid] < idx) lo = mid + 1; else hi = mid;
    }
    return (lo < n && index[lo] == idx) ? lo : -1;
}

/* dfdz of the holomorphic models: 2 z (mandelbrot_M2.py:599-602) or
 * N z^(N-1) by repeated products (mandelbrot_Mn.py:643-649) */
template <class T> T h_dfdz(int nexp, T z)
{
    if (nexp == 0) return 2. * z;
    T tmp = z;
    for (int k = 2; k < nexp; k++) tmp = tmp * z;
    return (double)nexp * tmp;
}

/* dZndc path, serial recurrence on the host (perturbation.py:2282-2336).
 * dZ[i] = 2 Z[i-1] dZ[i-1] + scale ; Xrange variant keeps (mantissa, exp). */
void host_dzndc_m2(const fsb_frame_desc &d, std::vector<C> &out, std::vector<int32_t> &oe)
{
    const long long L = d.L;
    const C *Zn = (const C *)d.Zn_path;
    const C *rxr = (const C *)d.ref_xr;
    long long valid = L < d.ref_div_iter ? L : d.ref_div_iter;
    XF scale_x = mkXF(d.scale_deriv, d.scale_deriv_e);
    double scale = to_std(scale_x);
    out.assign((size_t)L, mkC(0., 0.));
    oe.assign(d.xr_detect ? (size_t)L : 0, 0);
    if (valid < 2) return;
    if (d.xr_detect) {
        for (long long i = 1; i < valid; i++) {
            long long k = d.n_xr > 0 ? h_xr_find(d.ref_index_xr, d.n_xr, i - 1) : -1;
            XC rz = (k >= 0) ? mkXC(rxr[k], d.ref_xr_e[k]) : to_xr(Zn[i - 1]);
            XC v = h_dfdz(d.nexp, rz) * mkXC(out[i - 1], oe[i - 1]) + scale_x;
            out[i] = v.m; oe[i] = v.e;
        }
        long long i = valid - 1;
        if (i == d.ref_order - 1) {
            XC v = h_dfdz(d.nexp, Zn[i]) * mkXC(out[i], oe[i]) + scale_x;
            out[0] = v.m; oe[0] = v.e;
        }
    } else {
        for (long long i = 1; i < valid; i++) out[i] = h_dfdz(d.nexp, Zn[i - 1]) * out[i - 1] + scale;
        long long i = valid - 1;
        if (i == d.ref_order - 1) out[0] = h_dfdz(d.nexp, Zn[i]) * out[i] + scale;
    }
}

/* dZndz path (perturbation.py:2466-2516) */
void host_dzndz_m2(const fsb_frame_desc &d, std::vector<C> &out, std::vector<int32_t> &oe)
{
    const long long L = d.L;
    const C *Zn = (const C *)d.Zn_path;
    const C *rxr = (const C *)d.ref_xr;
    long long valid = L < d.ref_div_iter ? L : d.ref_div_iter;
    out.assign((size_t)L + 1, mkC(0., 0.));
    oe.assign(d.xr_detect ? (size_t)L + 1 : 0, 0);
    out[1] = mkC(1., 0.);
    if (valid < 3) return;
    if (d.xr_detect) {
        for (long long i = 2; i < valid; i++) {
            long long k = d.n_xr > 0 ? h_xr_find(d.ref_index_xr, d.n_xr, i - 1) : -1;
            XC rz = (k >= 0) ? mkXC(rxr[k], d.ref_xr_e[k]) : to_xr(Zn[i - 1]);
            XC v = h_dfdz(d.nexp, rz) * mkXC(out[i - 1], oe[i - 1]);
            out[i] = v.m; oe[i] = v.e;
        }
        long long i = valid - 1;
        long long k = d.n_xr > 0 ? h_xr_find(d.ref_index_xr, d.n_xr, i) : -1;
        XC rz = (k >= 0) ? mkXC(rxr[k], d.ref_xr_e[k]) : to_xr(Zn[i]);
        XC v = h_dfdz(d.nexp, rz) * mkXC(out[i], oe[i]);
        out[L] = v.m; oe[L] = v.e;
    } else {
        for (long long i = 2; i < valid; i++) out[i] = h_dfdz(d.nexp, Zn[i - 1]) * out[i - 1];
        long long i = valid - 1;
        out[L] = h_dfdz(d.nexp, Zn[i]) * out[i];
    }
}

/* burning_ship.py:441-532 on the host, plain double and Xrange */
void h_bs_jac(int flavor, double x, double y, double &fxx, double &fxy, double &fyx, double &fyy)
{
    switch (flavor) {
    case 1: fxx = 2. * x; fxy = -2. * y; fyx = 2. * sgn(x) * fabs(y); fyy = 2. * sgn(y) * fabs(x); break;
    case 2: fxx = 2. * x; fxy = -2. * y; fyx = 2. * fabs(y); fyy = 2. * sgn(y) * x; break;
    case 3: fxx = 2. * x; fxy = -2. * fabs(y); fyx = 2. * y; fyy = 2. * x; break;
    case 4: { double s = sgn(x * x - y * y); fxx = 2. * s * x; fxy = -2. * s * y; fyx = 2. * y; fyy = 2. * x; break; }
    default: { double s = sgn(x * x - y * y); fxx = 2. * s * x; fxy = -2. * s * y; fyx = 2. * sgn(x) * fabs(y); fyy = 2. * sgn(y) * fabs(x); break; }
    }
}
void h_bs_jac(int flavor, XF x, XF y, XF &fxx, XF &fxy, XF &fyx, XF &fyy)
{
    switch (flavor) {
    case 1: fxx = 2. * x; fxy = -2. * y; fyx = 2. * sgn_(x) * fabs_(y); fyy = 2. * sgn_(y) * fabs_(x); break;
    case 2: fxx = 2. * x; fxy = -2. * y; fyx = 2. * fabs_(y); fyy = 2. * sgn_(y) * x; break;
    case 3: fxx = 2. * x; fxy = -2. * fabs_(y); fyx = 2. * y; fyy = 2. * x; break;
    case 4: { double s = sgn_(x * x - y * y); fxx = 2. * s * x; fxy = -2. * s * y; fyx = 2. * y; fyy = 2. * x; break; }
    default: { double s = sgn_(x * x - y * y); fxx = 2. * s * x; fxy = -2. * s * y; fyx = 2. * sgn_(x) * fabs_(y); fyy = 2. * sgn_(y) * fabs_(x); break; }
    }
}

/* perturbation.py:2339-2463 ; out = [dXnda | dXndb | dYnda | dYndb] */
void host_dzndc_bs(const fsb_frame_desc &d, std::vector<double> &out, std::vector<int32_t> &oe)
{
    const long long L = d.L;
    const double *Zn = d.Zn_path;
    long long valid = L < d.ref_div_iter ? L : d.ref_div_iter;
    XF scale_x = mkXF(d.scale_deriv, d.scale_deriv_e);
    double scale = to_std(scale_x);
    out.assign((size_t)(4 * L), 0.);
    oe.assign(d.xr_detect ? (size_t)(4 * L) : 0, 0);
    if (valid < 2) return;
    double *A = out.data(), *B = A + L, *Cc = A + 2 * L, *D = A + 3 * L;
    int32_t *Ae = oe.data(), *Be = Ae + L, *Ce = Ae + 2 * L, *De = Ae + 3 * L;
    long long n_steps = valid - 1;
    bool wrap = ((valid - 1) == d.ref_order - 1);
    for (long long s = 0; s < n_steps + (wrap ? 1 : 0); s++) {
        long long from_i = (s < n_steps) ? s : valid - 1;
        long long to_i = (s < n_steps) ? s + 1 : 0;
        double X = Zn[2 * from_i], Y = Zn[2 * from_i + 1];
        if (d.xr_detect) {
            long long k = d.n_xr > 0 ? h_xr_find(d.ref_index_xr, d.n_xr, from_i) : -1;
            XF rx = (k >= 0) ? mkXF(d.ref_xr[k], d.ref_xr_e[k]) : to_xr(X);
            XF ry = (k >= 0) ? mkXF(d.refy_xr[k], d.refy_xr_e[k]) : to_xr(Y);
            XF fxx, fxy, fyx, fyy;
            h_bs_jac(d.flavor, rx, ry, fxx, fxy, fyx, fyy);
            XF a = mkXF(A[from_i], Ae[from_i]), b = mkXF(B[from_i], Be[from_i]);
            XF c = mkXF(Cc[from_i], Ce[from_i]), dd = mkXF(D[from_i], De[from_i]);
            XF na = fxx * a + fxy * c + scale_x;
            XF nb = fxx * b + fxy * dd;
            XF nc = fyx * a + fyy * c;
            XF nd = fyx * b + fyy * dd - scale_x;
            A[to_i] = na.m; Ae[to_i] = na.e; B[to_i] = nb.m; Be[to_i] = nb.e;
            Cc[to_i] = nc.m; Ce[to_i] = nc.e; D[to_i] = nd.m; De[to_i] = nd.e;
        } else {
            double fxx, fxy, fyx, fyy;
            h_bs_jac(d.flavor, X, Y, fxx, fxy, fyx, fyy);
            double a = A[from_i], b = B[from_i], c = Cc[from_i], dd = D[from_i];
            A[to_i] = fxx * a + fxy * c + scale;
            B[to_i] = fxx * b + fxy * dd;
            Cc[to_i] = fyx * a + fyy * c;
            D[to_i] = fyx * b + fyy * dd - scale;
        }
    }
}

/* fp64 mirror of an Xrange value for the fast path of the Xrange kernels:
 * exact when every component is a normal double, 0 for components below the
 * normal range (they are below half an ulp of anything the fast path adds them
 * to), NaN when a component is too large (forces the exact fallback). */
[[maybe_unused]] C xr_flushed_std(C m, int e)
{
    double out[2];
    const double in[2] = {m.re, m.im};
    for (int k = 0; k < 2; k++) {
        double nm; int ne;
        normalize_real(in[k], e, nm, ne);
        if (in[k] == 0.) out[k] = 0.;
        else if (!(in[k] == in[k]) || ne > 1000) out[k] = mk64(0x7ff80000, 0);
        else if (ne < -1022) out[k] = 0.;
        else out[k] = ldexp(nm, ne);
    }
    return mkC(out[0], out[1]);
}

int stages_bla_of(long long L)
{
    int s = 0;
    while ((1LL << s) < L) s++;
    return s;
}

/* K5: BLA tree on the device */
int build_bla(fsb_frame *f)
{
    const fsb_frame_desc &d = f->d;
    long long comp_len = d.L / 8;
    long long bla_len = 2 * comp_len;
    int stages = stages_bla_of(d.L);
    f->dev.bla_len = bla_len;
    f->dev.stages_bla = stages;
    if (comp_len == 0) return 0;
    double kc_std = to_std(mkXF(d.kc, d.kc_e));
    int width = (d.model == FSB_MODEL_M2) ? 4 : 8; /* doubles per node */
    void *dM = nullptr, *dr = nullptr;
    CK(pool_alloc(&dM, (size_t)(bla_len * width * 8)));
    f->owned.push_back(dM);
    CK(pool_alloc(&dr, (size_t)(bla_len * 8)));
    f->owned.push_back(dr);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, 0));
    int block = 128;
    int grid = (int)((comp_len + block - 1) / block);
    if (d.model == FSB_MODEL_M2)
        k_bla_leaf_m2<<<grid, block>>>(f->dev.Zn, comp_len, kc_std, d.BLA_eps, (C *)dM, (double *)dr,
                                       d.nexp);
    else
        k_bla_leaf_bs<<<grid, block>>>(d.flavor, f->dev.Zn, comp_len, kc_std, d.BLA_eps,
                                       (double *)dM, (double *)dr);
    for (int stg = 1; stg < stages - 3; stg++) {
        long long n_out = (comp_len >> stg) + 1;
        int g = (int)((n_out + block - 1) / block);
        if (d.model == FSB_MODEL_M2)
            k_bla_merge_m2<<<g, block>>>(comp_len, stg, kc_std, d.BLA_eps, (C *)dM, (double *)dr);
        else
            k_bla_merge_bs<<<g, block>>>(comp_len, stg, kc_std, d.BLA_eps, (double *)dM, (double *)dr);
    }
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    f->ms_bla = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    f->dev.M_bla = (const double *)dM;
    f->dev.r_bla = (const double *)dr;
    return 0;
}

template <class T> int dev_zeros(fsb_frame *f, long long n, T **dev)
{
    void *p = nullptr;
    CK(pool_alloc(&p, (size_t)(n * (long long)sizeof(T))));
    f->owned.push_back(p);
    CK(cudaMemset(p, 0, (size_t)(n * (long long)sizeof(T))));
    *dev = (T *)p;
    return 0;
}

/* K6: dZndc path of a holomorphic frame by a parallel affine scan on the GPU
 * (default build; the -fmad=false build keeps the serial host loop, which is
 * bit-exact with the oracle). */
int gpu_dzndc_m2(fsb_frame *f)
{
    const fsb_frame_desc &d = f->d;
    FrameDev &v = f->dev;
    const long long L = d.L;
    const long long valid = L < d.ref_div_iter ? L : d.ref_div_iter;
    const long long n_elem = valid - 1;
    C *dm = nullptr;
    int *de = nullptr;
    if (dev_zeros(f, L + 1, &dm)) return -1;
    if (d.xr_detect && dev_zeros(f, L + 1, &de)) return -1;
    v.dZndc = dm;
    v.dZndc_e = de;
    if (n_elem <= 0) return 0;
    const XF scale = mkXF(d.scale_deriv, d.scale_deriv_e);
    const long long n_thr = (n_elem + SCAN_E - 1) / SCAN_E;
    const int n_blk = (int)((n_thr + SCAN_T - 1) / SCAN_T);
    AffXC *thr_agg = nullptr, *blk_agg = nullptr;
    CK(pool_alloc(&thr_agg, (size_t)n_blk * SCAN_T * sizeof(AffXC)));
    CK(pool_alloc(&blk_agg, (size_t)n_blk * sizeof(AffXC)));
    k_dzndc_scan_local<<<n_blk, SCAN_T>>>(v, n_elem, scale, thr_agg, blk_agg);
    k_dzndc_scan_blocks<<<1, 1024>>>(n_blk, blk_agg);
    k_dzndc_scan_apply<<<n_blk, SCAN_T>>>(v, n_elem, scale, thr_agg, blk_agg, dm, de,
                                          d.xr_detect ? nullptr : dm, d.xr_detect ? 1 : 0);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    pool_free(thr_agg);
    pool_free(blk_agg);
    /* periodic reference: the wrapped value is stored at index 0
     * (perturbation.py:2315-2334) */
    const long long i = valid - 1;
    if (i == d.ref_order - 1) {
        const C *Zn = (const C *)d.Zn_path;
        C last; int last_e = 0;
        CK(cudaMemcpy(&last, dm + i, sizeof(C), cudaMemcpyDeviceToHost));
        if (d.xr_detect) {
            CK(cudaMemcpy(&last_e, de + i, sizeof(int), cudaMemcpyDeviceToHost));
            XC w = h_dfdz(d.nexp, to_xr(Zn[i])) * mkXC(last, last_e) + scale;
            CK(cudaMemcpy(dm, &w.m, sizeof(C), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(de, &w.e, sizeof(int), cudaMemcpyHostToDevice));
        } else {
            C w = h_dfdz(d.nexp, Zn[i]) * last + to_std(scale);
            CK(cudaMemcpy(dm, &w, sizeof(C), cudaMemcpyHostToDevice));
        }
    }
    return 0;
}

/* Same for the four Jacobian paths of a burning-ship frame (SURVEY f-2):
 * k_dzndc_bs_scan_* ; out = [dXnda | dXndb | dYnda | dYndb], L entries each. */
int gpu_dzndc_bs(fsb_frame *f)
{
    const fsb_frame_desc &d = f->d;
    FrameDev &v = f->dev;
    const long long L = d.L;
    const long long valid = L < d.ref_div_iter ? L : d.ref_div_iter;
    const long long n_elem = valid - 1;
    double *dm = nullptr;
    int *de = nullptr;
    const bool dbg = getenv("FSB200_DEBUG_TIMING") != nullptr;
    double tq = now_ms();
    if (dev_zeros(f, 4 * L + 4, &dm)) return -1;
    if (d.xr_detect && dev_zeros(f, 4 * L + 4, &de)) return -1;
    if (dbg) { cudaDeviceSynchronize(); fprintf(stderr, "bs scan: zeros %.3f ms\n", now_ms() - tq); tq = now_ms(); }
    for (int j = 0; j < 4; j++) {
        v.dP[j] = dm + j * L;
        v.dP_e[j] = de ? de + j * L : nullptr;
    }
    if (n_elem <= 0) return 0;
    const XF scale = mkXF(d.scale_deriv, d.scale_deriv_e);
    const long long n_thr = (n_elem + SCAN_E - 1) / SCAN_E;
    const int n_blk = (int)((n_thr + SCANB_T - 1) / SCANB_T);
    AffBS *thr_agg = nullptr, *blk_agg = nullptr;
    CK(pool_alloc(&thr_agg, (size_t)n_blk * SCANB_T * sizeof(AffBS)));
    CK(pool_alloc(&blk_agg, (size_t)n_blk * sizeof(AffBS)));
    k_dzndc_bs_scan_local<<<n_blk, SCANB_T>>>(v, n_elem, scale, thr_agg, blk_agg);
    k_dzndc_bs_scan_blocks<<<1, SCANB_T>>>(n_blk, blk_agg);
    k_dzndc_bs_scan_apply<<<n_blk, SCANB_T>>>(v, n_elem, scale, thr_agg, blk_agg, dm, de, L,
                                              d.xr_detect ? 1 : 0);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (dbg) { fprintf(stderr, "bs scan: kernels %.3f ms\n", now_ms() - tq); tq = now_ms(); }
    pool_free(thr_agg);
    pool_free(blk_agg);
    if (dbg) { fprintf(stderr, "bs scan: free %.3f ms\n", now_ms() - tq); tq = now_ms(); }
    /* periodic reference: the wrapped value goes to index 0 (perturbation.py:2440-2461) */
    const long long i = valid - 1;
    if (i == d.ref_order - 1) {
        double m4[4]; int e4[4] = {0, 0, 0, 0};
        for (int j = 0; j < 4; j++) {
            CK(cudaMemcpy(&m4[j], dm + j * L + i, sizeof(double), cudaMemcpyDeviceToHost));
            if (de) CK(cudaMemcpy(&e4[j], de + j * L + i, sizeof(int), cudaMemcpyDeviceToHost));
        }
        const double X = d.Zn_path[2 * i], Y = d.Zn_path[2 * i + 1];
        long long k = d.n_xr > 0 ? h_xr_find(d.ref_index_xr, d.n_xr, i) : -1;
        XF rx = (k >= 0) ? mkXF(d.ref_xr[k], d.ref_xr_e[k]) : to_xr(X);
        XF ry = (k >= 0) ? mkXF(d.refy_xr[k], d.refy_xr_e[k]) : to_xr(Y);
        XF fxx, fxy, fyx, fyy;
        h_bs_jac(d.flavor, rx, ry, fxx, fxy, fyx, fyy);
        XF a = mkXF(m4[0], e4[0]), b = mkXF(m4[1], e4[1]), c = mkXF(m4[2], e4[2]), dd = mkXF(m4[3], e4[3]);
        XF w[4] = {fxx * a + fxy * c + scale, fxx * b + fxy * dd, fyx * a + fyy * c,
                   fyx * b + fyy * dd - scale};
        for (int j = 0; j < 4; j++) {
            if (de) {
                CK(cudaMemcpy(dm + j * L, &w[j].m, sizeof(double), cudaMemcpyHostToDevice));
                CK(cudaMemcpy(de + j * L, &w[j].e, sizeof(int), cudaMemcpyHostToDevice));
            } else {
                const double ws = to_std(w[j]);
                CK(cudaMemcpy(dm + j * L, &ws, sizeof(double), cudaMemcpyHostToDevice));
            }
        }
    }
    return 0;
}

/* flushed fp64 mirror of an Xrange table, built on the device */
int gpu_flush_mirror(fsb_frame *f, const double *m, const int *e, long long n, int comps,
                     long long pad, const double **out)
{
    double *o = nullptr;
    if (dev_zeros(f, (n + pad) * comps, &o)) return -1;
    const long long tot = n * comps;
    k_flush_mirror<<<(int)((tot + 255) / 256), 256>>>(n, m, e, comps, o);
    CK(cudaGetLastError());
    *out = o;
    return 0;
}

int frame_nz(const fsb_frame_desc &d)
{
    if (d.model == FSB_MODEL_M2)
        return 1 + (d.calc_dzndz ? 1 : 0) + (d.calc_dzndc ? 1 : 0) + (d.calc_orbit ? 1 : 0);
    return 2 + (d.calc_dzndc ? 4 : 0) + (d.calc_orbit ? 2 : 0);
}

} /* namespace */

/* ======================================================================== */
extern "C" {

const char *fsb_last_error(void) { return g_err.c_str(); }

const char *fsb_build_info(void)
{
#ifdef FSB_STRICT
    return "fsb200 sm_100a fmad=off (IEEE-strict build)";
#else
    return "fsb200 sm_100a fmad=on (default build)";
#endif
}

int fsb_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        fail(-1, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return -1;
    }
    return n;
}

int fsb_init(int device)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(-2, "no CUDA device available (%s): libfsb200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(-2, "device %d out of range (0..%d)", device, n - 1);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(-2, "device %s is sm_%d%d; this library is built for sm_100a only",
                    prop.name, prop.major, prop.minor);
    g_device = device;
    g_sm_count = prop.multiProcessorCount;
    return 0;
}

void fsb_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_flush_buf) { cudaFree(g_flush_buf); g_flush_buf = nullptr; }
    pool_trim();
    if (t_ctx) { delete t_ctx; t_ctx = nullptr; }
    g_device = -1;
}

int fsb_device_info(char *name, int name_cap, int *sm_count, int64_t *mem_bytes)
{
    if (ensure_init() != 0) return -1;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g_device));
    if (name && name_cap > 0) snprintf(name, (size_t)name_cap, "%s", prop.name);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (mem_bytes) *mem_bytes = (int64_t)prop.totalGlobalMem;
    return 0;
}

void *fsb_host_alloc(int64_t bytes)
{
    if (ensure_init() != 0) return nullptr;
    void *p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
        fail(-1, "cudaHostAlloc(%lld) failed", (long long)bytes);
        return nullptr;
    }
    return p;
}
void fsb_host_free(void *p) { if (p) cudaFreeHost(p); }
void *fsb_dev_alloc(int64_t bytes)
{
    if (ensure_init() != 0) return nullptr;
    void *p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) {
        fail(-1, "cudaMalloc(%lld) failed", (long long)bytes);
        return nullptr;
    }
    return p;
}
void fsb_dev_free(void *p) { if (p) cudaFree(p); }
int fsb_memcpy_h2d(void *dst, const void *src, int64_t bytes)
{
    if (ensure_init() != 0) return -1;
    CK(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice));
    return 0;
}
int fsb_memcpy_d2h(void *dst, const void *src, int64_t bytes)
{
    if (ensure_init() != 0) return -1;
    CK(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return 0;
}
int fsb_dev_memset(void *dst, int value, int64_t bytes)
{
    if (ensure_init() != 0) return -1;
    CK(cudaMemset(dst, value, (size_t)bytes));
    return 0;
}
int fsb_flush_l2(void)
{
    if (ensure_init() != 0) return -1;
    if (!g_flush_buf) {
        CK(cudaMalloc(&g_flush_buf, (size_t)(FLUSH_DOUBLES * 8)));
        CK(cudaMemset(g_flush_buf, 0, (size_t)(FLUSH_DOUBLES * 8)));
    }
    k_flush<<<g_sm_count * 8, 256>>>(g_flush_buf, FLUSH_DOUBLES);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

/* ---- standard loops ------------------------------------------------------ */
int fsb_std_nz(const fsb_std_desc *d)
{
    if (d->model == FSB_MODEL_M2) return 3 + (d->calc_d2zndc2 ? 1 : 0) + (d->calc_orbit ? 1 : 0);
    return 6 + (d->calc_orbit ? 2 : 0);
}

static int proj_fill(const fsb_proj_desc &d, ProjDev &p, bool allow_modifier)
{
    if (d.kind != FSB_PROJ_CARTESIAN && d.kind != FSB_PROJ_EXPMAP)
        return fail(-3, "unsupported projection kind %d (no fallback)", d.kind);
    if (d.dzndc_modifier < FSB_DZNDC_MOD_NONE || d.dzndc_modifier > FSB_DZNDC_MOD_SEAM)
        return fail(-3, "unsupported dzndc modifier %d", d.dzndc_modifier);
    if (d.dzndc_modifier != FSB_DZNDC_MOD_NONE && !allow_modifier)
        return fail(-3, "dzndc modifier is only defined for perturbation frames");
    p.kind = d.kind; p.mod_kind = d.dzndc_modifier;
    p.hmoy = d.hmoy; p.k_re = d.pix_to_ht[0]; p.k_im = d.pix_to_ht[1];
    p.mod_param = d.mod_param;
    return 0;
}

static int std_fill(const fsb_std_desc *d, StdDev &p, long long zstride)
{
    if (d->model != FSB_MODEL_M2 && d->model != FSB_MODEL_BS)
        return fail(-3, "unsupported standard model %d", d->model);
    if (d->model == FSB_MODEL_BS && (d->flavor < 1 || d->flavor > 5))
        return fail(-3, "unsupported burning-ship flavor %d", d->flavor);
    if (d->calc_orbit && d->backshift <= 0) return fail(-3, "calc_orbit needs backshift > 0");
    if (d->nexp != 0) {
        if (d->model != FSB_MODEL_M2) return fail(-3, "nexp is only defined for the Mandelbrot model");
        if (d->nexp < 2 || d->nexp > 32) return fail(-3, "exponent %d out of the supported range [2, 32]", d->nexp);
    }
    p.nexp = d->nexp;
    p.center_re = d->center_re; p.center_im = d->center_im; p.dx = d->dx;
    for (int i = 0; i < 4; i++) p.lin_mat[i] = d->lin_mat[i];
    p.max_iter = d->max_iter; p.Mdiv_sq = d->M_divergence_sq; p.eps_sq = d->epsilon_stationnary_sq;
    p.calc_d2 = d->calc_d2zndc2; p.calc_orbit = d->calc_orbit; p.backshift = d->backshift;
    p.flavor = d->flavor;
    p.zstride = zstride;
    ProjDev chk;
    return proj_fill(d->proj, chk, false);
}

/* enqueue one kernel over the units [unit_lo, unit_hi) of the call's point list
 * (whole-list pointers; p.zstride = npts of the list) */
static int std_enqueue(Ctx *c, const fsb_std_desc *d, const StdDev &p, cudaStream_t st, int slot,
                       const Units &u, int unit_lo, int unit_hi, const C *d_c_pix, double *d_Z,
                       signed char *d_sr, int *d_si)
{
    unsigned long long *ctl = c->d_ctl + slot * CTL_WORDS;
    CK(cudaMemsetAsync(ctl, 0, CTL_WORDS * sizeof(unsigned long long), st));
    const int block = 256;
    int grid = 1;
    const long long n = 32LL * (unit_hi - unit_lo);
    if (d->model == FSB_MODEL_M2 && d->nexp != 0) { if (persistent_grid(k_std_mn, block, n, &grid)) return -1; }
    else if (d->model == FSB_MODEL_M2) { if (persistent_grid(k_std_m2, block, n, &grid)) return -1; }
    else { if (persistent_grid(k_std_bs, block, n, &grid)) return -1; }
    const Tiling t = tiling_of(u, unit_lo, unit_hi);
    ProjDev P;
    long long p_lo, p_hi;
    if (proj_fill(d->proj, P, false)) return -1;
    if (enqueue_projection(c, P, u, unit_lo, unit_hi, st, &d_c_pix, p_lo, p_hi)) return -1;
    if (d->model == FSB_MODEL_M2 && d->nexp != 0)
        k_std_mn<<<grid, block, 0, st>>>(p, u.npts, d_c_pix, d_Z, d_sr, d_si, ctl, ctl + 1,
                                         c->d_abort(), t);
    else if (d->model == FSB_MODEL_M2)
        k_std_m2<<<grid, block, 0, st>>>(p, u.npts, d_c_pix, d_Z, d_sr, d_si, ctl, ctl + 1,
                                         c->d_abort(), t);
    else
        k_std_bs<<<grid, block, 0, st>>>(p, u.npts, d_c_pix, d_Z, d_sr, d_si, ctl, ctl + 1,
                                         c->d_abort(), t);
    CK(cudaGetLastError());
    return 0;
}

static void gather_stats(Ctx *c, int n_slots, fsb_stats *stats)
{
    if (!stats) return;
    stats->n_iter_exec = stats->n_bla_steps = stats->n_rebase = stats->sum_stop_iter = 0;
    stats->n_iter_fast = 0;
    for (int k = 0; k < n_slots; k++) {
        const unsigned long long *h = c->h_ctl + k * CTL_WORDS;
        stats->n_iter_exec += (int64_t)h[1];
        stats->n_bla_steps += (int64_t)h[2];
        stats->n_rebase += (int64_t)h[3];
        stats->sum_stop_iter += (int64_t)h[4];
        stats->n_iter_fast += (int64_t)h[5];
    }
    stats->n_launches = n_slots;
}

} /* extern "C" */

static int std_run_device_impl(Ctx *c, const fsb_std_desc *d, const Units &u, const double *d_c_pix,
                               double *d_Z, int8_t *d_stop_reason, int32_t *d_stop_iter,
                               fsb_stats *stats)
{
    StdDev p;
    if (std_fill(d, p, u.npts)) return -3;
    CK(cudaMemsetAsync(c->d_abort(), 0, sizeof(int), c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    if (std_enqueue(c, d, p, c->stream, 0, u, 0, u.n_units, (const C *)d_c_pix, d_Z,
                    (signed char *)d_stop_reason, d_stop_iter)) return -1;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaMemcpyAsync(c->h_ctl, c->d_ctl, CTL_WORDS * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    gather_stats(c, 1, stats);
    if (stats) stats->kernel_ms = ms;
    return 0;
}

extern "C" {

int fsb_std_run_device(const fsb_std_desc *d, int64_t npts, const double *d_c_pix, double *d_Z,
                       int8_t *d_stop_reason, int32_t *d_stop_iter, fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (stats) memset(stats, 0, sizeof *stats);
    if (npts <= 0) return 0;
    Units u;
    if (units_flat(npts, u)) return -3;
    return std_run_device_impl(c, d, u, d_c_pix, d_Z, d_stop_reason, d_stop_iter, stats);
}

int fsb_std_run_tiles_device(const fsb_std_desc *d, int32_t n_tiles, const int32_t *tile_w,
                             const int32_t *tile_h, const double *d_c_pix, double *d_Z,
                             int8_t *d_stop_reason, int32_t *d_stop_iter, fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (stats) memset(stats, 0, sizeof *stats);
    Units u;
    if (units_tiled(c, n_tiles, tile_w, tile_h, c->stream, u)) return -3;
    return std_run_device_impl(c, d, u, d_c_pix, d_Z, d_stop_reason, d_stop_iter, stats);
}

} /* extern "C" */

/* One output plane of a host-buffer call */
struct Plane { char *host; long long dev_off; long long elem; };

/* Host-buffer call: H2D of c_pix, kernels and D2H of the planes, pipelined
 * over slabs of consecutive points.  `enqueue(stream, slot, unit_lo, unit_hi,
 * a, n)` launches the kernel for those units (= points [a, a+n)) into control
 * block `slot`. */
/* Grid calls: the pixel offsets come from per-tile axes (tile k: tile_w[k] x
 * values then tile_h[k] y values; pix(r, col) = x[col] + i y[r], the layout of
 * Fractal.chunk_pixel_pos without jitter, core.py:1767-1830) instead of a
 * 16-byte-per-point array: a few hundred kB cross PCIe instead of 133 MB per 4K
 * frame, and k_expand_grid writes the c_pix plane of each slab in HBM. */
static int upload_axes(Ctx *c, const Units &u, const double *axes, cudaStream_t st,
                       const long long **d_off, const double **d_ax)
{
    if (!u.tiled()) return fail(-3, "grid calls need a tile list");
    const size_t nt = u.tiles.size();
    std::vector<long long> off(nt);
    long long tot = 0;
    for (size_t k = 0; k < nt; k++) { off[k] = tot; tot += (long long)u.tiles[k].z + u.tiles[k].w; }
    const long long bytes = (long long)nt * 8 + tot * 8;
    if (bytes > c->axes_cap) {
        if (c->d_axes) { CK(cudaFree(c->d_axes)); c->d_axes = nullptr; c->axes_cap = 0; }
        CK(cudaMalloc(&c->d_axes, (size_t)(2 * bytes)));
        c->axes_cap = 2 * bytes;
    }
    /* pageable sources: the driver stages them before returning */
    CK(cudaMemcpyAsync(c->d_axes, off.data(), nt * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_axes + nt * 8, axes, (size_t)(tot * 8), cudaMemcpyHostToDevice, st));
    *d_off = (const long long *)c->d_axes;
    *d_ax = (const double *)(c->d_axes + nt * 8);
    return 0;
}

template <class Enqueue>
static int run_pipelined(Ctx *c, const Units &u, const double *c_pix, const double *axes, long long o_c,
                         const Plane *planes, int n_planes, long long o_zero_beg,
                         long long o_zero_end, long long o_sr, Enqueue enqueue,
                         const volatile uint8_t *interrupted, fsb_stats *stats, bool *was_int)
{
    char *base = (char *)c->d_buf;
    const long long npts = u.npts;
    const Chunks ch = make_chunks(u);
    CK(cudaMemsetAsync(c->d_abort(), 0, sizeof(int), c->stream));
    CK(cudaMemsetAsync(base + o_zero_beg, 0, (size_t)(o_zero_end - o_zero_beg), c->stream));
    CK(cudaMemsetAsync(base + o_sr, 0xff, (size_t)npts, c->stream));
    const long long *d_ax_off = nullptr;
    const double *d_ax = nullptr;
    if (!c_pix) {
        if (!axes) return fail(-3, "no pixel source");
        if (upload_axes(c, u, axes, c->stream, &d_ax_off, &d_ax)) return -1;
    }
    CK(cudaEventRecord(c->ev0, c->stream));
    CK(cudaStreamWaitEvent(c->s_alt, c->ev0, 0));
    CK(cudaStreamWaitEvent(c->s_d2h, c->ev0, 0));
    /* Slab k+1 (H2D + kernel) is enqueued before the host waits for slab k, so
     * the device never idles; the host polls the interruption flag while it
     * waits (with pageable host buffers the "async" copies block the host, which
     * is why the D2H of slab k is only issued once its kernel has finished). */
    auto enqueue_slab = [&](int k) -> int {
        const long long a = ch.beg[k], n = ch.beg[k + 1] - a;
        cudaStream_t st = (k & 1) ? c->s_alt : c->stream;
        if (c_pix) {
            CK(cudaMemcpyAsync(base + o_c + a * 16, c_pix + 2 * a, (size_t)(n * 16),
                               cudaMemcpyHostToDevice, c->s_h2d));
            CK(cudaEventRecord(c->ev_h[k], c->s_h2d));
            CK(cudaStreamWaitEvent(st, c->ev_h[k], 0));
        } else if (n > 0) {
            k_expand_grid<<<(int)((n + 255) / 256), 256, 0, st>>>(
                u.d_tiles, (int)u.tiles.size(), d_ax_off, d_ax, a, n, (C *)(base + o_c));
            CK(cudaGetLastError());
        }
        if (enqueue(st, k, ch.ubeg[k], ch.ubeg[k + 1], a, n)) return -1;
        CK(cudaEventRecord(c->ev_k[k], st));
        return 0;
    };
    if (enqueue_slab(0)) return -1;
    for (int k = 0; k < ch.n; k++) {
        if (k + 1 < ch.n && enqueue_slab(k + 1)) return -1;
        bool wi = false;
        if (wait_event(c, c->ev_k[k], interrupted, &wi)) return -1;
        *was_int = *was_int || wi;
        const long long a = ch.beg[k], n = ch.beg[k + 1] - a;
        for (int r = 0; r < n_planes; r++)
            CK(cudaMemcpyAsync(planes[r].host + a * planes[r].elem,
                               base + planes[r].dev_off + a * planes[r].elem,
                               (size_t)(n * planes[r].elem), cudaMemcpyDeviceToHost, c->s_d2h));
    }
    CK(cudaMemcpyAsync(c->h_ctl, c->d_ctl, (size_t)ch.n * CTL_WORDS * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, c->s_d2h));
    CK(cudaEventRecord(c->ev1, c->s_d2h));
    CK(cudaStreamSynchronize(c->s_d2h));
    gather_stats(c, ch.n, stats);
    if (stats) {
        float ms_k = 0, ms_k2 = 0, ms_all = 0;
        CK(cudaEventElapsedTime(&ms_k, c->ev0, c->ev_k[ch.n - 1]));
        if (ch.n > 1) CK(cudaEventElapsedTime(&ms_k2, c->ev0, c->ev_k[ch.n - 2]));
        CK(cudaEventElapsedTime(&ms_all, c->ev0, c->ev1));
        stats->kernel_ms = ms_k > ms_k2 ? ms_k : ms_k2;   /* span of the slab kernels */
        stats->h2d_ms = 0.;                               /* hidden behind the kernels */
        stats->d2h_ms = ms_all - stats->kernel_ms;        /* exposed tail of the last slab */
    }
    return 0;
}

extern "C" {

} /* extern "C" */

static int std_run_impl(Ctx *c, const fsb_std_desc *d, int32_t n_tiles, const int32_t *tile_w,
                        const int32_t *tile_h, int64_t npts_flat, const double *c_pix, double *Z,
                        int8_t *stop_reason, int32_t *stop_iter,
                        const volatile uint8_t *interrupted, fsb_stats *stats,
                        const double *axes = nullptr)
{
    if (stats) memset(stats, 0, sizeof *stats);
    if (interrupted && *interrupted) return FSB_USER_INTERRUPTED;
    Units u;
    if (n_tiles > 0) { if (units_tiled(c, n_tiles, tile_w, tile_h, c->stream, u)) return -3; }
    else if (units_flat(npts_flat, u)) return -3;
    const long long npts = u.npts;
    StdDev p;
    if (std_fill(d, p, npts)) return -3;
    const int nz = fsb_std_nz(d);
    const long long zelem = (d->model == FSB_MODEL_M2) ? 16 : 8;
    long long o_c = 0, o_Z = align256(o_c + npts * 16), o_si = align256(o_Z + nz * npts * zelem),
              o_sr = align256(o_si + npts * 4), total = align256(o_sr + npts);
    if (ctx_reserve(c, total)) return -1;
    char *base = (char *)c->d_buf;
    Plane planes[16];
    int np = 0;
    for (int r = 0; r < nz; r++)
        planes[np++] = Plane{(char *)Z + r * npts * zelem, o_Z + r * npts * zelem, zelem};
    planes[np++] = Plane{(char *)stop_iter, o_si, 4};
    planes[np++] = Plane{(char *)stop_reason, o_sr, 1};
    bool was_int = false;
    auto enqueue = [&](cudaStream_t st, int slot, int unit_lo, int unit_hi, long long, long long) {
        return std_enqueue(c, d, p, st, slot, u, unit_lo, unit_hi, (const C *)(base + o_c),
                           (double *)(base + o_Z), (signed char *)(base + o_sr),
                           (int *)(base + o_si));
    };
    int rc = run_pipelined(c, u, c_pix, axes, o_c, planes, np, o_Z, o_si + npts * 4, o_sr, enqueue,
                           interrupted, stats, &was_int);
    if (rc) return rc;
    return was_int ? FSB_USER_INTERRUPTED : 0;
}

extern "C" {

int fsb_std_run(const fsb_std_desc *d, int64_t npts, const double *c_pix, double *Z,
                int8_t *stop_reason, int32_t *stop_iter, const volatile uint8_t *interrupted,
                fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (npts <= 0) { if (stats) memset(stats, 0, sizeof *stats); return 0; }
    return std_run_impl(c, d, 0, nullptr, nullptr, npts, c_pix, Z, stop_reason, stop_iter,
                        interrupted, stats);
}

int fsb_std_run_tiles(const fsb_std_desc *d, int32_t n_tiles, const int32_t *tile_w,
                      const int32_t *tile_h, const double *c_pix, double *Z, int8_t *stop_reason,
                      int32_t *stop_iter, const volatile uint8_t *interrupted, fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (n_tiles <= 0) return fail(-3, "empty tile list");
    return std_run_impl(c, d, n_tiles, tile_w, tile_h, 0, c_pix, Z, stop_reason, stop_iter,
                        interrupted, stats);
}

/* ---- perturbation frames -------------------------------------------------- */
int fsb_frame_create(const fsb_frame_desc *desc, fsb_frame **out)
{
    if (ensure_init() != 0) return -1;
    CK(cudaSetDevice(g_device));
    if (!desc || !out) return fail(-3, "null argument");
    *out = nullptr;
    if (desc->model != FSB_MODEL_M2 && desc->model != FSB_MODEL_BS)
        return fail(-3, "unsupported model %d (no fallback)", desc->model);
    if (desc->model == FSB_MODEL_BS && (desc->flavor < 1 || desc->flavor > 5))
        return fail(-3, "unsupported burning-ship flavor %d", desc->flavor);
    if (desc->L < 2 || !desc->Zn_path) return fail(-3, "reference orbit missing or too short");
    if (desc->calc_orbit && desc->backshift <= 0) return fail(-3, "calc_orbit needs backshift > 0");
    if (desc->model == FSB_MODEL_BS && desc->calc_dzndz)
        return fail(-3, "interior detection is not defined for the burning-ship family");
    if (desc->n_xr > 0 && (!desc->ref_index_xr || !desc->ref_xr || !desc->ref_xr_e))
        return fail(-3, "Xrange reference arrays missing");
    /* orbit indices and iteration counts are 32-bit in the kernels (stop_iter,
     * U and ref_index_xr are int32 in the reference too) */
    if (desc->L >= (1LL << 30) || desc->max_iter >= (1LL << 30) || desc->max_iter < 1)
        return fail(-3, "orbit length / max_iter out of the supported range (< 2^30)");
    if (desc->ref_order < 1) return fail(-3, "ref_order must be >= 1");
    if (desc->nexp != 0) {
        if (desc->model != FSB_MODEL_M2) return fail(-3, "nexp is only defined for the holomorphic model");
        if (desc->nexp < 2 || desc->nexp > 32) return fail(-3, "exponent %d out of the supported range [2, 32]", desc->nexp);
    }
    {
        ProjDev chk;
        if (proj_fill(desc->proj, chk, true)) return -3;
    }

    fsb_frame *f = new fsb_frame();
    f->d = *desc;
    const fsb_frame_desc &d = f->d;
    FrameDev &v = f->dev;
    memset(&v, 0, sizeof v);
    const long long L = d.L;
    double t0 = now_ms();
    int rc = 0;
#define UP(expr) do { if ((rc = (expr)) != 0) { fsb_frame_destroy(f); return rc; } } while (0)
    v.L = L;
    UP(upload(f, (const C *)d.Zn_path, L, &v.Zn, 1));
    v.nexp = d.nexp;
    if (d.nexp > 0) {            /* comb(N, k) as float64, mandelbrot_Mn.py:636-639 */
        double cb[33];
        cb[0] = 1.;
        for (int k = 1; k <= d.nexp; k++) cb[k] = cb[k - 1] * (double)(d.nexp - k + 1) / (double)k;
        UP(upload(f, cb, (long long)d.nexp + 1, &v.cbinom));
    }
    v.n_xr = d.n_xr;
    UP(upload(f, d.ref_index_xr, d.n_xr, &v.ref_index_xr));
    if (d.model == FSB_MODEL_M2) {
        UP(upload(f, (const C *)d.ref_xr, d.n_xr, &v.ref_xr));
        UP(upload(f, d.ref_xr_e, d.n_xr, &v.ref_xr_e));
    } else {
        UP(upload(f, d.ref_xr, d.n_xr, &v.refx_xr));
        UP(upload(f, d.ref_xr_e, d.n_xr, &v.refx_xr_e));
        UP(upload(f, d.refy_xr, d.n_xr, &v.refy_xr));
        UP(upload(f, d.refy_xr_e, d.n_xr, &v.refy_xr_e));
    }
    v.ref_div_iter = d.ref_div_iter;
    v.ref_order = d.ref_order;
    v.drift[0] = d.drift[0]; v.drift[1] = d.drift[1];
    v.drift_e[0] = d.drift_e[0]; v.drift_e[1] = d.drift_e[1];
    v.lin_scale = d.lin_scale; v.lin_scale_e = d.lin_scale_e;
    for (int i = 0; i < 4; i++) v.lin_mat[i] = d.lin_mat[i];
    proj_fill(d.proj, f->proj, true);
    v.max_iter = d.max_iter;
    v.Mdiv_sq = d.M_divergence_sq;
    v.eps_sq = d.epsilon_stationnary_sq;
    v.calc_orbit = d.calc_orbit;
    v.backshift = d.backshift;
    v.flavor = d.flavor;
    {
        const long long big = (1LL << 30);
        v.Li = (int)L;
        v.ref_div_i = (int)(d.ref_div_iter < big ? d.ref_div_iter : big);
        v.ref_div_m1_i = v.ref_div_i - 1;
        v.order_i = (d.ref_order < big) ? (int)d.ref_order : 0;
        long long fi = L;
        if (d.ref_div_iter < fi) fi = d.ref_div_iter;
        if (d.ref_order < fi) fi = d.ref_order;
        v.first_invalid_i = (int)fi;
        v.max_iter_i = (int)d.max_iter;
        v.n_xr_i = (int)d.n_xr;
    }
    f->nz = frame_nz(d);
    f->ms_upload = now_ms() - t0;

    /* reference derivative paths */
    t0 = now_ms();
    {
#ifdef FSB_STRICT
        f->gpu_scan = false;              /* serial host loop: bit-exact with the oracle */
#else
        const char *host = getenv("FSB200_HOST_DZNDC");
        f->gpu_scan = d.model == FSB_MODEL_M2 && !(host && host[0] == '1');
        f->gpu_scan_bs = d.model == FSB_MODEL_BS && !(host && host[0] == '1');
#endif
    }
    if (d.calc_dzndc) {
        if (d.model == FSB_MODEL_M2) {
            if (d.dZndc) {
                UP(upload(f, (const C *)d.dZndc, L, &v.dZndc, 1));
                if (d.xr_detect) UP(upload(f, d.dZndc_e, L, &v.dZndc_e, 1));
            } else if (f->gpu_scan) {
                UP(gpu_dzndc_m2(f));
            } else {
                std::vector<C> p; std::vector<int32_t> pe;
                host_dzndc_m2(d, p, pe);
                UP(upload(f, p.data(), L, &v.dZndc, 1));
                if (d.xr_detect) UP(upload(f, (const int *)pe.data(), L, &v.dZndc_e, 1));
            }
        } else if (!d.dZndc && f->gpu_scan_bs) {
            UP(gpu_dzndc_bs(f));
        } else {
            std::vector<double> p; std::vector<int32_t> pe;
            const double *src = d.dZndc; const int32_t *srce = d.dZndc_e;
            if (!src) { host_dzndc_bs(d, p, pe); src = p.data(); srce = pe.data(); }
            const double *dp = nullptr; const int *dpe = nullptr;
            UP(upload(f, src, 4 * L, &dp));
            if (d.xr_detect) UP(upload(f, (const int *)srce, 4 * L, &dpe));
            for (int j = 0; j < 4; j++) {
                v.dP[j] = dp + j * L;
                v.dP_e[j] = dpe ? dpe + j * L : nullptr;
            }
        }
    }
    {
        const char *pure = getenv("FSB200_PURE_XR");
        f->fast_xr = d.xr_detect && !(pure && pure[0] == '1')
                     && ((d.model == FSB_MODEL_M2 && !d.calc_dzndz && d.nexp == 0)
                         || (d.model == FSB_MODEL_BS && d.flavor <= 3));
    }
    if (f->fast_xr && d.calc_dzndc && d.model == FSB_MODEL_BS) {
        const double *dp = nullptr;
        UP(gpu_flush_mirror(f, v.dP[0], v.dP_e[0], 4 * L, 1, 0, &dp));
        for (int j = 0; j < 4; j++) v.dP_std[j] = dp + j * L;
    }
    if (f->fast_xr && d.calc_dzndc && d.model == FSB_MODEL_M2) {
        const double *dp = nullptr;
        UP(gpu_flush_mirror(f, (const double *)v.dZndc, v.dZndc_e, L, 2, 1, &dp));
        v.dZndc_std = (const C *)dp;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { fsb_frame_destroy(f); return fail(-1, "derivative path kernels failed"); }
    if (d.calc_dzndz) {
        if (d.dZndz) {
            UP(upload(f, (const C *)d.dZndz, L + 1, &v.dZndz));
            if (d.xr_detect) UP(upload(f, d.dZndz_e, L + 1, &v.dZndz_e));
        } else {
            std::vector<C> p; std::vector<int32_t> pe;
            host_dzndz_m2(d, p, pe);
            UP(upload(f, p.data(), L + 1, &v.dZndz));
            if (d.xr_detect) UP(upload(f, (const int *)pe.data(), L + 1, &v.dZndz_e));
        }
    }
    f->ms_dzndc = now_ms() - t0;

    /* BLA tree */
    f->bla_on = d.bla_activated != 0;
    if (f->bla_on) {
        if (d.M_bla && d.r_bla) {
            int width = (d.model == FSB_MODEL_M2) ? 4 : 8;
            UP(upload(f, d.M_bla, d.bla_len * width, &v.M_bla));
            UP(upload(f, d.r_bla, d.bla_len, &v.r_bla));
            v.bla_len = d.bla_len;
            v.stages_bla = d.stages_bla;
        } else {
            UP(build_bla(f));
        }
        if (v.stages_bla <= 3 || v.bla_len == 0) f->bla_on = false;
    }
    /* event-driven kernel: interleaved orbit table */
    {
        const char *old = getenv("FSB200_KERNEL_V1");
        f->v2 = d.model == FSB_MODEL_M2 && d.nexp == 0 && !d.calc_dzndz && !d.calc_orbit
                && v.order_i == 0 && (!d.xr_detect || f->fast_xr) && !(old && old[0] == '1');
    }
    if (f->v2 && f->bla_on) {
        int *t = nullptr;
        if (dev_zeros(f, 3 * v.bla_len, &t)) { fsb_frame_destroy(f); return -1; }
        k_bla_r2hi<<<(int)((v.bla_len + 255) / 256), 256>>>(v.bla_len, v.r_bla, t, t + v.bla_len,
                                                           t + 2 * v.bla_len);
        v.r2hi = t;
        v.r2hi_up = t + v.bla_len;
        v.rhi = t + 2 * v.bla_len;
    }
    if (f->v2) {
        const long long n_rec = L + 16 + FSB_STAGE_WIN /* a staged window may start at the last index */, n_h3 = FSB_H3_DIRECT ? n_rec : L / 8 + 4;
        void *p = nullptr, *ph = nullptr;
        if (pool_alloc(&p, (size_t)(n_rec * 32)) != cudaSuccess
            || (f->owned.push_back(p), pool_alloc(&ph, (size_t)(n_h3 * 4))) != cudaSuccess) {
            fsb_frame_destroy(f);
            return fail(-1, "out of device memory for the orbit table (%lld records)", n_rec);
        }
        f->owned.push_back(ph);
        const C *dsrc = d.calc_dzndc ? (d.xr_detect ? v.dZndc_std : v.dZndc) : nullptr;
        k_build_t2<<<(int)((n_rec + 255) / 256), 256>>>(n_rec, v.Zn, L + 1, dsrc, L + 1, FSB_TSCALE,
                                                          (double4 *)p);
        /* a lookup needs its leaf: indices 8 j with j < bla_len / 2 */
        const bool bla = f->bla_on && v.stages_bla >= 4;
        const long long n_leaf = bla ? v.bla_len / 2 : 0;
        if (cudaMemsetAsync(ph, 0, (size_t)(n_h3 * 4)) != cudaSuccess) { fsb_frame_destroy(f); return fail(-1, "memset failed"); }
        if (n_leaf > 0)
            k_build_h3<<<(int)((n_leaf + 255) / 256), 256>>>(n_leaf, n_h3, v.r_bla, v.first_invalid_i, (unsigned *)ph);
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            fsb_frame_destroy(f);
            return fail(-1, "orbit table kernel failed");
        }
        v.T2 = (const double *)p;
        v.h3 = (const unsigned *)ph;
        v.esc_hi = esc_word(esc_hi_of(v.Mdiv_sq));
    }
#undef UP
    /* the descriptor copy must not keep caller pointers alive */
    f->d.Zn_path = nullptr; f->d.ref_index_xr = nullptr; f->d.ref_xr = nullptr;
    f->d.ref_xr_e = nullptr; f->d.refy_xr = nullptr; f->d.refy_xr_e = nullptr;
    f->d.dZndc = nullptr; f->d.dZndc_e = nullptr; f->d.dZndz = nullptr; f->d.dZndz_e = nullptr;
    f->d.M_bla = nullptr; f->d.r_bla = nullptr;
    *out = f;
    return 0;
}

int fsb_frame_destroy(fsb_frame *f)
{
    if (!f) return 0;
    for (void *p : f->owned) pool_free(p);
    delete f;
    return 0;
}

int fsb_frame_nz(const fsb_frame *f) { return f ? f->nz : -1; }
int64_t fsb_frame_bla_len(const fsb_frame *f) { return f ? f->dev.bla_len : -1; }
int fsb_frame_stages_bla(const fsb_frame *f) { return f ? f->dev.stages_bla : -1; }
double fsb_frame_setup_ms(const fsb_frame *f, int what)
{
    if (!f) return -1.;
    return what == 0 ? f->ms_upload : (what == 1 ? f->ms_dzndc : f->ms_bla);
}

int fsb_frame_get_bla(const fsb_frame *f, double *M_bla, double *r_bla)
{
    if (!f || !f->dev.M_bla) return fail(-3, "frame has no BLA table");
    int width = (f->d.model == FSB_MODEL_M2) ? 4 : 8;
    CK(cudaMemcpy(M_bla, f->dev.M_bla, (size_t)(f->dev.bla_len * width * 8), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r_bla, f->dev.r_bla, (size_t)(f->dev.bla_len * 8), cudaMemcpyDeviceToHost));
    return 0;
}

int fsb_frame_get_dzndc(const fsb_frame *f, double *dZndc, int32_t *dZndc_e)
{
    if (!f) return fail(-3, "null frame");
    const long long L = f->d.L;
    if (f->d.model == FSB_MODEL_M2) {
        if (!f->dev.dZndc) return fail(-3, "frame has no dZndc path");
        CK(cudaMemcpy(dZndc, f->dev.dZndc, (size_t)(L * 16), cudaMemcpyDeviceToHost));
        if (f->dev.dZndc_e && dZndc_e)
            CK(cudaMemcpy(dZndc_e, f->dev.dZndc_e, (size_t)(L * 4), cudaMemcpyDeviceToHost));
    } else {
        if (!f->dev.dP[0]) return fail(-3, "frame has no dZndc path");
        CK(cudaMemcpy(dZndc, f->dev.dP[0], (size_t)(4 * L * 8), cudaMemcpyDeviceToHost));
        if (f->dev.dP_e[0] && dZndc_e)
            CK(cudaMemcpy(dZndc_e, f->dev.dP_e[0], (size_t)(4 * L * 4), cudaMemcpyDeviceToHost));
    }
    return 0;
}

int fsb_frame_get_dzndz(const fsb_frame *f, double *dZndz, int32_t *dZndz_e)
{
    if (!f || !f->dev.dZndz) return fail(-3, "frame has no dZndz path");
    const long long L = f->d.L;
    CK(cudaMemcpy(dZndz, f->dev.dZndz, (size_t)((L + 1) * 16), cudaMemcpyDeviceToHost));
    if (f->dev.dZndz_e && dZndz_e)
        CK(cudaMemcpy(dZndz_e, f->dev.dZndz_e, (size_t)((L + 1) * 4), cudaMemcpyDeviceToHost));
    return 0;
}

} /* extern "C" */

/* enqueue one kernel over the units [unit_lo, unit_hi) of the call's point list */
static int frame_enqueue(Ctx *c, fsb_frame *f, cudaStream_t st, int slot, const Units &u,
                         int unit_lo, int unit_hi, const C *d_c_pix, double *d_Z, int *d_U,
                         signed char *d_sr, int *d_si)
{
    const fsb_frame_desc &d = f->d;
    perturb_kernel_t k = (d.model == FSB_MODEL_M2)
        ? (d.nexp != 0
               ? pick_mn(d.xr_detect != 0, d.calc_dzndc != 0, d.calc_dzndz != 0, f->bla_on)
               : pick_m2(d.xr_detect != 0, d.calc_dzndc != 0, d.calc_dzndz != 0, f->bla_on,
                         f->dev.order_i > 0 || d.calc_orbit != 0, f->fast_xr, f->v2))
        : pick_bs(d.xr_detect != 0, d.calc_dzndc != 0, f->bla_on, f->fast_xr, d.flavor);
    unsigned long long *ctl = c->d_ctl + slot * CTL_WORDS;
    CK(cudaMemsetAsync(ctl, 0, CTL_WORDS * sizeof(unsigned long long), st));
    const int block = 128;
    int grid = 1;
    if (persistent_grid(k, block, 32LL * (unit_hi - unit_lo), &grid)) return -1;
    FrameDev dv = f->dev;
    dv.zstride = u.npts;
    const C *d_c_raw = d_c_pix;
    long long p_lo, p_hi;
    if (enqueue_projection(c, f->proj, u, unit_lo, unit_hi, st, &d_c_pix, p_lo, p_hi)) return -1;
    k<<<grid, block, 0, st>>>(dv, u.npts, d_c_pix, d_Z, d_U, d_sr, d_si, ctl, ctl + 1,
                              c->d_abort(), tiling_of(u, unit_lo, unit_hi));
    CK(cudaGetLastError());
    /* Z[dzndc] *= proj_dzndc_modifier(c_pix), perturbation.py:1387-1388, 1772-1776 */
    if (f->proj.mod_kind != FSB_DZNDC_MOD_NONE && d.calc_dzndc && p_hi > p_lo) {
        const int holo = (d.model == FSB_MODEL_M2);
        const int row0 = holo ? (1 + (d.calc_dzndz ? 1 : 0)) : 2;
        k_modifier_range<<<(int)((p_hi - p_lo + 255) / 256), 256, 0, st>>>(
            f->proj, p_lo, p_hi, d_c_raw, d_Z, u.npts, holo, row0);
        CK(cudaGetLastError());
    }
    return 0;
}

static int frame_run_device_impl(Ctx *c, fsb_frame *f, const Units &u, const double *d_c_pix,
                                 double *d_Z, int32_t *d_U, int8_t *d_stop_reason,
                                 int32_t *d_stop_iter, fsb_stats *stats)
{
    CK(cudaMemsetAsync(c->d_abort(), 0, sizeof(int), c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    if (frame_enqueue(c, f, c->stream, 0, u, 0, u.n_units, (const C *)d_c_pix, d_Z, d_U,
                      (signed char *)d_stop_reason, d_stop_iter)) return -1;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaMemcpyAsync(c->h_ctl, c->d_ctl, CTL_WORDS * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    gather_stats(c, 1, stats);
    if (stats) stats->kernel_ms = ms;
    return 0;
}

/* fsb_postproc_desc -> device parameters (row stride = npts of the call) */
static int postproc_fill(const fsb_postproc_desc *d, long long zstride, int n_rows, PostprocDev &p)
{
    if (!d) return fail(-3, "null post-processing description");
    const int need = d->holomorphic ? 1 : 2, need_d = d->holomorphic ? 1 : 4;
    if (d->row_zn < 0 || d->row_zn + need > n_rows
        || (d->row_dzndc >= 0 && d->row_dzndc + need_d > n_rows))
        return fail(-3, "post-processing rows (%d, %d) outside the %d rows of Z", d->row_zn,
                    d->row_dzndc, n_rows);
    if (!(d->potential_d > 1.) || !(d->potential_M > 0.) || d->potential_a_d == 0.)
        return fail(-3, "unsupported potential (d = %g, a_d = %g, M = %g)", d->potential_d,
                    d->potential_a_d, d->potential_M);
    p.holomorphic = d->holomorphic; p.row_zn = d->row_zn; p.row_d = d->row_dzndc;
    p.zstride = zstride;
    p.k = pow(fabs(d->potential_a_d), 1. / (d->potential_d - 1.));
    p.log_Mk = log(d->potential_M * p.k);
    p.inv_log_d = 1. / log(d->potential_d);
    p.floor_iter = d->floor_iter; p.px_snap = d->px_snap;
    p.has_skew = d->has_skew;
    for (int i = 0; i < 4; i++) p.skew[i] = d->skew[i];
    p.out_f64 = d->out_f64;
    if (d->df_kind < 0 || d->df_kind > 3) return fail(-3, "unknown projection derivative %d", d->df_kind);
    p.df_kind = d->row_dzndc >= 0 ? d->df_kind : 0;
    p.df_kre = d->df_k[0]; p.df_kim = d->df_k[1];
    return 0;
}

static int postproc_enqueue(const PostprocDev &p, cudaStream_t st, long long first, long long n,
                            const double *d_Z, const int *d_si, const C *d_c, void *d_nu,
                            void *d_dem, void *d_nx, void *d_ny)
{
    if (n <= 0 || (!d_nu && !d_dem && !d_nx)) return 0;
    if (p.df_kind != 0 && (d_dem || d_nx) && !d_c)
        return fail(-3, "the projection derivative needs the pixel offsets (c_pix)");
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)g_sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_postproc<<<(int)blocks, 256, 0, st>>>(p, first, n, d_Z, d_si, d_c, d_nu, d_dem, d_nx, d_ny);
    CK(cudaGetLastError());
    return 0;
}

static int postproc_ext_fill(const fsb_postproc_desc *d, const fsb_postproc_ext *x, int n_rows,
                             bool want_fl, bool want_shade, PostprocExtDev &e)
{
    if (!x) return fail(-3, "null field-lines / shading description");
    memset(&e, 0, sizeof e);
    if (want_fl) {
        if (x->fl_n_iter < 1 || x->fl_n_iter > FSB_PP_MAX_FL)
            return fail(-3, "field lines: n_iter = %d outside 1..%d", x->fl_n_iter, FSB_PP_MAX_FL);
        const int need = d->holomorphic ? 1 : 2;
        if (x->fl_row_orbit >= 0 && x->fl_row_orbit + need > n_rows)
            return fail(-3, "field lines: orbit row %d outside the %d rows of Z", x->fl_row_orbit, n_rows);
        const bool ok_model = d->holomorphic ? (x->fl_model >= 2 && x->fl_model <= 32)
                                             : (x->fl_model <= -1 && x->fl_model >= -5);
        if (!ok_model) return fail(-3, "field lines: model code %d does not fit the Z rows", x->fl_model);
        e.fl_n = x->fl_n_iter; e.fl_row_orbit = x->fl_row_orbit; e.fl_backshift = x->fl_backshift;
        e.fl_model = x->fl_model;
        for (int i = 0; i < x->fl_n_iter; i++) { e.fl_k[i] = x->fl_k[i]; e.fl_phi[i] = x->fl_phi[i]; }
        e.P.kind = x->proj_kind == FSB_PROJ_EXPMAP ? 1 : 0;
        e.P.hmoy = x->proj_hmoy; e.P.k_re = x->proj_k[0]; e.P.k_im = x->proj_k[1];
        e.cx = x->c_center[0]; e.cy = x->c_center[1]; e.cs = x->c_scale;
        for (int i = 0; i < 4; i++) e.cm[i] = x->c_lin_mat[i];
    }
    if (want_shade) {
        if (x->n_lights < 1 || x->n_lights > FSB_PP_MAX_LIGHTS)
            return fail(-3, "shading: %d light sources outside 1..%d", x->n_lights, FSB_PP_MAX_LIGHTS);
        if (d->row_dzndc < 0) return fail(-3, "shading needs the derivative rows (normal map)");
        e.n_lights = x->n_lights; e.ncoeff = x->normal_coeff;
        for (int l = 0; l < x->n_lights; l++)
            for (int i = 0; i < 8; i++) e.light[l][i] = x->light[l][i];
    }
    return 0;
}

static int postproc_ext_enqueue(const PostprocDev &p, const PostprocExtDev &e, cudaStream_t st,
                                long long first, long long n, const double *d_Z, const int *d_si,
                                const C *d_c, void *d_fl, void *d_shade, long long shade_stride)
{
    if (n <= 0 || (!d_fl && !d_shade)) return 0;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)g_sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_postproc_ext<<<(int)blocks, 256, 0, st>>>(p, e, first, n, d_Z, d_si, d_c, d_fl, d_shade,
                                                shade_stride);
    CK(cudaGetLastError());
    return 0;
}

/* pp == nullptr: the raw planes come back (Z, U, stop_iter, stop_reason).
 * pp != nullptr: fused post-processing -- the raw planes stay in HBM and only
 * the requested fields (pp_out: nu, dem, nx, ny) plus the non-null ones of
 * stop_reason / stop_iter are copied to the host. */
static int frame_run_impl(Ctx *c, fsb_frame *f, int32_t n_tiles, const int32_t *tile_w,
                          const int32_t *tile_h, int64_t npts_flat, const double *c_pix, double *Z,
                          int32_t *U, int8_t *stop_reason, int32_t *stop_iter,
                          const volatile uint8_t *interrupted, fsb_stats *stats,
                          const fsb_postproc_desc *pp = nullptr, void *const *pp_out = nullptr,
                          const double *axes = nullptr, const fsb_postproc_ext *ext = nullptr,
                          void *fl_out = nullptr, void *shade_out = nullptr)
{
    if (stats) memset(stats, 0, sizeof *stats);
    if (interrupted && *interrupted) return FSB_USER_INTERRUPTED;
    Units u;
    if (n_tiles > 0) { if (units_tiled(c, n_tiles, tile_w, tile_h, c->stream, u)) return -3; }
    else if (units_flat(npts_flat, u)) return -3;
    const long long npts = u.npts;
    const int nz = f->nz;
    const long long zelem = (f->d.model == FSB_MODEL_M2) ? 16 : 8;
    PostprocDev pd;
    long long oelem = 4;
    PostprocExtDev pe;
    const bool use_ext = pp && ext && (fl_out || shade_out);
    int n_shade = 0;
    if (pp) {
        if (postproc_fill(pp, npts, nz, pd)) return -3;
        oelem = pp->out_f64 ? 8 : 4;
        if (use_ext) {
            if (postproc_ext_fill(pp, ext, nz, fl_out != nullptr, shade_out != nullptr, pe)) return -3;
            n_shade = shade_out ? 2 * pe.n_lights : 0;
        }
    }
    const long long pp_plane = align256(npts * oelem);
    long long o_c = 0, o_Z = align256(o_c + npts * 16), o_U = align256(o_Z + nz * npts * zelem),
              o_si = align256(o_U + npts * 4), o_sr = align256(o_si + npts * 4),
              o_pp = align256(o_sr + npts),
              o_fl = o_pp + (pp ? 4 * pp_plane : 0), o_sh = o_fl + (use_ext ? pp_plane : 0),
              total = align256(o_sh + n_shade * pp_plane);
    if (ctx_reserve(c, total)) return -1;
    char *base = (char *)c->d_buf;
    Plane planes[16];
    int np = 0;
    void *d_pp[4] = {nullptr, nullptr, nullptr, nullptr};
    void *d_fl = nullptr, *d_sh = nullptr;
    if (!pp) {
        for (int r = 0; r < nz; r++)
            planes[np++] = Plane{(char *)Z + r * npts * zelem, o_Z + r * npts * zelem, zelem};
        planes[np++] = Plane{(char *)U, o_U, 4};
        planes[np++] = Plane{(char *)stop_iter, o_si, 4};
        planes[np++] = Plane{(char *)stop_reason, o_sr, 1};
    } else {
        for (int k = 0; k < 4; k++) {
            if (!pp_out || !pp_out[k]) continue;
            const long long off = o_pp + k * align256(npts * oelem);
            d_pp[k] = base + off;
            planes[np++] = Plane{(char *)pp_out[k], off, oelem};
        }
        if ((d_pp[2] == nullptr) != (d_pp[3] == nullptr))
            return fail(-3, "the normal needs both of its output arrays");
        if ((d_pp[1] || d_pp[2]) && pp->row_dzndc < 0)
            return fail(-3, "distance estimate / normal need the derivative rows");
        if (use_ext && fl_out) {
            d_fl = base + o_fl;
            planes[np++] = Plane{(char *)fl_out, o_fl, oelem};
        }
        if (use_ext && shade_out) {
            d_sh = base + o_sh;
            /* rows of the caller's (2 n_lights, npts) array <-> device planes of pitch pp_plane */
            for (int r = 0; r < n_shade; r++)
                planes[np++] = Plane{(char *)shade_out + r * npts * oelem, o_sh + r * pp_plane, oelem};
        }
        if (stop_iter) planes[np++] = Plane{(char *)stop_iter, o_si, 4};
        if (stop_reason) planes[np++] = Plane{(char *)stop_reason, o_sr, 1};
    }
    bool was_int = false;
    auto enqueue = [&](cudaStream_t st, int slot, int unit_lo, int unit_hi, long long a,
                       long long n) {
        if (frame_enqueue(c, f, st, slot, u, unit_lo, unit_hi, (const C *)(base + o_c),
                          (double *)(base + o_Z), (int *)(base + o_U),
                          (signed char *)(base + o_sr), (int *)(base + o_si))) return -1;
        if (pp && postproc_enqueue(pd, st, a, n, (const double *)(base + o_Z),
                                   (const int *)(base + o_si), (const C *)(base + o_c), d_pp[0],
                                   d_pp[1], d_pp[2], d_pp[3]))
            return -1;
        if (use_ext) return postproc_ext_enqueue(pd, pe, st, a, n, (const double *)(base + o_Z),
                                                 (const int *)(base + o_si), (const C *)(base + o_c),
                                                 d_fl, d_sh, pp_plane / oelem);
        return 0;
    };
    int rc = run_pipelined(c, u, c_pix, axes, o_c, planes, np, o_Z, o_si + npts * 4, o_sr, enqueue,
                           interrupted, stats, &was_int);
    if (rc) return rc;
    if (stats && pp) stats->n_launches *= use_ext ? 3 : 2;
    return was_int ? FSB_USER_INTERRUPTED : 0;
}

extern "C" {

int fsb_frame_run_device(fsb_frame *f, int64_t npts, const double *d_c_pix, double *d_Z,
                         int32_t *d_U, int8_t *d_stop_reason, int32_t *d_stop_iter,
                         fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (!f) return fail(-3, "null frame");
    if (stats) memset(stats, 0, sizeof *stats);
    if (npts <= 0) return 0;
    Units u;
    if (units_flat(npts, u)) return -3;
    return frame_run_device_impl(c, f, u, d_c_pix, d_Z, d_U, d_stop_reason, d_stop_iter, stats);
}

int fsb_frame_run_tiles_device(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w,
                               const int32_t *tile_h, const double *d_c_pix, double *d_Z,
                               int32_t *d_U, int8_t *d_stop_reason, int32_t *d_stop_iter,
                               fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (!f) return fail(-3, "null frame");
    if (stats) memset(stats, 0, sizeof *stats);
    Units u;
    if (units_tiled(c, n_tiles, tile_w, tile_h, c->stream, u)) return -3;
    return frame_run_device_impl(c, f, u, d_c_pix, d_Z, d_U, d_stop_reason, d_stop_iter, stats);
}

int fsb_frame_run(fsb_frame *f, int64_t npts, const double *c_pix, double *Z, int32_t *U,
                  int8_t *stop_reason, int32_t *stop_iter, const volatile uint8_t *interrupted,
                  fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (!f) return fail(-3, "null frame");
    if (npts <= 0) { if (stats) memset(stats, 0, sizeof *stats); return 0; }
    return frame_run_impl(c, f, 0, nullptr, nullptr, npts, c_pix, Z, U, stop_reason, stop_iter,
                          interrupted, stats);
}

int fsb_frame_run_tiles(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w,
                        const int32_t *tile_h, const double *c_pix, double *Z, int32_t *U,
                        int8_t *stop_reason, int32_t *stop_iter,
                        const volatile uint8_t *interrupted, fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (!f) return fail(-3, "null frame");
    if (n_tiles <= 0) return fail(-3, "empty tile list");
    return frame_run_impl(c, f, n_tiles, tile_w, tile_h, 0, c_pix, Z, U, stop_reason, stop_iter,
                          interrupted, stats);
}

/* ---- post-processing (SURVEY f-3) ------------------------------------------ */
int fsb_frame_run_pp(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w, const int32_t *tile_h,
                     int64_t npts, const double *c_pix, const fsb_postproc_desc *pp, void *nu,
                     void *dem, void *normal_x, void *normal_y, int8_t *stop_reason,
                     int32_t *stop_iter, const volatile uint8_t *interrupted, fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (!f) return fail(-3, "null frame");
    if (!pp) return fail(-3, "null post-processing description");
    /* the projection's df (projection.py:375-453) is part of the description (df_kind;
     * 0 is what the stepped flow without rotation uses) */
    if (n_tiles <= 0 && npts <= 0) { if (stats) memset(stats, 0, sizeof *stats); return 0; }
    void *outs[4] = {nu, dem, normal_x, normal_y};
    return frame_run_impl(c, f, n_tiles, tile_w, tile_h, npts, c_pix, nullptr, nullptr,
                          stop_reason, stop_iter, interrupted, stats, pp, outs);
}

/* ---- grid calls: pixel offsets from per-tile axes ---------------------------- */
int fsb_std_run_grid(const fsb_std_desc *d, int32_t n_tiles, const int32_t *tile_w,
                     const int32_t *tile_h, const double *axes, double *Z, int8_t *stop_reason,
                     int32_t *stop_iter, const volatile uint8_t *interrupted, fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (n_tiles <= 0 || !axes) return fail(-3, "empty tile list / null axes");
    return std_run_impl(c, d, n_tiles, tile_w, tile_h, 0, nullptr, Z, stop_reason, stop_iter,
                        interrupted, stats, axes);
}

int fsb_frame_run_grid(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w, const int32_t *tile_h,
                       const double *axes, double *Z, int32_t *U, int8_t *stop_reason,
                       int32_t *stop_iter, const volatile uint8_t *interrupted, fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (!f) return fail(-3, "null frame");
    if (n_tiles <= 0 || !axes) return fail(-3, "empty tile list / null axes");
    return frame_run_impl(c, f, n_tiles, tile_w, tile_h, 0, nullptr, Z, U, stop_reason, stop_iter,
                          interrupted, stats, nullptr, nullptr, axes);
}

int fsb_frame_run_grid_pp(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w,
                          const int32_t *tile_h, const double *axes, const fsb_postproc_desc *pp,
                          void *nu, void *dem, void *normal_x, void *normal_y,
                          int8_t *stop_reason, int32_t *stop_iter,
                          const volatile uint8_t *interrupted, fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (!f) return fail(-3, "null frame");
    if (!pp) return fail(-3, "null post-processing description");
    if (n_tiles <= 0 || !axes) return fail(-3, "empty tile list / null axes");
    void *outs[4] = {nu, dem, normal_x, normal_y};
    return frame_run_impl(c, f, n_tiles, tile_w, tile_h, 0, nullptr, nullptr, nullptr, stop_reason,
                          stop_iter, interrupted, stats, pp, outs, axes);
}

int fsb_frame_run_grid_pp_ext(fsb_frame *f, int32_t n_tiles, const int32_t *tile_w,
                              const int32_t *tile_h, const double *axes,
                              const fsb_postproc_desc *pp, const fsb_postproc_ext *ext,
                              void *nu, void *dem, void *normal_x, void *normal_y,
                              void *fieldlines, void *shade, int8_t *stop_reason,
                              int32_t *stop_iter, const volatile uint8_t *interrupted,
                              fsb_stats *stats)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (!f) return fail(-3, "null frame");
    if (!pp) return fail(-3, "null post-processing description");
    if ((fieldlines || shade) && !ext) return fail(-3, "null field-lines / shading description");
    if (fieldlines && (int)f->d.proj.kind != ext->proj_kind)
        return fail(-3, "field lines: the projection of the description differs from the frame's");
    if (n_tiles <= 0 || !axes) return fail(-3, "empty tile list / null axes");
    void *outs[4] = {nu, dem, normal_x, normal_y};
    return frame_run_impl(c, f, n_tiles, tile_w, tile_h, 0, nullptr, nullptr, nullptr, stop_reason,
                          stop_iter, interrupted, stats, pp, outs, axes, ext, fieldlines, shade);
}

int fsb_postproc_ext_run_device(const fsb_postproc_desc *pp, const fsb_postproc_ext *ext,
                                int64_t npts, int32_t n_rows, const double *d_Z,
                                const int32_t *d_stop_iter, const double *d_c_pix,
                                void *d_fieldlines, void *d_shade)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (npts <= 0) return 0;
    PostprocDev pd;
    PostprocExtDev pe;
    if (postproc_fill(pp, npts, n_rows, pd)) return -3;
    if (postproc_ext_fill(pp, ext, n_rows, d_fieldlines != nullptr, d_shade != nullptr, pe)) return -3;
    if (d_fieldlines && !d_c_pix) return fail(-3, "field lines need the pixel offsets");
    if (d_shade && pd.df_kind != 0 && !d_c_pix)
        return fail(-3, "the projection derivative needs the pixel offsets (c_pix)");
    if (postproc_ext_enqueue(pd, pe, c->stream, 0, npts, d_Z, d_stop_iter, (const C *)d_c_pix,
                             d_fieldlines, d_shade, npts)) return -1;
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int fsb_postproc_ext_run(const fsb_postproc_desc *pp, const fsb_postproc_ext *ext, int64_t npts,
                         int32_t n_rows, const double *Z, const int32_t *stop_iter,
                         const double *c_pix, void *fieldlines, void *shade)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (npts <= 0) return 0;
    if (!pp || !ext) return fail(-3, "null post-processing description");
    if (fieldlines && !c_pix) return fail(-3, "field lines need the pixel offsets");
    const int n_shade = shade ? 2 * ext->n_lights : 0;
    if (n_shade < 0 || n_shade > 2 * FSB_PP_MAX_LIGHTS) return fail(-3, "bad number of light sources");
    const long long zelem = pp->holomorphic ? 16 : 8, oelem = pp->out_f64 ? 8 : 4;
    const long long o_Z = 0, o_si = align256(n_rows * npts * zelem), o_c = align256(o_si + npts * 4),
                    o_fl = align256(o_c + npts * 16), o_sh = align256(o_fl + npts * oelem),
                    total = o_sh + align256(n_shade * npts * oelem);
    if (ctx_reserve(c, total)) return -1;
    char *base = (char *)c->d_buf;
    CK(cudaMemcpyAsync(base + o_Z, Z, (size_t)(n_rows * npts * zelem), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(base + o_si, stop_iter, (size_t)(npts * 4), cudaMemcpyHostToDevice, c->stream));
    if (c_pix) CK(cudaMemcpyAsync(base + o_c, c_pix, (size_t)(npts * 16), cudaMemcpyHostToDevice, c->stream));
    int rc = fsb_postproc_ext_run_device(pp, ext, npts, n_rows, (const double *)(base + o_Z),
                                         (const int32_t *)(base + o_si),
                                         c_pix ? (const double *)(base + o_c) : nullptr,
                                         fieldlines ? base + o_fl : nullptr,
                                         shade ? base + o_sh : nullptr);
    if (rc) return rc;
    if (fieldlines) CK(cudaMemcpy(fieldlines, base + o_fl, (size_t)(npts * oelem), cudaMemcpyDeviceToHost));
    if (shade) CK(cudaMemcpy(shade, base + o_sh, (size_t)(n_shade * npts * oelem), cudaMemcpyDeviceToHost));
    return 0;
}

int fsb_postproc_run_proj_device(const fsb_postproc_desc *pp, int64_t npts, int32_t n_rows,
                                 const double *d_Z, const int32_t *d_stop_iter,
                                 const double *d_c_pix, void *d_nu, void *d_dem,
                                 void *d_normal_x, void *d_normal_y)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (npts <= 0) return 0;
    PostprocDev pd;
    if (postproc_fill(pp, npts, n_rows, pd)) return -3;
    if ((d_dem || d_normal_x) && pp->row_dzndc < 0)
        return fail(-3, "distance estimate / normal need the derivative rows");
    if ((d_normal_x == nullptr) != (d_normal_y == nullptr))
        return fail(-3, "the normal needs both of its output arrays");
    if (postproc_enqueue(pd, c->stream, 0, npts, d_Z, d_stop_iter, (const C *)d_c_pix, d_nu, d_dem,
                         d_normal_x, d_normal_y)) return -3;
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int fsb_postproc_run_device(const fsb_postproc_desc *pp, int64_t npts, int32_t n_rows,
                            const double *d_Z, const int32_t *d_stop_iter, void *d_nu,
                            void *d_dem, void *d_normal_x, void *d_normal_y)
{
    return fsb_postproc_run_proj_device(pp, npts, n_rows, d_Z, d_stop_iter, nullptr, d_nu, d_dem,
                                        d_normal_x, d_normal_y);
}

int fsb_postproc_run_proj(const fsb_postproc_desc *pp, int64_t npts, int32_t n_rows,
                          const double *Z, const int32_t *stop_iter, const double *c_pix,
                          void *nu, void *dem, void *normal_x, void *normal_y)
{
    Ctx *c;
    if (get_ctx(&c)) return -1;
    if (npts <= 0) return 0;
    if (!pp) return fail(-3, "null post-processing description");
    const long long zelem = pp->holomorphic ? 16 : 8, oelem = pp->out_f64 ? 8 : 4;
    const long long o_Z = 0, o_si = align256(n_rows * npts * zelem), o_c = align256(o_si + npts * 4),
                    o_pp = align256(o_c + npts * 16), total = o_pp + 4 * align256(npts * oelem);
    if (ctx_reserve(c, total)) return -1;
    char *base = (char *)c->d_buf;
    CK(cudaMemcpyAsync(base + o_Z, Z, (size_t)(n_rows * npts * zelem), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(base + o_si, stop_iter, (size_t)(npts * 4), cudaMemcpyHostToDevice, c->stream));
    if (c_pix) CK(cudaMemcpyAsync(base + o_c, c_pix, (size_t)(npts * 16), cudaMemcpyHostToDevice, c->stream));
    void *host[4] = {nu, dem, normal_x, normal_y};
    void *dev[4];
    for (int k = 0; k < 4; k++) dev[k] = host[k] ? base + o_pp + k * align256(npts * oelem) : nullptr;
    int rc = fsb_postproc_run_proj_device(pp, npts, n_rows, (const double *)(base + o_Z),
                                          (const int32_t *)(base + o_si),
                                          c_pix ? (const double *)(base + o_c) : nullptr, dev[0],
                                          dev[1], dev[2], dev[3]);
    if (rc) return rc;
    for (int k = 0; k < 4; k++)
        if (host[k]) CK(cudaMemcpy(host[k], dev[k], (size_t)(npts * oelem), cudaMemcpyDeviceToHost));
    return 0;
}

int fsb_postproc_run(const fsb_postproc_desc *pp, int64_t npts, int32_t n_rows, const double *Z,
                     const int32_t *stop_iter, void *nu, void *dem, void *normal_x,
                     void *normal_y)
{
    return fsb_postproc_run_proj(pp, npts, n_rows, Z, stop_iter, nullptr, nu, dem, normal_x, normal_y);
}

/* ---- unit-test / calibration entry points --------------------------------- */
int fsb_xr_binop_c(int op, int64_t n, const double *a, const int32_t *ae, const double *b,
                   const int32_t *be, double *out, int32_t *oute)
{
    if (ensure_init() != 0) return -1;
    if (n <= 0) return 0;
    C *da, *db, *dout; int *dae, *dbe, *doe;
    CK(cudaMalloc(&da, n * 16)); CK(cudaMalloc(&db, n * 16)); CK(cudaMalloc(&dout, n * 16));
    CK(cudaMalloc(&dae, n * 4)); CK(cudaMalloc(&dbe, n * 4)); CK(cudaMalloc(&doe, n * 4));
    CK(cudaMemcpy(da, a, n * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b, n * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dae, ae, n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbe, be, n * 4, cudaMemcpyHostToDevice));
    k_xr_binop_c<<<(int)((n + 127) / 128), 128>>>(op, n, da, dae, db, dbe, dout, doe);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, dout, n * 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(oute, doe, n * 4, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dae); cudaFree(dbe); cudaFree(doe);
    return 0;
}

int fsb_xr_to_standard_c(int64_t n, const double *a, const int32_t *ae, double *out)
{
    if (ensure_init() != 0) return -1;
    if (n <= 0) return 0;
    C *da, *dout; int *dae;
    CK(cudaMalloc(&da, n * 16)); CK(cudaMalloc(&dout, n * 16)); CK(cudaMalloc(&dae, n * 4));
    CK(cudaMemcpy(da, a, n * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dae, ae, n * 4, cudaMemcpyHostToDevice));
    k_xr_to_standard_c<<<(int)((n + 127) / 128), 128>>>(n, da, dae, dout);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, dout, n * 16, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(dout); cudaFree(dae);
    return 0;
}

int fsb_proj_apply(const fsb_proj_desc *p, int64_t npts, const double *c_pix, double *out_pix,
                   double *out_modifier)
{
    if (ensure_init() != 0) return -1;
    if (!p) return fail(-3, "null argument");
    ProjDev P;
    if (proj_fill(*p, P, true)) return -3;
    if (npts <= 0) return 0;
    C *dp, *dq = nullptr; double *dm = nullptr;
    CK(cudaMalloc(&dp, npts * 16));
    if (out_pix) CK(cudaMalloc(&dq, npts * 16));
    if (out_modifier) CK(cudaMalloc(&dm, npts * 8));
    CK(cudaMemcpy(dp, c_pix, npts * 16, cudaMemcpyHostToDevice));
    k_proj_apply<<<(int)((npts + 127) / 128), 128>>>(P, npts, dp, dq, dm);
    CK(cudaGetLastError());
    if (out_pix) CK(cudaMemcpy(out_pix, dq, npts * 16, cudaMemcpyDeviceToHost));
    if (out_modifier) CK(cudaMemcpy(out_modifier, dm, npts * 8, cudaMemcpyDeviceToHost));
    cudaFree(dp); cudaFree(dq); cudaFree(dm);
    return 0;
}

int fsb_hypot_test(int64_t n, const double *x, const double *y, double *out)
{
    if (ensure_init() != 0) return -1;
    if (n <= 0) return 0;
    double *dx, *dy, *dout;
    CK(cudaMalloc(&dx, n * 8)); CK(cudaMalloc(&dy, n * 8)); CK(cudaMalloc(&dout, n * 8));
    CK(cudaMemcpy(dx, x, n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dy, y, n * 8, cudaMemcpyHostToDevice));
    k_hypot<<<(int)((n + 127) / 128), 128>>>(n, dx, dy, dout);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, dout, n * 8, cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    return 0;
}

double fsb_fp64_peak_tflops(int iters)
{
    if (ensure_init() != 0) return -1.;
    double *dout = nullptr;
    if (cudaMalloc(&dout, 8) != cudaSuccess) return -1.;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int block = 256, grid = g_sm_count * 8;
    k_fp64_peak<<<grid, block>>>(1000, dout); /* warm-up */
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0, 0);
        k_fp64_peak<<<grid, block>>>(iters, dout);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dout);
    if (cudaGetLastError() != cudaSuccess) return -1.;
    double flops = (double)grid * block * 8.0 * (double)iters * 2.0;
    return flops / (best * 1e-3) / 1e12;
}

} /* extern "C" */
