// The extern "C" surface declared in include/b200_bls.h.
// No CPU compute path exists in this library: every entry point either launches the CUDA
// kernels of msm.cuh or returns an error.
#include <thread>

#include "engine.cuh"

namespace b200 {

void engine_teardown(Engine &E);

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// The engine is shared by every host thread of the process (cgo callers arrive on arbitrary threads): entry points copy
// the shared_ptr under g_engine_mu, then serialise on Engine::mu; b200_shutdown unpublishes it first, waits for the
// call in flight by taking Engine::mu, tears the CUDA objects down and marks the object dead for late holders.
static std::shared_ptr<Engine> g_engine;
static std::vector<std::shared_ptr<Engine>> g_group;      // b200_init_devices: one engine per GPU, g_group[0] == g_engine
static std::mutex g_engine_mu;
static std::vector<std::shared_ptr<Engine>> engine_group() {
    std::lock_guard<std::mutex> lk(g_engine_mu);
    return g_group;
}
static std::shared_ptr<Engine> engine_ref() {
    std::lock_guard<std::mutex> lk(g_engine_mu);
    return g_engine;
}

struct CurveInfo {
    size_t coord_bytes;   // one coordinate
    size_t scalar_bytes;
    size_t jac_bytes;
};
static bool curve_info(int curve, CurveInfo &ci) {
    switch (curve) {
    case B200_BLS12_377_G1: ci = {48, 32, 144}; return true;
    case B200_BLS12_377_G2: ci = {96, 32, 288}; return true;
    case B200_BW6_761_G1:
    case B200_BW6_761_G2: ci = {96, 48, 288}; return true;
    default: return false;
    }
}

#define DISPATCH_CURVE(curve, FN, ...)                                                     \
    ((curve) == B200_BLS12_377_G1   ? FN<G1_377>(__VA_ARGS__)                              \
     : (curve) == B200_BLS12_377_G2 ? FN<G2_377>(__VA_ARGS__)                              \
                                    : FN<G_761>(__VA_ARGS__))

}  // namespace b200

using namespace b200;

extern "C" {

const char *b200_last_error(void) { return g_err.c_str(); }
uint64_t b200_launch_count(void) { return g_launches.load(); }

// creates the engine of one device (streams, events); the caller publishes it
static int create_engine(int device, std::shared_ptr<Engine> &out) {
    CUDA_TRY(cudaSetDevice(device));
    // a failure below leaves through CUDA_TRY: the deleter releases whatever was created so far
    std::shared_ptr<Engine> E(new Engine(), [](Engine *e) {
        engine_teardown(*e);
        delete e;
    });
    E->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(B200_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    E->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&E->done, cudaEventDisableTiming));
    {
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        // sort and tail blocks are few and short: high priority lets them slip in between the long-lived
        // accumulate blocks instead of waiting for that kernel to run out of blocks (measured: 8.38 ->
        // 7.58 ms per MSM at n = 2^20).  B200_PIPE_PRIO: bit 0 = sort stream high, bit 1 = tail stream high.
        const int mask = getenv("B200_PIPE_PRIO") ? atoi(getenv("B200_PIPE_PRIO")) : 3;
        for (int i = 0; i < 3; i++) {
            const bool hi = (i == 0 && (mask & 1)) || (i == 2 && (mask & 2));
            CUDA_TRY(cudaStreamCreateWithPriority(&E->pipe_stream[i], cudaStreamNonBlocking, hi ? prio_hi : prio_lo));
        }
        std::vector<cudaEvent_t *> evs = {&E->ev_fork, &E->ev_join};
        for (int i = 0; i < Engine::PIPE_SETS_MAX; i++) {
            evs.push_back(&E->ev_sorted[i]);
            evs.push_back(&E->ev_acc[i]);
            evs.push_back(&E->ev_tail[i]);
        }
        for (cudaEvent_t *ev : evs)
            CUDA_TRY(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
        CUDA_TRY(cudaStreamCreateWithFlags(&E->copy_stream, cudaStreamNonBlocking));
        for (cudaEvent_t &ev : E->ev_chunk) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (cudaEvent_t &ev : E->ev_slot) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (cudaEvent_t &ev : E->ev_split) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CUDA_TRY(cudaStreamCreateWithFlags(&E->split_stream, cudaStreamNonBlocking));
    }
    out = E;
    return B200_OK;
}

int b200_init(int device) {
    std::lock_guard<std::mutex> lk(g_engine_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(B200_ERR_CUDA, "no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    if (device >= count) return fail(B200_ERR_ARG, "device %d out of range (%d devices)", device, count);
    if (g_engine && g_engine->device == device) return B200_OK;
    if (g_engine) return fail(B200_ERR_STATE, "engine already bound to device %d; call b200_shutdown first", g_engine->device);
    std::shared_ptr<Engine> E;
    int rc = create_engine(device, E);
    if (rc) return rc;
    g_engine = E;
    g_group.assign(1, E);
    return B200_OK;
}

// One process, several GPUs: one engine per listed device.  devices[0] becomes the primary engine -- every single-device
// entry point keeps using it -- and the *_sharded entry points spread one call over all of them.
int b200_init_devices(const int *devices, int count) {
    if (!devices || count <= 0) return fail(B200_ERR_ARG, "no devices");
    std::lock_guard<std::mutex> lk(g_engine_mu);
    int have = 0;
    cudaError_t e = cudaGetDeviceCount(&have);
    if (e != cudaSuccess || have == 0)
        return fail(B200_ERR_CUDA, "no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    for (int i = 0; i < count; i++) {
        if (devices[i] < 0 || devices[i] >= have) return fail(B200_ERR_ARG, "device %d out of range (%d devices)", devices[i], have);
        // a device listed twice gets two engines (tests drive the sharded path on a one-GPU box this way)
        static const bool allow_dup = getenv("B200_ALLOW_DUP_DEVICES") != nullptr;
        for (int j = 0; j < i && !allow_dup; j++)
            if (devices[j] == devices[i]) return fail(B200_ERR_ARG, "device %d listed twice", devices[i]);
    }
    if (g_engine && g_engine->device != devices[0])
        return fail(B200_ERR_STATE, "engine already bound to device %d; call b200_shutdown first", g_engine->device);
    std::vector<std::shared_ptr<Engine>> group;
    for (int i = 0; i < count; i++) {
        std::shared_ptr<Engine> E;
        for (const std::shared_ptr<Engine> &old : g_group) {
            bool taken = false;
            for (const std::shared_ptr<Engine> &g : group) taken = taken || g == old;
            if (old->device == devices[i] && !taken && !E) E = old;
        }
        if (!E) {
            int rc = create_engine(devices[i], E);
            if (rc) return rc;
        }
        group.push_back(E);
    }
    // partial results travel to the primary device by peer copy: direct over NVLink when peer access can be enabled
    for (int i = 1; i < count; i++) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) == cudaSuccess && can) {
            cudaSetDevice(devices[i]);
            cudaError_t pe = cudaDeviceEnablePeerAccess(devices[0], 0);
            if (pe != cudaSuccess) cudaGetLastError();       // already enabled, or not possible: the copy is staged instead
        }
    }
    cudaSetDevice(devices[0]);
    g_group = group;
    g_engine = group[0];
    return B200_OK;
}

int b200_device_count(void) {
    std::lock_guard<std::mutex> lk(g_engine_mu);
    return (int)g_group.size();
}

int b200_ensure_init(void) { return engine_ref() ? B200_OK : b200_init(-1); }

int b200_bound_device(void) {
    std::shared_ptr<Engine> e = engine_ref();
    return e ? e->device : -1;
}

void b200_shutdown(void) {
    std::vector<std::shared_ptr<Engine>> all;
    {
        std::lock_guard<std::mutex> lk(g_engine_mu);
        all.swap(g_group);
        if (g_engine && all.empty()) all.push_back(g_engine);
        g_engine.reset();
    }
    for (std::shared_ptr<Engine> &e : all) {
        std::lock_guard<std::mutex> lk(e->mu);         // waits for the call in flight
        engine_teardown(*e);
    }
}

}  // extern "C"

namespace b200 {
// idempotent: runs from b200_shutdown and again (as a no-op) from the shared_ptr deleter
void engine_teardown(Engine &En) {
    Engine *E = &En;
    if (E->dead) return;
    E->dead = true;
    cudaSetDevice(E->device);
    if (E->stream) cudaStreamSynchronize(E->stream);
    for (MsmWs &w : E->ws)
        for (b200::Buffer *b : {&w.counts, &w.offsets, &w.cursor, &w.tile_sums, &w.bins, &w.order, &w.sorted, &w.buckets, &w.partials,
                          &w.window_sums, &w.ones, &w.huge_slices, &w.aff_a, &w.aff_b})
            b->release();
    for (b200::Buffer *b : {&E->h2d_bases, &E->native_bases, &E->scalars, &E->result,
                      &E->miller, &E->g2_packed, &E->h2d_g2, &E->gather, &E->v_sum, &E->v_pairs1, &E->v_pairs2, &E->v_offsets, &E->v_flags, &E->v_g1jac, &E->v_g2jac, &E->v_g1aff, &E->v_g2aff, &E->g16_h, &E->g16_tmp, &E->g16_part, &E->census, &E->bh_table, &E->hash_ws, &E->sqrt_tables})
        b->release();
    for (NttDomain &d : E->ntt)
        for (b200::Buffer *b : {&d.consts, &d.pw, &d.tw}) b->release();
    for (auto &ev : E->prof_ev)
        if (ev) cudaEventDestroy(ev);
    if (E->done) cudaEventDestroy(E->done);
    std::vector<cudaEvent_t> evs = {E->ev_fork, E->ev_join};
    for (int i = 0; i < Engine::PIPE_SETS_MAX; i++) {
        evs.push_back(E->ev_sorted[i]);
        evs.push_back(E->ev_acc[i]);
        evs.push_back(E->ev_tail[i]);
    }
    for (cudaEvent_t ev : evs)
        if (ev) cudaEventDestroy(ev);
    for (cudaStream_t s : E->pipe_stream)
        if (s) cudaStreamDestroy(s);
    for (cudaEvent_t ev : E->ev_chunk)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : E->ev_slot)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : E->ev_split)
        if (ev) cudaEventDestroy(ev);
    if (E->split_stream) cudaStreamDestroy(E->split_stream);
    for (void *&p : E->pinned) {
        if (p) cudaFreeHost(p);
        p = nullptr;
    }
    if (E->copy_stream) cudaStreamDestroy(E->copy_stream);
    if (E->stream) cudaStreamDestroy(E->stream);
}
}  // namespace b200

extern "C" {

#define REQUIRE_ENGINE()                                                              \
    std::shared_ptr<Engine> Ep = engine_ref();                                        \
    if (!Ep) return fail(B200_ERR_STATE, "b200_init has not been called");            \
    Engine &E = *Ep;                                                                  \
    std::lock_guard<std::mutex> lk(E.mu);                                             \
    if (E.dead) return fail(B200_ERR_STATE, "engine was shut down");                  \
    CUDA_TRY(cudaSetDevice(E.device));                                                \
    ENGINE_ORDER(E.stream)

int b200_msm_device(int curve, const void *d_bases_packed, const void *d_scalars, size_t n, void *d_out_jacobian,
                    void *stream) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!d_out_jacobian || (n && (!d_bases_packed || !d_scalars))) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    return DISPATCH_CURVE(curve, msm_device, E, d_bases_packed, 2 * ci.coord_bytes, d_scalars, n, d_out_jacobian, st);
}

int b200_msm_prepared_device(int curve, const void *d_bases_prepared, const void *d_scalars, size_t n,
                             void *d_out_jacobian, void *stream) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!d_out_jacobian || (n && (!d_bases_prepared || !d_scalars))) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);
    return DISPATCH_CURVE(curve, msm_native, E, d_bases_prepared, d_scalars, n, d_out_jacobian, st);
}

int b200_msm_batch_device(int curve, const b200_msm_job *jobs, size_t count, void *stream) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (count && !jobs) return fail(B200_ERR_ARG, "null pointer");
    if (count == 0) return B200_OK;
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    return DISPATCH_CURVE(curve, msm_batch, E, jobs, count, st, nullptr);
}

int b200_pack_bases_device(int curve, const void *src, size_t stride, size_t n, int src_on_device, void *d_dst_packed,
                           void *stream) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (n && (!src || !d_dst_packed)) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    const void *dsrc = src;
    if (!src_on_device && n) {
        ENGINE_ORDER(st);                             // the staging buffer may still feed a prior call
        int rc = E.h2d_bases.reserve(n * stride);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(E.h2d_bases.p, src, n * stride, cudaMemcpyHostToDevice, st));
        dsrc = E.h2d_bases.p;
    }
    int rc = DISPATCH_CURVE(curve, pack_bases, dsrc, stride, n, d_dst_packed, st);
    if (rc) return rc;
    if (!src_on_device && n) ENGINE_MARK(st);
    return B200_OK;
}

// ---- host-pointer MSM ------------------------------------------------------------------------------------
// Large inputs are cut into chunks (about 2^18 pairs): chunk k + 1 crosses PCIe on the copy stream while chunk k
// is sorted and bucketed; all chunks feed ONE bucket set and one tail (msm_chunks_* in curve_impl.cuh).
// Page-locked callers (cudaHostAlloc / cudaHostRegister) are copied from directly.  Ordinary (pageable) memory --
// what a Rust Vec is -- goes through two pinned staging slots filled by a few host threads, so the PCIe copy
// runs at its pinned rate and overlaps the host-side memcpy of the next chunk.
constexpr size_t HOST_CHUNK_MIN = (size_t)1 << 17;       // below this: one plain copy + one MSM
constexpr size_t HOST_CHUNK_PAIRS = (size_t)1 << 18;
constexpr int STAGE_THREADS = 6;

static bool host_is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// result (one GroupProjective image) is left in E.result[0]; asynchronous on E.stream, except that pageable inputs
// have been fully read when the call returns.  Caller holds E.mu and has selected E.device.
static int msm_host(Engine &E, int curve, const void *bases, size_t stride, const uint64_t *scalars, size_t n) {
    CurveInfo ci;
    curve_info(curve, ci);
    cudaStream_t st = E.stream;
    int rc;
    const size_t packed = 2 * ci.coord_bytes;
    if ((rc = E.result.reserve(4 * ci.jac_bytes))) return rc;
    char *res = E.result.as<char>();
    if (n) {
        if (stride % 4 || stride < packed) return fail(B200_ERR_ARG, "bad base stride %zu", stride);
        if ((rc = E.scalars.reserve(n * ci.scalar_bytes)) || (rc = E.h2d_bases.reserve(n * stride))) return rc;
        if (stride != packed && (rc = E.native_bases.reserve(n * packed))) return rc;
    }
    static const bool no_chunks = getenv("B200_HOST_NOCHUNK") != nullptr;
    if (n < HOST_CHUNK_MIN || no_chunks) {
        const void *d_bases = nullptr;
        if (n) {
            CUDA_TRY(cudaMemcpyAsync(E.scalars.p, scalars, n * ci.scalar_bytes, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(E.h2d_bases.p, bases, n * stride, cudaMemcpyHostToDevice, st));
            d_bases = E.h2d_bases.p;
        }
        return DISPATCH_CURVE(curve, msm_device, E, d_bases, stride, E.scalars.p, n, res, st);
    }
    // chunk bounds: about 2^18 pairs each, but the first two are a quarter and a half of that -- the accumulation can only
    // start once the first chunk has crossed PCIe and been sorted, so a short first chunk trims the exposed latency
    size_t bound[Engine::MAX_CHUNKS + 1];
    int chunks = 0;
    {
        const size_t body = std::max<size_t>(HOST_CHUNK_PAIRS, (n + Engine::MAX_CHUNKS - 3) / (Engine::MAX_CHUNKS - 2));
        size_t at = 0;
        bound[0] = 0;
        for (size_t want : {HOST_CHUNK_PAIRS / 4, HOST_CHUNK_PAIRS / 2}) {
            if (n - at > want + HOST_CHUNK_PAIRS / 2) bound[++chunks] = (at += want);
        }
        while (n - at > body + body / 4 && chunks < Engine::MAX_CHUNKS - 1) bound[++chunks] = (at += body);
        bound[++chunks] = n;
    }
    size_t chunk_max = 0;
    for (int c = 0; c < chunks; c++) chunk_max = std::max(chunk_max, bound[c + 1] - bound[c]);
    chunk_max += 1;
    const size_t rec = stride + ci.scalar_bytes;
    const bool pinned = host_is_pinned(bases) && host_is_pinned(scalars);
    if (!pinned && E.pinned_cap < chunk_max * rec) {          // staging slots: grow-only
        for (void *&p : E.pinned) {
            if (p) cudaFreeHost(p);
            p = nullptr;
        }
        E.pinned_cap = 0;
        for (void *&p : E.pinned) CUDA_TRY(cudaHostAlloc(&p, chunk_max * rec + chunk_max * rec / 8, cudaHostAllocDefault));
        E.pinned_cap = chunk_max * rec + chunk_max * rec / 8;
    }
    cudaStream_t cs = E.copy_stream;
    CUDA_TRY(cudaEventRecord(E.ev_fork, st));                 // the copy stream starts after everything queued on st
    CUDA_TRY(cudaStreamWaitEvent(cs, E.ev_fork, 0));
    if ((rc = DISPATCH_CURVE(curve, msm_chunks_begin, E, n, chunk_max, st))) return rc;
    const char *hb = reinterpret_cast<const char *>(bases), *hs = reinterpret_cast<const char *>(scalars);
    char *db = E.h2d_bases.as<char>(), *ds = E.scalars.as<char>(), *dn = E.native_bases.as<char>();

    // pageable input: STAGE_THREADS host threads copy chunk k into slot k % 2 as soon as the slot's previous H2D is done
    std::atomic<int> staged[Engine::MAX_CHUNKS];
    std::atomic<int> free_upto{2};                            // chunks below this index may be written into their slot
    std::vector<std::thread> workers;
    if (!pinned) {
        for (int c = 0; c < chunks; c++) staged[c].store(0);
        for (int t = 0; t < STAGE_THREADS; t++)
            workers.emplace_back([&, t]() {
                for (int c = 0; c < chunks; c++) {
                    while (free_upto.load(std::memory_order_acquire) <= c) std::this_thread::yield();
                    const size_t lo = bound[c], cnt = bound[c + 1] - lo;
                    char *slot = reinterpret_cast<char *>(E.pinned[c & 1]);
                    const size_t bbytes = cnt * stride, sbytes = cnt * ci.scalar_bytes, total = bbytes + sbytes;
                    const size_t a = total * t / STAGE_THREADS, b = total * (t + 1) / STAGE_THREADS;   // this thread's byte range
                    if (a < bbytes) memcpy(slot + a, hb + lo * stride + a, std::min(b, bbytes) - a);
                    if (b > bbytes) {
                        const size_t sa = std::max(a, bbytes) - bbytes;
                        memcpy(slot + bbytes + sa, hs + lo * ci.scalar_bytes + sa, b - bbytes - sa);
                    }
                    staged[c].fetch_add(1, std::memory_order_release);
                }
            });
    }
    auto join = [&]() {
        free_upto.store(1 << 30, std::memory_order_release);
        for (std::thread &w : workers) w.join();
        workers.clear();
    };
    rc = B200_OK;
    for (int c = 0; c < chunks && rc == B200_OK; c++) {
        const size_t lo = bound[c], cnt = bound[c + 1] - lo;
        const char *src_b = hb + lo * stride, *src_s = hs + lo * ci.scalar_bytes;
        if (!pinned) {
            while (staged[c].load(std::memory_order_acquire) < STAGE_THREADS) std::this_thread::yield();
            src_b = reinterpret_cast<const char *>(E.pinned[c & 1]);
            src_s = src_b + cnt * stride;
        }
        cudaError_t e = cudaMemcpyAsync(ds + lo * ci.scalar_bytes, src_s, cnt * ci.scalar_bytes, cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(db + lo * stride, src_b, cnt * stride, cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess && !pinned) e = cudaEventRecord(E.ev_slot[c & 1], cs);
        if (e != cudaSuccess) {
            rc = fail(B200_ERR_CUDA, "host-to-device copy: %s", cudaGetErrorString(e));
            break;
        }
        const void *chunk_bases = db + lo * stride;
        if (stride != packed) {
            if ((rc = DISPATCH_CURVE(curve, pack_bases, db + lo * stride, stride, cnt, dn + lo * packed, cs))) break;
            chunk_bases = dn + lo * packed;
        }
        if ((e = cudaEventRecord(E.ev_chunk[c], cs)) != cudaSuccess) {
            rc = fail(B200_ERR_CUDA, "event record: %s", cudaGetErrorString(e));
            break;
        }
        if ((rc = DISPATCH_CURVE(curve, msm_chunks_add, E, chunk_bases, ds + lo * ci.scalar_bytes, cnt, E.ev_chunk[c], c == chunks - 1, res)))
            break;
        if (!pinned && c + 2 < chunks) {                      // slot c % 2 is reusable once its copy has left the host
            if ((e = cudaEventSynchronize(E.ev_slot[c & 1])) != cudaSuccess) {
                rc = fail(B200_ERR_CUDA, "event sync: %s", cudaGetErrorString(e));
                break;
            }
            free_upto.store(c + 3, std::memory_order_release);
        }
    }
    join();
    if (rc) return rc;
    if (!pinned) CUDA_TRY(cudaStreamSynchronize(cs));          // the caller's buffers and the slots are free again
    return DISPATCH_CURVE(curve, msm_chunks_finish, E, res, st);
}

int b200_msm(int curve, const void *bases, size_t stride, const uint64_t *scalars, size_t n, void *out_jacobian) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!out_jacobian || (n && (!bases || !scalars))) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    int rc = msm_host(E, curve, bases, stride, scalars, n);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_jacobian, E.result.p, ci.jac_bytes, cudaMemcpyDeviceToHost, E.stream));
    CUDA_TRY(cudaStreamSynchronize(E.stream));
    return B200_OK;
}

// ---- one process, several GPUs ------------------------------------------------------------------------------
// The input is cut into one contiguous slice per engine; a host thread per GPU runs that slice through the host-pointer
// MSM above, every GPU peer-copies its partial point (144 / 288 bytes) to the primary GPU, which adds them up.
}  // extern "C"

template <class Fn>
static int run_on_group(const std::vector<std::shared_ptr<Engine>> &group, Fn fn) {
    std::vector<int> rcs(group.size(), B200_OK);
    std::vector<std::string> errs(group.size());
    std::vector<std::thread> th;
    for (size_t d = 0; d < group.size(); d++)
        th.emplace_back([&, d]() {
            Engine &E = *group[d];
            std::lock_guard<std::mutex> lk(E.mu);
            if (E.dead) {
                rcs[d] = fail(B200_ERR_STATE, "engine was shut down");
            } else if (cudaSetDevice(E.device) != cudaSuccess) {
                rcs[d] = fail(B200_ERR_CUDA, "cudaSetDevice(%d)", E.device);
            } else {
                rcs[d] = fn(E, d);
            }
            if (rcs[d]) errs[d] = g_err;                       // the error text is thread-local
        });
    for (std::thread &t : th) t.join();
    for (size_t d = 0; d < group.size(); d++)
        if (rcs[d]) return fail(rcs[d], "GPU %d: %s", group[d]->device, errs[d].c_str());
    return B200_OK;
}

extern "C" {

static int sharded_gather(Engine &E, Engine &E0, size_t d, size_t bytes, const void *d_src) {
    // E0.gather was sized before the worker threads started
    CUDA_TRY(cudaMemcpyPeerAsync(E0.gather.as<char>() + d * bytes, E0.device, d_src, E.device, bytes, E.stream));
    CUDA_TRY(cudaStreamSynchronize(E.stream));
    return B200_OK;
}

int b200_msm_sharded(int curve, const void *bases, size_t stride, const uint64_t *scalars, size_t n, void *out_jacobian) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!out_jacobian || (n && (!bases || !scalars))) return fail(B200_ERR_ARG, "null pointer");
    std::vector<std::shared_ptr<Engine>> group = engine_group();
    if (group.empty()) return fail(B200_ERR_STATE, "b200_init / b200_init_devices has not been called");
    if (group.size() == 1 || n < group.size() * 1024) return b200_msm(curve, bases, stride, scalars, n, out_jacobian);
    const size_t G = group.size();
    Engine &E0 = *group[0];
    {
        std::lock_guard<std::mutex> lk(E0.mu);
        CUDA_TRY(cudaSetDevice(E0.device));
        int rc = E0.gather.reserve((G + 1) * ci.jac_bytes);
        if (rc) return rc;
    }
    const char *hb = reinterpret_cast<const char *>(bases);
    int rc = run_on_group(group, [&](Engine &E, size_t d) -> int {
        const size_t lo = n * d / G, cnt = n * (d + 1) / G - lo;
        ENGINE_ORDER(E.stream);
        int r = msm_host(E, curve, hb + lo * stride, stride, scalars + lo * (ci.scalar_bytes / 8), cnt);
        if (r) return r;
        return sharded_gather(E, E0, d, ci.jac_bytes, E.result.p);
    });
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(E0.mu);
    CUDA_TRY(cudaSetDevice(E0.device));
    char *g = E0.gather.as<char>();
    if ((rc = DISPATCH_CURVE(curve, sum_jacobian, g, G, g + G * ci.jac_bytes, E0.stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_jacobian, g + G * ci.jac_bytes, ci.jac_bytes, cudaMemcpyDeviceToHost, E0.stream));
    CUDA_TRY(cudaStreamSynchronize(E0.stream));
    return B200_OK;
}

int b200_msm_bls12_377_g1(const void *b, const uint64_t *s, size_t n, void *o) { return b200_msm(B200_BLS12_377_G1, b, 104, s, n, o); }
int b200_msm_bls12_377_g2(const void *b, const uint64_t *s, size_t n, void *o) { return b200_msm(B200_BLS12_377_G2, b, 200, s, n, o); }
int b200_msm_bw6_761_g1(const void *b, const uint64_t *s, size_t n, void *o) { return b200_msm(B200_BW6_761_G1, b, 200, s, n, o); }
int b200_msm_bw6_761_g2(const void *b, const uint64_t *s, size_t n, void *o) { return b200_msm(B200_BW6_761_G2, b, 200, s, n, o); }

int b200_sum_jacobian_device(int curve, const void *d_points, size_t count, void *d_out, void *stream) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!d_out || (count && !d_points)) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    return DISPATCH_CURVE(curve, sum_jacobian, d_points, count, d_out, st);
}

int b200_sum_jacobian_batch_device(int curve, const void *d_points, size_t count, size_t batch, void *d_out, void *stream) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (batch && (!d_out || (count && !d_points))) return fail(B200_ERR_ARG, "null pointer");
    if (batch > 65535) return fail(B200_ERR_ARG, "batch too large");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    return DISPATCH_CURVE(curve, sum_jacobian_batch, d_points, count, batch, d_out, st);
}

// host-pointer form: `count` contiguous arkworks GroupProjective images in host memory -> their sum (host memory).
// Everything is queued on the engine's stream (copies included), so the kernel can never see a stale staging buffer.
int b200_sum_jacobian(int curve, const void *points, size_t count, void *out) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!out || (count && !points)) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    cudaStream_t st = E.stream;
    int rc = E.v_sum.reserve((count + 1) * ci.jac_bytes);
    if (rc) return rc;
    char *d = E.v_sum.as<char>();
    if (count) CUDA_TRY(cudaMemcpyAsync(d, points, count * ci.jac_bytes, cudaMemcpyHostToDevice, st));
    if ((rc = DISPATCH_CURVE(curve, sum_jacobian, d, count, d + count * ci.jac_bytes, st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, d + count * ci.jac_bytes, ci.jac_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B200_OK;
}

// out = scalar * base for ONE point, host pointers, GroupProjective images in and out (out is (x, y, 1) or zero()): the
// scalar multiplications of PrivateKey::sign_raw / to_public (crates/bls-crypto/src/bls/secret.rs:65-72)
int b200_scalar_mul(int curve, const void *base_jacobian, const uint64_t *scalar, void *out_jacobian) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (!base_jacobian || !scalar || !out_jacobian) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    cudaStream_t st = E.stream;
    const size_t packed = 2 * ci.coord_bytes;
    int rc = E.v_sum.reserve(ci.jac_bytes + 2 * packed + ci.scalar_bytes + 64);
    if (rc) return rc;
    char *d_j = E.v_sum.as<char>(), *d_a = d_j + ci.jac_bytes, *d_o = d_a + packed, *d_s = d_o + packed;
    CUDA_TRY(cudaMemcpyAsync(d_j, base_jacobian, ci.jac_bytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_s, scalar, ci.scalar_bytes, cudaMemcpyHostToDevice, st));
    if ((rc = DISPATCH_CURVE(curve, batch_to_affine, d_j, (size_t)1, d_a, st))) return rc;
    if ((rc = DISPATCH_CURVE(curve, fixed_base_mul, E, d_a, d_s, (size_t)1, d_o, st))) return rc;
    std::vector<unsigned char> aff(packed);
    CUDA_TRY(cudaMemcpyAsync(aff.data(), d_o, packed, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    // Montgomery form of 1 in the coordinate field (c0 of an Fq2 coordinate; c1 = 0)
    unsigned char one[96] = {0};
    if (ci.coord_bytes == 96 && curve != B200_BLS12_377_G2)
        for (int i = 0; i < 24; i++) reinterpret_cast<uint32_t *>(one)[i] = Fq761Params::one(i);
    else
        for (int i = 0; i < 12; i++) reinterpret_cast<uint32_t *>(one)[i] = Fq377Params::one(i);
    unsigned char *o = reinterpret_cast<unsigned char *>(out_jacobian);
    bool inf = true;
    for (size_t i = 0; i < packed; i++) inf = inf && aff[i] == 0;
    memset(o, 0, ci.jac_bytes);
    if (inf) {                                        // GroupProjective::zero() = (1, 1, 0)
        memcpy(o, one, ci.coord_bytes);
        memcpy(o + ci.coord_bytes, one, ci.coord_bytes);
    } else {
        memcpy(o, aff.data(), packed);
        memcpy(o + packed, one, ci.coord_bytes);
    }
    return B200_OK;
}

int b200_fixed_base_mul_device(int curve, const void *d_base_packed, const void *d_scalars, size_t n,
                               void *d_out_packed, void *stream) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (n && (!d_base_packed || !d_scalars || !d_out_packed)) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);
    int rc = DISPATCH_CURVE(curve, fixed_base_mul, E, d_base_packed, d_scalars, n, d_out_packed, st);
    if (rc) return rc;
    ENGINE_MARK(st);
    return B200_OK;
}

int b200_point_runs_device(int curve, const void *d_base_packed, const void *d_start_scalars, size_t runs, size_t run_len,
                           void *d_out_packed, void *stream) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (runs && run_len && (!d_base_packed || !d_start_scalars || !d_out_packed)) return fail(B200_ERR_ARG, "null pointer");
    if (runs * run_len > ((size_t)1 << 28)) return fail(B200_ERR_ARG, "too many points");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);
    int rc = DISPATCH_CURVE(curve, point_runs, E, d_base_packed, d_start_scalars, runs, run_len, d_out_packed, st);
    if (rc) return rc;
    ENGINE_MARK(st);
    return B200_OK;
}

int b200_batch_to_affine_device(int curve, const void *d_jacobian, size_t n, void *d_out_packed, void *stream) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (n && (!d_jacobian || !d_out_packed)) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    return DISPATCH_CURVE(curve, batch_to_affine, d_jacobian, n, d_out_packed, st);
}

int b200_miller_product_bls12_377_device(const void *d_g1_packed, const void *d_g2_packed, size_t n, void *d_out_fq12,
                                         void *stream) {
    if (!d_out_fq12 || (n && (!d_g1_packed || !d_g2_packed))) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);
    int rc = miller_product(E, d_g1_packed, d_g2_packed, n, d_out_fq12, st);
    if (rc) return rc;
    ENGINE_MARK(st);
    return B200_OK;
}

int b200_final_exp_bls12_377_device(const void *d_fq12_vals, size_t count, void *d_out_fq12, int *d_is_one,
                                    void *stream) {
    if (!d_fq12_vals || count == 0) return fail(B200_ERR_ARG, "null pointer / empty input");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);
    int rc = final_exp(E, d_fq12_vals, count, d_out_fq12, d_is_one, st);
    if (rc) return rc;
    ENGINE_MARK(st);
    return B200_OK;
}

// Miller product of n host pairs -> E.result[0 .. 576); asynchronous on E.stream (caller holds E.mu)
static int pairing_host_miller(Engine &E, const void *g1, size_t stride1, const void *g2, size_t stride2, size_t n) {
    cudaStream_t st = E.stream;
    int rc;
    if ((rc = E.result.reserve(576 + 16))) return rc;
    if (n) {
        if (stride1 % 4 || stride1 < 96 || stride2 % 4 || stride2 < 192) return fail(B200_ERR_ARG, "bad stride");
        if ((rc = E.h2d_bases.reserve(n * stride1)) || (rc = E.h2d_g2.reserve(n * stride2)) ||
            (rc = E.native_bases.reserve(n * 96)) || (rc = E.g2_packed.reserve(n * 192)))
            return rc;
        CUDA_TRY(cudaMemcpyAsync(E.h2d_bases.p, g1, n * stride1, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(E.h2d_g2.p, g2, n * stride2, cudaMemcpyHostToDevice, st));
        if ((rc = pack_bases<G1_377>(E.h2d_bases.p, stride1, n, E.native_bases.p, st))) return rc;
        if ((rc = pack_bases<G2_377>(E.h2d_g2.p, stride2, n, E.g2_packed.p, st))) return rc;
    }
    return miller_product(E, E.native_bases.p, E.g2_packed.p, n, E.result.p, st);
}

// final exponentiation of `count` Miller values at d_vals (engine's device) -> host outputs
static int pairing_host_finish(Engine &E, const void *d_vals, size_t count, void *out_fq12, int *out_is_one) {
    cudaStream_t st = E.stream;
    int rc;
    if ((rc = E.result.reserve(576 + 16))) return rc;
    char *res = E.result.as<char>();
    if ((rc = final_exp(E, d_vals, count, res, reinterpret_cast<int *>(res + 576), st))) return rc;
    if (out_fq12) CUDA_TRY(cudaMemcpyAsync(out_fq12, res, 576, cudaMemcpyDeviceToHost, st));
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, res + 576, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (out_is_one) *out_is_one = flag;
    return B200_OK;
}

int b200_multi_pairing_bls12_377(const void *g1, size_t stride1, const void *g2, size_t stride2, size_t n,
                                 void *out_fq12, int *out_is_one) {
    if (n && (!g1 || !g2)) return fail(B200_ERR_ARG, "null pointer");
    if (!out_fq12 && !out_is_one) return fail(B200_ERR_ARG, "no output requested");
    REQUIRE_ENGINE();
    int rc = pairing_host_miller(E, g1, stride1, g2, stride2, n);
    if (rc) return rc;
    return pairing_host_finish(E, E.result.p, 1, out_fq12, out_is_one);
}

// the pairs split over the GPUs of b200_init_devices: one Miller value (576 bytes) per GPU travels to the primary GPU,
// which multiplies them and runs the one final exponentiation -- the same GT element as the single-GPU call
int b200_multi_pairing_bls12_377_sharded(const void *g1, size_t stride1, const void *g2, size_t stride2, size_t n,
                                         void *out_fq12, int *out_is_one) {
    if (n && (!g1 || !g2)) return fail(B200_ERR_ARG, "null pointer");
    if (!out_fq12 && !out_is_one) return fail(B200_ERR_ARG, "no output requested");
    std::vector<std::shared_ptr<Engine>> group = engine_group();
    if (group.empty()) return fail(B200_ERR_STATE, "b200_init / b200_init_devices has not been called");
    if (group.size() == 1 || n < group.size() * 64) return b200_multi_pairing_bls12_377(g1, stride1, g2, stride2, n, out_fq12, out_is_one);
    const size_t G = group.size();
    Engine &E0 = *group[0];
    {
        std::lock_guard<std::mutex> lk(E0.mu);
        CUDA_TRY(cudaSetDevice(E0.device));
        int rc = E0.gather.reserve(G * 576);
        if (rc) return rc;
    }
    const char *h1 = reinterpret_cast<const char *>(g1), *h2 = reinterpret_cast<const char *>(g2);
    int rc = run_on_group(group, [&](Engine &E, size_t d) -> int {
        const size_t lo = n * d / G, cnt = n * (d + 1) / G - lo;
        ENGINE_ORDER(E.stream);
        int r = pairing_host_miller(E, h1 + lo * stride1, stride1, h2 + lo * stride2, stride2, cnt);
        if (r) return r;
        return sharded_gather(E, E0, d, 576, E.result.p);
    });
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(E0.mu);
    CUDA_TRY(cudaSetDevice(E0.device));
    return pairing_host_finish(E0, E0.gather.p, G, out_fq12, out_is_one);
}

int b200_multi_pairing_bw6_761(const void *g1, size_t stride1, const void *g2, size_t stride2, size_t n,
                               void *out_fq6, int *out_is_one) {
    if (n && (!g1 || !g2)) return fail(B200_ERR_ARG, "null pointer");
    if (!out_fq6 && !out_is_one) return fail(B200_ERR_ARG, "no output requested");
    REQUIRE_ENGINE();
    return bw6_multi_pairing_host(E, g1, stride1, g2, stride2, n, out_fq6, out_is_one);
}

int b200_miller_values_bw6_761_device(const void *d_g1_packed, const void *d_g2_packed, size_t n, void *d_out_vals,
                                      void *stream) {
    if (n && (!d_g1_packed || !d_g2_packed || !d_out_vals)) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);
    int rc = bw6_miller_values(E, d_g1_packed, d_g2_packed, n, d_out_vals, st);
    if (rc) return rc;
    ENGINE_MARK(st);
    return B200_OK;
}

int b200_final_exp_bw6_761_device(const void *d_vals, size_t count, void *d_out_fq6, int *d_is_one, void *stream) {
    if (!d_vals || count == 0) return fail(B200_ERR_ARG, "null pointer / empty input");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);
    int rc = bw6_final_exp(E, d_vals, count, d_out_fq6, d_is_one, st);
    if (rc) return rc;
    ENGINE_MARK(st);
    return B200_OK;
}

int b200_groth16_verify_bw6_761(const b200_groth16_vk *vk, const void *proof_a, const void *proof_b, const void *proof_c,
                                const uint64_t *public_inputs, size_t num_inputs, int *out_verified) {
    if (!vk || !proof_a || !proof_b || !proof_c || !out_verified || !vk->alpha_g1 || !vk->beta_g2 || !vk->gamma_g2 ||
        !vk->delta_g2 || !vk->gamma_abc_g1 || (num_inputs && !public_inputs))
        return fail(B200_ERR_ARG, "null pointer");
    if (num_inputs + 1 != vk->num_gamma_abc)
        return fail(B200_ERR_ARG, "malformed verifying key: %zu public inputs against %zu gamma_abc entries", num_inputs,
                    vk->num_gamma_abc);
    REQUIRE_ENGINE();
    return bw6_groth16_verify(E, vk, proof_a, proof_b, proof_c, public_inputs, num_inputs, out_verified);
}

int b200_deserialize_points(int kind, const void *bytes, size_t n, int check_subgroup, void *out_packed, int *out_status) {
    if (n && (!bytes || (!out_packed && !out_status))) return fail(B200_ERR_ARG, "null pointer");
    if (kind < 0 || kind > 3) return fail(B200_ERR_ARG, "unknown point kind %d", kind);
    REQUIRE_ENGINE();
    return decode_points_host(E, kind, bytes, n, check_subgroup, out_packed, out_status);
}

void b200_blake2s_personal(const uint8_t *data, size_t len, const uint8_t *personal8, uint8_t *out32) {
    blake2s_personal(data, len, personal8, out32);
}
void b200_blake2s_param(const uint8_t *data, size_t len, int digest_len, int fanout, int depth, uint32_t leaf_len, uint64_t node_offset,
                        int inner_len, const uint8_t *personal8, uint8_t *out) {
    blake2s_param(data, len, digest_len, fanout, depth, leaf_len, node_offset, inner_len, personal8, out);
}
int b200_encode_epoch_block(int cip22, uint16_t index, uint8_t round, const uint8_t *epoch_entropy, const uint8_t *parent_entropy,
                            uint32_t maximum_non_signers, size_t maximum_validators, const uint8_t *keys96, size_t nkeys,
                            uint8_t **out_inner, size_t *out_inner_len, uint8_t **out_extra, size_t *out_extra_len) {
    if (!out_inner || !out_inner_len || (nkeys && !keys96) || (cip22 && (!out_extra || !out_extra_len))) return fail(B200_ERR_ARG, "null pointer");
    std::vector<uint8_t> inner, extra;
    epoch_block_encode(cip22, index, round, epoch_entropy, parent_entropy, maximum_non_signers, maximum_validators, keys96, nkeys, &inner,
                       &extra);
    auto leak = [](const std::vector<uint8_t> &v) {
        uint8_t *p = (uint8_t *)malloc(v.size() ? v.size() : 1);
        if (p && !v.empty()) memcpy(p, v.data(), v.size());
        return p;
    };
    *out_inner = leak(inner);
    *out_inner_len = inner.size();
    if (!*out_inner) return fail(B200_ERR_ARG, "out of memory");
    if (cip22) {
        *out_extra = leak(extra);
        *out_extra_len = extra.size();
        if (!*out_extra) {
            free(*out_inner);
            return fail(B200_ERR_ARG, "out of memory");
        }
    }
    return B200_OK;
}

int b200_verify_epochs(const uint8_t *vk, size_t vk_len, const uint8_t *proof, size_t proof_len, const void *first_epoch,
                       const void *last_epoch, int *out_ok) {
    if (!out_ok || !first_epoch || !last_epoch) return fail(B200_ERR_ARG, "null pointer");
    *out_ok = 0;
    {                                                 // the reference's init() is optional: bind on first use
        int rc = b200_ensure_init();
        if (rc) return rc;
    }
    REQUIRE_ENGINE();
    std::string why;
    int rc = epoch_verify(E, vk, vk_len, proof, proof_len, *reinterpret_cast<const EpochBlockFFI *>(first_epoch),
                          *reinterpret_cast<const EpochBlockFFI *>(last_epoch), out_ok, &why);
    if (rc == B200_OK && !*out_ok) fail(B200_OK, "verify: %s", why.c_str());      // keep the reason readable
    return rc;
}

int b200_epoch_public_inputs(const void *first_epoch, const void *last_epoch, uint64_t *out_inputs, size_t capacity,
                             size_t *out_count, int *out_ok) {
    if (!out_ok || !out_count || !first_epoch || !last_epoch) return fail(B200_ERR_ARG, "null pointer");
    *out_ok = 0;
    *out_count = 0;
    {
        int rc = b200_ensure_init();
        if (rc) return rc;
    }
    REQUIRE_ENGINE();
    std::string why;
    std::vector<uint64_t> inputs;
    int rc = epoch_public_inputs(E, *reinterpret_cast<const EpochBlockFFI *>(first_epoch),
                                 *reinterpret_cast<const EpochBlockFFI *>(last_epoch), &inputs, out_ok, &why);
    if (rc) return rc;
    if (!*out_ok) {
        fail(B200_OK, "epoch blocks: %s", why.c_str());
        return B200_OK;
    }
    *out_count = inputs.size() / 6;
    if (*out_count > capacity || (inputs.size() && !out_inputs)) return fail(B200_ERR_ARG, "output holds %zu scalars, %zu needed", capacity, *out_count);
    memcpy(out_inputs, inputs.data(), inputs.size() * sizeof(uint64_t));
    return B200_OK;
}

// bls-snark-sys' own entry point (crates/bls-snark-sys/src/snark/mod.rs:23-45): bool, errors logged and mapped to false
bool verify(const uint8_t *vk, uint32_t vk_len, const uint8_t *proof, uint32_t proof_len, EpochBlockFFI first_epoch,
            EpochBlockFFI last_epoch) {
    int ok = 0;
    int rc = b200_verify_epochs(vk, vk_len, proof, proof_len, &first_epoch, &last_epoch, &ok);
    if (rc != B200_OK || !ok) fprintf(stderr, "[b200] verify -> false: %s\n", b200_last_error());
    return rc == B200_OK && ok;
}

int b200_batch_verify_hashes(const void *signature, const void *pubkeys, const void *message_hashes, size_t n,
                             int *out_verified) {
    if (!signature || !out_verified || (n && (!pubkeys || !message_hashes))) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    return batch_verify_hashes(E, signature, pubkeys, message_hashes, n, out_verified);
}

int b200_batch_verify_strict_hash(const void *pubkeys, const void *signatures, const uint64_t *exponents, size_t n,
                                  const void *message_hash, int *out_verified) {
    if (!message_hash || !out_verified || (n && (!pubkeys || !signatures || !exponents)))
        return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    return batch_verify_strict_hash(E, pubkeys, signatures, exponents, n, message_hash, out_verified);
}

int b200_batch_verify_strict_many(const b200_strict_batch *batches, size_t count, int *out_verified) {
    if (count && (!batches || !out_verified)) return fail(B200_ERR_ARG, "null pointer");
    if (count == 0) return B200_OK;
    REQUIRE_ENGINE();
    return batch_verify_strict_many(E, batches, count, out_verified);
}

int b200_serialize_points(int kind, const void *jacobian_images, size_t n, void *out_bytes) {
    if (n && (!jacobian_images || !out_bytes)) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    return encode_points_host(E, kind, jacobian_images, n, out_bytes);
}

int b200_hash_to_g1(int hasher, int flags, const uint8_t *domain, size_t domain_len, const b200_hash_input *inputs, size_t n,
                    void *out_jacobian, uint32_t *out_attempts) {
    if ((n && (!inputs || !out_jacobian)) || (domain_len && !domain)) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    return hash_to_g1(E, hasher, flags, domain, domain_len, inputs, n, out_jacobian, out_attempts);
}

int b200_ntt_device(int field, void *d_data, unsigned log_n, int inverse, int coset, void *stream) {
    if (field != B200_FR_BLS12_377 && field != B200_FR_BW6_761) return fail(B200_ERR_ARG, "unknown scalar field id %d", field);
    if (!d_data) return fail(B200_ERR_ARG, "null pointer");
    if (log_n > 26) return fail(B200_ERR_ARG, "log_n = %u exceeds the 2^26 per-call limit", log_n);
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);                                 // the cached domain tables may have been built on another stream
    int rc = ntt_transform(E, field, d_data, (int)log_n, inverse, coset, st);
    if (rc) return rc;
    ENGINE_MARK(st);
    return B200_OK;
}

int b200_witness_map_device(int field, void *d_a, void *d_b, void *d_c, unsigned log_n, void *d_h, void *stream) {
    if (field != B200_FR_BLS12_377 && field != B200_FR_BW6_761) return fail(B200_ERR_ARG, "unknown scalar field id %d", field);
    if (!d_a || !d_b || !d_c || !d_h) return fail(B200_ERR_ARG, "null pointer");
    if (log_n > 26) return fail(B200_ERR_ARG, "log_n = %u exceeds the 2^26 per-call limit", log_n);
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);
    int rc = witness_map(E, field, d_a, d_b, d_c, (int)log_n, d_h, st);
    if (rc) return rc;
    ENGINE_MARK(st);
    return B200_OK;
}

int b200_groth16_prove_device(int family, const b200_groth16_pk *pk, const void *d_assignment, size_t num_assign,
                              size_t num_aux, void *d_a, void *d_b, void *d_c, unsigned log_n, void *d_proof,
                              void *stream) {
    if (family != B200_GROTH16_BLS12_377 && family != B200_GROTH16_BW6_761) return fail(B200_ERR_ARG, "unknown Groth16 family %d", family);
    if (!pk || !pk->a_query || !pk->b_g2_query || !pk->h_query || !pk->alpha_g1 || !pk->beta_g2 || !d_a || !d_b || !d_c ||
        !d_proof || (num_assign && !d_assignment) || (num_aux && !pk->l_query))
        return fail(B200_ERR_ARG, "null pointer");
    if (num_aux > num_assign) return fail(B200_ERR_ARG, "num_aux exceeds num_assign");
    if (log_n == 0 || log_n > 26) return fail(B200_ERR_ARG, "log_n = %u out of range [1, 26]", log_n);
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);
    return groth16_prove(E, family, pk, d_assignment, num_assign, num_aux, d_a, d_b, d_c, log_n, d_proof, st);
}

int b200_groth16_prove_partial_device(int family, const b200_groth16_pk *pk, const void *d_assignment, size_t num_assign,
                                      size_t num_aux, void *d_a, void *d_b, void *d_c, unsigned log_n, unsigned shard,
                                      unsigned shards, void *d_partials, void *stream) {
    if (family != B200_GROTH16_BLS12_377 && family != B200_GROTH16_BW6_761) return fail(B200_ERR_ARG, "unknown Groth16 family %d", family);
    if (!pk || !pk->a_query || !pk->b_g2_query || !pk->h_query || !d_a || !d_b || !d_c || !d_partials ||
        (num_assign && !d_assignment) || (num_aux && !pk->l_query))
        return fail(B200_ERR_ARG, "null pointer");
    if (num_aux > num_assign) return fail(B200_ERR_ARG, "num_aux exceeds num_assign");
    if (log_n == 0 || log_n > 26) return fail(B200_ERR_ARG, "log_n = %u out of range [1, 26]", log_n);
    if (shards == 0 || shard >= shards) return fail(B200_ERR_ARG, "shard %u of %u", shard, shards);
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);
    return groth16_partial(E, family, pk, d_assignment, num_assign, num_aux, d_a, d_b, d_c, log_n, shard, shards, d_partials, st);
}

int b200_groth16_assemble_device(int family, const b200_groth16_pk *pk, const void *d_partials, unsigned shards, void *d_proof,
                                 void *stream) {
    if (family != B200_GROTH16_BLS12_377 && family != B200_GROTH16_BW6_761) return fail(B200_ERR_ARG, "unknown Groth16 family %d", family);
    if (!pk || !pk->a_query || !pk->b_g2_query || !pk->alpha_g1 || !pk->beta_g2 || !d_partials || !d_proof || shards == 0)
        return fail(B200_ERR_ARG, "null pointer / no shards");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    ENGINE_ORDER(st);
    return groth16_assemble(E, family, pk, d_partials, shards, d_proof, st);
}

int b200_field_op_device(int curve, int op, const void *d_a, const void *d_b, size_t n, void *d_out, void *stream) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    if (op < 0 || op > 14 || op == 7 || op == 12) return fail(B200_ERR_ARG, "unknown field op %d", op);   // 8..14: cooperative forms (no inverse)
    if (n && (!d_a || !d_b || !d_out)) return fail(B200_ERR_ARG, "null pointer");
    REQUIRE_ENGINE();
    cudaStream_t st = stream ? (cudaStream_t)stream : E.stream;
    return DISPATCH_CURVE(curve, field_op, op, d_a, d_b, n, d_out, st);
}

int b200_profile_enable(int on) {
    REQUIRE_ENGINE();
    if (on && !E.prof_ev[0])
        for (auto &ev : E.prof_ev) CUDA_TRY(cudaEventCreate(&ev));
    E.profile = on != 0;
    E.prof_used = 0;
    E.prof_units = 0;
    return B200_OK;
}

int b200_profile_read(double *accumulate_ms, int *launches, uint64_t *pairs) {
    REQUIRE_ENGINE();
    double ms = 0;
    for (int i = 0; i < E.prof_used; i++) {
        CUDA_TRY(cudaEventSynchronize(E.prof_ev[2 * i + 1]));
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, E.prof_ev[2 * i], E.prof_ev[2 * i + 1]));
        if (i > 0) {
            // consecutive launches of a pipelined batch overlap at their ends (two accumulate streams): a launch is charged
            // from the completion of the previous one, not from the moment it was queued behind it
            float since_prev = 0;
            if (cudaEventElapsedTime(&since_prev, E.prof_ev[2 * i - 1], E.prof_ev[2 * i + 1]) == cudaSuccess && since_prev > 0 && since_prev < t)
                t = since_prev;
        }
        ms += t;
    }
    if (accumulate_ms) *accumulate_ms = ms;
    if (launches) *launches = E.prof_used;
    if (pairs) *pairs = E.prof_units;
    E.prof_used = 0;
    E.prof_units = 0;
    return B200_OK;
}

int b200_sync(void *stream) {
    std::shared_ptr<Engine> Ep = engine_ref();         // no Engine::mu: a sync must not queue behind other callers
    if (!Ep) return fail(B200_ERR_STATE, "b200_init has not been called");
    CUDA_TRY(cudaSetDevice(Ep->device));
    CUDA_TRY(cudaStreamSynchronize(stream ? (cudaStream_t)stream : Ep->stream));
    return B200_OK;
}

int b200_msm_plan(int curve, size_t n, int *window_bits, int *windows, uint32_t *buckets_per_window) {
    CurveInfo ci;
    if (!curve_info(curve, ci)) return fail(B200_ERR_ARG, "unknown curve id %d", curve);
    return DISPATCH_CURVE(curve, plan_query, n, window_bits, windows, buckets_per_window);
}

}  // extern "C"
