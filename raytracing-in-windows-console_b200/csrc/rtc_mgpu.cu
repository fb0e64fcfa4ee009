// rtc_mgpu.cu -- RayTracingManager::Update across 1/2/4/8 GPUs of one box: row bands, ONE process.
//
// The reference's frame driver (RayTracingManager.cu:76-154) can only ever use one GPU.  Here the frame shards by rows:
// pixels are independent given the replicated scene (<= 0.26 MB) and the 96-byte camera block, so device g traces rows
// [rows[g], rows[g+1]) with no data-path collective.  One worker thread and one stream per device (a CUDA launch costs
// 2-4 us of host time; three launches x eight devices from one thread would be most of a 0.17 ms frame), frames pipelined
// kSlots deep, no Python, no process boundary, no shared-memory polling between processes.
//
// Two ways from bands to one frame (rtc_mgpu_create's `gather`):
//   RTC_GATHER_HOST  every device also ENCODES its band (rtc_encode_band semantics: the band above contributes one
//                    context row, so MinimizeRGB's latestColor carry-over across the seam is exact and the band streams
//                    concatenate to the frame's stream) and copies its piece over ITS OWN PCIe link into one pinned host
//                    frame at the offset given by the lengths of the devices before it (published through atomics).  The
//                    frame is assembled where the reference's sink wants it (PrintMachine::SetDataInBackBuffer) by N
//                    copy engines instead of one.
//   RTC_GATHER_P2P   the north-star layout: every device's ray kernel stores its quantised band straight into device 0's
//                    frame planes over NVLink (peer stores from the tile epilogue, 16-byte pieces), device 0 waits on the
//                    other devices' events, encodes the whole frame and copies the stream out over its one PCIe link.
//                    Planes are double-buffered so that device g may write frame k+1 while device 0 still encodes frame k.
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "rtc_ctx.h"

using rtc::fail;

namespace {

constexpr int kSlots = 3;          // frames in flight (submit k+2 before collect k)
constexpr int kPlaneSlots = 2;     // P2P: frame planes on device 0
constexpr int kMaxGpus = 16;
constexpr size_t kFlushBytes = 256u << 20;   // > 126 MB of L2

// The entry points below run on the CALLER's thread and touch several devices; the caller's current device (torch's, for
// one) must come back unchanged.
struct DeviceGuard {
    int dev = -1;
    DeviceGuard() { if (cudaGetDevice(&dev) != cudaSuccess) { dev = -1; (void)cudaGetLastError(); } }
    ~DeviceGuard() { if (dev >= 0) cudaSetDevice(dev); }
};

inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}

// A scene mutation.  Mutations are queued by the caller's thread and applied by every worker to its own context at the
// start of the next submitted frame, in order -- so each context's host copy is only ever touched by one thread, and the
// Scene3D calls stay legal while frames are in flight (Engine3D::Run adds a sphere every second).
struct SceneOp {
    enum Kind { SET_OBJECTS, CLEAR, ADD_SPHERE, ADD_PLANE, SET_LIGHT } kind = CLEAR;
    std::vector<rtc_object> objs;     // SET_OBJECTS
    rtc_object one{};                 // ADD_SPHERE / ADD_PLANE (plane: `normal` still un-normalised)
    rtc_light light{};                // SET_LIGHT
    bool default_light = false;
};
typedef std::vector<SceneOp> SceneOps;

int apply_scene_op(rtc_ctx* c, const SceneOp& op)
{
    switch (op.kind) {
    case SceneOp::SET_OBJECTS: return rtc_scene_set_objects(c, op.objs.data(), (uint32_t)op.objs.size());
    case SceneOp::CLEAR: return rtc_scene_clear(c);
    case SceneOp::ADD_SPHERE: return rtc_scene_add_sphere(c, op.one.center, op.one.radius, op.one.color, op.one.speed, op.one.mover);
    case SceneOp::ADD_PLANE: return rtc_scene_add_plane(c, op.one.center, op.one.normal, op.one.color, op.one.width, op.one.height);
    case SceneOp::SET_LIGHT: return rtc_set_light(c, op.default_light ? nullptr : &op.light);
    }
    return RTC_OK;
}

struct Cmd {
    enum Type { SUBMIT, QUIT } type = SUBMIT;
    long long frame = 0;
    rtc_params p{};
    int mode = RTC_RGB_PIXEL;
    double dt = 0.0;
    uint32_t flags = 0;
    bool flush = false;
    std::shared_ptr<const SceneOps> scene;   // scene mutations that take effect with this frame
};

struct FrameSlot {                                   // shared between the workers and the caller, one per frame slot
    std::atomic<long long> len_tag[kMaxGpus];        // frame id + 1 whose stream length device g has published
    std::atomic<long long> done_tag[kMaxGpus];       // frame id + 1 whose piece device g has landed in `host`
    unsigned long long len[kMaxGpus];
    float ms[kMaxGpus];                              // device time of the frame's kernels on device g (CUDA events)
    float enc_ms;                                    // P2P: encode time on device 0 ...
    float own_ms;                                    // ... and the time of its own band
    int rc[kMaxGpus];
    uint32_t rows[kMaxGpus + 1];
    uint32_t x = 0, y = 0;
    int mode = 0;
    char* host = nullptr;                            // pinned frame buffer
    size_t host_cap = 0;
    bool host_registered = false;                    // mmap + cudaHostRegister (NUMA placement) instead of cudaHostAlloc
};

struct InFlight {
    long long frame = 0;
    int stage = 0;                                   // 0 kernels running, 1 length published, 2 copy issued
    std::chrono::steady_clock::time_point t_done, t_issue;
};

struct Worker {
    int g = 0, device = 0;
    rtc_ctx* ctx = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Cmd> q;
    std::deque<InFlight> inflight;                   // frames enqueued on the device, oldest first
    rtc::DevBuf<uint8_t> d_color, d_glyph;           // HOST gather: this device's band planes (+1 context row)
    rtc::DevBuf<char> d_out[kSlots];
    unsigned long long* h_total = nullptr;           // [kSlots], mapped pinned: the emit kernel writes the length here
    cudaEvent_t ev_t0[kSlots] = {}, ev_t1[kSlots] = {}, ev_own[kSlots] = {}, ev_mid[kSlots] = {}, ev_copy[kSlots] = {};
    cudaEvent_t ev_band[kPlaneSlots] = {};           // P2P: this device's band of the frame in plane slot s has been written
    cudaStream_t copy_stream = nullptr;
    rtc::DevBuf<unsigned char> d_flush;
    // host-side accounting (rtc_mgpu_host_stats): microseconds spent enqueueing frames, from "kernels done" to "copy issued"
    // (waiting for the lengths of the devices before this one), and from "copy issued" to "copy landed"
    double us_enqueue = 0.0, us_wait_len = 0.0, us_copy = 0.0;
    unsigned long long n_frames = 0;
    double trace[64][6] = {};                        // rtc_mgpu_debug_trace: per frame % 64, host-clock us since the driver's epoch
};

}  // namespace

struct rtc_mgpu {
    std::chrono::steady_clock::time_point epoch = std::chrono::steady_clock::now();
    double main_trace[64][2] = {};                   // submit called, collect returned
    int n = 0;
    int gather = RTC_GATHER_HOST;
    Worker w[kMaxGpus];
    FrameSlot fr[kSlots];
    long long n_sub = 0, n_col = 0;
    bool flush_next = false;
    std::shared_ptr<SceneOps> pending_scene;         // scene mutations not yet handed to the workers
    // P2P
    rtc::DevBuf<uint8_t> plane_color[kPlaneSlots], plane_glyph[kPlaneSlots];   // on device 0
    cudaEvent_t ev_enc[kPlaneSlots] = {};                                      // device 0 has encoded the frame in plane slot s
    std::atomic<long long> enq_tag[kMaxGpus];        // frame id + 1 whose band event worker g has recorded
    std::atomic<long long> enc_tag{0};               // frame id + 1 whose encode event worker 0 has recorded
    double deficit_rows = 0.0;                       // P2P: rows device 0 gives up to pay for the encoder
    bool calibrated = false;
    bool user_bands = false;
    uint32_t user_rows[kMaxGpus + 1] = {};
    uint32_t user_y = 0;
    std::atomic<bool> failed{false};
    std::mutex err_mu;
    std::string err;
    int err_code = 0;
    float last_ms[kMaxGpus] = {};
    float last_enc_ms = 0.f, last_own_ms = 0.f;
    uint32_t last_rows[kMaxGpus + 1] = {};
};

namespace {

inline double us_since(const rtc_mgpu* m)
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - m->epoch).count();
}

void set_error(rtc_mgpu* m, int g, int code, const char* what)
{
    std::lock_guard<std::mutex> lk(m->err_mu);
    if (!m->failed.load()) {
        char buf[640];
        snprintf(buf, sizeof buf, "device slot %d: %s", g, what);
        m->err = buf;
        m->err_code = code;
        m->failed.store(true);
    }
}

// Alternative backings of the pinned frame buffer (experiments, RTC_MGPU_HOSTBUF): anonymous memory that is
// cudaHostRegister'ed -- "thp": transparent huge pages requested (fewer IOMMU translations for eight concurrent DMA
// streams); "interleave": pages interleaved over all NUMA nodes (mbind(2) without libnuma).  Measured on this pool's
// 8-GPU box (one socket, one NUMA node): no gain over cudaHostAlloc -- the fan-in of the eight D2H streams tops out at
// ~85 GB/s whatever the backing (profiles/r02_d2h_fanin.md), so cudaHostAlloc stays the default.
char* alloc_anonymous(size_t bytes, bool thp, bool interleave)
{
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
#ifdef MADV_HUGEPAGE
    if (thp) (void)madvise(p, bytes, MADV_HUGEPAGE);
#endif
#ifdef SYS_mbind
    if (interleave) {
        int n_nodes = 0;
        for (int i = 0; i < 64; ++i) {
            char path[64];
            snprintf(path, sizeof path, "/sys/devices/system/node/node%d", i);
            if (access(path, F_OK) == 0) n_nodes = i + 1;
        }
        if (n_nodes > 1) {
            unsigned long mask[16];
            memset(mask, 0, sizeof mask);
            for (int i = 0; i < n_nodes; ++i) mask[i / (8 * sizeof(long))] |= 1ul << (i % (8 * sizeof(long)));
            (void)syscall(SYS_mbind, p, bytes, 3 /* MPOL_INTERLEAVE */, mask, (unsigned long)n_nodes + 1, 0u);
        }
    }
#endif
    return static_cast<char*>(p);
}

int ensure_host_frame(rtc_mgpu* m, FrameSlot& f, size_t cap)
{
    if (cap <= f.host_cap) return RTC_OK;
    if (f.host) {
        if (f.host_registered) { cudaHostUnregister(f.host); munmap(f.host, f.host_cap); }
        else cudaFreeHost(f.host);
        f.host = nullptr; f.host_cap = 0;
    }
    const size_t want = (cap + cap / 8 + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    const char* hb = getenv("RTC_MGPU_HOSTBUF");
    if (hb && (!strcmp(hb, "interleave") || !strcmp(hb, "thp"))) {
        char* p = alloc_anonymous(want, !strcmp(hb, "thp"), !strcmp(hb, "interleave"));
        if (p) {
            memset(p, 0, want);                                  // fault the pages in under the interleave policy
            if (cudaHostRegister(p, want, cudaHostRegisterPortable) == cudaSuccess) {
                f.host = p; f.host_cap = want; f.host_registered = true;
                return RTC_OK;
            }
            (void)cudaGetLastError();
            munmap(p, want);
        }
    }
    void* p = nullptr;
    CK(cudaHostAlloc(&p, want, cudaHostAllocPortable));
    f.host = static_cast<char*>(p); f.host_cap = want; f.host_registered = false;
    return RTC_OK;
}

// Spin until cond() or failure / timeout; false on failure.
template <class F>
bool wait_for(rtc_mgpu* m, F cond)
{
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (!cond()) {
        if (m->failed.load(std::memory_order_relaxed)) return false;
        cpu_relax();
        if ((++spins & 0xffffu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) return false;
    }
    return true;
}

// ---- worker: enqueue one frame's device work --------------------------------------------------------------------------
int enqueue_frame(rtc_mgpu* m, Worker& w, const Cmd& cmd)
{
    const int g = w.g, n = m->n;
    const int slot = (int)(cmd.frame % kSlots);
    FrameSlot& f = m->fr[slot];
    rtc_ctx* c = w.ctx;
    const uint32_t x = cmd.p.x, W = x - 1u;
    const int mode = cmd.mode;
    const uint32_t bpp = rtc::mode_bpp(mode);
    const bool gl = rtc::mode_has_glyph(mode);
    const uint32_t r0 = f.rows[g], r1 = f.rows[g + 1];
    CK(cudaSetDevice(w.device));
    if (cmd.flush) {
        CK(w.d_flush.ensure(kFlushBytes));
        CK(cudaMemsetAsync(w.d_flush.p, 0, kFlushBytes, c->stream));
    }
    int rc = RTC_OK;
    if (cmd.scene) {                                            // scene mutations queued before this frame (uploads are staged in stream order)
        for (const SceneOp& op : *cmd.scene) {
            rc = apply_scene_op(c, op);
            if (rc) return rc;
        }
    }
    rc = rtc_update_objects(c, cmd.dt, cmd.flags);              // RayTracingManager::Update runs the physics step first
    if (rc) return rc;
    CK(cudaEventRecord(w.ev_t0[slot], c->stream));
    if (m->gather == RTC_GATHER_HOST) {
        const uint32_t c0 = r0 > 0 ? r0 - 1u : 0u;              // one context row above the band (seam carry-over)
        const size_t px_t = (size_t)(r1 - c0) * W;
        CK(w.d_color.ensure(px_t * bpp + 64));
        if (gl) CK(w.d_glyph.ensure(px_t + 64));
        const size_t cap = rtc_encode_capacity(x, r1 > r0 ? r1 - r0 : 1u, (rtc_mode)mode);
        CK(w.d_out[slot].ensure(cap));
        if (r1 > r0) {
            rc = rtc::trace_shade(c, &cmd.p, mode, cmd.flags & ~(uint32_t)RTC_FLAG_KEEP_HITS, c0, r1, w.d_color.p, gl ? w.d_glyph.p : nullptr, false);
            if (rc) return rc;
            const size_t skip = (size_t)(r0 - c0) * W;
            rc = rtc::do_encode(c, w.d_color.p + skip * bpp, gl ? w.d_glyph.p + skip : nullptr, x, r1 - r0, mode, w.d_out[slot].p, cap,
                                w.h_total + slot, r0 > 0);
            if (rc) return rc;
        } else {
            w.h_total[slot] = 0ull;                              // an empty band contributes an empty stream
        }
        CK(cudaEventRecord(w.ev_t1[slot], c->stream));
        return RTC_OK;
    }
    // ---- P2P: bands straight into device 0's planes ----------------------------------------------------------------
    const int ps = (int)(cmd.frame % kPlaneSlots);
    if (g != 0 && cmd.frame >= kPlaneSlots) {                    // device 0 must have encoded the frame that used this plane slot
        if (!wait_for(m, [&] { return m->enc_tag.load(std::memory_order_acquire) >= cmd.frame - kPlaneSlots + 1; }))
            return fail(RTC_ERR_CUDA, "timed out waiting for device 0's encoder");
        CK(cudaStreamWaitEvent(c->stream, m->ev_enc[ps], 0));
    }
    if (r1 > r0) {
        rc = rtc::trace_shade(c, &cmd.p, mode, cmd.flags & ~(uint32_t)RTC_FLAG_KEEP_HITS, r0, r1, m->plane_color[ps].p + (size_t)r0 * W * bpp,
                              gl ? m->plane_glyph[ps].p + (size_t)r0 * W : nullptr, false);
        if (rc) return rc;
    }
    CK(cudaEventRecord(w.ev_band[ps], c->stream));
    m->enq_tag[g].store(cmd.frame + 1, std::memory_order_release);
    if (g != 0) {
        CK(cudaEventRecord(w.ev_t1[slot], c->stream));
        return RTC_OK;
    }
    CK(cudaEventRecord(w.ev_own[slot], c->stream));            // device 0's own band is done
    for (int h = 1; h < n; ++h) {
        if (!wait_for(m, [&] { return m->enq_tag[h].load(std::memory_order_acquire) >= cmd.frame + 1; }))
            return fail(RTC_ERR_CUDA, "timed out waiting for device slot %d to enqueue its band", h);
        CK(cudaStreamWaitEvent(c->stream, m->w[h].ev_band[ps], 0));
    }
    CK(cudaEventRecord(w.ev_mid[slot], c->stream));            // every band has landed: what follows is the encode alone
    const size_t cap = rtc_encode_capacity(x, f.y, (rtc_mode)mode);
    CK(w.d_out[slot].ensure(cap));
    rc = rtc::do_encode(c, m->plane_color[ps].p, gl ? m->plane_glyph[ps].p : nullptr, x, f.y, mode, w.d_out[slot].p, cap, w.h_total + slot, false);
    if (rc) return rc;
    CK(cudaEventRecord(w.ev_t1[slot], c->stream));
    CK(cudaEventRecord(m->ev_enc[ps], c->stream));
    m->enc_tag.store(cmd.frame + 1, std::memory_order_release);
    return RTC_OK;
}

// One non-blocking pass over this worker's in-flight frames; true if something moved.  Every frame walks through
//   0 kernels running -> 1 length published -> 2 copy issued -> landed (popped)
// on its own: the length of frame k+1 is published the moment its kernels finish, even while the copy of frame k is
// still on the wire (the devices behind this one need that length to place THEIR pieces of frame k+1 -- holding it back
// until the older copy has landed cost them > 100 us per frame).  Copies are issued in frame order.
bool progress(rtc_mgpu* m, Worker& w)
{
    if (w.inflight.empty()) return false;
    const int g = w.g;
    const bool owns_stream = m->gather == RTC_GATHER_HOST || g == 0;     // has a piece of the stream to land
    bool moved = false;
    bool older_copies_issued = true;
    for (size_t i = 0; i < w.inflight.size();) {
        InFlight& e = w.inflight[i];
        const long long j = e.frame;
        const int slot = (int)(j % kSlots);
        FrameSlot& f = m->fr[slot];
        auto finish = [&](int rc) {                              // only ever reached for the oldest frame (stream order)
            f.rc[g] = rc;
            if (rc) {
                f.len[g] = 0;
                set_error(m, g, rc, rtc_last_error());
                f.len_tag[g].store(j + 1, std::memory_order_release);
            }
            f.done_tag[g].store(j + 1, std::memory_order_release);
            w.inflight.erase(w.inflight.begin() + (long)i);
            moved = true;
        };
        if (m->failed.load(std::memory_order_relaxed)) { finish(f.rc[g] ? f.rc[g] : RTC_ERR_CUDA); continue; }
        if (e.stage == 0) {
            static const bool serial = getenv("RTC_MGPU_SERIAL_STAGES") != nullptr;   // experiment: the first version's behaviour
            if (serial && i != 0) break;
            const cudaError_t q = cudaEventQuery(w.ev_t1[slot]);
            if (q == cudaErrorNotReady) break;                   // the frames behind it run on the same stream: not ready either
            if (q != cudaSuccess) { fail(RTC_ERR_CUDA, "frame %lld failed on the device: %s", j, cudaGetErrorString(q)); finish(RTC_ERR_CUDA); continue; }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, w.ev_t0[slot], w.ev_t1[slot]);
            f.ms[g] = ms;
            if (m->gather == RTC_GATHER_P2P && g == 0) {
                float enc = 0.f, own = 0.f;
                cudaEventElapsedTime(&enc, w.ev_mid[slot], w.ev_t1[slot]);
                cudaEventElapsedTime(&own, w.ev_t0[slot], w.ev_own[slot]);
                f.enc_ms = enc;
                f.own_ms = own;
            }
            f.len[g] = owns_stream ? w.h_total[slot] : 0ull;
            f.len_tag[g].store(j + 1, std::memory_order_release);
            e.stage = 1;
            e.t_done = std::chrono::steady_clock::now();
            w.trace[j & 63][2] = us_since(m);
            w.trace[j & 63][5] = ms * 1000.0;
            moved = true;
            if (!owns_stream) { finish(RTC_OK); continue; }
        }
        if (e.stage == 1) {
            bool ready = older_copies_issued;
            unsigned long long off = 0;
            if (ready && m->gather == RTC_GATHER_HOST) {
                for (int h = 0; h < g; ++h) {
                    if (f.len_tag[h].load(std::memory_order_acquire) < j + 1) { ready = false; break; }    // a length before mine is not known yet
                    off += f.len[h];
                }
            }
            if (ready) {
                const unsigned long long nbytes = f.len[g];
                if (off + nbytes > f.host_cap) { fail(RTC_ERR_CAPACITY, "stream piece [%llu, +%llu) exceeds the frame buffer (%zu B)", off, nbytes, f.host_cap); finish(RTC_ERR_CAPACITY); continue; }
                cudaError_t ce = cudaSuccess;
                if (nbytes) ce = cudaMemcpyAsync(f.host + off, w.d_out[slot].p, nbytes, cudaMemcpyDeviceToHost, w.copy_stream);
                if (ce == cudaSuccess) ce = cudaEventRecord(w.ev_copy[slot], w.copy_stream);
                if (ce != cudaSuccess) { fail(RTC_ERR_CUDA, "D2H of the band stream failed: %s", cudaGetErrorString(ce)); finish(RTC_ERR_CUDA); continue; }
                e.stage = 2;
                e.t_issue = std::chrono::steady_clock::now();
                w.trace[j & 63][3] = us_since(m);
                w.us_wait_len += std::chrono::duration<double, std::micro>(e.t_issue - e.t_done).count();
                moved = true;
            } else {
                older_copies_issued = false;
            }
        }
        if (e.stage == 2 && i == 0) {
            const cudaError_t q = cudaEventQuery(w.ev_copy[slot]);
            if (q != cudaErrorNotReady) {
                if (q != cudaSuccess) { fail(RTC_ERR_CUDA, "D2H of the band stream failed: %s", cudaGetErrorString(q)); finish(RTC_ERR_CUDA); continue; }
                w.us_copy += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - e.t_issue).count();
                w.trace[j & 63][4] = us_since(m);
                finish(RTC_OK);
                continue;
            }
        }
        ++i;
    }
    return moved;
}

void worker_main(rtc_mgpu* m, Worker* w)
{
    cudaSetDevice(w->device);
    for (;;) {
        Cmd cmd;
        bool have = false;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            if (w->q.empty() && w->inflight.empty()) w->cv.wait(lk, [&] { return !w->q.empty(); });
            if (!w->q.empty()) { cmd = w->q.front(); w->q.pop_front(); have = true; }
        }
        if (have) {
            if (cmd.type == Cmd::QUIT) break;
            const int slot = (int)(cmd.frame % kSlots);
            const auto tq0 = std::chrono::steady_clock::now();
            w->trace[cmd.frame & 63][0] = us_since(m);
            int rc = m->failed.load() ? RTC_ERR_CUDA : enqueue_frame(m, *w, cmd);
            w->trace[cmd.frame & 63][1] = us_since(m);
            w->us_enqueue += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tq0).count();
            w->n_frames++;
            if (rc) {                                            // publish the failure so that nobody waits for this band
                FrameSlot& f = m->fr[slot];
                if (!m->failed.load()) set_error(m, w->g, rc, rtc_last_error());
                f.rc[w->g] = rc; f.len[w->g] = 0;
                m->enq_tag[w->g].store(cmd.frame + 1, std::memory_order_release);
                if (w->g == 0) m->enc_tag.store(cmd.frame + 1, std::memory_order_release);
                f.len_tag[w->g].store(cmd.frame + 1, std::memory_order_release);
                f.done_tag[w->g].store(cmd.frame + 1, std::memory_order_release);
            } else {
                InFlight fl;
                fl.frame = cmd.frame;
                w->inflight.push_back(fl);
            }
        }
        const bool moved = progress(m, *w);
        if (!have && !moved) cpu_relax();
    }
    // drain
    while (!w->inflight.empty()) { if (!progress(m, *w)) cpu_relax(); }
}

void plan_rows(rtc_mgpu* m, uint32_t x, uint32_t y, uint32_t* rows)
{
    if (m->user_bands && m->user_y == y) { memcpy(rows, m->user_rows, sizeof(uint32_t) * (m->n + 1)); return; }
    if (m->gather == RTC_GATHER_P2P && m->n > 1) {
        // device 0 also encodes: it gets `deficit_rows` fewer rows; boundaries on the ray kernel's 16-row tile rows, and
        // never so that the other devices need a second tile wave when one would do
        const uint32_t tiles_x = (x - 1u + 15u) / 16u;
        const uint32_t wave_units = tiles_x ? (uint32_t)(m->w[0].ctx->sm_count * 28) / tiles_x : 0u;
        rtc_plan_bands(y, m->n, 16, m->deficit_rows, wave_units, rows);
        return;
    }
    rtc_plan_bands(y, m->n, 1, 0.0, 0, rows);
}

}  // namespace

extern "C" {

// Contiguous row bands that tile [0, y) exactly.  deficit_rows > 0: band 0 gets that many rows fewer than the others
// (device 0 also runs the encoder in the P2P gather); align > 1 puts the boundaries on multiples of `align` rows (16 = the
// ray kernel's tile height: a band that ends inside a tile row pays for the whole row); wave_units > 0: the number of
// align-row units one tile wave covers -- if the deficit would push the other bands just over one wave while everything
// fits into one wave per device, band 0 takes the excess instead.  Host-only (no GPU needed).
int rtc_plan_bands(uint32_t y, int n, uint32_t align, double deficit_rows, uint32_t wave_units, uint32_t* rows_out)
{
    if (n < 1 || n > kMaxGpus || !rows_out) return fail(RTC_ERR_INVALID, "rtc_plan_bands: invalid argument");
    const double d = deficit_rows > 0.0 ? deficit_rows : 0.0;
    if (n == 1) { rows_out[0] = 0; rows_out[1] = y; return RTC_OK; }
    if (d == 0.0 && align <= 1) {
        for (int g = 0; g <= n; ++g) rows_out[g] = (uint32_t)(((uint64_t)y * (uint64_t)g) / (uint64_t)n);
        return RTC_OK;
    }
    if (align < 1) align = 1;
    const uint64_t units = ((uint64_t)y + align - 1) / align;
    const double du = d / (double)align;
    double n0d = ((double)units + du) / (double)n - du;
    long long n0 = (long long)(n0d + 0.5);
    if (n0d < 0.0) n0 = 0;
    if (n0 < 0) n0 = 0;
    if ((uint64_t)n0 > units) n0 = (long long)units;
    if (wave_units > 0 && units <= (uint64_t)wave_units * (uint64_t)n) {
        const uint64_t rest = units - (uint64_t)n0;
        const uint64_t per_other = (rest + (uint64_t)(n - 2)) / (uint64_t)(n - 1);
        if (per_other > wave_units) n0 = (long long)(units - (uint64_t)wave_units * (uint64_t)(n - 1));
    }
    const uint64_t rest = units - (uint64_t)n0;
    rows_out[0] = 0;
    for (int g = 1; g <= n; ++g) {
        const uint64_t e = (uint64_t)n0 + (rest * (uint64_t)(g - 1)) / (uint64_t)(n - 1);
        const uint64_t r = e * align;
        rows_out[g] = (uint32_t)(r < y ? r : y);
    }
    rows_out[n] = y;
    return RTC_OK;
}

int rtc_mgpu_create(rtc_mgpu** out, int n_gpus, const int* device_ids, int gather)
{
    if (!out) return fail(RTC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_gpus < 1 || n_gpus > kMaxGpus) return fail(RTC_ERR_INVALID, "n_gpus %d out of range (1..%d)", n_gpus, kMaxGpus);
    if (gather != RTC_GATHER_HOST && gather != RTC_GATHER_P2P) return fail(RTC_ERR_INVALID, "unknown gather mode %d", gather);
    DeviceGuard guard;
    rtc_mgpu* m = new (std::nothrow) rtc_mgpu();
    if (!m) return fail(RTC_ERR_NOMEM, "out of host memory");
    m->n = n_gpus;
    m->gather = gather;
    for (int s = 0; s < kSlots; ++s)
        for (int g = 0; g < kMaxGpus; ++g) { m->fr[s].len_tag[g].store(0); m->fr[s].done_tag[g].store(0); m->fr[s].rc[g] = 0; }
    for (int g = 0; g < kMaxGpus; ++g) m->enq_tag[g].store(0);
#define CKM(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e2_ = (call);                                                                          \
        if (e2_ != cudaSuccess) {                                                                          \
            fail(RTC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e2_), __FILE__, __LINE__); \
            rtc_mgpu_destroy(m);                                                                           \
            return RTC_ERR_CUDA;                                                                           \
        }                                                                                                  \
    } while (0)
    for (int g = 0; g < n_gpus; ++g) {
        Worker& w = m->w[g];
        w.g = g;
        w.device = device_ids ? device_ids[g] : g;
        int rc = rtc_create(&w.ctx, w.device);
        if (rc) { rtc_mgpu_destroy(m); return rc; }
        CKM(cudaSetDevice(w.device));
        CKM(cudaStreamCreateWithFlags(&w.copy_stream, cudaStreamNonBlocking));
        for (int s = 0; s < kSlots; ++s) {
            CKM(cudaEventCreate(&w.ev_t0[s]));
            CKM(cudaEventCreate(&w.ev_t1[s]));
            CKM(cudaEventCreate(&w.ev_mid[s]));
            CKM(cudaEventCreate(&w.ev_own[s]));
            CKM(cudaEventCreateWithFlags(&w.ev_copy[s], cudaEventDisableTiming));
        }
        for (int s = 0; s < kPlaneSlots; ++s) CKM(cudaEventCreateWithFlags(&w.ev_band[s], cudaEventDisableTiming));
        void* ht = nullptr;
        CKM(cudaHostAlloc(&ht, sizeof(unsigned long long) * kSlots, cudaHostAllocPortable | cudaHostAllocMapped));
        w.h_total = static_cast<unsigned long long*>(ht);
        memset(w.h_total, 0, sizeof(unsigned long long) * kSlots);
        if (gather == RTC_GATHER_P2P && g > 0 && w.device != m->w[0].device) {
            int can = 0;
            CKM(cudaDeviceCanAccessPeer(&can, w.device, m->w[0].device));
            if (!can) {
                fail(RTC_ERR_CUDA, "device %d cannot access device %d's memory: no P2P gather on this box", w.device, m->w[0].device);
                rtc_mgpu_destroy(m);
                return RTC_ERR_CUDA;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(m->w[0].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CKM(e);
            (void)cudaGetLastError();
        }
    }
    if (gather == RTC_GATHER_P2P) {
        CKM(cudaSetDevice(m->w[0].device));
        for (int s = 0; s < kPlaneSlots; ++s) CKM(cudaEventCreateWithFlags(&m->ev_enc[s], cudaEventDisableTiming));
    }
#undef CKM
    for (int g = 0; g < n_gpus; ++g) m->w[g].th = std::thread(worker_main, m, &m->w[g]);
    *out = m;
    return RTC_OK;
}

void rtc_mgpu_destroy(rtc_mgpu* m)
{
    if (!m) return;
    DeviceGuard guard;
    for (int g = 0; g < m->n; ++g) {
        Worker& w = m->w[g];
        if (w.th.joinable()) {
            { std::lock_guard<std::mutex> lk(w.mu); Cmd q; q.type = Cmd::QUIT; w.q.push_back(q); }
            w.cv.notify_one();
            w.th.join();
        }
    }
    for (int g = 0; g < m->n; ++g) {
        Worker& w = m->w[g];
        cudaSetDevice(w.device);
        if (w.ctx) cudaStreamSynchronize(w.ctx->stream);
        if (w.copy_stream) { cudaStreamSynchronize(w.copy_stream); cudaStreamDestroy(w.copy_stream); }
        w.d_color.release(); w.d_glyph.release(); w.d_flush.release();
        for (int s = 0; s < kSlots; ++s) {
            w.d_out[s].release();
            if (w.ev_t0[s]) cudaEventDestroy(w.ev_t0[s]);
            if (w.ev_t1[s]) cudaEventDestroy(w.ev_t1[s]);
            if (w.ev_mid[s]) cudaEventDestroy(w.ev_mid[s]);
            if (w.ev_own[s]) cudaEventDestroy(w.ev_own[s]);
            if (w.ev_copy[s]) cudaEventDestroy(w.ev_copy[s]);
        }
        for (int s = 0; s < kPlaneSlots; ++s) if (w.ev_band[s]) cudaEventDestroy(w.ev_band[s]);
        if (w.h_total) cudaFreeHost(w.h_total);
    }
    if (m->n > 0) {
        cudaSetDevice(m->w[0].device);
        for (int s = 0; s < kPlaneSlots; ++s) {
            m->plane_color[s].release(); m->plane_glyph[s].release();
            if (m->ev_enc[s]) cudaEventDestroy(m->ev_enc[s]);
        }
    }
    for (int s = 0; s < kSlots; ++s) {
        FrameSlot& f = m->fr[s];
        if (f.host) {
            if (f.host_registered) { cudaHostUnregister(f.host); munmap(f.host, f.host_cap); }
            else cudaFreeHost(f.host);
        }
    }
    for (int g = 0; g < m->n; ++g) if (m->w[g].ctx) rtc_destroy(m->w[g].ctx);
    delete m;
}

int rtc_mgpu_count(rtc_mgpu* m) { return m ? m->n : 0; }

int rtc_mgpu_context(rtc_mgpu* m, int i, rtc_ctx** out)
{
    if (!m || !out || i < 0 || i >= m->n) return fail(RTC_ERR_INVALID, "rtc_mgpu_context: invalid argument");
    *out = m->w[i].ctx;
    return RTC_OK;
}

// ---- scene: replicated on every device (<= 0.26 MB) ------------------------------------------------------------------
// Mutations are only queued here (see SceneOp); they travel to the workers with the next submit.  With an idle
// pipeline (get_objects) they are applied on the spot.
static int apply_pending_scene(rtc_mgpu* m)
{
    if (!m->pending_scene) return RTC_OK;
    for (int g = 0; g < m->n; ++g)
        for (const SceneOp& op : *m->pending_scene) {
            const int rc = apply_scene_op(m->w[g].ctx, op);
            if (rc) return rc;
        }
    m->pending_scene.reset();
    return RTC_OK;
}
static SceneOp& queue_op(rtc_mgpu* m, SceneOp::Kind kind)
{
    if (!m->pending_scene) m->pending_scene = std::make_shared<SceneOps>();
    if (kind == SceneOp::SET_OBJECTS || kind == SceneOp::CLEAR) {   // everything queued before a replacement is dead
        SceneOps kept;
        for (SceneOp& op : *m->pending_scene) if (op.kind == SceneOp::SET_LIGHT) kept.push_back(std::move(op));
        m->pending_scene->swap(kept);
    }
    m->pending_scene->emplace_back();
    m->pending_scene->back().kind = kind;
    return m->pending_scene->back();
}

int rtc_mgpu_scene_clear(rtc_mgpu* m)
{
    if (!m) return fail(RTC_ERR_INVALID, "mgpu is NULL");
    queue_op(m, SceneOp::CLEAR);
    return RTC_OK;
}
int rtc_mgpu_scene_add_sphere(rtc_mgpu* m, const float center[3], float radius, const float rgb[3], float speed, int mover)
{
    if (!m || !center || !rgb) return fail(RTC_ERR_INVALID, "NULL argument");
    SceneOp& op = queue_op(m, SceneOp::ADD_SPHERE);
    for (int i = 0; i < 3; ++i) { op.one.center[i] = center[i]; op.one.color[i] = rgb[i]; }
    op.one.radius = radius; op.one.speed = speed; op.one.mover = mover;
    return RTC_OK;
}
int rtc_mgpu_scene_add_plane(rtc_mgpu* m, const float center[3], const float normal[3], const float rgb[3], float width, float height)
{
    if (!m || !center || !normal || !rgb) return fail(RTC_ERR_INVALID, "NULL argument");
    SceneOp& op = queue_op(m, SceneOp::ADD_PLANE);
    for (int i = 0; i < 3; ++i) { op.one.center[i] = center[i]; op.one.color[i] = rgb[i]; op.one.normal[i] = normal[i]; }
    op.one.width = width; op.one.height = height;
    return RTC_OK;
}
int rtc_mgpu_set_light(rtc_mgpu* m, const rtc_light* light)
{
    if (!m) return fail(RTC_ERR_INVALID, "mgpu is NULL");
    SceneOp& op = queue_op(m, SceneOp::SET_LIGHT);
    if (light) op.light = *light; else op.default_light = true;
    return RTC_OK;
}

int rtc_mgpu_scene_set_objects(rtc_mgpu* m, const rtc_object* objs, uint32_t n)
{
    if (!m || (n && !objs)) return fail(RTC_ERR_INVALID, "NULL argument");
    for (uint32_t i = 0; i < n; ++i)
        if (objs[i].type != RTC_OBJ_SPHERE && objs[i].type != RTC_OBJ_PLANE)
            return fail(RTC_ERR_INVALID, "object %u has unknown type %d", i, objs[i].type);
    queue_op(m, SceneOp::SET_OBJECTS).objs.assign(objs, objs + n);
    return RTC_OK;
}

int rtc_mgpu_scene_get_objects(rtc_mgpu* m, rtc_object* out, uint32_t cap, uint32_t* n)
{
    if (!m) return fail(RTC_ERR_INVALID, "mgpu is NULL");
    if (m->n_sub != m->n_col) return fail(RTC_ERR_INVALID, "frames are in flight: collect them first");
    DeviceGuard guard;
    const int rc = apply_pending_scene(m);
    if (rc) return rc;
    return rtc_scene_get_objects(m->w[0].ctx, out, cap, n);     // every replica ran the same physics steps
}

int rtc_mgpu_set_bands(rtc_mgpu* m, uint32_t y, const uint32_t* rows)
{
    if (!m) return fail(RTC_ERR_INVALID, "mgpu is NULL");
    if (!rows) { m->user_bands = false; return RTC_OK; }
    if (rows[0] != 0 || rows[m->n] != y) return fail(RTC_ERR_INVALID, "bands must tile [0, y)");
    for (int g = 0; g < m->n; ++g) if (rows[g] > rows[g + 1]) return fail(RTC_ERR_INVALID, "bands must be ascending");
    memcpy(m->user_rows, rows, sizeof(uint32_t) * (m->n + 1));
    m->user_y = y;
    m->user_bands = true;
    return RTC_OK;
}

int rtc_mgpu_flush_l2(rtc_mgpu* m)
{
    if (!m) return fail(RTC_ERR_INVALID, "mgpu is NULL");
    m->flush_next = true;
    return RTC_OK;
}

int rtc_mgpu_submit(rtc_mgpu* m, const rtc_params* p, rtc_mode mode, double dt, uint32_t flags)
{
    if (!m || !p) return fail(RTC_ERR_INVALID, "NULL argument");
    if (m->failed.load()) { std::lock_guard<std::mutex> lk(m->err_mu); return fail(m->err_code ? m->err_code : RTC_ERR_CUDA, "%s", m->err.c_str()); }
    if (mode < RTC_BIT_ASCII || mode > RTC_SDL) return fail(RTC_ERR_INVALID, "invalid rendering mode %d", (int)mode);
    if (p->x < 1 || p->y < 1) return fail(RTC_ERR_INVALID, "invalid console size %ux%u", p->x, p->y);
    if ((uint64_t)(p->x - 1u) * p->y >= (1ull << 31)) return fail(RTC_ERR_CAPACITY, "console size too large");
    if (m->n_sub - m->n_col >= kSlots) return fail(RTC_ERR_INVALID, "%d frames are already in flight: rtc_mgpu_collect one first", kSlots);
    DeviceGuard guard;
    const long long j = m->n_sub;
    const int slot = (int)(j % kSlots);
    FrameSlot& f = m->fr[slot];
    f.x = p->x; f.y = p->y; f.mode = mode;
    // P2P: once a few frames have been timed, give device 0 a smaller band to pay for the encoder
    if (m->gather == RTC_GATHER_P2P && m->n > 1 && !m->calibrated && !m->user_bands && m->n_col >= 3) {
        const uint32_t rows0 = m->last_rows[1] - m->last_rows[0];
        const float band_ms = m->last_own_ms;
        if (rows0 > 0 && band_ms > 0.f) m->deficit_rows = (double)m->last_enc_ms / ((double)band_ms / (double)rows0);
        m->calibrated = true;
    }
    plan_rows(m, p->x, p->y, f.rows);
    const size_t cap = rtc_encode_capacity(p->x, p->y, mode);
    if (cap > f.host_cap) {
        int rc = ensure_host_frame(m, f, cap);
        if (rc) return rc;
    }
    if (m->gather == RTC_GATHER_P2P) {
        const int ps = (int)(j % kPlaneSlots);
        const size_t n_px = (size_t)(p->x - 1u) * p->y;
        const size_t need_c = n_px * 3u + 64, need_g = n_px + 64;     // sized for every rendering mode: modes may alternate frame by frame
        if (need_c > m->plane_color[ps].cap || need_g > m->plane_glyph[ps].cap) {
            if (m->n_sub != m->n_col) return fail(RTC_ERR_INVALID, "the frame grew: collect the frames in flight before submitting a larger one");
            CK(cudaSetDevice(m->w[0].device));
            for (int s = 0; s < kPlaneSlots; ++s) {
                CK(m->plane_color[s].ensure(need_c));
                CK(m->plane_glyph[s].ensure(need_g));
            }
        }
    }
    Cmd cmd;
    cmd.type = Cmd::SUBMIT; cmd.frame = j; cmd.p = *p; cmd.mode = mode; cmd.dt = dt; cmd.flags = flags; cmd.flush = m->flush_next;
    cmd.scene = std::move(m->pending_scene);
    m->pending_scene.reset();
    m->flush_next = false;
    for (int g = 0; g < m->n; ++g) {
        Worker& w = m->w[g];
        { std::lock_guard<std::mutex> lk(w.mu); w.q.push_back(cmd); }
        w.cv.notify_one();
    }
    m->main_trace[j & 63][0] = us_since(m);
    m->n_sub = j + 1;
    return RTC_OK;
}

int rtc_mgpu_collect(rtc_mgpu* m, const char** host_ptr, size_t* n_bytes)
{
    if (!m || !host_ptr || !n_bytes) return fail(RTC_ERR_INVALID, "NULL argument");
    if (m->n_col >= m->n_sub) return fail(RTC_ERR_INVALID, "no frame in flight");
    const long long j = m->n_col;
    const int slot = (int)(j % kSlots);
    FrameSlot& f = m->fr[slot];
    bool ok = true;
    for (int g = 0; g < m->n; ++g) {
        const auto t0 = std::chrono::steady_clock::now();
        unsigned spins = 0;
        while (f.done_tag[g].load(std::memory_order_acquire) < j + 1) {
            cpu_relax();
            if ((++spins & 0xffffu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) { ok = false; break; }
        }
        if (!ok) break;
    }
    m->n_col = j + 1;
    if (!ok) { set_error(m, -1, RTC_ERR_CUDA, "timed out waiting for a frame"); }
    if (m->failed.load()) { std::lock_guard<std::mutex> lk(m->err_mu); return fail(m->err_code ? m->err_code : RTC_ERR_CUDA, "%s", m->err.c_str()); }
    size_t total = 0;
    for (int g = 0; g < m->n; ++g) { total += (size_t)f.len[g]; m->last_ms[g] = f.ms[g]; }
    m->last_enc_ms = f.enc_ms;
    m->last_own_ms = f.own_ms;
    memcpy(m->last_rows, f.rows, sizeof(uint32_t) * (m->n + 1));
    *host_ptr = f.host;
    *n_bytes = total;
    m->main_trace[j & 63][1] = us_since(m);
    return RTC_OK;
}

int rtc_mgpu_update(rtc_mgpu* m, const rtc_params* p, rtc_mode mode, double dt, uint32_t flags, const char** host_ptr, size_t* n_bytes)
{
    if (!m) return fail(RTC_ERR_INVALID, "mgpu is NULL");
    if (m->n_sub != m->n_col) return fail(RTC_ERR_INVALID, "frames are in flight: use rtc_mgpu_submit / rtc_mgpu_collect");
    int rc = rtc_mgpu_submit(m, p, mode, dt, flags);
    if (rc) return rc;
    return rtc_mgpu_collect(m, host_ptr, n_bytes);
}

// Host-side accounting since the last call (idle pipeline only): per device, average microseconds per frame spent
// enqueueing (scene staging + launches), between "kernels done" and "copy issued", and in the D2H copy.  out[3 * n_gpus].
int rtc_mgpu_host_stats(rtc_mgpu* m, float* out)
{
    if (!m || !out) return fail(RTC_ERR_INVALID, "NULL argument");
    if (m->n_sub != m->n_col) return fail(RTC_ERR_INVALID, "frames are in flight: collect them first");
    for (int g = 0; g < m->n; ++g) {
        Worker& w = m->w[g];
        if (!wait_for(m, [&] { std::lock_guard<std::mutex> lk(w.mu); return w.q.empty(); })) return fail(RTC_ERR_CUDA, "workers are busy");
        const double n = w.n_frames ? (double)w.n_frames : 1.0;
        out[3 * g + 0] = (float)(w.us_enqueue / n); out[3 * g + 1] = (float)(w.us_wait_len / n); out[3 * g + 2] = (float)(w.us_copy / n);
        w.us_enqueue = w.us_wait_len = w.us_copy = 0.0; w.n_frames = 0;
    }
    return RTC_OK;
}

// Debug: host-clock timeline (us since creation) of the last 64 frames.  out[(g * 64 + k) * 6 + i], k = frame % 64:
// i = 0 enqueue start, 1 enqueue end, 2 kernels seen finished, 3 copy issued, 4 copy landed, 5 device time of the frame
// (us); then out[n * 64 * 6 + k * 2 + {0, 1}] = rtc_mgpu_submit called / rtc_mgpu_collect returned.  Idle pipeline only.
int rtc_mgpu_debug_trace(rtc_mgpu* m, double* out)
{
    if (!m || !out) return fail(RTC_ERR_INVALID, "NULL argument");
    if (m->n_sub != m->n_col) return fail(RTC_ERR_INVALID, "frames are in flight: collect them first");
    for (int g = 0; g < m->n; ++g) memcpy(out + (size_t)g * 64 * 6, m->w[g].trace, sizeof(double) * 64 * 6);
    memcpy(out + (size_t)m->n * 64 * 6, m->main_trace, sizeof(double) * 64 * 2);
    return RTC_OK;
}

int rtc_mgpu_last_frame(rtc_mgpu* m, float* device_ms, uint32_t* rows, float* encode_ms)
{
    if (!m) return fail(RTC_ERR_INVALID, "mgpu is NULL");
    if (m->n_col == 0) return fail(RTC_ERR_INVALID, "no frame collected");
    if (device_ms) memcpy(device_ms, m->last_ms, sizeof(float) * m->n);
    if (rows) memcpy(rows, m->last_rows, sizeof(uint32_t) * (m->n + 1));
    if (encode_ms) *encode_ms = m->last_enc_ms;
    return RTC_OK;
}

}  // extern "C"
