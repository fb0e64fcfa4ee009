/*
 * rtc.h -- C-ABI of the B200-native console ray tracer hot path ("rtc").
 *
 * This is the drop-in boundary for ONE path of EmilHogstedt/Raytracing-in-Windows-Console:
 * everything inside RayTracingManager::Update (reference RayTracingManager.cu:76-154):
 * camera ray generation -> nearest-hit ray/sphere + ray/plane -> Blinn-Phong shading ->
 * integer quantisation -> ANSI escape stream (the bytes handed to
 * PrintMachine::SetDataInBackBuffer, reference PrintMachine.cpp:178-192).
 *
 * The reference has no FFI; the seam is a C++ call.  The C++ facade classes in
 * raytracing-in-windows-console_b200/host/ keep the reference's class names and signatures
 * (Engine3D / Scene3D / Object3D / Sphere / Plane / Camera3D / RayTracingManager /
 * PrintMachine) and forward to the entry points below.  Each entry point cites the
 * reference interface it replaces.
 *
 * Conventions
 *   - plain C types only (pointers, sizes, PODs); no C++/torch types cross this boundary;
 *   - every function returns 0 on success and a negative rtc_status on failure; the
 *     message is available from rtc_last_error() (thread-local).  Nothing here calls
 *     exit() (the reference's gpuErrchk does, pch.h:45-53) and nothing throws;
 *   - one context = one GPU = one caller thread (the reference is single-threaded on
 *     this path); multi-GPU = one context (and normally one process) per GPU,
 *     each tracing a row band (rtc_trace_band), gathered by the caller;
 *   - there is NO CPU fallback: with no usable CUDA device rtc_create fails with
 *     RTC_ERR_CUDA.
 *   - "x" is the console width INCLUDING the newline column, exactly as in the reference:
 *     traced cells per row = x-1 (RayTracing.cu:491), rays per frame = (x-1)*y.
 */
#ifndef RTC_H_
#define RTC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RTC_API __attribute__((visibility("default")))
#else
#define RTC_API
#endif

typedef struct rtc_ctx rtc_ctx;

typedef enum rtc_status {
    RTC_OK = 0,
    RTC_ERR_INVALID = -1, /* bad argument / bad state                                  */
    RTC_ERR_CUDA = -2,    /* a CUDA runtime call failed (message has the CUDA string)   */
    RTC_ERR_NOMEM = -3,   /* host or device allocation failed                           */
    RTC_ERR_CAPACITY = -4 /* scene/output capacity exceeded                             */
} rtc_status;

/* == RenderingMode, same order (reference RayTracingManager.h:21). */
typedef enum rtc_mode {
    RTC_BIT_ASCII = 0,   /* ESC[38;5;Nm<ch>   12-byte cells, xterm-256 index           */
    RTC_BIT_PIXEL = 1,   /* ESC[48;5;Nm<sp>                                            */
    RTC_RGB_ASCII = 2,   /* ESC[38;2;R;G;Bm<ch> 20-byte cells                          */
    RTC_RGB_PIXEL = 3,   /* ESC[48;2;R;G;Bm<sp>                                        */
    RTC_RGB_NORMALS = 4, /* as RGB_PIXEL with colour = (uint8)(normal*255)             */
    RTC_SDL = 5          /* reference stub: traces, writes nothing (RayTracing.cu:755) */
} rtc_mode;

/* == ObjectType (reference Object3D.h:14). */
enum { RTC_OBJ_NONE = 0, RTC_OBJ_PLANE = 1, RTC_OBJ_SPHERE = 2 };

/* Flags for rtc_render / rtc_trace_band / rtc_update_objects. */
enum {
    RTC_FLAG_NONE = 0u,
    /* Extension (NOT in the reference, which casts no shadow rays -- SURVEY F1): one shadow ray
     * per shaded pixel, cast from the light (1,50,0) toward the shaded point lifted 1e-3 along
     * its normal; a point with an object strictly nearer to the light keeps only the ambient
     * term.  Off by default; all parity claims are made with it off.                      */
    RTC_FLAG_SHADOWS = 1u << 0,
    /* rtc_update_objects: mirror the reference's launch bug -- its UpdateObjects launch
     * uses block = count threads, which CUDA rejects when count > 1024, so no object moves
     * (RayTracingManager.cu:89-107).  Without this flag all objects are updated.        */
    RTC_FLAG_UPDATE_REF_LAUNCH_LIMIT = 1u << 1,
    /* Per-tile sphere culling (the reference's author planned a Culling kernel and never wrote it:
     * RayTracingManager.cu:46-51, :100-117).  Spheres are grouped by 4 along a Morton curve; per 16x16-pixel warp
     * tile, groups whose bounding cone misses the tile's ray cone are skipped.  Results are identical bit for bit;
     * only the number of ray-sphere tests executed changes (rtc_timings.sphere_tests), which is why the FP32 roofline
     * is quoted with this flag off.                                                                               */
    RTC_FLAG_CULL = 1u << 2,
    /* Parity hook: also keep the hit records (distance, object index) of every ray for rtc_frame_hits.  Without it
     * the ray kernel shades in its tile epilogue and the records never leave the SM (with RTC_FLAG_SHADOWS they are
     * kept anyway: the shadow pass reads them).                                                                    */
    RTC_FLAG_KEEP_HITS = 1u << 3,
    /* RGB_NORMALS only.  The reference converts normal * 255 -- negative for half of the normals -- with a plain
     * (uint8_t) cast (RayTracing.cu:669-671), which is undefined behaviour: its CUDA build saturates (cvt.rzi.u8.f32:
     * negative -> 0), its sources compiled for x86 wrap (cvttss2si, low byte).  The default follows the x86 build, the
     * one every parity fixture of this repository is pinned to; this flag selects the CUDA platform's conversion
     * (cross-checked against the reference's own kernels on a B200: tests/test_ref_cuda_crosscheck.py).           */
    RTC_FLAG_NORMALS_SATURATE = 1u << 4,
    /* Packet filter: the ray kernel's conservative ray-sphere filter runs once per PACKET -- the 8 (4) rays of one thread:
     * one column, consecutive rows -- on the packet's first and last ray instead of on every ray; packets it flags go
     * through the per-ray filter and the exact path as before.  Accepted hits, hence every output byte, are identical
     * (tests/test_gpu_parity.py::test_packet_filter_is_invisible); 14 instead of 50 packed operations per 4 spheres x 8
     * rays.  Like RTC_FLAG_CULL it changes what "a test" costs, so the brute-force roofline is quoted without it; the
     * facade turns both on.  Primary rays under a camera matrix only (any other matrix: ignored).                  */
    RTC_FLAG_PACKET = 1u << 5
};

/* == RayTracingCPUToGPUData (reference RayTracingManager.h:9-19) without the vptrs.
 * inv_view is row-major: inv_view[4*r + c] = Matrix::row{r+1}.{x,y,z,w}[c].            */
typedef struct rtc_params {
    float inv_view[16];
    float cam_pos[3];
    uint32_t x;       /* console width incl. newline column  (PrintMachine::GetWidth)  */
    uint32_t y;       /* console height                      (PrintMachine::GetHeight) */
    float element1;   /* projection [0][0] (Engine3D.cpp:94)                           */
    float element2;   /* projection [1][1] (Engine3D.cpp:95)                           */
    float cam_far;    /* far-plane distance (Engine3D.cpp:96)                          */
} rtc_params;

/* One scene object, 64 bytes.  The union of the reference's Sphere (Sphere.h:5-25) and
 * Plane (Plane.h:5-38) state; colour is float 0..255 per channel as in the reference.  */
typedef struct rtc_object {
    int32_t type;      /* RTC_OBJ_SPHERE / RTC_OBJ_PLANE                                 */
    float center[3];   /* Object3D::m_center                                             */
    float color[3];    /* Object3D::m_color                                              */
    float radius;      /* Sphere::m_radius                                               */
    float normal[3];   /* Plane::m_normal -- stored ALREADY normalised (Plane.cu:9)       */
    float width;       /* Plane::m_width  (extent along world X)                          */
    float height;      /* Plane::m_height (extent along world Z)                          */
    float speed;       /* Sphere::speed                                                  */
    int32_t mover;     /* Sphere::mover (+1/-1)                                          */
    int32_t reserved_;
} rtc_object;

/* The light and material constants of the reference's shading call site (RayTracing.cu:143-152, :69, :77), which the
 * reference hard-codes; rtc_set_light makes them per-context state.  The defaults are the reference's values.  The
 * shininess stays 32 (RayTracing.cu:151).                                                                           */
typedef struct rtc_light {
    float pos[3];          /* (1, 50, 0)       lightPos       RayTracing.cu:146; also the origin of the shadow rays */
    float diffuse_color;   /* 1                diffuseColor   :147                                                  */
    float diffuse_power;   /* 2000             diffusePower   :147                                                  */
    float specular_color;  /* 1                specColor      :148                                                  */
    float specular_power;  /* 3000             specPower      :148                                                  */
    float ambient[3];      /* (0.2, 0.2, 0.2)  ambient term   :77                                                   */
    float object_specular; /* 1                the object's specular colour :78                                     */
} rtc_light;

/* Per-stage device timings of the last rtc_render on this context (CUDA events). */
typedef struct rtc_timings {
    float prep_ms;    /* scene upload wait (the hoist runs inside kernel 1) */
    float trace_ms;   /* kernel 1: ray generation + nearest hit            */
    float shade_ms;   /* kernel 2: shade + quantise                        */
    float encode_ms;  /* kernel 3: ANSI encode (scan + scatter)            */
    float total_ms;   /* first launch to last kernel end                   */
    uint32_t launches;/* kernels launched by the last rtc_render           */
    uint32_t reserved_;
    uint64_t sphere_tests; /* packed ray-sphere tests executed (rays x spheres without RTC_FLAG_CULL, tile-granular) */
} rtc_timings;

/* ---- context ------------------------------------------------------------------------ */
/* Replaces RayTracingManager::RayTracingManager (RayTracingManager.cu:53-67) + the device
 * side of Scene3D::Init (Scene3D.cpp:7-26).  `device` is a CUDA ordinal.                */
RTC_API int rtc_create(rtc_ctx** out, int device);
/* Replaces ~RayTracingManager (RayTracingManager.cu:69-74) + Scene3D::CleanUp (:94-99). */
RTC_API void rtc_destroy(rtc_ctx* ctx);
RTC_API const char* rtc_last_error(void);
RTC_API const char* rtc_version(void);
/* Run all work of this context on an externally owned cudaStream_t (e.g. a torch.cuda.Stream),
 * or pass NULL to go back to the context's own non-blocking stream.  (NULL never means the
 * legacy default stream: hand over an explicit stream.)                                   */
RTC_API int rtc_set_stream(rtc_ctx* ctx, void* cuda_stream);
RTC_API int rtc_device_info(rtc_ctx* ctx, int* sm_count, int* clock_khz, size_t* smem_optin);
/* Wait for everything enqueued on the context's stream (the cudaDeviceSynchronize of the reference's Update,
 * RayTracingManager.cu:126, restricted to this context).                                                              */
RTC_API int rtc_synchronize(rtc_ctx* ctx);

/* Light / material block used by the shading stage and the shadow pass of every later frame; NULL restores the
 * reference's constants (RayTracing.cu:143-152).                                                                    */
RTC_API int rtc_set_light(rtc_ctx* ctx, const rtc_light* light);

/* ---- sink geometry: PrintMachine::Start / ChangeSize (PrintMachine.cpp:108-152,:216) - */
RTC_API int rtc_resize(rtc_ctx* ctx, uint32_t x, uint32_t y);

/* ---- scene: Scene3D (Scene3D.cpp:36-105) ---------------------------------------------- */
RTC_API int rtc_scene_clear(rtc_ctx* ctx);
/* Scene3D::CreateSphere (Scene3D.cpp:36-60).  speed/mover: the reference draws speed from
 * host rand() (Sphere.cu:11-12) and starts mover at -1; here they are explicit.          */
RTC_API int rtc_scene_add_sphere(rtc_ctx* ctx, const float center[3], float radius,
                                 const float rgb[3], float speed, int mover);
/* Scene3D::CreatePlane (Scene3D.cpp:62-86); `normal` is normalised here like Plane.cu:9. */
RTC_API int rtc_scene_add_plane(rtc_ctx* ctx, const float center[3], const float normal[3],
                                const float rgb[3], float width, float height);
/* Bulk replace (objects taken verbatim, plane normals must already be normalised).       */
RTC_API int rtc_scene_set_objects(rtc_ctx* ctx, const rtc_object* objs, uint32_t n);
/* Scene3D::GetObjects (Scene3D.cpp:102-105) -- host copy of the current object state
 * (after any rtc_update_objects).                                                        */
RTC_API int rtc_scene_get_objects(rtc_ctx* ctx, rtc_object* out, uint32_t cap, uint32_t* n);
RTC_API int rtc_scene_count(rtc_ctx* ctx, uint32_t* n);
/* UpdateObjects kernel / Sphere::Update (RayTracingManager.cu:10-44, Sphere.cu:15-23).   */
RTC_API int rtc_update_objects(rtc_ctx* ctx, double dt, uint32_t flags);

/* ---- frame: RayTracingManager::Update (RayTracingManager.cu:76-154) -------------------- */
/* Asynchronous: enqueue trace (+ shade/quantise) + ANSI encode (3 launches) for the whole frame
 * on the context's stream (replaces :83-127 and the host minimiser :146).                */
RTC_API int rtc_render(rtc_ctx* ctx, const rtc_params* params, rtc_mode mode, uint32_t flags);
/* Synchronise and return the minimised ANSI stream in a pinned host buffer owned by the
 * context (valid until the next rtc_render).  These are exactly the `size` bytes the
 * reference passes to PrintMachine::SetDataInBackBuffer (RayTracingManager.cu:150).      */
RTC_API int rtc_frame_ansi(rtc_ctx* ctx, const char** host_ptr, size_t* n_bytes);
/* Headless: the stream stays in HBM.  Synchronises only to learn n_bytes.                */
RTC_API int rtc_frame_ansi_device(rtc_ctx* ctx, const char** dev_ptr, size_t* n_bytes);
/* Parity hooks: quantised colour plane ((x-1)*y*bpp bytes, bpp = 3 for the RGB modes, 1 =
 * xterm-256 index for the 8-bit modes), glyph plane ((x-1)*y bytes, ' ' on a miss; NULL in
 * the PIXEL/NORMALS modes) and hit records (distance; object index or -1) of the last
 * rtc_render, copied to host buffers owned by the context.  rtc_frame_hits needs a frame
 * rendered with RTC_FLAG_KEEP_HITS (or RTC_FLAG_SHADOWS).                                 */
RTC_API int rtc_frame_color(rtc_ctx* ctx, const uint8_t** host_color, uint32_t* bpp,
                            const uint8_t** host_glyph);
RTC_API int rtc_frame_hits(rtc_ctx* ctx, const float** host_dist, const int32_t** host_index);
/* Parity hook: the shade kernel's xterm-256 quantiser (== ansi256_from_rgb, ANSIRGB.h:141-189) evaluated over the whole
 * RGB cube into 2^24 bytes of DEVICE memory, dev_out[(r << 16) | (g << 8) | b]; asynchronous on the context's stream. */
RTC_API int rtc_debug_ansi256_cube(rtc_ctx* ctx, uint8_t* dev_out);
/* RayTracingManager::Update as one synchronous call: optional physics step (dt != 0),
 * render, copy the stream to host.                                                        */
RTC_API int rtc_update(rtc_ctx* ctx, const rtc_params* params, rtc_mode mode, double dt,
                       uint32_t flags, const char** host_ptr, size_t* n_bytes);
/* The same frame driver, pipelined two deep (the reference's own sink is asynchronous too: Update hands the frame
 * to a print thread, PrintMachine.cpp:178-192, :257-306).  rtc_submit == rtc_update without the wait: physics step,
 * render, and the stream length on its way to the host.  rtc_collect returns the OLDEST submitted frame: it waits
 * for that frame only and copies its stream to pinned host memory on a separate copy stream, so the copy overlaps
 * the kernels of the frame submitted after it.  At most two frames in flight; the returned buffer stays valid
 * until the second rtc_submit after this call.
 *     rtc_submit(f0); for (k = 0; ; ++k) { rtc_submit(f[k+1]); rtc_collect(&ptr, &n); sink(ptr, n); }            */
RTC_API int rtc_submit(rtc_ctx* ctx, const rtc_params* params, rtc_mode mode, double dt, uint32_t flags);
RTC_API int rtc_collect(rtc_ctx* ctx, const char** host_ptr, size_t* n_bytes);
RTC_API int rtc_last_timings(rtc_ctx* ctx, rtc_timings* out);

/* ---- stage-level entry points on caller-owned DEVICE buffers ---------------------------
 * (used for multi-GPU row bands and for the encode-only workload; all asynchronous on the
 * context's stream).                                                                      */
/* RayTracing::RayTrace (RayTracing.h:31-38, RayTracing.cu:797-867) restricted to rows
 * [row0,row1): writes (row1-row0)*(x-1)*bpp colour bytes to dev_color and, in the ASCII
 * modes, (row1-row0)*(x-1) glyph bytes to dev_glyph (may be NULL otherwise).  The pointers
 * may be peer-mapped memory of another GPU.                                               */
RTC_API int rtc_trace_band(rtc_ctx* ctx, const rtc_params* params, rtc_mode mode,
                           uint32_t flags, uint32_t row0, uint32_t row1,
                           uint8_t* dev_color, uint8_t* dev_glyph);
/* RayTracing::RayTrace as the reference's INNER seam delivers it (RayTracing.h:31-38, RayTracing.cu:797-867): the RAW
 * cell buffer, rtc_raw_size(x, y) = 20*x*y bytes in DEVICE memory (4-byte aligned), one 20-byte (8-bit modes: 12-byte)
 * cell per console position at (row*x + col)*SIZE, the newline column and everything past SIZE*x*y zero -- exactly what
 * the reference's kernels leave in m_deviceResultArray after its per-frame memset (RayTracingManager.cu:161-165).  For
 * callers that keep the reference's own host-side MinimizeRGB; the product path (rtc_render / rtc_update) never builds
 * this buffer.  Asynchronous on the context's stream.                                                               */
RTC_API int rtc_trace_raw(rtc_ctx* ctx, const rtc_params* params, rtc_mode mode, uint32_t flags, char* dev_result);
RTC_API size_t rtc_raw_size(uint32_t x, uint32_t y);
/* MinimizeRGB / Minimize8bit (RayTracingManager.cu:251-319 / :181-249) on the device:
 * colour plane (+glyph plane) of an x-by-y frame -> minimised stream in dev_out (capacity
 * cap bytes; rtc_encode_capacity gives the worst case).  *dev_total (8-byte device or
 * mapped-host location) receives the stream length.                                       */
RTC_API int rtc_encode(rtc_ctx* ctx, const uint8_t* dev_color, const uint8_t* dev_glyph,
                       uint32_t x, uint32_t y, rtc_mode mode, char* dev_out, size_t cap,
                       unsigned long long* dev_total);
/* The same for a ROW BAND of a frame, so that every GPU can encode its own band and the per-band streams simply
 * concatenate to the frame's stream: `rows` rows starting at dev_color.  continues = 0: the band starts the frame
 * (its first cell always emits, as in rtc_encode).  continues != 0: the band continues a frame -- the colour key of
 * the cell stored immediately BEFORE dev_color (dev_color - bpp; the last cell of the previous row, which the caller
 * traces as one extra context row: rtc_trace_band(row0 - 1, ...)) decides whether the band's first cell emits, exactly
 * as MinimizeRGB carries latestColor across rows (RayTracingManager.cu:262-301).                                   */
RTC_API int rtc_encode_band(rtc_ctx* ctx, const uint8_t* dev_color, const uint8_t* dev_glyph, uint32_t x,
                            uint32_t rows, rtc_mode mode, int continues, char* dev_out, size_t cap,
                            unsigned long long* dev_total);
RTC_API size_t rtc_encode_capacity(uint32_t x, uint32_t y, rtc_mode mode);
RTC_API uint32_t rtc_mode_bpp(rtc_mode mode);
RTC_API uint32_t rtc_mode_has_glyph(rtc_mode mode);

/* ---- multi-GPU frame driver: RayTracingManager::Update across the GPUs of one box ------------------------------------
 * The reference's frame driver (RayTracingManager.h:31-35, RayTracingManager.cu:76-154) can only use one GPU.  rtc_mgpu
 * splits the frame into contiguous row bands, one per device, in ONE process: a context, a stream and a worker thread per
 * device, frames pipelined three deep.  The result is byte-identical to a single-GPU frame (MinimizeRGB's colour
 * carry-over across band seams included, RayTracingManager.cu:262-301).
 *   RTC_GATHER_HOST  every device encodes its own band and copies its piece of the stream over its own PCIe link into
 *                    one pinned host frame, at the offset given by the lengths of the devices before it;
 *   RTC_GATHER_P2P   every device's ray kernel stores its quantised band straight into device 0's frame planes over
 *                    NVLink (peer stores), device 0 encodes the whole frame and copies the stream out.
 * device_ids == NULL means 0..n_gpus-1; an id may appear more than once (several bands on one GPU: tests).            */
typedef struct rtc_mgpu rtc_mgpu;
enum { RTC_GATHER_HOST = 0, RTC_GATHER_P2P = 1 };
RTC_API int rtc_mgpu_create(rtc_mgpu** out, int n_gpus, const int* device_ids, int gather);
RTC_API void rtc_mgpu_destroy(rtc_mgpu* m);
RTC_API int rtc_mgpu_count(rtc_mgpu* m);
/* Borrowed per-device context (device info, timings); do not destroy it.                                             */
RTC_API int rtc_mgpu_context(rtc_mgpu* m, int i, rtc_ctx** out);
/* Scene3D (Scene3D.cpp:36-105), replicated on every device.  The mutators may be called with frames in flight: they
 * are queued and take effect, in order, with the next submitted frame.  get_objects needs an idle pipeline.           */
RTC_API int rtc_mgpu_scene_clear(rtc_mgpu* m);
RTC_API int rtc_mgpu_scene_add_sphere(rtc_mgpu* m, const float center[3], float radius, const float rgb[3], float speed, int mover);
RTC_API int rtc_mgpu_scene_add_plane(rtc_mgpu* m, const float center[3], const float normal[3], const float rgb[3],
                                     float width, float height);
RTC_API int rtc_mgpu_scene_set_objects(rtc_mgpu* m, const rtc_object* objs, uint32_t n);
RTC_API int rtc_mgpu_scene_get_objects(rtc_mgpu* m, rtc_object* out, uint32_t cap, uint32_t* n);
RTC_API int rtc_mgpu_set_light(rtc_mgpu* m, const rtc_light* light);
/* RayTracingManager::Update, pipelined like rtc_submit / rtc_collect (at most three frames in flight): submit enqueues
 * physics step + band trace + shade (+ band encode) on every device and returns at once; collect returns the OLDEST
 * submitted frame's stream in pinned host memory, valid until the third rtc_mgpu_submit after the one that made it.   */
RTC_API int rtc_mgpu_submit(rtc_mgpu* m, const rtc_params* params, rtc_mode mode, double dt, uint32_t flags);
RTC_API int rtc_mgpu_collect(rtc_mgpu* m, const char** host_ptr, size_t* n_bytes);
/* submit + collect as one synchronous call (== RayTracingManager::Update).                                            */
RTC_API int rtc_mgpu_update(rtc_mgpu* m, const rtc_params* params, rtc_mode mode, double dt, uint32_t flags,
                            const char** host_ptr, size_t* n_bytes);
/* Of the last collected frame: per-device time of its kernels (CUDA events on each device's stream), the band
 * boundaries rows[0..n_gpus], and (P2P) device 0's encode time.  Any pointer may be NULL.                             */
RTC_API int rtc_mgpu_last_frame(rtc_mgpu* m, float* device_ms, uint32_t* rows, float* encode_ms);
/* Host-side accounting since the previous call (idle pipeline only): out[3*g + {0,1,2}] = device g's average microseconds
 * per frame spent enqueueing (scene staging + launches), waiting for the stream lengths of the devices before it, and in
 * its D2H copy.                                                                                                       */
RTC_API int rtc_mgpu_host_stats(rtc_mgpu* m, float* out);
/* Debug: host-clock timeline of the last 64 frames (layout in rtc_mgpu.cu); out holds (n_gpus * 6 + 2) * 64 doubles.    */
RTC_API int rtc_mgpu_debug_trace(rtc_mgpu* m, double* out);
/* Explicit band boundaries rows[0..n_gpus] for frames of height y (NULL: automatic -- equal bands; in the P2P gather
 * device 0's band shrinks by the measured cost of the encoder).                                                       */
RTC_API int rtc_mgpu_set_bands(rtc_mgpu* m, uint32_t y, const uint32_t* rows);
/* Measurement helper: the next submit first overwrites 256 MiB on every device (evicts the 126 MB L2), untimed.       */
RTC_API int rtc_mgpu_flush_l2(rtc_mgpu* m);
/* The band planner itself (host only, no GPU): see rtc_mgpu.cu.                                                       */
RTC_API int rtc_plan_bands(uint32_t y, int n, uint32_t align, double deficit_rows, uint32_t wave_units, uint32_t* rows_out);

/* ---- host-side helpers that need no GPU (reference host code on the path) -------------- */
/* Camera3D::Init + Update + GetInverseVMatrix + Engine3D::Render's parameter block
 * (Camera3D.cpp:8-48, :51-98, :207-376; Engine3D.cpp:88-97).  pixel_aspect is the
 * reference's hard-coded 0.01f (Camera3D.cpp:17); pass 0 for that default.               */
RTC_API int rtc_camera_params(uint32_t x, uint32_t y, const float pos[3], const float rot[3],
                              float pixel_aspect, rtc_params* out);

/* ---- microbenchmark: FP32 pipe peak measured on this GPU (roofline denominator) -------- */
/* variant 0: scalar FFMA chain, 1: packed FFMA2 chain.  Returns achieved TFLOP/s.          */
RTC_API int rtc_fp32_peak(rtc_ctx* ctx, int variant, int iters, float* tflops, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* RTC_H_ */
