/* ==========================================================================================
 * essa_b200.h -- C ABI of the B200-native World::update hot path.
 *
 * The reference (essa-software/essa) has no plugin / FFI interface: its boundary for this path
 * is the C++ method `void World::update(int steps)` (src/World.hpp:29, src/World.cpp:86-139)
 * observed through the Object accessors (src/Object.hpp:72-99).  This header is what the
 * replacement World::update binds instead of running the loop itself:
 *
 *     reference                                            this ABI
 *     ---------------------------------------------------  ---------------------------------
 *     World::m_object_list state (src/World.hpp:77,        essa_upload / essa_download
 *       src/Object.hpp:152-191)                            (SoA, plain double* / int*)
 *     World::update(int)           src/World.cpp:86-139    essa_step
 *     World::set_forces()          src/World.cpp:42-57     essa_set_forces
 *     World::m_date / deleted()    src/World.cpp:59-84,    essa_set_date / essa_get_date
 *                                  src/Object.cpp:60-62
 *     Object::m_history            src/History.cpp:6-59    essa_history_read / _write
 *     Object::m_trail              src/Trail.cpp:19-110    essa_trail_read / _write
 *     Trail::draw's vertex data    src/Trail.cpp:90-103    essa_trail_device_view (zero-copy), essa_snapshot_*
 *     clone_for_forward_simulation src/World.cpp:230-238,  essa_ensemble_init / _step / _read
 *       + update(ticks)            src/essagui/EssaCreateObject.cpp:166-189
 *     (none: new diagnostic required by the north star)    essa_energy
 *
 * Conventions: plain C types; every array is caller-owned host memory unless the name says
 * `_device`; int return code (0 = ok, <0 = error, see essa_status) with a message available
 * from essa_last_error(); no exceptions cross the boundary; one context per host thread;
 * every call is synchronous on return unless it ends in `_async`.
 *
 * There is NO CPU fallback: without a CUDA device essa_ctx_create fails with ESSA_ERR_CUDA.
 * ========================================================================================== */
#ifndef ESSA_B200_H
#define ESSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESSA_ABI_VERSION 2

typedef struct essa_ctx essa_ctx;

typedef enum essa_status {
    ESSA_OK = 0,
    ESSA_ERR_ARG = -1,   /* bad argument (null pointer, steps == 0, N out of range ...) */
    ESSA_ERR_CUDA = -2,  /* CUDA runtime error, incl. "no device" */
    ESSA_ERR_NCCL = -3,  /* NCCL missing or failed */
    ESSA_ERR_STATE = -4, /* call not valid in the current state (nothing uploaded ...) */
    ESSA_ERR_NOMEM = -5
} essa_status;

/* Util::Constants::Gravity and ::AU as used by the reference (src/Object.cpp:38,
 * src/Trail.cpp:53); G is pinned by the reference's tests/accuracy.ods. */
#define ESSA_GRAVITY 6.67408e-11
#define ESSA_AU 149597870700.0
/* Util::SimulationTime::create(1990, 4, 20) (src/World.cpp:24-27) as seconds, from the traces. */
#define ESSA_START_DATE 640566000LL
/* History(1000, ...) in the Object constructor (src/Object.cpp:34). */
#define ESSA_HISTORY_SEGMENTS 1000
/* Largest world the resident single-CTA kernel takes (one thread per body). */
#define ESSA_RESIDENT_MAX 1024
/* Widest template world essa_ensemble_init takes (one copy must fit a 320-thread CTA). */
#define ESSA_ENSEMBLE_MAX 320

/* Which kernel essa_step runs. */
typedef enum essa_kernel {
    ESSA_KERNEL_AUTO = 0,     /* N <= ESSA_RESIDENT_MAX: resident; N <= 5120: grid (else tiled: recording, two_pass); N > 12288: paired (tracking: up to 2^21 bodies); else tiled */
    ESSA_KERNEL_TILED = 1,    /* K1: j-tiled FP64 force pass + fused kick/drift (any N) */
    ESSA_KERNEL_RESIDENT = 2, /* K3: one CTA keeps the world in shared memory for all steps */
    ESSA_KERNEL_PAIRED = 3,   /* K1p: pair-shared force pass (each unordered pair once, both bodies updated) */
    ESSA_KERNEL_GRID = 4      /* K1g: 1,025 .. 5,120 bodies, all steps of a call in one cooperative launch over all SMs */
} essa_kernel;

/* SoA view of the per-Object state (src/Object.hpp:152-191).  A null member is skipped.
 * `most` is the list index of Object::m_most_attracting_object, -1 for nullptr. */
typedef struct essa_bodies {
    double *x, *y, *z;       /* m_pos */
    double *vx, *vy, *vz;    /* m_vel */
    double *ax, *ay, *az;    /* m_attraction_factor (output of set_forces; on upload: optional value kept by masked bodies) */
    double *gf;              /* m_gravity_factor = mass * G */
    int64_t *creation_date;  /* m_creation_date, seconds */
    int64_t *deletion_date;  /* m_deletion_date, seconds */
    uint8_t *deleted;        /* m_deleted */
    int32_t *most;           /* m_most_attracting_object */
    double *max_attraction;  /* m_max_attraction (same remark) */
    double *ap, *pe, *ap_vel, *pe_vel; /* m_ap, m_pe, m_ap_vel, m_pe_vel */
} essa_bodies;

/* Options of one essa_step call.  Zero-initialise, then set what differs. */
typedef struct essa_step_opts {
    int dt_seconds;            /* World::m_simulation_seconds_per_tick (an int, src/World.hpp:78) */
    int forward_simulated;     /* World::m_is_forward_simulated: date frozen, no History, trails not offset */
    int offset_trails;         /* SimulationView::offset_trails() (default true, SimulationView.hpp:182) */
    int record;                /* run Object::update (History + Trail + ap/pe) for the tracked bodies */
    int track_most_attracting; /* evaluate Object.cpp:84-94 (forced on by `record`) */
    int exact_order;           /* reference operation order, IEEE div/sqrt, no FMA: bit-identical to the
                                  reference restatement; resident kernel only */
    int two_pass;              /* evaluate both set_forces() calls of every step instead of reusing the
                                  acceleration left by the previous step (same results, 2x the work) */
    int kernel;                /* essa_kernel */
    double eps2;               /* Plummer softening^2; 0 = reference semantics */
} essa_step_opts;

/* ---- context ----------------------------------------------------------------------------- */
int essa_abi_version(void);
/* device: CUDA ordinal.  capacity: bodies to reserve for (grows on demand). */
int essa_ctx_create(int device, int capacity, essa_ctx** out);
void essa_ctx_destroy(essa_ctx* ctx);
/* Message of the last failure on this context (or of the last failed create when ctx is null). */
const char* essa_last_error(const essa_ctx* ctx);
/* Kernels launched by this context so far / force passes / directed interactions evaluated. */
int64_t essa_launch_count(const essa_ctx* ctx);
int64_t essa_pass_count(const essa_ctx* ctx);
double essa_interaction_count(const essa_ctx* ctx);
/* Device-side milliseconds of the last essa_step / essa_ensemble_run (CUDA events on its stream). */
double essa_last_step_ms(const essa_ctx* ctx);
/* Milliseconds spent in force-pass kernels during the last essa_step (events around each launch). */
double essa_last_force_ms(const essa_ctx* ctx);
/* Force passes the last essa_step / essa_ensemble_step evaluated. */
int essa_last_force_passes(const essa_ctx* ctx);
/* out = {step ms, force-pass ms, ms before the first force pass, ms after the last} of the last essa_step. */
int essa_last_step_breakdown(const essa_ctx* ctx, double out[4]);
/* Phases of the last pair-shared force pass that was timed by itself (a single-step essa_step / essa_set_forces), in ms:
 * out = {before the kernel: position gather wait, threshold exchange, mass sort; the pair kernel; the exchange behind it
 * (sharded worlds: the barrier after the peer stores, or the NCCL reduce-scatter fallback); k_pair_finish}.
 * New diagnostic (no reference counterpart): the per-kernel timeline of a sharded pass for bench.py. */
int essa_last_pass_phases(const essa_ctx* ctx, double out[4]);
int essa_sync(essa_ctx* ctx);

/* ---- world state --------------------------------------------------------------------------- */
/* Replace the world by `n` bodies (list order = index order).  x..vz and gf are required; dates
 * default to "always active", most to -1, ap to 0, pe to DBL_MAX.  Recording state is kept. */
int essa_upload(essa_ctx* ctx, int n, const essa_bodies* host);
/* Copy current state out (members that are null are skipped). */
int essa_download(essa_ctx* ctx, essa_bodies* host);
int essa_set_date(essa_ctx* ctx, int64_t date);
int64_t essa_get_date(const essa_ctx* ctx);
int essa_body_count(const essa_ctx* ctx);

/* World::update(steps): |steps| KDK iterations, sign = direction (src/World.cpp:86-139). */
int essa_step(essa_ctx* ctx, int steps, const essa_step_opts* opts);
/* The same, but a call that one resident kernel serves returns after the launch: the next synchronising call on the
 * context (essa_download, essa_sync, essa_last_step_ms, the next step ...) waits for it and reports its errors.
 * What the GUI's tick needs: update(steps) immediately followed by a read-back pays for one synchronisation. */
int essa_step_async(essa_ctx* ctx, int steps, const essa_step_opts* opts);
/* World::set_forces() alone (src/World.cpp:42-57): refreshes acc / most / max_attraction. */
int essa_set_forces(essa_ctx* ctx, const essa_step_opts* opts);

/* Persistent resident kernel.  The reference's caller, SimulationView::update (src/SimulationView.cpp:374-397), runs
 * World::update(iterations) once per GUI tick (10 iterations by default, src/essagui/EssaSettings.cpp:34-47) and reads the
 * objects back.  For worlds of up to 32 bodies in fast arithmetic the kernel that served one essa_step therefore stays
 * resident on its SMs: the next essa_step / essa_step_async with the same options is handed to it through a mailbox in
 * mapped pinned host memory (no launch, no stream synchronisation) and the following essa_download reads the state the
 * kernel published there.  Any other call on the context retires the kernel first (it stores its state to device memory
 * as a one-shot launch does), and a kernel nobody rings for `idle_timeout_us` retires by itself -- so the mode is
 * invisible apart from its speed.  enable: 0 = one launch per call; idle_timeout_us: 0 keeps the current value
 * (default 1000; environment: ESSA_PERSIST=0 / ESSA_PERSIST_IDLE_US). */
int essa_persist_configure(essa_ctx* ctx, int enable, int idle_timeout_us);
/* out = {kernel resident now (0/1), ticks served through the mailbox, persistent kernels launched,
 *        kernels found retired by their idle timeout when the next tick came,
 *        host nanoseconds spent between a doorbell (or launch) and its result, ... between a result and the next doorbell}. */
int essa_persist_stats(const essa_ctx* ctx, int64_t out[6]);

/* ---- recording: Object::m_history / m_trail (src/History.cpp, src/Trail.cpp) ---------------- */
/* Track bodies [0, n_tracked): allocate History rings (ESSA_HISTORY_SEGMENTS entries) and Trail
 * vertex rings of trail_capacity[k] vertices (already clamped as in src/Object.cpp:33 and
 * src/Trail.cpp:19-34) and run the constructor-time pushes from the uploaded state.  Bodies
 * [0, keep) keep their existing rings (an object was appended), the rest are (re)created. */
int essa_record_configure(essa_ctx* ctx, int n_tracked, const int32_t* trail_capacity, int keep);
int essa_tracked_count(const essa_ctx* ctx);
/* Object::reset_history() / Trail::reset() for one body (src/Object.cpp:268-271). */
int essa_record_reset(essa_ctx* ctx, int body, int history, int trail);
/* History entries oldest first, 6 doubles each (px py pz vx vy vz).  Returns the entry count
 * (<= ESSA_HISTORY_SEGMENTS) or <0; `out` may be null to query the count. */
int essa_history_read(essa_ctx* ctx, int body, double* out);
/* set_first != 0 also makes entries[0..6) the History's m_first_entry (what History::reset() restores). */
int essa_history_write(essa_ctx* ctx, int body, const double* entries, int count, int set_first);
typedef struct essa_trail_meta {
    int32_t capacity;      /* m_vertexes.size() */
    int32_t append_offset; /* m_append_offset */
    int32_t length;        /* m_length */
    int32_t pad;
    double offset[3];      /* m_offset */
    double last_xy_angle;  /* m_last_xy_angle */
} essa_trail_meta;
/* verts: 3 * capacity floats (x y z per vertex, already divided by AU); may be null. */
int essa_trail_read(essa_ctx* ctx, int body, float* verts, essa_trail_meta* meta);
int essa_trail_write(essa_ctx* ctx, int body, const float* verts, const essa_trail_meta* meta);

/* ---- consumers of the recording: what src/Trail.cpp:90-103 (Trail::draw) and src/World.cpp:150-161 (World::draw) do next ------ */
/* Zero-copy view of a tracked body's Trail ring IN DEVICE MEMORY, for a renderer that draws from a CUDA-mapped vertex buffer
 * (cudaGraphicsGLRegisterBuffer / external memory) or copies device-to-device: `verts_device` = the ring's first vertex
 * (3 floats each, already divided by AU, `capacity` of them), plus the cursor Trail::draw needs to split the ring into its one
 * or two line strips (draw [1, append_offset) while length != capacity, else [0, append_offset) and [append_offset, length)) and
 * the offset its model matrix translates by.  The pointer stays valid until the next essa_record_configure / essa_ctx_destroy;
 * the contents change with every essa_step (retire the view's reader first: the call synchronises the context). */
typedef struct essa_trail_view {
    const float* verts_device;
    int32_t capacity, append_offset, length, pad;
    double offset[3];
} essa_trail_view;
int essa_trail_device_view(essa_ctx* ctx, int body, essa_trail_view* out);
/* Flat snapshot of the whole context -- bodies, date, History and Trail rings of the tracked bodies -- as one self-describing
 * little-endian blob (the reference has no on-disk or wire format for simulation state; this one is
 * "ESSASNAP" | u32 version | u32 n | u32 tracked | i64 date | per-field arrays, see capi.cu).  essa_snapshot_size returns the
 * bytes essa_snapshot_write needs; essa_snapshot_read restores a context from a blob (upload + recording state). */
int64_t essa_snapshot_size(essa_ctx* ctx);
int essa_snapshot_write(essa_ctx* ctx, void* blob, int64_t bytes);
int essa_snapshot_read(essa_ctx* ctx, const void* blob, int64_t bytes);

/* ---- closest approaches of a forward-simulated world (src/Object.cpp:405-417) ------------------ */
/* Allocate (enable != 0) or drop the N x N table; essa_step with forward_simulated set then keeps,
 * for every ordered pair, the smallest distance seen at a step's start and both positions there. */
int essa_closest_enable(essa_ctx* ctx, int enable);
/* out = {distance, this x y z, other x y z}; returns 1 if the pair was evaluated, 0 if not, <0 on error. */
int essa_closest_read(essa_ctx* ctx, int body, int other, double out[7]);

/* ---- diagnostics ----------------------------------------------------------------------------- */
/* Kinetic and potential energy (J) and total momentum (kg m/s) of the active bodies, FP64,
 * fixed reduction order.  m = gf / ESSA_GRAVITY. */
int essa_energy(essa_ctx* ctx, double* kinetic, double* potential, double momentum[3]);

/* ---- ensemble: many perturbed copies of the uploaded template -------------------------------- */
typedef struct essa_ensemble_opts {
    int copies;            /* number of copies held by THIS context (a rank's share) */
    int64_t first_copy;    /* global index of this context's copy 0 (keys the perturbation) */
    int dt_seconds;
    uint64_t seed;
    double rel_sigma;      /* every pos/vel component *= 1 + rel_sigma * n,  n ~ Irwin-Hall(12)-6 from
                              Philox4x32-10 keyed (seed; copy, body, component); global copy 0 unperturbed */
    int exact_order;       /* as essa_step_opts.exact_order */
    int closest_approaches;/* Object::update_closest_approaches (src/Object.cpp:405-417) for the designated body
                              `approach_to` (the candidate of src/essagui/EssaCreateObject.cpp:166-189) against EVERY other
                              body of its copy: minimum distance over the run and both positions there */
    int approach_to;
    int trail_every;       /* > 0 (with closest_approaches): also keep the designated body's position after every
                              trail_every-th step -- the trail the GUI draws for the candidate ... */
    int trail_capacity;    /* ... up to this many points per copy */
} essa_ensemble_opts;
/* Build the copies on the device from the uploaded template (positions, velocities, gf). */
int essa_ensemble_init(essa_ctx* ctx, const essa_ensemble_opts* opts);
/* Advance every copy by |steps| KDK iterations (forward-simulated semantics: no date, no History). */
int essa_ensemble_step(essa_ctx* ctx, int steps);
/* State of `count` copies starting at local copy `first`: arrays of count * N doubles, copy-major. */
int essa_ensemble_read(essa_ctx* ctx, int first, int count, double* x, double* y, double* z,
                       double* vx, double* vy, double* vz, double* min_distance);
/* Closest approaches of `count` copies: this_pos / other_pos are count * N * 3 doubles -- for body k of a copy, where body k
 * (this) and the designated body (other) were when their distance was smallest (Object::ClosestApproachEntry,
 * src/Object.hpp:184-188); the distances are essa_ensemble_read's min_distance.  Either pointer may be null. */
int essa_ensemble_read_approaches(essa_ctx* ctx, int first, int count, double* this_pos, double* other_pos);
/* The designated body's trail of `count` copies: xyz = count * trail_capacity * 3 doubles, lengths = count points filled. */
int essa_ensemble_read_trail(essa_ctx* ctx, int first, int count, double* xyz, int32_t* lengths);
/* The perturbation factor (1 + rel_sigma * n) for one (copy, body, component 0..5), on the host,
 * bit-identical to what the device applies: lets a checker rebuild any copy's initial state. */
double essa_ensemble_factor(uint64_t seed, double rel_sigma, int64_t copy, int body, int component);

/* ---- multi-GPU: targets sharded over ranks, positions all-gathered each force pass ----------- */
#define ESSA_NCCL_ID_BYTES 128
/* Rank 0 fills `id` (an ncclUniqueId); the caller broadcasts it (torch.distributed, MPI ...). */
int essa_nccl_unique_id(uint8_t id[ESSA_NCCL_ID_BYTES]);
/* Join the communicator.  After this, essa_upload takes the FULL world on every rank and
 * essa_step integrates bodies [rank*N/world, (rank+1)*N/world) with one all-gather of the
 * drifted positions per force pass.  essa_download is collective (every rank calls it with the same members set) and
 * returns the full world on every rank: positions are current everywhere, velocities, accelerations, most-attracting
 * links, max_attraction and ap / pe are gathered from their owners.  Recording (essa_step_opts.record) runs for the
 * tracked bodies a rank owns: History / Trail rings of body k are read on the rank whose essa_shard_range holds k. */
int essa_shard_attach(essa_ctx* ctx, int rank, int world, const uint8_t id[ESSA_NCCL_ID_BYTES]);
/* A sharded host program keeps only its own bodies current on the host.  essa_upload_shard: positions / velocities of bodies
 * [i0, i1) (essa_shard_range) from host arrays indexed by GLOBAL body index -- only that slice is read; the records of the
 * other ranks' bodies arrive over NVLink (collective); gravity factors, dates and masks stay as the last essa_upload left
 * them.  essa_download_shard: the requested members of bodies [i0, i1) only (same indexing, only the slice is written), no
 * collective.  Per tick a rank then moves 1 / world of the state over its PCIe link instead of all of it. */
int essa_upload_shard(essa_ctx* ctx, const essa_bodies* host);
int essa_download_shard(essa_ctx* ctx, essa_bodies* host);
/* The target range [*i0, *i1) of a rank (pure function; what every kernel of a sharded context uses). */
int essa_shard_range(int n, int rank, int world, int* i0, int* i1);
int essa_shard_rank(const essa_ctx* ctx);
int essa_shard_world(const essa_ctx* ctx);

/* ---- measurement helpers ----------------------------------------------------------------------- */
/* Register-resident independent-DFMA-chain kernel on every SM for about `ms` milliseconds:
 * returns the achieved FP64 rate in TFLOP/s (FMA = 2 flops) -- the roofline denominator that
 * MEASURED_PEAKS.json lacks -- or <0. */
double essa_measure_fp64_peak(essa_ctx* ctx, double ms);
/* Maximum relative error of the MUFU.RSQ64H seed and of the refined r^-3 over a log sweep. */
int essa_measure_rsqrt_error(essa_ctx* ctx, double* seed_rel_err, double* refined_rel_err);
/* Dependent-issue latencies in SM cycles: {DFMA, DADD, DMUL, MUFU.RSQ64H seed, LDS chase, shuffle of a double}, then
 * the single-warp FP64 issue interval with 4 and 8 independent chains (cycles per instruction):
 * what bounds a single small-N trajectory (one step is a dependent chain issued by one warp). */
int essa_measure_latency(essa_ctx* ctx, double out[8]);
/* Evict the L2 between timed iterations (writes a 256 MiB scratch buffer). */
int essa_flush_l2(essa_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* ESSA_B200_H */
