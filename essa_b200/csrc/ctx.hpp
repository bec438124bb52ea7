// Host-side context behind the C ABI (include/essa_b200.h).  Owns every device buffer of one world.
#pragma once

#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/essa_b200.h"

namespace essa {

constexpr int kTileJ = 256;       // j-records per shared-memory tile (8 KiB)
constexpr int kChunkQ = 32;       // K1's j-chunks are cut at multiples of this many records
constexpr int kStages = 4;        // bulk-copy pipeline depth
constexpr int kForceThreads = 256;
constexpr int kEnsembleMaxBodies = 320; // widest template one ensemble CTA takes
constexpr int kGridMaxBodies = 5120;     // K1g (force_grid.cu): every SM holds all positions in shared memory
constexpr int kEnsembleSoloMax = 16;    // templates up to this size run one thread per copy (ensemble.cu k_ensemble_solo)

// Arguments of the fused kick / drift epilogue (src/World.cpp:106-116,121-127).
struct KickArgs {
    double h;      // dt / 2
    double dt;     // (double)seconds_per_tick
    double mul;    // +1 / -1
    int second_kick; // v = vh + (a h) mul    -- finishes the step whose drift produced these positions
    int advance;     // vh = v + (a h) mul ; x' = x + (vh dt) mul  -> other position buffer
};

// Device pointers of the world (all arrays have `cap` elements).
struct WorldPtrs {
    double4* pos_cur;  // {x, y, z, active ? gf : 0}: the j-record the force pass streams
    double4* pos_next;
    double* v[3];      // velocity at the integer step (what Object::vel() shows)
    double* vh[3];     // velocity after the first half kick
    double* a[3];      // m_attraction_factor at pos_cur
    double const* gf;  // m_gravity_factor
    uint8_t const* active;
    int32_t* most;
    int32_t* old_most;
    double* maxattr;
};

struct ForceArgs {
    WorldPtrs w;
    int n;          // bodies in the world
    int i0, i1;     // targets owned by this context (whole world unless sharded)
    int j0, j1;     // sources streamed by this launch
    int S;          // j-chunks in this launch (gridDim.x = iblocks * S)
    int slot0;      // first partial slot this launch writes
    int slots_total; // slots per i-block that must be present before the epilogue runs
    int ipad;       // iblocks * 256 * R
    double* part;   // [slots_total][3][ipad]
    double* part_m; // [slots_total][ipad]   (argmax variant)
    int* part_j;    // [slots_total][ipad]
    unsigned* counters; // [iblocks]
    double eps2;
    int save_old;   // Object::before_update: old_most = most (first pass of a step)
    KickArgs k;
};

// ---- recording state of the tracked bodies (History + Trail + ap/pe) -----------------------
struct RecordPtrs {
    int tracked;
    double* hist;        // [tracked][ESSA_HISTORY_SEGMENTS][6]
    int32_t* hist_start; // ring start
    int32_t* hist_len;
    float* verts;        // concatenated vertex rings, 3 floats per vertex
    int64_t const* vert_off; // [tracked] first vertex of body k
    int32_t const* cap;  // [tracked]
    int32_t* append_off;
    int32_t* length;
    double* toff;        // [tracked][3]  Trail::m_offset
    double* last_angle;  // [tracked]
    double* ap;
    double* pe;
    double* ap_vel;
    double* pe_vel;
};

// ---- persistent resident kernel (resident_quad.cu K3q): host <-> device mailbox in mapped pinned memory ----
// The host rings the doorbell by writing `cmd`; the kernel answers by filling `ll` with self-validating words
// {tick number (high half), 32 bits of payload (low half)}: the result is complete when every word carries the tick's number.
constexpr int kPersistMaxBodies = 128; // K3q: up to 32 bodies, K3x: up to 128
// Two layouts of the published words (32 bits of payload each, doubles as (low, high) halves):
//  * worlds of up to 32 bodies (K3q, one publisher warp): field-major -- x y z vx vy vz ax ay az max_attraction ap pe ap_vel pe_vel
//    as 2 n words each, then n links, then two words with the tick's device nanoseconds;
//  * larger worlds (K3x, every warp publishes its own body): one 32-word record per body (256 bytes, so that a body's words leave
//    the GPU as one or two contiguous stores) -- words 0 .. 17 x .. az, 18 .. 27 max_attraction ap pe ap_vel pe_vel, 28 the link,
//    29 .. 31 unused -- then the two time words.
constexpr int kPersistFieldMajorMax = 32;
constexpr int kPersistRecWords = 32, kPersistRecUsed = 29;
constexpr int kPersistWords = kPersistRecWords * kPersistMaxBodies + 2;
struct PersistMailbox {
    unsigned long long cmd;   // host -> device: (seq << 32) | steps;  steps == 0: retire
    unsigned long long pad0[15];
    unsigned ack[4];          // device -> host: [1] 1 = retired on request, 2 = retired after the idle timeout
    unsigned pad1[28];
    unsigned long long ll[kPersistWords];
};

} // namespace essa

struct ncclComm;

struct essa_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_f0 = nullptr, ev_f1 = nullptr, ev_gather = nullptr, ev_ready = nullptr;

    int n = 0;
    int cap = 0;
    double4* pos[2] = { nullptr, nullptr };
    int cur = 0;
    double* vel = nullptr;  // 3*cap
    double* velh = nullptr; // 3*cap
    double* acc = nullptr;  // 3*cap
    double* gf = nullptr;
    int64_t* cdate = nullptr;
    int64_t* ddate = nullptr;
    uint8_t* delflag = nullptr;
    uint8_t* active = nullptr;
    int32_t* most = nullptr;
    int32_t* old_most = nullptr;
    double* maxattr = nullptr;
    double* appe = nullptr; // 4*cap: ap, pe, ap_vel, pe_vel (also used by untracked bodies' upload/download)
    bool acc_valid = false;     // acc[] holds set_forces() of pos[cur] under the current active mask
    bool acc_has_most = false;  // ... and most / max_attraction were evaluated with it
    bool acc_exact = false;     // ... in exact_order arithmetic
    double acc_eps2 = 0.0;      // ... with this softening
    bool gather_pending = false;
    // paired_plan's last answer (a pure function of n, the sharding and the SM count)
    mutable struct {
        int n = -1, world = 0, rank = 0, cap = 0, nb = 0, S = 0, ri = 0;
        int db = 0, batches = 1; // offsets per launch / launches per pass (slot budget, force_paired.cu)
        bool ok = false;
        size_t slot_doubles = 0, part_doubles = 0;
    } plan;
    bool gf_uniform = false; // every uploaded gravity factor is the same value ...
    double gf_uniform_value = 0.0; // ... this one
    bool timing_pending = false; // essa_step_async: ev_end recorded, elapsed time not read yet
    bool attr_pair[8] = { false, false, false, false, false, false, false, false };
    bool attr_cta = false; // opt-in shared-memory sizes set on this device
    bool attr_cluster_wide[3] = { false, false, false };
    bool attr_cluster16[5] = { false, false, false, false, false }; // non-portable cluster size allowed for K3x<PER>
    bool use_paired = false;    // this essa_step runs its force passes through force_paired.cu
    std::vector<cudaEvent_t> force_events;
    int last_force_passes = 0;
    int64_t date = ESSA_START_DATE;
    std::vector<int64_t> events; // sorted creation / deletion instants at which the active mask flips
    bool any_inactive_possible = false;

    // force-pass scratch
    double* part = nullptr;
    double* part_m = nullptr;
    int* part_j = nullptr;
    size_t part_doubles = 0;
    size_t part_slots_cap = 0;
    unsigned* counters = nullptr;
    int counters_cap = 0;

    // pair-shared force pass scratch (force_paired.cu)
    double* pair_slots = nullptr;
    size_t pair_slots_n = 0;
    double* pair_part = nullptr;
    size_t pair_part_n = 0;
    unsigned long long* pair_kslots = nullptr; // most-attracting candidate keys (pair-shared kernel with tracking)
    size_t pair_kslots_bytes = 0;
    unsigned long long* pair_kpart = nullptr;
    // sharded pair-shared pass: every rank's J-side slot buffers mapped into this address space (cudaIpc, same node), so that
    // the force kernel stores J-side partials straight into the owning rank's memory over NVLink (capi.cu pair_peer_setup)
    static constexpr int kMaxPeers = 8;
    double* pair_slot_peer[kMaxPeers] = {};
    unsigned long long* pair_kslot_peer[kMaxPeers] = {};
    bool pair_peer_ok = false;      // the tables above are valid
    // ... and both position buffers of every rank (same exchange): the sharded pair kernel then takes the J tiles of remote
    // blocks straight from their owner over NVLink and does not wait for the gather of the next positions (capi.cu force_pass)
    double4* pair_pos_peer[2][kMaxPeers] = {};
    double4* pair_pos_exported[2] = { nullptr, nullptr }; // this rank's buffers as the peers have them mapped
    bool pair_pospeer_ok = false;
    // ... and a word per peer for the barrier in front of such a pass ("every rank's shard of these positions is written"): a
    // device-side flag exchange over the same mappings -- an NCCL collective would queue behind the gather on the communicator
    unsigned long long* pair_flags = nullptr;               // [kMaxPeers] epochs announced TO this rank, one word per sender
    unsigned long long* pair_flags_peer[kMaxPeers] = {};
    unsigned long long pair_epoch = 0;
    bool gather_deferred = false;   // gather_next has marked the buffer; the NCCL calls are still to be issued (gather_issue)
    double4* gather_buf = nullptr;
    bool pair_peer_pos_pass = false; // the pass being launched reads remote blocks from their owners
    std::vector<void*> pos_graveyard; // exported buffers that a grown world replaced: freed with the context (peers may map them)
    bool pair_peer_failed = false;  // peer mapping is not available here: NCCL reduce-scatter fallback
    unsigned long long* peer_xchg = nullptr; // device scratch for the handle exchange / the per-pass barrier
    unsigned long long* pair_kall = nullptr; // sharded: every rank's candidate key for each of this rank's bodies [world][n / world]
    size_t pair_kall_n = 0;
    unsigned* pair_prior = nullptr; // candidate-filter thresholds carried from pass to pass (force_paired.cu)
    size_t pair_prior_n = 0;
    int pair_prior_npad = 0;
    bool pair_prior_valid = false;
    size_t pair_kpart_bytes = 0;
    // mass-sorted layout of the tracked pair-shared pass (single GPU; force_paired.cu): slot s of the sorted domain holds body
    // pair_perm[s]; records gathered into pair_pos_s before every pass; sort scratch
    int* pair_perm = nullptr;
    size_t pair_perm_n = 0;
    double4* pair_pos_s = nullptr;
    size_t pair_pos_s_n = 0;
    void* pair_sort_tmp = nullptr;
    size_t pair_sort_tmp_bytes = 0;
    bool pair_sorted_pass = false; // the pass in flight (compute + finish) works in the sorted domain
    // K1g (force_grid.cu): grid-barrier arrival counter, its timeout flag and the host's copy of the arrivals so far
    unsigned* grid_bar = nullptr;
    unsigned* grid_timed_out = nullptr; // mapped pinned host word
    unsigned grid_bar_count = 0;
    bool attr_grid[2] = { false, false };
    size_t grid_smem[2] = { 0, 0 };

    // recording
    int tracked = 0;
    std::vector<int32_t> trail_cap;
    std::vector<int64_t> vert_off_h;
    double* hist = nullptr;
    double* hist_first = nullptr; // History::m_first_entry, [tracked][6]
    int32_t* hist_start = nullptr;
    int32_t* hist_len = nullptr;
    float* verts = nullptr;
    int64_t* vert_off = nullptr;
    int32_t* tcap = nullptr;
    int32_t* append_off = nullptr;
    int32_t* tlength = nullptr;
    double* toff = nullptr;
    double* last_angle = nullptr;
    int64_t verts_total = 0;

    // ensemble
    int ens_copies = 0;
    int ens_n = 0;
    essa_ensemble_opts ens {};
    double* ens_state = nullptr; // [copies][6][n]
    double* ens_gf = nullptr;    // [n]
    double* ens_mind = nullptr;  // [copies][n]
    uint8_t* ens_active = nullptr; // [n]  Object::deleted() of the template
    double* ens_apos = nullptr;  // [copies][n][6] positions at the closest approach (closest_approaches)
    double* ens_trail = nullptr; // [copies][trail_capacity][3] the designated body's trail
    int* ens_trail_len = nullptr; // [copies]
    long long ens_steps_done = 0;
    double ens_gf_host[essa::kEnsembleSoloMax] = {}; // the template's weights / live mask on the host (kernel parameters of k_ensemble_solo)
    unsigned ens_on_mask = 0;
    bool ens_solo = true;                      // ESSA_ENSEMBLE_SOLO=0: measurement switch, small templates take the duo kernel

    // closest approaches of a forward-simulated world: [n][n][7]
    double* closest = nullptr;
    int closest_n = 0;

    // energy scratch
    double* escratch = nullptr;
    size_t escratch_n = 0;

    // sharding
    int rank = 0, world = 1;
    ncclComm* comm = nullptr;

    // persistent resident kernel
    struct {
        bool enabled = true;
        int idle_us = 1000;        // the kernel retires by itself after this long without a doorbell
        essa::PersistMailbox* mailbox = nullptr;
        bool live = false;         // a kernel was launched and has not been seen to retire
        bool pending = false;      // a tick was posted and its acknowledgement not yet consumed
        int pending_steps = 0;     // ... this one (signed)
        bool result_valid = false; // `staged` holds the state as of the device's current state
        double staged[18 * essa::kPersistMaxBodies]; // the last tick's result in essa_download's staging layout
        unsigned seq = 0;
        long long granted = 0;     // steps handed to the live kernel so far
        essa_step_opts opts {};    // what the live kernel was launched with
        int sign = 1;
        int64_t ticks = 0, launches = 0, idle_exits = 0; // doorbell ticks served / kernels started / kernels found retired
        // host clock: nanoseconds between a doorbell (or launch) and its result / between a result and the next doorbell
        int64_t wait_ns = 0, host_ns = 0;
        std::chrono::steady_clock::time_point t_post, t_landed;
    } persist;

    // stats
    int64_t launches = 0;
    int64_t passes = 0;
    double interactions = 0;
    double last_step_ms = 0;
    double last_force_ms = 0;
    // phases of the LAST pair-shared force pass that was timed by itself: before the kernel (waits on the position gather, the
    // threshold broadcast, the sort / gather of the mass-sorted layout), the pair kernel, the exchange after it (sharded: the
    // one-word barrier, or the NCCL reduce-scatter fallback), k_pair_finish
    cudaEvent_t ev_phase[3] = { nullptr, nullptr, nullptr };
    cudaEvent_t ev_phase_first = nullptr, ev_phase_last = nullptr; // the pass's own e0 / e1 (owned by force_events)
    bool phases_recorded = false;
    double last_phase_ms[4] = { 0, 0, 0, 0 };
    double last_pre_ms = 0, last_post_ms = 0; // before the first / after the last force pass of the last essa_step
    std::string err;

    void* l2_scratch = nullptr;
    int l2_value = 0;
    void* pinned = nullptr; // staging for uploads / downloads
    size_t pinned_bytes = 0;
};

namespace essa {

// kernels' host launchers (defined in the .cu files)
cudaError_t launch_force_pass(essa_ctx* c, int j0, int j1, int S, int slot0, int slots_total, bool argmax, bool save_old,
    KickArgs const& k, double eps2, cudaStream_t st);
cudaError_t launch_kick_drift(essa_ctx* c, KickArgs const& k, cudaStream_t st);
cudaError_t launch_refresh_active(essa_ctx* c, cudaStream_t st);
cudaError_t launch_record(essa_ctx* c, int offset_trails, int forward_simulated, cudaStream_t st);
cudaError_t launch_record_init(essa_ctx* c, int first, cudaStream_t st);
cudaError_t launch_fill_f64(essa_ctx* c, double* dst, double v, size_t n, cudaStream_t st);
cudaError_t launch_fill_i64(essa_ctx* c, int64_t* dst, int64_t v, size_t n, cudaStream_t st);
cudaError_t launch_pack_range(essa_ctx* c, double const* x, double const* y, double const* z, int i0, int i1, cudaStream_t st);
cudaError_t launch_pack(essa_ctx* c, double const* x, double const* y, double const* z, cudaStream_t st);
cudaError_t launch_unpack(essa_ctx* c, double* x, double* y, double* z, cudaStream_t st);
cudaError_t pair_reserve(essa_ctx* c, bool argmax, bool* changed);
bool pair_reserve_would_change(essa_ctx const* c, bool argmax);
bool paired_plan(essa_ctx const* c, int* nb_out, int* S_out, size_t* slot_doubles, size_t* part_doubles, int* ri_out = nullptr);
cudaError_t launch_force_paired_compute(essa_ctx* c, double eps2, cudaStream_t st, double** acc_out, int* npad_out, bool argmax,
    unsigned long long** kacc_out);
cudaError_t launch_force_paired_finish(essa_ctx* c, KickArgs const& k, cudaStream_t st, double const* ajred, bool argmax_noop, bool save_old,
    bool argmax, unsigned long long const* akred, double eps2);
bool force_grid_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only);
cudaError_t launch_force_grid(essa_ctx* c, int steps, essa_step_opts const& o, bool first_pass, bool argmax, cudaStream_t st);
int force_iblocks(essa_ctx const* c, int* R_out);
void choose_split(essa_ctx const* c, int iblocks, int nquanta, int* S_out);

cudaError_t launch_resident(essa_ctx* c, int steps, essa_step_opts const& o, bool forces_only, cudaStream_t st);
cudaError_t launch_resident_warp(essa_ctx* c, int steps, essa_step_opts const& o, bool mask_may_change, cudaStream_t st,
    bool persist = false);
bool resident_persist_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only);
bool resident_cta_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only);
cudaError_t launch_resident_cta(essa_ctx* c, int steps, essa_step_opts const& o, cudaStream_t st);
cudaError_t launch_download_small(essa_ctx* c, void* pinned_out, cudaStream_t st);
bool resident_cluster_wide_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only);
cudaError_t launch_resident_cluster_wide(essa_ctx* c, int steps, essa_step_opts const& o, cudaStream_t st);
bool resident_cluster_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only);
cudaError_t launch_resident_cluster(essa_ctx* c, int steps, essa_step_opts const& o, cudaStream_t st, bool persist = false);
bool resident_cluster_persist_applicable(essa_ctx const* c, essa_step_opts const& o, bool forces_only);
cudaError_t launch_ensemble_init(essa_ctx* c, cudaStream_t st);
cudaError_t launch_ensemble_step(essa_ctx* c, int steps, cudaStream_t st);
cudaError_t launch_energy(essa_ctx* c, cudaStream_t st); // result in c->escratch[0..5)
double run_fp64_peak(essa_ctx* c, double ms);
int run_rsqrt_error(essa_ctx* c, double* seed, double* refined);

RecordPtrs record_ptrs(essa_ctx* c);
WorldPtrs world_ptrs(essa_ctx* c);

} // namespace essa
