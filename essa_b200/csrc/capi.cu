// C ABI of the World::update hot path (see include/essa_b200.h for the reference interfaces each
// entry point replaces).  Host code only: context, transfers, step scheduling, NCCL plumbing.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <climits>
#include <functional>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ctx.hpp"

using namespace essa;

static thread_local std::string g_create_error;

// Object::deleted() on the host (src/Object.cpp:60-62), same expression as common.cuh body_active.
static inline bool body_active_host(bool del_flag, int64_t ddate, int64_t cdate, int64_t date) {
    return !((del_flag && ddate <= date) || cdate > date);
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                              \
            return ESSA_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

#define ARG(cond, msg)                                                                                   \
    do {                                                                                                 \
        if (!(cond)) {                                                                                   \
            if (ctx)                                                                                     \
                ctx->err = msg;                                                                          \
            return ESSA_ERR_ARG;                                                                         \
        }                                                                                                \
    } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen: single-GPU users never need the library; inside a torch process the
// already loaded libnccl.so.2 is picked up by soname.
// ---------------------------------------------------------------------------------------------
namespace {
struct NcclId {
    char internal[ESSA_NCCL_ID_BYTES];
};
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(ncclComm**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(ncclComm*) = nullptr;
    int (*AllGather)(void const*, void*, size_t, int, ncclComm*, cudaStream_t) = nullptr;
    int (*Broadcast)(void const*, void*, size_t, int, int, ncclComm*, cudaStream_t) = nullptr;
    int (*AllReduce)(void const*, void*, size_t, int, int, ncclComm*, cudaStream_t) = nullptr;
    int (*ReduceScatter)(void const*, void*, size_t, int, int, ncclComm*, cudaStream_t) = nullptr;
    int (*Send)(void const*, size_t, int, int, ncclComm*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    char const* (*GetErrorString)(int) = nullptr;
    std::string why;
};
constexpr int kNcclFloat64 = 8, kNcclUint64 = 5, kNcclUint32 = 3, kNcclSum = 0; // ncclDataType_t / ncclRedOp_t values (nccl.h)

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried)
        return api;
    tried = true;
    char const* names[] = { "libnccl.so.2", "libnccl.so" };
    for (char const* nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib)
            break;
    }
    if (!api.lib) {
        api.why = std::string("libnccl.so.2 not loadable: ") + dlerror();
        return api;
    }
#define SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name))
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather");
    SYM(Broadcast, "ncclBroadcast");
    SYM(AllReduce, "ncclAllReduce");
    SYM(ReduceScatter, "ncclReduceScatter");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    if (!api.GetUniqueId || !api.CommInitRank || !api.Broadcast || !api.GroupStart || !api.GroupEnd || !api.AllReduce || !api.ReduceScatter || !api.Send || !api.Recv) {
        api.why = "libnccl.so.2 lacks required symbols";
        dlclose(api.lib);
        api.lib = nullptr;
    }
    return api;
}
} // namespace

#define NK(call)                                                                                         \
    do {                                                                                                 \
        int r__ = (call);                                                                                \
        if (r__ != 0) {                                                                                  \
            ctx->err = std::string(#call) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r__) : "nccl error"); \
            return ESSA_ERR_NCCL;                                                                        \
        }                                                                                                \
    } while (0)

// ---------------------------------------------------------------------------------------------
// persistent resident kernel (resident_warp.cu, K3q): the kernel that served the previous
// World::update(steps) stays on its two SMs and takes the next tick from a mailbox in mapped pinned
// memory -- no launch, no stream synchronisation and no packing kernel per tick.  What the GUI's
// cadence needs (SimulationView::update calls World::update(iterations) once per tick and reads the
// objects back, src/SimulationView.cpp:374-397).  Every other entry point parks the kernel first
// (PARK): it stores its state to HBM exactly as a one-shot launch does at its end, so nothing else in
// this file knows about it.  A kernel nobody rings for `idle_us` retires by itself.
// ---------------------------------------------------------------------------------------------
#include <chrono>

static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}

static int persist_launch(essa_ctx* ctx, int steps, essa_step_opts const& o);

// The tick in flight has been published (ack == seq).  A kernel found retired (idle timeout raced with the doorbell)
// never saw the tick: it is started again with that tick as its first.
static int persist_wait_ack(essa_ctx* ctx) {
    auto& p = ctx->persist;
    if (!p.live || !p.pending)
        return ESSA_OK;
    PersistMailbox* mb = p.mailbox;
    auto const t0 = std::chrono::steady_clock::now();
    double const limit_s = 10.0 + 5e-6 * (double)p.pending_steps;
    int const n = ctx->n;
    bool const field_major = n <= kPersistFieldMajorMax;
    auto landed = [&]() { // every word of the result carries this tick's number (checked back to front)
        if (field_major) {
            for (int w = 29 * n + 1; w >= 0; w--)
                if ((unsigned)(__atomic_load_n(&mb->ll[w], __ATOMIC_RELAXED) >> 32) != p.seq)
                    return false;
            __atomic_thread_fence(__ATOMIC_ACQUIRE);
            return true;
        }
        for (int w = kPersistRecWords * n + 1; w >= kPersistRecWords * n; w--)
            if ((unsigned)(__atomic_load_n(&mb->ll[w], __ATOMIC_RELAXED) >> 32) != p.seq)
                return false;
        // the time words are the last the kernel writes: the records (256 bytes per body, fresh from the PCIe DMA and in no
        // cache of this core) are here or about to be -- have all their lines on the way before the first one is looked at
        for (int k = 0; k < n; k++)
            for (int line = 0; line < kPersistRecWords * 8; line += 64)
                __builtin_prefetch(reinterpret_cast<char const*>(mb->ll + kPersistRecWords * k) + line, 0, 3);
        for (int k = n - 1; k >= 0; k--)
            for (int w = kPersistRecUsed - 1; w >= 0; w--)
                if ((unsigned)(__atomic_load_n(&mb->ll[kPersistRecWords * k + w], __ATOMIC_RELAXED) >> 32) != p.seq)
                    return false;
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        return true;
    };
    for (unsigned spin = 0;; spin++) {
        if (landed())
            break;
        if (__atomic_load_n(&mb->ack[1], __ATOMIC_ACQUIRE) != 0) {
            // retired without seeing the doorbell
            CK(cudaStreamSynchronize(ctx->stream));
            if (landed())
                break;
            p.live = false;
            p.idle_exits++;
            p.ticks--; // the doorbell was not answered: this tick is the new kernel's first
            int rc = persist_launch(ctx, p.pending_steps, p.opts);
            if (rc != ESSA_OK)
                return rc;
            continue;
        }
        if ((spin & 0xfff) == 0xfff) {
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s) {
                ctx->err = "persistent resident kernel does not answer";
                return ESSA_ERR_CUDA;
            }
            cudaError_t e = cudaStreamQuery(ctx->stream); // a kernel that died shows up here
            if (e != cudaSuccess && e != cudaErrorNotReady) {
                p.live = false;
                ctx->err = std::string("persistent resident kernel: ") + cudaGetErrorString(e);
                return ESSA_ERR_CUDA;
            }
        }
        cpu_relax();
    }
    p.pending = false;
    p.result_valid = true;
    p.t_landed = std::chrono::steady_clock::now();
    p.wait_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(p.t_landed - p.t_post).count();
    // unpack into essa_download's staging layout: 15 n doubles (field 9, gf, is not carried) ... links at double offset 17 n
    if (field_major) {
        size_t const N = (size_t)n;
        for (int f = 0; f < 14; f++) {
            double* dst = p.staged + (size_t)(f < 9 ? f : f + 1) * N;
            unsigned long long const* src = mb->ll + (size_t)f * 2 * N;
            for (size_t k = 0; k < N; k++) {
                unsigned long long const bits = (src[2 * k] & 0xffffffffull) | (src[2 * k + 1] << 32);
                memcpy(dst + k, &bits, sizeof(double));
            }
        }
        int32_t* most = reinterpret_cast<int32_t*>(p.staged + 17 * N);
        for (size_t k = 0; k < N; k++)
            most[k] = (int32_t)(unsigned)mb->ll[28 * N + k];
        ctx->last_step_ms = (double)((mb->ll[29 * N] & 0xffffffffull) | (mb->ll[29 * N + 1] << 32)) * 1e-6;
    } else {
        size_t const N = (size_t)n;
        int32_t* most = reinterpret_cast<int32_t*>(p.staged + 17 * N);
        for (size_t k = 0; k < N; k++) {
            unsigned long long const* rec = mb->ll + kPersistRecWords * k;
            for (int f = 0; f < 14; f++) {
                unsigned long long const bits = (rec[2 * f] & 0xffffffffull) | (rec[2 * f + 1] << 32);
                memcpy(p.staged + (size_t)(f < 9 ? f : f + 1) * N + k, &bits, sizeof(double));
            }
            most[k] = (int32_t)(unsigned)rec[28];
        }
        unsigned long long const* tw = mb->ll + kPersistRecWords * N;
        ctx->last_step_ms = (double)((tw[0] & 0xffffffffull) | (tw[1] << 32)) * 1e-6;
    }
    return ESSA_OK;
}

// Retire the kernel (its epilogue stores the world and the recording cursors to HBM) and wait for it.
static int persist_park(essa_ctx* ctx) {
    auto& p = ctx->persist;
    if (!p.live)
        return ESSA_OK;
    int rc = persist_wait_ack(ctx);
    if (p.live) {
        p.seq++;
        __atomic_store_n(&p.mailbox->cmd, (unsigned long long)p.seq << 32, __ATOMIC_RELEASE);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    p.live = false;
    p.pending = false;
    p.result_valid = false;
    if (rc != ESSA_OK)
        return rc;
    CK(e);
    return ESSA_OK;
}

#define PARK()                                                                                           \
    do {                                                                                                 \
        if (ctx->persist.live) {                                                                         \
            int prc__ = persist_park(ctx);                                                               \
            if (prc__ != ESSA_OK)                                                                        \
                return prc__;                                                                            \
        }                                                                                                \
    } while (0)

// Start a persistent kernel whose first tick is `steps`.
static int persist_launch(essa_ctx* ctx, int steps, essa_step_opts const& o) {
    auto& p = ctx->persist;
    if (!p.mailbox) {
        void* m = nullptr;
        CK(cudaHostAlloc(&m, sizeof(PersistMailbox), cudaHostAllocMapped | cudaHostAllocPortable));
        memset(m, 0, sizeof(PersistMailbox));
        p.mailbox = static_cast<PersistMailbox*>(m);
    }
    p.seq++;
    int const k = steps < 0 ? -steps : steps;
    p.mailbox->ack[1] = 0;
    __atomic_store_n(&p.mailbox->cmd, ((unsigned long long)p.seq << 32) | (unsigned)k, __ATOMIC_RELEASE);
    if (ctx->n <= 32)
        CK(launch_resident_warp(ctx, steps, o, false, ctx->stream, true)); // K3q
    else
        CK(launch_resident_cluster(ctx, steps, o, ctx->stream, true));     // K3x
    p.t_post = std::chrono::steady_clock::now();
    p.live = true;
    p.pending = true;
    p.pending_steps = steps;
    p.result_valid = false;
    p.opts = o;
    p.sign = steps >= 0 ? 1 : -1;
    p.granted = k;
    p.launches++;
    return ESSA_OK;
}

// Can the live kernel take this call as its next tick?
static bool persist_takes(essa_ctx const* ctx, int steps, essa_step_opts const& o) {
    auto const& p = ctx->persist;
    long long const k = steps < 0 ? -(long long)steps : steps;
    return p.live && p.enabled && (steps >= 0 ? 1 : -1) == p.sign && o.dt_seconds == p.opts.dt_seconds
        && o.forward_simulated == p.opts.forward_simulated && o.offset_trails == p.opts.offset_trails && o.record == p.opts.record
        && o.track_most_attracting == p.opts.track_most_attracting && o.exact_order == p.opts.exact_order && o.two_pass == p.opts.two_pass
        && o.eps2 == p.opts.eps2 && (o.kernel == ESSA_KERNEL_AUTO || o.kernel == ESSA_KERNEL_RESIDENT)
        && p.granted + k < (1ll << 30); // the kernel counts steps in 32-bit integers
}

// Ring the doorbell.
static void persist_post(essa_ctx* ctx, int steps) {
    auto& p = ctx->persist;
    unsigned const k = (unsigned)(steps < 0 ? -steps : steps);
    p.seq++;
    p.t_post = std::chrono::steady_clock::now();
    if (p.ticks > 0 || p.launches > 0)
        p.host_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(p.t_post - p.t_landed).count();
    __atomic_store_n(&p.mailbox->cmd, ((unsigned long long)p.seq << 32) | k, __ATOMIC_RELEASE);
    p.pending = true;
    p.pending_steps = steps;
    p.result_valid = false;
    p.granted += k;
    p.ticks++;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
static void pair_peer_close(essa_ctx* ctx);
constexpr int kXchgWords = 64; // words per rank of the handle exchange (pair_peer_setup)

static void free_world(essa_ctx* c) {
    for (int k = 0; k < 2; k++) {
        if (c->pos[k] && c->pos[k] == c->pair_pos_exported[k])
            c->pos_graveyard.push_back(c->pos[k]); // other ranks may still have it mapped: lives until the context goes
        else
            cudaFree(c->pos[k]);
    }
    cudaFree(c->vel);
    cudaFree(c->velh);
    cudaFree(c->acc);
    cudaFree(c->gf);
    cudaFree(c->cdate);
    cudaFree(c->ddate);
    cudaFree(c->delflag);
    cudaFree(c->active);
    cudaFree(c->most);
    cudaFree(c->old_most);
    cudaFree(c->maxattr);
    cudaFree(c->appe);
    c->pos[0] = c->pos[1] = nullptr;
    c->vel = c->velh = c->acc = c->gf = c->maxattr = c->appe = nullptr;
    c->cdate = c->ddate = nullptr;
    c->delflag = c->active = nullptr;
    c->most = c->old_most = nullptr;
    c->cap = 0;
}

static int reserve_world(essa_ctx* ctx, int cap) {
    if (cap <= ctx->cap)
        return ESSA_OK;
    // Growing keeps nothing: callers re-upload (essa_upload is the only entry that grows).
    free_world(ctx);
    cap = (cap + 2047) / 2048 * 2048; // whole blocks of the pair-shared kernel (force_paired.cu)
    size_t n = (size_t)cap;
    CK(cudaMalloc(&ctx->pos[0], n * sizeof(double4)));
    CK(cudaMalloc(&ctx->pos[1], n * sizeof(double4)));
    CK(cudaMalloc(&ctx->vel, 3 * n * sizeof(double)));
    CK(cudaMalloc(&ctx->velh, 3 * n * sizeof(double)));
    CK(cudaMalloc(&ctx->acc, 3 * n * sizeof(double)));
    CK(cudaMalloc(&ctx->gf, n * sizeof(double)));
    CK(cudaMalloc(&ctx->cdate, n * sizeof(int64_t)));
    CK(cudaMalloc(&ctx->ddate, n * sizeof(int64_t)));
    CK(cudaMalloc(&ctx->delflag, n));
    CK(cudaMalloc(&ctx->active, n));
    CK(cudaMalloc(&ctx->most, n * sizeof(int32_t)));
    CK(cudaMalloc(&ctx->old_most, n * sizeof(int32_t)));
    CK(cudaMalloc(&ctx->maxattr, n * sizeof(double)));
    CK(cudaMalloc(&ctx->appe, 4 * n * sizeof(double)));
    CK(cudaMemsetAsync(ctx->pos[0], 0, n * sizeof(double4), ctx->stream));
    CK(cudaMemsetAsync(ctx->pos[1], 0, n * sizeof(double4), ctx->stream));
    CK(cudaMemsetAsync(ctx->acc, 0, 3 * n * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->velh, 0, 3 * n * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->maxattr, 0, n * sizeof(double), ctx->stream));
    ctx->cap = cap;
    return ESSA_OK;
}

static int reserve_pinned(essa_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->pinned_bytes)
        return ESSA_OK;
    if (ctx->pinned)
        cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    bytes = (bytes + (1u << 20)) & ~((size_t)(1u << 20) - 1);
    CK(cudaHostAlloc(&ctx->pinned, bytes, cudaHostAllocDefault));
    ctx->pinned_bytes = bytes;
    return ESSA_OK;
}

extern "C" int essa_abi_version(void) { return ESSA_ABI_VERSION; }

extern "C" int essa_ctx_create(int device, int capacity, essa_ctx** out) {
    if (!out) {
        g_create_error = "out is null";
        return ESSA_ERR_ARG;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
        return ESSA_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        g_create_error = "device ordinal out of range";
        return ESSA_ERR_ARG;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return ESSA_ERR_CUDA;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return ESSA_ERR_CUDA;
    }
    if (prop.major != 10) {
        g_create_error = "device is not sm_100 (this library carries sm_100a code only)";
        return ESSA_ERR_CUDA;
    }
    essa_ctx* ctx = new essa_ctx();
    ctx->device = device;
    if (char const* e = getenv("ESSA_PERSIST"))
        ctx->persist.enabled = atoi(e) != 0;
    if (char const* e = getenv("ESSA_PERSIST_IDLE_US"))
        ctx->persist.idle_us = std::max(1, std::min(1000000, atoi(e)));
    ctx->sm_count = prop.multiProcessorCount;
    auto fail = [&](char const* what, cudaError_t ce) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        essa_ctx_destroy(ctx);
        return ESSA_ERR_CUDA;
    };
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return fail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return fail("cudaStreamCreate", e);
    cudaEvent_t* evs[] = { &ctx->ev_begin, &ctx->ev_end, &ctx->ev_f0, &ctx->ev_f1 };
    for (auto* ev : evs)
        if ((e = cudaEventCreate(ev)) != cudaSuccess)
            return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_gather, cudaEventDisableTiming)) != cudaSuccess)
        return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming)) != cudaSuccess)
        return fail("cudaEventCreate", e);
    if (capacity > 0) {
        int rc = reserve_world(ctx, capacity);
        if (rc != ESSA_OK) {
            g_create_error = ctx->err;
            essa_ctx_destroy(ctx);
            return rc;
        }
    }
    *out = ctx;
    return ESSA_OK;
}

static void free_recording(essa_ctx* c) {
    cudaFree(c->hist);
    cudaFree(c->hist_first);
    cudaFree(c->hist_start);
    cudaFree(c->hist_len);
    cudaFree(c->verts);
    cudaFree(c->vert_off);
    cudaFree(c->tcap);
    cudaFree(c->append_off);
    cudaFree(c->tlength);
    cudaFree(c->toff);
    cudaFree(c->last_angle);
    c->hist = c->hist_first = nullptr;
    c->hist_start = c->hist_len = c->tcap = c->append_off = c->tlength = nullptr;
    c->verts = nullptr;
    c->vert_off = nullptr;
    c->toff = c->last_angle = nullptr;
    c->tracked = 0;
    c->verts_total = 0;
    c->trail_cap.clear();
    c->vert_off_h.clear();
}

extern "C" void essa_ctx_destroy(essa_ctx* c) {
    if (!c)
        return;
    cudaSetDevice(c->device);
    (void)persist_park(c);
    if (c->pair_peer_ok) { // let go of the peers' buffers, and give them the chance to let go of ours before it is freed
        pair_peer_close(c);
        if (c->comm && c->peer_xchg && nccl().AllReduce
            && nccl().AllReduce(c->peer_xchg, c->peer_xchg, 1, kNcclUint32, kNcclSum, c->comm, c->stream) == 0)
            cudaStreamSynchronize(c->stream);
    }
    cudaFree(c->peer_xchg);
    cudaFree(c->pair_flags);
    for (void* g : c->pos_graveyard)
        cudaFree(g);
    c->pos_graveyard.clear();
    c->pair_pos_exported[0] = c->pair_pos_exported[1] = nullptr; // the peers have let go: free_world below frees for real
    if (c->persist.mailbox)
        cudaFreeHost(c->persist.mailbox);
    if (c->stream)
        cudaStreamSynchronize(c->stream);
    if (c->comm && nccl().CommDestroy)
        nccl().CommDestroy(c->comm);
    free_world(c);
    free_recording(c);
    cudaFree(c->pair_slots);
    cudaFree(c->pair_part);
    cudaFree(c->pair_kslots);
    cudaFree(c->pair_kpart);
    cudaFree(c->pair_prior);
    cudaFree(c->pair_perm);
    cudaFree(c->pair_pos_s);
    cudaFree(c->pair_sort_tmp);
    cudaFree(c->grid_bar);
    if (c->grid_timed_out)
        cudaFreeHost(c->grid_timed_out);
    cudaFree(c->pair_kall);
    cudaFree(c->part);
    cudaFree(c->part_m);
    cudaFree(c->part_j);
    cudaFree(c->counters);
    cudaFree(c->ens_state);
    cudaFree(c->ens_gf);
    cudaFree(c->ens_mind);
    cudaFree(c->ens_active);
    cudaFree(c->ens_apos);
    cudaFree(c->ens_trail);
    cudaFree(c->ens_trail_len);
    cudaFree(c->escratch);
    cudaFree(c->closest);
    cudaFree(c->l2_scratch);
    if (c->pinned)
        cudaFreeHost(c->pinned);
    cudaEvent_t evs[] = { c->ev_begin, c->ev_end, c->ev_f0, c->ev_f1, c->ev_gather, c->ev_ready, c->ev_phase[0], c->ev_phase[1], c->ev_phase[2] };
    for (auto ev : evs)
        if (ev)
            cudaEventDestroy(ev);
    for (auto ev : c->force_events)
        cudaEventDestroy(ev);
    if (c->stream)
        cudaStreamDestroy(c->stream);
    if (c->comm_stream)
        cudaStreamDestroy(c->comm_stream);
    delete c;
}

extern "C" const char* essa_last_error(const essa_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }
extern "C" int64_t essa_launch_count(const essa_ctx* c) { return c ? c->launches : 0; }
extern "C" int64_t essa_pass_count(const essa_ctx* c) { return c ? c->passes : 0; }
extern "C" double essa_interaction_count(const essa_ctx* c) { return c ? c->interactions : 0; }
// essa_step_async leaves the end event unread: settle it the first time anybody asks
static void settle_timing(essa_ctx const* cc) {
    essa_ctx* c = const_cast<essa_ctx*>(cc);
    if (!c || !c->timing_pending)
        return;
    c->timing_pending = false;
    float ms = 0;
    if (cudaEventSynchronize(c->ev_end) == cudaSuccess && cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end) == cudaSuccess)
        c->last_step_ms = ms;
}
extern "C" double essa_last_step_ms(const essa_ctx* c) {
    settle_timing(c);
    return c ? c->last_step_ms : 0;
}
extern "C" double essa_last_force_ms(const essa_ctx* c) { return c ? c->last_force_ms : 0; }
extern "C" int essa_last_force_passes(const essa_ctx* c) { return c ? c->last_force_passes : 0; }
extern "C" int essa_last_step_breakdown(const essa_ctx* c, double out[4]) {
    if (!c || !out)
        return ESSA_ERR_ARG;
    settle_timing(c);
    out[0] = c->last_step_ms;
    out[1] = c->last_force_ms;
    out[2] = c->last_pre_ms;
    out[3] = c->last_post_ms;
    return ESSA_OK;
}
extern "C" int essa_last_pass_phases(const essa_ctx* c, double out[4]) {
    if (!c || !out)
        return ESSA_ERR_ARG;
    settle_timing(c);
    for (int i = 0; i < 4; i++)
        out[i] = c->last_phase_ms[i];
    return ESSA_OK;
}
extern "C" int essa_body_count(const essa_ctx* c) { return c ? c->n : 0; }
extern "C" int64_t essa_get_date(const essa_ctx* c) { return c ? c->date : 0; }
extern "C" int essa_tracked_count(const essa_ctx* c) { return c ? c->tracked : 0; }
extern "C" int essa_shard_rank(const essa_ctx* c) { return c ? c->rank : 0; }
extern "C" int essa_shard_world(const essa_ctx* c) { return c ? c->world : 1; }

extern "C" int essa_sync(essa_ctx* ctx) {
    ARG(ctx, "ctx is null");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->comm_stream));
    return ESSA_OK;
}

// ---------------------------------------------------------------------------------------------
// active mask bookkeeping (Object::deleted(), src/Object.cpp:60-62): the mask flips exactly when
// the date crosses a creation date, or the deletion date of a body whose m_deleted is set.
// ---------------------------------------------------------------------------------------------
static bool mask_changes(essa_ctx const* c, int64_t d_from, int64_t d_to) {
    if (c->events.empty() || d_from == d_to)
        return false;
    int64_t lo = std::min(d_from, d_to), hi = std::max(d_from, d_to);
    auto it = std::upper_bound(c->events.begin(), c->events.end(), lo); // first t > lo
    return it != c->events.end() && *it <= hi;
}

extern "C" int essa_set_date(essa_ctx* ctx, int64_t date) {
    ARG(ctx, "ctx is null");
    CK(cudaSetDevice(ctx->device));
    PARK();
    bool changed = mask_changes(ctx, ctx->date, date);
    ctx->date = date;
    if (changed && ctx->n > 0) {
        CK(launch_refresh_active(ctx, ctx->stream));
        ctx->pair_prior_valid = false;
        ctx->acc_valid = false;
    }
    return ESSA_OK;
}

// ---------------------------------------------------------------------------------------------
// upload / download
// ---------------------------------------------------------------------------------------------
// Host edge of large worlds: the caller's SoA arrays are ordinary pageable memory, the DMA engine wants pinned memory.
// Copying 59 MB (N = 1,048,576: x y z vx vy vz gf) into the staging buffer with one thread ran at 7 GB/s and was the
// longest part of an end-to-end tick's host side; a handful of threads bring it to memory bandwidth.
static void parallel_chunks(size_t total, size_t chunk, int max_threads, std::function<void(size_t, size_t)> const& fn) {
    size_t const nchunks = (total + chunk - 1) / chunk;
    int threads = (int)std::min<size_t>(nchunks, (size_t)max_threads);
    if (threads <= 1) {
        if (total)
            fn(0, total);
        return;
    }
    std::atomic<size_t> next { 0 };
    auto work = [&]() {
        for (;;) {
            size_t c = next.fetch_add(1);
            if (c >= nchunks)
                return;
            size_t lo = c * chunk, hi = std::min(total, lo + chunk);
            fn(lo, hi);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++)
        pool.emplace_back(work);
    work();
    for (auto& th : pool)
        th.join();
}

static int host_threads() {
    static int const v = [] {
        if (char const* e = getenv("ESSA_HOST_THREADS"))
            return std::max(1, atoi(e));
        unsigned hc = std::thread::hardware_concurrency();
        return (int)std::max(1u, std::min(8u, hc ? hc / 2 : 4u));
    }();
    return v;
}

extern "C" int essa_upload(essa_ctx* ctx, int n, const essa_bodies* h) {
    ARG(ctx, "ctx is null");
    ARG(h && n >= 0, "bad arguments");
    ARG(n == 0 || (h->x && h->y && h->z && h->vx && h->vy && h->vz && h->gf), "x,y,z,vx,vy,vz,gf are required");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    int rc = reserve_world(ctx, n);
    if (rc != ESSA_OK)
        return rc;
    ctx->n = n;
    ctx->acc_valid = false;
    ctx->acc_has_most = false;
    ctx->pair_prior_valid = false;
    ctx->events.clear();
    if (n == 0)
        return ESSA_OK;
    size_t N = (size_t)n, cap = (size_t)ctx->cap;
    if (h->most)
        for (size_t i = 0; i < N; i++)
            if (h->most[i] < -1 || h->most[i] >= n) {
                ctx->err = "most[] must hold list indices in [0, n) or -1";
                ctx->n = 0; // nothing was uploaded
                return ESSA_ERR_ARG;
            }
    // staging layout: 7 N doubles | 4 N doubles (ap pe apv pev) | 2 N int64 | N int32 | 2 N bytes
    size_t bytes = 11 * N * sizeof(double) + 2 * N * sizeof(int64_t) + N * sizeof(int32_t) + 2 * N;
    rc = reserve_pinned(ctx, bytes);
    if (rc != ESSA_OK)
        return rc;
    double* d = static_cast<double*>(ctx->pinned);
    double const* src[7] = { h->x, h->y, h->z, h->vx, h->vy, h->vz, h->gf };
    // required arrays: staged by several threads (chunks of 256 Ki elements of all seven arrays); the same pass finds out
    // whether every gravity factor is the same value (then no body has a heavier partner: Object.cpp:84-94 can never fire)
    std::atomic<bool> uniform { true };
    double const g0 = h->gf[0];
    parallel_chunks(N, (size_t)1 << 18, N >= ((size_t)1 << 17) ? host_threads() : 1, [&](size_t lo, size_t hi) {
        for (int k = 0; k < 7; k++)
            memcpy(d + k * N + lo, src[k] + lo, (hi - lo) * sizeof(double));
        bool u = true;
        for (size_t i = lo; i < hi && u; i++)
            u = h->gf[i] == g0;
        if (!u)
            uniform.store(false, std::memory_order_relaxed);
    });
    ctx->gf_uniform = uniform.load();
    ctx->gf_uniform_value = g0;
    cudaStream_t st = ctx->stream;
    // x y z go through velh (scratch between steps) and are packed into 32-byte j-records
    for (int k = 0; k < 3; k++) {
        CK(cudaMemcpyAsync(ctx->velh + k * cap, d + k * N, N * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->vel + k * cap, d + (3 + k) * N, N * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemcpyAsync(ctx->gf, d + 6 * N, N * sizeof(double), cudaMemcpyHostToDevice, st));
    // ap / pe: members the caller left out get the Object constructor's values ON THE DEVICE (0, DBL_MAX, 0, 0: no host loop, no copy)
    double const* appe_in[4] = { h->ap, h->pe, h->ap_vel, h->pe_vel };
    double const appe_def[4] = { 0.0, DBL_MAX, 0.0, 0.0 };
    for (int k = 0; k < 4; k++) {
        if (appe_in[k]) {
            memcpy(d + (7 + k) * N, appe_in[k], N * sizeof(double));
            CK(cudaMemcpyAsync(ctx->appe + k * cap, d + (7 + k) * N, N * sizeof(double), cudaMemcpyHostToDevice, st));
        } else {
            CK(launch_fill_f64(ctx, ctx->appe + k * cap, appe_def[k], N, st));
        }
    }
    int64_t* dates = reinterpret_cast<int64_t*>(d + 11 * N);
    int32_t* most = reinterpret_cast<int32_t*>(dates + 2 * N);
    uint8_t* flags = reinterpret_cast<uint8_t*>(most + N);
    if (h->creation_date || h->deletion_date || h->deleted) {
        for (size_t i = 0; i < N; i++) {
            int64_t cd = h->creation_date ? h->creation_date[i] : INT64_MIN;
            int64_t dd = h->deletion_date ? h->deletion_date[i] : 0;
            uint8_t del = h->deleted ? (h->deleted[i] != 0) : 0;
            dates[i] = cd;
            dates[N + i] = dd;
            flags[i] = del;
            flags[N + i] = body_active_host(del, dd, cd, ctx->date) ? 1 : 0;
            if (cd != INT64_MIN)
                ctx->events.push_back(cd);
            if (del)
                ctx->events.push_back(dd);
        }
        std::sort(ctx->events.begin(), ctx->events.end());
        ctx->events.erase(std::unique(ctx->events.begin(), ctx->events.end()), ctx->events.end());
        CK(cudaMemcpyAsync(ctx->cdate, dates, N * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->ddate, dates + N, N * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->delflag, flags, N, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->active, flags + N, N, cudaMemcpyHostToDevice, st));
    } else {
        // "always there": creation date in the infinite past, never deleted, active
        CK(launch_fill_i64(ctx, ctx->cdate, INT64_MIN, N, st));
        CK(cudaMemsetAsync(ctx->ddate, 0, N * sizeof(int64_t), st));
        CK(cudaMemsetAsync(ctx->delflag, 0, N, st));
        CK(cudaMemsetAsync(ctx->active, 1, N, st));
    }
    if (h->most) {
        memcpy(most, h->most, N * sizeof(int32_t));
        CK(cudaMemcpyAsync(ctx->most, most, N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->old_most, most, N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    } else { // -1: no most attracting object
        CK(cudaMemsetAsync(ctx->most, 0xff, N * sizeof(int32_t), st));
        CK(cudaMemsetAsync(ctx->old_most, 0xff, N * sizeof(int32_t), st));
    }
    // m_attraction_factor / m_max_attraction are outputs of set_forces(), except for bodies that
    // Object::deleted() masks: those keep whatever they held (clear_forces skips them), so a caller
    // that re-uploads a world with masked bodies may pass them along.  Default 0.
    double const* acc_in[3] = { h->ax, h->ay, h->az };
    for (int k = 0; k < 3; k++) {
        if (acc_in[k])
            CK(cudaMemcpyAsync(ctx->acc + k * cap, acc_in[k], N * sizeof(double), cudaMemcpyHostToDevice, st));
        else
            CK(cudaMemsetAsync(ctx->acc + k * cap, 0, N * sizeof(double), st));
    }
    if (h->max_attraction)
        CK(cudaMemcpyAsync(ctx->maxattr, h->max_attraction, N * sizeof(double), cudaMemcpyHostToDevice, st));
    else
        CK(cudaMemsetAsync(ctx->maxattr, 0, N * sizeof(double), st));
    CK(launch_pack(ctx, ctx->velh, ctx->velh + cap, ctx->velh + 2 * cap, st));
    CK(cudaStreamSynchronize(st));
    return ESSA_OK;
}

constexpr size_t kDownloadSmallMax = 4096; // bodies: one packing kernel instead of one copy per field

extern "C" int essa_download(essa_ctx* ctx, essa_bodies* h) {
    ARG(ctx, "ctx is null");
    ARG(h, "host view is null");
    CK(cudaSetDevice(ctx->device));
    size_t N = (size_t)ctx->n, cap = (size_t)ctx->cap;
    if (ctx->persist.live && (ctx->persist.pending || ctx->persist.result_valid) && !h->gf && !h->creation_date && !h->deletion_date
        && !h->deleted) {
        // the persistent kernel has published (or is about to publish) exactly this state in the mailbox
        int prc = persist_wait_ack(ctx);
        if (prc != ESSA_OK)
            return prc;
        if (ctx->persist.live && ctx->persist.result_valid) {
            double const* d = ctx->persist.staged;
            double* dst[15] = { h->x, h->y, h->z, h->vx, h->vy, h->vz, h->ax, h->ay, h->az, nullptr, h->max_attraction, h->ap, h->pe,
                h->ap_vel, h->pe_vel };
            for (int k = 0; k < 15; k++)
                if (dst[k])
                    memcpy(dst[k], d + k * N, N * sizeof(double));
            if (h->most)
                memcpy(h->most, d + 17 * N, N * sizeof(int32_t));
            return ESSA_OK;
        }
    }
    PARK();
    if (N == 0)
        return ESSA_OK;
    size_t bytes = 15 * N * sizeof(double) + 2 * N * sizeof(int64_t) + N * sizeof(int32_t) + N;
    int rc = reserve_pinned(ctx, bytes);
    if (rc != ESSA_OK)
        return rc;
    cudaStream_t st = ctx->stream;
    double* d = static_cast<double*>(ctx->pinned);
    if (ctx->world > 1) {
        // every rank holds all positions (gathered each pass) but only its own targets' v, a, most-attracting link,
        // max_attraction and ap / pe (post-conditions of World::update that Object::most_attracting_object() / get_info()
        // show, src/Object.hpp:99, src/Object.cpp:248-266): gather what the caller asked for
        bool const want_v = h->vx || h->vy || h->vz, want_a = h->ax || h->ay || h->az;
        bool const want_appe = h->ap || h->pe || h->ap_vel || h->pe_vel;
        NK(nccl().GroupStart());
        for (int r = 0; r < ctx->world; r++) {
            long long a = (long long)ctx->n * r / ctx->world, b = (long long)ctx->n * (r + 1) / ctx->world;
            if (b <= a)
                continue;
            size_t const cnt = (size_t)(b - a);
            for (int k = 0; k < 3; k++) {
                if (want_v)
                    NK(nccl().Broadcast(ctx->vel + k * cap + a, ctx->vel + k * cap + a, cnt, kNcclFloat64, r, ctx->comm, st));
                if (want_a)
                    NK(nccl().Broadcast(ctx->acc + k * cap + a, ctx->acc + k * cap + a, cnt, kNcclFloat64, r, ctx->comm, st));
            }
            if (h->most)
                NK(nccl().Broadcast(ctx->most + a, ctx->most + a, cnt, kNcclUint32, r, ctx->comm, st));
            if (h->max_attraction)
                NK(nccl().Broadcast(ctx->maxattr + a, ctx->maxattr + a, cnt, kNcclFloat64, r, ctx->comm, st));
            if (want_appe)
                for (int k = 0; k < 4; k++)
                    NK(nccl().Broadcast(ctx->appe + k * cap + a, ctx->appe + k * cap + a, cnt, kNcclFloat64, r, ctx->comm, st));
        }
        NK(nccl().GroupEnd());
    }
    int64_t* dates = reinterpret_cast<int64_t*>(d + 15 * N);
    int32_t* most = reinterpret_cast<int32_t*>(dates + 2 * N);
    uint8_t* flags = reinterpret_cast<uint8_t*>(most + N);
    bool const small = ctx->world == 1 && N <= kDownloadSmallMax;
    if (small)
        CK(launch_download_small(ctx, ctx->pinned, st)); // one kernel writes the whole staging area
    bool want_pos = !small && (h->x || h->y || h->z);
    if (want_pos) {
        CK(launch_unpack(ctx, ctx->velh, ctx->velh + cap, ctx->velh + 2 * cap, st));
        for (int k = 0; k < 3; k++)
            CK(cudaMemcpyAsync(d + k * N, ctx->velh + k * cap, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (!small && (h->vx || h->vy || h->vz))
        for (int k = 0; k < 3; k++)
            CK(cudaMemcpyAsync(d + (3 + k) * N, ctx->vel + k * cap, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!small && (h->ax || h->ay || h->az))
        for (int k = 0; k < 3; k++)
            CK(cudaMemcpyAsync(d + (6 + k) * N, ctx->acc + k * cap, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!small && h->gf)
        CK(cudaMemcpyAsync(d + 9 * N, ctx->gf, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!small && h->max_attraction)
        CK(cudaMemcpyAsync(d + 10 * N, ctx->maxattr, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!small && (h->ap || h->pe || h->ap_vel || h->pe_vel))
        for (int k = 0; k < 4; k++)
            CK(cudaMemcpyAsync(d + (11 + k) * N, ctx->appe + k * cap, N * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!small && h->creation_date)
        CK(cudaMemcpyAsync(dates, ctx->cdate, N * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (!small && h->deletion_date)
        CK(cudaMemcpyAsync(dates + N, ctx->ddate, N * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (!small && h->most)
        CK(cudaMemcpyAsync(most, ctx->most, N * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (!small && h->deleted)
        CK(cudaMemcpyAsync(flags, ctx->delflag, N, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    double* dst[15] = { h->x, h->y, h->z, h->vx, h->vy, h->vz, h->ax, h->ay, h->az, h->gf, h->max_attraction, h->ap, h->pe, h->ap_vel,
        h->pe_vel };
    parallel_chunks(N, (size_t)1 << 18, N >= ((size_t)1 << 17) ? host_threads() : 1, [&](size_t lo, size_t hi) {
        for (int k = 0; k < 15; k++)
            if (dst[k])
                memcpy(dst[k] + lo, d + k * N + lo, (hi - lo) * sizeof(double));
    });
    if (h->creation_date)
        memcpy(h->creation_date, dates, N * sizeof(int64_t));
    if (h->deletion_date)
        memcpy(h->deletion_date, dates + N, N * sizeof(int64_t));
    if (h->most)
        memcpy(h->most, most, N * sizeof(int32_t));
    if (h->deleted)
        memcpy(h->deleted, flags, N);
    return ESSA_OK;
}

// Sharded host programs: every rank holds (and moves over its own PCIe link) only the bodies it integrates.
//   essa_upload_shard    positions / velocities of bodies [i0, i1) from the host; the j-records of the other ranks' bodies
//                        arrive over NVLink (one grouped broadcast-gather), gravity factors / dates / masks stay as the last
//                        essa_upload left them
//   essa_download_shard  members of bodies [i0, i1) only, no collective
// (i0, i1) = essa_shard_range.  Host arrays are indexed by GLOBAL body index (length n); only the slice is touched.
extern "C" int essa_upload_shard(essa_ctx* ctx, const essa_bodies* h) {
    ARG(ctx, "ctx is null");
    ARG(h && h->x && h->y && h->z && h->vx && h->vy && h->vz, "x,y,z,vx,vy,vz are required");
    ARG(ctx->n > 0, "essa_upload the whole world once first");
    CK(cudaSetDevice(ctx->device));
    PARK();
    size_t const cap = (size_t)ctx->cap;
    int const i0 = (int)((long long)ctx->n * ctx->rank / ctx->world), i1 = (int)((long long)ctx->n * (ctx->rank + 1) / ctx->world);
    size_t const M = (size_t)(i1 - i0);
    cudaStream_t st = ctx->stream;
    CK(cudaStreamSynchronize(st));
    int rc = reserve_pinned(ctx, 6 * M * sizeof(double));
    if (rc != ESSA_OK)
        return rc;
    double* d = static_cast<double*>(ctx->pinned);
    double const* src[6] = { h->x, h->y, h->z, h->vx, h->vy, h->vz };
    parallel_chunks(M, (size_t)1 << 18, M >= ((size_t)1 << 17) ? host_threads() : 1, [&](size_t lo, size_t hi) {
        for (int k = 0; k < 6; k++)
            memcpy(d + k * M + lo, src[k] + i0 + lo, (hi - lo) * sizeof(double));
    });
    for (int k = 0; k < 3; k++) {
        CK(cudaMemcpyAsync(ctx->velh + k * cap + i0, d + k * M, M * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->vel + k * cap + i0, d + (3 + k) * M, M * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    CK(launch_pack_range(ctx, ctx->velh, ctx->velh + cap, ctx->velh + 2 * cap, i0, i1, st));
    if (ctx->world > 1) { // the other ranks' records
        double4* buf = ctx->pos[ctx->cur];
        NK(nccl().GroupStart());
        for (int r = 0; r < ctx->world; r++) {
            long long a = (long long)ctx->n * r / ctx->world, b = (long long)ctx->n * (r + 1) / ctx->world;
            if (b > a)
                NK(nccl().Broadcast(buf + a, buf + a, (size_t)(b - a) * 4, kNcclFloat64, r, ctx->comm, st));
        }
        NK(nccl().GroupEnd());
    }
    ctx->acc_valid = false;
    ctx->acc_has_most = false;
    CK(cudaStreamSynchronize(st));
    return ESSA_OK;
}

extern "C" int essa_download_shard(essa_ctx* ctx, essa_bodies* h) {
    ARG(ctx, "ctx is null");
    ARG(h, "host view is null");
    CK(cudaSetDevice(ctx->device));
    PARK();
    size_t const cap = (size_t)ctx->cap;
    int const i0 = (int)((long long)ctx->n * ctx->rank / ctx->world), i1 = (int)((long long)ctx->n * (ctx->rank + 1) / ctx->world);
    size_t const M = (size_t)(i1 - i0);
    if (M == 0)
        return ESSA_OK;
    int rc = reserve_pinned(ctx, 15 * M * sizeof(double) + M * sizeof(int32_t));
    if (rc != ESSA_OK)
        return rc;
    cudaStream_t st = ctx->stream;
    double* d = static_cast<double*>(ctx->pinned);
    int32_t* most = reinterpret_cast<int32_t*>(d + 15 * M);
    if (h->x || h->y || h->z) {
        CK(launch_unpack(ctx, ctx->velh, ctx->velh + cap, ctx->velh + 2 * cap, st));
        for (int k = 0; k < 3; k++)
            CK(cudaMemcpyAsync(d + k * M, ctx->velh + k * cap + i0, M * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    double const* dev[15] = { nullptr, nullptr, nullptr, ctx->vel, ctx->vel + cap, ctx->vel + 2 * cap, ctx->acc, ctx->acc + cap, ctx->acc + 2 * cap,
        ctx->gf, ctx->maxattr, ctx->appe, ctx->appe + cap, ctx->appe + 2 * cap, ctx->appe + 3 * cap };
    double* dst[15] = { h->x, h->y, h->z, h->vx, h->vy, h->vz, h->ax, h->ay, h->az, h->gf, h->max_attraction, h->ap, h->pe, h->ap_vel,
        h->pe_vel };
    for (int k = 3; k < 15; k++)
        if (dst[k])
            CK(cudaMemcpyAsync(d + k * M, dev[k] + i0, M * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (h->most)
        CK(cudaMemcpyAsync(most, ctx->most + i0, M * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    parallel_chunks(M, (size_t)1 << 18, M >= ((size_t)1 << 17) ? host_threads() : 1, [&](size_t lo, size_t hi) {
        for (int k = 0; k < 15; k++)
            if (dst[k])
                memcpy(dst[k] + i0 + lo, d + k * M + lo, (hi - lo) * sizeof(double));
    });
    if (h->most)
        memcpy(h->most + i0, most, M * sizeof(int32_t));
    return ESSA_OK;
}

// ---------------------------------------------------------------------------------------------
// stepping
// ---------------------------------------------------------------------------------------------
static cudaEvent_t force_event(essa_ctx* c, size_t idx) {
    while (c->force_events.size() <= idx) {
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) != cudaSuccess)
            return nullptr;
        c->force_events.push_back(ev);
    }
    return c->force_events[idx];
}

// ---------------------------------------------------------------------------------------------
// Sharded pair-shared pass: peer-mapped slot buffers.
// Tile (I, d) of the pair kernel produces J-side partial accelerations for block J = I + d, which another rank may own.
// Instead of folding them locally and reduce-scattering 3 x N doubles with NCCL after the kernel, every rank maps every other
// rank's slot buffer into its address space once (cudaIpcGetMemHandle / cudaIpcOpenMemHandle, the handles travel through an
// ncclAllGather on the context's own communicator) and the kernel stores each tile's J-side sums directly into the owner's
// memory over NVLink while it computes.  What is left per pass is a one-word all-reduce as a barrier ("every rank's kernel
// has finished, its stores have landed") before k_pair_finish reads the slots -- and the owner adds a body's J-side slots in
// the same fixed order (ascending d) as a single GPU does.  The I-side partial sums are still cut in as many chunks as the
// plan of the rank count at hand chooses, so results on different rank counts agree to rounding (measured at N = 1,048,576
// against the oracle's extended-precision rows: 4.4e-15 on 1 GPU, 2.2e-15 on 2, 6.2e-15 on 8), not bit for bit.
// ---------------------------------------------------------------------------------------------
static void pair_peer_close(essa_ctx* ctx) {
    for (int r = 0; r < essa_ctx::kMaxPeers; r++) {
        if (r != ctx->rank) {
            for (int k = 0; k < 2; k++)
                if (ctx->pair_pos_peer[k][r])
                    cudaIpcCloseMemHandle(ctx->pair_pos_peer[k][r]);
            if (ctx->pair_flags_peer[r])
                cudaIpcCloseMemHandle(ctx->pair_flags_peer[r]);
            if (ctx->pair_slot_peer[r])
                cudaIpcCloseMemHandle(ctx->pair_slot_peer[r]);
            if (ctx->pair_kslot_peer[r])
                cudaIpcCloseMemHandle(ctx->pair_kslot_peer[r]);
        }
        ctx->pair_slot_peer[r] = nullptr;
        ctx->pair_kslot_peer[r] = nullptr;
        ctx->pair_pos_peer[0][r] = ctx->pair_pos_peer[1][r] = nullptr;
        ctx->pair_flags_peer[r] = nullptr;
    }
    ctx->pair_peer_ok = false;
    ctx->pair_pospeer_ok = false;
}

// Barrier over the peer mappings, without NCCL: every rank announces an epoch in each peer's flag array (one word per sender),
// then waits until its own array shows that epoch from everybody.  Stream-ordered behind the kernel whose results it announces.
__global__ void k_peer_wait(unsigned long long const* flags, int world, unsigned long long epoch) {
    int const r = threadIdx.x;
    if (r < world) {
        long long const t0 = clock64();
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
        } while (v < epoch && clock64() - t0 < 20000000000ll); // (~10 s: a rank that died must not hang the others' GPUs)
    }
}
struct PeerFlagTable {
    unsigned long long* p[essa_ctx::kMaxPeers];
};
__global__ void k_peer_signal_t(PeerFlagTable const T, int rank, int world, unsigned long long epoch) {
    int const r = threadIdx.x;
    if (r < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(T.p[r] + rank), "l"(epoch) : "memory");
    }
}
static int peer_flag_barrier(essa_ctx* ctx) {
    PeerFlagTable T;
    for (int r = 0; r < essa_ctx::kMaxPeers; r++)
        T.p[r] = ctx->pair_flags_peer[r];
    unsigned long long const epoch = ++ctx->pair_epoch;
    k_peer_signal_t<<<1, 32, 0, ctx->stream>>>(T, ctx->rank, ctx->world, epoch);
    k_peer_wait<<<1, 32, 0, ctx->stream>>>(ctx->pair_flags, ctx->world, epoch);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return ESSA_OK;
}

// every rank's kernels issued so far on the context's stream have finished (device-side: one word through NCCL)
static int shard_barrier(essa_ctx* ctx) {
    if (!ctx->peer_xchg)
        CK(cudaMalloc(&ctx->peer_xchg, (size_t)essa_ctx::kMaxPeers * kXchgWords * sizeof(unsigned long long)));
    NK(nccl().AllReduce(ctx->peer_xchg, ctx->peer_xchg, 1, kNcclUint32, kNcclSum, ctx->comm, ctx->stream));
    return ESSA_OK;
}

// Collective over the ranks of the context (all of them run the same plan, so all of them get here together).
static int pair_peer_setup(essa_ctx* ctx, bool keys) {
    if (ctx->world <= 1 || ctx->pair_peer_failed)
        return ESSA_OK;
    static bool const disabled = [] {
        char const* e = getenv("ESSA_PAIR_PEER"); // diagnostics: 0 = NCCL reduce-scatter instead of peer stores
        return e && atoi(e) == 0;
    }();
    if (disabled || ctx->world > essa_ctx::kMaxPeers || !nccl().AllGather) {
        ctx->pair_peer_failed = true;
        return ESSA_OK;
    }
    bool const will_change = pair_reserve_would_change(ctx, keys) || ctx->pos[0] != ctx->pair_pos_exported[0]
        || ctx->pos[1] != ctx->pair_pos_exported[1]; // (a grown world has new position buffers: on every rank, uploads are collective)
    if (ctx->pair_peer_ok && !will_change)
        return ESSA_OK;
    if (ctx->pair_peer_ok) { // the buffers the peers have mapped are about to be freed: everybody lets go first
        pair_peer_close(ctx);
        int rc = shard_barrier(ctx);
        if (rc != ESSA_OK)
            return rc;
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(pair_reserve(ctx, keys, nullptr));
    if (!ctx->peer_xchg)
        CK(cudaMalloc(&ctx->peer_xchg, (size_t)essa_ctx::kMaxPeers * kXchgWords * sizeof(unsigned long long)));
    // 64 words per rank: [0] ok flag, [1] has keys, [2] position buffers exported, [8..16) handle of the slots, [16..24) handle of
    // the key slots, [24..32) / [32..40) handles of the two position buffers
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    unsigned long long mine[kXchgWords] = {};
    cudaIpcMemHandle_t hs, hk, hp[2];
    if (!ctx->pair_flags) { // (never freed before the context: the epochs only grow)
        CK(cudaMalloc(&ctx->pair_flags, (size_t)2 << 20)); // a whole 2 MiB page of its own: what an IPC handle exposes
        CK(cudaMemsetAsync(ctx->pair_flags, 0, essa_ctx::kMaxPeers * sizeof(unsigned long long), ctx->stream));
    }
    cudaIpcMemHandle_t hf;
    bool const pos_ok = cudaIpcGetMemHandle(&hp[0], ctx->pos[0]) == cudaSuccess && cudaIpcGetMemHandle(&hp[1], ctx->pos[1]) == cudaSuccess
        && cudaIpcGetMemHandle(&hf, ctx->pair_flags) == cudaSuccess;
    mine[2] = pos_ok ? 1 : 0;
    if (pos_ok) {
        memcpy(&mine[24], &hp[0], 64);
        memcpy(&mine[32], &hp[1], 64);
        memcpy(&mine[40], &hf, 64);
    }
    bool ok = cudaIpcGetMemHandle(&hs, ctx->pair_slots) == cudaSuccess;
    bool const have_k = ctx->pair_kslots != nullptr;
    if (ok && have_k)
        ok = cudaIpcGetMemHandle(&hk, ctx->pair_kslots) == cudaSuccess;
    (void)cudaGetLastError();
    mine[0] = ok ? 1 : 0;
    mine[1] = have_k ? 1 : 0;
    if (ok) {
        memcpy(&mine[8], &hs, 64);
        if (have_k)
            memcpy(&mine[16], &hk, 64);
    }
    CK(cudaMemcpyAsync(ctx->peer_xchg + (size_t)ctx->rank * kXchgWords, mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
    NK(nccl().AllGather(ctx->peer_xchg + (size_t)ctx->rank * kXchgWords, ctx->peer_xchg, kXchgWords, kNcclUint64, ctx->comm, ctx->stream));
    std::vector<unsigned long long> all((size_t)ctx->world * kXchgWords);
    CK(cudaMemcpyAsync(all.data(), ctx->peer_xchg, all.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    bool all_ok = true, all_pos = true;
    for (int r = 0; r < ctx->world; r++) {
        all_ok = all_ok && all[(size_t)r * kXchgWords] == 1 && all[(size_t)r * kXchgWords + 1] == (have_k ? 1u : 0u);
        all_pos = all_pos && all[(size_t)r * kXchgWords + 2] == 1;
    }
    if (all_ok) {
        for (int r = 0; r < ctx->world && all_ok; r++) {
            if (r == ctx->rank) {
                ctx->pair_slot_peer[r] = ctx->pair_slots;
                ctx->pair_kslot_peer[r] = ctx->pair_kslots;
                continue;
            }
            cudaIpcMemHandle_t h;
            void* ptr = nullptr;
            memcpy(&h, &all[(size_t)r * kXchgWords + 8], 64);
            all_ok = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            ctx->pair_slot_peer[r] = static_cast<double*>(ptr);
            if (all_ok && have_k) {
                memcpy(&h, &all[(size_t)r * kXchgWords + 16], 64);
                ptr = nullptr;
                all_ok = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
                ctx->pair_kslot_peer[r] = static_cast<unsigned long long*>(ptr);
            }
            for (int k = 0; k < 2 && all_ok && all_pos; k++) {
                memcpy(&h, &all[(size_t)r * kXchgWords + 24 + 8 * k], 64);
                ptr = nullptr;
                all_pos = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
                ctx->pair_pos_peer[k][r] = static_cast<double4*>(ptr);
            }
            if (all_ok && all_pos) {
                memcpy(&h, &all[(size_t)r * kXchgWords + 40], 64);
                ptr = nullptr;
                all_pos = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
                ctx->pair_flags_peer[r] = static_cast<unsigned long long*>(ptr);
            }
        }
        (void)cudaGetLastError();
    }
    if (all_ok && all_pos) {
        ctx->pair_pos_peer[0][ctx->rank] = ctx->pos[0];
        ctx->pair_pos_peer[1][ctx->rank] = ctx->pos[1];
        ctx->pair_flags_peer[ctx->rank] = ctx->pair_flags;
    }
    // whatever came of it, the peers were handed these two buffers: they must outlive a grown world (free_world)
    ctx->pair_pos_exported[0] = ctx->pos[0];
    ctx->pair_pos_exported[1] = ctx->pos[1];
    // the decision must be the same everywhere: one more word around (low half: slots, high half: positions too)
    unsigned long long flag = (all_ok ? 1ull : 0ull) | (all_ok && all_pos ? 1ull << 32 : 0ull);
    CK(cudaMemcpyAsync(ctx->peer_xchg, &flag, sizeof(flag), cudaMemcpyHostToDevice, ctx->stream));
    NK(nccl().AllReduce(ctx->peer_xchg, ctx->peer_xchg, 1, kNcclUint64, kNcclSum, ctx->comm, ctx->stream));
    CK(cudaMemcpyAsync(&flag, ctx->peer_xchg, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if ((flag & 0xffffffffull) == (unsigned long long)ctx->world) {
        ctx->pair_peer_ok = true;
        static bool const pos_allowed = [] {
            char const* e = getenv("ESSA_PAIR_PEER_POS"); // diagnostics: 0 = wait for the gathered positions instead
            return !(e && atoi(e) == 0);
        }();
        ctx->pair_pospeer_ok = pos_allowed && (flag >> 32) == (unsigned long long)ctx->world;
    } else {
        pair_peer_close(ctx);
        ctx->pair_peer_failed = true;
    }
    return ESSA_OK;
}

// In a sharded context every rank has just written its own targets' records into the `next`
// position buffer: gather the other ranks' shards (ragged shards -> one broadcast per rank,
// grouped).  Runs on the comm stream so that the force pass can start on the local shard.
static int gather_issue(essa_ctx* ctx);
static int gather_next(essa_ctx* ctx) {
    if (ctx->world == 1)
        return ESSA_OK;
    CK(cudaEventRecord(ctx->ev_ready, ctx->stream));
    ctx->gather_buf = ctx->pos[ctx->cur ^ 1];
    ctx->gather_deferred = true;
    ctx->gather_pending = true;
    // Issued at once.  (Measured and rejected: issuing the NCCL calls only behind the launch of a force kernel that reads its
    // remote tiles from their owners, so that the host's ~0.1 ms in the group call does not hold that launch back -- the
    // gather's own kernel then finds every SM taken and runs at the force kernel's tail, in front of the barrier that follows
    // it: 16,384 bodies on 2 GPUs 156 -> 166 us per step.)
    return gather_issue(ctx);
}
static int gather_issue(essa_ctx* ctx) {
    if (!ctx->gather_deferred)
        return ESSA_OK;
    ctx->gather_deferred = false;
    CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_ready, 0));
    double4* buf = ctx->gather_buf;
    NK(nccl().GroupStart());
    for (int r = 0; r < ctx->world; r++) {
        long long a = (long long)ctx->n * r / ctx->world, b = (long long)ctx->n * (r + 1) / ctx->world;
        if (b > a)
            NK(nccl().Broadcast(buf + a, buf + a, (size_t)(b - a) * 4, kNcclFloat64, r, ctx->comm, ctx->comm_stream));
    }
    NK(nccl().GroupEnd());
    CK(cudaEventRecord(ctx->ev_gather, ctx->comm_stream));
    return ESSA_OK;
}
// ... and the wait for it on the context's stream
static int gather_wait(essa_ctx* ctx) {
    if (!ctx->gather_pending)
        return ESSA_OK;
    int rc = gather_issue(ctx);
    if (rc != ESSA_OK)
        return rc;
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_gather, 0));
    ctx->gather_pending = false;
    return ESSA_OK;
}

// One force pass over pos[cur] for this context's targets.  Sharded: local-shard sources first,
// the other shards once the gather of this buffer has landed.
static int force_pass(essa_ctx* ctx, bool argmax, bool save_old, KickArgs const& k, double eps2, size_t* ev_idx) {
    int n = ctx->n;
    int R;
    int iblocks = force_iblocks(ctx, &R);
    // ev_idx == nullptr: a call of many steps brackets ALL its passes with one pair of events (step_tiled) -- two timestamp
    // records around every pass keep back-to-back launches from overlapping their ramp-up, microseconds that matter at mid sizes
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ev_idx) {
        e0 = force_event(ctx, (*ev_idx)++);
        e1 = force_event(ctx, (*ev_idx)++);
        if (!e0 || !e1) {
            ctx->err = "cudaEventCreate failed";
            return ESSA_ERR_CUDA;
        }
        CK(cudaEventRecord(e0, ctx->stream));
    }
    if (ctx->use_paired) {
        // most-attracting tracking: real candidate keys, unless every gravity factor is equal (then it is a no-op)
        bool const keys = argmax && !ctx->gf_uniform;
        if (ctx->world > 1) {
            int prc = pair_peer_setup(ctx, keys);
            if (prc != ESSA_OK)
                return prc;
        }
        // The pair kernel walks the whole ring of blocks.  With peer-mapped position buffers it reads a remote block from the
        // rank that owns it, so all it needs is that every rank's previous kernel (the one that wrote its shard of these
        // positions) is through: a one-word barrier instead of the wait for the gathered copy, which lands in the background
        // and is waited for only in front of k_pair_finish (whose rare repair path reads all positions from local memory).
        bool const peer_pos = ctx->world > 1 && ctx->pair_peer_ok && ctx->pair_pospeer_ok && !keys;
        bool gather_late = false;
        if (ctx->gather_pending) {
            if (peer_pos) {
                int brc = peer_flag_barrier(ctx);
                if (brc != ESSA_OK)
                    return brc;
                gather_late = true;
            } else {
                int grc = gather_wait(ctx);
                if (grc != ESSA_OK)
                    return grc;
            }
        }
        ctx->pair_peer_pos_pass = gather_late; // without a pending gather the local copy is complete: nothing to fetch from peers
        double* acc = nullptr;
        unsigned long long* kacc = nullptr;
        int npad = 0;
        if (keys && ctx->world > 1 && ctx->pair_prior_valid) {
            // every rank's k_pair_finish rewrote the filter thresholds of its own targets: all ranks need all of them
            size_t cnt = (size_t)n / ctx->world;
            NK(nccl().GroupStart());
            for (int r = 0; r < ctx->world; r++)
                NK(nccl().Broadcast(ctx->pair_prior + r * cnt, ctx->pair_prior + r * cnt, cnt, kNcclUint32, r, ctx->comm, ctx->stream));
            NK(nccl().GroupEnd());
        }
        // a pass timed by itself: also where its time goes (three more event records; not for small worlds, whose steps are
        // bounded by how fast the host enqueues work: 16,384 bodies 210 -> 217 us per step with them)
        bool const phases = e0 != nullptr && ctx->n >= 65536;
        if (phases) {
            for (auto& pe : ctx->ev_phase)
                if (!pe)
                    CK(cudaEventCreate(&pe));
            CK(cudaEventRecord(ctx->ev_phase[0], ctx->stream));
        }
        CK(launch_force_paired_compute(ctx, eps2, ctx->stream, &acc, &npad, keys, &kacc));
        if (phases)
            CK(cudaEventRecord(ctx->ev_phase[1], ctx->stream));
        if (ctx->world > 1 && ctx->pair_peer_ok) {
            // the J-side partials went straight into their owners' slot buffers: wait until every rank's kernel is through
            int brc = shard_barrier(ctx);
            if (brc != ESSA_OK)
                return brc;
            acc = nullptr;
        } else if (ctx->world > 1) {
            // every rank holds J-side sums (and candidate keys) for every body: reduce-scatter them onto the owners (in place)
            size_t cnt = (size_t)n / ctx->world;
            if (keys && ctx->pair_kall_n < (size_t)n) {
                cudaFree(ctx->pair_kall);
                ctx->pair_kall = nullptr;
                ctx->pair_kall_n = 0;
                CK(cudaMalloc(&ctx->pair_kall, (size_t)n * sizeof(unsigned long long)));
                ctx->pair_kall_n = (size_t)n;
            }
            NK(nccl().GroupStart());
            for (int c3 = 0; c3 < 3; c3++) {
                double* base = acc + (size_t)c3 * npad;
                NK(nccl().ReduceScatter(base, base + (size_t)ctx->rank * cnt, cnt, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream));
            }
            if (keys) {
                // every rank's candidate for each of my bodies (not an ncclMax: two keys that differ in the index only are
                // re-decided exactly by k_pair_finish, which therefore has to see both)
                for (int r = 0; r < ctx->world; r++) {
                    NK(nccl().Send(kacc + (size_t)r * cnt, cnt, kNcclUint64, r, ctx->comm, ctx->stream));
                    NK(nccl().Recv(ctx->pair_kall + (size_t)r * cnt, cnt, kNcclUint64, r, ctx->comm, ctx->stream));
                }
            }
            NK(nccl().GroupEnd());
        }
        if (gather_late) { // the gather's NCCL calls go out behind the force kernel's launch; the wait is a formality by then
            int grc = gather_wait(ctx);
            if (grc != ESSA_OK)
                return grc;
        }
        if (phases) {
            CK(cudaEventRecord(ctx->ev_phase[2], ctx->stream));
            ctx->ev_phase_first = e0;
            ctx->ev_phase_last = e1;
            ctx->phases_recorded = true;
        }
        CK(launch_force_paired_finish(ctx, k, ctx->stream, acc, argmax && !keys, save_old, keys,
            keys && ctx->world > 1 && !ctx->pair_peer_ok ? ctx->pair_kall : nullptr, eps2));
    } else if (ctx->world == 1 || !ctx->gather_pending) {
        int S;
        choose_split(ctx, iblocks, (n + kChunkQ - 1) / kChunkQ, &S);
        CK(launch_force_pass(ctx, 0, n, S, 0, S, argmax, save_old, k, eps2, ctx->stream));
    } else {
        int i0 = (int)((long long)n * ctx->rank / ctx->world), i1 = (int)((long long)n * (ctx->rank + 1) / ctx->world);
        int SA, SB0 = 0, SB1 = 0;
        choose_split(ctx, iblocks, (i1 - i0 + kChunkQ - 1) / kChunkQ, &SA);
        if (i0 > 0)
            choose_split(ctx, iblocks, (i0 + kChunkQ - 1) / kChunkQ, &SB0);
        if (i1 < n)
            choose_split(ctx, iblocks, (n - i1 + kChunkQ - 1) / kChunkQ, &SB1);
        int total = SA + SB0 + SB1;
        // slot order follows source index so that the ordered reduction is the same on every rank count
        CK(launch_force_pass(ctx, i0, i1, SA, SB0, total, argmax, save_old, k, eps2, ctx->stream));
        {
            int grc = gather_wait(ctx); // (the NCCL calls go out behind the local-shard launch)
            if (grc != ESSA_OK)
                return grc;
        }
        if (SB0)
            CK(launch_force_pass(ctx, 0, i0, SB0, 0, total, argmax, save_old, k, eps2, ctx->stream));
        if (SB1)
            CK(launch_force_pass(ctx, i1, n, SB1, SB0 + SA, total, argmax, save_old, k, eps2, ctx->stream));
    }
    if (e1)
        CK(cudaEventRecord(e1, ctx->stream));
    int i0 = (int)((long long)n * ctx->rank / ctx->world), i1 = (int)((long long)n * (ctx->rank + 1) / ctx->world);
    ctx->passes++;
    ctx->interactions += (double)(i1 - i0) * (double)(n - 1);
    return ESSA_OK;
}

constexpr int kPairedAutoMin = 12288; // AUTO takes the pair-shared kernel above this many bodies

// K1g: a mid-size world's whole call in one cooperative launch (force_grid.cu).
static int step_grid(essa_ctx* ctx, int steps, essa_step_opts const& o, bool* not_placed) {
    int const K = std::abs(steps);
    *not_placed = false;
    bool const argmax = o.track_most_attracting != 0;
    bool const first_pass = !ctx->acc_valid || (argmax && !ctx->acc_has_most);
    cudaEvent_t e0 = force_event(ctx, 0), e1 = force_event(ctx, 1);
    if (!e0 || !e1) {
        ctx->err = "cudaEventCreate failed";
        return ESSA_ERR_CUDA;
    }
    CK(cudaEventRecord(e0, ctx->stream));
    {
        // the grid must be co-resident (one CTA per SM): another context's persistent kernel on this device, or a partitioned
        // GPU, can make that impossible -- then the tiled kernel takes the call
        cudaError_t const le = launch_force_grid(ctx, steps, o, first_pass, argmax, ctx->stream);
        if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorLaunchOutOfResources) {
            (void)cudaGetLastError();
            *not_placed = true;
            return ESSA_OK;
        }
        CK(le);
    }
    if (!o.forward_simulated)
        ctx->date += (int64_t)steps * o.dt_seconds; // no creation / deletion instant lies in between (the caller checked)
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventRecord(ctx->ev_end, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->grid_timed_out && *ctx->grid_timed_out) { // a CTA waited two seconds for the others: the state is not a step's result
        *ctx->grid_timed_out = 0u;
        ctx->acc_valid = false;
        ctx->err = "grid kernel: a grid barrier timed out";
        return ESSA_ERR_CUDA;
    }
    ctx->cur ^= K & 1;
    ctx->acc_valid = true;
    ctx->acc_has_most = argmax;
    ctx->acc_exact = false;
    int const passes = K + (first_pass ? 1 : 0);
    ctx->passes += passes;
    ctx->interactions += (double)passes * (double)ctx->n * (double)(ctx->n - 1);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ctx->last_force_ms = ms;
    ctx->last_force_passes = passes;
    ctx->last_pre_ms = ctx->last_post_ms = 0;
    return ESSA_OK;
}

static int step_tiled(essa_ctx* ctx, int steps, essa_step_opts const& o, bool forces_only) {
    int const K = forces_only ? 0 : std::abs(steps);
    double const mul = steps >= 0 ? 1.0 : -1.0;
    bool const argmax = o.track_most_attracting || o.record;
    if ((o.kernel == ESSA_KERNEL_AUTO || o.kernel == ESSA_KERNEL_GRID) && force_grid_applicable(ctx, o, forces_only)
        && (o.forward_simulated || !mask_changes(ctx, ctx->date, ctx->date + (int64_t)steps * o.dt_seconds))) {
        bool not_placed = false;
        int const rc = step_grid(ctx, steps, o, &not_placed);
        if (rc != ESSA_OK || !not_placed)
            return rc;
        if (o.kernel == ESSA_KERNEL_GRID) {
            ctx->err = "grid kernel: the device cannot hold one CTA per SM at once (another resident kernel is running)";
            return ESSA_ERR_CUDA;
        }
    } else if (o.kernel == ESSA_KERNEL_GRID) {
        ctx->err = "grid kernel: 1,025 .. 5,120 bodies, unsharded, no recording, no two_pass, no creation / deletion instant inside the call";
        return ESSA_ERR_ARG;
    }
    {
        // pair-shared kernel: slots must fit comfortably.  Most-attracting tracking rides along as sortable candidate
        // keys (worlds of up to 2^21 bodies), or costs nothing where it is provably a no-op (all gravity factors
        // equal: the "heavier partner" test of Object.cpp:84-94 fails for every pair, so max_attraction stays 0 and
        // the links stay as they are -- k_pair_finish writes exactly that)
        int nb, S;
        size_t slot_d = 0, part_d = 0;
        bool can = (!argmax || ctx->gf_uniform || ctx->n <= (1 << 21)) && paired_plan(ctx, &nb, &S, &slot_d, &part_d);
        if (can && ctx->plan.batches > 1 && argmax && !ctx->gf_uniform)
            can = false; // candidate keys are not batched (with the default budget only worlds beyond 2^21 bodies batch at all)
        if (can && (slot_d > ctx->pair_slots_n || part_d > ctx->pair_part_n)) { // first use at this size: will it fit?
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            can = (slot_d + part_d) * sizeof(double) < free_b / 2 + (ctx->pair_slots_n + ctx->pair_part_n) * sizeof(double);
        }
        if (o.kernel == ESSA_KERNEL_PAIRED && !can) {
            ctx->err = "pair-shared kernel needs >= 1025 bodies (whole 1024- or 2048-body blocks per rank when sharded), at most 2^21 bodies when the most-attracting links are tracked, and room for its slots";
            return ESSA_ERR_ARG;
        }
        // AUTO: measured crossover on B200 (tools/crossover.py): 12,288 bodies 185 (tiled) vs 190 us per step (pair-shared),
        // 16,383 bodies 313 vs 192
        ctx->use_paired = can && (o.kernel == ESSA_KERNEL_PAIRED || (o.kernel == ESSA_KERNEL_AUTO && ctx->n > kPairedAutoMin));
    }
    KickArgs base;
    base.dt = (double)o.dt_seconds;
    base.h = base.dt / 2.0;
    base.mul = mul;
    base.second_kick = 0;
    base.advance = 0;
    size_t ev = 0;
    // calls of more than two steps time their passes as one span (see force_pass)
    bool const span = K > 2;
    size_t* const evp = span ? nullptr : &ev;
    int span_passes = 0;
    if (span) {
        cudaEvent_t e0 = force_event(ctx, 0), e1 = force_event(ctx, 1);
        if (!e0 || !e1) {
            ctx->err = "cudaEventCreate failed";
            return ESSA_ERR_CUDA;
        }
        CK(cudaEventRecord(e0, ctx->stream));
    }
    if (ctx->gather_pending) { // never left pending between calls, defensive
        int grc = gather_wait(ctx);
        if (grc != ESSA_OK)
            return grc;
    }
    if (forces_only) {
        int rc = force_pass(ctx, argmax, false, base, o.eps2, &ev);
        if (rc != ESSA_OK)
            return rc;
        ctx->acc_valid = true;
        ctx->acc_has_most = argmax;
        ctx->acc_exact = false;
    }
    bool fused_prev = false;
    for (int k = 0; k < K; k++) {
        // update_history_and_date (src/World.cpp:59-84): the date moves before anything else
        if (!o.forward_simulated) {
            int64_t nd = ctx->date + (int64_t)(mul > 0 ? 1 : -1) * o.dt_seconds;
            bool flip = mask_changes(ctx, ctx->date, nd);
            ctx->date = nd;
            if (flip) {
                CK(launch_refresh_active(ctx, ctx->stream));
                ctx->pair_prior_valid = false;
                ctx->acc_valid = false;
            }
        }
        bool need_save_old = true;
        if (!fused_prev) {
            if (!ctx->acc_valid || o.two_pass || (argmax && !ctx->acc_has_most)) {
                KickArgs ka = base; // first set_forces() of the step: store only
                int rc = force_pass(ctx, argmax, need_save_old, ka, o.eps2, evp);
                span_passes++;
                if (rc != ESSA_OK)
                    return rc;
                need_save_old = false;
            }
            KickArgs kd = base;
            kd.advance = 1;
            CK(launch_kick_drift(ctx, kd, ctx->stream));
            int rc = gather_next(ctx);
            if (rc != ESSA_OK)
                return rc;
        }
        ctx->cur ^= 1;
        bool advance = false;
        if (k + 1 < K && !o.two_pass) {
            int64_t nd = ctx->date + (int64_t)(mul > 0 ? 1 : -1) * o.dt_seconds;
            advance = o.forward_simulated || !mask_changes(ctx, ctx->date, nd);
        }
        KickArgs ks = base;
        ks.second_kick = 1;
        ks.advance = advance ? 1 : 0;
        int rc = force_pass(ctx, argmax, need_save_old, ks, o.eps2, evp);
        span_passes++;
        if (rc != ESSA_OK)
            return rc;
        ctx->acc_valid = true;
        ctx->acc_has_most = argmax;
        ctx->acc_exact = false;
        if (advance) {
            rc = gather_next(ctx);
            if (rc != ESSA_OK)
                return rc;
        }
        fused_prev = advance;
        if (o.record && ctx->tracked > 0)
            CK(launch_record(ctx, o.offset_trails, o.forward_simulated, ctx->stream));
    }
    if (span) {
        CK(cudaEventRecord(force_event(ctx, 1), ctx->stream));
        ev = 2;
    }
    CK(cudaEventRecord(ctx->ev_end, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0, total = 0;
    for (size_t i = 0; i + 1 < ev; i += 2) {
        CK(cudaEventElapsedTime(&ms, ctx->force_events[i], ctx->force_events[i + 1]));
        total += ms;
    }
    ctx->last_force_ms = total;
    if (ctx->phases_recorded) {
        cudaEvent_t const seq[5] = { ctx->ev_phase_first, ctx->ev_phase[0], ctx->ev_phase[1], ctx->ev_phase[2], ctx->ev_phase_last };
        for (int i = 0; i < 4; i++) {
            CK(cudaEventElapsedTime(&ms, seq[i], seq[i + 1]));
            ctx->last_phase_ms[i] = ms;
        }
        ctx->phases_recorded = false;
    }
    ctx->last_force_passes = span ? span_passes : (int)(ev / 2);
    ctx->last_pre_ms = ctx->last_post_ms = 0;
    if (ev >= 2) {
        CK(cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->force_events[0]));
        ctx->last_pre_ms = ms;
        CK(cudaEventElapsedTime(&ms, ctx->force_events[ev - 1], ctx->ev_end));
        ctx->last_post_ms = ms;
    }
    return ESSA_OK;
}

static int resident_finish(essa_ctx* ctx, int steps, essa_step_opts const& o, bool forces_only) {
    if (!forces_only && !o.forward_simulated)
        ctx->date += (int64_t)steps * o.dt_seconds;
    ctx->acc_valid = true;
    ctx->acc_has_most = o.track_most_attracting || o.record || o.exact_order;
    ctx->acc_exact = o.exact_order != 0;
    ctx->last_force_ms = 0;
    return ESSA_OK;
}

constexpr int kResidentAutoMax = 320;
constexpr int kGridAutoMin = 400; // AUTO prefers K1g to the resident cluster kernels from here on (calls without recording)

static int step_common(essa_ctx* ctx, int steps, const essa_step_opts* opts, bool forces_only, bool async = false) {
    ARG(ctx, "ctx is null");
    ARG(opts, "opts is null");
    ARG(forces_only || steps != 0, "steps must not be 0 (assert in src/World.cpp:87)");
    CK(cudaSetDevice(ctx->device));
    essa_step_opts o = *opts;
    if (o.record)
        o.track_most_attracting = 1;
    if (ctx->persist.live) {
        // the kernel of the previous call is still resident: same kind of tick, no date at which Object::deleted() flips -> doorbell
        if (!forces_only && persist_takes(ctx, steps, o)
            && (o.forward_simulated || !mask_changes(ctx, ctx->date, ctx->date + (int64_t)steps * o.dt_seconds))) {
            int prc = persist_wait_ack(ctx); // (a tick nobody has read back yet)
            if (prc != ESSA_OK)
                return prc;
            if (ctx->persist.live) {
                persist_post(ctx, steps);
                int const k = steps < 0 ? -steps : steps;
                ctx->passes += o.two_pass ? 2 * k : k;
                ctx->interactions += (double)(o.two_pass ? 2 * k : k) * (double)ctx->n * (double)(ctx->n - 1);
                ctx->last_force_passes = o.two_pass ? 2 * k : k;
                if (!async && (prc = persist_wait_ack(ctx)) != ESSA_OK)
                    return prc;
                return resident_finish(ctx, steps, o, false);
            }
        }
        PARK();
    }
    settle_timing(ctx); // the events are about to be reused
    if (o.eps2 != ctx->acc_eps2) { // accelerations (and filter thresholds) left by a call with another softening are not reusable
        ctx->acc_valid = false;
        ctx->pair_prior_valid = false;
        ctx->acc_eps2 = o.eps2;
    }
    if (ctx->n == 0) {
        if (!forces_only && !opts->forward_simulated)
            ctx->date += (int64_t)steps * opts->dt_seconds;
        return ESSA_OK;
    }
    // AUTO: one resident launch for all steps up to ESSA_RESIDENT_MAX bodies.  Plain fast-arithmetic calls run on a
    // 16-SM cluster (K3x / K3xw: 0.7 .. 20 us per step, always below the tiled path's ~24 us of launches per step);
    // what the cluster kernels do not take (a mask that flips during the call, two_pass) goes to the single-CTA K3
    // only up to ~320 bodies -- beyond that one SM takes longer per step than the tiled launches do (measured on
    // B200: 384 bodies 27.7 vs 23.7 us, 1024 bodies 176 vs 24 us per step).
    bool const needs_resident = o.exact_order || (o.forward_simulated && ctx->closest && ctx->closest_n == ctx->n);
    bool const wide_ok = ctx->world == 1 && ctx->n > kResidentAutoMax && resident_cluster_wide_applicable(ctx, o, forces_only)
        && (o.forward_simulated || !mask_changes(ctx, ctx->date, ctx->date + (int64_t)steps * o.dt_seconds));
    // ... except plain calls (no recording) on 400 .. 1,024 bodies: K1g spreads them over ALL SMs with one grid barrier per
    // step (force_grid.cu; measured on B200, us per step, K1g vs K3xw: 384 bodies 3.3 vs 3.9, 512: 3.2 vs 6.1, 1,024: 4.2 vs
    // 14.9; with most-attracting links 5.0 vs 4.9, 4.9 vs 6.8, 6.2 vs 18.8; below that the 16-SM cluster is ahead: 256: 2.9 vs 2.4)
    bool const grid_better = o.kernel == ESSA_KERNEL_AUTO && ctx->n >= kGridAutoMin && !needs_resident && force_grid_applicable(ctx, o, forces_only)
        && (o.forward_simulated || !mask_changes(ctx, ctx->date, ctx->date + (int64_t)steps * o.dt_seconds));
    bool resident = o.kernel == ESSA_KERNEL_RESIDENT
        || (o.kernel == ESSA_KERNEL_AUTO && ctx->n <= ESSA_RESIDENT_MAX && !grid_better && (ctx->n <= kResidentAutoMax || needs_resident || wide_ok));
    if (ctx->world > 1)
        resident = false;
    if (resident && ctx->n > ESSA_RESIDENT_MAX) {
        ctx->err = "resident kernel takes at most ESSA_RESIDENT_MAX bodies";
        return ESSA_ERR_ARG;
    }
    if (o.exact_order && !resident) {
        ctx->err = "exact_order is a resident-kernel mode (N <= ESSA_RESIDENT_MAX, unsharded)";
        return ESSA_ERR_ARG;
    }
    if (o.record && ctx->tracked == 0) {
        ctx->err = "record requested but essa_record_configure was not called";
        return ESSA_ERR_STATE;
    }
    CK(cudaEventRecord(ctx->ev_begin, ctx->stream));
    int rc;
    if (resident) {
        // worlds of up to 32 bodies in fast arithmetic: one-warp kernel with a recorder warp (resident_warp.cu)
        bool const closest = o.forward_simulated && ctx->closest && ctx->closest_n == ctx->n;
        if (ctx->n <= 32 && !closest && !forces_only) {
            bool const flips = !o.forward_simulated && mask_changes(ctx, ctx->date, ctx->date + (int64_t)steps * o.dt_seconds);
            long long const k = steps < 0 ? -(long long)steps : steps;
            if (ctx->persist.enabled && !flips && k < (1ll << 30) && resident_persist_applicable(ctx, o, forces_only)) {
                // stays resident after this tick and takes the following ones from its mailbox
                rc = persist_launch(ctx, steps, o);
                if (rc != ESSA_OK)
                    return rc;
                if (!async && (rc = persist_wait_ack(ctx)) != ESSA_OK)
                    return rc;
                return resident_finish(ctx, steps, o, forces_only);
            }
            CK(launch_resident_warp(ctx, steps, o, flips, ctx->stream));
        } else if (resident_cta_applicable(ctx, o, forces_only)
            && (o.forward_simulated || !mask_changes(ctx, ctx->date, ctx->date + (int64_t)steps * o.dt_seconds))) {
            // 33 .. 128 bodies: a cluster of 8 / 16 SMs (resident_cluster.cu); one SM if the cluster cannot be placed
            bool clustered = false;
            long long const k = steps < 0 ? -(long long)steps : steps;
            if (ctx->persist.enabled && k < (1ll << 30) && resident_cluster_persist_applicable(ctx, o, forces_only)) {
                // stays resident after this tick and takes the following ones from its mailbox (if the cluster can be placed)
                rc = persist_launch(ctx, steps, o);
                if (rc == ESSA_OK) {
                    if (!async && (rc = persist_wait_ack(ctx)) != ESSA_OK)
                        return rc;
                    return resident_finish(ctx, steps, o, forces_only);
                }
                (void)cudaGetLastError();
                ctx->persist.live = false;
            }
            if (resident_cluster_applicable(ctx, o, forces_only)) {
                clustered = launch_resident_cluster(ctx, steps, o, ctx->stream) == cudaSuccess;
                if (!clustered)
                    (void)cudaGetLastError();
            }
            if (!clustered)
                CK(launch_resident_cta(ctx, steps, o, ctx->stream));
        } else {
            // 129 .. 1024 bodies in plain fast arithmetic: the wide cluster kernel; everything else (and a cluster
            // that cannot be placed): the single-CTA kernel
            bool clustered = false;
            if (resident_cluster_wide_applicable(ctx, o, forces_only)
                && (o.forward_simulated || !mask_changes(ctx, ctx->date, ctx->date + (int64_t)steps * o.dt_seconds))) {
                clustered = launch_resident_cluster_wide(ctx, steps, o, ctx->stream) == cudaSuccess;
                if (!clustered)
                    (void)cudaGetLastError();
            }
            if (!clustered)
                CK(launch_resident(ctx, steps, o, forces_only, ctx->stream));
        }
        CK(cudaEventRecord(ctx->ev_end, ctx->stream));
        if (async) { // host-side bookkeeping does not need the results; the next synchronising call settles the rest
            ctx->timing_pending = true;
            return resident_finish(ctx, steps, o, forces_only);
        }
        CK(cudaStreamSynchronize(ctx->stream));
        rc = resident_finish(ctx, steps, o, forces_only);
    } else {
        rc = step_tiled(ctx, steps, o, forces_only);
    }
    if (rc != ESSA_OK)
        return rc;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end));
    ctx->last_step_ms = ms;
    return ESSA_OK;
}

extern "C" int essa_persist_configure(essa_ctx* ctx, int enable, int idle_timeout_us) {
    ARG(ctx, "ctx is null");
    ARG(idle_timeout_us >= 0 && idle_timeout_us <= 1000000, "idle timeout must be in [0, 1e6] microseconds");
    CK(cudaSetDevice(ctx->device));
    PARK();
    ctx->persist.enabled = enable != 0;
    if (idle_timeout_us > 0)
        ctx->persist.idle_us = idle_timeout_us;
    return ESSA_OK;
}

extern "C" int essa_persist_stats(const essa_ctx* c, int64_t out[6]) {
    if (!c || !out)
        return ESSA_ERR_ARG;
    out[0] = c->persist.live ? 1 : 0;
    out[1] = c->persist.ticks;
    out[2] = c->persist.launches;
    out[3] = c->persist.idle_exits;
    out[4] = c->persist.wait_ns;
    out[5] = c->persist.host_ns;
    return ESSA_OK;
}

extern "C" int essa_step(essa_ctx* ctx, int steps, const essa_step_opts* opts) { return step_common(ctx, steps, opts, false); }
extern "C" int essa_set_forces(essa_ctx* ctx, const essa_step_opts* opts) { return step_common(ctx, 1, opts, true); }
extern "C" int essa_step_async(essa_ctx* ctx, int steps, const essa_step_opts* opts) { return step_common(ctx, steps, opts, false, true); }

// ---------------------------------------------------------------------------------------------
// recording
// ---------------------------------------------------------------------------------------------
extern "C" int essa_record_configure(essa_ctx* ctx, int n_tracked, const int32_t* trail_capacity, int keep) {
    ARG(ctx, "ctx is null");
    ARG(n_tracked >= 0 && n_tracked <= ctx->n, "n_tracked out of range");
    ARG(n_tracked == 0 || trail_capacity, "trail_capacity is null");
    ARG(keep >= 0 && keep <= ctx->tracked && keep <= n_tracked, "keep out of range");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < keep; k++)
        ARG(trail_capacity[k] == ctx->trail_cap[k], "kept bodies must keep their trail capacity");
    if (n_tracked == 0) {
        free_recording(ctx);
        return ESSA_OK;
    }
    std::vector<int64_t> off(n_tracked);
    int64_t total = 0;
    for (int k = 0; k < n_tracked; k++) {
        ARG(trail_capacity[k] >= 3 && trail_capacity[k] <= 0xfffff, "trail capacity must be in [3, 0xfffff] (src/Trail.cpp:19-34)");
        off[k] = total;
        total += trail_capacity[k];
    }
    // new buffers, old contents of the kept prefix copied over
    essa_ctx old = *ctx; // shallow: used only for the old device pointers
    size_t T = (size_t)n_tracked;
    double *hist, *hfirst;
    int32_t *hs, *hl, *tc, *ao, *tl;
    float* verts;
    int64_t* voff;
    double *toff, *la;
    CK(cudaMalloc(&hist, T * ESSA_HISTORY_SEGMENTS * 6 * sizeof(double)));
    CK(cudaMalloc(&hfirst, T * 6 * sizeof(double)));
    CK(cudaMalloc(&hs, T * sizeof(int32_t)));
    CK(cudaMalloc(&hl, T * sizeof(int32_t)));
    CK(cudaMalloc(&tc, T * sizeof(int32_t)));
    CK(cudaMalloc(&ao, T * sizeof(int32_t)));
    CK(cudaMalloc(&tl, T * sizeof(int32_t)));
    CK(cudaMalloc(&verts, (size_t)total * 3 * sizeof(float)));
    CK(cudaMalloc(&voff, T * sizeof(int64_t)));
    CK(cudaMalloc(&toff, T * 3 * sizeof(double)));
    CK(cudaMalloc(&la, T * sizeof(double)));
    CK(cudaMemsetAsync(verts, 0, (size_t)total * 3 * sizeof(float), ctx->stream));
    CK(cudaMemsetAsync(toff, 0, T * 3 * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(la, 0, T * sizeof(double), ctx->stream));
    if (keep > 0) {
        size_t Kp = (size_t)keep;
        cudaStream_t st = ctx->stream;
        CK(cudaMemcpyAsync(hist, old.hist, Kp * ESSA_HISTORY_SEGMENTS * 6 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(hfirst, old.hist_first, Kp * 6 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(hs, old.hist_start, Kp * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(hl, old.hist_len, Kp * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(ao, old.append_off, Kp * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(tl, old.tlength, Kp * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(toff, old.toff, Kp * 3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(la, old.last_angle, Kp * sizeof(double), cudaMemcpyDeviceToDevice, st));
        int64_t kv = old.vert_off_h[keep - 1] + old.trail_cap[keep - 1];
        CK(cudaMemcpyAsync(verts, old.verts, (size_t)kv * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaMemcpyAsync(tc, trail_capacity, T * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(voff, off.data(), T * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    free_recording(ctx);
    ctx->hist = hist;
    ctx->hist_first = hfirst;
    ctx->hist_start = hs;
    ctx->hist_len = hl;
    ctx->tcap = tc;
    ctx->append_off = ao;
    ctx->tlength = tl;
    ctx->verts = verts;
    ctx->vert_off = voff;
    ctx->toff = toff;
    ctx->last_angle = la;
    ctx->tracked = n_tracked;
    ctx->verts_total = total;
    ctx->trail_cap.assign(trail_capacity, trail_capacity + n_tracked);
    ctx->vert_off_h = off;
    if (keep < n_tracked) {
        CK(launch_record_init(ctx, keep, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return ESSA_OK;
}

extern "C" int essa_record_reset(essa_ctx* ctx, int body, int history, int trail) {
    ARG(ctx, "ctx is null");
    ARG(body >= 0 && body < ctx->tracked, "body is not tracked");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    if (history) { // History::reset(): the list becomes { m_first_entry } (src/History.cpp:55-59)
        int32_t one = 1, zero = 0;
        CK(cudaMemcpy(ctx->hist + (size_t)body * ESSA_HISTORY_SEGMENTS * 6, ctx->hist_first + (size_t)body * 6, 6 * sizeof(double),
            cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(ctx->hist_start + body, &zero, sizeof(int32_t), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->hist_len + body, &one, sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    if (trail) { // Trail::reset(): m_append_offset = 1, m_length = 1 (src/Trail.cpp:77-80)
        int32_t one = 1;
        CK(cudaMemcpy(ctx->append_off + body, &one, sizeof(int32_t), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->tlength + body, &one, sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    return ESSA_OK;
}

extern "C" int essa_history_read(essa_ctx* ctx, int body, double* out) {
    ARG(ctx, "ctx is null");
    ARG(body >= 0 && body < ctx->tracked, "body is not tracked");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    int32_t start = 0, len = 0;
    CK(cudaMemcpy(&start, ctx->hist_start + body, sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&len, ctx->hist_len + body, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (!out)
        return len;
    double const* ring = ctx->hist + (size_t)body * ESSA_HISTORY_SEGMENTS * 6;
    int first = std::min(len, ESSA_HISTORY_SEGMENTS - start);
    if (first > 0)
        CK(cudaMemcpy(out, ring + (size_t)start * 6, (size_t)first * 6 * sizeof(double), cudaMemcpyDeviceToHost));
    if (len > first)
        CK(cudaMemcpy(out + (size_t)first * 6, ring, (size_t)(len - first) * 6 * sizeof(double), cudaMemcpyDeviceToHost));
    return len;
}

extern "C" int essa_history_write(essa_ctx* ctx, int body, const double* entries, int count, int set_first) {
    ARG(ctx, "ctx is null");
    ARG(body >= 0 && body < ctx->tracked, "body is not tracked");
    ARG(entries && count >= 1 && count <= ESSA_HISTORY_SEGMENTS, "count must be in [1, ESSA_HISTORY_SEGMENTS]");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    int32_t zero = 0;
    CK(cudaMemcpy(ctx->hist + (size_t)body * ESSA_HISTORY_SEGMENTS * 6, entries, (size_t)count * 6 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->hist_start + body, &zero, sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->hist_len + body, &count, sizeof(int32_t), cudaMemcpyHostToDevice));
    if (set_first)
        CK(cudaMemcpy(ctx->hist_first + (size_t)body * 6, entries, 6 * sizeof(double), cudaMemcpyHostToDevice));
    return ESSA_OK;
}

extern "C" int essa_trail_read(essa_ctx* ctx, int body, float* verts, essa_trail_meta* meta) {
    ARG(ctx, "ctx is null");
    ARG(body >= 0 && body < ctx->tracked, "body is not tracked");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    if (meta) {
        meta->capacity = ctx->trail_cap[body];
        meta->pad = 0;
        CK(cudaMemcpy(&meta->append_offset, ctx->append_off + body, sizeof(int32_t), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&meta->length, ctx->tlength + body, sizeof(int32_t), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(meta->offset, ctx->toff + (size_t)body * 3, 3 * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&meta->last_xy_angle, ctx->last_angle + body, sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (verts)
        CK(cudaMemcpy(verts, ctx->verts + (size_t)ctx->vert_off_h[body] * 3, (size_t)ctx->trail_cap[body] * 3 * sizeof(float),
            cudaMemcpyDeviceToHost));
    return ESSA_OK;
}

extern "C" int essa_trail_write(essa_ctx* ctx, int body, const float* verts, const essa_trail_meta* meta) {
    ARG(ctx, "ctx is null");
    ARG(body >= 0 && body < ctx->tracked, "body is not tracked");
    ARG(meta, "meta is null");
    ARG(meta->capacity == ctx->trail_cap[body], "capacity differs from essa_record_configure");
    ARG(meta->append_offset >= 0 && meta->append_offset < meta->capacity && meta->length >= 1 && meta->length <= meta->capacity,
        "append_offset / length out of range");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(ctx->append_off + body, &meta->append_offset, sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->tlength + body, &meta->length, sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->toff + (size_t)body * 3, meta->offset, 3 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->last_angle + body, &meta->last_xy_angle, sizeof(double), cudaMemcpyHostToDevice));
    if (verts)
        CK(cudaMemcpy(ctx->verts + (size_t)ctx->vert_off_h[body] * 3, verts, (size_t)ctx->trail_cap[body] * 3 * sizeof(float),
            cudaMemcpyHostToDevice));
    return ESSA_OK;
}

// ---------------------------------------------------------------------------------------------
// consumers of the recording: zero-copy Trail view for a renderer, flat snapshot
// ---------------------------------------------------------------------------------------------
extern "C" int essa_trail_device_view(essa_ctx* ctx, int body, essa_trail_view* out) {
    ARG(ctx, "ctx is null");
    ARG(out, "out is null");
    ARG(body >= 0 && body < ctx->tracked, "body is not tracked");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    out->verts_device = ctx->verts + (size_t)ctx->vert_off_h[body] * 3;
    out->capacity = ctx->trail_cap[body];
    out->pad = 0;
    CK(cudaMemcpy(&out->append_offset, ctx->append_off + body, sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&out->length, ctx->tlength + body, sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->offset, ctx->toff + (size_t)body * 3, 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return ESSA_OK;
}

// Snapshot layout (all little-endian, naturally aligned because every section is a multiple of 8 bytes):
//   char[8] "ESSASNAP" | u32 version = 1 | u32 n | u32 tracked | u32 0 | i64 date
//   f64[15][n]  x y z vx vy vz ax ay az gf max_attraction ap pe ap_vel pe_vel
//   i64[2][n]   creation_date deletion_date | i32[n] most (+ pad to 8) | u8[n] deleted (+ pad to 8)
//   per tracked body: i32 capacity, i32 append_offset, i32 length, i32 history_entries, f64 offset[3], f64 last_xy_angle,
//                     f64[6] History::m_first_entry, f64[history_entries][6] (oldest first), f32[capacity][3] (+ pad to 8)
namespace {
constexpr char kSnapMagic[8] = { 'E', 'S', 'S', 'A', 'S', 'N', 'A', 'P' };
inline size_t pad8(size_t b) { return (b + 7) & ~(size_t)7; }
}

extern "C" int64_t essa_snapshot_size(essa_ctx* ctx) {
    if (!ctx)
        return ESSA_ERR_ARG;
    size_t const N = (size_t)ctx->n;
    size_t bytes = 8 + 4 * 4 + 8 + 15 * N * 8 + 2 * N * 8 + pad8(N * 4) + pad8(N);
    for (int k = 0; k < ctx->tracked; k++) {
        int len = essa_history_read(ctx, k, nullptr);
        if (len < 0)
            return len;
        bytes += 4 * 4 + 4 * 8 + 6 * 8 + (size_t)len * 48 + pad8((size_t)ctx->trail_cap[k] * 12);
    }
    return (int64_t)bytes;
}

extern "C" int essa_snapshot_write(essa_ctx* ctx, void* blob, int64_t bytes) {
    ARG(ctx, "ctx is null");
    ARG(blob, "blob is null");
    int64_t const need = essa_snapshot_size(ctx);
    if (need < 0)
        return (int)need;
    ARG(bytes >= need, "blob is smaller than essa_snapshot_size()");
    size_t const N = (size_t)ctx->n;
    unsigned char* p = static_cast<unsigned char*>(blob);
    memset(p, 0, (size_t)need);
    memcpy(p, kSnapMagic, 8);
    uint32_t hdr[4] = { 1u, (uint32_t)ctx->n, (uint32_t)ctx->tracked, 0u };
    memcpy(p + 8, hdr, sizeof(hdr));
    int64_t const date = ctx->date;
    memcpy(p + 24, &date, 8);
    p += 32;
    essa_bodies b;
    memset(&b, 0, sizeof(b));
    double* f = reinterpret_cast<double*>(p);
    double** dm[15] = { &b.x, &b.y, &b.z, &b.vx, &b.vy, &b.vz, &b.ax, &b.ay, &b.az, &b.gf, &b.max_attraction, &b.ap, &b.pe, &b.ap_vel, &b.pe_vel };
    for (int k = 0; k < 15; k++)
        *dm[k] = f + (size_t)k * N;
    p += 15 * N * 8;
    b.creation_date = reinterpret_cast<int64_t*>(p);
    b.deletion_date = b.creation_date + N;
    p += 2 * N * 8;
    b.most = reinterpret_cast<int32_t*>(p);
    p += pad8(N * 4);
    b.deleted = p;
    p += pad8(N);
    if (N) {
        int rc = essa_download(ctx, &b);
        if (rc != ESSA_OK)
            return rc;
    }
    for (int k = 0; k < ctx->tracked; k++) {
        essa_trail_meta meta;
        int32_t* ih = reinterpret_cast<int32_t*>(p);
        double* dh = reinterpret_cast<double*>(p + 16);
        int len = essa_history_read(ctx, k, dh + 4 + 6);
        if (len < 0)
            return len;
        CK(cudaMemcpy(dh + 4, ctx->hist_first + (size_t)k * 6, 6 * sizeof(double), cudaMemcpyDeviceToHost));
        float* verts = reinterpret_cast<float*>(p + 16 + 4 * 8 + 6 * 8 + (size_t)len * 48);
        int rc = essa_trail_read(ctx, k, verts, &meta);
        if (rc != ESSA_OK)
            return rc;
        ih[0] = meta.capacity;
        ih[1] = meta.append_offset;
        ih[2] = meta.length;
        ih[3] = len;
        memcpy(dh, meta.offset, 3 * sizeof(double));
        dh[3] = meta.last_xy_angle;
        p += 16 + 4 * 8 + 6 * 8 + (size_t)len * 48 + pad8((size_t)meta.capacity * 12);
    }
    return ESSA_OK;
}

extern "C" int essa_snapshot_read(essa_ctx* ctx, const void* blob, int64_t bytes) {
    ARG(ctx, "ctx is null");
    ARG(blob && bytes >= 32, "blob is too short");
    unsigned char const* p = static_cast<unsigned char const*>(blob);
    unsigned char const* const end = p + bytes;
    ARG(memcmp(p, kSnapMagic, 8) == 0, "not an ESSASNAP blob");
    uint32_t hdr[4];
    memcpy(hdr, p + 8, sizeof(hdr));
    ARG(hdr[0] == 1u, "unknown snapshot version");
    size_t const N = hdr[1];
    int const tracked = (int)hdr[2];
    int64_t date;
    memcpy(&date, p + 24, 8);
    p += 32;
    ARG((size_t)(end - p) >= 15 * N * 8 + 2 * N * 8 + pad8(N * 4) + pad8(N) && (size_t)tracked <= N, "blob is truncated");
    int rc = essa_set_date(ctx, date);
    if (rc != ESSA_OK)
        return rc;
    essa_bodies b;
    memset(&b, 0, sizeof(b));
    double* f = const_cast<double*>(reinterpret_cast<double const*>(p));
    double** dm[15] = { &b.x, &b.y, &b.z, &b.vx, &b.vy, &b.vz, &b.ax, &b.ay, &b.az, &b.gf, &b.max_attraction, &b.ap, &b.pe, &b.ap_vel, &b.pe_vel };
    for (int k = 0; k < 15; k++)
        *dm[k] = f + (size_t)k * N;
    p += 15 * N * 8;
    b.creation_date = const_cast<int64_t*>(reinterpret_cast<int64_t const*>(p));
    b.deletion_date = b.creation_date + N;
    p += 2 * N * 8;
    b.most = const_cast<int32_t*>(reinterpret_cast<int32_t const*>(p));
    p += pad8(N * 4);
    b.deleted = const_cast<uint8_t*>(p);
    p += pad8(N);
    rc = essa_upload(ctx, (int)N, &b);
    if (rc != ESSA_OK)
        return rc;
    if (tracked == 0)
        return essa_record_configure(ctx, 0, nullptr, 0);
    // first pass: capacities, then the rings
    std::vector<int32_t> caps(tracked);
    unsigned char const* q = p;
    for (int k = 0; k < tracked; k++) {
        ARG((size_t)(end - q) >= 16 + 4 * 8 + 6 * 8, "blob is truncated");
        int32_t ih[4];
        memcpy(ih, q, sizeof(ih));
        ARG(ih[0] >= 3 && ih[3] >= 1 && ih[3] <= ESSA_HISTORY_SEGMENTS, "bad recording header");
        caps[k] = ih[0];
        size_t const sz = 16 + 4 * 8 + 6 * 8 + (size_t)ih[3] * 48 + pad8((size_t)ih[0] * 12);
        ARG((size_t)(end - q) >= sz, "blob is truncated");
        q += sz;
    }
    rc = essa_record_configure(ctx, tracked, caps.data(), 0);
    if (rc != ESSA_OK)
        return rc;
    for (int k = 0; k < tracked; k++) {
        int32_t ih[4];
        memcpy(ih, p, sizeof(ih));
        double const* dh = reinterpret_cast<double const*>(p + 16);
        if ((rc = essa_history_write(ctx, k, dh + 4, 1, 1)) != ESSA_OK) // History::m_first_entry
            return rc;
        if ((rc = essa_history_write(ctx, k, dh + 4 + 6, ih[3], 0)) != ESSA_OK)
            return rc;
        essa_trail_meta meta;
        meta.capacity = ih[0];
        meta.append_offset = ih[1];
        meta.length = ih[2];
        meta.pad = 0;
        memcpy(meta.offset, dh, 3 * sizeof(double));
        meta.last_xy_angle = dh[3];
        float const* verts = reinterpret_cast<float const*>(p + 16 + 4 * 8 + 6 * 8 + (size_t)ih[3] * 48);
        if ((rc = essa_trail_write(ctx, k, verts, &meta)) != ESSA_OK)
            return rc;
        p += 16 + 4 * 8 + 6 * 8 + (size_t)ih[3] * 48 + pad8((size_t)ih[0] * 12);
    }
    return ESSA_OK;
}

// ---------------------------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------------------------
extern "C" int essa_energy(essa_ctx* ctx, double* kinetic, double* potential, double momentum[3]) {
    ARG(ctx, "ctx is null");
    CK(cudaSetDevice(ctx->device));
    PARK();
    double out[5] = { 0, 0, 0, 0, 0 };
    if (ctx->n > 0) {
        CK(launch_energy(ctx, ctx->stream));
        CK(cudaMemcpyAsync(out, ctx->escratch, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->world > 1) { // each rank summed its own targets' rows
            CK(cudaMemcpyAsync(ctx->escratch, out, 5 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            NK(nccl().AllReduce(ctx->escratch, ctx->escratch, 5, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream));
            CK(cudaMemcpyAsync(out, ctx->escratch, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    if (kinetic)
        *kinetic = out[0];
    if (potential)
        *potential = out[1];
    if (momentum) {
        momentum[0] = out[2];
        momentum[1] = out[3];
        momentum[2] = out[4];
    }
    return ESSA_OK;
}

extern "C" double essa_measure_fp64_peak(essa_ctx* ctx, double ms) {
    if (!ctx)
        return -1;
    if (cudaSetDevice(ctx->device) != cudaSuccess || persist_park(ctx) != ESSA_OK)
        return -1;
    return run_fp64_peak(ctx, ms);
}

extern "C" int essa_measure_rsqrt_error(essa_ctx* ctx, double* seed_rel_err, double* refined_rel_err) {
    ARG(ctx, "ctx is null");
    CK(cudaSetDevice(ctx->device));
    PARK();
    return run_rsqrt_error(ctx, seed_rel_err, refined_rel_err);
}

// ---------------------------------------------------------------------------------------------
// ensemble
// ---------------------------------------------------------------------------------------------
extern "C" double essa_ensemble_factor(uint64_t seed, double rel_sigma, int64_t copy, int body, int component) {
    return ensemble_factor(seed, rel_sigma, (long long)copy, body, component);
}

extern "C" int essa_ensemble_init(essa_ctx* ctx, const essa_ensemble_opts* o) {
    ARG(ctx, "ctx is null");
    ARG(o && o->copies > 0, "copies must be positive");
    ARG(ctx->n > 0, "upload the template world first");
    ARG(ctx->n <= kEnsembleMaxBodies, "ensemble template must have at most ESSA_ENSEMBLE_MAX bodies");
    ARG(!o->closest_approaches || (o->approach_to >= 0 && o->approach_to < ctx->n), "approach_to out of range");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    ARG(o->trail_every >= 0 && o->trail_capacity >= 0 && (o->trail_every == 0 || (o->closest_approaches && o->trail_capacity > 0)),
        "trail_every needs closest_approaches and a positive trail_capacity");
    cudaFree(ctx->ens_state);
    cudaFree(ctx->ens_gf);
    cudaFree(ctx->ens_mind);
    cudaFree(ctx->ens_active);
    cudaFree(ctx->ens_apos);
    cudaFree(ctx->ens_trail);
    cudaFree(ctx->ens_trail_len);
    ctx->ens_state = ctx->ens_gf = ctx->ens_mind = ctx->ens_apos = ctx->ens_trail = nullptr;
    ctx->ens_active = nullptr;
    ctx->ens_trail_len = nullptr;
    ctx->ens_steps_done = 0;
    ctx->ens_copies = o->copies;
    ctx->ens_n = ctx->n;
    ctx->ens = *o;
    size_t total = (size_t)o->copies * ctx->n;
    CK(cudaMalloc(&ctx->ens_state, total * 6 * sizeof(double)));
    CK(cudaMalloc(&ctx->ens_gf, (size_t)ctx->n * sizeof(double)));
    CK(cudaMalloc(&ctx->ens_mind, total * sizeof(double)));
    CK(cudaMalloc(&ctx->ens_active, (size_t)ctx->n));
    if (o->closest_approaches) {
        CK(cudaMalloc(&ctx->ens_apos, total * 6 * sizeof(double)));
        CK(cudaMemsetAsync(ctx->ens_apos, 0, total * 6 * sizeof(double), ctx->stream));
    }
    if (o->trail_every > 0) {
        CK(cudaMalloc(&ctx->ens_trail, (size_t)o->copies * o->trail_capacity * 3 * sizeof(double)));
        CK(cudaMalloc(&ctx->ens_trail_len, (size_t)o->copies * sizeof(int)));
        CK(cudaMemsetAsync(ctx->ens_trail, 0, (size_t)o->copies * o->trail_capacity * 3 * sizeof(double), ctx->stream));
        CK(cudaMemsetAsync(ctx->ens_trail_len, 0, (size_t)o->copies * sizeof(int), ctx->stream));
    }
    CK(launch_ensemble_init(ctx, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->ens_on_mask = 0;
    if (ctx->n <= kEnsembleSoloMax) { // small templates: weights and live mask travel as kernel parameters (k_ensemble_solo)
        uint8_t on[kEnsembleSoloMax];
        CK(cudaMemcpy(ctx->ens_gf_host, ctx->ens_gf, (size_t)ctx->n * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(on, ctx->ens_active, (size_t)ctx->n, cudaMemcpyDeviceToHost));
        for (int k = 0; k < ctx->n; k++)
            ctx->ens_on_mask |= on[k] ? 1u << k : 0u;
    }
    char const* solo = getenv("ESSA_ENSEMBLE_SOLO");
    ctx->ens_solo = !(solo && solo[0] == '0');
    return ESSA_OK;
}

extern "C" int essa_ensemble_step(essa_ctx* ctx, int steps) {
    ARG(ctx, "ctx is null");
    ARG(steps != 0, "steps must not be 0");
    ARG(ctx->ens_copies > 0, "essa_ensemble_init first");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaEventRecord(ctx->ev_begin, ctx->stream));
    CK(launch_ensemble_step(ctx, steps, ctx->stream));
    CK(cudaEventRecord(ctx->ev_end, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end));
    ctx->last_step_ms = ms;
    ctx->last_force_ms = ms;
    return ESSA_OK;
}

extern "C" int essa_ensemble_read(essa_ctx* ctx, int first, int count, double* x, double* y, double* z, double* vx, double* vy,
    double* vz, double* min_distance) {
    ARG(ctx, "ctx is null");
    ARG(first >= 0 && count >= 0 && first + count <= ctx->ens_copies, "copy range out of bounds");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    size_t n = (size_t)ctx->ens_n;
    double* dst[6] = { x, y, z, vx, vy, vz };
    for (int c = 0; c < count; c++) {
        double const* src = ctx->ens_state + (size_t)(first + c) * 6 * n;
        for (int k = 0; k < 6; k++)
            if (dst[k])
                CK(cudaMemcpyAsync(dst[k] + (size_t)c * n, src + k * n, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (min_distance)
        CK(cudaMemcpyAsync(min_distance, ctx->ens_mind + (size_t)first * n, (size_t)count * n * sizeof(double), cudaMemcpyDeviceToHost,
            ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ESSA_OK;
}

extern "C" int essa_ensemble_read_approaches(essa_ctx* ctx, int first, int count, double* this_pos, double* other_pos) {
    ARG(ctx, "ctx is null");
    ARG(first >= 0 && count >= 0 && first + count <= ctx->ens_copies, "copy range out of bounds");
    ARG(ctx->ens_apos, "the ensemble was initialised without closest_approaches");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    size_t const n = (size_t)ctx->ens_n, rows = (size_t)count * n;
    std::vector<double> tmp(rows * 6);
    CK(cudaMemcpy(tmp.data(), ctx->ens_apos + (size_t)first * n * 6, rows * 6 * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < rows; r++)
        for (int c = 0; c < 3; c++) {
            if (this_pos)
                this_pos[3 * r + c] = tmp[6 * r + c];
            if (other_pos)
                other_pos[3 * r + c] = tmp[6 * r + 3 + c];
        }
    return ESSA_OK;
}

extern "C" int essa_ensemble_read_trail(essa_ctx* ctx, int first, int count, double* xyz, int32_t* lengths) {
    ARG(ctx, "ctx is null");
    ARG(first >= 0 && count >= 0 && first + count <= ctx->ens_copies, "copy range out of bounds");
    ARG(ctx->ens_trail, "the ensemble was initialised without trail_every");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    size_t const cap = (size_t)ctx->ens.trail_capacity;
    if (xyz)
        CK(cudaMemcpy(xyz, ctx->ens_trail + (size_t)first * cap * 3, (size_t)count * cap * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (lengths)
        CK(cudaMemcpy(lengths, ctx->ens_trail_len + first, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return ESSA_OK;
}

// ---------------------------------------------------------------------------------------------
// sharding
// ---------------------------------------------------------------------------------------------
extern "C" int essa_nccl_unique_id(uint8_t id[ESSA_NCCL_ID_BYTES]) {
    if (!id)
        return ESSA_ERR_ARG;
    NcclApi& api = nccl();
    if (!api.lib) {
        g_create_error = api.why;
        return ESSA_ERR_NCCL;
    }
    NcclId nid;
    if (api.GetUniqueId(&nid) != 0) {
        g_create_error = "ncclGetUniqueId failed";
        return ESSA_ERR_NCCL;
    }
    memcpy(id, nid.internal, ESSA_NCCL_ID_BYTES);
    return ESSA_OK;
}

extern "C" int essa_shard_attach(essa_ctx* ctx, int rank, int world, const uint8_t id[ESSA_NCCL_ID_BYTES]) {
    ARG(ctx, "ctx is null");
    ARG(world >= 1 && rank >= 0 && rank < world && id, "bad rank / world / id");
    ARG(!ctx->comm, "already attached");
    CK(cudaSetDevice(ctx->device));
    PARK();
    if (world == 1)
        return ESSA_OK;
    NcclApi& api = nccl();
    if (!api.lib) {
        ctx->err = api.why;
        return ESSA_ERR_NCCL;
    }
    NcclId nid;
    memcpy(nid.internal, id, ESSA_NCCL_ID_BYTES);
    NK(api.CommInitRank(&ctx->comm, world, nid, rank));
    ctx->rank = rank;
    ctx->world = world;
    ctx->acc_valid = false;
    return ESSA_OK;
}

// ---------------------------------------------------------------------------------------------
// measurement hygiene: evict the L2 (126 MB) by writing a larger scratch buffer
// ---------------------------------------------------------------------------------------------
extern "C" int essa_flush_l2(essa_ctx* ctx) {
    ARG(ctx, "ctx is null");
    CK(cudaSetDevice(ctx->device));
    PARK();
    size_t const bytes = 256u << 20;
    if (!ctx->l2_scratch)
        CK(cudaMalloc(&ctx->l2_scratch, bytes));
    ctx->l2_value++;
    CK(cudaMemsetAsync(ctx->l2_scratch, ctx->l2_value & 0xff, bytes, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ESSA_OK;
}

// ---------------------------------------------------------------------------------------------
// closest approaches (Object::m_closest_approaches, src/Object.hpp:184-189, src/Object.cpp:405-417)
// ---------------------------------------------------------------------------------------------
extern "C" int essa_closest_enable(essa_ctx* ctx, int enable) {
    ARG(ctx, "ctx is null");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->closest);
    ctx->closest = nullptr;
    ctx->closest_n = 0;
    if (!enable || ctx->n == 0)
        return ESSA_OK;
    ARG(ctx->n <= ESSA_RESIDENT_MAX, "closest approaches are tracked by the resident kernel only");
    size_t bytes = (size_t)ctx->n * ctx->n * 7 * sizeof(double);
    CK(cudaMalloc(&ctx->closest, bytes));
    CK(cudaMemset(ctx->closest, 0, bytes));
    ctx->closest_n = ctx->n;
    return ESSA_OK;
}

extern "C" int essa_closest_read(essa_ctx* ctx, int body, int other, double out[7]) {
    ARG(ctx, "ctx is null");
    ARG(out, "out is null");
    ARG(ctx->closest && body >= 0 && other >= 0 && body < ctx->closest_n && other < ctx->closest_n, "closest approaches not enabled / index");
    CK(cudaSetDevice(ctx->device));
    PARK();
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out, ctx->closest + ((size_t)body * ctx->closest_n + other) * 7, 7 * sizeof(double), cudaMemcpyDeviceToHost));
    return out[0] != 0.0 ? 1 : 0;
}

// Targets [*i0, *i1) that rank `rank` of `world` integrates in an n-body world (ragged shards allowed).
extern "C" int essa_shard_range(int n, int rank, int world, int* i0, int* i1) {
    if (n < 0 || world < 1 || rank < 0 || rank >= world || !i0 || !i1)
        return ESSA_ERR_ARG;
    *i0 = (int)((long long)n * rank / world);
    *i1 = (int)((long long)n * (rank + 1) / world);
    return ESSA_OK;
}

namespace essa {
int run_latency(essa_ctx* c, double* out6);
}
extern "C" int essa_measure_latency(essa_ctx* ctx, double out[8]) {
    ARG(ctx, "ctx is null");
    ARG(out, "out is null");
    CK(cudaSetDevice(ctx->device));
    PARK();
    return run_latency(ctx, out);
}
