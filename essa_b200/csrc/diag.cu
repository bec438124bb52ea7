// K5 -- energy / momentum diagnostic, and the two measurement kernels behind the roofline
// denominators (FP64 FMA peak, MUFU.RSQ64H seed accuracy).  Not in the reference: the north star
// asks for energy drift next to every throughput number.
#include <vector>

#include "common.cuh"
#include "ctx.hpp"

namespace essa {

constexpr int kEThreads = 256;

// 1/r with the same seed + third-order correction as the force kernel.
__device__ __forceinline__ double inv_r(double r2) {
    double y0 = rsqrt_seed(r2);
    double t = r2 * y0;
    double e = fma(-t, y0, 1.0);
    double p = fma(0.375, e, 0.5);
    double q = y0 * e;
    return fma(q, p, y0);
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
    int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int s = kEThreads / 2; s > 0; s >>= 1) {
        if (tid < s)
            sh[tid] = sh[tid] + sh[tid + s];
        __syncthreads();
    }
    double r = sh[0];
    __syncthreads();
    return r;
}

// Potential energy with every unordered pair evaluated ONCE (the pair-shared enumeration of K1p, force_paired.cu): bodies
// are cut into tiles of 256; the CTA of target tile T meets source tiles T + d (mod ntiles), d = 0 .. ntiles / 2 -- the tile
// against itself with j > i only, and for an even tile count the opposite tile from the lower half only -- so that every tile
// does the same amount of work and no pair is seen twice: 12 FP64 instructions per PAIR (the energy needs no reaction term:
// sum_pairs gf_i gf_j / r = sum_i gf_i sum_{j met by i} gf_j / r) where the one-sided version spent 13 per directed interaction.
// grid = (tiles of this rank, S chunks of the offsets); CTA (T, sj) writes 5 partial sums:
//   sum_i gf_i * sum_{j met in chunk} gf_j / r_ij ,   and for sj == 0:  sum gf v^2, sum gf v.
__global__ void __launch_bounds__(kEThreads) k_energy(double4 const* pos, double const* vx, double const* vy, double const* vz, int n,
    int tile0, int S, double* partial) {
    __shared__ double4 tile[kEThreads];
    __shared__ double red[kEThreads];
    int tid = threadIdx.x;
    int ib = blockIdx.x / S, sj = blockIdx.x - ib * S;
    int const T = tile0 + ib;
    int i = T * kEThreads + tid;
    bool live = i < n;
    double4 pi = pos[live ? i : n - 1];
    int ntiles = (n + kEThreads - 1) / kEThreads;
    int const nd = ntiles / 2 + 1; // offsets 0 .. ntiles / 2
    int d0 = (int)((long long)nd * sj / S), d1 = (int)((long long)nd * (sj + 1) / S);
    double phi = 0;
    for (int d = d0; d < d1; d++) {
        if (d > 0 && 2 * d == ntiles && T >= d)
            continue; // the opposite tile of an even ring: met from the lower half
        int t = T + d;
        t = t >= ntiles ? t - ntiles : t;
        if (d > 0 && t == T)
            continue; // a ring of one tile
        int j = t * kEThreads + tid;
        tile[tid] = j < n ? pos[j] : make_double4(0, 0, 0, 0);
        __syncthreads();
        int cnt = min(kEThreads, n - t * kEThreads);
        if (d == 0) {
            for (int jj = 0; jj < cnt; jj++) {
                double4 q = tile[jj];
                double dx = q.x - pi.x, dy = q.y - pi.y, dz = q.z - pi.z;
                double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                phi = fma(jj > tid ? q.w : 0.0, inv_r(r2), phi);
            }
        } else {
#pragma unroll 4
            for (int jj = 0; jj < cnt; jj++) {
                double4 q = tile[jj];
                double dx = q.x - pi.x, dy = q.y - pi.y, dz = q.z - pi.z;
                double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                phi = fma(q.w, inv_r(r2), phi);
            }
        }
        __syncthreads();
    }
    double w = live ? pi.w : 0.0; // gf_eff: inactive bodies carry 0
    double pe = block_sum(w * phi, red);
    double ke = 0, px = 0, py = 0, pz = 0;
    if (sj == 0) {
        double ux = live ? vx[i] : 0, uy = live ? vy[i] : 0, uz = live ? vz[i] : 0;
        ke = block_sum(w * (ux * ux + uy * uy + uz * uz), red);
        px = block_sum(w * ux, red);
        py = block_sum(w * uy, red);
        pz = block_sum(w * uz, red);
    }
    if (tid == 0) {
        double* o = partial + (size_t)blockIdx.x * 5;
        o[0] = ke;
        o[1] = pe;
        o[2] = px;
        o[3] = py;
        o[4] = pz;
    }
}

// Fixed-order final sum; converts gf-weighted sums to SI (m = gf / G).
__global__ void k_energy_final(double const* partial, int blocks, double* out) {
    __shared__ double red[kEThreads];
    double acc[5] = { 0, 0, 0, 0, 0 };
    for (int b = threadIdx.x; b < blocks; b += kEThreads)
        for (int k = 0; k < 5; k++)
            acc[k] += partial[(size_t)b * 5 + k];
    double r[5];
    for (int k = 0; k < 5; k++)
        r[k] = block_sum(acc[k], red);
    if (threadIdx.x == 0) {
        double const invG = 1.0 / ESSA_GRAVITY;
        out[0] = 0.5 * r[0] * invG;          // sum 1/2 m v^2
        out[1] = -r[1] * invG;               // - sum_{i<j} G m_i m_j / r  (every pair was met once)
        out[2] = r[2] * invG;
        out[3] = r[3] * invG;
        out[4] = r[4] * invG;
    }
}

cudaError_t launch_energy(essa_ctx* c, cudaStream_t st) {
    int n = c->n;
    int ntiles = (n + kEThreads - 1) / kEThreads;
    // whole tiles per rank (the sum over ranks is an all-reduce: which rank owns a body does not matter here)
    int tile0 = (int)((long long)ntiles * c->rank / c->world), tile1 = (int)((long long)ntiles * (c->rank + 1) / c->world);
    int iblocks = tile1 - tile0;
    int const nd = ntiles / 2 + 1;
    int S = 1;
    while ((long long)iblocks * S < 4LL * c->sm_count && S * 2 <= nd)
        S *= 2;
    int blocks = iblocks * S;
    size_t need = 8 + (size_t)blocks * 5;
    if (c->escratch_n < need) {
        cudaFree(c->escratch);
        c->escratch = nullptr;
        c->escratch_n = 0;
        cudaError_t e = cudaMalloc(&c->escratch, need * sizeof(double));
        if (e != cudaSuccess)
            return e;
        c->escratch_n = need;
    }
    if (iblocks == 0)
        return cudaMemsetAsync(c->escratch, 0, 8 * sizeof(double), st);
    double* partial = c->escratch + 8;
    k_energy<<<blocks, kEThreads, 0, st>>>(c->pos[c->cur], c->vel, c->vel + c->cap, c->vel + 2 * (size_t)c->cap, n, tile0, S, partial);
    k_energy_final<<<1, kEThreads, 0, st>>>(partial, blocks, c->escratch);
    c->launches += 2;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// FP64 FMA peak: 8 independent register-resident DFMA chains per thread.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            r0 = fma(r0, a, b);
            r1 = fma(r1, a, b);
            r2 = fma(r2, a, b);
            r3 = fma(r3, a, b);
            r4 = fma(r4, a, b);
            r5 = fma(r5, a, b);
            r6 = fma(r6, a, b);
            r7 = fma(r7, a, b);
        }
    }
    double s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    if (s == 123.456)
        out[0] = s;
}

double run_fp64_peak(essa_ctx* c, double ms_target) {
    double* dummy;
    if (cudaMalloc(&dummy, 8) != cudaSuccess)
        return -1;
    int blocks = c->sm_count * 8;
    auto run = [&](int iters) -> double {
        cudaEventRecord(c->ev_f0, c->stream);
        k_dfma_peak<<<blocks, 256, 0, c->stream>>>(dummy, iters, 0.999999, 1e-9);
        cudaEventRecord(c->ev_f1, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess)
            return -1;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev_f0, c->ev_f1);
        c->launches++;
        return ms;
    };
    run(64); // warm-up
    int iters = 256;
    double ms = run(iters);
    if (ms <= 0) {
        cudaFree(dummy);
        return -1;
    }
    double want = ms_target > 0 ? ms_target : 200.0;
    long long it2 = (long long)(iters * want / ms);
    if (it2 < 256)
        it2 = 256;
    if (it2 > 2000000000LL)
        it2 = 2000000000LL;
    ms = run((int)it2);
    cudaFree(dummy);
    if (ms <= 0)
        return -1;
    double flops = (double)blocks * 256.0 * (double)it2 * 16.0 * 8.0 * 2.0;
    return flops / (ms * 1e-3) / 1e12;
}

// ------------------------------------------------------------------------------------------
// Accuracy of the MUFU.RSQ64H seed and of the refined gf / r^3 over 2^-60 .. 2^120.
// ------------------------------------------------------------------------------------------
__global__ void k_rsqrt_error(double* out) {
    double worst_seed = 0, worst_ref = 0;
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    int total = gridDim.x * blockDim.x;
    for (int k = tid; k < 180 * 4096; k += total) {
        double r2 = exp2((double)k / 4096.0 - 60.0) * (1.0 + 1e-3 * (k % 7));
        double y0 = rsqrt_seed(r2);
        double ex = 1.0 / sqrt(r2);
        double es = fabs(y0 - ex) / ex;
        double s = inv_r3_times(r2, 1.0);
        double ex3 = 1.0 / (r2 * sqrt(r2));
        double er = fabs(s - ex3) / ex3;
        worst_seed = fmax(worst_seed, es);
        worst_ref = fmax(worst_ref, er);
    }
    out[2 * tid] = worst_seed;
    out[2 * tid + 1] = worst_ref;
}

int run_rsqrt_error(essa_ctx* c, double* seed, double* refined) {
    int blocks = 64, threads = 256, total = blocks * threads;
    double* d;
    if (cudaMalloc(&d, (size_t)total * 2 * sizeof(double)) != cudaSuccess)
        return ESSA_ERR_CUDA;
    k_rsqrt_error<<<blocks, threads, 0, c->stream>>>(d);
    c->launches++;
    std::vector<double> h((size_t)total * 2);
    cudaError_t e = cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) {
        c->err = cudaGetErrorString(e);
        return ESSA_ERR_CUDA;
    }
    double ws = 0, wr = 0;
    for (int i = 0; i < total; i++) {
        ws = h[2 * i] > ws ? h[2 * i] : ws;
        wr = h[2 * i + 1] > wr ? h[2 * i + 1] : wr;
    }
    if (seed)
        *seed = ws;
    if (refined)
        *refined = wr;
    return ESSA_OK;
}

} // namespace essa

// ------------------------------------------------------------------------------------------
// Dependent-issue latencies (cycles) of the instructions on a single trajectory's critical
// path: DFMA, DADD, DMUL, MUFU.RSQ64H (+ its guard), LDS.64 pointer chase, warp shuffle of a
// double.  One warp, one chain, timed with clock64.
// ------------------------------------------------------------------------------------------
namespace essa {

__global__ void k_latency(double* out, double a, double b, int iters) {
    __shared__ int chase[32];
    chase[threadIdx.x] = (threadIdx.x + 1) & 31;
    __syncwarp();
    double r = a;
    long long t0, t1;
    // DFMA
    t0 = clock64();
    for (int i = 0; i < iters; i++)
        r = fma(r, a, b);
    t1 = clock64();
    double l_fma = double(t1 - t0) / iters;
    // DADD
    t0 = clock64();
    for (int i = 0; i < iters; i++)
        r = __dadd_rn(r, b);
    t1 = clock64();
    double l_add = double(t1 - t0) / iters;
    // DMUL
    t0 = clock64();
    for (int i = 0; i < iters; i++)
        r = __dmul_rn(r, a);
    t1 = clock64();
    double l_mul = double(t1 - t0) / iters;
    // rsqrt seed (MUFU.RSQ64H + integer guard) chained through a DADD to keep the value in range
    double s = fabs(r) + 2.0;
    t0 = clock64();
    for (int i = 0; i < iters; i++)
        s = __dadd_rn(rsqrt_seed(s), 2.0);
    t1 = clock64();
    double l_rsq = double(t1 - t0) / iters - l_add;
    // LDS pointer chase
    int p = threadIdx.x;
    t0 = clock64();
    for (int i = 0; i < iters; i++)
        p = chase[p];
    t1 = clock64();
    double l_lds = double(t1 - t0) / iters;
    // shuffle of a double (2 x SHFL.32)
    double v = s;
    t0 = clock64();
    for (int i = 0; i < iters; i++)
        v = __shfl_sync(0xffffffffu, v, (threadIdx.x + 1) & 31);
    t1 = clock64();
    double l_shfl = double(t1 - t0) / iters;
    // single-warp FP64 issue rate: 4 and 8 independent DFMA chains (cycles per instruction)
    double c0 = a, c1 = a + 1, c2 = a + 2, c3 = a + 3, c4 = a + 4, c5 = a + 5, c6 = a + 6, c7 = a + 7;
    t0 = clock64();
    for (int i = 0; i < iters; i++) {
        c0 = fma(c0, a, b);
        c1 = fma(c1, a, b);
        c2 = fma(c2, a, b);
        c3 = fma(c3, a, b);
    }
    t1 = clock64();
    double thr4 = double(t1 - t0) / (4.0 * iters);
    t0 = clock64();
    for (int i = 0; i < iters; i++) {
        c0 = fma(c0, a, b);
        c1 = fma(c1, a, b);
        c2 = fma(c2, a, b);
        c3 = fma(c3, a, b);
        c4 = fma(c4, a, b);
        c5 = fma(c5, a, b);
        c6 = fma(c6, a, b);
        c7 = fma(c7, a, b);
    }
    t1 = clock64();
    double thr8 = double(t1 - t0) / (8.0 * iters);
    if (threadIdx.x == 0) {
        out[7] = thr4;
        out[8] = thr8;
        out[9] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
        out[0] = l_fma;
        out[1] = l_add;
        out[2] = l_mul;
        out[3] = l_rsq;
        out[4] = l_lds;
        out[5] = l_shfl;
        out[6] = r + s + v + p; // keep everything alive
    }
}

int run_latency(essa_ctx* c, double* out6) {
    double* d;
    if (cudaMalloc(&d, 16 * sizeof(double)) != cudaSuccess)
        return ESSA_ERR_CUDA;
    k_latency<<<1, 32, 0, c->stream>>>(d, 0.9999999, 1e-9, 4096);
    c->launches++;
    double h[16];
    cudaError_t e = cudaMemcpyAsync(h, d, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) {
        c->err = cudaGetErrorString(e);
        return ESSA_ERR_CUDA;
    }
    for (int i = 0; i < 6; i++)
        out6[i] = h[i];
    out6[6] = h[7];
    out6[7] = h[8];
    return ESSA_OK;
}

} // namespace essa
