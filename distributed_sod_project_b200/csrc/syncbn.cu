// syncbn.cu — synchronized batch-norm forward / backward with the cross-GPU statistics exchange and
// the (pre-add → normalize/affine → residual-add → ReLU) elementwise chain in ONE kernel per direction.
//
// Reference behaviour being replaced: apex.parallel.SyncBatchNorm as installed by
// convert_syncbn_model (reference train.py:180): local Welford → 2 NCCL all_gathers → elementwise
// forward; local reduce → 2 NCCL all_reduces → elementwise backward — four kernels and two latency-
// bound collectives per layer per direction, 84 layers.  Arithmetic spec (readable in-container
// equivalent): torch/nn/modules/_functions.py:7-209.
//
// Layout: channels-last matrix [rows = N·H·W, C].  A CTA owns a (slab of ≤256 channels) × (strip of
// rows).  Each thread owns 8 consecutive channels (one 16-byte packet for 16-bit data) and walks rows.
//
//   phase 1  per-thread fp32 (Σ, Σ²) → warp shuffles → shared memory → one partial per CTA in L2;
//            the LAST CTA of a slab to arrive (atomic ticket) sums the slab's partials in fixed order
//            and publishes the LOCAL totals to every rank: 8-byte {value, tag} packets written with
//            multimem.st (NVLS multicast: one store, the switch replicates) or per-peer stores.
//   phase 2  every CTA of every rank spins on the {value, tag} packets of its slab in ITS OWN memory
//            (no flag round trip, no fence: an 8-byte store is single-copy atomic), adds the W
//            contributions in rank order (bit-identical on all ranks), derives mean / invstd, and
//            normalizes its strip, which it just read and which is still in L1/L2.
// The same packet mechanism is the intra-GPU broadcast when world == 1, so there is no grid barrier.
// All CTAs must be co-resident: the host caps the grid at (SMs × occupancy).  Every spin is bounded.
#include "common.cuh"

namespace sod {
namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kSlabMax = 256;  // channels per slab → at most 32 lanes of 8 channels
constexpr int kMaxSlabs = 64;

struct BnGeom {
    int C, slabC, L, R, slabs, strips;
    long long rows, rows_per_strip;
};

struct BnWork {
    unsigned* counters;  // [kMaxSlabs]
    float* partials;     // [slabs][strips][16*L]
    uint2* ll_local;     // [2C] packets, used when world == 1
};

struct BnFwd {
    const void *x, *pre, *res;
    void* y;
    const float *gamma, *beta;
    float *rmean, *rvar, *smean, *sinvstd;
    float momentum, eps;
    int relu, training;
    BnGeom g;
    BnWork w;
    CommDev c;
    uint64_t stats_off;
    uint32_t tag;
    int use_mc;
};

struct BnBwd {
    const void *dy, *x, *pre, *y;
    void *dz, *dres;
    const float *gamma, *smean, *sinvstd;
    float *dgamma, *dbeta;
    int relu;
    BnGeom g;
    BnWork w;
    CommDev c;
    uint64_t stats_off;
    uint32_t tag;
    int use_mc;
};

// ---- packet exchange ---------------------------------------------------------------------------------
// entry index space of one layer call: [src_rank][2C]; within a slab the order is j = k*L + l with
// k in [0,16): k<8 → first statistic of channel l*8+k, k>=8 → second statistic of channel l*8+(k-8).
__device__ __forceinline__ void publish(const CommDev& c, int use_mc, uint64_t stats_off, uint2* ll_local, int C,
                                        int entry, float value, uint32_t tag) {
    const uint32_t bits = __float_as_uint(value);
    if (c.world == 1) {
        st_relaxed_sys_v2(ll_local + entry, bits, tag);
        return;
    }
    const uint64_t off = stats_off + (static_cast<uint64_t>(c.rank) * 2u * C + entry) * 8u;
    if (use_mc) {
        multimem_st_b64(reinterpret_cast<void*>(c.mc + off), bits, tag);
    } else {
        for (int q = 0; q < c.world; ++q) st_relaxed_sys_v2(reinterpret_cast<void*>(c.peer[q] + off), bits, tag);
    }
}

// sum over ranks (rank order) of entry `entry`; false on timeout
__device__ __forceinline__ bool collect(const CommDev& c, uint64_t stats_off, const uint2* ll_local, int C, int entry,
                                        uint32_t tag, unsigned long long timeout, float& out) {
    float acc = 0.f;
    const long long t0 = clock64();
    for (int q = 0; q < c.world; ++q) {
        const void* p = (c.world == 1)
                            ? static_cast<const void*>(ll_local + entry)
                            : reinterpret_cast<const void*>(c.peer[c.rank] + stats_off +
                                                            (static_cast<uint64_t>(q) * 2u * C + entry) * 8u);
        uint2 v = ld_relaxed_sys_v2(p);
        while (v.y != tag) {
            if (static_cast<unsigned long long>(clock64() - t0) > timeout) return false;
            v = ld_relaxed_sys_v2(p);
        }
        acc += __uint_as_float(v.x);
    }
    out = acc;
    return true;
}

// ---- CTA-level reduction of 16 per-thread accumulators over the row-lanes --------------------------
// on return threads j < 16*L hold (in `out`) the CTA total of entry j = k*L + l
__device__ __forceinline__ float cta_reduce16(float (&a)[16], int L, float* red /*[kWarps][16*L]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        float v = a[k];
        for (int o = 16; o >= L; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        a[k] = v;
    }
    const int n16 = 16 * L;
    if (lane < L) {
#pragma unroll
        for (int k = 0; k < 16; ++k) red[warp * n16 + k * L + lane] = a[k];
    }
    __syncthreads();
    float tot = 0.f;
    if (tid < n16) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) tot += red[w * n16 + tid];
    }
    return tot;
}

// Writes this CTA's partial, takes a ticket, and if last of its slab: reduces the slab's partials in
// fixed order. Returns true for the finisher, whose threads j < 16L then hold the slab total in `tot`.
__device__ __forceinline__ bool slab_finish(const BnGeom& g, const BnWork& w, int slab, int strip, float* red,
                                            float& tot) {
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int n16 = 16 * g.L;
    float* mine = w.partials + (static_cast<size_t>(slab) * g.strips + strip) * n16;
    if (g.strips == 1) return true;  // tot already is the slab total
    if (tid < n16) mine[tid] = tot;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(&w.counters[slab], 1u);
        s_last = (t == static_cast<unsigned>(g.strips - 1));
        if (s_last) w.counters[slab] = 0;  // everyone has arrived; ready for the next launch
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    // parts × n16 threads, each sums a residue class of strips; then a fixed-order combine
    const int parts = kThreads / n16;  // ≥ 1 (n16 ≤ 512)
    const int j = tid % n16, part = tid / n16;
    float acc = 0.f;
    if (part < parts) {
        const float* base = w.partials + static_cast<size_t>(slab) * g.strips * n16 + j;
        for (int t = part; t < g.strips; t += parts) acc += __ldcg(base + static_cast<size_t>(t) * n16);
    }
    __syncthreads();  // red is free again (cta_reduce16 readers are done: they passed the barriers above)
    if (part < parts) red[part * n16 + j] = acc;
    __syncthreads();
    tot = 0.f;
    if (tid < n16)
        for (int p = 0; p < parts; ++p) tot += red[p * n16 + tid];
    return true;
}

// =================================================================================================
// forward
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads) syncbn_fwd_kernel(const __grid_constant__ BnFwd prm) {
    __shared__ float red[kWarps * 16 * 32];  // 32 KB
    __shared__ float s_scale[kSlabMax], s_shift[kSlabMax];
    __shared__ int s_fail;

    const BnGeom& g = prm.g;
    const int tid = threadIdx.x;
    const int L = g.L, R = g.R, C = g.C;
    const int l = tid % L, rl = tid / L;
    const int slab = blockIdx.x % g.slabs, strip = blockIdx.x / g.slabs;
    const long long r0 = strip * g.rows_per_strip;
    const long long r1 = (r0 + g.rows_per_strip < g.rows) ? r0 + g.rows_per_strip : g.rows;
    const size_t coff = static_cast<size_t>(slab) * g.slabC + static_cast<size_t>(l) * 8;
    const T* __restrict__ x = static_cast<const T*>(prm.x);
    const T* __restrict__ pre = static_cast<const T*>(prm.pre);
    const T* __restrict__ res = static_cast<const T*>(prm.res);
    T* __restrict__ y = static_cast<T*>(prm.y);
    const int n16 = 16 * L;
    if (tid == 0) s_fail = 0;

    if (prm.training) {
        // ---- phase 1 ---------------------------------------------------------------------------------
        float a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = 0.f;
        long long r = r0 + rl;
        for (; r + 3ll * R < r1; r += 4ll * R) {
            float z[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) IO<T>::load8(x + (r + static_cast<long long>(u) * R) * C + coff, z[u]);
            if (pre) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float t[8];
                    IO<T>::load8(pre + (r + static_cast<long long>(u) * R) * C + coff, t);
#pragma unroll
                    for (int k = 0; k < 8; ++k) z[u][k] += t[k];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    a[k] += z[u][k];
                    a[8 + k] = fmaf(z[u][k], z[u][k], a[8 + k]);
                }
        }
        for (; r < r1; r += R) {
            float z[8];
            IO<T>::load8(x + r * C + coff, z);
            if (pre) {
                float t[8];
                IO<T>::load8(pre + r * C + coff, t);
#pragma unroll
                for (int k = 0; k < 8; ++k) z[k] += t[k];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a[k] += z[k];
                a[8 + k] = fmaf(z[k], z[k], a[8 + k]);
            }
        }
        float tot = cta_reduce16(a, L, red);
        if (slab_finish(g, prm.w, slab, strip, red, tot)) {
            if (tid < n16) publish(prm.c, prm.use_mc, prm.stats_off, prm.w.ll_local, C, slab * n16 + tid, tot, prm.tag);
        }
        // ---- phase 2a: global statistics -------------------------------------------------------------
        __syncthreads();
        if (tid < n16) {
            float v;
            if (!collect(prm.c, prm.stats_off, prm.w.ll_local, C, slab * n16 + tid, prm.tag,
                         prm.c.timeout_cycles ? prm.c.timeout_cycles : 4000000000ull, v)) {
                s_fail = 1;
                if (prm.c.error_flag) atomicExch(prm.c.error_flag, 0xDEAD0001u);
                v = 0.f;
            }
            red[tid] = v;
        }
        __syncthreads();
        if (s_fail) return;
        if (tid < g.slabC) {
            const int cl = tid, ll = cl >> 3, k = cl & 7, ch = slab * g.slabC + cl;
            const float n = static_cast<float>(g.rows) * static_cast<float>(prm.c.world);
            const float mean = red[k * L + ll] / n;
            const float var = fmaxf(red[(8 + k) * L + ll] / n - mean * mean, 0.f);
            const float invstd = 1.0f / sqrtf(var + prm.eps);
            const float sc = invstd * prm.gamma[ch];
            s_scale[cl] = sc;
            s_shift[cl] = prm.beta[ch] - mean * sc;
            if (strip == 0) {
                prm.smean[ch] = mean;
                prm.sinvstd[ch] = invstd;
                if (prm.rmean) {
                    const float unbiased = var * (n / fmaxf(n - 1.f, 1.f));
                    prm.rmean[ch] = (1.f - prm.momentum) * prm.rmean[ch] + prm.momentum * mean;
                    prm.rvar[ch] = (1.f - prm.momentum) * prm.rvar[ch] + prm.momentum * unbiased;
                }
            }
        }
    } else {
        if (tid < g.slabC) {
            const int ch = slab * g.slabC + tid;
            const float invstd = 1.0f / sqrtf(prm.rvar[ch] + prm.eps);
            const float sc = invstd * prm.gamma[ch];
            s_scale[tid] = sc;
            s_shift[tid] = prm.beta[ch] - prm.rmean[ch] * sc;
        }
    }
    __syncthreads();

    // ---- phase 2b: normalize / affine / residual / ReLU over the same strip -------------------------
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sc[k] = s_scale[l * 8 + k];
        sh[k] = s_shift[l * 8 + k];
    }
    const bool relu = prm.relu != 0;
    long long r = r0 + rl;
    for (; r + 3ll * R < r1; r += 4ll * R) {
        float z[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) IO<T>::load8(x + (r + static_cast<long long>(u) * R) * C + coff, z[u]);
        if (pre) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float t[8];
                IO<T>::load8(pre + (r + static_cast<long long>(u) * R) * C + coff, t);
#pragma unroll
                for (int k = 0; k < 8; ++k) z[u][k] += t[k];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) z[u][k] = fmaf(z[u][k], sc[k], sh[k]);
        if (res) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float t[8];
                IO<T>::load8(res + (r + static_cast<long long>(u) * R) * C + coff, t);
#pragma unroll
                for (int k = 0; k < 8; ++k) z[u][k] += t[k];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (relu) {
#pragma unroll
                for (int k = 0; k < 8; ++k) z[u][k] = fmaxf(z[u][k], 0.f);
            }
            IO<T>::store8(y + (r + static_cast<long long>(u) * R) * C + coff, z[u]);
        }
    }
    for (; r < r1; r += R) {
        float z[8];
        IO<T>::load8(x + r * C + coff, z);
        if (pre) {
            float t[8];
            IO<T>::load8(pre + r * C + coff, t);
#pragma unroll
            for (int k = 0; k < 8; ++k) z[k] += t[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) z[k] = fmaf(z[k], sc[k], sh[k]);
        if (res) {
            float t[8];
            IO<T>::load8(res + r * C + coff, t);
#pragma unroll
            for (int k = 0; k < 8; ++k) z[k] += t[k];
        }
        if (relu) {
#pragma unroll
            for (int k = 0; k < 8; ++k) z[k] = fmaxf(z[k], 0.f);
        }
        IO<T>::store8(y + r * C + coff, z);
    }
}

// =================================================================================================
// backward
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads) syncbn_bwd_kernel(const __grid_constant__ BnBwd prm) {
    __shared__ float red[kWarps * 16 * 32];
    __shared__ float s_a[kSlabMax], s_b[kSlabMax], s_d[kSlabMax];
    __shared__ int s_fail;

    const BnGeom& g = prm.g;
    const int tid = threadIdx.x;
    const int L = g.L, R = g.R, C = g.C;
    const int l = tid % L, rl = tid / L;
    const int slab = blockIdx.x % g.slabs, strip = blockIdx.x / g.slabs;
    const long long r0 = strip * g.rows_per_strip;
    const long long r1 = (r0 + g.rows_per_strip < g.rows) ? r0 + g.rows_per_strip : g.rows;
    const size_t coff = static_cast<size_t>(slab) * g.slabC + static_cast<size_t>(l) * 8;
    const int ch0 = slab * g.slabC + l * 8;
    const T* __restrict__ dy = static_cast<const T*>(prm.dy);
    const T* __restrict__ x = static_cast<const T*>(prm.x);
    const T* __restrict__ pre = static_cast<const T*>(prm.pre);
    const T* __restrict__ yy = static_cast<const T*>(prm.y);
    T* __restrict__ dz = static_cast<T*>(prm.dz);
    T* __restrict__ dres = static_cast<T*>(prm.dres);
    const int n16 = 16 * L;
    const bool relu = prm.relu != 0;
    if (tid == 0) s_fail = 0;

    float mean[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) mean[k] = prm.smean[ch0 + k];

    // ---- phase 1: Σ dy_m and Σ dy_m (z - mean) ----------------------------------------------------------
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = 0.f;
    for (long long r = r0 + rl; r < r1; r += 2ll * R) {
        float d[2][8], z[2][8];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long rr = r + static_cast<long long>(u) * R;
            ok[u] = rr < r1;
            if (ok[u]) {
                IO<T>::load8(dy + rr * C + coff, d[u]);
                IO<T>::load8(x + rr * C + coff, z[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long rr = r + static_cast<long long>(u) * R;
            if (ok[u]) {
                if (pre) {
                    float t[8];
                    IO<T>::load8(pre + rr * C + coff, t);
#pragma unroll
                    for (int k = 0; k < 8; ++k) z[u][k] += t[k];
                }
                if (relu) {
                    float o[8];
                    IO<T>::load8(yy + rr * C + coff, o);
#pragma unroll
                    for (int k = 0; k < 8; ++k) d[u][k] = o[k] > 0.f ? d[u][k] : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    a[k] += d[u][k];
                    a[8 + k] = fmaf(d[u][k], z[u][k] - mean[k], a[8 + k]);
                }
            }
        }
    }
    float tot = cta_reduce16(a, L, red);
    if (slab_finish(g, prm.w, slab, strip, red, tot)) {
        if (tid < n16) {
            // local parameter gradients (the gradient all-reduce averages them later)
            const int k = tid / L, ll = tid % L;
            const int ch = slab * g.slabC + ll * 8 + (k & 7);
            if (k < 8) prm.dbeta[ch] = tot;
            else prm.dgamma[ch] = tot * prm.sinvstd[ch];
            publish(prm.c, prm.use_mc, prm.stats_off, prm.w.ll_local, C, slab * n16 + tid, tot, prm.tag);
        }
    }
    __syncthreads();
    if (tid < n16) {
        float v;
        if (!collect(prm.c, prm.stats_off, prm.w.ll_local, C, slab * n16 + tid, prm.tag,
                     prm.c.timeout_cycles ? prm.c.timeout_cycles : 4000000000ull, v)) {
            s_fail = 1;
            if (prm.c.error_flag) atomicExch(prm.c.error_flag, 0xDEAD0002u);
            v = 0.f;
        }
        red[tid] = v;
    }
    __syncthreads();
    if (s_fail) return;
    if (tid < g.slabC) {
        const int cl = tid, ll = cl >> 3, k = cl & 7, ch = slab * g.slabC + cl;
        const float n = static_cast<float>(g.rows) * static_cast<float>(prm.c.world);
        const float mean_dy = red[k * L + ll] / n;
        const float mean_dy_xmu = red[(8 + k) * L + ll] / n;
        const float invstd = prm.sinvstd[ch];
        const float A = prm.gamma[ch] * invstd;
        const float B = -A * invstd * invstd * mean_dy_xmu;
        s_a[cl] = A;
        s_b[cl] = B;
        s_d[cl] = -A * mean_dy - B * prm.smean[ch];
    }
    __syncthreads();

    // ---- phase 2: dz = A dy_m + B z + D ; dres = dy_m ----------------------------------------------------
    float A[8], B[8], D[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        A[k] = s_a[l * 8 + k];
        B[k] = s_b[l * 8 + k];
        D[k] = s_d[l * 8 + k];
    }
    for (long long r = r0 + rl; r < r1; r += 2ll * R) {
        float d[2][8], z[2][8];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long rr = r + static_cast<long long>(u) * R;
            ok[u] = rr < r1;
            if (ok[u]) {
                IO<T>::load8(dy + rr * C + coff, d[u]);
                IO<T>::load8(x + rr * C + coff, z[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long rr = r + static_cast<long long>(u) * R;
            if (ok[u]) {
                if (pre) {
                    float t[8];
                    IO<T>::load8(pre + rr * C + coff, t);
#pragma unroll
                    for (int k = 0; k < 8; ++k) z[u][k] += t[k];
                }
                if (relu) {
                    float o[8];
                    IO<T>::load8(yy + rr * C + coff, o);
#pragma unroll
                    for (int k = 0; k < 8; ++k) d[u][k] = o[k] > 0.f ? d[u][k] : 0.f;
                }
                if (dres) IO<T>::store8(dres + rr * C + coff, d[u]);
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = fmaf(A[k], d[u][k], fmaf(B[k], z[u][k], D[k]));
                IO<T>::store8(dz + rr * C + coff, o);
            }
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------
static int make_geom(int64_t rows, int C, int dtype, int max_ctas, BnGeom& g) {
    if (rows <= 0 || C <= 0 || (C % 8) != 0) return SOD_EUNSUPPORTED;
    (void)dtype;
    g.C = C;
    g.rows = rows;
    g.slabC = C < kSlabMax ? C : kSlabMax;
    if (C % g.slabC) return SOD_EUNSUPPORTED;
    g.L = g.slabC / 8;
    if (g.L & (g.L - 1)) return SOD_EUNSUPPORTED;  // lanes per row must be a power of two (≤32)
    g.R = kThreads / g.L;
    g.slabs = C / g.slabC;
    if (g.slabs > kMaxSlabs) return SOD_EUNSUPPORTED;
    long long strips = (rows + 4ll * g.R - 1) / (4ll * g.R);  // ≥ ~4 row-iterations per CTA
    long long cap = max_ctas / g.slabs;
    if (cap < 1) cap = 1;
    if (strips > cap) strips = cap;
    if (strips < 1) strips = 1;
    g.rows_per_strip = (rows + strips - 1) / strips;
    g.strips = static_cast<int>((rows + g.rows_per_strip - 1) / g.rows_per_strip);
    return SOD_OK;
}

template <typename K>
static int max_resident(K kern) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * dev_info().sm_count;
}

static size_t bn_ws_layout(const BnGeom* g, int C, BnWork* w, void* base) {
    // [counters 256 B][ll_local 2C*8][partials]
    size_t off = 0;
    if (w) w->counters = reinterpret_cast<unsigned*>(static_cast<char*>(base) + off);
    off += kMaxSlabs * sizeof(unsigned);
    if (w) w->ll_local = reinterpret_cast<uint2*>(static_cast<char*>(base) + off);
    off += static_cast<size_t>(2) * C * sizeof(uint2);
    if (w) w->partials = reinterpret_cast<float*>(static_cast<char*>(base) + off);
    if (g) off += static_cast<size_t>(g->slabs) * g->strips * 16 * g->L * sizeof(float);
    return off;
}

}  // namespace
}  // namespace sod

extern "C" size_t sod_syncbn_workspace_bytes(int64_t rows, int channels) {
    using namespace sod;
    (void)rows;
    // upper bound independent of the device: every CTA of a 2-CTA/SM grid on ≤ 256 SMs writes 2*slabC floats
    const size_t slabC = channels < kSlabMax ? channels : kSlabMax;
    return kMaxSlabs * sizeof(unsigned) + static_cast<size_t>(2) * channels * sizeof(uint2) +
           static_cast<size_t>(1024) * 2 * slabC * sizeof(float);
}

extern "C" size_t sod_syncbn_exchange_bytes(int channels) {
    return static_cast<size_t>(SOD_MAX_WORLD) * 2u * channels * 8u;
}

extern "C" int sod_syncbn_fwd(const void* x, const void* pre_add, const void* residual, void* y, int dtype,
                              const float* gamma, const float* beta, float* running_mean, float* running_var,
                              float* save_mean, float* save_invstd, int64_t rows, int channels, float momentum,
                              float eps, int relu, int training, const sod_comm* comm, uint64_t stats_off,
                              uint32_t seq, void* workspace, size_t workspace_bytes, int flags, void* stream) {
    using namespace sod;
    SOD_CHECK_ARG(x && y && gamma && beta && workspace, SOD_EINVAL);
    SOD_CHECK_ARG(training ? (save_mean && save_invstd) : (running_mean && running_var), SOD_EINVAL);
    SOD_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), SOD_EINVAL);
    SOD_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(workspace) && (!pre_add || aligned16(pre_add)) &&
                      (!residual || aligned16(residual)), SOD_EALIGN);
    if (dev_info().cc_major != 10) return SOD_EUNSUPPORTED;
    BnFwd p{};
    int rc = make_comm_dev(training ? comm : nullptr, p.c);
    if (rc != SOD_OK) return rc;
    if (p.c.world > 1) {
        SOD_CHECK_ARG((stats_off & 15) == 0, SOD_EALIGN);
        SOD_CHECK_ARG(stats_off >= sod_comm_flag_bytes() &&
                          stats_off + sod_syncbn_exchange_bytes(channels) <= comm->arena_bytes, SOD_ECOMM);
    }
    return SOD_DISPATCH_DTYPE(dtype, T, [&]() -> int {
        auto kern = syncbn_fwd_kernel<T>;
        rc = make_geom(rows, channels, dtype, max_resident(kern), p.g);
        if (rc != SOD_OK) return rc;
        if (bn_ws_layout(&p.g, channels, &p.w, workspace) > workspace_bytes) return SOD_EWORKSPACE;
        p.x = x; p.pre = pre_add; p.res = residual; p.y = y;
        p.gamma = gamma; p.beta = beta; p.rmean = running_mean; p.rvar = running_var;
        p.smean = save_mean; p.sinvstd = save_invstd;
        p.momentum = momentum; p.eps = eps; p.relu = relu; p.training = training;
        p.stats_off = stats_off; p.tag = seq;
        p.use_mc = (p.c.mc != 0) && !(flags & SOD_ALGO_NO_MULTIMEM);
        kern<<<p.g.slabs * p.g.strips, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
        return static_cast<int>(cudaGetLastError());
    });
}

extern "C" int sod_syncbn_bwd(const void* dy, const void* x, const void* pre_add, const void* y, void* dz, void* dres,
                              int dtype, const float* gamma, const float* save_mean, const float* save_invstd,
                              float* dgamma, float* dbeta, int64_t rows, int channels, int relu, const sod_comm* comm,
                              uint64_t stats_off, uint32_t seq, void* workspace, size_t workspace_bytes, int flags,
                              void* stream) {
    using namespace sod;
    SOD_CHECK_ARG(dy && x && dz && gamma && save_mean && save_invstd && dgamma && dbeta && workspace, SOD_EINVAL);
    SOD_CHECK_ARG(!relu || y, SOD_EINVAL);
    SOD_CHECK_ARG(aligned16(dy) && aligned16(x) && aligned16(dz) && aligned16(workspace) &&
                      (!pre_add || aligned16(pre_add)) && (!y || aligned16(y)) && (!dres || aligned16(dres)), SOD_EALIGN);
    if (dev_info().cc_major != 10) return SOD_EUNSUPPORTED;
    BnBwd p{};
    int rc = make_comm_dev(comm, p.c);
    if (rc != SOD_OK) return rc;
    if (p.c.world > 1) {
        SOD_CHECK_ARG((stats_off & 15) == 0, SOD_EALIGN);
        SOD_CHECK_ARG(stats_off >= sod_comm_flag_bytes() &&
                          stats_off + sod_syncbn_exchange_bytes(channels) <= comm->arena_bytes, SOD_ECOMM);
    }
    return SOD_DISPATCH_DTYPE(dtype, T, [&]() -> int {
        auto kern = syncbn_bwd_kernel<T>;
        rc = make_geom(rows, channels, dtype, max_resident(kern), p.g);
        if (rc != SOD_OK) return rc;
        if (bn_ws_layout(&p.g, channels, &p.w, workspace) > workspace_bytes) return SOD_EWORKSPACE;
        p.dy = dy; p.x = x; p.pre = pre_add; p.y = y; p.dz = dz; p.dres = dres;
        p.gamma = gamma; p.smean = save_mean; p.sinvstd = save_invstd; p.dgamma = dgamma; p.dbeta = dbeta;
        p.relu = relu; p.stats_off = stats_off; p.tag = seq;
        p.use_mc = (p.c.mc != 0) && !(flags & SOD_ALGO_NO_MULTIMEM);
        kern<<<p.g.slabs * p.g.strips, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
        return static_cast<int>(cudaGetLastError());
    });
}
