// sm_100a kernels + C ABI of the batched Overcooked simulator (include/overcooked_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/overcooked_b200.h"
#include "oc_device.cuh"
#include "oc_host.hpp"

using namespace ock;

// =============================================================================================
// kernels
// =============================================================================================

// A/B knob (compile time): -DOC_STEP_MIN_CTAS=n adds a minimum-CTAs-per-SM hint to the step kernel, i.e. a register
// cap of 65536 / (256 n); unset = ptxas' own choice (see profiles/r1_ptxas_sass.txt for what it picks)
#ifdef OC_STEP_MIN_CTAS
#define OC_STEP_BOUNDS __launch_bounds__(256, OC_STEP_MIN_CTAS)
#else
#define OC_STEP_BOUNDS __launch_bounds__(256)
#endif

// dynamic shared memory: [table blob][per warp: nb env rows (float or biased-byte format)]
template <int A, int NOBJ, int NF, int MODE>
__global__ void OC_STEP_BOUNDS
oc_step_kernel(const __grid_constant__ OcParams p, uint4* __restrict__ state,
               const int32_t* __restrict__ actions, float* __restrict__ obs,
               float* __restrict__ rew32, double* __restrict__ rew64, uint8_t* __restrict__ done_out,
               float* __restrict__ term_obs, uint32_t flags) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // Programmatic dependent launch: let the NEXT launch in the stream start its own prologue as
    // soon as every CTA of this grid is running; do our prologue (tables, row clear) before
    // waiting for the previous grid -- only what follows the wait touches its outputs (state,
    // actions).  Both instructions are no-ops when the launch carries no PDL attribute.
    asm volatile("griddepcontrol.launch_dependents;");
    load_tables(p, smem);
    uint8_t* wrows = smem + p.blob_bytes + (size_t)warp * p.warp_row_bytes;
    warp_clear_rows<MODE != 0>(wrows, p.warp_row_bytes, lane);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    const Tables tb = make_tables(p, smem);

    // the grid is sized to ONE resident wave (148 SMs x CTAs that fit); with more envs than that
    // each CTA walks several chunks so the tables are loaded once per CTA, not once per chunk
    bool first = true;                              // rows still clear from the prologue
    for (int base = blockIdx.x * blockDim.x; base < p.E; base += gridDim.x * blockDim.x) {
        const int env = base + threadIdx.x;
        const bool valid = env < p.E;
        Env<A, NOBJ> e;
        Info in;
        bool done = false;
        if (valid) {                                    // dynamics: no row access, may overlap the previous chunk's TMA read
            int nav[A], comm[A];
            load_env<A, NOBJ>(e, state, p.E, env);
            const int2* a2 = reinterpret_cast<const int2*>(actions) + (size_t)env * A;
#pragma unroll
            for (int k = 0; k < A; ++k) { const int2 v = __ldg(a2 + k); nav[k] = v.x & 3; comm[k] = v.y; }
            in = step_logic<A, NOBJ, NF>(e, p, tb, nav, comm[0], comm[1], (uint32_t)env, rew32, rew64, done_out, done);
        }
        const bool fin = valid && done && (flags & OC_FLAG_AUTO_RESET);
        if (term_obs != nullptr && __any_sync(0xFFFFFFFFu, fin))      // rare: some env of this warp finished
            warp_terminal_obs<A, NOBJ, NF, MODE>(e, in, fin, p, tb, wrows, lane, term_obs + (size_t)env * p.row_bytes);
        if (fin) {
            finish_episode<A, NOBJ>(e, p, tb, (uint32_t)env);
            in = gather_info<A, NOBJ, NF>(e, p, tb);
        }
        if (valid) store_env<A, NOBJ>(e, state, p.E, env);
        const int env0 = base + warp * 32;
        emit_obs<A, NOBJ, NF, MODE>(e, in, valid, p, tb, wrows, lane, obs + (size_t)env0 * p.row_bytes, min(32, p.E - env0), first);
        first = false;
    }
    rows_wait_done(p);
}

// n_steps steps per launch, state in registers, Philox actions (SURVEY section 8d synthetic inputs)
template <int A, int NOBJ, int NF, int MODE>
__global__ void __launch_bounds__(256)
oc_rollout_kernel(const __grid_constant__ OcParams p, uint4* __restrict__ state, int n_steps, uint32_t step0,
                  float* __restrict__ obs, float* __restrict__ rew32, uint8_t* __restrict__ done_out,
                  int32_t* __restrict__ actions_out, const int32_t* __restrict__ actions_in) {
    constexpr bool ROWF = MODE != 0;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = env < p.E;
    Env<A, NOBJ> e;
    if (valid) load_env<A, NOBJ>(e, state, p.E, env);
    load_tables(p, smem);
    uint8_t* wrows = smem + p.blob_bytes + (size_t)warp * p.warp_row_bytes;
    __syncthreads();
    const Tables tb = make_tables(p, smem);
    const int env0 = blockIdx.x * blockDim.x + warp * 32;
    const int nvalid = min(32, p.E - env0);
    const size_t step_floats = (size_t)p.E * p.row_bytes;

    for (int s = 0; s < n_steps; ++s) {
        // dynamics first: they do not touch the rows, so the copy engine may still be reading the
        // previous step's rows out of shared memory while this runs
        Info in;
        if (valid) in = rollout_logic<A, NOBJ, NF, ROWF>(e, p, tb, (uint32_t)env, (uint32_t)s, step0, rew32, done_out, actions_out, actions_in);
        if (obs != nullptr) {
            emit_obs<A, NOBJ, NF, MODE>(e, in, valid, p, tb, wrows, lane,
                                    obs + (size_t)s * step_floats + (size_t)env0 * p.row_bytes, nvalid);
        }
    }
    if (valid) store_env<A, NOBJ>(e, state, p.E, env);
    rows_wait_done(p);
}

// reset (masked) + observation of every env
template <int A, int NOBJ, int NF, int MODE>
__global__ void __launch_bounds__(256)
oc_reset_kernel(const __grid_constant__ OcParams p, uint4* __restrict__ state, const uint8_t* __restrict__ mask,
                const int32_t* __restrict__ placements, float* __restrict__ obs, int initial) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = env < p.E;
    Env<A, NOBJ> e;
    if (valid && !initial) load_env<A, NOBJ>(e, state, p.E, env);
    load_tables(p, smem);
    uint8_t* wrows = smem + p.blob_bytes + (size_t)warp * p.warp_row_bytes;
    __syncthreads();
    const Tables tb = make_tables(p, smem);
    Info in;
    if (valid) {
        reset_logic<A, NOBJ>(e, p, tb, (uint32_t)env, initial != 0, mask, placements);
        store_env<A, NOBJ>(e, state, p.E, env);
        if (obs != nullptr) in = gather_info<A, NOBJ, NF>(e, p, tb);
    }
    const int env0 = blockIdx.x * blockDim.x + warp * 32;
    if (obs != nullptr)
        emit_obs<A, NOBJ, NF, MODE>(e, in, valid, p, tb, wrows, lane, obs + (size_t)env0 * p.row_bytes, min(32, p.E - env0));
    rows_wait_done(p);
}

// packed state <-> [E, 16] u32 rows
__global__ void oc_state_export_kernel(const uint4* __restrict__ planes, uint32_t* __restrict__ rows, int E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * 4) return;
    const int env = i >> 2, pl = i & 3;
    reinterpret_cast<uint4*>(rows)[(size_t)env * 4 + pl] = planes[(size_t)pl * E + env];
}
__global__ void oc_state_import_kernel(uint4* __restrict__ planes, const uint32_t* __restrict__ rows, int E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * 4) return;
    const int env = i >> 2, pl = i & 3;
    planes[(size_t)pl * E + env] = reinterpret_cast<const uint4*>(rows)[(size_t)env * 4 + pl];
}
__global__ void oc_stats_kernel(const uint4* __restrict__ planes, uint32_t* __restrict__ episodes,
                                uint32_t* __restrict__ last_completed, int E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    if (episodes) episodes[i] = planes[i].y;
    if (last_completed) last_completed[i] = planes[(size_t)E + i].y & 0xFFu;
}

// float rows [E, A, F] -> int8 rows [E, A, F-1] + clock f32 [E]; one thread per output word
__global__ void __launch_bounds__(256)
oc_pack_i8_kernel(const __grid_constant__ OcParams p, const float* __restrict__ obs, int8_t* __restrict__ out,
                  float* __restrict__ ts, uint32_t nwords) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nwords) pack_i8_word(p, obs, out, ts, w);
}

// rows of the n finished envs idx[0..n) -> dense [n, ...] buffer; one 128-thread CTA per row
__global__ void __launch_bounds__(128)
oc_gather_term_kernel(const __grid_constant__ OcParams p, const float* __restrict__ term, const int32_t* __restrict__ idx,
                      int n, float* __restrict__ out_f32, int8_t* __restrict__ out_i8, float* __restrict__ out_ts) {
    for (int i = blockIdx.x; i < n; i += gridDim.x)
        gather_term_row(p, term, idx[i], i, out_f32, out_i8, out_ts, (int)threadIdx.x, (int)blockDim.x);
}

// =============================================================================================
// host side
// =============================================================================================
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fail(OC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
    } while (0)

struct oc_env {
    OcParams p;
    int device = 0;
    int threads = 64;
    size_t smem_bytes = 0;
    uint4* state = nullptr;
    uint8_t* blob = nullptr;
    float* ts = nullptr;
    uint64_t launches = 0;
    uint32_t rollout_step = 0;
    int pdl = 1;
    int tma_rows_in_step = 1;
    int step_grid = 1;
    int obs_off[OC_NUM_OBS_KEYS], obs_size[OC_NUM_OBS_KEYS];
    // device-side staging of the host-buffer entry points (oc_step_host / oc_reset_host), allocated on first use
    struct HostPath {
        int32_t* actions = nullptr; float* obs = nullptr; float* rew32 = nullptr; double* rew64 = nullptr;
        uint8_t* done = nullptr; float* term = nullptr; uint8_t* mask = nullptr; int32_t* place = nullptr;
        int8_t* obs8 = nullptr; float* ts = nullptr;          // compact format (oc_*_host_i8)
        // terminal rows: indices of the finished envs (host -> device), their rows gathered densely (device -> host)
        int32_t* idx = nullptr; float* gather = nullptr; float* gather_ts = nullptr;
        int32_t* h_idx = nullptr; float* h_gather = nullptr; float* h_gather_ts = nullptr;   // pinned
    } hp;
};

// kernel template MODE: 0 byte rows, 1 float rows (all 32 envs of a warp in one pass), 2 float rows in passes
static int row_mode(const OcParams& p) { return !p.rowf ? 0 : (p.obs_passes > 1 ? 2 : 1); }

template <int N, int NFOOD> constexpr int shape_nobj(std::integer_sequence<int, N, NFOOD>) { return N; }
template <int N, int NFOOD> constexpr int shape_nf(std::integer_sequence<int, N, NFOOD>) { return NFOOD; }

template <typename F>
static int dispatch(int A, int NOBJ, int mode, F&& f) {
    // shapes: (2 object slots, 1 food channel), (4, 2), (6, 3) -- chosen by compile_config
#define OC_CASE(a, n, nfood)                                                                                            \
    if (A == a && NOBJ == n) {                                                                                          \
        using IA = std::integral_constant<int, a>;                                                                      \
        using IS = std::integer_sequence<int, n, nfood>;                                                                \
        if (mode == 2) return f(IA(), IS(), std::integral_constant<int, 2>());                                          \
        if (mode == 1) return f(IA(), IS(), std::integral_constant<int, 1>());                                          \
        return f(IA(), IS(), std::integral_constant<int, 0>());                                                         \
    }
    OC_CASE(2, 2, 1) OC_CASE(3, 2, 1) OC_CASE(4, 2, 1) OC_CASE(2, 4, 2) OC_CASE(3, 4, 2) OC_CASE(4, 4, 2)
    OC_CASE(2, 6, 3) OC_CASE(3, 6, 3) OC_CASE(4, 6, 3)
#undef OC_CASE
    return fail(OC_ERR_INVALID, "unsupported (num_agents, num_objects)");
}

extern "C" int oc_abi_version(void) { return OC_ABI_VERSION; }
extern "C" const char* oc_last_error(void) { return g_err.c_str(); }

extern "C" int oc_create(const oc_config* c, oc_env** out) {
    if (!c || !out) return fail(OC_ERR_INVALID, "null argument");
    *out = nullptr;
    oc_env* h = new (std::nothrow) oc_env();
    if (!h) return fail(OC_ERR_ALLOC, "out of host memory");
    HostImage img;
    std::string err;
    if (compile_config(c, img, err) != OC_OK) { delete h; return fail(OC_ERR_INVALID, err); }
    h->p = img.p;
    memcpy(h->obs_off, img.obs_off, sizeof(h->obs_off));
    memcpy(h->obs_size, img.obs_size, sizeof(h->obs_size));
    OcParams& p = h->p;
    const std::vector<uint8_t>& blob = img.blob;
    const std::vector<float>& ts = img.ts;
    int dev = 0;
    {   // there is no CPU fallback: without a CUDA device creation fails
        cudaError_t ce0 = cudaGetDevice(&dev);
        if (ce0 != cudaSuccess) { delete h; return fail(OC_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(ce0)); }
    }
    h->device = dev;

    if (const char* pe = getenv("OC_PDL")) h->pdl = atoi(pe) != 0;
    if (const char* te = getenv("OC_TMA")) h->tma_rows_in_step = atoi(te) != 0;
    cudaDeviceProp prop;
    {
        cudaError_t cep = cudaGetDeviceProperties(&prop, dev);
        if (cep != cudaSuccess) { delete h; return fail(OC_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(cep)); }
    }
    const int num_sm = prop.multiProcessorCount;
    const size_t smem_cta_max = prop.sharedMemPerBlockOptin;
    auto smem_for = [&](int threads) { return (size_t)p.blob_bytes + (size_t)(threads / 32) * (size_t)p.warp_row_bytes; };

    // opt in to large dynamic shared memory for the instantiations we launch.  The attribute is per
    // function, not per handle, so it is always set to the device maximum: a second handle with a
    // smaller footprint must not lower it under the first one's feet.  Then ask the runtime how many
    // CTAs of each candidate size are resident per SM (shared memory AND registers).
    const int smem_optin = (int)smem_cta_max;
    int caps[9] = {0};
    int rc = dispatch(p.A, p.NOBJ, row_mode(p), [&](auto a, auto nobj, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        constexpr int RF = decltype(rf)::value;
        CUDA_TRY(cudaFuncSetAttribute(oc_step_kernel<AA, NN, FF, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(oc_rollout_kernel<AA, NN, FF, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(oc_reset_kernel<AA, NN, FF, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
        for (int t = 32; t <= 256; t += 32) {
            if (smem_for(t) > smem_cta_max) continue;
            int cs = 0, cr = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cs, oc_step_kernel<AA, NN, FF, RF>, t, smem_for(t)));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cr, oc_rollout_kernel<AA, NN, FF, RF>, t, smem_for(t)));
            caps[t / 32] = std::min(cs, cr);
        }
        return OC_OK;
    });
    if (rc != OC_OK) { std::string m = g_err; delete h; return fail(rc, m); }

    // CTA shape: the work is one warp per 32 envs; pick the CTA size whose resident wave covers the
    // envs with the smallest makespan, preferring fewer table copies on ties.
    const char* tenv = getenv("OC_BLOCK_THREADS");
    // resident warps per SM that hide the latencies: byte rows run LDS -> convert -> STG chains, float rows
    // leave through the copy engine, multi-pass float rows wait on it most of the time
    const double want_warps = !p.rowf ? 24.0 : (p.obs_passes > 1 ? 6.0 : 14.0);
    int best_t = 0, best_cap = 1; double best_cost = 1e30;
    for (int t = 32; t <= 256; t += 32) {
        if (tenv && atoi(tenv) != t) continue;
        const int cap = caps[t / 32];
        if (cap < 1) continue;
        const long long ctas = ((long long)p.E + t - 1) / t;
        const long long per_sm = (ctas + num_sm - 1) / num_sm;              // CTAs of work on the busiest SM
        const double conc = (double)std::min<long long>(per_sm, cap) * (t / 32); // warps resident together
        const double eff = std::min(1.0, conc / want_warps);
        double cost = (double)per_sm * (t / 32) / eff + 1e-3 * (double)(256 - t) / 256.0;
        // several waves through ONE resident CTA per SM: the SM idles while each new CTA loads its tables
        if (per_sm > cap && cap < 2) cost *= 1.25;
        if (cost < best_cost) { best_cost = cost; best_t = t; best_cap = cap; }
    }
    if (best_t == 0) { delete h; return fail(OC_ERR_INVALID, "observation row too wide for shared memory"); }
    h->threads = best_t;
    h->smem_bytes = smem_for(best_t);
    h->step_grid = (int)std::min<long long>(((long long)p.E + best_t - 1) / best_t, (long long)num_sm * best_cap);

    cudaError_t ce;
    if ((ce = cudaMalloc(&h->state, (size_t)p.E * 64)) != cudaSuccess ||
        (ce = cudaMalloc(&h->blob, blob.size())) != cudaSuccess ||
        (ce = cudaMalloc(&h->ts, ts.size() * 4)) != cudaSuccess ||
        (ce = cudaMemcpy(h->blob, blob.data(), blob.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (ce = cudaMemcpy(h->ts, ts.data(), ts.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
        std::string m = std::string("device allocation/copy failed: ") + cudaGetErrorString(ce);
        oc_destroy(h);
        return fail(ce == cudaErrorMemoryAllocation ? OC_ERR_ALLOC : OC_ERR_CUDA, m);
    }
    p.blob = h->blob; p.ts_table = h->ts;

    rc = dispatch(p.A, p.NOBJ, row_mode(p), [&](auto a, auto nobj, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        constexpr int RF = decltype(rf)::value;
        const int grid = (p.E + h->threads - 1) / h->threads;
        oc_reset_kernel<AA, NN, FF, RF><<<grid, h->threads, h->smem_bytes, 0>>>(p, h->state, nullptr, nullptr, nullptr, 1);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
        return OC_OK;
    });
    if (rc != OC_OK) { std::string m = g_err; oc_destroy(h); return fail(rc, m); }
    h->launches += 1;
    *out = h;
    return OC_OK;
}

extern "C" int oc_destroy(oc_env* h) {
    if (!h) return OC_OK;
    if (h->state) cudaFree(h->state);
    if (h->blob) cudaFree(h->blob);
    if (h->ts) cudaFree(h->ts);
    for (void* q : {(void*)h->hp.actions, (void*)h->hp.obs, (void*)h->hp.rew32, (void*)h->hp.rew64, (void*)h->hp.done,
                    (void*)h->hp.term, (void*)h->hp.mask, (void*)h->hp.place, (void*)h->hp.obs8, (void*)h->hp.ts,
                    (void*)h->hp.idx, (void*)h->hp.gather, (void*)h->hp.gather_ts})
        if (q) cudaFree(q);
    for (void* q : {(void*)h->hp.h_idx, (void*)h->hp.h_gather, (void*)h->hp.h_gather_ts})
        if (q) cudaFreeHost(q);
    delete h;
    return OC_OK;
}

extern "C" int oc_obs_width(const oc_env* h) { return h ? h->p.F : OC_ERR_INVALID; }

extern "C" int oc_obs_layout(const oc_env* h, int32_t* offsets, int32_t* sizes) {
    if (!h || !offsets || !sizes) return fail(OC_ERR_INVALID, "null argument");
    for (int i = 0; i < OC_NUM_OBS_KEYS; ++i) { offsets[i] = h->obs_off[i]; sizes[i] = h->obs_size[i]; }
    return OC_OK;
}

static bool misaligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) != 0; }

// a handle is bound to the device that was current at oc_create
static int check_device(const oc_env* h) {
    int cur = -1;
    CUDA_TRY(cudaGetDevice(&cur));
    if (cur != h->device) return fail(OC_ERR_INVALID, "handle was created on device " + std::to_string(h->device) +
                                                      " but device " + std::to_string(cur) + " is current");
    return OC_OK;
}

extern "C" int oc_reset(oc_env* h, const uint8_t* mask, const int32_t* placements, float* obs, void* stream) {
    if (!h) return fail(OC_ERR_INVALID, "null handle");
    if (obs && misaligned16(obs)) return fail(OC_ERR_INVALID, "obs must be 16-byte aligned");
    if (int dc = check_device(h)) return dc;
    const OcParams& p = h->p;
    int rc = dispatch(p.A, p.NOBJ, row_mode(p), [&](auto a, auto nobj, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        constexpr int RF = decltype(rf)::value;
        const int grid = (p.E + h->threads - 1) / h->threads;
        oc_reset_kernel<AA, NN, FF, RF><<<grid, h->threads, h->smem_bytes, (cudaStream_t)stream>>>(p, h->state, mask, placements, obs, 0);
        CUDA_TRY(cudaGetLastError());
        return OC_OK;
    });
    if (rc == OC_OK) h->launches += 1;
    return rc;
}

extern "C" int oc_step(oc_env* h, const int32_t* actions, float* obs, float* rew_f32, double* rew_f64,
                       uint8_t* done, float* term_obs, uint32_t flags, void* stream) {
    if (!h || !actions || !obs || !done) return fail(OC_ERR_INVALID, "null argument");
    if (misaligned16(obs) || (reinterpret_cast<uintptr_t>(actions) & 7)) return fail(OC_ERR_INVALID, "obs must be 16-byte and actions 8-byte aligned");
    if (int dc = check_device(h)) return dc;
    const OcParams& p = h->p;
    int rc = dispatch(p.A, p.NOBJ, row_mode(p), [&](auto a, auto nobj, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        constexpr int RF = decltype(rf)::value;
        const int grid = h->step_grid;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(h->threads);
        cfg.dynamicSmemBytes = h->smem_bytes; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = h->pdl ? 1 : 0;
        OcParams ps = p;
        if (ps.rowf && !ps.use_tma && h->tma_rows_in_step) ps.use_tma = 2;      // padded float rows: per-row bulk copies
        CUDA_TRY(cudaLaunchKernelEx(&cfg, oc_step_kernel<AA, NN, FF, RF>, ps, h->state, actions, obs, rew_f32, rew_f64,
                                    done, term_obs, flags));
        return OC_OK;
    });
    if (rc == OC_OK) h->launches += 1;
    return rc;
}

static int launch_rollout(oc_env* h, int32_t n_steps, float* obs, float* rew_f32, uint8_t* done,
                          int32_t* actions_out, const int32_t* actions_in, void* stream);

extern "C" int oc_rollout(oc_env* h, int32_t n_steps, float* obs, float* rew_f32, uint8_t* done,
                          int32_t* actions_out, void* stream) {
    return launch_rollout(h, n_steps, obs, rew_f32, done, actions_out, nullptr, stream);
}

extern "C" int oc_replay(oc_env* h, int32_t n_steps, const int32_t* actions, float* obs, float* rew_f32,
                         uint8_t* done, void* stream) {
    if (!actions) return fail(OC_ERR_INVALID, "null actions");
    if (reinterpret_cast<uintptr_t>(actions) & 7) return fail(OC_ERR_INVALID, "actions must be 8-byte aligned");
    return launch_rollout(h, n_steps, obs, rew_f32, done, nullptr, actions, stream);
}

static int launch_rollout(oc_env* h, int32_t n_steps, float* obs, float* rew_f32, uint8_t* done,
                          int32_t* actions_out, const int32_t* actions_in, void* stream) {
    if (!h || n_steps <= 0) return fail(OC_ERR_INVALID, "bad argument");
    if (obs && misaligned16(obs)) return fail(OC_ERR_INVALID, "obs must be 16-byte aligned");
    if (int dc = check_device(h)) return dc;
    const OcParams& p = h->p;
    int rc = dispatch(p.A, p.NOBJ, row_mode(p), [&](auto a, auto nobj, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        constexpr int RF = decltype(rf)::value;
        const int grid = (p.E + h->threads - 1) / h->threads;
        oc_rollout_kernel<AA, NN, FF, RF><<<grid, h->threads, h->smem_bytes, (cudaStream_t)stream>>>(
            p, h->state, n_steps, h->rollout_step, obs, rew_f32, done, actions_out, actions_in);
        CUDA_TRY(cudaGetLastError());
        return OC_OK;
    });
    if (rc == OC_OK) { h->launches += 1; if (!actions_in) h->rollout_step += (uint32_t)n_steps; }
    return rc;
}

extern "C" int oc_get_state(oc_env* h, uint32_t* state, void* stream) {
    if (!h || !state) return fail(OC_ERR_INVALID, "null argument");
    if (misaligned16(state)) return fail(OC_ERR_INVALID, "state must be 16-byte aligned");
    if (int dc = check_device(h)) return dc;
    const int n = h->p.E * 4;
    oc_state_export_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->state, state, h->p.E);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return OC_OK;
}

extern "C" int oc_set_state(oc_env* h, const uint32_t* state, void* stream) {
    if (!h || !state) return fail(OC_ERR_INVALID, "null argument");
    if (misaligned16(state)) return fail(OC_ERR_INVALID, "state must be 16-byte aligned");
    if (int dc = check_device(h)) return dc;
    const int n = h->p.E * 4;
    oc_state_import_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->state, state, h->p.E);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return OC_OK;
}

extern "C" int oc_get_stats(oc_env* h, uint32_t* episodes, uint32_t* last_completed, void* stream) {
    if (!h) return fail(OC_ERR_INVALID, "null handle");
    if (int dc = check_device(h)) return dc;
    oc_stats_kernel<<<(h->p.E + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->state, episodes, last_completed, h->p.E);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return OC_OK;
}

extern "C" uint64_t oc_launch_count(const oc_env* h) { return h ? h->launches : 0; }

// ---- host-buffer entry points: what a caller without device memory of its own (numpy, SB3 on the CPU) binds

extern "C" int oc_set_device(int device) {
    CUDA_TRY(cudaSetDevice(device));
    return OC_OK;
}

extern "C" int oc_host_alloc(uint64_t bytes, void** out) {
    if (!out) return fail(OC_ERR_INVALID, "null argument");
    *out = nullptr;
    cudaError_t ce = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (ce != cudaSuccess) return fail(OC_ERR_ALLOC, std::string("cudaHostAlloc: ") + cudaGetErrorString(ce));
    return OC_OK;
}

extern "C" int oc_host_free(void* ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return OC_OK;
}

template <typename T>
static int ensure_dev(T*& ptr, size_t count) {
    if (ptr) return OC_OK;
    cudaError_t ce = cudaMalloc(&ptr, count * sizeof(T));
    if (ce != cudaSuccess) { ptr = nullptr; return fail(OC_ERR_ALLOC, std::string("cudaMalloc (host path staging): ") + cudaGetErrorString(ce)); }
    return OC_OK;
}

template <typename T>
static int ensure_host(T*& ptr, size_t count) {
    if (ptr) return OC_OK;
    cudaError_t ce = cudaHostAlloc((void**)&ptr, count * sizeof(T), cudaHostAllocDefault);
    if (ce != cudaSuccess) { ptr = nullptr; return fail(OC_ERR_ALLOC, std::string("cudaHostAlloc (host path staging): ") + cudaGetErrorString(ce)); }
    return OC_OK;
}

// device-side repack of float rows into the compact integer format (also usable on its own)
extern "C" int oc_pack_obs_i8(oc_env* h, const float* obs, int8_t* obs_i8, float* timestep, void* stream) {
    if (!h || !obs || !obs_i8) return fail(OC_ERR_INVALID, "null argument");
    if (reinterpret_cast<uintptr_t>(obs_i8) & 3) return fail(OC_ERR_INVALID, "obs_i8 must be 4-byte aligned");
    if (int dc = check_device(h)) return dc;
    const OcParams& p = h->p;
    if (p.off_ts != p.F - 1) return fail(OC_ERR_INVALID, "timestep is not the last key of the row");
    const uint64_t nbytes = (uint64_t)p.E * p.A * (p.F - 1);
    if ((uint64_t)p.E * p.A * p.F >= (1ull << 32)) return fail(OC_ERR_INVALID, "batch too large for the 32-bit pack index");
    const uint32_t nwords = (uint32_t)std::max<uint64_t>((nbytes + 3) / 4, (uint64_t)p.E);
    oc_pack_i8_kernel<<<(nwords + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, obs, obs_i8, timestep, nwords);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return OC_OK;
}

// i8 = compact format: obs / term_obs are int8 [E, A, F-1] and the clocks go to ts / term_ts (f32 [E])
static int reset_host_impl(oc_env* h, const uint8_t* mask, const int32_t* placements, void* obs, float* ts,
                           bool i8, void* stream) {
    if (!h) return fail(OC_ERR_INVALID, "null handle");
    if (int dc = check_device(h)) return dc;
    const OcParams& p = h->p;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t E = (size_t)p.E, row = (size_t)p.row_bytes, row8 = (size_t)p.A * (p.F - 1);
    int rc;
    if (mask) {
        if ((rc = ensure_dev(h->hp.mask, E))) return rc;
        CUDA_TRY(cudaMemcpyAsync(h->hp.mask, mask, E, cudaMemcpyHostToDevice, st));
    }
    if (placements && p.nrandom > 0) {
        if ((rc = ensure_dev(h->hp.place, E * p.nrandom))) return rc;
        CUDA_TRY(cudaMemcpyAsync(h->hp.place, placements, E * p.nrandom * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    if (obs && (rc = ensure_dev(h->hp.obs, E * row))) return rc;
    if (obs && i8 && ((rc = ensure_dev(h->hp.obs8, E * row8 + 4)) || (rc = ensure_dev(h->hp.ts, E)))) return rc;
    if ((rc = oc_reset(h, mask ? h->hp.mask : nullptr, (placements && p.nrandom > 0) ? h->hp.place : nullptr,
                       obs ? h->hp.obs : nullptr, stream))) return rc;
    if (obs && i8) {
        if ((rc = oc_pack_obs_i8(h, h->hp.obs, h->hp.obs8, h->hp.ts, stream))) return rc;
        CUDA_TRY(cudaMemcpyAsync(obs, h->hp.obs8, E * row8, cudaMemcpyDeviceToHost, st));
        if (ts) CUDA_TRY(cudaMemcpyAsync(ts, h->hp.ts, E * sizeof(float), cudaMemcpyDeviceToHost, st));
    } else if (obs) {
        CUDA_TRY(cudaMemcpyAsync(obs, h->hp.obs, E * row * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return OC_OK;
}

static int step_host_impl(oc_env* h, const int32_t* actions, void* obs, float* ts, float* rew_f32, double* rew_f64,
                          uint8_t* done, void* term_obs, float* term_ts, bool i8, uint32_t flags, void* stream) {
    if (!h || !actions || !obs || !done) return fail(OC_ERR_INVALID, "null argument");
    if (int dc = check_device(h)) return dc;
    const OcParams& p = h->p;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t E = (size_t)p.E, A = (size_t)p.A, row = (size_t)p.row_bytes, row8 = A * (size_t)(p.F - 1);
    int rc;
    if ((rc = ensure_dev(h->hp.actions, E * A * 2)) || (rc = ensure_dev(h->hp.obs, E * row)) ||
        (rc = ensure_dev(h->hp.done, E))) return rc;
    if (i8 && ((rc = ensure_dev(h->hp.obs8, E * row8 + 4)) || (rc = ensure_dev(h->hp.ts, E)))) return rc;
    if (rew_f32 && (rc = ensure_dev(h->hp.rew32, E * A))) return rc;
    if (rew_f64 && (rc = ensure_dev(h->hp.rew64, E))) return rc;
    const bool want_term = term_obs != nullptr && (flags & OC_FLAG_AUTO_RESET);
    if (want_term && !h->hp.term) {                    // rows of envs that never finished stay zero
        if ((rc = ensure_dev(h->hp.term, E * row))) return rc;
        CUDA_TRY(cudaMemsetAsync(h->hp.term, 0, E * row * sizeof(float), st));
    }
    CUDA_TRY(cudaMemcpyAsync(h->hp.actions, actions, E * A * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if ((rc = oc_step(h, h->hp.actions, h->hp.obs, rew_f32 ? h->hp.rew32 : nullptr, rew_f64 ? h->hp.rew64 : nullptr,
                      h->hp.done, want_term ? h->hp.term : nullptr, flags, stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(done, h->hp.done, E, cudaMemcpyDeviceToHost, st));
    if (rew_f32) CUDA_TRY(cudaMemcpyAsync(rew_f32, h->hp.rew32, E * A * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (rew_f64) CUDA_TRY(cudaMemcpyAsync(rew_f64, h->hp.rew64, E * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (i8) {
        if ((rc = oc_pack_obs_i8(h, h->hp.obs, h->hp.obs8, h->hp.ts, stream))) return rc;
        CUDA_TRY(cudaMemcpyAsync(obs, h->hp.obs8, E * row8, cudaMemcpyDeviceToHost, st));
        if (ts) CUDA_TRY(cudaMemcpyAsync(ts, h->hp.ts, E * sizeof(float), cudaMemcpyDeviceToHost, st));
    } else {
        CUDA_TRY(cudaMemcpyAsync(obs, h->hp.obs, E * row * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (want_term) {                                   // only the rows of envs that just finished reach the caller's buffer
        if ((rc = ensure_host(h->hp.h_idx, E))) return rc;
        size_t nfin = 0, e = 0;
        for (; e + 8 <= E; e += 8) {                   // finished envs are rare: skip eight clear flags at a time
            uint64_t w8;
            memcpy(&w8, done + e, 8);
            if (w8 == 0) continue;
            for (size_t k = e; k < e + 8; ++k)
                if (done[k]) h->hp.h_idx[nfin++] = (int32_t)k;
        }
        for (; e < E; ++e)
            if (done[e]) h->hp.h_idx[nfin++] = (int32_t)e;
        if (nfin == 0) return OC_OK;
        // gather them on the device into a dense [nfin, row] buffer (the compact format is packed on the way), one
        // copy to pinned host memory, then row-wise into the caller's (possibly pageable) buffer.  The staging
        // buffers are sized for the worst case: lock-step envs all hit the time limit in the same step.
        if ((rc = ensure_dev(h->hp.idx, E)) || (rc = ensure_dev(h->hp.gather, E * row)) || (rc = ensure_host(h->hp.h_gather, E * row))) return rc;
        if (i8 && ((rc = ensure_dev(h->hp.gather_ts, E)) || (rc = ensure_host(h->hp.h_gather_ts, E)))) return rc;
        const size_t rb = i8 ? row8 : row * sizeof(float);                // bytes of one env's rows
        CUDA_TRY(cudaMemcpyAsync(h->hp.idx, h->hp.h_idx, nfin * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        const int grid = (int)std::min<size_t>(nfin, 148 * 16);
        oc_gather_term_kernel<<<grid, 128, 0, st>>>(p, h->hp.term, h->hp.idx, (int)nfin, i8 ? nullptr : h->hp.gather,
                                                    i8 ? (int8_t*)h->hp.gather : nullptr, h->hp.gather_ts);
        CUDA_TRY(cudaGetLastError());
        h->launches += 1;
        CUDA_TRY(cudaMemcpyAsync(h->hp.h_gather, h->hp.gather, nfin * rb, cudaMemcpyDeviceToHost, st));
        if (i8) CUDA_TRY(cudaMemcpyAsync(h->hp.h_gather_ts, h->hp.gather_ts, nfin * sizeof(float), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const uint8_t* hb = (const uint8_t*)h->hp.h_gather;
        uint8_t* dst = (uint8_t*)term_obs;
        for (size_t i = 0; i < nfin; ++i) {
            const size_t e = (size_t)h->hp.h_idx[i];
            memcpy(dst + e * rb, hb + i * rb, rb);
            if (i8 && term_ts) term_ts[e] = h->hp.h_gather_ts[i];
        }
    }
    return OC_OK;
}

extern "C" int oc_reset_host(oc_env* h, const uint8_t* mask, const int32_t* placements, float* obs, void* stream) {
    return reset_host_impl(h, mask, placements, obs, nullptr, false, stream);
}
extern "C" int oc_step_host(oc_env* h, const int32_t* actions, float* obs, float* rew_f32, double* rew_f64,
                            uint8_t* done, float* term_obs, uint32_t flags, void* stream) {
    return step_host_impl(h, actions, obs, nullptr, rew_f32, rew_f64, done, term_obs, nullptr, false, flags, stream);
}
extern "C" int oc_reset_host_i8(oc_env* h, const uint8_t* mask, const int32_t* placements, int8_t* obs_i8,
                                float* timestep, void* stream) {
    return reset_host_impl(h, mask, placements, obs_i8, timestep, true, stream);
}
extern "C" int oc_step_host_i8(oc_env* h, const int32_t* actions, int8_t* obs_i8, float* timestep, float* rew_f32,
                               double* rew_f64, uint8_t* done, int8_t* term_obs_i8, float* term_timestep,
                               uint32_t flags, void* stream) {
    return step_host_impl(h, actions, obs_i8, timestep, rew_f32, rew_f64, done, term_obs_i8, term_timestep, true, flags, stream);
}
