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

// Launch bounds of the step kernel: the float-row instantiations are limited by shared memory to 2-3 CTAs per SM, so
// they may use up to 128 registers (no spills); the byte-row and compact-row ones want many resident warps (64 registers).
// A/B knob (compile time): -DOC_STEP_MIN_CTAS=n forces one value for all of them.
#ifdef OC_STEP_MIN_CTAS
#define OC_STEP_BOUNDS(MODE) __launch_bounds__(256, OC_STEP_MIN_CTAS)
#else
#define OC_STEP_BOUNDS(MODE) __launch_bounds__(256, ((MODE) == 1 || (MODE) == 2) ? 2 : 4)
#endif

// A/B knob (compile time): -DOC_ROW_UNDO=0 makes the fused kernel clear its single-pass float rows in every step again
#ifndef OC_ROW_UNDO
#define OC_ROW_UNDO 1
#endif

// Phase probe (tools/probe_step.py, -DOC_PHASE_PROBE builds only; never in the shipped library): lane 0 of every
// warp stamps %clock64 at the phase boundaries of the step kernel and %globaltimer at entry / exit.
#ifdef OC_PHASE_PROBE
__device__ unsigned long long* g_probe = nullptr;
__device__ __forceinline__ unsigned long long probe_clock(uint32_t dep) {
    __shared__ uint32_t scratch[32];
    unsigned long long c;
    // the store cannot issue before `dep` has arrived; the clock read issues after it (in-order issue)
    asm volatile("st.shared.u32 [%1], %2;\n\tmov.u64 %0, %%clock64;"
                 : "=l"(c) : "r"((uint32_t)__cvta_generic_to_shared(scratch + (threadIdx.x >> 5))), "r"(dep) : "memory");
    return c;
}
__device__ __forceinline__ unsigned long long probe_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
    return t;
}
#define OC_PROBE(k, dep) do { if (lane == 0) pr[k] = probe_clock(dep); } while (0)
#else
#define OC_PROBE(k, dep) do { } while (0)
#endif

// One oc_step launch.  MODE 0-2: float rows [E, A, F] to `obs`; MODE 3: compact rows int8 [E, A, F-1] to `obs`
// and the clock f32 [E] to `ts`.  Any output pointer may be device memory or page-locked host memory.
struct StepIO {
    const void* actions;      // int32 [E, A, 2], or u8 [E, A, 2] with OC_FLAG_ACTIONS_U8
    void* obs;
    float* ts;                // MODE 3 only
    float* rew32;             // f32 [E, A], or f32 [E] with OC_FLAG_REWARD_PER_ENV
    double* rew64;            // f64 [E]
    uint8_t* done;
    void* term_obs;           // like obs: rows of envs that finished (auto-reset on)
    float* term_ts;           // MODE 3 only
    uint32_t flags;
    int32_t env_lo, env_hi;   // this launch steps envs [env_lo, env_hi); env_lo is a multiple of 32
    // chained steps (OC_FLAG_CHAIN_HEAD / OC_FLAG_CHAINED): one flag per warp-chunk of 32 envs; the launch at chain
    // position k (head = 0) sets flag[c] = k + 1 once chunk c's NEW state is in memory, and a chained launch lets chunk c
    // start as soon as flag[c] == k -- chunk by chunk, instead of waiting for the previous grid to drain and retire
    uint32_t* chain_flags;    // nullptr: this launch is not part of a chain
    uint32_t chain_pos;       // 0: depend on the previous grid as a whole (griddepcontrol.wait)
};

// bounded spin on a device counter (a chain that was put together wrongly must fail loudly, not hang the GPU)
__device__ __forceinline__ void chain_wait_for(const uint32_t* cnt, uint32_t target) {
    for (uint32_t spins = 0;; ++spins) {
        uint32_t v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
        if (v == target) return;
        if (spins > (1u << 22)) __trap();       // ~ a second: the predecessor never stored its state
        __nanosleep(64);
    }
}

// dynamic shared memory: [table blob][per warp: nb env rows (float / biased-byte / compact int8 format)]
template <int A, int NOBJ, int NF, int MODE>
__global__ void OC_STEP_BOUNDS(MODE)
oc_step_kernel(const __grid_constant__ OcParams p, uint4* __restrict__ state, const __grid_constant__ StepIO io) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef OC_PHASE_PROBE
    unsigned long long pr[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned long long g0 = probe_gtime();
#endif
    OC_PROBE(0, 0u);

    // Prologue, all of it in front of the dependency wait, where it is free: with programmatic dependent launch a CTA of
    // this grid enters the moment a CTA of the previous grid leaves its slot, and then has nothing to do until that grid
    // has completed (tools/probe_step.py: median 0.4-1.3 us).  (1) One thread hands the table blob to the copy engine
    // (TMA bulk load, completion on an mbarrier): nobody stalls on it until the dynamics need the tables.  (2) The
    // warps clear their rows.  (3) griddepcontrol.wait: everything below may touch what the previous grid wrote
    // (state, actions).  Both griddepcontrol instructions are no-ops when the launch carries no PDL attribute.
    __shared__ __align__(8) uint64_t tbar;
#ifndef OC_TABLES_GLOBAL      // A/B build (tools/ab_tables_global.sh): tables read from global memory through L1, no staging
    if (threadIdx.x == 0) {
        mbar_init(&tbar, 1);
        tma_load(smem, p.blob, (uint32_t)p.blob_bytes, &tbar);
    }
#endif
    asm volatile("griddepcontrol.launch_dependents;");
    uint8_t* wrows = smem + p.blob_bytes + (size_t)warp * p.warp_row_bytes;
    warp_clear_rows<MODE != 0>(wrows, p.warp_row_bytes, lane);
    __syncthreads();                                // mbarrier initialised by thread 0 -> visible to everyone
    OC_PROBE(1, 0u);
    if (io.chain_pos == 0) asm volatile("griddepcontrol.wait;" ::: "memory");     // chained launches wait per chunk, below
    OC_PROBE(2, 0u);
#ifdef OC_TABLES_GLOBAL
    const Tables tb = make_tables(p, p.blob);
#else
    const Tables tb = make_tables(p, smem);
#endif
    const uint32_t flags = io.flags;

    // the grid is sized to ONE resident wave (148 SMs x CTAs that fit); with more envs than that
    // each CTA walks several chunks so the tables are loaded once per CTA, not once per chunk
    bool first = true;                              // rows still clear from the prologue
    for (int base = io.env_lo + blockIdx.x * blockDim.x; base < io.env_hi; base += gridDim.x * blockDim.x) {
        const int env = base + threadIdx.x;
        const bool valid = env < io.env_hi;
        Env<A, NOBJ> e;
        Info in;
        bool done = false;
        uint4 s0, s1, s2, s3;
        int nav[A], comm[A];
        const int wchunk = (base >> 5) + warp;          // this warp's chunk of 32 envs
        if (io.chain_pos != 0 && base + warp * 32 < io.env_hi) {        // chained: our chunk's state from the previous step
            if (lane == 0) chain_wait_for(io.chain_flags + wchunk, io.chain_pos);
            __syncwarp();
        }
        if (valid) {                                    // raw loads: state planes + this env's actions
            // L2 loads (ld.global.cg): in a chain the previous writer of these lines may be a grid that is still running
            s0 = __ldcg(state + env); s1 = __ldcg(state + p.E + env);
            s2 = __ldcg(state + 2 * p.E + env); s3 = __ldcg(state + 3 * p.E + env);
            if (flags & OC_FLAG_ACTIONS_U8) {
                const uchar2* a2 = reinterpret_cast<const uchar2*>(io.actions) + (size_t)env * A;
#pragma unroll
                for (int k = 0; k < A; ++k) { const uchar2 v = __ldg(a2 + k); nav[k] = v.x & 3; comm[k] = v.y; }
            } else {
                const int2* a2 = reinterpret_cast<const int2*>(io.actions) + (size_t)env * A;
#pragma unroll
                for (int k = 0; k < A; ++k) { const int2 v = __ldg(a2 + k); nav[k] = v.x & 3; comm[k] = v.y; }
            }
        }
        if (first) {
#ifndef OC_TABLES_GLOBAL
            mbar_wait(&tbar, 0);                        // tables have landed (long ago, as a rule)
#endif
            OC_PROBE(3, 0u);
        }
        if (valid) {                                    // dynamics: no row access, may overlap the previous chunk's TMA read
            unpack_env<A, NOBJ>(e, s0, s1, s2, s3);
            OC_PROBE(4, e.w0 + e.comm + e.obj[0] + e.acell[0] + (uint32_t)nav[A - 1]);
            in = step_logic<A, NOBJ, NF>(e, p, tb, nav, comm[0], comm[1], (uint32_t)env, io.rew32, io.rew64, io.done, done,
                                         (flags & OC_FLAG_REWARD_PER_ENV) != 0);
            OC_PROBE(5, in.pres + in.holdmask + (uint32_t)done);
        }
        const bool fin = valid && done && (flags & OC_FLAG_AUTO_RESET);
        if (io.term_obs != nullptr && __any_sync(0xFFFFFFFFu, fin)) {  // rare: some env of this warp finished
            if (MODE == 3)
                warp_terminal_obs_packed<A, NOBJ, NF>(e, in, fin, p, tb, wrows, lane,
                                                      reinterpret_cast<uint8_t*>(io.term_obs) + (size_t)env * p.row_bytes,
                                                      io.term_ts ? io.term_ts + env : nullptr);
            else
                warp_terminal_obs<A, NOBJ, NF, (MODE == 3 ? 0 : MODE)>(e, in, fin, p, tb, wrows, lane,
                                                      reinterpret_cast<float*>(io.term_obs) + (size_t)env * p.row_bytes);
        }
        if (fin) {
            finish_episode<A, NOBJ>(e, p, tb, (uint32_t)env);
            in = gather_info<A, NOBJ, NF>(e, p, tb);
        }
        if (valid) store_env<A, NOBJ>(e, state, p.E, env);
        const int env0 = base + warp * 32;
        const int nvalid = max(0, min(32, io.env_hi - env0));
        OC_PROBE(6, 0u);
        // chained steps: the successor of this chunk (the same chunk of the next launch) may start once the new state is
        // in memory; the flag is released inside emit_obs, between filling the rows and handing them to the copy engine
        uint32_t* cflag = (io.chain_flags != nullptr && nvalid > 0) ? io.chain_flags + wchunk : nullptr;
        if (MODE == 3)
            emit_obs_packed<A, NOBJ, NF>(e, in, valid, p, tb, wrows, lane,
                                         reinterpret_cast<uint8_t*>(io.obs) + (size_t)env0 * p.row_bytes,
                                         io.ts ? io.ts + env0 : nullptr, nvalid, first, cflag, io.chain_pos + 1u);
        else
            emit_obs<A, NOBJ, NF, (MODE == 3 ? 0 : MODE), true>(e, in, valid, p, tb, wrows, lane,
                                         reinterpret_cast<float*>(io.obs) + (size_t)env0 * p.row_bytes, nvalid, first,
                                         cflag, io.chain_pos + 1u);
        OC_PROBE(7, 0u);
        first = false;
    }
    rows_wait_done(p);
    OC_PROBE(8, 0u);
#ifdef OC_PHASE_PROBE
    if (lane == 0 && g_probe != nullptr) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long* o = g_probe + ((size_t)blockIdx.x * 8 + warp) * 16;
        for (int k = 0; k < 9; ++k) o[k] = pr[k];
        o[9] = g0; o[10] = probe_gtime(); o[11] = smid;
    }
#endif
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

    // Single-pass float rows (MODE 1): after the first step of the launch the warp's rows are never cleared again -- they
    // still hold the same envs' previous observation, and emit_obs_undo takes back / overwrites it (shown_*: what the row
    // shows at its data-dependent places).  The first step is peeled off the loop so that the loop body carries ONE copy of
    // the fill code: the hot loop is far larger than the L0 instruction cache, and its size shows up directly in the step time.
    constexpr bool UNDO = (MODE == 1) && (OC_ROW_UNDO != 0);
    uint32_t shown_comm = 0, shown_completed = 0;
    int s = 0;
    if (UNDO && n_steps > 0) {
        Info in;
        if (valid) in = rollout_logic<A, NOBJ, NF, ROWF>(e, p, tb, (uint32_t)env, 0u, step0, rew32, done_out, actions_out, actions_in);
        if (obs != nullptr) {
            emit_obs<A, NOBJ, NF, MODE>(e, in, valid, p, tb, wrows, lane, obs + (size_t)env0 * p.row_bytes, nvalid);
            shown_comm = e.comm; shown_completed = e.completed;
        }
        s = 1;
    }
    for (; s < n_steps; ++s) {
        // dynamics first: they do not touch the rows, so the copy engine may still be reading the
        // previous step's rows out of shared memory while this runs
        Info in;
        if (valid) in = rollout_logic<A, NOBJ, NF, ROWF>(e, p, tb, (uint32_t)env, (uint32_t)s, step0, rew32, done_out, actions_out, actions_in);
        if (obs != nullptr) {
            float* out = obs + (size_t)s * step_floats + (size_t)env0 * p.row_bytes;
            if (UNDO) {
                emit_obs_undo<A, NOBJ, NF>(e, in, valid, p, tb, wrows, lane, out, nvalid, shown_comm, shown_completed);
                shown_comm = e.comm; shown_completed = e.completed;
            } else {
                emit_obs<A, NOBJ, NF, MODE>(e, in, valid, p, tb, wrows, lane, out, nvalid);
            }
        }
    }
    if (valid) store_env<A, NOBJ>(e, state, p.E, env);
    rows_wait_done(p);
}

// reset (masked) + observation of every env (MODE 3: compact rows to `obs`, clocks to `ts`)
template <int A, int NOBJ, int NF, int MODE>
__global__ void __launch_bounds__(256)
oc_reset_kernel(const __grid_constant__ OcParams p, uint4* __restrict__ state, const uint8_t* __restrict__ mask,
                const int32_t* __restrict__ placements, void* __restrict__ obs, float* __restrict__ ts, int initial) {
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
    const int nvalid = max(0, min(32, p.E - env0));
    if (obs != nullptr) {
        if (MODE == 3)
            emit_obs_packed<A, NOBJ, NF>(e, in, valid, p, tb, wrows, lane,
                                         reinterpret_cast<uint8_t*>(obs) + (size_t)env0 * p.row_bytes, ts ? ts + env0 : nullptr, nvalid);
        else
            emit_obs<A, NOBJ, NF, (MODE == 3 ? 0 : MODE)>(e, in, valid, p, tb, wrows, lane,
                                                          reinterpret_cast<float*>(obs) + (size_t)env0 * p.row_bytes, nvalid);
    }
    rows_wait_done(p);
}

// packed state <-> [E, 16] u32 rows
__global__ void oc_state_export_kernel(const uint4* __restrict__ planes, uint32_t* __restrict__ rows, int E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * 4) return;
    const int env = i >> 2, pl = i & 3;
    reinterpret_cast<uint4*>(rows)[(size_t)env * 4 + pl] = planes[(size_t)pl * E + env];
}
// imported state is sanitised on the way in (a checkpoint of another level must not drive the table look-ups out
// of the CTA's shared memory): agent cells and live object cells are clamped to the grid, holders to {agents, none},
// empty slots become the canonical dead word
__global__ void oc_state_import_kernel(const __grid_constant__ OcParams p, uint4* __restrict__ planes,
                                       const uint32_t* __restrict__ rows, int E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * 4) return;
    const int env = i >> 2, pl = i & 3;
    const uint4 v = sanitize_state_plane(p, pl, reinterpret_cast<const uint4*>(rows)[(size_t)env * 4 + pl]);
    planes[(size_t)pl * E + env] = v;
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

// compact terminal rows of the n finished envs idx[0..n) -> dense [n, A*(F-1)] bytes + [n] clocks (staged host path
// with pageable caller buffers; page-locked ones are written by the step kernel itself)
__global__ void __launch_bounds__(128)
oc_gather_term8_kernel(const uint8_t* __restrict__ term8, const float* __restrict__ term_ts, const int32_t* __restrict__ idx,
                       int n, int row8, uint8_t* __restrict__ out8, float* __restrict__ out_ts) {
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint8_t* src = term8 + (size_t)idx[i] * row8;
        uint8_t* dst = out8 + (size_t)i * row8;
        for (int j = threadIdx.x; j < row8; j += blockDim.x) dst[j] = src[j];
        if (threadIdx.x == 0) out_ts[i] = term_ts[idx[i]];
    }
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

// one kernel configuration of a handle: parameter block + CTA shape
struct LaunchCfg {
    OcParams p;
    int threads = 64;
    size_t smem_bytes = 0;
    int step_grid = 1;
    bool ok = false;
};

struct oc_env {
    OcParams p;               // float rows (kernel MODE 0-2); == f.p
    int device = 0;
    int threads = 64;
    size_t smem_bytes = 0;
    LaunchCfg c;              // compact int8 rows (kernel MODE 3); c.ok false when the rows are too wide for one pass
    uint4* state = nullptr;
    uint8_t* blob = nullptr;
    float* ts = nullptr;
    uint64_t launches = 0;
    uint32_t rollout_step = 0;
    int pdl = 1;
    int tma_rows_in_step = 1;
    int step_grid = 1;
    int chain_threads = 0, chain_grid = 0;   // CTA shape of chained float-row launches (0: same as plain launches)
    size_t chain_smem = 0;
    int zero_copy = 1;        // OC_HOST_ZEROCOPY: page-locked caller buffers are read / written by the kernels directly where that pays
    // chained steps: one flag per warp-chunk of 32 envs + what the next OC_FLAG_CHAINED call must match
    uint32_t* chain_cnt = nullptr;
    struct Chain {
        bool active = false, compact = false;
        uint32_t pos = 0;                     // launches of the chain so far
        cudaStream_t stream = nullptr;
        int lo = 0, hi = 0;
        const void *obs = nullptr, *rew32 = nullptr, *rew64 = nullptr, *done = nullptr, *term = nullptr;
    } chain;
    int obs_off[OC_NUM_OBS_KEYS], obs_size[OC_NUM_OBS_KEYS];
    // device-side staging of the host-buffer entry points (oc_step_host* / oc_reset_host*), allocated on first use
    struct HostPath {
        int32_t* actions = nullptr; float* obs = nullptr; float* rew32 = nullptr; double* rew64 = nullptr;
        uint8_t* done = nullptr; float* term = nullptr; uint8_t* mask = nullptr; int32_t* place = nullptr;
        int8_t* obs8 = nullptr; float* ts = nullptr;          // compact format (oc_*_host_i8)
        uint8_t* term8 = nullptr; float* term_ts = nullptr;   // compact terminal rows [E, A*(F-1)] + [E]
        uint8_t* block = nullptr; uint8_t* actions8 = nullptr; // oc_*_host_block: one staging block, u8 actions
        uint32_t* state_rows = nullptr;                        // oc_get_state_host / oc_set_state_host
        // terminal rows: indices of the finished envs (host -> device), their rows gathered densely (device -> host)
        int32_t* idx = nullptr; float* gather = nullptr; float* gather_ts = nullptr;
        int32_t* h_idx = nullptr; float* h_gather = nullptr; float* h_gather_ts = nullptr;   // pinned
    } hp;
};

// kernel template MODE: 0 byte rows, 1 float rows (all 32 envs of a warp in one pass), 2 float rows in passes,
// 3 compact int8 rows (step / reset kernels only)
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
// the compact-row instantiations (MODE 3) exist for the step and reset kernels only
template <typename F>
static int dispatch_shape(int A, int NOBJ, F&& f) {
#define OC_CASE(a, n, nfood)                                                                                            \
    if (A == a && NOBJ == n) return f(std::integral_constant<int, a>(), std::integer_sequence<int, n, nfood>());
    OC_CASE(2, 2, 1) OC_CASE(3, 2, 1) OC_CASE(4, 2, 1) OC_CASE(2, 4, 2) OC_CASE(3, 4, 2) OC_CASE(4, 4, 2)
    OC_CASE(2, 6, 3) OC_CASE(3, 6, 3) OC_CASE(4, 6, 3)
#undef OC_CASE
    return fail(OC_ERR_INVALID, "unsupported (num_agents, num_objects)");
}

extern "C" int oc_abi_version(void) { return OC_ABI_VERSION; }
#ifdef OC_PHASE_PROBE
extern "C" int oc_debug_set_probe(unsigned long long* dev_buf) {      // [grid * 8 warps * 16] u64, or NULL
    CUDA_TRY(cudaMemcpyToSymbol(g_probe, &dev_buf, sizeof(dev_buf)));
    return OC_OK;
}
#endif
extern "C" const char* oc_last_error(void) { return g_err.c_str(); }

// CTA shape: the work is one warp per 32 envs; pick the CTA size whose resident wave covers the envs with the
// smallest makespan, preferring fewer table copies on ties.  caps[t / 32] = resident CTAs of t threads per SM.
static bool pick_cta(const OcParams& p, const int* caps, int num_sm, double want_warps, const char* tenv,
                     int& best_t, int& best_cap) {
    best_t = 0; best_cap = 1;
    double best_cost = 1e30;
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
    return best_t != 0;
}

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
    if (const char* ze = getenv("OC_HOST_ZEROCOPY")) h->zero_copy = atoi(ze) != 0;
    cudaDeviceProp prop;
    {
        cudaError_t cep = cudaGetDeviceProperties(&prop, dev);
        if (cep != cudaSuccess) { delete h; return fail(OC_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(cep)); }
    }
    const int num_sm = prop.multiProcessorCount;
    const size_t smem_cta_max = prop.sharedMemPerBlockOptin;

    // compact-row configuration (MODE 3): the same parameter block with byte rows of A * (F-1), contiguous, one pass
    OcParams& p8 = h->c.p;
    p8 = p;
    make_compact_params(p8);
    const bool want8 = p8.warp_row_bytes <= 64 * 1024 && p.off_ts == p.F - 1;

    // opt in to large dynamic shared memory for the instantiations we launch.  The attribute is per
    // function, not per handle, so it is always set to the device maximum: a second handle with a
    // smaller footprint must not lower it under the first one's feet.  Then ask the runtime how many
    // CTAs of each candidate size are resident per SM (shared memory AND registers).
    const int smem_optin = (int)smem_cta_max - 1024;      // the kernels keep a few static words (mbarrier)
    auto smem_for = [&](const OcParams& q, int threads) { return (size_t)q.blob_bytes + (size_t)(threads / 32) * (size_t)q.warp_row_bytes; };
    int caps[9] = {0}, caps8[9] = {0};
    int rc = dispatch(p.A, p.NOBJ, row_mode(p), [&](auto a, auto nobj, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        constexpr int RF = decltype(rf)::value;
        CUDA_TRY(cudaFuncSetAttribute(oc_step_kernel<AA, NN, FF, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(oc_rollout_kernel<AA, NN, FF, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(oc_reset_kernel<AA, NN, FF, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
        for (int t = 32; t <= 256; t += 32) {
            if (smem_for(p, t) > (size_t)smem_optin) continue;
            int cs = 0, cr = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cs, oc_step_kernel<AA, NN, FF, RF>, t, smem_for(p, t)));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cr, oc_rollout_kernel<AA, NN, FF, RF>, t, smem_for(p, t)));
            caps[t / 32] = std::min(cs, cr);
        }
        return OC_OK;
    });
    if (rc == OC_OK && want8) rc = dispatch_shape(p.A, p.NOBJ, [&](auto a, auto nobj) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        CUDA_TRY(cudaFuncSetAttribute(oc_step_kernel<AA, NN, FF, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(oc_reset_kernel<AA, NN, FF, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
        for (int t = 32; t <= 256; t += 32) {
            if (smem_for(p8, t) > (size_t)smem_optin) continue;
            int cs = 0, cr = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cs, oc_step_kernel<AA, NN, FF, 3>, t, smem_for(p8, t)));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cr, oc_reset_kernel<AA, NN, FF, 3>, t, smem_for(p8, t)));
            caps8[t / 32] = std::min(cs, cr);
        }
        return OC_OK;
    });
    if (rc != OC_OK) { std::string m = g_err; delete h; return fail(rc, m); }

    const char* tenv = getenv("OC_BLOCK_THREADS");
    // resident warps per SM that hide the latencies: byte rows run LDS -> convert -> STG chains, float rows
    // leave through the copy engine, multi-pass float rows wait on it most of the time
    const double want_warps = !p.rowf ? 24.0 : (p.obs_passes > 1 ? 6.0 : 14.0);
    int best_t = 0, best_cap = 1;
    if (!pick_cta(p, caps, num_sm, want_warps, tenv, best_t, best_cap)) { delete h; return fail(OC_ERR_INVALID, "observation row too wide for shared memory"); }
    h->threads = best_t;
    h->smem_bytes = smem_for(p, best_t);
    h->step_grid = (int)std::min<long long>(((long long)p.E + best_t - 1) / best_t, (long long)num_sm * best_cap);
    // Chained launches overlap with their predecessor CTA by CTA: when the batch is a single wave of contiguous
    // single-pass float rows, four-warp CTAs (four or more resident per SM) hand their slots over in finer steps than
    // the seven-warp CTAs plain launches like best (cfg2: 6.7 us vs 7.6 us per chained step; tools/r2_chain_sweep.sh).
    {
        const char* ce = getenv("OC_BLOCK_THREADS_CHAIN");
        const int ct = ce ? atoi(ce) : 128;
        if (ct >= 32 && ct <= 256 && ct % 32 == 0 && caps[ct / 32] >= 1 && (ce || ((p.use_tma == 1 || p.use_tma == 3) && p.obs_passes == 1 && !tenv &&
            caps[ct / 32] >= 3 && ((long long)p.E + ct - 1) / ct <= (long long)num_sm * caps[ct / 32]))) {
            h->chain_threads = ct;
            h->chain_smem = smem_for(p, ct);
            h->chain_grid = (int)std::min<long long>(((long long)p.E + ct - 1) / ct, (long long)num_sm * caps[ct / 32]);
            // OC_CHAIN_CHUNKS = m: a chained grid of 1/m of the CTAs, each walking m chunk groups.  The table load and the
            // final drain are then paid once per m chunks, a warp's dynamics of chunk i+1 overlap the drain of chunk i (as
            // in the fused kernel), and the slots the smaller grid leaves free are taken by the NEXT step of the chain.
            const char* me = getenv("OC_CHAIN_CHUNKS");
            const int m = me ? std::min(16, std::max(1, atoi(me))) : 1;
            h->chain_grid = std::max(1, (h->chain_grid + m - 1) / m);
        }
    }
    // compact rows: small CTAs when the batch is a fraction of a wave (<= 16 warps of work per SM) -- the launch is then
    // bound by the slowest CTA, and two-warp CTAs hand their slots over soonest (cfg2: 6.1 us vs 8.3 us with 256 threads,
    // tools/step_sweep.py); larger batches amortise the table copy over more warps
    const char* tenv8 = getenv("OC_BLOCK_THREADS_I8");
    if (!tenv8 && (long long)p.E <= (long long)num_sm * 16 * 32 && caps8[2] >= 1) tenv8 = "64";
    if (want8 && pick_cta(p8, caps8, num_sm, 32.0, tenv8, best_t, best_cap)) {
        h->c.threads = best_t;
        h->c.smem_bytes = smem_for(p8, best_t);
        h->c.step_grid = (int)std::min<long long>(((long long)p.E + best_t - 1) / best_t, (long long)num_sm * best_cap);
        h->c.ok = true;
    }

    cudaError_t ce;
    if ((ce = cudaMalloc(&h->state, (size_t)p.E * 64)) != cudaSuccess ||
        (ce = cudaMalloc(&h->chain_cnt, (size_t)((p.E + 31) / 32) * sizeof(uint32_t))) != cudaSuccess ||   // here, not on first use:
        (ce = cudaMemset(h->chain_cnt, 0, (size_t)((p.E + 31) / 32) * sizeof(uint32_t))) != cudaSuccess ||  // that may be inside a capture
        (ce = cudaMalloc(&h->blob, blob.size())) != cudaSuccess ||
        (ce = cudaMalloc(&h->ts, ts.size() * 4)) != cudaSuccess ||
        (ce = cudaMemcpy(h->blob, blob.data(), blob.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (ce = cudaMemcpy(h->ts, ts.data(), ts.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
        std::string m = std::string("device allocation/copy failed: ") + cudaGetErrorString(ce);
        oc_destroy(h);
        return fail(ce == cudaErrorMemoryAllocation ? OC_ERR_ALLOC : OC_ERR_CUDA, m);
    }
    p.blob = h->blob; p.ts_table = h->ts;
    p8.blob = h->blob; p8.ts_table = h->ts;

    rc = dispatch(p.A, p.NOBJ, row_mode(p), [&](auto a, auto nobj, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        constexpr int RF = decltype(rf)::value;
        const int grid = (p.E + h->threads - 1) / h->threads;
        oc_reset_kernel<AA, NN, FF, RF><<<grid, h->threads, h->smem_bytes, 0>>>(p, h->state, nullptr, nullptr, nullptr, nullptr, 1);
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
    if (h->chain_cnt) cudaFree(h->chain_cnt);
    if (h->blob) cudaFree(h->blob);
    if (h->ts) cudaFree(h->ts);
    for (void* q : {(void*)h->hp.actions, (void*)h->hp.obs, (void*)h->hp.rew32, (void*)h->hp.rew64, (void*)h->hp.done,
                    (void*)h->hp.term, (void*)h->hp.mask, (void*)h->hp.place, (void*)h->hp.obs8, (void*)h->hp.ts,
                    (void*)h->hp.term8, (void*)h->hp.term_ts, (void*)h->hp.block, (void*)h->hp.actions8,
                    (void*)h->hp.state_rows, (void*)h->hp.idx, (void*)h->hp.gather, (void*)h->hp.gather_ts})
        if (q) cudaFree(q);
    for (void* q : {(void*)h->hp.h_idx, (void*)h->hp.h_gather, (void*)h->hp.h_gather_ts})
        if (q) cudaFreeHost(q);
    delete h;
    return OC_OK;
}

extern "C" int oc_obs_width(const oc_env* h) { return h ? h->p.F : OC_ERR_INVALID; }

extern "C" int oc_compact_supported(const oc_env* h) { return (h && h->c.ok) ? 1 : 0; }

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

// ---- launches ------------------------------------------------------------------------------------------------
// compact = kernel MODE 3 (obs int8 [E, A, F-1] + ts f32 [E]); otherwise float rows
static int launch_reset(oc_env* h, bool compact, const uint8_t* mask, const int32_t* placements, void* obs, float* ts,
                        cudaStream_t st) {
    h->chain.active = false;                          // anything but a step ends a chain of steps
    const OcParams& p = compact ? h->c.p : h->p;
    const int threads = compact ? h->c.threads : h->threads;
    const size_t smem = compact ? h->c.smem_bytes : h->smem_bytes;
    const int grid = (p.E + threads - 1) / threads;
    int rc;
    if (compact) rc = dispatch_shape(p.A, p.NOBJ, [&](auto a, auto nobj) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        oc_reset_kernel<AA, NN, FF, 3><<<grid, threads, smem, st>>>(p, h->state, mask, placements, obs, ts, 0);
        CUDA_TRY(cudaGetLastError());
        return OC_OK;
    });
    else rc = dispatch(p.A, p.NOBJ, row_mode(p), [&](auto a, auto nobj, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        constexpr int RF = decltype(rf)::value;
        oc_reset_kernel<AA, NN, FF, RF><<<grid, threads, smem, st>>>(p, h->state, mask, placements, obs, ts, 0);
        CUDA_TRY(cudaGetLastError());
        return OC_OK;
    });
    if (rc == OC_OK) h->launches += 1;
    return rc;
}

static int launch_step(oc_env* h, bool compact, StepIO io, cudaStream_t st) {
    const OcParams& p = compact ? h->c.p : h->p;
    if (io.env_hi <= 0) { io.env_lo = 0; io.env_hi = p.E; }
    // ---- chained steps.  HEAD: zero the chain counter (stream-ordered, before the kernel), run with the ordinary
    // grid-wide dependency, count this launch's state chunks.  CHAINED: must directly follow a HEAD / CHAINED step of
    // the same kind on the same stream with DIFFERENT output buffers (the two grids overlap: nothing orders their
    // observation / reward / done stores); it starts as soon as every state chunk of the chain so far is in memory.
    const uint32_t chain_flags = io.flags & (OC_FLAG_CHAIN_HEAD | OC_FLAG_CHAINED);
    io.chain_flags = nullptr; io.chain_pos = 0;
    if (chain_flags) {
        if (chain_flags == (OC_FLAG_CHAIN_HEAD | OC_FLAG_CHAINED)) return fail(OC_ERR_INVALID, "OC_FLAG_CHAIN_HEAD and OC_FLAG_CHAINED exclude each other");
        oc_env::Chain& c = h->chain;
        if (chain_flags & OC_FLAG_CHAIN_HEAD) {
            // stale flags of an earlier chain (or an earlier replay of this graph) must not look like this chain's
            CUDA_TRY(cudaMemsetAsync(h->chain_cnt, 0, (size_t)((p.E + 31) / 32) * sizeof(uint32_t), st));
            c.active = true; c.compact = compact; c.pos = 0; c.stream = st; c.lo = io.env_lo; c.hi = io.env_hi;
        } else {
            if (!c.active || c.compact != compact || c.stream != st || c.lo != io.env_lo || c.hi != io.env_hi)
                return fail(OC_ERR_INVALID, "OC_FLAG_CHAINED: the previous launch of this handle must be a chain head or a chained "
                                            "step of the same kind (float / compact rows), on the same stream, over the same envs");
            if (io.obs == c.obs || (io.rew32 && io.rew32 == c.rew32) || (io.rew64 && io.rew64 == c.rew64) || io.done == c.done ||
                (io.term_obs && io.term_obs == c.term))
                return fail(OC_ERR_INVALID, "OC_FLAG_CHAINED: consecutive steps of a chain overlap in time and must write different "
                                            "obs / reward / done / terminal buffers (e.g. consecutive rollout-buffer slots)");
            if (c.pos >= (1u << 30)) return fail(OC_ERR_INVALID, "chain too long: start a new one with OC_FLAG_CHAIN_HEAD");
            io.chain_pos = c.pos;
        }
        io.chain_flags = h->chain_cnt;
        c.pos += 1;
        c.obs = io.obs; c.rew32 = io.rew32; c.rew64 = io.rew64; c.done = io.done; c.term = io.term_obs;
    } else {
        h->chain.active = false;                      // a plain step ends the chain
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    const bool chain_shape = !compact && io.chain_flags != nullptr && h->chain_threads != 0;
    cfg.gridDim = dim3(compact ? h->c.step_grid : (chain_shape ? h->chain_grid : h->step_grid));
    cfg.blockDim = dim3(compact ? h->c.threads : (chain_shape ? h->chain_threads : h->threads));
    cfg.dynamicSmemBytes = compact ? h->c.smem_bytes : (chain_shape ? h->chain_smem : h->smem_bytes);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = h->pdl ? 1 : 0;
    OcParams ps = p;
    if (!compact && ps.rowf && !ps.use_tma && h->tma_rows_in_step) ps.use_tma = 2;      // padded float rows: per-row bulk copies
    int rc;
    if (compact) rc = dispatch_shape(p.A, p.NOBJ, [&](auto a, auto nobj) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        CUDA_TRY(cudaLaunchKernelEx(&cfg, oc_step_kernel<AA, NN, FF, 3>, ps, h->state, io));
        return OC_OK;
    });
    else rc = dispatch(p.A, p.NOBJ, row_mode(p), [&](auto a, auto nobj, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(nobj), FF = shape_nf(nobj);
        constexpr int RF = decltype(rf)::value;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, oc_step_kernel<AA, NN, FF, RF>, ps, h->state, io));
        return OC_OK;
    });
    if (rc == OC_OK) h->launches += 1;
    return rc;
}

extern "C" int oc_reset(oc_env* h, const uint8_t* mask, const int32_t* placements, float* obs, void* stream) {
    if (!h) return fail(OC_ERR_INVALID, "null handle");
    if (obs && misaligned16(obs)) return fail(OC_ERR_INVALID, "obs must be 16-byte aligned");
    if (int dc = check_device(h)) return dc;
    return launch_reset(h, false, mask, placements, obs, nullptr, (cudaStream_t)stream);
}

extern "C" int oc_reset_i8(oc_env* h, const uint8_t* mask, const int32_t* placements, int8_t* obs_i8, float* timestep,
                           void* stream) {
    if (!h) return fail(OC_ERR_INVALID, "null handle");
    if (!h->c.ok) return fail(OC_ERR_INVALID, "observation rows too wide for the compact format kernels; use oc_reset + oc_pack_obs_i8");
    if (obs_i8 && misaligned16(obs_i8)) return fail(OC_ERR_INVALID, "obs_i8 must be 16-byte aligned");
    if (int dc = check_device(h)) return dc;
    return launch_reset(h, true, mask, placements, obs_i8, timestep, (cudaStream_t)stream);
}

extern "C" int oc_step(oc_env* h, const int32_t* actions, float* obs, float* rew_f32, double* rew_f64,
                       uint8_t* done, float* term_obs, uint32_t flags, void* stream) {
    if (!h || !actions || !obs || !done) return fail(OC_ERR_INVALID, "null argument");
    if (misaligned16(obs) || (reinterpret_cast<uintptr_t>(actions) & 7)) return fail(OC_ERR_INVALID, "obs must be 16-byte and actions 8-byte aligned");
    if (flags & ~(OC_FLAG_AUTO_RESET | OC_FLAG_CHAIN_HEAD | OC_FLAG_CHAINED))
        return fail(OC_ERR_INVALID, "oc_step takes OC_FLAG_AUTO_RESET | OC_FLAG_CHAIN_HEAD | OC_FLAG_CHAINED");
    if (int dc = check_device(h)) return dc;
    StepIO io{actions, obs, nullptr, rew_f32, rew_f64, done, term_obs, nullptr, flags, 0, h->p.E};
    return launch_step(h, false, io, (cudaStream_t)stream);
}

extern "C" int oc_step_i8(oc_env* h, const void* actions, int8_t* obs_i8, float* timestep, float* rew_f32, double* rew_f64,
                          uint8_t* done, int8_t* term_obs_i8, float* term_timestep, uint32_t flags, void* stream) {
    if (!h || !actions || !obs_i8 || !done) return fail(OC_ERR_INVALID, "null argument");
    if (!h->c.ok) return fail(OC_ERR_INVALID, "observation rows too wide for the compact format kernels; use oc_step + oc_pack_obs_i8");
    if (misaligned16(obs_i8)) return fail(OC_ERR_INVALID, "obs_i8 must be 16-byte aligned");
    if (!(flags & OC_FLAG_ACTIONS_U8) && (reinterpret_cast<uintptr_t>(actions) & 7)) return fail(OC_ERR_INVALID, "int32 actions must be 8-byte aligned");
    if ((flags & OC_FLAG_ACTIONS_U8) && (reinterpret_cast<uintptr_t>(actions) & 1)) return fail(OC_ERR_INVALID, "u8 actions must be 2-byte aligned");
    if (flags & ~(OC_FLAG_AUTO_RESET | OC_FLAG_ACTIONS_U8 | OC_FLAG_REWARD_PER_ENV | OC_FLAG_CHAIN_HEAD | OC_FLAG_CHAINED))
        return fail(OC_ERR_INVALID, "unknown flag");
    if (int dc = check_device(h)) return dc;
    StepIO io{actions, obs_i8, timestep, rew_f32, rew_f64, done, term_obs_i8, term_timestep, flags, 0, h->p.E};
    return launch_step(h, true, io, (cudaStream_t)stream);
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
    h->chain.active = false;
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
    h->chain.active = false;
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
    h->chain.active = false;
    const int n = h->p.E * 4;
    oc_state_import_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->p, h->state, state, h->p.E);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return OC_OK;
}

extern "C" int oc_get_stats(oc_env* h, uint32_t* episodes, uint32_t* last_completed, void* stream) {
    if (!h) return fail(OC_ERR_INVALID, "null handle");
    if (int dc = check_device(h)) return dc;
    h->chain.active = false;
    oc_stats_kernel<<<(h->p.E + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->state, episodes, last_completed, h->p.E);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return OC_OK;
}

extern "C" uint64_t oc_launch_count(const oc_env* h) { return h ? h->launches : 0; }

extern "C" int oc_sync(oc_env* h, void* stream) {
    if (!h) return fail(OC_ERR_INVALID, "null handle");
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return OC_OK;
}

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

// Page-locked host memory (oc_host_alloc / cudaHostAlloc / cudaHostRegister) is addressable from the device under
// unified virtual addressing: returns the device alias of `ptr`, or nullptr for pageable memory.
template <typename T>
static T* device_alias(T* ptr) {
    if (!ptr) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, (const void*)ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return nullptr;
    return (T*)at.devicePointer;
}

extern "C" int oc_get_state_host(oc_env* h, uint32_t* state, void* stream) {
    if (!h || !state) return fail(OC_ERR_INVALID, "null argument");
    int rc;
    const size_t n = (size_t)h->p.E * OC_STATE_WORDS;
    if ((rc = ensure_dev(h->hp.state_rows, n))) return rc;
    if ((rc = oc_get_state(h, h->hp.state_rows, stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(state, h->hp.state_rows, n * 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return OC_OK;
}

extern "C" int oc_set_state_host(oc_env* h, const uint32_t* state, void* stream) {
    if (!h || !state) return fail(OC_ERR_INVALID, "null argument");
    int rc;
    const size_t n = (size_t)h->p.E * OC_STATE_WORDS;
    if ((rc = ensure_dev(h->hp.state_rows, n))) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->hp.state_rows, state, n * 4, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if ((rc = oc_set_state(h, h->hp.state_rows, stream))) return rc;
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
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
    const uint8_t* dmask = mask ? h->hp.mask : nullptr;
    const int32_t* dplace = (placements && p.nrandom > 0) ? h->hp.place : nullptr;
    if (obs && i8) {
        if ((rc = ensure_dev(h->hp.obs8, E * row8 + 16)) || (rc = ensure_dev(h->hp.ts, E))) return rc;
        if (h->c.ok) {                                  // the reset kernel emits the compact rows itself
            if ((rc = launch_reset(h, true, dmask, dplace, h->hp.obs8, h->hp.ts, st))) return rc;
        } else {
            if ((rc = ensure_dev(h->hp.obs, E * row))) return rc;
            if ((rc = launch_reset(h, false, dmask, dplace, h->hp.obs, nullptr, st))) return rc;
            if ((rc = oc_pack_obs_i8(h, h->hp.obs, h->hp.obs8, h->hp.ts, stream))) return rc;
        }
        CUDA_TRY(cudaMemcpyAsync(obs, h->hp.obs8, E * row8, cudaMemcpyDeviceToHost, st));
        if (ts) CUDA_TRY(cudaMemcpyAsync(ts, h->hp.ts, E * sizeof(float), cudaMemcpyDeviceToHost, st));
    } else {
        if (obs && (rc = ensure_dev(h->hp.obs, E * row))) return rc;
        if ((rc = launch_reset(h, false, dmask, dplace, obs ? h->hp.obs : nullptr, nullptr, st))) return rc;
        if (obs) CUDA_TRY(cudaMemcpyAsync(obs, h->hp.obs, E * row * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return OC_OK;
}

// indices of the set flags of a host done buffer -> h_idx; finished envs are rare, eight clear flags are skipped at a time
static size_t scan_done(const uint8_t* done, size_t E, int32_t* h_idx) {
    size_t nfin = 0, e = 0;
    for (; e + 8 <= E; e += 8) {
        uint64_t w8;
        memcpy(&w8, done + e, 8);
        if (w8 == 0) continue;
        for (size_t k = e; k < e + 8; ++k)
            if (done[k]) h_idx[nfin++] = (int32_t)k;
    }
    for (; e < E; ++e)
        if (done[e]) h_idx[nfin++] = (int32_t)e;
    return nfin;
}

static int step_host_impl(oc_env* h, const int32_t* actions, void* obs, float* ts, float* rew_f32, double* rew_f64,
                          uint8_t* done, void* term_obs, float* term_ts, bool i8, uint32_t flags, void* stream) {
    if (!h || !actions || !obs || !done) return fail(OC_ERR_INVALID, "null argument");
    if (flags & ~OC_FLAG_AUTO_RESET) return fail(OC_ERR_INVALID, "the host entry points take OC_FLAG_AUTO_RESET only");
    if (int dc = check_device(h)) return dc;
    const OcParams& p = h->p;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t E = (size_t)p.E, A = (size_t)p.A, row = (size_t)p.row_bytes, row8 = A * (size_t)(p.F - 1);
    const bool direct8 = i8 && h->c.ok;                // the step kernel emits the compact rows itself (MODE 3)
    int rc;
    if ((rc = ensure_dev(h->hp.actions, E * A * 2)) || (rc = ensure_dev(h->hp.done, E))) return rc;
    if (!direct8 && (rc = ensure_dev(h->hp.obs, E * row))) return rc;
    if (i8 && ((rc = ensure_dev(h->hp.obs8, E * row8 + 16)) || (rc = ensure_dev(h->hp.ts, E)))) return rc;
    if (rew_f32 && (rc = ensure_dev(h->hp.rew32, E * A))) return rc;
    if (rew_f64 && (rc = ensure_dev(h->hp.rew64, E))) return rc;
    const bool want_term = term_obs != nullptr && (flags & OC_FLAG_AUTO_RESET);
    // compact terminal rows into page-locked caller buffers: the kernel writes the few finished rows across PCIe itself
    int8_t* term8_alias = (want_term && direct8 && h->zero_copy) ? device_alias((int8_t*)term_obs) : nullptr;
    float* term_ts_alias = term8_alias ? device_alias(term_ts) : nullptr;
    const bool term_direct = term8_alias != nullptr && (term_ts == nullptr || term_ts_alias != nullptr);
    if (want_term && !term_direct) {                   // device-side terminal buffer; rows of envs that never finished stay zero
        if (direct8) {
            if (!h->hp.term8) {
                if ((rc = ensure_dev(h->hp.term8, E * row8)) || (rc = ensure_dev(h->hp.term_ts, E))) return rc;
                CUDA_TRY(cudaMemsetAsync(h->hp.term8, 0, E * row8, st));
                CUDA_TRY(cudaMemsetAsync(h->hp.term_ts, 0, E * sizeof(float), st));
            }
        } else if (!h->hp.term) {
            if ((rc = ensure_dev(h->hp.term, E * row))) return rc;
            CUDA_TRY(cudaMemsetAsync(h->hp.term, 0, E * row * sizeof(float), st));
        }
    }
    CUDA_TRY(cudaMemcpyAsync(h->hp.actions, actions, E * A * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (direct8) {
        StepIO io{h->hp.actions, h->hp.obs8, h->hp.ts, rew_f32 ? h->hp.rew32 : nullptr, rew_f64 ? h->hp.rew64 : nullptr,
                  h->hp.done, nullptr, nullptr, flags, 0, p.E};
        if (want_term) { io.term_obs = term_direct ? (void*)term8_alias : (void*)h->hp.term8; io.term_ts = term_direct ? term_ts_alias : h->hp.term_ts; }
        if ((rc = launch_step(h, true, io, st))) return rc;
    } else {
        StepIO io{h->hp.actions, h->hp.obs, nullptr, rew_f32 ? h->hp.rew32 : nullptr, rew_f64 ? h->hp.rew64 : nullptr,
                  h->hp.done, want_term ? h->hp.term : nullptr, nullptr, flags, 0, p.E};
        if ((rc = launch_step(h, false, io, st))) return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(done, h->hp.done, E, cudaMemcpyDeviceToHost, st));
    if (rew_f32) CUDA_TRY(cudaMemcpyAsync(rew_f32, h->hp.rew32, E * A * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (rew_f64) CUDA_TRY(cudaMemcpyAsync(rew_f64, h->hp.rew64, E * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (i8) {
        if (!direct8 && (rc = oc_pack_obs_i8(h, h->hp.obs, h->hp.obs8, h->hp.ts, stream))) return rc;
        CUDA_TRY(cudaMemcpyAsync(obs, h->hp.obs8, E * row8, cudaMemcpyDeviceToHost, st));
        if (ts) CUDA_TRY(cudaMemcpyAsync(ts, h->hp.ts, E * sizeof(float), cudaMemcpyDeviceToHost, st));
    } else {
        CUDA_TRY(cudaMemcpyAsync(obs, h->hp.obs, E * row * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (want_term && !term_direct) {                   // only the rows of envs that just finished reach the caller's buffer
        if ((rc = ensure_host(h->hp.h_idx, E))) return rc;
        const size_t nfin = scan_done(done, E, h->hp.h_idx);
        if (nfin == 0) return OC_OK;
        // gather them on the device into a dense [nfin, row] buffer (the compact format is packed on the way), one
        // copy to pinned host memory, then row-wise into the caller's (possibly pageable) buffer.  The staging
        // buffers are sized for the worst case: lock-step envs all hit the time limit in the same step.
        if ((rc = ensure_dev(h->hp.idx, E)) || (rc = ensure_dev(h->hp.gather, E * row)) || (rc = ensure_host(h->hp.h_gather, E * row))) return rc;
        if (i8 && ((rc = ensure_dev(h->hp.gather_ts, E)) || (rc = ensure_host(h->hp.h_gather_ts, E)))) return rc;
        const size_t rb = i8 ? row8 : row * sizeof(float);                // bytes of one env's rows
        CUDA_TRY(cudaMemcpyAsync(h->hp.idx, h->hp.h_idx, nfin * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        const int grid = (int)std::min<size_t>(nfin, 148 * 16);
        if (direct8)
            oc_gather_term8_kernel<<<grid, 128, 0, st>>>(h->hp.term8, h->hp.term_ts, h->hp.idx, (int)nfin, (int)row8,
                                                         (uint8_t*)h->hp.gather, h->hp.gather_ts);
        else
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

// ---- one-block host path: a single page-locked block per step ------------------------------------------------
static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

extern "C" int oc_host_block_layout(const oc_env* h, oc_host_block* out) {
    if (!h || !out) return fail(OC_ERR_INVALID, "null argument");
    const size_t E = (size_t)h->p.E, row8 = (size_t)h->p.A * (h->p.F - 1);
    size_t off = 0;
    out->obs_i8 = off;   off += align256(E * row8);
    out->timestep = off; off += align256(E * 4);
    out->reward = off;   off += align256(E * 4);
    out->done = off;     off += align256(E);
    out->total_bytes = off;
    return OC_OK;
}

extern "C" int oc_reset_host_block(oc_env* h, const uint8_t* mask, const int32_t* placements, void* block, void* stream) {
    if (!h || !block) return fail(OC_ERR_INVALID, "null argument");
    oc_host_block L;
    oc_host_block_layout(h, &L);
    uint8_t* b = (uint8_t*)block;
    return reset_host_impl(h, mask, placements, b + L.obs_i8, (float*)(b + L.timestep), true, stream);
}

extern "C" int oc_step_host_block(oc_env* h, const uint8_t* actions_u8, void* block, int8_t* term_obs_i8,
                                  float* term_timestep, uint32_t flags, void* stream) {
    if (!h || !actions_u8 || !block) return fail(OC_ERR_INVALID, "null argument");
    if (!h->c.ok) return fail(OC_ERR_INVALID, "observation rows too wide for the compact format kernels; use oc_step_host_i8");
    if (flags & ~(OC_FLAG_AUTO_RESET | OC_FLAG_NO_SYNC)) return fail(OC_ERR_INVALID, "oc_step_host_block takes OC_FLAG_AUTO_RESET | OC_FLAG_NO_SYNC");
    if (int dc = check_device(h)) return dc;
    const OcParams& p = h->p;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t E = (size_t)p.E, A = (size_t)p.A;
    oc_host_block L;
    oc_host_block_layout(h, &L);
    int rc;
    if ((rc = ensure_dev(h->hp.block, (size_t)L.total_bytes))) return rc;
    const bool want_term = term_obs_i8 != nullptr && (flags & OC_FLAG_AUTO_RESET);
    int8_t* term8_alias = want_term ? device_alias(term_obs_i8) : nullptr;
    float* term_ts_alias = want_term ? device_alias(term_timestep) : nullptr;
    if (want_term && (!term8_alias || (term_timestep && !term_ts_alias)))
        return fail(OC_ERR_INVALID, "oc_step_host_block: terminal buffers must be page-locked (oc_host_alloc); pageable callers use oc_step_host_i8");
    // actions: page-locked u8 pairs are read by the kernel across PCIe (0.26 MB at cfg2: no copy call, no dependent
    // launch); pageable ones are staged with one copy
    const uint8_t* act = h->zero_copy ? device_alias(actions_u8) : nullptr;
    if (!act) {
        if ((rc = ensure_dev(h->hp.actions8, E * A * 2))) return rc;
        CUDA_TRY(cudaMemcpyAsync(h->hp.actions8, actions_u8, E * A * 2, cudaMemcpyHostToDevice, st));
        act = h->hp.actions8;
    }
    uint8_t* d = h->hp.block;
    StepIO io{act, d + L.obs_i8, (float*)(d + L.timestep), (float*)(d + L.reward), nullptr, d + L.done,
              term8_alias, term_ts_alias,
              (flags & OC_FLAG_AUTO_RESET) | OC_FLAG_ACTIONS_U8 | OC_FLAG_REWARD_PER_ENV, 0, p.E};
    if ((rc = launch_step(h, true, io, st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(block, d, (size_t)L.total_bytes, cudaMemcpyDeviceToHost, st));      // ONE copy for everything
    if (!(flags & OC_FLAG_NO_SYNC)) CUDA_TRY(cudaStreamSynchronize(st));
    return OC_OK;
}
