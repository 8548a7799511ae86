// Device-side Overcooked dynamics for sm_100a.
//
// Mapping: one THREAD owns one env for the branchy integer logic (a warp instruction advances 32
// envs), and the WARP cooperatively streams the 32 envs' observation rows from shared memory to
// global memory with coalesced 16-byte stores.  The kernels are bound by the integer ALU pipe
// (ncu: sm__pipe_alu_cycles_active ~50 % = its issue limit), so the code below
//   * tests object slots with one mask-compare on the packed word (dead slot = OCK_DEAD, which
//     can never match a holder or a cell),
//   * walks the objects ONCE per step to gather everything reward, shaping and observation need,
//   * does the observation arithmetic in float on the FMA pipe (FADD.SAT / FFMA) and keeps
//     float rows in shared memory when they fit, so the expansion is a plain 16-byte copy.
// Reference semantics and their file:line anchors are listed next to each block.
#pragma once
#ifdef OCK_HOST_EMU
#include "oc_emu_shim.h"   // tests/emu only: runs this header on the CPU to debug the logic without a GPU
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include "oc_params.h"

namespace ock {

enum : uint32_t { TILE_FLOOR = 0, TILE_COUNTER = 1, TILE_CUTBOARD = 2, TILE_DELIVERY = 3 };

// ---------------------------------------------------------------------------------------------
// shared-memory view of the per-level tables (one copy per CTA)
struct Tables {
    const double*   q;        // q[n] = n / MAX_PATH as the reference's Python float (f64)
    const uint32_t* tmlut;    // [128] object signature -> bitmask of subtasks whose goal template it equals
    const float2*   xyf;      // [ncell] (x, y) of a cell as floats (only ever indexed with live cells)
    const float*    ts;       // [T+1] float32(t / T) when small enough for shared memory (else nullptr)
    const uint16_t* mvt;      // [ncell*4] inbounds(cell + NAV[a]) | tile(target) << 8     world.py:317-320
    const uint16_t* xy16;     // [ncell] x | y << 8
    const uint8_t*  dmin;     // [ncell] min over Delivery tiles of pd + manhattan   overcooked_environment.py:383-388
    const uint8_t*  counters; // [ncounters] Counter cells, reading order            overcooked_environment.py:164
    const uint8_t*  pd;       // [ncell*ncell] World.get_path_distance_between       world.py:114-131
    const uint8_t*  pdm;      // [ncell*ncell] pd + manhattan distance               overcooked_environment.py:380
};

__device__ __forceinline__ Tables make_tables(const OcParams& p, const uint8_t* smem) {
    Tables t;
    t.q = reinterpret_cast<const double*>(smem + p.o_q);
    t.tmlut = reinterpret_cast<const uint32_t*>(smem + p.o_tmlut);
    t.xyf = reinterpret_cast<const float2*>(smem + p.o_xyf);
    t.mvt = reinterpret_cast<const uint16_t*>(smem + p.o_mvt);
    t.xy16 = reinterpret_cast<const uint16_t*>(smem + p.o_xy16);
    t.dmin = smem + p.o_dmin;
    t.counters = smem + p.o_counters;
    t.pd = smem + p.o_pd;
    t.pdm = smem + p.o_pdm;
    t.ts = (p.o_ts >= 0) ? reinterpret_cast<const float*>(smem + p.o_ts) : nullptr;
    return t;
}

// cooperative copy of the table blob into shared memory (16-byte chunks)
__device__ __forceinline__ void load_tables(const OcParams& p, uint8_t* smem) {
    const uint4* src = reinterpret_cast<const uint4*>(p.blob);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < (p.blob_bytes >> 4); i += blockDim.x) dst[i] = __ldg(src + i);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: same function in oracle/oc_oracle.c
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ---------------------------------------------------------------------------------------------
template <int A, int NOBJ>
struct Env {
    uint32_t w0, episodes, completed, countbits;
    uint32_t acell[A];
    uint32_t w5;
    uint64_t ranks;
    uint32_t obj[NOBJ];
    uint32_t comm;      // agent0 | agent1 << 16
    uint32_t w15;
};

template <int A, int NOBJ>
__device__ __forceinline__ void unpack_env(Env<A, NOBJ>& e, const uint4 a, const uint4 b, const uint4 c, const uint4 d);

template <int A, int NOBJ>
__device__ __forceinline__ void load_env(Env<A, NOBJ>& e, const uint4* __restrict__ st, int E, int i) {
    unpack_env<A, NOBJ>(e, st[i], st[E + i], st[2 * E + i], st[3 * E + i]);
}

// the four 16-byte state planes of one env -> fields
template <int A, int NOBJ>
__device__ __forceinline__ void unpack_env(Env<A, NOBJ>& e, const uint4 a, const uint4 b, const uint4 c, const uint4 d) {
    e.w0 = a.x; e.episodes = a.y; e.completed = a.z; e.countbits = a.w;
#pragma unroll
    for (int k = 0; k < A; ++k) e.acell[k] = (b.x >> (8 * k)) & 0xFF;
    e.w5 = b.y;
    e.ranks = (uint64_t)b.z | ((uint64_t)b.w << 32);
    const uint32_t o[6] = {c.x, c.y, c.z, c.w, d.x, d.y};
#pragma unroll
    for (int s = 0; s < NOBJ; ++s) e.obj[s] = o[s];
    e.comm = d.z; e.w15 = d.w;
}

template <int A, int NOBJ>
__device__ __forceinline__ void store_env(const Env<A, NOBJ>& e, uint4* __restrict__ st, int E, int i) {
    uint32_t cells = 0;
#pragma unroll
    for (int k = 0; k < A; ++k) cells |= e.acell[k] << (8 * k);
    uint32_t o[6] = {OCK_DEAD, OCK_DEAD, OCK_DEAD, OCK_DEAD, OCK_DEAD, OCK_DEAD};
#pragma unroll
    for (int s = 0; s < NOBJ; ++s) o[s] = e.obj[s];
    st[i] = make_uint4(e.w0, e.episodes, e.completed, e.countbits);
    st[E + i] = make_uint4(cells, e.w5, (uint32_t)e.ranks, (uint32_t)(e.ranks >> 32));
    st[2 * E + i] = make_uint4(o[0], o[1], o[2], o[3]);
    st[3 * E + i] = make_uint4(o[4], o[5], e.comm, e.w15);
}

// object word: contents[0:4] | chopped[4:7] | holder[8:11] (7 = not held) | cell[16:24] | stamp[24:32]
__device__ __forceinline__ uint32_t obj_contents(uint32_t o) { return o & 0xFu; }
__device__ __forceinline__ uint32_t obj_chopped(uint32_t o) { return (o >> 4) & 7u; }
__device__ __forceinline__ uint32_t obj_holder(uint32_t o) { return (o >> 8) & 7u; }
__device__ __forceinline__ uint32_t obj_cell(uint32_t o) { return (o >> 16) & 0xFFu; }
__device__ __forceinline__ bool obj_alive(uint32_t o) { return (o & 0xFu) != 0; }
__device__ __forceinline__ bool obj_held(uint32_t o) { return (o & 0x700u) != 0x700u; }
__device__ __forceinline__ uint32_t obj_set_cell(uint32_t o, uint32_t c) { return (o & ~0x00FF0000u) | (c << 16); }

// ---------------------------------------------------------------------------------------------
// load_level phase 4 (overcooked_environment.py:157-173): each random object goes to a Counter
// drawn uniformly from ALL Counter tiles, rejecting tiles already taken by an earlier phase-4
// object == sequential sampling without replacement: draw j picks the k-th still-free counter,
// k = mulhi(Philox word j of counter (env, episode, 'RESE', j / 4), ncounters - j) -- the same
// algorithm as oracle/oc_oracle.c.  "The k-th still-free counter" = k, bumped once for every taken index
// it reaches, the taken indices visited in ascending order; at most five are ever taken, so they live
// sorted in registers and the whole draw is a few dozen straight-line compare / add / min / max
// instructions: no loop, no local array, no call.  (A finishing env holds up its whole warp, and in a
// launch of a few steps the slowest warp IS the launch: a stack-based version cost cfg4 1.3 us per step
// of a 20-step rollout, a 128-bit taken-mask version spent a data-dependent loop per object on "drop the k lowest free bits".)
__device__ __forceinline__ void draw_random_cells(const OcParams& p, const uint8_t* __restrict__ counters,
                                                  uint32_t env_id, uint32_t episode, uint32_t (&cell)[OCK_MAX_OBJECTS]) {
    uint32_t r[8];
    philox4x32_10(env_id, episode, 0x52455345u, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r);
    if (p.nrandom > 4)       // every shipped random level places 3 objects: one block of four words is enough
        philox4x32_10(env_id, episode, 0x52455345u, 1u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r + 4);
    else
        r[4] = r[5] = r[6] = r[7] = 0u;
    uint32_t taken[OCK_MAX_OBJECTS];                         // ascending; fully unrolled -> registers
#pragma unroll
    for (int j = 0; j < OCK_MAX_OBJECTS; ++j) {
        if (j < p.nrandom) {
            uint32_t idx = __umulhi(r[j], (uint32_t)(p.ncounters - j));
#pragma unroll
            for (int i = 0; i < j; ++i) idx += (idx >= taken[i]) ? 1u : 0u;
            cell[j] = counters[idx];
            uint32_t v = idx;                                // insert into the sorted list
#pragma unroll
            for (int i = 0; i < j; ++i) { const uint32_t lo = min(taken[i], v); v = max(taken[i], v); taken[i] = lo; }
            taken[j] = v;
        }
    }
}

// OvercookedEnvironment.reset (overcooked_environment.py:180-206).  placements: this env's R
// cells, or nullptr -> draw on device.  Comm buffers are kept (overcooked_env.py:284-297 never
// touches per_agent_communications).
template <int A, int NOBJ>
__device__ __forceinline__ void env_reset(Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                          const int32_t* __restrict__ placements, uint32_t env_id) {
    e.w0 = p.init_w0;
    e.completed = 0;
    e.countbits = 0;
    e.ranks = p.init_ranks;
#pragma unroll
    for (int k = 0; k < A; ++k) e.acell[k] = p.start_cell[k];
#pragma unroll
    for (int s = 0; s < NOBJ; ++s) e.obj[s] = p.init_obj[s];
    if (p.nrandom > 0) {
        uint32_t cell[OCK_MAX_OBJECTS] = {0, 0, 0, 0, 0, 0};
        if (placements != nullptr) {
            for (int j = 0; j < p.nrandom; ++j)          // clamped: a bad index must not run the table look-ups out of bounds
                cell[j] = min((uint32_t)placements[j], (uint32_t)(p.ncell - 1));
        } else {
            draw_random_cells(p, tb.counters, env_id, e.episodes, cell);
        }
#pragma unroll
        for (int j = 0; j < OCK_MAX_OBJECTS; ++j) {
            if (j < p.nrandom) {
                const int slot = p.random_slot[j];
#pragma unroll
                for (int s = 0; s < NOBJ; ++s)
                    if (s == slot) e.obj[s] = obj_set_cell(e.obj[s], cell[j]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One pass over the object slots gathering what reward, shaping and observation need.
// Domain rule: every Food exists at most once, so "the object containing food f" is unique;
// only the Plate channel can have several candidates, and the LAST one in world.objects
// iteration order wins the observation (last writer, overcooked_env.py:121-131): order = (key
// creation rank of the object's name, insertion stamp) -- SURVEY A.8-2.
struct Info {
    uint32_t fword[3];      // object word holding Tomato / Lettuce / Onion (OCK_DEAD if none)
    uint32_t pword;         // plate-channel winner (OCK_DEAD if none)
    uint32_t holdmask;      // bit k: agent k carries something
    uint32_t pres, deliv;   // subtasks whose goal template exists somewhere / on the first Delivery tile
};

template <int A, int NOBJ, int NF>
__device__ __forceinline__ Info gather_info(const Env<A, NOBJ>& e, const OcParams& p, const Tables& tb) {
    Info in;
    in.fword[0] = in.fword[1] = in.fword[2] = OCK_DEAD;
    in.pword = OCK_DEAD;
    in.holdmask = 0; in.pres = 0; in.deliv = 0;
    uint32_t pkey = 0;
#pragma unroll
    for (int s = 0; s < NOBJ; ++s) {
        const uint32_t o = e.obj[s];
        const uint32_t tm = tb.tmlut[o & 0x7Fu];             // dead slot: signature 0 -> no subtask
        in.pres |= tm;
        if ((o & 0x00FF0000u) == ((uint32_t)p.delivery0 << 16)) in.deliv |= tm;   // first Delivery tile only (:259,:402)
#pragma unroll
        for (int f = 0; f < NF; ++f)
            if ((o >> f) & 1u) in.fword[f] = o;
        if (o & 8u) {
            const uint32_t key = ((uint32_t)((e.ranks >> (4 * (o & 0xFu))) & 15ull) << 8) | (o >> 24);
            if (key > pkey) { pkey = key; in.pword = o; }
        }
        in.holdmask |= 1u << obj_holder(o);
    }
    in.holdmask &= 0xFu;                                     // bit 7 = "nobody"
    return in;
}

// ---------------------------------------------------------------------------------------------
// One env step.  nav[k] in [0,4); returns the returned (shaped) reward and done.
//   comm write + CAN_MOVE          gym_comm/envs/overcooked_env.py:227-262
//   t += 1                          overcooked_environment.py:213
//   check_collisions/is_collision   :543-613
//   interact per agent in order     gym_cooking/utils/interact.py:4-75
//   done                            :243-270
//   reward/subtask_reward           :399-432
//   calculate_reward_shaping x2     :272-397
template <int A, int NOBJ, int NF>
__device__ __forceinline__ Info env_step(Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                         const int (&nav)[A], int comm0, int comm1,
                                         double& reward, bool& done) {
    // ---- comm channel write (overcooked_env.py:227-246)
    {
        const uint32_t c0 = p.comm_on ? (uint32_t)comm0 : OCK_COMM_NONE;
        const uint32_t c1 = (p.comm_on && !p.ego_led) ? (uint32_t)comm1 : OCK_COMM_NONE;
        e.comm = (c0 & 0xFFFFu) | (c1 << 16);
    }
    if ((e.w0 & 0xFFFFu) != 0xFFFFu) e.w0 += 1;   // t += 1 (t lives in the low 16 bits and saturates instead of carrying into the stamp)
    const uint32_t t = e.w0 & 0xFFFFu;

    // ---- collisions, on the ORIGINAL actions for every pair (:543-613)
    uint32_t tgt[A], ttile[A], nxt[A];
    bool act[A], ex[A];
#pragma unroll
    for (int k = 0; k < A; ++k) {
        act[k] = p.can_move[k] != 0;                        // (0,0) iff CAN_MOVE false (:250-262)
        const uint32_t m = tb.mvt[e.acell[k] * 4 + nav[k]]; // inbounds(loc + action) | tile << 8
        tgt[k] = m & 0xFFu;
        ttile[k] = m >> 8;
        // off-grid targets assert in the reference (world.py:314); here the clamped target is the
        // agent's own (floor) cell, i.e. the agent stays.
        nxt[k] = (act[k] && ttile[k] == TILE_FLOOR) ? tgt[k] : e.acell[k];
        ex[k] = true;
    }
#pragma unroll
    for (int i = 0; i < A; ++i) {
#pragma unroll
        for (int j = i + 1; j < A; ++j) {
            if (nxt[i] == nxt[j]) {
                if (nxt[i] == e.acell[i] && act[i]) ex[j] = false;
                else if (nxt[j] == e.acell[j] && act[j]) ex[i] = false;
                else { ex[i] = false; ex[j] = false; }
            } else if (e.acell[i] == nxt[j] && e.acell[j] == nxt[i]) {
                ex[i] = false; ex[j] = false;
            }
        }
    }

    // ---- interact, sequentially in agent order on the already-mutated world (interact.py:4-75)
    uint32_t next_stamp = (e.w0 >> 16) & 0xFFu, nkeys = e.w0 >> 24;
#pragma unroll
    for (int k = 0; k < A; ++k) {
        if (!(act[k] && ex[k])) continue;
        const uint32_t tg = tgt[k], tt = ttile[k];
        uint32_t hv = OCK_DEAD, hm = 0;    // held object word / slot mask
#pragma unroll
        for (int s = 0; s < NOBJ; ++s)
            if ((e.obj[s] & 0x700u) == ((uint32_t)k << 8)) { hv = e.obj[s]; hm = 1u << s; }
        if (tt == TILE_FLOOR) {            // move; held object moves along (agent.py:311-314)
            e.acell[k] = tg;
#pragma unroll
            for (int s = 0; s < NOBJ; ++s)
                if ((hm >> s) & 1u) e.obj[s] = obj_set_cell(e.obj[s], tg);
            continue;
        }
        uint32_t ov = OCK_DEAD, om = 0;    // un-held object on the target tile (world.py:217-222)
        const uint32_t want = (tg << 16) | 0x700u;
#pragma unroll
        for (int s = 0; s < NOBJ; ++s)
            if ((e.obj[s] & 0x00FF0700u) == want) { ov = e.obj[s]; om = 1u << s; }
        uint32_t newh = hv, newo = ov;
        if (hm != 0) {
            const uint32_t hc = obj_contents(hv), hch = obj_chopped(hv);
            const bool h_done = (hc & 7u) == hch;                    // every Food in its last state
            const uint32_t put = obj_set_cell(hv, tg) | 0x700u;      // on the tile, nobody holds it
            if (tt == TILE_DELIVERY) {                               // :25-30, is_deliverable core.py:232-237
                if (__popc(hc) > 1 && h_done) newh = put;
            } else if (om != 0) {                                    // merge :33-42, mergeable core.py:240-257
                const uint32_t oc = obj_contents(ov), och = obj_chopped(ov);
                if (!(hc & oc & 8u) && h_done && (oc & 7u) == och) {
                    next_stamp += 1;                                 // world.insert under the NEW name
                    newh = ((hv | (ov & 0x7Fu)) & 0x00FFFFFFu) | (next_stamp << 24);
                    const uint32_t nm = newh & 0xFu;
                    if (((e.ranks >> (4 * nm)) & 15ull) == 0ull) {   // key created on first insert (world.py:236-237)
                        nkeys += 1;
                        e.ranks |= (uint64_t)nkeys << (4 * nm);
                    }
                    newo = OCK_DEAD;                                 // absorbed object leaves the world
                }
            } else {                                                 // :48-59
                if (tt == TILE_CUTBOARD && (hc == 1u || hc == 2u || hc == 4u) && hch == 0u)
                    newh = hv | (hc << 4);                           // chop in hand (core.py:201-206)
                else
                    newh = put;                                      // put down
            }
        } else if (om != 0 && tt != TILE_DELIVERY && !p.allergic[k]) {   // pick up :64-71, agent.py:296-305
            newo = (ov & ~0x00FF0700u) | (e.acell[k] << 16) | ((uint32_t)k << 8);
        }
#pragma unroll
        for (int s = 0; s < NOBJ; ++s) {
            if ((hm >> s) & 1u) e.obj[s] = newh;
            if ((om >> s) & 1u) e.obj[s] = newo;
        }
    }
    e.w0 = t | (next_stamp << 16) | (nkeys << 24);

    // ---- done + sparse reward through the signature -> subtask-mask table
    const Info in = gather_info<A, NOBJ, NF>(e, p, tb);
    const uint32_t dl = in.deliv & p.deliver_mask;
    done = (t >= (uint32_t)p.T) || (dl == p.deliver_mask);                   // :243-270 (T >= 1: oc_create rejects 0)
    const uint32_t nw = in.pres & ~e.countbits & p.nondeliver_mask;          // count rose (:409-415)
    const int sparse = 3 * __popc(dl) + __popc(nw);
    e.countbits = in.pres & p.nondeliver_mask;
    e.completed |= dl | nw;                                                  // :425-426

    // ---- reward shaping (:272-397); f64 additions in reference order
    // (1) Chop subtasks still open: U = [pd(agent, FreshX)], needs min(U) and len(U)
    int lenU = 0, nf[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) nf[f] = 0;
#pragma unroll
    for (int f = 0; f < NF; ++f) {                                          // NF = food channels this level can use
        const bool fresh = (in.fword[f] & 0x7Fu) == (1u << f);              // X alone and un-chopped
        nf[f] = fresh ? __popc(~e.completed & p.chop_mask[f]) : 0;
        lenU += nf[f];
    }
    // (2) item-pair distances (:319-363): agent independent.  Items = Plate + the Foods of
    // recipes[0]; only min(P) and len(P) are used.  pd(src, .) == MAX_PATH unless src is a floor
    // cell, i.e. unless the FIRST item's object is being carried (world.py:126-127).
    int lenP = p.npairs, minP = p.M;
    if (in.holdmask != 0) {
        lenP = 0;
        int mP[3] = {p.M, p.M, p.M};
#pragma unroll
        for (int s = 0; s < NOBJ; ++s) {
            const uint32_t o = e.obj[s];
            if ((o & 8u) && obj_held(o)) {                                   // a carried plate-bearing object
                const uint8_t* row = tb.pd + obj_cell(o) * p.ncell;
#pragma unroll
                for (int f = 0; f < NF; ++f)
                    if (((p.item_foods >> f) & 1u) && obj_alive(in.fword[f]))
                        mP[f] = min(mP[f], (int)row[obj_cell(in.fword[f])]);
            }
        }
#pragma unroll
        for (int f = 0; f < NF; ++f)
            if (((p.item_foods >> f) & 1u) && mP[f] != 0) { lenP += 1; minP = min(minP, mP[f]); }
        // Food-Food pairs, first item alphabetically first: (Lettuce, Onion) (Lettuce, Tomato) (Onion, Tomato)
        const int px[3] = {1, 1, 2}, py[3] = {2, 0, 0};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (px[q] >= NF || py[q] >= NF) continue;                            // compile-time after unrolling
            if (((p.item_foods >> px[q]) & 1u) && ((p.item_foods >> py[q]) & 1u)) {
                const uint32_t ox = in.fword[px[q]], oy = in.fword[py[q]];
                int m = p.M;
                if (obj_alive(ox) && obj_held(ox) && obj_alive(oy)) m = tb.pd[obj_cell(ox) * p.ncell + obj_cell(oy)];
                if (m != 0) { lenP += 1; minP = min(minP, m); }
            }
        }
    }
    double tp[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        tp[a] = 0.0;
        if (lenU > 0) {
            const uint8_t* row = tb.pd + e.acell[a] * p.ncell;
            int minU = 1 << 20;
#pragma unroll
            for (int f = 0; f < NF; ++f)
                if (nf[f] > 0) minU = min(minU, (int)row[obj_cell(in.fword[f])]);
            tp[a] = tb.q[minU + p.M + (lenU - 1) * 2 * p.M];                         // :303-304
            if (lenP > 0) tp[a] = __dadd_rn(tp[a], (double)lenP);                    // :363
        } else if (lenP > 0) {
            tp[a] = tb.q[minP + (lenP - 1) * p.M];                                   // :361
        }
    }
    // (3) Deliver subtasks still open, table order (:370-395)
    for (int j = 0; j < p.ndeliver; ++j) {
        if ((e.completed >> p.deliver_idx[j]) & 1u) continue;
        const uint32_t sig = p.deliver_sig[j];
        uint32_t dw = OCK_DEAD;
#pragma unroll
        for (int s = 0; s < NOBJ; ++s)
            if ((e.obj[s] & 0x7Fu) == sig) dw = e.obj[s];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (!obj_alive(dw)) { tp[a] = __dadd_rn(tp[a], 2.0); continue; }             // :377-378
            const int d = tb.pdm[e.acell[a] * p.ncell + obj_cell(dw)];                    // pd + manhattan (:380)
            if (d == 0) tp[a] = __dadd_rn(tp[a], tb.q[tb.dmin[e.acell[a]]]);              // :381-389
            else tp[a] = __dadd_rn(tp[a], __dadd_rn(tb.q[d], 1.0));                       // :393
        }
    }
    reward = __dsub_rn(__dsub_rn((double)sparse, tp[0]), tp[1]);                     // overcooked_env.py:282
    return in;
}

// =============================================================================================
// Observation rows.  get_observation2 (gym_comm/envs/overcooked_env.py:105-159) for every
// observer of one env, written into this env's shared-memory row by its owning thread, then
// streamed to global memory by the whole warp.  Two row formats:
//   ROWF = true : float32 rows -> leave shared memory as they are (one bulk copy, or 16-byte stores);
//                 rows too wide for 32 of them per warp go out in passes of 16 / 8 / 4 envs (emit_obs)
//   ROWF = false: one biased byte per feature (value + 128) -> PRMT + FADD per float on the way out
//                 (rows whose length is not a multiple of 4 floats, or wider than 1920 floats)
// =============================================================================================
#define OCK_BIAS 128u

// per observer: is_hidden / object_encodings_x / _y / state_encodings of channel c, as floats
__device__ __forceinline__ void channel_features(const Tables& tb, uint32_t word, float2 me, float fow, int c,
                                                 float& hid, float& ex, float& ey, float& st) {
    const float2 o = tb.xyf[obj_cell(word)];
    const float dx = o.x - me.x, dy = o.y - me.y;
    st = (c < 3) ? (float)((word >> (4 + c)) & 1u) : 0.0f;         // Food state_index; Plate has none (:127-128)
    hid = __saturatef(fabsf(dx) + fabsf(dy) - fow);                // 0 if |dx|+|dy| <= radius else 1 (:133)
    ex = __fmaf_rn(dx, hid, 0.0f);                                 // near objects report (0,0) (:135); +0.0, never -0.0
    ey = __fmaf_rn(dy, hid, 0.0f);
}

// `rot`: the observer this thread writes FIRST.  Lanes 8 apart share their banks when the row stride is 4 (mod 8) words
// (cfg2: 92, cfg3: 156), so the emit path hands lanes of different octets different starting observers: at any one
// store instruction they then write rows that are F words apart (F is not a multiple of 4 for those strides),
// i.e. on different banks -- the 4-way conflict of the fixed-offset stores becomes 2-way (A = 2, 3) or none (A = 4).
//
// UNDO (the fused kernel, single-pass rows): the row is NOT zero-filled but still holds this env's previous
// observation.  Every feature at a fixed offset is then written unconditionally (zeros included), and the three
// data-dependent ones are taken back first: the two message one-hots of the previous step (`shown_comm`) and the
// completed-subtask bits that are no longer set (`shown_completed`, i.e. after an episode boundary).  That replaces
// the 12 KB row clear of every step (92 shared-memory wavefronts per warp) by a handful of scattered stores.
template <int A, int NOBJ, int NF, bool ROT = false /* compile-time: without it k stays a constant of the unrolled loop */,
          bool UNDO = false>
__device__ __forceinline__ void build_rows_f32(const Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                               const Info& in, float ts, float* __restrict__ row /* zero-filled */,
                                               int rot = 0, uint32_t shown_comm = 0, uint32_t shown_completed = 0) {
    const float2 a0 = tb.xyf[e.acell[0]], a1 = tb.xyf[e.acell[1]];
    const uint32_t c0 = e.comm & 0xFFFFu, c1 = e.comm >> 16;
    const float fow = (float)p.fow;
    const uint32_t blind_bits = (p.blind[0] ? 1u : 0u) | (p.blind[1] ? 2u : 0u) | (p.blind[2] ? 4u : 0u) | (p.blind[3] ? 8u : 0u);
#pragma unroll
    for (int k0 = 0; k0 < A; ++k0) {
        int k = k0;
        if (ROT) { k += rot; k -= (k >= A) ? A : 0; }
        float* r = row + k * p.F;
        const bool blind = ((blind_bits >> k) & 1u) != 0;                      // :115-118
        if (UNDO) {                                                            // the previous step's messages
            const uint32_t s0 = shown_comm & 0xFFFFu, s1 = shown_comm >> 16;
            if (s0 != OCK_COMM_NONE) r[p.off_a1comm + s0] = 0.0f;
            if (s1 != OCK_COMM_NONE) r[p.off_a2comm + s1] = 0.0f;
        }
        if (c0 != OCK_COMM_NONE) r[p.off_a1comm + c0] = 1.0f;
        if (c1 != OCK_COMM_NONE) r[p.off_a2comm + c1] = 1.0f;
        if (!blind) {                                                          // :139-143
            r[p.off_a1loc] = a0.x; r[p.off_a1loc + 1] = a0.y;
            r[p.off_a2loc] = a1.x; r[p.off_a2loc + 1] = a1.y;
        }
        if (UNDO) { if (!p.ego_blind) r[p.off_hold] = ((in.holdmask >> k) & 1u) ? 1.0f : 0.0f; }
        else if (!p.ego_blind && ((in.holdmask >> k) & 1u)) r[p.off_hold] = 1.0f;   // :154
        if (blind) {
#pragma unroll
            for (int c = 0; c < 4; ++c) r[p.off_hidden + c] = 1.0f;            // :109
        } else {
            float2 me = a0;
            if (k == 1) me = a1;
            if (A > 2 && k >= 2) {
                uint32_t cellk = e.acell[2];
                if (A > 3 && k == 3) cellk = e.acell[A > 3 ? 3 : 2];
                me = tb.xyf[cellk];
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < 3 && c >= NF) continue;                 // a food this level never holds: features stay 0
                const uint32_t w = (c < 3) ? in.fword[c] : in.pword;
                if (obj_alive(w)) {                                            // absent channel: all zeros already
                    float hid, ex, ey, st;
                    channel_features(tb, w, me, fow, c, hid, ex, ey, st);
                    r[p.off_hidden + c] = hid;
                    r[p.off_encx + c] = ex;
                    r[p.off_ency + c] = ey;
                    if (c < 3) r[p.off_state + c] = st;
                } else if (UNDO) {                                             // ... unless the row still holds the last step
                    r[p.off_hidden + c] = 0.0f;
                    r[p.off_encx + c] = 0.0f;
                    r[p.off_ency + c] = 0.0f;
                    if (c < 3) r[p.off_state + c] = 0.0f;
                }
            }
        }
        r[p.off_ts] = ts;                                                      // :146
    }
    if (UNDO) {                                     // bits shown last step that are gone (a new episode began)
        for (uint32_t m = shown_completed & ~e.completed; m != 0; m &= m - 1) {
            float* r = row + p.off_completed + (__ffs((int)m) - 1);
#pragma unroll
            for (int k = 0; k < A; ++k) r[k * p.F] = 0.0f;
        }
    }
    // completed_subtasks: the same bits for every observer -- walk the set bits once
    for (uint32_t m = (UNDO ? (e.completed & ~shown_completed) : e.completed); m != 0; m &= m - 1) {
        float* r = row + p.off_completed + (__ffs((int)m) - 1);
#pragma unroll
        for (int k = 0; k < A; ++k) r[k * p.F] = 1.0f;
    }
}

// one observer's row as bytes: value + BIAS.  BIAS = 128 with 0x80-filled rows of F bytes (expanded to floats on
// the way out); BIAS = 0 with zero-filled rows of F - 1 bytes = the compact integer format itself (int8, the
// `timestep` column -- the last key -- cut out)
template <int A, int NOBJ, int NF, uint32_t BIAS = 128u>
__device__ __forceinline__ void build_row_u8_one(const Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                                 const Info& in, int k, uint8_t* __restrict__ r) {
    const uint32_t xy0 = tb.xy16[e.acell[0]], xy1 = tb.xy16[e.acell[1]];
    const uint32_t c0 = e.comm & 0xFFFFu, c1 = e.comm >> 16;
    const float fow = (float)p.fow;
    const bool blind = p.blind[k] != 0;
    if (c0 != OCK_COMM_NONE) r[p.off_a1comm + c0] = BIAS + 1;
    if (c1 != OCK_COMM_NONE) r[p.off_a2comm + c1] = BIAS + 1;
    if (!blind) {
        r[p.off_a1loc] = BIAS + (xy0 & 0xFF);  r[p.off_a1loc + 1] = BIAS + (xy0 >> 8);
        r[p.off_a2loc] = BIAS + (xy1 & 0xFF);  r[p.off_a2loc + 1] = BIAS + (xy1 >> 8);
    }
    if (!p.ego_blind && ((in.holdmask >> k) & 1u)) r[p.off_hold] = BIAS + 1;
    if (blind) {
#pragma unroll
        for (int c = 0; c < 4; ++c) r[p.off_hidden + c] = BIAS + 1;
    } else {
        const float2 me = tb.xyf[e.acell[k]];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < 3 && c >= NF) continue;                     // a food this level never holds: features stay 0
            const uint32_t w = (c < 3) ? in.fword[c] : in.pword;
            if (obj_alive(w)) {
                float hid, ex, ey, st;
                channel_features(tb, w, me, fow, c, hid, ex, ey, st);
                r[p.off_hidden + c] = (uint8_t)(BIAS + (int)hid);
                r[p.off_encx + c] = (uint8_t)(BIAS + (int)ex);
                r[p.off_ency + c] = (uint8_t)(BIAS + (int)ey);
                if (c < 3) r[p.off_state + c] = (uint8_t)(BIAS + (int)st);
            }
        }
    }
    for (uint32_t m = e.completed; m != 0; m &= m - 1) r[p.off_completed + (__ffs((int)m) - 1)] = BIAS + 1;
}

template <int A, int NOBJ, int NF>
__device__ __forceinline__ void build_rows_u8(const Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                              const Info& in, uint8_t* __restrict__ row /* 0x80-filled */) {
#pragma unroll
    for (int k = 0; k < A; ++k) build_row_u8_one<A, NOBJ, NF>(e, p, tb, in, k, row + k * p.F);
}

// compact integer rows: int8 [A, F-1] per env, zero-filled on entry
template <int A, int NOBJ, int NF>
__device__ __forceinline__ void build_rows_i8(const Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                              const Info& in, uint8_t* __restrict__ row /* zero-filled */) {
#pragma unroll
    for (int k = 0; k < A; ++k) build_row_u8_one<A, NOBJ, NF, 0u>(e, p, tb, in, k, row + k * (p.F - 1));
}

// byte offset of row slot `slot` inside a warp's row buffer (grouped rows: grp_pad bytes after every 2^grp_shift rows)
__device__ __forceinline__ uint32_t row_offset(const OcParams& p, int slot) {
    return (uint32_t)slot * (uint32_t)p.row_stride + (uint32_t)(slot >> p.grp_shift) * (uint32_t)p.grp_pad;
}

// warp-cooperative fill of the warp's 32 rows with the "all features 0.0" pattern
template <bool ROWF>
__device__ __forceinline__ void warp_clear_rows(uint8_t* wrows, int bytes, int lane) {
    const uint32_t z = ROWF ? 0u : 0x80808080u;
    uint4* d = reinterpret_cast<uint4*>(wrows);
    for (int i = lane; i < (bytes >> 4); i += 32) d[i] = make_uint4(z, z, z, z);
}

// byte k of w (value + 128) -> float, on the integer/FP32 pipes only: PRMT builds 0x4B0000bb
// (= 8388608 + bb as a float), one FADD removes 8388608 + 128.  (I2F runs at quarter rate.)
__device__ __forceinline__ float biased_byte_to_float(uint32_t w, int k) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + (uint32_t)k)) - 8388736.0f;
}

#ifndef OCK_HOST_EMU
// TMA bulk store (cp.async.bulk, SASS UBLKCP): the copy engine streams a contiguous shared-memory
// region to global memory while the warp goes on with the next step; no LDS/STG instructions.
__device__ __forceinline__ void tma_store(void* gdst, const void* ssrc, uint32_t bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// ---- TMA bulk LOAD of a contiguous global region into shared memory, completion on an mbarrier (one thread issues,
// everybody waits on the barrier's phase bit): the table blob arrives while the threads are busy with other things
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");     // visible to the async proxy
}
__device__ __forceinline__ void tma_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(sdst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .pred p;\n\tOCK_MBAR_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra OCK_MBAR_WAIT;\n\t}" :: "r"(b), "r"(parity) : "memory");
}
#endif
// rows may be overwritten again once the copy engine has READ them
__device__ __forceinline__ void rows_wait_read(const OcParams& p) {
#ifndef OCK_HOST_EMU
    if (p.use_tma) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
    __syncwarp();
}
// two row buffers used alternately: a buffer may be refilled once every bulk copy but the most
// recent one (which reads the OTHER buffer) has been read
__device__ __forceinline__ void rows_wait_read_but_one(const OcParams& p) {
#ifndef OCK_HOST_EMU
    if (p.use_tma) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
#endif
    __syncwarp();
}
// before the CTA exits: the copy engine must have read the rows out of its shared memory; the
// global writes themselves complete, like any store, by the end of the grid
__device__ __forceinline__ void rows_wait_done(const OcParams& p) {
#ifndef OCK_HOST_EMU
    if (p.use_tma) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}

// warp-cooperative expansion: the warp's 32 rows in shared memory -> float32 rows in global
// memory, consecutive lanes writing consecutive 16-byte (or 4-byte) pieces of one contiguous region.
template <bool ROWF>
__device__ __forceinline__ void warp_expand_rows(const OcParams& p, const uint8_t* __restrict__ wrows,
                                                 float* __restrict__ out /* warp's first env row */,
                                                 int nvalid, int lane) {
    if (ROWF) {                                     // float rows: plain 16-byte copy
        const int r4 = p.row_bytes >> 2;
        const int total = nvalid * r4;
        const float4* i4 = reinterpret_cast<const float4*>(wrows);
        float4* o4 = reinterpret_cast<float4*>(out);
#ifndef OCK_HOST_EMU
        if (p.use_tma) {
            // generic-proxy writes of every lane -> visible to the async proxy, then ONE bulk store
            // of the warp's 32 contiguous rows (11.8 KB for cfg2) by one lane
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (p.use_tma == 1) {                    // contiguous rows: one copy for the warp
                if (lane == 0) tma_store(out, wrows, (uint32_t)(total * 16));
            } else if (p.use_tma == 3) {             // grouped rows: one copy per group, by the group's first lane
                const int g = 1 << p.grp_shift;
                if ((lane & (g - 1)) == 0 && lane < nvalid)
                    tma_store(out + (size_t)lane * p.row_bytes, wrows + row_offset(p, lane),
                              (uint32_t)(min(g, nvalid - lane) * p.row_bytes * 4));
            } else if (lane < nvalid) {              // padded rows (use_tma == 2): one copy per env row
                tma_store(out + (size_t)lane * p.row_bytes, wrows + (size_t)lane * p.row_stride, (uint32_t)(p.row_bytes * 4));
            }
            return;
        }
#endif
        if (p.row_stride == p.row_bytes * 4 && p.grp_pad == 0) {      // contiguous rows
#pragma unroll 4
            for (int idx = lane; idx < total; idx += 32) __stcs(o4 + idx, i4[idx]);   // streaming: written once, read later by the learner
        } else {                                    // padded / grouped rows: env = idx / r4 by magic multiply
#pragma unroll 4
            for (int idx = lane; idx < total; idx += 32) {
                const int env = (int)__umulhi((uint32_t)idx, p.r4_magic);
                __stcs(o4 + idx, i4[(row_offset(p, env) >> 4) + (idx - env * r4)]);
            }
        }
    } else if ((p.row_bytes & 3) == 0) {
        const int r4 = p.row_bytes >> 2;            // float4 per env row
        const int total = nvalid * r4;
        float4* o4 = reinterpret_cast<float4*>(out);
        const bool contiguous = p.row_stride == p.row_bytes;
#pragma unroll 2
        for (int idx = lane; idx < total; idx += 32) {
            int off = idx;                                          // word offset inside the warp's rows
            if (!contiguous) {                                      // padded rows: env = idx / r4 by magic multiply
                const int env = (int)__umulhi((uint32_t)idx, p.r4_magic);
                off = env * (p.row_stride >> 2) + (idx - env * r4);
            }
            const uint32_t w = reinterpret_cast<const uint32_t*>(wrows)[off];
            float4 v;
            v.x = biased_byte_to_float(w, 0);
            v.y = biased_byte_to_float(w, 1);
            v.z = biased_byte_to_float(w, 2);
            v.w = biased_byte_to_float(w, 3);
            __stcs(o4 + idx, v);
        }
    } else {
        const int rf = p.row_bytes;
        const int total = nvalid * rf;
        for (int idx = lane; idx < total; idx += 32) {
            const int env = (int)__umulhi((uint32_t)idx, p.rf_magic);
            const uint8_t b = wrows[env * p.row_stride + (idx - env * rf)];
            out[idx] = (float)((int)b - 128);
        }
    }
}

// timestep = float32(t / max_num_timesteps) (overcooked_env.py:146): from the shared-memory copy of the
// table when it fits -- a global load here would queue behind this warp's own obs stores -- else
// computed (IEEE f64 division + one rounding, identical to the host table)
template <int A, int NOBJ>
__device__ __forceinline__ float timestep_of(const Env<A, NOBJ>& e, const OcParams& p, const Tables& tb) {
    const uint32_t t = e.w0 & 0xFFFFu;
    // the table holds t = 0..T; an env stepped past done without a reset (auto-reset off) reports t / T > 1 like the reference
    return (p.o_ts >= 0 && t <= (uint32_t)p.T) ? tb.ts[t] : (float)__ddiv_rn((double)t, (double)p.T);
}

template <int A, int NOBJ, int NF, bool ROWF, bool ROT = false>
__device__ __forceinline__ void fill_rows(const Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                          const Info& in, float ts, uint8_t* myrow, int rot = 0) {
    if (ROWF) build_rows_f32<A, NOBJ, NF, ROT, false>(e, p, tb, in, ts, reinterpret_cast<float*>(myrow), rot);
    else build_rows_u8<A, NOBJ, NF>(e, p, tb, in, myrow);
}

// byte rows only: after the expansion (and a __syncwarp) each thread stores the timestep feature
// of its own env rows (the byte rows carry 0.0 there)
template <int A>
__device__ __forceinline__ void store_timesteps(const OcParams& p, float* __restrict__ env_row, float ts) {
#pragma unroll
    for (int k = 0; k < A; ++k) env_row[k * p.F + p.off_ts] = ts;
}

// Chained steps (oc_kernels.cu): tell the successor of this warp-chunk that its new state is in memory.  Called after
// the rows were filled and BEFORE they are handed to the copy engine: the release waits for the warp's earlier stores
// (state, reward, done -- acknowledged long ago at this point) but not for a bulk copy of its own.
__device__ __forceinline__ void chain_release(uint32_t* flag, uint32_t val, int lane) {
#ifndef OCK_HOST_EMU
    if (flag != nullptr && lane == 0)
        asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flag), "r"(val) : "memory");
#endif
}

// ---- observation emission of one warp's 32 envs.  The warp owns nbuf buffers of p.nb env rows
// (1 x 32, or 2 x 16 / 1 x 16 / 2 x 8 ... for float rows); the envs go out in 32 / nb passes, lanes
// [pass * nb, pass * nb + nb) filling buffer (pass mod nbuf) in their pass, one bulk copy per pass.
// With two buffers a pass only waits for the copy issued TWO passes ago, so filling and dynamics
// overlap the copy engine's drain.  Each pass waits for its buffer to be read out, clears it (unless
// the caller says the rows are clean), fills it, and hands it to the copy engine (or stores it with
// the warp).
template <int A, int NOBJ, int NF, int MODE /* 0 byte rows, 1 float rows, 2 float rows in several passes */,
          bool ROT = false /* the single-step kernel: observer rotation per lane octet, see build_rows_f32 */>
__device__ __forceinline__ void emit_obs(const Env<A, NOBJ>& e, const Info& in, bool valid, const OcParams& p,
                                         const Tables& tb, uint8_t* wrows, int lane,
                                         float* __restrict__ out_env0 /* warp's first env row */, int nvalid,
                                         bool clean_on_entry = false /* rows already clear and not in flight */,
                                         uint32_t* chain_flag = nullptr, uint32_t chain_val = 0 /* see chain_release */) {
    constexpr bool ROWF = MODE != 0;
    constexpr bool MULTI = MODE == 2;               // compile-time: e / in stay live across passes only here
    const float ts = valid ? timestep_of<A, NOBJ>(e, p, tb) : 0.0f;
    const int passes = MULTI ? p.obs_passes : 1;
    for (int pass = 0; pass < passes; ++pass) {
        uint8_t* buf = wrows;
        if (MULTI && p.nbuf == 2) buf += (pass & 1) * p.buf_bytes;
        if (!(clean_on_entry && pass < p.nbuf)) {
            if (MULTI && p.nbuf == 2) rows_wait_read_but_one(p); else rows_wait_read(p);
            warp_clear_rows<ROWF>(buf, MULTI ? p.buf_bytes : p.warp_row_bytes, lane);
            __syncwarp();
        }
        uint8_t* myrow = buf + row_offset(p, MULTI ? (lane & (p.nb - 1)) : lane);
        // starting observer per lane octet (single-pass, ungrouped float rows; see build_rows_f32)
        constexpr bool R = ROT && ROWF && !MULTI;
        const int rot = (R && p.grp_pad == 0 && p.obs_rot) ? ((lane >> 3) % A) : 0;
        if (valid && (!MULTI || (lane >> p.nb_shift) == pass)) fill_rows<A, NOBJ, NF, ROWF, R>(e, p, tb, in, ts, myrow, rot);
        __syncwarp();
        if (pass == 0) chain_release(chain_flag, chain_val, lane);
        const int first = MULTI ? (pass << p.nb_shift) : 0;
        const int nv = min(MULTI ? p.nb : 32, nvalid - first);
        if (nv > 0) warp_expand_rows<ROWF>(p, buf, out_env0 + (size_t)first * p.row_bytes, nv, lane);
    }
    if (!ROWF) {                                    // byte rows are never split into passes
        __syncwarp();                               // orders the float4 stores before the timestep patch
        if (valid) store_timesteps<A>(p, out_env0 + (size_t)lane * p.row_bytes, ts);
    }
}

// The fused kernel's steady state for single-pass float rows: this warp's rows still hold the same envs' previous
// observation (nobody else ever writes them), so instead of clearing 12 KB the threads take back what the previous fill
// put at data-dependent places and overwrite the rest (build_rows_f32<UNDO>).
template <int A, int NOBJ, int NF>
__device__ __forceinline__ void emit_obs_undo(const Env<A, NOBJ>& e, const Info& in, bool valid, const OcParams& p,
                                              const Tables& tb, uint8_t* wrows, int lane, float* __restrict__ out_env0,
                                              int nvalid, uint32_t shown_comm, uint32_t shown_completed) {
    const float ts = valid ? timestep_of<A, NOBJ>(e, p, tb) : 0.0f;
    rows_wait_read(p);                              // the copy engine has read the previous step's rows
    if (valid) build_rows_f32<A, NOBJ, NF, false, true>(e, p, tb, in, ts, reinterpret_cast<float*>(wrows + row_offset(p, lane)),
                                                        0, shown_comm, shown_completed);
    __syncwarp();
    if (nvalid > 0) warp_expand_rows<true>(p, wrows, out_env0, nvalid, lane);
}

// a single thread emits its own env's rows (rare path: terminal observations); its row must be
// clear on entry and is clear again on exit
template <int A, int NOBJ, int NF, bool ROWF>
__device__ __forceinline__ void thread_emit_rows(const Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                                 const Info& in, uint8_t* myrow, float* __restrict__ out) {
    const float ts = timestep_of<A, NOBJ>(e, p, tb);
    fill_rows<A, NOBJ, NF, ROWF>(e, p, tb, in, ts, myrow);
    if (ROWF) {
        for (int j = 0; j < p.row_bytes; ++j) out[j] = reinterpret_cast<const float*>(myrow)[j];
    } else {
        for (int j = 0; j < p.row_bytes; ++j) out[j] = (float)((int)myrow[j] - 128);
        store_timesteps<A>(p, out, ts);
    }
    const uint32_t z = ROWF ? 0u : 0x80808080u;
    for (int j = 0; j < (p.row_stride >> 2); ++j) reinterpret_cast<uint32_t*>(myrow)[j] = z;
}

// infos["terminal_observation"] of the envs of this warp that just finished (SB3 VecEnv contract);
// rare path: waits for every pending copy, clears the warp's rows and lets the finished lanes emit
// their rows one by one (lanes that share a row slot, nb < 32, take turns).
template <int A, int NOBJ, int NF, int MODE>
__device__ __forceinline__ void warp_terminal_obs(const Env<A, NOBJ>& e, const Info& in, bool fin, const OcParams& p,
                                                  const Tables& tb, uint8_t* wrows, int lane,
                                                  float* __restrict__ term_row) {
    constexpr bool ROWF = MODE != 0;
    rows_wait_read(p);
    warp_clear_rows<ROWF>(wrows, p.warp_row_bytes, lane);
    __syncwarp();
    if (MODE != 2) {
        if (fin) thread_emit_rows<A, NOBJ, NF, ROWF>(e, p, tb, in, wrows + row_offset(p, lane), term_row);
        __syncwarp();
        return;
    }
    uint8_t* myrow = wrows + row_offset(p, lane & (p.nb - 1));
    for (int pass = 0; pass < p.obs_passes; ++pass) {
        if (fin && (lane >> p.nb_shift) == pass) thread_emit_rows<A, NOBJ, NF, ROWF>(e, p, tb, in, myrow, term_row);
        __syncwarp();
    }
}

// =============================================================================================
// Compact integer format produced by the step / reset kernels themselves (kernel MODE 3): int8 [E, A, F-1]
// rows + the f32 clock [E] -- what a consumer on the far side of PCIe gets (oc_step_host_i8 /
// oc_step_host_block).  A warp's 32 rows are 32 * A * (F-1) contiguous bytes in shared memory (a quarter of
// the float rows, so four times the warps fit on an SM) and leave through ONE bulk copy; the destination may
// be device memory or page-locked host memory (the copy engine then writes across PCIe while the kernel runs).
// p.row_bytes = A * (F-1) = bytes of one env's rows, p.row_stride == p.row_bytes (contiguous).
// =============================================================================================
__device__ __forceinline__ void warp_store_packed(const OcParams& p, const uint8_t* __restrict__ wrows,
                                                  uint8_t* __restrict__ out /* warp's first env row */, int nvalid, int lane) {
    const int total = nvalid * p.row_bytes;                      // bytes; rows are contiguous
#ifndef OCK_HOST_EMU
    if (p.use_tma && (total & 15) == 0) {                        // every full warp: 32 * row_bytes is a multiple of 32
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) tma_store(out, wrows, (uint32_t)total);
        return;
    }
#endif
    const int nw = total >> 2;                                   // ragged last warp: word copy + byte tail
    const uint32_t* i32 = reinterpret_cast<const uint32_t*>(wrows);
    uint32_t* o32 = reinterpret_cast<uint32_t*>(out);            // out = base + env0 * row_bytes with env0 % 32 == 0: 4-byte aligned
    for (int idx = lane; idx < nw; idx += 32) o32[idx] = i32[idx];
    for (int idx = (nw << 2) + lane; idx < total; idx += 32) out[idx] = wrows[idx];
}

template <int A, int NOBJ, int NF>
__device__ __forceinline__ void emit_obs_packed(const Env<A, NOBJ>& e, const Info& in, bool valid, const OcParams& p,
                                                const Tables& tb, uint8_t* wrows, int lane,
                                                uint8_t* __restrict__ out_env0, float* __restrict__ ts_env0, int nvalid,
                                                bool clean_on_entry = false, uint32_t* chain_flag = nullptr, uint32_t chain_val = 0) {
    const float ts = valid ? timestep_of<A, NOBJ>(e, p, tb) : 0.0f;
    if (!clean_on_entry) {
        rows_wait_read(p);
        warp_clear_rows<true>(wrows, p.warp_row_bytes, lane);
        __syncwarp();
    }
    if (valid) build_rows_i8<A, NOBJ, NF>(e, p, tb, in, wrows + (size_t)lane * p.row_stride);
    __syncwarp();
    chain_release(chain_flag, chain_val, lane);
    if (nvalid > 0) warp_store_packed(p, wrows, out_env0, nvalid, lane);
    if (valid && ts_env0 != nullptr) ts_env0[lane] = ts;         // 128 contiguous bytes per warp
}

// terminal rows of the envs of this warp that just finished, compact format; rare path
template <int A, int NOBJ, int NF>
__device__ __forceinline__ void warp_terminal_obs_packed(const Env<A, NOBJ>& e, const Info& in, bool fin, const OcParams& p,
                                                         const Tables& tb, uint8_t* wrows, int lane,
                                                         uint8_t* __restrict__ term_row, float* __restrict__ term_ts) {
    rows_wait_read(p);
    warp_clear_rows<true>(wrows, p.warp_row_bytes, lane);
    __syncwarp();
    if (fin) {
        uint8_t* myrow = wrows + (size_t)lane * p.row_stride;
        build_rows_i8<A, NOBJ, NF>(e, p, tb, in, myrow);
        for (int j = 0; j < p.row_bytes; ++j) { term_row[j] = myrow[j]; myrow[j] = 0; }
        if (term_ts != nullptr) *term_ts = timestep_of<A, NOBJ>(e, p, tb);
    }
    __syncwarp();
}

// terminal bookkeeping + in-place reset of one finished env (SB3 VecEnv auto-reset contract)
template <int A, int NOBJ>
__device__ __forceinline__ void finish_episode(Env<A, NOBJ>& e, const OcParams& p, const Tables& tb, uint32_t env_id) {
    e.w5 = (e.w5 & ~0xFFu) | (uint32_t)__popc(e.completed);    // episode_recorder.py:29
    e.episodes += 1;
    env_reset<A, NOBJ>(e, p, tb, nullptr, env_id);
}

// oc_step, part 1: dynamics + reward / done outputs (does not touch the observation rows)
template <int A, int NOBJ, int NF>
__device__ __forceinline__ Info step_logic(Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                           const int (&nav)[A], int comm0, int comm1, uint32_t env,
                                           float* __restrict__ rew32, double* __restrict__ rew64,
                                           uint8_t* __restrict__ done_out, bool& done,
                                           bool rew_per_env = false /* rew32 is f32 [E], not [E, A] */) {
    double reward;
    // out-of-range message index -> zero vector (the reference raises IndexError)
    const int c0 = ((uint32_t)comm0 < (uint32_t)p.C) ? comm0 : (int)OCK_COMM_NONE;
    const int c1 = ((uint32_t)comm1 < (uint32_t)p.C) ? comm1 : (int)OCK_COMM_NONE;
    const Info in = env_step<A, NOBJ, NF>(e, p, tb, nav, c0, c1, reward, done);
    if (rew64 != nullptr) rew64[env] = reward;
    if (rew32 != nullptr) {
        const float r = (float)reward;
        if (rew_per_env) rew32[env] = r;
        else {
#pragma unroll
            for (int k = 0; k < A; ++k) rew32[(size_t)env * A + k] = r;
        }
    }
    done_out[env] = done ? 1 : 0;
    return in;
}

// one env, one step of the fused synthetic rollout: Philox actions (nav ~ U{0..3}, comm ~ U{0..C-1}),
// auto-reset always on.  Draw layout (same in oracle/oc_oracle.c): counter (env, global step,
// 'ACTS', 0); nav_k = bits [2k, 2k+2) of word 0; comm_0/1 = mulhi(word 1/2, C).
template <int A, int NOBJ, int NF, bool ROWF>
__device__ __forceinline__ Info rollout_logic(Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                              uint32_t env, uint32_t s, uint32_t step0,
                                              float* __restrict__ rew32, uint8_t* __restrict__ done_out,
                                              int32_t* __restrict__ actions_out,
                                              const int32_t* __restrict__ actions_in = nullptr) {
    int nav[A];
    int c0, c1;
    if (actions_in != nullptr) {                     // oc_replay: open-loop action sequence from HBM
        const int2* a2 = reinterpret_cast<const int2*>(actions_in) + ((size_t)s * p.E + env) * A;
        int cm[A];
#pragma unroll
        for (int k = 0; k < A; ++k) { const int2 v = __ldg(a2 + k); nav[k] = v.x & 3; cm[k] = v.y; }
        c0 = ((uint32_t)cm[0] < (uint32_t)p.C) ? cm[0] : (int)OCK_COMM_NONE;
        c1 = ((uint32_t)cm[1] < (uint32_t)p.C) ? cm[1] : (int)OCK_COMM_NONE;
    } else {
        uint32_t r[4];
        philox4x32_10(env, step0 + s, 0x41435453u /*'ACTS'*/, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r);
#pragma unroll
        for (int k = 0; k < A; ++k) nav[k] = (r[0] >> (2 * k)) & 3;
        c0 = (int)__umulhi(r[1], (uint32_t)p.C); c1 = (int)__umulhi(r[2], (uint32_t)p.C);
    }
    if (actions_out != nullptr) {
        int32_t* ao = actions_out + ((size_t)s * p.E + env) * A * 2;
#pragma unroll
        for (int k = 0; k < A; ++k) { ao[2 * k] = nav[k]; ao[2 * k + 1] = (k == 0) ? c0 : (k == 1 ? c1 : 0); }
    }
    double reward; bool done;
    Info in = env_step<A, NOBJ, NF>(e, p, tb, nav, c0, c1, reward, done);
    if (rew32 != nullptr) {
        const float rr = (float)reward;
#pragma unroll
        for (int k = 0; k < A; ++k) rew32[((size_t)s * p.E + env) * A + k] = rr;
    }
    if (done_out != nullptr) done_out[(size_t)s * p.E + env] = done ? 1 : 0;
    if (done) {                                      // no terminal observation in the fused rollout: rows untouched
        finish_episode<A, NOBJ>(e, p, tb, env);
        in = gather_info<A, NOBJ, NF>(e, p, tb);
    }
    return in;
}

// oc_reset / initial bring-up of one env
template <int A, int NOBJ>
__device__ __forceinline__ void reset_logic(Env<A, NOBJ>& e, const OcParams& p, const Tables& tb, uint32_t env,
                                            bool initial, const uint8_t* __restrict__ mask,
                                            const int32_t* __restrict__ placements) {
    const int32_t* pl = placements ? placements + (size_t)env * p.nrandom : nullptr;
    if (initial) {
        e.episodes = 0; e.w5 = 0; e.w15 = 0;
        e.comm = 0;                               // one-hot at index 0 for both agents (overcooked_env.py:89-91)
        env_reset<A, NOBJ>(e, p, tb, pl, env);
    } else if (mask == nullptr || mask[env]) {
        e.episodes += 1;
        env_reset<A, NOBJ>(e, p, tb, pl, env);
    }
}

// ---- oc_set_state: one 16-byte plane of an imported env, sanitised (agent cells and live object cells clamped to the
// grid, holders to {0..A-1, none}, empty slots -> the canonical dead word, subtask bits and message indices cut to
// the configured S and C): whatever the words are, the kernels stay inside their tables and rows
__device__ __forceinline__ uint32_t sanitize_object_word(const OcParams& p, uint32_t o) {
    if ((o & 0xFu) == 0u) return OCK_DEAD;
    uint32_t holder = (o >> 8) & 7u, cell = (o >> 16) & 0xFFu;
    if (holder >= (uint32_t)p.A) holder = OCK_HOLDER_NONE;
    cell = min(cell, (uint32_t)p.ncell - 1u);
    return (o & 0xFF0000FFu) | (holder << 8) | (cell << 16);
}
__device__ __forceinline__ uint4 sanitize_state_plane(const OcParams& p, int plane, uint4 v) {
    if (plane == 0) {                                    // completed / count bits index the observation row
        const uint32_t smask = p.S >= 32 ? 0xFFFFFFFFu : ((1u << p.S) - 1u);
        v.z &= smask; v.w &= smask;
    } else if (plane == 1) {
        uint32_t cells = 0;
        for (int k = 0; k < 4; ++k) cells |= min((v.x >> (8 * k)) & 0xFFu, (uint32_t)p.ncell - 1u) << (8 * k);
        v.x = cells;
    } else if (plane == 2) {
        v.x = sanitize_object_word(p, v.x); v.y = sanitize_object_word(p, v.y);
        v.z = sanitize_object_word(p, v.z); v.w = sanitize_object_word(p, v.w);
    } else if (plane == 3) {
        v.x = sanitize_object_word(p, v.x); v.y = sanitize_object_word(p, v.y);
        uint32_t c0 = v.z & 0xFFFFu, c1 = v.z >> 16;     // message indices write one-hots into the row
        if (c0 >= (uint32_t)p.C) c0 = OCK_COMM_NONE;
        if (c1 >= (uint32_t)p.C) c1 = OCK_COMM_NONE;
        v.z = c0 | (c1 << 16);
    }
    return v;
}

// ---- compact integer format (oc_pack_obs_i8): every key of a row except the clock is a small
// integer (get_observation2 builds them as int64 arrays, overcooked_env.py:145-157), so a consumer
// on the far side of PCIe gets them as int8 [E, A, F-1] (key order kept, the `timestep` column cut
// out) plus the clock as f32 [E] (the same for every agent of an env): 4x fewer bytes, same values.
// One thread produces word `w` = 4 consecutive output bytes (rows are F-1 bytes, so a word may
// straddle two rows); the first E threads also copy their env's clock.
__device__ __forceinline__ void pack_i8_word(const OcParams& p, const float* __restrict__ obs,
                                             int8_t* __restrict__ out, float* __restrict__ ts, uint32_t w) {
    const uint32_t Fm = (uint32_t)p.F - 1u, F = (uint32_t)p.F, off_ts = (uint32_t)p.off_ts;
    const uint32_t nbytes = (uint32_t)p.E * (uint32_t)p.A * Fm;
    if (ts != nullptr && w < (uint32_t)p.E) ts[w] = obs[(size_t)w * p.row_bytes + off_ts];
    const uint32_t j0 = w * 4u;
    if (j0 >= nbytes) return;
    uint32_t r = j0 / Fm, c = j0 - r * Fm;            // observer row, column inside the packed row
    uint32_t word = 0;
    const uint32_t n = min(4u, nbytes - j0);
    for (uint32_t k = 0; k < n; ++k) {
        const float v = obs[(size_t)r * F + c + (c >= off_ts ? 1u : 0u)];
        word |= ((uint32_t)(int)v & 0xFFu) << (8u * k);
        if (++c == Fm) { c = 0; ++r; }
    }
    if (n == 4u) *reinterpret_cast<uint32_t*>(out + j0) = word;
    else for (uint32_t k = 0; k < n; ++k) out[j0 + k] = (int8_t)(word >> (8u * k));
}

// ---- terminal rows of the host path: of the E rows in `term` only those of envs that just finished are
// wanted on the host.  Finished env number i (env index idx[i]) is copied to slot i of a dense buffer, as
// float rows (out_f32 [n, A*F]) or in the compact format (out_i8 [n, A*(F-1)] + out_ts [n]); thread `tid` of
// `nthreads` cooperating on one row.
__device__ __forceinline__ void gather_term_row(const OcParams& p, const float* __restrict__ term, int e, int i,
                                                float* __restrict__ out_f32, int8_t* __restrict__ out_i8,
                                                float* __restrict__ out_ts, int tid, int nthreads) {
    const float* src = term + (size_t)e * p.row_bytes;
    if (out_f32 != nullptr) {
        float* dst = out_f32 + (size_t)i * p.row_bytes;
        for (int j = tid; j < p.row_bytes; j += nthreads) dst[j] = src[j];
    } else {
        const int Fm = p.F - 1, n8 = p.A * Fm;
        int8_t* dst = out_i8 + (size_t)i * n8;
        for (int j = tid; j < n8; j += nthreads) {
            const int k = j / Fm, c = j - k * Fm;
            dst[j] = (int8_t)(int)src[k * p.F + c + (c >= p.off_ts ? 1 : 0)];
        }
        if (tid == 0) out_ts[i] = src[p.off_ts];
    }
}

}  // namespace ock
