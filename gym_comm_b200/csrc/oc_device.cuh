// Device-side Overcooked dynamics for sm_100a: one THREAD owns one env for the branchy integer
// logic (a warp instruction then advances 32 envs), and the WARP cooperatively expands the 32
// envs' compact observation rows into coalesced 16-byte stores.  Reference semantics and their
// file:line anchors are listed next to each block; DESIGN.md explains the mapping.
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
    const uint8_t*  tile;     // [ncell]
    const uint8_t*  mv;       // [ncell*4] inbounds(cell + NAV[a])                   world.py:317-320
    const uint8_t*  xy;       // [ncell*2]
    const uint8_t*  dmin;     // [ncell] min over Delivery tiles of pd + manhattan   overcooked_environment.py:383-388
    const uint8_t*  counters; // [ncounters] Counter cells, reading order            overcooked_environment.py:164
    const uint8_t*  pd;       // [ncell*ncell] World.get_path_distance_between       world.py:114-131
};

__device__ __forceinline__ Tables make_tables(const OcParams& p, const uint8_t* smem) {
    Tables t;
    t.q = reinterpret_cast<const double*>(smem + p.o_q);
    t.tmlut = reinterpret_cast<const uint32_t*>(smem + p.o_tmlut);
    t.tile = smem + p.o_tile;
    t.mv = smem + p.o_mv;
    t.xy = smem + p.o_xy;
    t.dmin = smem + p.o_dmin;
    t.counters = smem + p.o_counters;
    t.pd = smem + p.o_pd;
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
__device__ __forceinline__ void load_env(Env<A, NOBJ>& e, const uint4* __restrict__ st, int E, int i) {
    const uint4 a = st[i], b = st[E + i], c = st[2 * E + i], d = st[3 * E + i];
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
    uint32_t o[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int s = 0; s < NOBJ; ++s) o[s] = e.obj[s];
    st[i] = make_uint4(e.w0, e.episodes, e.completed, e.countbits);
    st[E + i] = make_uint4(cells, e.w5, (uint32_t)e.ranks, (uint32_t)(e.ranks >> 32));
    st[2 * E + i] = make_uint4(o[0], o[1], o[2], o[3]);
    st[3 * E + i] = make_uint4(o[4], o[5], e.comm, e.w15);
}

__device__ __forceinline__ uint32_t obj_contents(uint32_t o) { return o & 0xFu; }
__device__ __forceinline__ uint32_t obj_chopped(uint32_t o) { return (o >> 4) & 7u; }
__device__ __forceinline__ uint32_t obj_holder(uint32_t o) { return (o >> 8) & 7u; }
__device__ __forceinline__ uint32_t obj_cell(uint32_t o) { return (o >> 16) & 0xFFu; }
__device__ __forceinline__ uint32_t obj_set_cell(uint32_t o, uint32_t c) { return (o & ~0x00FF0000u) | (c << 16); }
__device__ __forceinline__ uint32_t obj_set_holder(uint32_t o, uint32_t h) { return (o & ~0x00000700u) | (h << 8); }

// ---------------------------------------------------------------------------------------------
// load_level phase 4 (overcooked_environment.py:157-173): each random object goes to a Counter
// drawn uniformly from ALL Counter tiles, rejecting tiles already taken by an earlier phase-4
// object == sequential sampling without replacement.  Rare path (once per episode), kept out of
// line so its index arithmetic does not cost registers in the step loop.  Same algorithm in
// oracle/oc_oracle.c: draw j uses Philox word j of counter (env, episode, 'RESE', j / 4).
__device__ __noinline__ void draw_random_cells(const OcParams& p, const uint8_t* __restrict__ counters,
                                               uint32_t env_id, uint32_t episode, uint32_t* cell) {
    uint32_t r[8];
    philox4x32_10(env_id, episode, 0x52455345u, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r);
    philox4x32_10(env_id, episode, 0x52455345u, 1u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r + 4);
    uint32_t sorted[OCK_MAX_OBJECTS];
    int n = 0;
    for (int j = 0; j < p.nrandom; ++j) {
        uint32_t idx = __umulhi(r[j], (uint32_t)(p.ncounters - j));      // uniform in [0, n - j)
        for (int a = 0; a < n; ++a)
            if (sorted[a] <= idx) ++idx;                                   // skip counters already taken
        int pos = n;
        while (pos > 0 && sorted[pos - 1] > idx) { sorted[pos] = sorted[pos - 1]; --pos; }
        sorted[pos] = idx;
        ++n;
        cell[j] = counters[idx];
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
            for (int j = 0; j < p.nrandom; ++j) cell[j] = (uint32_t)placements[j];
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
// One env step.  nav[k] in [0,4); returns sparse reward, the returned (shaped) reward and done.
//   comm write + CAN_MOVE          gym_comm/envs/overcooked_env.py:227-262
//   t += 1                          overcooked_environment.py:213
//   check_collisions/is_collision   :543-613
//   interact per agent in order     gym_cooking/utils/interact.py:4-75
//   done                            :243-270
//   reward/subtask_reward           :399-432
//   calculate_reward_shaping x2     :272-397
template <int A, int NOBJ>
__device__ __forceinline__ void env_step(Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                         const int (&nav)[A], int comm0, int comm1,
                                         double& reward, bool& done) {
    // ---- comm channel write (overcooked_env.py:227-246)
    {
        const uint32_t c0 = p.comm_on ? (uint32_t)comm0 : OCK_COMM_NONE;
        const uint32_t c1 = (p.comm_on && !p.ego_led) ? (uint32_t)comm1 : OCK_COMM_NONE;
        e.comm = (c0 & 0xFFFFu) | (c1 << 16);
    }
    e.w0 += 1;   // t += 1 (t lives in the low 16 bits)
    const uint32_t t = e.w0 & 0xFFFFu;

    // ---- collisions, on the ORIGINAL actions for every pair (:543-613)
    uint32_t tgt[A], nxt[A];
    bool act[A], ex[A];
#pragma unroll
    for (int k = 0; k < A; ++k) {
        act[k] = p.can_move[k] != 0;                        // (0,0) iff CAN_MOVE false (:250-262)
        tgt[k] = tb.mv[e.acell[k] * 4 + nav[k]];            // inbounds(loc + action)
        // off-grid targets assert in the reference (world.py:314); here the clamped target is the
        // agent's own (floor) cell, i.e. the agent stays.
        nxt[k] = (act[k] && tb.tile[tgt[k]] == TILE_FLOOR) ? tgt[k] : e.acell[k];
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
        const uint32_t tg = tgt[k];
        const uint32_t tt = tb.tile[tg];
        uint32_t hv = 0, hm = 0;           // held object word / slot mask
#pragma unroll
        for (int s = 0; s < NOBJ; ++s) {
            const uint32_t o = e.obj[s];
            if (obj_contents(o) != 0 && obj_holder(o) == (uint32_t)k) { hv = o; hm = 1u << s; }
        }
        if (tt == TILE_FLOOR) {            // move; held object moves along (agent.py:311-314)
            e.acell[k] = tg;
#pragma unroll
            for (int s = 0; s < NOBJ; ++s)
                if ((hm >> s) & 1u) e.obj[s] = obj_set_cell(e.obj[s], tg);
            continue;
        }
        uint32_t ov = 0, om = 0;           // un-held object on the target tile (world.py:217-222)
#pragma unroll
        for (int s = 0; s < NOBJ; ++s) {
            const uint32_t o = e.obj[s];
            if (obj_contents(o) != 0 && obj_holder(o) == OCK_HOLDER_NONE && obj_cell(o) == tg) { ov = o; om = 1u << s; }
        }
        uint32_t newh = hv, newo = ov;
        if (hm != 0) {
            const uint32_t hc = obj_contents(hv), hch = obj_chopped(hv);
            const bool h_done = (hc & 7u) == hch;                    // every Food in its last state
            const uint32_t put = obj_set_holder(obj_set_cell(hv, tg), OCK_HOLDER_NONE);
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
                    newo = 0;                                        // absorbed object leaves the world
                }
            } else {                                                 // :48-59
                if (tt == TILE_CUTBOARD && (hc == 1u || hc == 2u || hc == 4u) && hch == 0u)
                    newh = hv | (hc << 4);                           // chop in hand (core.py:201-206)
                else
                    newh = put;                                      // put down
            }
        } else if (om != 0 && tt != TILE_DELIVERY && !p.allergic[k]) {   // pick up :64-71, agent.py:296-305
            newo = obj_set_holder(obj_set_cell(ov, e.acell[k]), (uint32_t)k);
        }
#pragma unroll
        for (int s = 0; s < NOBJ; ++s) {
            if ((hm >> s) & 1u) e.obj[s] = newh;
            if ((om >> s) & 1u) e.obj[s] = newo;
        }
    }
    e.w0 = t | (next_stamp << 16) | (nkeys << 24);

    // ---- done + sparse reward through the signature -> subtask-mask table
    uint32_t pres = 0, deliv = 0;
    int fresh_cell[3] = {-1, -1, -1};
#pragma unroll
    for (int s = 0; s < NOBJ; ++s) {
        const uint32_t o = e.obj[s];
        if (obj_contents(o) != 0) {
            const uint32_t sig = o & 0x7Fu;
            const uint32_t tm = tb.tmlut[sig];
            pres |= tm;
            if (obj_cell(o) == p.delivery0) deliv |= tm;     // only the first Delivery tile (:259,:402)
            if (sig == 1u) fresh_cell[0] = (int)obj_cell(o);
            if (sig == 2u) fresh_cell[1] = (int)obj_cell(o);
            if (sig == 4u) fresh_cell[2] = (int)obj_cell(o);
        }
    }
    const uint32_t dl = deliv & p.deliver_mask;
    done = (p.T != 0 && t >= (uint32_t)p.T) || (dl == p.deliver_mask);       // :243-270
    const uint32_t nw = pres & ~e.countbits & p.nondeliver_mask;             // count rose (:409-415)
    const int sparse = 3 * __popc(dl) + __popc(nw);
    e.countbits = pres & p.nondeliver_mask;
    e.completed |= dl | nw;                                                  // :425-426

    // ---- reward shaping (:272-397); f64 additions in reference order
    // (1) Chop subtasks still open
    int lenU = 0;
    int nf[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
        nf[f] = (fresh_cell[f] >= 0) ? __popc(~e.completed & p.chop_mask[f]) : 0;
        lenU += nf[f];
    }
    // (2) item-pair distances (:319-363): agent independent.  Items = Plate + the Foods of
    // recipes[0]; pairs in combinations() order (P,f0) (P,f1) (P,f2) (f0,f1) (f0,f2) (f1,f2).
    // pd(src, .) == MAX_PATH unless src is a floor cell, i.e. unless the FIRST item's object is
    // being carried (world.py:126-127), and each Food lives in exactly one object (domain rule), so
    // only a handful of table look-ups are ever needed.
    int lenP = p.npairs, minP = p.M;
    {
        bool any_held = false;
#pragma unroll
        for (int s = 0; s < NOBJ; ++s) any_held |= (obj_contents(e.obj[s]) != 0 && obj_holder(e.obj[s]) != OCK_HOLDER_NONE);
        if (any_held) {
            int fc[3] = {-1, -1, -1};
            bool fh[3] = {false, false, false};
#pragma unroll
            for (int s = 0; s < NOBJ; ++s) {
                const uint32_t o = e.obj[s];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (i < p.nfi && (o & p.fi_bit[i])) { fc[i] = (int)obj_cell(o); fh[i] = obj_holder(o) != OCK_HOLDER_NONE; }
            }
            int mP[3] = {p.M, p.M, p.M};
#pragma unroll
            for (int s = 0; s < NOBJ; ++s) {
                const uint32_t o = e.obj[s];
                if ((o & 8u) && obj_holder(o) != OCK_HOLDER_NONE) {          // a carried plate-bearing object
                    const uint8_t* row = tb.pd + obj_cell(o) * p.ncell;
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (i < p.nfi && fc[i] >= 0) mP[i] = min(mP[i], (int)row[fc[i]]);
                }
            }
            lenP = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (i < p.nfi && mP[i] != 0) { lenP += 1; minP = min(minP, mP[i]); }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = i + 1; j < 3; ++j) {
                    if (j < p.nfi) {
                        int m = p.M;
                        if (fh[i] && fc[j] >= 0) m = tb.pd[fc[i] * p.ncell + fc[j]];
                        if (m != 0) { lenP += 1; minP = min(minP, m); }
                    }
                }
            }
        }
    }
    // (3) Deliver subtasks: dish cell per subtask (table order)
    int dcell[OCK_MAX_DELIVER];
#pragma unroll
    for (int j = 0; j < OCK_MAX_DELIVER; ++j) {
        dcell[j] = -2;                                        // -2: subtask absent or completed
        if (j < p.ndeliver && !((e.completed >> p.deliver_idx[j]) & 1u)) {
            dcell[j] = -1;                                    // -1: no such dish in the world
#pragma unroll
            for (int s = 0; s < NOBJ; ++s)
                if ((e.obj[s] & 0x7Fu) == p.deliver_sig[j]) dcell[j] = (int)obj_cell(e.obj[s]);
        }
    }
    double shaped = (double)sparse;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const uint32_t ac = e.acell[a];
        const uint8_t* row = tb.pd + ac * p.ncell;
        double tp = 0.0;
        if (lenU > 0) {
            int minU = 1 << 20;
#pragma unroll
            for (int f = 0; f < 3; ++f)
                if (nf[f] > 0) minU = min(minU, (int)row[fresh_cell[f]]);
            tp = tb.q[minU + p.M + (lenU - 1) * 2 * p.M];                       // :303-304
            if (lenP > 0) tp = __dadd_rn(tp, (double)lenP);                       // :363
        } else if (lenP > 0) {
            tp = tb.q[minP + (lenP - 1) * p.M];                                   // :361
        }
        const int ax = tb.xy[ac * 2], ay = tb.xy[ac * 2 + 1];
#pragma unroll
        for (int j = 0; j < OCK_MAX_DELIVER; ++j) {
            if (dcell[j] == -2) continue;
            if (dcell[j] == -1) { tp = __dadd_rn(tp, 2.0); continue; }            // :377-378
            const int dc = dcell[j];
            const int d = row[dc] + abs(ax - (int)tb.xy[dc * 2]) + abs(ay - (int)tb.xy[dc * 2 + 1]);
            if (d == 0) tp = __dadd_rn(tp, tb.q[tb.dmin[ac]]);                    // :381-389
            else tp = __dadd_rn(tp, __dadd_rn(tb.q[d], 1.0));                     // :393
        }
        shaped = __dsub_rn(shaped, tp);                                           // overcooked_env.py:282
    }
    reward = shaped;
}

// ---------------------------------------------------------------------------------------------
// get_observation2 (gym_comm/envs/overcooked_env.py:105-159) for every observer of one env,
// written as ONE BYTE PER FEATURE (value + 128) into this env's shared-memory row.  Rows are
// pre-filled with 0x80 (= 0.0).  The timestep feature is left at 0.0 here; its float is stored
// by the owning thread after the warp's cooperative expansion (store_timesteps).
#define OCK_BIAS 128u

template <int A, int NOBJ>
__device__ __forceinline__ void env_build_rows(const Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                               uint8_t* __restrict__ row /* 0x80-filled, row_stride bytes */) {
    // per channel the object that is LAST in world.objects iteration order among those containing
    // it (last writer wins, :121-131).  Each Food lives in exactly one object; only the Plate
    // channel can have several candidates, ordered by (key creation rank, insertion stamp).
    int wx[4] = {0, 0, 0, 0}, wy[4] = {0, 0, 0, 0};
    uint32_t wst[4] = {0, 0, 0, 0};
    bool has[4] = {false, false, false, false};
    uint32_t pkey = 0;
    uint32_t holdmask = 0;                 // bit k: agent k holds something
#pragma unroll
    for (int s = 0; s < NOBJ; ++s) {
        const uint32_t o = e.obj[s];
        const uint32_t oc = obj_contents(o);
        if (oc == 0) continue;
        const uint32_t cell = obj_cell(o);
        const int x = tb.xy[cell * 2], y = tb.xy[cell * 2 + 1];
        holdmask |= (1u << obj_holder(o));
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if ((oc >> c) & 1u) { has[c] = true; wx[c] = x; wy[c] = y; wst[c] = (o >> (4 + c)) & 1u; }
        if (oc & 8u) {
            const uint32_t key = ((uint32_t)((e.ranks >> (4 * oc)) & 15ull) << 8) | (o >> 24);
            if (key > pkey) { pkey = key; has[3] = true; wx[3] = x; wy[3] = y; }
        }
    }
    const int x0 = tb.xy[e.acell[0] * 2], y0 = tb.xy[e.acell[0] * 2 + 1];
    const int x1 = tb.xy[e.acell[1] * 2], y1 = tb.xy[e.acell[1] * 2 + 1];
    const uint32_t c0 = e.comm & 0xFFFFu, c1 = e.comm >> 16;
#pragma unroll
    for (int k = 0; k < A; ++k) {
        uint8_t* r = row + k * p.F;
        const bool blind = p.blind[k] != 0;
        if (c0 != OCK_COMM_NONE) r[p.off_a1comm + c0] = OCK_BIAS + 1;
        if (c1 != OCK_COMM_NONE) r[p.off_a2comm + c1] = OCK_BIAS + 1;
        r[p.off_a1loc] = OCK_BIAS + (blind ? 0 : x0);  r[p.off_a1loc + 1] = OCK_BIAS + (blind ? 0 : y0);   // :139-143
        r[p.off_a2loc] = OCK_BIAS + (blind ? 0 : x1);  r[p.off_a2loc + 1] = OCK_BIAS + (blind ? 0 : y1);
        r[p.off_hold] = OCK_BIAS + ((!p.ego_blind && ((holdmask >> k) & 1u)) ? 1 : 0);                      // :154
        for (int i = 0; i < p.S; ++i) r[p.off_completed + i] = OCK_BIAS + ((e.completed >> i) & 1u);
        int ax = x0, ay = y0;
        if (k == 1) { ax = x1; ay = y1; }
        if (k >= 2) { ax = tb.xy[e.acell[k] * 2]; ay = tb.xy[e.acell[k] * 2 + 1]; }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int dx = 0, dy = 0; uint32_t st = 0, hid = 1;
            if (!blind) {
                if (has[c]) { dx = wx[c] - ax; dy = wy[c] - ay; st = wst[c]; }
                const bool near = (abs(dx) + abs(dy)) <= p.fow;                     // :133-135
                hid = near ? 0u : 1u;
                if (near) { dx = 0; dy = 0; }
            }
            r[p.off_hidden + c] = (uint8_t)(OCK_BIAS + hid);
            r[p.off_encx + c] = (uint8_t)(OCK_BIAS + dx);
            r[p.off_ency + c] = (uint8_t)(OCK_BIAS + dy);
            r[p.off_state + c] = (uint8_t)(OCK_BIAS + st);
        }
    }
}

// warp-cooperative fill of the warp's 32 byte-rows with 0x80 (= 0.0)
__device__ __forceinline__ void warp_zero_rows(uint8_t* wrows, int bytes, int lane) {
    uint4* d = reinterpret_cast<uint4*>(wrows);
    for (int i = lane; i < (bytes >> 4); i += 32) d[i] = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
}

// byte k of w (value + 128) -> float, on the integer/FP32 pipes only: PRMT builds 0x4B0000bb
// (= 8388608 + bb as a float), one FADD removes 8388608 + 128.  (I2F runs at quarter rate.)
__device__ __forceinline__ float biased_byte_to_float(uint32_t w, int k) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + (uint32_t)k)) - 8388736.0f;
}

// warp-cooperative expansion: 32 byte-rows in shared memory -> float32 rows in global memory,
// consecutive lanes writing consecutive 16-byte (or 4-byte) pieces of one contiguous region.
__device__ __forceinline__ void warp_expand_rows(const OcParams& p, const uint8_t* __restrict__ wrows,
                                                 float* __restrict__ out /* warp's first env row */,
                                                 int nvalid, int lane) {
    if ((p.row_bytes & 3) == 0) {
        const int r4 = p.row_bytes >> 2;             // float4 per env row
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
            __stcs(o4 + idx, v);                                    // streaming: written once, read later by the learner
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

// after the expansion (and a __syncwarp): each thread stores the timestep feature of its own env
// rows, timestep = float32(t / max_num_timesteps)  (overcooked_env.py:146)
template <int A>
__device__ __forceinline__ void store_timesteps(const OcParams& p, float* __restrict__ env_row, float ts) {
#pragma unroll
    for (int k = 0; k < A; ++k) env_row[k * p.F + p.off_ts] = ts;
}

// a single thread expands its own row (rare path: terminal observations)
__device__ __forceinline__ void thread_expand_row(const OcParams& p, const uint8_t* __restrict__ row, float ts,
                                                  float* __restrict__ out) {
    for (int j = 0; j < p.row_bytes; ++j) out[j] = (float)((int)row[j] - 128);
    for (int k = 0; k < p.A; ++k) out[k * p.F + p.off_ts] = ts;
}

template <int A, int NOBJ>
__device__ __forceinline__ float finish_obs(const Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                            uint8_t* myrow) {
    env_build_rows<A, NOBJ>(e, p, tb, myrow);
    return __ldg(p.ts_table + (e.w0 & 0xFFFFu));
}

// terminal bookkeeping + in-place reset of one finished env (SB3 VecEnv auto-reset contract)
template <int A, int NOBJ>
__device__ __forceinline__ void finish_episode(Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                               uint8_t* myrow, float* __restrict__ term_row, uint32_t env_id) {
    if (term_row != nullptr) {                       // infos["terminal_observation"]
        env_build_rows<A, NOBJ>(e, p, tb, myrow);
        thread_expand_row(p, myrow, __ldg(p.ts_table + (e.w0 & 0xFFFFu)), term_row);
        for (int j = 0; j < (p.row_stride >> 2); ++j) reinterpret_cast<uint32_t*>(myrow)[j] = 0x80808080u;
    }
    e.w5 = (e.w5 & ~0xFFu) | (uint32_t)__popc(e.completed);    // episode_recorder.py:29
    e.episodes += 1;
    env_reset<A, NOBJ>(e, p, tb, nullptr, env_id);
}

// everything one thread does for its env in oc_step between loading and storing the state
template <int A, int NOBJ>
__device__ __forceinline__ float step_one_env(Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                              const int (&nav)[A], int comm0, int comm1, uint32_t env,
                                              uint8_t* myrow,
                                             float* __restrict__ rew32, double* __restrict__ rew64,
                                             uint8_t* __restrict__ done_out, float* __restrict__ term_obs,
                                             uint32_t flags) {
    double reward; bool done;
    // out-of-range message index -> zero vector (the reference raises IndexError)
    const int c0 = ((uint32_t)comm0 < (uint32_t)p.C) ? comm0 : (int)OCK_COMM_NONE;
    const int c1 = ((uint32_t)comm1 < (uint32_t)p.C) ? comm1 : (int)OCK_COMM_NONE;
    env_step<A, NOBJ>(e, p, tb, nav, c0, c1, reward, done);
    if (rew64 != nullptr) rew64[env] = reward;
    if (rew32 != nullptr) {
        const float r = (float)reward;
#pragma unroll
        for (int k = 0; k < A; ++k) rew32[(size_t)env * A + k] = r;
    }
    done_out[env] = done ? 1 : 0;
    if (done && (flags & 1u /*OC_FLAG_AUTO_RESET*/))
        finish_episode<A, NOBJ>(e, p, tb, myrow, term_obs ? term_obs + (size_t)env * p.row_bytes : nullptr, env);
    return finish_obs<A, NOBJ>(e, p, tb, myrow);
}

// one env, one step of the fused synthetic rollout: Philox actions (nav ~ U{0..3}, comm ~ U{0..C-1}),
// auto-reset always on.  Draw layout (same in oracle/oc_oracle.c): counter (env, global step,
// 'ACTS', 0); nav_k = bits [2k, 2k+2) of word 0; comm_0/1 = mulhi(word 1/2, C).
template <int A, int NOBJ>
__device__ __forceinline__ float rollout_one_env(Env<A, NOBJ>& e, const OcParams& p, const Tables& tb,
                                                 uint32_t env, uint32_t s, uint32_t step0,
                                                 uint8_t* myrow, bool want_obs,
                                                float* __restrict__ rew32, uint8_t* __restrict__ done_out,
                                                int32_t* __restrict__ actions_out) {
    uint32_t r[4];
    philox4x32_10(env, step0 + s, 0x41435453u /*'ACTS'*/, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r);
    int nav[A];
#pragma unroll
    for (int k = 0; k < A; ++k) nav[k] = (r[0] >> (2 * k)) & 3;
    const int c0 = (int)__umulhi(r[1], (uint32_t)p.C), c1 = (int)__umulhi(r[2], (uint32_t)p.C);
    if (actions_out != nullptr) {
        int32_t* ao = actions_out + ((size_t)s * p.E + env) * A * 2;
#pragma unroll
        for (int k = 0; k < A; ++k) { ao[2 * k] = nav[k]; ao[2 * k + 1] = (k == 0) ? c0 : (k == 1 ? c1 : 0); }
    }
    double reward; bool done;
    env_step<A, NOBJ>(e, p, tb, nav, c0, c1, reward, done);
    if (rew32 != nullptr) {
        const float rr = (float)reward;
#pragma unroll
        for (int k = 0; k < A; ++k) rew32[((size_t)s * p.E + env) * A + k] = rr;
    }
    if (done_out != nullptr) done_out[(size_t)s * p.E + env] = done ? 1 : 0;
    if (done) finish_episode<A, NOBJ>(e, p, tb, myrow, nullptr, env);
    return want_obs ? finish_obs<A, NOBJ>(e, p, tb, myrow) : 0.0f;
}

// oc_reset / initial bring-up of one env
template <int A, int NOBJ>
__device__ __forceinline__ float reset_one_env(Env<A, NOBJ>& e, const OcParams& p, const Tables& tb, uint32_t env,
                                               bool initial, const uint8_t* __restrict__ mask,
                                               const int32_t* __restrict__ placements, bool want_obs,
                                               uint8_t* myrow) {
    const int32_t* pl = placements ? placements + (size_t)env * p.nrandom : nullptr;
    if (initial) {
        e.episodes = 0; e.w5 = 0; e.w15 = 0;
        e.comm = 0;                               // one-hot at index 0 for both agents (overcooked_env.py:89-91)
        env_reset<A, NOBJ>(e, p, tb, pl, env);
    } else if (mask == nullptr || mask[env]) {
        e.episodes += 1;
        env_reset<A, NOBJ>(e, p, tb, pl, env);
    }
    return want_obs ? finish_obs<A, NOBJ>(e, p, tb, myrow) : 0.0f;
}

}  // namespace ock
