// Host-side compilation of an oc_config into the kernel parameter block + table blob.
// Pure C++ (no CUDA calls) so the same code feeds oc_create and the CPU emulation harness that
// tests/emu uses to debug the device logic without a GPU (tests only; never a product path).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/overcooked_b200.h"
#include "oc_params.h"

namespace ock {

struct HostImage {
    OcParams p;
    std::vector<uint8_t> blob;
    std::vector<float> ts;
    int obs_off[OC_NUM_OBS_KEYS], obs_size[OC_NUM_OBS_KEYS];
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline void bfs_path_dist(const oc_config* c, std::vector<uint8_t>& pd) {
    // World.get_path_distance_between (gym_cooking/utils/world.py:61-92,114-131) tabulated.
    const int W = c->width, H = c->height, n = W * H, M = 2 * (W + H) + 1;
    pd.assign((size_t)n * n, (uint8_t)M);
    static const int dx[4] = {0, 0, -1, 1}, dy[4] = {1, -1, 0, 0};
    std::vector<int> dist(n), queue(n);
    for (int src = 0; src < n; ++src) {
        if (c->tiles[src] != 0) continue;            // source node missing -> MAX_PATH (world.py:126-127)
        std::fill(dist.begin(), dist.end(), -1);
        int qh = 0, qt = 0;
        dist[src] = 0; queue[qt++] = src;
        while (qh < qt) {
            const int u = queue[qh++], ux = u % W, uy = u / W;
            for (int a = 0; a < 4; ++a) {
                const int vx = ux + dx[a], vy = uy + dy[a];
                if (vx < 0 || vy < 0 || vx >= W || vy >= H) continue;
                const int v = vy * W + vx;
                if (c->tiles[v] == 0 && dist[v] < 0) { dist[v] = dist[u] + 1; queue[qt++] = v; }
            }
        }
        for (int dst = 0; dst < n; ++dst) {
            int best = M;
            if (c->tiles[dst] == 0) { if (dist[dst] >= 0) best = dist[dst]; }
            else {
                const int tx = dst % W, ty = dst / W;
                for (int a = 0; a < 4; ++a) {
                    const int vx = tx + dx[a], vy = ty + dy[a];
                    if (vx < 0 || vy < 0 || vx >= W || vy >= H) continue;
                    const int v = vy * W + vx;
                    if (c->tiles[v] == 0 && dist[v] >= 0) best = std::min(best, dist[v] + 1);
                }
            }
            pd[(size_t)src * n + dst] = (uint8_t)std::min(best, M);
        }
    }
}


// the compact-row variant of a compiled parameter block (kernel MODE 3): one env's rows are the A * (F-1) int8
// values of the compact integer format, contiguous, all 32 envs of a warp in one pass, one bulk copy per warp
static inline void make_compact_params(OcParams& p) {
    p.rowf = 0;
    p.row_bytes = p.A * (p.F - 1);
    p.row_stride = p.row_bytes;
    p.nb = 32; p.nb_shift = 5; p.obs_passes = 1; p.nbuf = 1;
    p.grp_shift = 5; p.grp_pad = 0;
    p.buf_bytes = (int)align_up((size_t)32 * p.row_bytes, 16);
    p.warp_row_bytes = p.buf_bytes;
    p.use_tma = 1;
    if (const char* t = getenv("OC_TMA")) p.use_tma = atoi(t) != 0 ? 1 : 0;
}

// returns OC_OK or OC_ERR_INVALID (message in err)
static inline int compile_config(const oc_config* c, HostImage& h, std::string& err) {
#define OC_BAD(msg) do { err = (msg); return OC_ERR_INVALID; } while (0)
    if (!c) OC_BAD("null argument");
    if (c->abi_version != OC_ABI_VERSION) OC_BAD("abi_version mismatch");
    const int W = c->width, H = c->height, n = W * H, A = c->num_agents, S = c->num_subtasks, C = c->num_communication;
    const int M = 2 * (W + H) + 1;
    if (c->num_envs <= 0) OC_BAD("num_envs must be positive");
    if (A < 2 || A > OC_MAX_AGENTS) OC_BAD("num_agents must be 2..4");
    if (W <= 0 || H <= 0 || n > OC_MAX_CELLS || M + W + H > 255) OC_BAD("grid too large");
    if (c->max_path != 0 && c->max_path != M) OC_BAD("max_path must be 2*(w+h)+1");
    if (S <= 0 || S > OC_MAX_SUBTASKS) OC_BAD("num_subtasks out of range");
    if (C <= 0 || C > OC_MAX_COMM) OC_BAD("num_communication out of range");
    if (c->max_num_timesteps < 0 || c->max_num_timesteps > 65534) OC_BAD("max_num_timesteps out of range");
    if (c->max_num_timesteps == 0) OC_BAD("max_num_timesteps == 0 makes the timestep observation t/0 (the reference raises ZeroDivisionError)");
    if (c->num_objects <= 0 || c->num_objects > OC_MAX_OBJECTS) OC_BAD("num_objects out of range");
    if (c->num_items < 1 || c->num_items > 4) OC_BAD("num_items out of range");
    if (!c->tiles) OC_BAD("tiles is null");
    if (c->fow_radius < 0) OC_BAD("fow_radius must be >= 0");
    for (int i = 0; i < n; ++i) if (c->tiles[i] > 3) OC_BAD("bad tile code");
    for (int k = 0; k < A; ++k)
        if (c->start_cell[k] >= n || c->tiles[c->start_cell[k]] != 0) OC_BAD("agent start must be a floor cell");
    // domain: every Food at most once (otherwise the reference itself is ill-defined: it indexes
    // list(set(locations))[0], overcooked_environment.py:288,380) -- see DESIGN.md
    int food_seen = 0;
    for (int s = 0; s < c->num_objects; ++s) {
        const int b = c->object_contents[s];
        if (b != 1 && b != 2 && b != 4 && b != 8) OC_BAD("object_contents must be a single content bit");
        if (b != 8) { if (food_seen & b) OC_BAD("each Food may appear at most once in a level"); food_seen |= b; }
        if (c->object_cell[s] >= 0 && (c->object_cell[s] >= n || c->tiles[c->object_cell[s]] == 0))
            OC_BAD("fixed objects must sit on a non-floor tile");
    }

    OcParams& p = h.p;
    memset(&p, 0, sizeof(p));
    p.E = c->num_envs; p.A = A;
    p.W = W; p.H = H; p.ncell = n; p.T = c->max_num_timesteps; p.C = C; p.S = S;
    p.F = 23 + S + 2 * C; p.fow = c->fow_radius; p.M = M;
    p.row_bytes = A * p.F;
    // byte rows in shared memory: an ODD number of 32-bit words per env makes the per-thread byte
    // scatter bank-conflict free; when row_bytes itself is such a number the rows stay contiguous
    // and the expansion needs no index arithmetic at all
    p.row_stride = (int)align_up(p.row_bytes, 4);
    if (((p.row_stride >> 2) & 1) == 0) p.row_stride += 4;
    // float rows (bulk-stored by the copy engine) when rows are 16-byte multiples and a warp's buffer stays
    // small: all 32 env rows when they fit in 24 KB, else 16 / 8 / 4 rows (<= 30 KB) emitted in 2 / 4 / 8 passes
    p.nb = 32;
    p.rowf = 0;
    if ((p.row_bytes & 3) == 0) {
        if (32 * p.row_bytes * 4 <= 24 * 1024) p.rowf = 1;
        else for (int nb = 16; nb >= 4 && !p.rowf; nb >>= 1)
            if (nb * p.row_bytes * 4 <= 30 * 1024) { p.rowf = 1; p.nb = nb; }
    }
    if (const char* f = getenv("OC_ROW_FORMAT")) {
        if (f[0] != 'f') { p.rowf = 0; p.nb = 32; }
        else if ((p.row_bytes & 3) == 0) p.rowf = 1;
    }
    if (const char* o = getenv("OC_ROW_ENVS")) {
        const int nb = atoi(o);
        if (p.rowf && (nb == 32 || nb == 16 || nb == 8 || nb == 4)) p.nb = nb;
    }
    // float rows stay 16-byte aligned, so the best the per-thread scatter can do is a stride of
    // 4 (mod 8) words (8 distinct banks, 4-way conflict); a multiple of 8 words would put all 32 lanes
    // on 1-4 banks (cfg4: 96 floats -> 32-way).  Such rows are GROUPED: four rows stay contiguous, 16 bytes of
    // padding follow every group -- lanes of one group share a bank (4-way, like everybody else), the eight groups
    // sit on eight different banks, and a group leaves in ONE bulk copy (8 copies per warp and step; padding every
    // single row, OC_ROW_GROUP=1, needs 32 small copies and made the fill phase of a cfg4 step 2.4x cfg2's).
    // (multi-pass rows stay contiguous: one bulk copy per pass measured faster than conflict-free fills)
    bool pad = (p.row_bytes & 7) == 0 && p.nb == 32;
    if (const char* o = getenv("OC_ROW_PAD")) pad = pad && atoi(o) != 0;
    int group = 4;
    if (const char* o = getenv("OC_ROW_GROUP")) { const int g = atoi(o); if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16) group = g; }
    p.grp_shift = 5; p.grp_pad = 0;
    if (p.rowf) {
        p.row_stride = p.row_bytes * 4;
        if (pad && group == 1) p.row_stride += 16;                 // every row padded
        else if (pad) { p.grp_pad = 16; p.grp_shift = group == 2 ? 1 : group == 4 ? 2 : group == 8 ? 3 : 4; }
    }
    // 1: one bulk copy per pass (contiguous rows).  3: one bulk copy per row group.  Rows padded one by one would need
    // one small copy per row: that measured slower in the fused rollout but faster in the single-step kernel, which
    // sets 2 itself.
    p.use_tma = !p.rowf ? 0 : (p.grp_pad != 0 ? 3 : (p.row_stride == p.row_bytes * 4 ? 1 : 0));
    if (const char* t = getenv("OC_TMA")) p.use_tma = (p.use_tma && atoi(t) != 0) ? p.use_tma : 0;
    // A/B knob OC_ROW_BUFS=2: two half-size buffers used alternately, so that a pass only waits for the copy
    // issued two passes ago.  Measured slower everywhere (cfg2 5.1 vs 4.2 us, cfg3 35.4 vs 25.7, cfg5 35.0 vs
    // 33.9): every extra pass re-executes the per-lane fill with half of the lanes idle, which costs more
    // than the overlap gains.  Default: one buffer.
    p.nbuf = 1;
    if (const char* o = getenv("OC_ROW_BUFS"))
        if (atoi(o) == 2 && p.use_tma == 1 && p.nb >= 8) { p.nb /= 2; p.nbuf = 2; }
    p.nb_shift = p.nb == 32 ? 5 : p.nb == 16 ? 4 : p.nb == 8 ? 3 : 2;
    p.obs_passes = 32 / p.nb;
    p.obs_rot = (p.rowf && ((p.row_stride >> 2) & 7) == 4 && (p.F & 3) != 0) ? 1 : 0;
    if (const char* o = getenv("OC_OBS_ROT")) p.obs_rot = p.obs_rot && atoi(o) != 0;
    p.buf_bytes = p.nb * p.row_stride + (p.grp_pad ? (p.nb >> p.grp_shift) * p.grp_pad : 0);
    p.warp_row_bytes = p.nbuf * p.buf_bytes;
    p.r4_magic = (p.row_bytes >= 4) ? (uint32_t)((1ull << 32) / (uint64_t)(p.row_bytes >> 2)) + 1u : 0u;
    p.rf_magic = (uint32_t)((1ull << 32) / (uint64_t)p.row_bytes) + 1u;
    for (int k = 0; k < OC_MAX_AGENTS; ++k) {
        p.can_move[k] = c->can_move[k]; p.allergic[k] = c->allergic[k]; p.blind[k] = c->blind[k];
        p.start_cell[k] = c->start_cell[k];
    }
    p.comm_on = c->communication_on != 0; p.ego_led = c->ego_led != 0; p.ego_blind = c->blind[0] != 0;
    p.seed = c->seed;

    // observation layout: key-sorted order of the reference's Dict space (overcooked_env.py:66-78)
    {
        const int sizes[OC_NUM_OBS_KEYS] = {C, 2, C, 2, 2, S, 4, 4, 4, 4, 1};
        int off = 0;
        for (int i = 0; i < OC_NUM_OBS_KEYS; ++i) { h.obs_off[i] = off; h.obs_size[i] = sizes[i]; off += sizes[i]; }
        p.off_a1comm = h.obs_off[0]; p.off_a1loc = h.obs_off[1]; p.off_a2comm = h.obs_off[2];
        p.off_a2loc = h.obs_off[3]; p.off_hold = h.obs_off[4]; p.off_completed = h.obs_off[5];
        p.off_hidden = h.obs_off[6]; p.off_encx = h.obs_off[7]; p.off_ency = h.obs_off[8];
        p.off_state = h.obs_off[9]; p.off_ts = h.obs_off[10];
    }

    // delivery / counters
    std::vector<uint8_t> counters; int delivery0 = -1; std::vector<int> deliveries;
    for (int i = 0; i < n; ++i) {
        if (c->tiles[i] == 1) counters.push_back((uint8_t)i);
        if (c->tiles[i] == 3) { deliveries.push_back(i); if (delivery0 < 0) delivery0 = i; }
    }
    if (delivery0 < 0) OC_BAD("level has no Delivery tile");
    p.delivery0 = (uint8_t)delivery0;
    p.ncounters = (int)counters.size();

    // reset image: objects in world insertion order; stamp = insertion index + 1; key ranks in
    // first-insert order (world.py:236-237)
    uint64_t ranks = 0; uint32_t nkeys = 0; p.nrandom = 0;
    for (int s = 0; s < OCK_MAX_OBJECTS; ++s) p.init_obj[s] = OCK_DEAD;
    for (int s = 0; s < c->num_objects; ++s) {
        const uint32_t b = c->object_contents[s];
        uint32_t cell = 0;
        if (c->object_cell[s] < 0) p.random_slot[p.nrandom++] = (uint8_t)s; else cell = (uint32_t)c->object_cell[s];
        p.init_obj[s] = b | (OCK_HOLDER_NONE << 8) | (cell << 16) | ((uint32_t)(s + 1) << 24);
        if (((ranks >> (4 * b)) & 15) == 0) { ++nkeys; ranks |= (uint64_t)nkeys << (4 * b); }
    }
    if (p.nrandom > p.ncounters) OC_BAD("more random objects than Counter tiles");
    p.init_ranks = ranks;
    p.init_w0 = ((uint32_t)c->num_objects << 16) | (nkeys << 24);

    // subtask tables
    std::vector<uint32_t> tmlut(128, 0);
    int nchop = 0;
    for (int i = 0; i < S; ++i) {
        const int kind = c->subtask_kind[i], goal = c->subtask_goal[i];
        if (kind < 0 || kind > 2 || goal <= 0 || goal > 127 || (goal & 7) == 0) OC_BAD("bad subtask entry");
        tmlut[goal] |= 1u << i;
        if (kind == 2) {
            p.deliver_mask |= 1u << i;
            if (p.ndeliver >= OCK_MAX_DELIVER) OC_BAD("too many Deliver subtasks");
            p.deliver_sig[p.ndeliver] = (uint8_t)goal; p.deliver_idx[p.ndeliver] = (uint8_t)i; ++p.ndeliver;
        } else {
            p.nondeliver_mask |= 1u << i;
            if (kind == 0) {
                const int f = c->subtask_arg0[i];
                if (f != 1 && f != 2 && f != 4) OC_BAD("Chop subtask needs a single food bit");
                p.chop_mask[f == 1 ? 0 : (f == 2 ? 1 : 2)] |= 1u << i;
                ++nchop;
            }
        }
    }
    if (p.ndeliver == 0) OC_BAD("no delivery subtask (overcooked_environment.py:251 asserts)");
    // items = ['Plate'] + Foods of recipes[0] (overcooked_environment.py:319-321)
    if (c->items[0] != 8) OC_BAD("items[0] must be Plate");
    p.item_foods = 0;
    for (int i = 1; i < c->num_items; ++i) {
        const int b = c->items[i];
        if ((b != 1 && b != 2 && b != 4) || (p.item_foods & b)) OC_BAD("items[1..] must be distinct single Food bits");
        p.item_foods |= (uint32_t)b;
    }
    p.npairs = c->num_items * (c->num_items - 1) / 2;
    // kernel shape (template instantiation): object slots and food channels the device loops run over.
    // A food can only matter if some object, subtask goal, Chop subtask or recipe item names it.
    {
        uint32_t foods = p.item_foods;
        for (int s = 0; s < c->num_objects; ++s) foods |= c->object_contents[s] & 7u;
        for (int i = 0; i < S; ++i) foods |= (c->subtask_goal[i] & 7u) | (c->subtask_kind[i] == 0 ? (c->subtask_arg0[i] & 7u) : 0u);
        const int nf = (foods & 4u) ? 3 : ((foods & 2u) ? 2 : 1);
        if (c->num_objects <= 2 && nf == 1) { p.NOBJ = 2; p.NF = 1; }
        else if (c->num_objects <= 4 && nf <= 2) { p.NOBJ = 4; p.NF = 2; }
        else { p.NOBJ = 6; p.NF = 3; }
        if (const char* sh = getenv("OC_KERNEL_SHAPE")) {       // A/B knob: force a larger shape
            const int v = atoi(sh);
            if (v >= 6) { p.NOBJ = 6; p.NF = 3; } else if (v >= 4 && p.NOBJ <= 4) { p.NOBJ = 4; p.NF = 2; }
        }
    }

    // tables
    std::vector<uint8_t> pd;
    if (c->path_dist) pd.assign(c->path_dist, c->path_dist + (size_t)n * n); else bfs_path_dist(c, pd);
    std::vector<uint16_t> mvt(n * 4), xy16(n);
    std::vector<float> xyf((size_t)2 * n, 0.0f);
    std::vector<uint8_t> dmin(n), pdm((size_t)n * n);
    static const int dx[4] = {0, 0, -1, 1}, dy[4] = {1, -1, 0, 0};
    for (int i = 0; i < n; ++i) {
        const int x = i % W, y = i / W;
        xy16[i] = (uint16_t)(x | (y << 8));
        xyf[2 * i] = (float)x; xyf[2 * i + 1] = (float)y;
        for (int a = 0; a < 4; ++a) {
            const int vx = std::min(std::max(x + dx[a], 0), W - 1), vy = std::min(std::max(y + dy[a], 0), H - 1);
            const int v = vy * W + vx;
            mvt[4 * i + a] = (uint16_t)(v | (c->tiles[v] << 8));
        }
        int best = 1 << 20;
        for (int j = 0; j < n; ++j) pdm[(size_t)i * n + j] = (uint8_t)((int)pd[(size_t)i * n + j] + abs(x - j % W) + abs(y - j / W));
        for (int d : deliveries) best = std::min(best, (int)pdm[(size_t)i * n + d]);
        dmin[i] = (uint8_t)best;
    }
    const int nq = std::max(std::max(2 * M + std::max(nchop - 1, 0) * 2 * M, p.npairs * M), M + W + H) + 1;
    std::vector<double> q(nq);
    for (int i = 0; i < nq; ++i) q[i] = (double)i / (double)M;      // == Python int / int (correctly rounded)
    std::vector<float>& ts = h.ts; ts.assign(p.T + 1, 0.f);
    for (int t = 0; t <= p.T; ++t) ts[t] = (float)((double)t / (double)p.T);   // np.float32(t / T)

    size_t off = 0;
    p.o_q = (int)off; off += align_up(q.size() * 8, 16);
    p.o_tmlut = (int)off; off += 128 * 4;
    p.o_xyf = (int)off; off += align_up((size_t)n * 8, 16);
    p.o_mvt = (int)off; off += align_up((size_t)n * 8, 16);
    p.o_xy16 = (int)off; off += align_up((size_t)n * 2, 16);
    p.o_dmin = (int)off; off += align_up(n, 16);
    p.o_counters = (int)off; off += align_up(std::max<size_t>(counters.size(), 1), 16);
    p.o_pd = (int)off; off += align_up((size_t)n * n, 16);
    p.o_pdm = (int)off; off += align_up((size_t)n * n, 16);
    p.o_ts = -1;
    if (ts.size() * 4 <= 4096) { p.o_ts = (int)off; off += align_up(ts.size() * 4, 16); }
    p.blob_bytes = (int)off;
    std::vector<uint8_t>& blob = h.blob; blob.assign(off, 0);
    memcpy(blob.data() + p.o_q, q.data(), q.size() * 8);
    memcpy(blob.data() + p.o_tmlut, tmlut.data(), 128 * 4);
    memcpy(blob.data() + p.o_xyf, xyf.data(), xyf.size() * 4);
    memcpy(blob.data() + p.o_mvt, mvt.data(), mvt.size() * 2);
    memcpy(blob.data() + p.o_xy16, xy16.data(), xy16.size() * 2);
    memcpy(blob.data() + p.o_dmin, dmin.data(), dmin.size());
    if (!counters.empty()) memcpy(blob.data() + p.o_counters, counters.data(), counters.size());
    memcpy(blob.data() + p.o_pd, pd.data(), pd.size());
    memcpy(blob.data() + p.o_pdm, pdm.data(), pdm.size());
    if (p.o_ts >= 0) memcpy(blob.data() + p.o_ts, ts.data(), ts.size() * 4);
    return OC_OK;
#undef OC_BAD
}

}  // namespace ock
