// TEST INFRASTRUCTURE ONLY -- CPU emulation of the CUDA kernels' control flow.
//
// Executes the device functions of gym_comm_b200/csrc/oc_device.cuh (compiled for the host via
// oc_emu_shim.h) warp by warp, lane by lane, behind the same C ABI names with an `emu_` prefix.
// Purpose: debug the packed-state logic against the oracle in a container that has no GPU,
// before spending GPU minutes.  It is NOT a fallback: the product package never loads it and
// `oc_create` fails without a CUDA device.  Host-memory pointers everywhere.
#define OCK_HOST_EMU 1
#include "../../gym_comm_b200/csrc/oc_device.cuh"
#include "../../gym_comm_b200/csrc/oc_host.hpp"

#include <string>
#include <utility>
#include <vector>

using namespace ock;

struct emu_env {
    OcParams p;
    std::vector<uint8_t> blob;
    std::vector<float> ts;
    std::vector<uint4> state;
    uint32_t rollout_step = 0;
    int obs_off[OC_NUM_OBS_KEYS], obs_size[OC_NUM_OBS_KEYS];
};
static std::string g_err;

template <int N, int NFOOD> constexpr int shape_nobj(std::integer_sequence<int, N, NFOOD>) { return N; }
template <int N, int NFOOD> constexpr int shape_nf(std::integer_sequence<int, N, NFOOD>) { return NFOOD; }

template <typename F>
static int dispatch(int A, int NOBJ, int rowf, F&& f) {
#define OC_CASE(a, n, nfood)                                                                                       \
    if (A == a && NOBJ == n) {                                                                                     \
        using IS = std::integer_sequence<int, n, nfood>;                                                           \
        if (rowf) return f(std::integral_constant<int, a>(), IS(), std::true_type());                              \
        return f(std::integral_constant<int, a>(), IS(), std::false_type());                                       \
    }
    OC_CASE(2, 2, 1) OC_CASE(3, 2, 1) OC_CASE(4, 2, 1) OC_CASE(2, 4, 2) OC_CASE(3, 4, 2) OC_CASE(4, 4, 2)
    OC_CASE(2, 6, 3) OC_CASE(3, 6, 3) OC_CASE(4, 6, 3)
#undef OC_CASE
    return OC_ERR_INVALID;
}

// one warp at a time: every lane runs the logic (body; returns true when its env just finished
// and wants the auto-reset), then terminal observations, auto-reset and the observation passes run
// the way the kernels order them
template <int A, int NOBJ, int NF, bool ROWF, typename Body>
static void for_each_warp(emu_env* h, float* obs, float* term_obs, Body&& body) {
    const OcParams& p = h->p;
    const Tables tb = make_tables(p, h->blob.data());
    std::vector<uint8_t> rows((size_t)p.warp_row_bytes);
    Env<A, NOBJ> we[32];
    Info win[32];
    for (int env0 = 0; env0 < p.E; env0 += 32) {
        const int nvalid = std::min(32, p.E - env0);
        for (int lane = 0; lane < 32; ++lane) warp_clear_rows<ROWF>(rows.data(), p.warp_row_bytes, lane);
        for (int lane = 0; lane < nvalid; ++lane) {
            const int env = env0 + lane;
            const bool fin = body(tb, env, we[lane], win[lane]);
            if (fin) {
                if (term_obs)        // warp_terminal_obs: rows clear, finished lanes emit one by one
                    thread_emit_rows<A, NOBJ, NF, ROWF>(we[lane], p, tb, win[lane],
                                                        rows.data() + row_offset(p, lane & (p.nb - 1)),
                                                        term_obs + (size_t)env * p.row_bytes);
                finish_episode<A, NOBJ>(we[lane], p, tb, (uint32_t)env);
                win[lane] = gather_info<A, NOBJ, NF>(we[lane], p, tb);
            }
            store_env<A, NOBJ>(we[lane], h->state.data(), p.E, env);
        }
        if (!obs) continue;
        float* out0 = obs + (size_t)env0 * p.row_bytes;
        for (int pass = 0; pass < p.obs_passes; ++pass) {          // emit_obs, lane by lane
            uint8_t* buf = rows.data() + (size_t)(pass & (p.nbuf - 1)) * p.buf_bytes;
            for (int lane = 0; lane < 32; ++lane) warp_clear_rows<ROWF>(buf, p.buf_bytes, lane);
            for (int lane = 0; lane < nvalid; ++lane)
                if ((lane >> p.nb_shift) == pass)
                    fill_rows<A, NOBJ, NF, ROWF, true>(we[lane], p, tb, win[lane], timestep_of<A, NOBJ>(we[lane], p, tb),
                                                 buf + row_offset(p, lane & (p.nb - 1)),
                                                 (ROWF && p.obs_passes == 1 && p.grp_pad == 0 && p.obs_rot) ? ((lane >> 3) % A) : 0);
            const int first = pass << p.nb_shift;
            const int nv = std::min(p.nb, nvalid - first);
            if (nv > 0)
                for (int lane = 0; lane < 32; ++lane)
                    warp_expand_rows<ROWF>(p, buf, out0 + (size_t)first * p.row_bytes, nv, lane);
        }
        if (!ROWF)
            for (int lane = 0; lane < nvalid; ++lane)
                store_timesteps<A>(p, out0 + (size_t)lane * p.row_bytes, timestep_of<A, NOBJ>(we[lane], p, tb));
    }
}

// the same for the compact-row kernels (MODE 3): rows = int8 [A, F-1] per env, clock per env
template <int A, int NOBJ, int NF, typename Body>
static void for_each_warp_packed(emu_env* h, uint8_t* obs8, float* ts, uint8_t* term8, float* term_ts, Body&& body) {
    OcParams p = h->p;
    make_compact_params(p);
    const Tables tb = make_tables(p, h->blob.data());
    std::vector<uint8_t> rows((size_t)p.warp_row_bytes);
    Env<A, NOBJ> we[32];
    Info win[32];
    for (int env0 = 0; env0 < p.E; env0 += 32) {
        const int nvalid = std::min(32, p.E - env0);
        for (int lane = 0; lane < 32; ++lane) warp_clear_rows<true>(rows.data(), p.warp_row_bytes, lane);
        for (int lane = 0; lane < nvalid; ++lane) {
            const int env = env0 + lane;
            const bool fin = body(tb, env, we[lane], win[lane]);
            if (fin) {
                if (term8) {          // warp_terminal_obs_packed, one finished lane at a time (rows clear before and after)
                    uint8_t* myrow = rows.data() + (size_t)lane * p.row_stride;
                    build_rows_i8<A, NOBJ, NF>(we[lane], p, tb, win[lane], myrow);
                    for (int j = 0; j < p.row_bytes; ++j) { term8[(size_t)env * p.row_bytes + j] = myrow[j]; myrow[j] = 0; }
                    if (term_ts) term_ts[env] = timestep_of<A, NOBJ>(we[lane], p, tb);
                }
                finish_episode<A, NOBJ>(we[lane], p, tb, (uint32_t)env);
                win[lane] = gather_info<A, NOBJ, NF>(we[lane], p, tb);
            }
            store_env<A, NOBJ>(we[lane], h->state.data(), p.E, env);
        }
        if (!obs8) continue;
        // emit_obs_packed, lane by lane: fill, then the warp's copy, then the clocks
        for (int lane = 0; lane < nvalid; ++lane)
            build_rows_i8<A, NOBJ, NF>(we[lane], p, tb, win[lane], rows.data() + (size_t)lane * p.row_stride);
        for (int lane = 0; lane < 32; ++lane)
            warp_store_packed(p, rows.data(), obs8 + (size_t)env0 * p.row_bytes, nvalid, lane);
        if (ts)
            for (int lane = 0; lane < nvalid; ++lane) ts[env0 + lane] = timestep_of<A, NOBJ>(we[lane], p, tb);
    }
}

template <typename F>
static int dispatch_shape(int A, int NOBJ, F&& f) {
#define OC_CASE(a, n, nfood)                                                                                       \
    if (A == a && NOBJ == n) return f(std::integral_constant<int, a>(), std::integer_sequence<int, n, nfood>());
    OC_CASE(2, 2, 1) OC_CASE(3, 2, 1) OC_CASE(4, 2, 1) OC_CASE(2, 4, 2) OC_CASE(3, 4, 2) OC_CASE(4, 4, 2)
    OC_CASE(2, 6, 3) OC_CASE(3, 6, 3) OC_CASE(4, 6, 3)
#undef OC_CASE
    return OC_ERR_INVALID;
}

// oc_rollout_kernel with single-pass float rows (kernel MODE 1), warp by warp: the warp's rows are cleared and filled in
// the first step only; every later step takes back / overwrites the previous observation (build_rows_f32<UNDO>) exactly
// as emit_obs_undo does on the device.  (Envs are independent, so walking the steps warp-major gives the same results as
// the device's step-major order.)
template <int A, int NOBJ, int NF>
static void rollout_undo(emu_env* h, int32_t n_steps, float* obs, float* rew32, uint8_t* done, int32_t* actions_out) {
    const OcParams& p = h->p;
    const Tables tb = make_tables(p, h->blob.data());
    const size_t step_floats = (size_t)p.E * p.row_bytes;
    std::vector<uint8_t> rows((size_t)p.warp_row_bytes);
    Env<A, NOBJ> we[32];
    Info win[32];
    uint32_t shown_comm[32], shown_completed[32];
    for (int env0 = 0; env0 < p.E; env0 += 32) {
        const int nvalid = std::min(32, p.E - env0);
        for (int lane = 0; lane < nvalid; ++lane) load_env<A, NOBJ>(we[lane], h->state.data(), p.E, env0 + lane);
        for (int s = 0; s < n_steps; ++s) {
            for (int lane = 0; lane < nvalid; ++lane)
                win[lane] = rollout_logic<A, NOBJ, NF, true>(we[lane], p, tb, (uint32_t)(env0 + lane), (uint32_t)s, h->rollout_step,
                                                             rew32, done, actions_out);
            if (!obs) continue;
            if (s == 0)
                for (int lane = 0; lane < 32; ++lane) warp_clear_rows<true>(rows.data(), p.warp_row_bytes, lane);
            for (int lane = 0; lane < nvalid; ++lane) {
                float* myrow = reinterpret_cast<float*>(rows.data() + row_offset(p, lane));
                const float ts = timestep_of<A, NOBJ>(we[lane], p, tb);
                if (s == 0) build_rows_f32<A, NOBJ, NF, false, false>(we[lane], p, tb, win[lane], ts, myrow);
                else build_rows_f32<A, NOBJ, NF, false, true>(we[lane], p, tb, win[lane], ts, myrow, 0, shown_comm[lane], shown_completed[lane]);
                shown_comm[lane] = we[lane].comm; shown_completed[lane] = we[lane].completed;
            }
            for (int lane = 0; lane < 32; ++lane)
                warp_expand_rows<true>(p, rows.data(), obs + (size_t)s * step_floats + (size_t)env0 * p.row_bytes, nvalid, lane);
        }
        for (int lane = 0; lane < nvalid; ++lane) store_env<A, NOBJ>(we[lane], h->state.data(), p.E, env0 + lane);
    }
}

extern "C" {

const char* emu_last_error(void) { return g_err.c_str(); }

int emu_create(const oc_config* c, emu_env** out) {
    emu_env* h = new emu_env();
    HostImage img;
    if (compile_config(c, img, g_err) != OC_OK) { delete h; return OC_ERR_INVALID; }
    h->p = img.p; h->blob = img.blob; h->ts = img.ts;
    memcpy(h->obs_off, img.obs_off, sizeof(h->obs_off));
    memcpy(h->obs_size, img.obs_size, sizeof(h->obs_size));
    h->p.blob = h->blob.data(); h->p.ts_table = h->ts.data();
    h->state.assign((size_t)h->p.E * 4, uint4{0, 0, 0, 0});
    dispatch(h->p.A, h->p.NOBJ, h->p.rowf, [&](auto a, auto n, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(n), FF = shape_nf(n);
        constexpr bool RF = decltype(rf)::value;
        for_each_warp<AA, NN, FF, RF>(h, nullptr, nullptr, [&](const Tables& tb, int env, Env<AA, NN>& e, Info&) {
            reset_logic<AA, NN>(e, h->p, tb, (uint32_t)env, true, nullptr, nullptr);
            return false;
        });
        return 0;
    });
    *out = h;
    return OC_OK;
}
int emu_destroy(emu_env* h) { delete h; return OC_OK; }
int emu_obs_width(const emu_env* h) { return h->p.F; }
int emu_obs_layout(const emu_env* h, int32_t* off, int32_t* sz) {
    for (int i = 0; i < OC_NUM_OBS_KEYS; ++i) { off[i] = h->obs_off[i]; sz[i] = h->obs_size[i]; }
    return OC_OK;
}

int emu_reset(emu_env* h, const uint8_t* mask, const int32_t* placements, float* obs, void*) {
    return dispatch(h->p.A, h->p.NOBJ, h->p.rowf, [&](auto a, auto n, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(n), FF = shape_nf(n);
        constexpr bool RF = decltype(rf)::value;
        for_each_warp<AA, NN, FF, RF>(h, obs, nullptr, [&](const Tables& tb, int env, Env<AA, NN>& e, Info& in) {
            load_env<AA, NN>(e, h->state.data(), h->p.E, env);
            reset_logic<AA, NN>(e, h->p, tb, (uint32_t)env, false, mask, placements);
            if (obs) in = gather_info<AA, NN, FF>(e, h->p, tb);
            return false;
        });
        return OC_OK;
    });
}

int emu_step(emu_env* h, const int32_t* actions, float* obs, float* rew32, double* rew64, uint8_t* done,
             float* term_obs, uint32_t flags, void*) {
    return dispatch(h->p.A, h->p.NOBJ, h->p.rowf, [&](auto a, auto n, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(n), FF = shape_nf(n);
        constexpr bool RF = decltype(rf)::value;
        for_each_warp<AA, NN, FF, RF>(h, obs, term_obs, [&](const Tables& tb, int env, Env<AA, NN>& e, Info& in) {
            load_env<AA, NN>(e, h->state.data(), h->p.E, env);
            int nav[AA], comm[AA];
            for (int k = 0; k < AA; ++k) { nav[k] = actions[((size_t)env * AA + k) * 2] & 3; comm[k] = actions[((size_t)env * AA + k) * 2 + 1]; }
            bool fin;
            in = step_logic<AA, NN, FF>(e, h->p, tb, nav, comm[0], comm[1], (uint32_t)env, rew32, rew64, done, fin);
            return fin && (flags & OC_FLAG_AUTO_RESET);
        });
        return OC_OK;
    });
}

// oc_reset_i8 / oc_step_i8: the compact-row kernels (MODE 3), incl. u8 actions and the per-env reward
int emu_reset_i8(emu_env* h, const uint8_t* mask, const int32_t* placements, int8_t* obs8, float* ts, void*) {
    return dispatch_shape(h->p.A, h->p.NOBJ, [&](auto a, auto n) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(n), FF = shape_nf(n);
        for_each_warp_packed<AA, NN, FF>(h, (uint8_t*)obs8, ts, nullptr, nullptr, [&](const Tables& tb, int env, Env<AA, NN>& e, Info& in) {
            load_env<AA, NN>(e, h->state.data(), h->p.E, env);
            reset_logic<AA, NN>(e, h->p, tb, (uint32_t)env, false, mask, placements);
            if (obs8) in = gather_info<AA, NN, FF>(e, h->p, tb);
            return false;
        });
        return OC_OK;
    });
}

int emu_step_i8(emu_env* h, const void* actions, int8_t* obs8, float* ts, float* rew32, double* rew64, uint8_t* done,
                int8_t* term8, float* term_ts, uint32_t flags, void*) {
    return dispatch_shape(h->p.A, h->p.NOBJ, [&](auto a, auto n) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(n), FF = shape_nf(n);
        for_each_warp_packed<AA, NN, FF>(h, (uint8_t*)obs8, ts, (uint8_t*)term8, term_ts,
                                         [&](const Tables& tb, int env, Env<AA, NN>& e, Info& in) {
            load_env<AA, NN>(e, h->state.data(), h->p.E, env);
            int nav[AA], comm[AA];
            for (int k = 0; k < AA; ++k) {
                if (flags & OC_FLAG_ACTIONS_U8) {
                    const uint8_t* a8 = (const uint8_t*)actions + ((size_t)env * AA + k) * 2;
                    nav[k] = a8[0] & 3; comm[k] = a8[1];
                } else {
                    const int32_t* a32 = (const int32_t*)actions + ((size_t)env * AA + k) * 2;
                    nav[k] = a32[0] & 3; comm[k] = a32[1];
                }
            }
            bool fin;
            in = step_logic<AA, NN, FF>(e, h->p, tb, nav, comm[0], comm[1], (uint32_t)env, rew32, rew64, done, fin,
                                        (flags & OC_FLAG_REWARD_PER_ENV) != 0);
            return fin && (flags & OC_FLAG_AUTO_RESET);
        });
        return OC_OK;
    });
}

int emu_rollout(emu_env* h, int32_t n_steps, float* obs, float* rew32, uint8_t* done, int32_t* actions_out, void*) {
    const size_t step_floats = (size_t)h->p.E * h->p.row_bytes;
    int rc = dispatch(h->p.A, h->p.NOBJ, h->p.rowf, [&](auto a, auto n, auto rf) -> int {
        constexpr int AA = decltype(a)::value, NN = shape_nobj(n), FF = shape_nf(n);
        constexpr bool RF = decltype(rf)::value;
        if (RF && h->p.obs_passes == 1) {                        // kernel MODE 1: rows keep the previous observation
            rollout_undo<AA, NN, FF>(h, n_steps, obs, rew32, done, actions_out);
            return OC_OK;
        }
        for (int s = 0; s < n_steps; ++s)
            for_each_warp<AA, NN, FF, RF>(h, obs ? obs + (size_t)s * step_floats : nullptr, nullptr,
                                  [&](const Tables& tb, int env, Env<AA, NN>& e, Info& in) {
                load_env<AA, NN>(e, h->state.data(), h->p.E, env);
                in = rollout_logic<AA, NN, FF, RF>(e, h->p, tb, (uint32_t)env, (uint32_t)s, h->rollout_step, rew32, done, actions_out);
                return false;
            });
        return OC_OK;
    });
    h->rollout_step += (uint32_t)n_steps;
    return rc;
}

int emu_get_state(emu_env* h, uint32_t* state, void*) {
    for (int env = 0; env < h->p.E; ++env)
        for (int pl = 0; pl < 4; ++pl) reinterpret_cast<uint4*>(state)[(size_t)env * 4 + pl] = h->state[(size_t)pl * h->p.E + env];
    return OC_OK;
}
int emu_set_state(emu_env* h, const uint32_t* state, void*) {
    for (int env = 0; env < h->p.E; ++env)
        for (int pl = 0; pl < 4; ++pl)            // oc_state_import_kernel: sanitised on the way in
            h->state[(size_t)pl * h->p.E + env] = sanitize_state_plane(h->p, pl, reinterpret_cast<const uint4*>(state)[(size_t)env * 4 + pl]);
    return OC_OK;
}
int emu_pack_obs_i8(emu_env* h, const float* obs, int8_t* obs_i8, float* timestep, void*) {   // oc_pack_i8_kernel, thread by thread
    const OcParams& p = h->p;
    const uint64_t nbytes = (uint64_t)p.E * p.A * (p.F - 1);
    const uint32_t nwords = (uint32_t)std::max<uint64_t>((nbytes + 3) / 4, (uint64_t)p.E);
    for (uint32_t w = 0; w < nwords; ++w) pack_i8_word(p, obs, obs_i8, timestep, w);
    return OC_OK;
}
int emu_gather_term(emu_env* h, const float* term, const int32_t* idx, int n, float* out_f32, int8_t* out_i8, float* out_ts) {
    for (int i = 0; i < n; ++i)                       // oc_gather_term_kernel: one 128-thread CTA per finished env
        for (int tid = 0; tid < 128; ++tid) gather_term_row(h->p, term, idx[i], i, out_f32, out_i8, out_ts, tid, 128);
    return OC_OK;
}
int emu_get_stats(emu_env* h, uint32_t* episodes, uint32_t* last_completed, void*) {
    for (int i = 0; i < h->p.E; ++i) {
        if (episodes) episodes[i] = h->state[i].y;
        if (last_completed) last_completed[i] = h->state[(size_t)h->p.E + i].y & 0xFFu;
    }
    return OC_OK;
}
}  // extern "C"
