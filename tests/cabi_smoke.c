/* The C ABI used from plain C (no Python, no torch): builds the 5x4 kitchen below by hand, steps 64
 * envs for 200 random steps through oc_step / oc_rollout / oc_reset and checks structural facts of
 * the outputs, then checks the host-buffer entry points (oc_step_host) against the device-pointer path.  Compiled and run by tests/test_gpu_cabi_c.py (nvcc only links the CUDA runtime).
 *
 *      - - * - -          tiles: 1 Counter, 3 Delivery, 2 Cutboard, 0 Floor
 *      t       p          objects: Tomato on (0,1), Plate on (4,1)
 *      /       -
 *      - - - - -
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/overcooked_b200.h"

#define CHECK(x) do { int _r = (x); if (_r != 0) { fprintf(stderr, "%s failed (%d): %s\n", #x, _r, oc_last_error()); return 1; } } while (0)
#define CU(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(_e)); return 1; } } while (0)

int main(void) {
    enum { W = 5, H = 4, E = 64, A = 2, C = 4, S = 3, F = 23 + S + 2 * C };
    static const uint8_t tiles[W * H] = {1, 1, 3, 1, 1,
                                         1, 0, 0, 0, 1,
                                         2, 0, 0, 0, 1,
                                         1, 1, 1, 1, 1};
    oc_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = OC_ABI_VERSION;
    cfg.num_envs = E; cfg.num_agents = A; cfg.width = W; cfg.height = H;
    cfg.max_num_timesteps = 50; cfg.num_communication = C; cfg.communication_on = 1; cfg.fow_radius = 1;
    for (int k = 0; k < A; ++k) { cfg.can_move[k] = 1; cfg.start_cell[k] = (uint8_t)(1 * W + 1 + 2 * k); }
    cfg.tiles = tiles;
    cfg.path_dist = NULL;                       /* let the library tabulate get_path_distance_between */
    cfg.num_objects = 2;
    cfg.object_contents[0] = 1; cfg.object_cell[0] = 1 * W + 0;      /* Tomato */
    cfg.object_contents[1] = 8; cfg.object_cell[1] = 1 * W + 4;      /* Plate  */
    cfg.num_subtasks = S;                       /* SimpleTomato: Deliver(Plate-Tomato), Merge(Tomato, Plate), Chop(Tomato) */
    cfg.subtask_kind[0] = 2; cfg.subtask_goal[0] = 9 | (1 << 4);
    cfg.subtask_kind[1] = 1; cfg.subtask_goal[1] = 9 | (1 << 4);
    cfg.subtask_kind[2] = 0; cfg.subtask_goal[2] = 1 | (1 << 4); cfg.subtask_arg0[2] = 1;
    cfg.num_items = 2; cfg.items[0] = 8; cfg.items[1] = 1;
    cfg.seed = 7;

    oc_env* env = NULL;
    CHECK(oc_create(&cfg, &env));
    if (oc_obs_width(env) != F) { fprintf(stderr, "obs width %d != %d\n", oc_obs_width(env), F); return 1; }
    int32_t off[OC_NUM_OBS_KEYS], size[OC_NUM_OBS_KEYS];
    CHECK(oc_obs_layout(env, off, size));

    int32_t* d_act; float *d_obs, *d_rew; double* d_rew64; uint8_t* d_done;
    CU(cudaMalloc((void**)&d_act, E * A * 2 * sizeof(int32_t)));
    CU(cudaMalloc((void**)&d_obs, (size_t)8 * E * A * F * sizeof(float)));
    CU(cudaMalloc((void**)&d_rew, (size_t)8 * E * A * sizeof(float)));
    CU(cudaMalloc((void**)&d_rew64, E * sizeof(double)));
    CU(cudaMalloc((void**)&d_done, (size_t)8 * E));
    static int32_t h_act[E * A * 2];
    static float h_obs[E * A * F];
    static uint8_t h_done[E];
    static double h_rew[E];
    CHECK(oc_reset(env, NULL, NULL, d_obs, NULL));
    long dones = 0;
    srand(1);
    for (int t = 0; t < 200; ++t) {
        for (int i = 0; i < E * A; ++i) { h_act[2 * i] = rand() % 4; h_act[2 * i + 1] = rand() % C; }
        CU(cudaMemcpy(d_act, h_act, sizeof(h_act), cudaMemcpyHostToDevice));
        CHECK(oc_step(env, d_act, d_obs, d_rew, d_rew64, d_done, NULL, OC_FLAG_AUTO_RESET, NULL));
        CU(cudaMemcpy(h_obs, d_obs, sizeof(h_obs), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(h_done, d_done, sizeof(h_done), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(h_rew, d_rew64, sizeof(h_rew), cudaMemcpyDeviceToHost));
        for (int e = 0; e < E; ++e) {
            dones += h_done[e];
            if (!(h_rew[e] <= 5.0 && h_rew[e] >= -20.0)) { fprintf(stderr, "reward out of range %f\n", h_rew[e]); return 1; }
            for (int k = 0; k < A; ++k) {
                const float* row = h_obs + ((size_t)e * A + k) * F;
                float s1 = 0, s2 = 0;
                for (int j = 0; j < C; ++j) { s1 += row[off[OC_OBS_AGENT1_COMM] + j]; s2 += row[off[OC_OBS_AGENT2_COMM] + j]; }
                if (s1 != 1.0f || s2 != 1.0f) { fprintf(stderr, "message one-hot broken\n"); return 1; }
                const float ts = row[off[OC_OBS_TIMESTEP]];
                if (ts < 0.0f || ts > 1.0f) { fprintf(stderr, "timestep %f\n", ts); return 1; }
                const float x = row[off[OC_OBS_AGENT1_LOCATION]], y = row[off[OC_OBS_AGENT1_LOCATION] + 1];
                if (x < 1 || x > 3 || y < 1 || y > 2) { fprintf(stderr, "agent outside the floor (%f, %f)\n", x, y); return 1; }
            }
        }
    }
    if (dones != (long)E * 4) { fprintf(stderr, "expected every env to hit the 50-step limit 4 times, got %ld dones\n", dones); return 1; }
    CHECK(oc_rollout(env, 8, d_obs, d_rew, d_done, NULL, NULL));
    CU(cudaDeviceSynchronize());
    if (oc_launch_count(env) < 203) { fprintf(stderr, "launch count %llu\n", (unsigned long long)oc_launch_count(env)); return 1; }
    /* host-buffer entry points: same env continued through oc_step_host with pinned buffers from
     * oc_host_alloc must agree with the device-pointer path run on a second handle */
    {
        cfg.seed = 11;
        oc_env *ha = NULL, *hb = NULL, *hc = NULL;     /* hc: the compact integer format (oc_step_host_i8) */
        CHECK(oc_create(&cfg, &ha));
        CHECK(oc_create(&cfg, &hb));
        CHECK(oc_create(&cfg, &hc));
        int8_t *p_o8, *p_t8; float *p_ts, *p_tts; uint8_t* p_done8;
        CHECK(oc_host_alloc((size_t)E * A * (F - 1), (void**)&p_o8));
        CHECK(oc_host_alloc((size_t)E * A * (F - 1), (void**)&p_t8));
        CHECK(oc_host_alloc(E * sizeof(float), (void**)&p_ts));
        CHECK(oc_host_alloc(E * sizeof(float), (void**)&p_tts));
        CHECK(oc_host_alloc(E, (void**)&p_done8));
        CHECK(oc_reset_host_i8(hc, NULL, NULL, p_o8, p_ts, NULL));
        int32_t* p_act; float *p_obs, *p_rew, *p_term; uint8_t* p_done;
        CHECK(oc_host_alloc(sizeof(h_act), (void**)&p_act));
        CHECK(oc_host_alloc(sizeof(h_obs), (void**)&p_obs));
        CHECK(oc_host_alloc(sizeof(h_obs), (void**)&p_term));
        CHECK(oc_host_alloc(E * A * sizeof(float), (void**)&p_rew));
        CHECK(oc_host_alloc(E, (void**)&p_done));
        memset(p_term, 0, sizeof(h_obs));
        CHECK(oc_reset_host(ha, NULL, NULL, p_obs, NULL));
        CHECK(oc_reset(hb, NULL, NULL, d_obs, NULL));
        long term_rows = 0;
        for (int t = 0; t < 120; ++t) {
            for (int i = 0; i < E * A; ++i) { p_act[2 * i] = rand() % 4; p_act[2 * i + 1] = rand() % C; }
            CHECK(oc_step_host(ha, p_act, p_obs, p_rew, NULL, p_done, p_term, OC_FLAG_AUTO_RESET, NULL));
            CU(cudaMemcpy(d_act, p_act, sizeof(h_act), cudaMemcpyHostToDevice));
            CHECK(oc_step(hb, d_act, d_obs, d_rew, NULL, d_done, NULL, OC_FLAG_AUTO_RESET, NULL));
            CU(cudaMemcpy(h_obs, d_obs, sizeof(h_obs), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(h_done, d_done, sizeof(h_done), cudaMemcpyDeviceToHost));
            if (memcmp(h_obs, p_obs, sizeof(h_obs)) != 0 || memcmp(h_done, p_done, E) != 0) {
                fprintf(stderr, "oc_step_host differs from oc_step at t=%d\n", t); return 1;
            }
            CHECK(oc_step_host_i8(hc, p_act, p_o8, p_ts, NULL, NULL, p_done8, p_t8, p_tts, OC_FLAG_AUTO_RESET, NULL));
            if (memcmp(p_done8, p_done, E) != 0) { fprintf(stderr, "oc_step_host_i8: done differs at t=%d\n", t); return 1; }
            for (int e = 0; e < E; ++e)
                for (int k = 0; k < A; ++k) {
                    const float* fr = p_obs + ((size_t)e * A + k) * F;
                    const int8_t* br = p_o8 + ((size_t)e * A + k) * (F - 1);
                    for (int c = 0; c < F - 1; ++c)
                        if ((float)br[c] != fr[c]) { fprintf(stderr, "oc_step_host_i8 differs at t=%d env %d col %d\n", t, e, c); return 1; }
                    if (p_ts[e] != fr[F - 1]) { fprintf(stderr, "oc_step_host_i8 clock differs at t=%d env %d\n", t, e); return 1; }
                    if (p_done[e] && ((float)p_t8[((size_t)e * A + k) * (F - 1)] != p_term[((size_t)e * A + k) * F] ||
                                      p_tts[e] != p_term[((size_t)e * A + k) * F + F - 1])) {
                        fprintf(stderr, "oc_step_host_i8 terminal row differs at t=%d env %d\n", t, e); return 1;
                    }
                }
            for (int e = 0; e < E; ++e)
                if (p_done[e]) {                      /* terminal observation: the clock feature of the last step = 49/50 or earlier */
                    const float ts = p_term[((size_t)e * A) * F + off[OC_OBS_TIMESTEP]];
                    if (!(ts > 0.0f && ts <= 1.0f)) { fprintf(stderr, "terminal observation missing (ts=%f)\n", ts); return 1; }
                    term_rows += 1;
                }
        }
        if (term_rows != (long)E * 2) { fprintf(stderr, "expected %d terminal observations, got %ld\n", E * 2, term_rows); return 1; }
        CHECK(oc_host_free(p_act)); CHECK(oc_host_free(p_obs)); CHECK(oc_host_free(p_term));
        CHECK(oc_host_free(p_rew)); CHECK(oc_host_free(p_done));
        CHECK(oc_host_free(p_o8)); CHECK(oc_host_free(p_t8)); CHECK(oc_host_free(p_ts)); CHECK(oc_host_free(p_tts));
        CHECK(oc_host_free(p_done8));
        CHECK(oc_destroy(ha)); CHECK(oc_destroy(hb)); CHECK(oc_destroy(hc));
        cfg.seed = 7;
    }
    /* error path: a level with two tomatoes is outside the supported domain */
    cfg.object_contents[1] = 1;
    oc_env* bad = NULL;
    if (oc_create(&cfg, &bad) != OC_ERR_INVALID || bad != NULL) { fprintf(stderr, "duplicate food was not rejected\n"); return 1; }
    CHECK(oc_destroy(env));
    printf("cabi_smoke ok: %d envs, 200 steps + 8 fused, %ld episode ends\n", E, dones);
    return 0;
}
