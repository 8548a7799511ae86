// Kernel-side view of one compiled env configuration (passed BY VALUE as a __grid_constant__
// kernel parameter: uniform reads come from the constant bank, no global __constant__ state, so
// several handles with different levels can coexist in one process).
#pragma once
#include <stdint.h>

#define OCK_MAX_AGENTS 4
#define OCK_MAX_OBJECTS 6
#define OCK_MAX_PAIRS 6
#define OCK_MAX_DELIVER 4

// ---- packed per-env state: 16 x u32, stored as 4 SoA planes of uint4 (plane p, env e at
//      state[p * E + e]) so a warp's 32 envs load/store 512 contiguous bytes per plane.
// w0  t[0:16] | next_stamp[16:24] | nkeys[24:32]
// w1  episodes finished
// w2  completed_subtasks bits            (overcooked_environment.py:202)
// w3  goal_objects_count bits (0/1 each) (overcooked_environment.py:201; see DESIGN.md domain note)
// w4  agent cells, 4 x u8
// w5  last_completed[0:8] (popcount of completed at the end of the last episode)
// w6,w7  world.objects key ranks: 16 x 4 bit, index = contents mask, 0 = key absent, else rank+1
// w8..w13 objects: contents[0:4] | chopped[4:7] | holder[8:11] (7 = not held) | cell[16:24] | stamp[24:32]
//         (OCK_DEAD = empty slot: no contents, nobody holds it, cell 0xFF -- matches no holder / cell test)
// w14 comm index agent0 [0:16] | agent1 [16:32]   (0xFFFF = all-zero vector)
// w15 reserved
#define OCK_HOLDER_NONE 7u
#define OCK_DEAD 0x00FF0700u
#define OCK_COMM_NONE 0xFFFFu

struct OcParams {
    int32_t E, A, NOBJ, NF;   // NOBJ / NF: object slots (2 / 4 / 6) and food channels (1 / 2 / 3) the kernels are instantiated for
    int32_t W, H, ncell;
    int32_t T, C, S, F;
    int32_t fow, M;
    int32_t row_bytes;        // A * F = features (floats) of one env row in global memory
    int32_t use_tma;          // float rows leave shared memory through cp.async.bulk (TMA) instead of LDS/STG
    int32_t rowf;             // 1: float32 rows in shared memory (row_stride = 4 * row_bytes), 0: biased-byte rows
    int32_t row_stride;       // shared-memory bytes per env row.  Byte rows: == row_bytes when that is an odd
                              // number of words (contiguous AND conflict-free), else padded to one
    uint8_t can_move[OCK_MAX_AGENTS];
    uint8_t allergic[OCK_MAX_AGENTS];
    uint8_t blind[OCK_MAX_AGENTS];
    uint8_t start_cell[OCK_MAX_AGENTS];
    uint8_t comm_on, ego_led, ego_blind, delivery0;
    // reset image
    uint32_t init_obj[OCK_MAX_OBJECTS];     // object words at reset (cell = 0 for random ones)
    uint8_t  random_slot[OCK_MAX_OBJECTS];  // slots placed on a random counter, phase-4 order
    int32_t  nrandom, ncounters;
    uint32_t init_w0;                       // t=0 | next_stamp | nkeys
    uint64_t init_ranks;                    // ranks AFTER all initial inserts
    // reward tables
    uint32_t deliver_mask, nondeliver_mask;
    uint32_t chop_mask[3];                  // per food bit: the Chop(food) subtasks
    int32_t  ndeliver;
    uint8_t  deliver_sig[OCK_MAX_DELIVER];  // in table order
    uint8_t  deliver_idx[OCK_MAX_DELIVER];
    int32_t  npairs;                        // C(num_items, 2) item pairs of calculate_reward_shaping
    uint32_t item_foods;                    // Food bits among the shaping items (items[0] is always Plate)
    uint32_t r4_magic, rf_magic;            // floor(2^32 / d) + 1 for d = row_bytes / 4 and row_bytes
    int32_t nb, nb_shift;                   // env rows in one warp's shared-memory buffer: 32, or 16 / 8 / 4 (wide float rows)
    int32_t obs_passes;                     // 32 / nb: a warp emits its 32 envs in this many passes
    int32_t nbuf, buf_bytes;                // row buffers per warp (1 or 2, used alternately) of nb * row_stride bytes each
    int32_t warp_row_bytes;                 // nbuf * buf_bytes
    // grouped float rows (row sizes that are a multiple of 8 words would put every lane's scatter on the same bank):
    // rows stay contiguous in groups of 2^grp_shift, grp_pad bytes follow every group -- a group leaves in ONE bulk copy
    // and lanes of different groups hit different banks.  grp_pad == 0: no grouping.
    int32_t grp_shift, grp_pad;
    int32_t obs_rot;                        // 1: lanes of different octets start their rows with different observers (OC_OBS_ROT)
    // observation layout (float offsets inside one observer row)
    int32_t off_a1comm, off_a1loc, off_a2comm, off_a2loc, off_hold, off_completed,
            off_hidden, off_encx, off_ency, off_state, off_ts;
    uint64_t seed;
    // table blob (device pointer) and the byte offsets of its sections; copied to smem per CTA
    const uint8_t* blob;
    int32_t blob_bytes;        // multiple of 16
    int32_t o_q, o_tmlut, o_xyf, o_mvt, o_xy16, o_dmin, o_counters, o_pd, o_pdm, o_ts /* -1: timestep table not in the blob */;
    const float* ts_table;     // [T+1] float32(t / T)   (overcooked_env.py:146)
};
