/*
 * overcooked_b200.h -- C ABI of the B200-native batched Overcooked simulator.
 *
 * This is the drop-in boundary for ONE hot path of kyle-he/gym-comm: the environment
 * step + observation featurisation.  The reference has no native/FFI boundary of its own
 * (it is 100 % Python), so each entry point below cites the reference *Python* interface it
 * replaces (file:line relative to the reference root); INTEGRATION.md shows the ctypes stub
 * a maintainer of the reference would add.
 *
 * Conventions
 *  - plain C: pointers and sizes only, no torch/C++ types; never throws across the ABI;
 *  - every function returns OC_OK (0) or a negative oc_status; oc_last_error() gives text;
 *  - the CALLER owns every buffer passed in (device pointers unless stated otherwise); the
 *    handle owns only its packed per-env state and its constant tables;
 *  - all launches are asynchronous on the passed stream (a cudaStream_t cast to void*,
 *    e.g. torch.cuda.current_stream().cuda_stream); nothing here synchronises the host
 *    except oc_create / oc_destroy;
 *  - one handle per device; a handle is not thread-safe;
 *  - there is NO CPU fallback: without a CUDA device oc_create fails with OC_ERR_CUDA.
 *
 * Encodings (shared with DESIGN.md section "HBM layout")
 *  - cell  = y * width + x  (x = column, y = row, origin top-left; overcooked_environment.py:113-131)
 *  - tile  : 0 Floor, 1 Counter, 2 Cutboard, 3 Delivery                       (utils/core.py:18-26)
 *  - nav   : 0 (0,+1)  1 (0,-1)  2 (-1,0)  3 (+1,0)                            (utils/world.py:16)
 *  - content bits: Tomato 1, Lettuce 2, Onion 4, Plate 8 (1 << ObjectChannel)  (utils/core.py:383-388)
 *  - subtask kind: 0 Chop, 1 Merge, 2 Deliver                                  (recipe_planner/utils.py)
 */
#ifndef OVERCOOKED_B200_H
#define OVERCOOKED_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OC_ABI_VERSION 2

#define OC_MAX_AGENTS   4
#define OC_MAX_OBJECTS  6
#define OC_MAX_SUBTASKS 32
#define OC_MAX_CELLS    128
#define OC_MAX_COMM     65534
#define OC_STATE_WORDS  16     /* packed state: 16 x uint32 = 64 B per env */
#define OC_NUM_OBS_KEYS 11

typedef enum oc_status {
    OC_OK = 0,
    OC_ERR_INVALID = -1,   /* bad argument / config outside the supported domain */
    OC_ERR_CUDA = -2,      /* CUDA runtime error (message in oc_last_error)       */
    OC_ERR_ALLOC = -3
} oc_status;

/* step flags */
#define OC_FLAG_AUTO_RESET 1u  /* on done: reset the env in place and return the first obs of the new
                                  episode (SB3 VecEnv contract; DummyVecEnv.step_wait)             */
#define OC_FLAG_ACTIONS_U8 2u      /* oc_step_i8: actions are u8 [E, A, 2] (nav < 4, comm < C <= 256) instead of int32  */
#define OC_FLAG_REWARD_PER_ENV 4u  /* oc_step_i8: rew_f32 is f32 [E] (one value per env) instead of f32 [E, A]           */
#define OC_FLAG_NO_SYNC 8u         /* oc_step_host_block: return after enqueueing; oc_sync() before reading the block     */
/* Chained steps (oc_step / oc_step_i8) -- for runs of steps whose actions are ALL in memory before the first of them is
 * enqueued (recorded, scripted or pre-drawn action sequences; oc_replay is the same thing in one launch when the
 * outputs form one contiguous block).  The first step of the run carries OC_FLAG_CHAIN_HEAD, the following ones
 * OC_FLAG_CHAINED: a chained launch does not wait for the previous grid to drain and retire; each warp of it starts as
 * soon as the states of ITS 32 envs from the previous step are in memory, so the observation stores of step N overlap
 * the dynamics of step N+1.  Rules (violations fail with OC_ERR_INVALID, never silently): a chained step directly follows a
 * head / chained step of the same entry point on the same handle and stream, with nothing else enqueued on that stream
 * in between; consecutive steps write DIFFERENT obs / reward / done / terminal buffers (rollout-buffer slots); the
 * first step of a captured CUDA graph is a head.  A step with neither flag, and any other call, ends the chain. */
#define OC_FLAG_CHAIN_HEAD 16u
#define OC_FLAG_CHAINED 32u

/* Order of the 11 observation keys inside one flat feature row = the key-sorted order a gym
 * spaces.Dict gives the dict built at gym_comm/envs/overcooked_env.py:66-78 (what
 * FlattenedDictExtractor concatenates, gym_comm/extractors/CustomExtractor.py:119-128). */
typedef enum oc_obs_key {
    OC_OBS_AGENT1_COMM = 0, OC_OBS_AGENT1_LOCATION, OC_OBS_AGENT2_COMM, OC_OBS_AGENT2_LOCATION,
    OC_OBS_AGENT_IS_HOLDING, OC_OBS_COMPLETED_SUBTASKS, OC_OBS_IS_HIDDEN, OC_OBS_OBJECT_ENCODINGS_X,
    OC_OBS_OBJECT_ENCODINGS_Y, OC_OBS_STATE_ENCODINGS, OC_OBS_TIMESTEP
} oc_obs_key;

/* One compiled environment configuration.  Host pointers; copied by oc_create.
 * Replaces: the argparse Namespace of arglist.py:96-121 + the level file parsed by
 * OvercookedEnvironment.load_level (gym_cooking/envs/overcooked_environment.py:100-178) + the
 * static subtask table of run_recipes (:452-459) + World.get_path_distance_between
 * (gym_cooking/utils/world.py:114-131) tabulated for every cell pair. */
typedef struct oc_config {
    uint32_t abi_version;          /* = OC_ABI_VERSION */
    int32_t  num_envs;             /* E */
    int32_t  num_agents;           /* A, 2..4  (arglist.num_agents) */
    int32_t  width, height;        /* grid; width*height <= OC_MAX_CELLS */
    int32_t  max_num_timesteps;    /* T, 1..65534 (0 is rejected: the timestep observation is t / T,
                                      overcooked_env.py:146)  overcooked_environment.py:245 */
    int32_t  num_communication;    /* C  (arglist.num_communication) */
    int32_t  communication_on;     /* overcooked_env.py:229 */
    int32_t  ego_led;              /* overcooked_env.py:234 */
    int32_t  fow_radius;           /* overcooked_env.py:133-135 */
    /* per agent: agent 0 = ego_config, others = partner_config (overcooked_environment.py:140-143) */
    uint8_t  can_move[OC_MAX_AGENTS];
    uint8_t  allergic[OC_MAX_AGENTS];
    uint8_t  blind[OC_MAX_AGENTS];
    uint8_t  start_cell[OC_MAX_AGENTS];

    const uint8_t* tiles;          /* [width*height] tile codes */
    /* pd[src*ncell+dst] = get_path_distance_between(src, dst); MAX_PATH when src is not floor or
     * dst is unreachable (world.py:114-131).  MAX_PATH = 2*(w+h)+1 (overcooked_environment.py:274) */
    const uint8_t* path_dist;      /* [ncell*ncell] */
    int32_t  max_path;

    /* objects present at reset, in world insertion order (load_level phase 1, then phase 4) */
    int32_t  num_objects;                      /* <= OC_MAX_OBJECTS */
    uint8_t  object_contents[OC_MAX_OBJECTS];  /* one content bit each */
    int16_t  object_cell[OC_MAX_OBJECTS];      /* -1 = placed at reset on a random Counter (phase 4) */

    /* static subtask table, reference order (PYTHONHASHSEED=0 canonical) */
    int32_t  num_subtasks;                       /* S <= OC_MAX_SUBTASKS */
    uint8_t  subtask_kind[OC_MAX_SUBTASKS];      /* 0 Chop 1 Merge 2 Deliver */
    uint8_t  subtask_goal[OC_MAX_SUBTASKS];      /* goal template: contents | chopped << 4 */
    uint8_t  subtask_arg0[OC_MAX_SUBTASKS];      /* Chop: the food bit */

    /* calculate_reward_shaping item list: Plate + recipes[0].contents sorted by name
     * (overcooked_environment.py:319-321) as content bits */
    int32_t  num_items;
    uint8_t  items[4];

    uint64_t seed;                 /* device RNG seed (random-level placement, oc_rollout actions) */
} oc_config;

typedef struct oc_env oc_env;      /* opaque handle */

/* Replaces gym.make('OvercookedMultiCommEnv-v0', arglist=ns) -> OvercookedMultiEnv.__init__
 * (gym_comm/__init__.py:3-6, gym_comm/envs/overcooked_env.py:16-100) for E envs at once.
 * The envs come up reset (t=0) with comm buffers one-hot at index 0 (overcooked_env.py:89-91). */
int oc_create(const oc_config* cfg, oc_env** out);
int oc_destroy(oc_env* env);

/* Feature width F = 23 + S + 2*C of one observer's flat row (overcooked_env.py:145-157). */
int oc_obs_width(const oc_env* env);
/* offsets[OC_NUM_OBS_KEYS], sizes[OC_NUM_OBS_KEYS]: where each obs key sits inside a row. */
int oc_obs_layout(const oc_env* env, int32_t* offsets, int32_t* sizes);

/* Replaces OvercookedMultiEnv.multi_reset (overcooked_env.py:284-297) /
 * OvercookedEnvironment.reset (overcooked_environment.py:180-206).
 *  mask        u8[E] device or NULL (= all): envs to reset
 *  placements  int32[E, R] device or NULL: cell of each of the R random (phase-4) objects per env,
 *              in phase-4 string order; NULL = draw on device, uniform over all Counter tiles
 *              without replacement (overcooked_environment.py:157-173)
 *  obs         f32[E, A, F] device or NULL: observation of every agent after the reset
 * Comm buffers are NOT cleared (reference behaviour). */
int oc_reset(oc_env* env, const uint8_t* mask, const int32_t* placements, float* obs, void* stream);

/* Replaces OvercookedMultiEnv.multi_step (overcooked_env.py:207-282) = comm write, action decode,
 * CAN_MOVE, OvercookedEnvironment.step (overcooked_environment.py:211-241: check_collisions,
 * interact per agent, done, reward, calculate_reward_shaping x2) and get_observation2 for every
 * agent (overcooked_env.py:105-159), for all E envs in one launch.
 *  actions   int32[E, A, 2] device: (nav in [0,4), comm in [0,C)) per agent
 *  obs       f32[E, A, F] device: flat observation of every agent (post-step; post-reset when
 *            auto-reset fired)
 *  rew_f32   f32[E, A] device or NULL: returned reward, the same value for every agent
 *            (overcooked_env.py:282), rounded once from f64
 *  rew_f64   f64[E] device or NULL: the same reward in the reference's own precision
 *  done      u8[E] device
 *  term_obs  f32[E, A, F] device or NULL: with OC_FLAG_AUTO_RESET, rows of envs that finished
 *            receive the terminal observation (SB3 infos["terminal_observation"]); other rows
 *            are left untouched */
int oc_step(oc_env* env, const int32_t* actions, float* obs, float* rew_f32, double* rew_f64,
            uint8_t* done, float* term_obs, uint32_t flags, void* stream);

/* oc_step / oc_reset with the observations in the COMPACT INTEGER FORMAT (see oc_pack_obs_i8 below), produced by
 * the step / reset kernel itself: obs_i8 int8 [E, A, F-1] (16-byte aligned) + timestep f32 [E] (may be NULL);
 * term_obs_i8 / term_timestep likewise for envs that finished.  Every pointer may be device memory or
 * page-locked host memory (cudaHostAlloc / oc_host_alloc: the kernel then reads / writes across PCIe).
 * flags: OC_FLAG_AUTO_RESET | OC_FLAG_ACTIONS_U8 (actions u8 [E, A, 2]) | OC_FLAG_REWARD_PER_ENV (rew_f32 f32 [E]).
 * Fails with OC_ERR_INVALID when 32 rows of A * (F-1) bytes exceed 64 KB (use oc_step + oc_pack_obs_i8 then). */
int oc_compact_supported(const oc_env* env);   /* 1 when oc_reset_i8 / oc_step_i8 / oc_*_host_block can run on this handle */
int oc_reset_i8(oc_env* env, const uint8_t* mask, const int32_t* placements, int8_t* obs_i8, float* timestep,
                void* stream);
int oc_step_i8(oc_env* env, const void* actions, int8_t* obs_i8, float* timestep, float* rew_f32, double* rew_f64,
               uint8_t* done, int8_t* term_obs_i8, float* term_timestep, uint32_t flags, void* stream);

/* Fused synthetic rollout (the throughput benchmark of SURVEY section 8d): n_steps env steps in
 * ONE launch with state kept on chip; actions nav~U{0..3}, comm~U{0..C-1} from Philox4x32-10
 * keyed by (seed; env index, global step); auto-reset always on.  Step s writes
 *  obs[s]  f32[n_steps, E, A, F],  rew_f32[s] f32[n_steps, E, A],  done[s] u8[n_steps, E]
 * (any of them may be NULL = not written);  actions_out int32[n_steps, E, A, 2] or NULL.
 * The global step counter of a handle is 32-bit: its action stream repeats after 2^32 fused steps. */
int oc_rollout(oc_env* env, int32_t n_steps, float* obs, float* rew_f32, uint8_t* done,
               int32_t* actions_out, void* stream);

/* Open-loop replay: like oc_rollout, but step s applies the caller's actions[s] (int32
 * [n_steps, E, A, 2], device) instead of drawing them -- n_steps env steps in ONE launch with the
 * state kept on chip, auto-reset on.  For evaluating recorded / scripted / planned action sequences
 * (the reference's tester.py:72-106 loop with a fixed action list). */
int oc_replay(oc_env* env, int32_t n_steps, const int32_t* actions, float* obs, float* rew_f32,
              uint8_t* done, void* stream);

/* Packed state export / injection (parity tests; checkpointing).  state: u32[E, OC_STATE_WORDS]
 * device, layout in DESIGN.md.  Replaces OvercookedEnvironment.__copy__/get_repr
 * (overcooked_environment.py:59-84). */
int oc_get_state(oc_env* env, uint32_t* state, void* stream);
/* Imported words are sanitised: agent cells and live object cells are clamped to the grid, holders to
 * {0..A-1, none}, empty slots become the canonical dead word -- a state exported from another level cannot
 * make the kernels index outside their tables. */
int oc_set_state(oc_env* env, const uint32_t* state, void* stream);

/* Per-env statistics kept on device: episodes finished, and completed-subtask count of the last
 * finished episode (episode_recorder.py:29).  Either pointer may be NULL. */
int oc_get_stats(oc_env* env, uint32_t* episodes /*[E]*/, uint32_t* last_completed /*[E]*/, void* stream);

/* Number of kernel launches issued through this handle so far (bench.py "gpu_launches"). */
uint64_t oc_launch_count(const oc_env* env);

/* ---- Host-buffer entry points: for a caller that keeps its buffers in HOST memory (numpy arrays, SB3 rollout
 * buffers on the CPU -- exactly what the reference's DummyVecEnv.step_wait hands to
 * sb3_contrib/ppo_recurrent/ppo_recurrent.py:233-252).  Same arguments and semantics as oc_reset / oc_step,
 * but every pointer is a HOST pointer; the handle stages through device buffers it owns (allocated on first
 * use), copies host->device and device->host on `stream`, and SYNCHRONISES the stream before returning, so
 * the results are valid on return.  Pinned buffers (oc_host_alloc) make the copies run at PCIe speed;
 * pageable memory works but is slower.  term_obs: only the rows of envs that finished in this step are
 * written; other rows of the caller's buffer are left untouched. */
int oc_reset_host(oc_env* env, const uint8_t* mask, const int32_t* placements, float* obs, void* stream);
int oc_step_host(oc_env* env, const int32_t* actions, float* obs, float* rew_f32, double* rew_f64,
                 uint8_t* done, float* term_obs, uint32_t flags, void* stream);
/* ---- Compact integer format.  Every key of an observation row except the clock holds small integers -- the
 * reference builds them as int64 arrays (get_observation2, gym_comm/envs/overcooked_env.py:145-157; the spaces
 * at :56-78 are Box(int64) / MultiBinary) and only SB3's preprocessing turns them into float32.  For a consumer
 * on the far side of PCIe the same values travel as
 *    obs_i8    int8[E, A, F-1]  the row with the `timestep` column (the last key, oc_obs_layout) cut out
 *    timestep  f32[E]           float32(t / max_num_timesteps), identical for every agent of an env
 * i.e. 4x fewer bytes than the float rows; `obs_i8.astype(float32)` with `timestep` appended IS the float row.
 * oc_pack_obs_i8 converts float rows already on the DEVICE (all pointers device, obs_i8 4-byte aligned,
 * timestep may be NULL); the _host_i8 entry points are oc_reset_host / oc_step_host with the observation
 * buffers in this format (HOST pointers; term_timestep[e] is written only for envs that finished). */
int oc_pack_obs_i8(oc_env* env, const float* obs, int8_t* obs_i8, float* timestep, void* stream);
int oc_reset_host_i8(oc_env* env, const uint8_t* mask, const int32_t* placements, int8_t* obs_i8, float* timestep,
                     void* stream);
int oc_step_host_i8(oc_env* env, const int32_t* actions, int8_t* obs_i8, float* timestep, float* rew_f32,
                    double* rew_f64, uint8_t* done, int8_t* term_obs_i8, float* term_timestep, uint32_t flags,
                    void* stream);
/* ---- One-block host path: everything a step returns in ONE page-locked block, so that one device->host copy
 * per step carries it (the host-side ceiling of this path is the PCIe link).  The block holds, at
 * the 256-byte aligned offsets oc_host_block_layout reports,
 *    obs_i8 int8 [E, A, F-1] | timestep f32 [E] | reward f32 [E] | done u8 [E]
 * (reward: ONE value per env -- the reference hands the same reward to both players, overcooked_env.py:282).
 * actions_u8: u8 [E, A, 2] host (nav, comm per agent; C <= 256).  When it is page-locked the kernel reads it
 * across PCIe itself and the step is one launch + one copy.  term_obs_i8 / term_timestep (NULL or page-locked
 * host, int8 [E, A, F-1] / f32 [E]): rows of envs that finished are written by the kernel directly.
 * flags: OC_FLAG_AUTO_RESET | OC_FLAG_NO_SYNC (return after enqueueing; call oc_sync before reading). */
typedef struct oc_host_block {
    uint64_t obs_i8, timestep, reward, done;   /* byte offsets */
    uint64_t total_bytes;
} oc_host_block;
int oc_host_block_layout(const oc_env* env, oc_host_block* out);
int oc_reset_host_block(oc_env* env, const uint8_t* mask, const int32_t* placements, void* block, void* stream);
int oc_step_host_block(oc_env* env, const uint8_t* actions_u8, void* block, int8_t* term_obs_i8,
                       float* term_timestep, uint32_t flags, void* stream);
/* cudaStreamSynchronize(stream) for callers that do not link the CUDA runtime. */
int oc_sync(oc_env* env, void* stream);
/* oc_get_state / oc_set_state with a HOST buffer u32 [E, OC_STATE_WORDS] (checkpointing from numpy); synchronous. */
int oc_get_state_host(oc_env* env, uint32_t* state, void* stream);
int oc_set_state_host(oc_env* env, const uint32_t* state, void* stream);

/* Page-locked host memory for those buffers (cudaHostAlloc / cudaFreeHost without linking the CUDA runtime). */
int oc_host_alloc(uint64_t bytes, void** out);
int oc_host_free(void* ptr);
/* Select the CUDA device of the calling thread (cudaSetDevice) -- for callers that do not otherwise touch
 * the CUDA runtime; a handle stays bound to the device that was current at oc_create. */
int oc_set_device(int device);

const char* oc_last_error(void);
int oc_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OVERCOOKED_B200_H */
