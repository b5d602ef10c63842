/*
 * blockpuzzle_b200.h -- C-ABI of the B200-native batched gym_blocks environment
 * hot path (libblockpuzzle_b200.so, built from blockpuzzle_gym_b200/csrc).
 *
 * The reference (matthew9671/BlockPuzzle-gym) has no FFI layer: the hot path sits
 * behind the Python gym GoalEnv API.  Each entry point below names the reference
 * interface it replaces (paths relative to /root/reference/gym_blocks); the
 * Python binding a maintainer would add is shown in INTEGRATION.md and shipped
 * in blockpuzzle_gym_b200/_lib.py.
 *
 * Conventions: plain C, no torch types.  Every function returns 0 on success or
 * a negative bp_status; bp_last_error() returns the message of the calling
 * thread's last failure.  Pointers named d_* are DEVICE pointers owned by the
 * caller (e.g. torch tensors); h_* are HOST pointers.  `stream` is a
 * cudaStream_t passed as void* (0 = legacy default stream).  Calls enqueue work
 * on `stream` and return without synchronising unless stated otherwise.  One
 * handle per (GPU, env id); a handle is not thread-safe.
 */
#ifndef BLOCKPUZZLE_B200_H
#define BLOCKPUZZLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_ABI_VERSION 2   /* 2: bp_step_host takes the caller's stream; bp_rollout_begin / bp_rollout_step; BP_ERR_NO_ATTRIBUTE */

typedef enum {
    BP_OK = 0,
    BP_ERR_INVALID_ARG = -1,
    BP_ERR_CUDA = -2,
    BP_ERR_NOT_IMPLEMENTED = -3, /* the reference method raises NotImplementedError */
    BP_ERR_NO_DEVICE = -4,
    BP_ERR_NO_ATTRIBUTE = -5     /* the reference method raises AttributeError (fetch_env.py:413-415,420) */
} bp_status;

/* env ids in the registration order of __init__.py:6-53 */
typedef enum {
    BP_GRIPPER_TOUCH = 0,                   /* 'GripperTouch-v0'                __init__.py:7  */
    BP_BLOCKS_TOUCH = 1,                    /* 'BlocksTouch-v0'                 __init__.py:14 */
    BP_TOPPLE_TOWER = 2,                    /* 'ToppleTower-v0'                 __init__.py:21 */
    BP_BLOCKS_TOUCH_CURRICULUM = 3,         /* 'BlocksTouchCurriculum-v0'       __init__.py:28 */
    BP_BLOCKS_TOUCH_CHOOSE = 4,             /* 'BlocksTouchChoose-v0'           __init__.py:35 */
    BP_BLOCKS_TOUCH_CHOOSE_CURRICULUM = 5,  /* 'BlocksTouchChooseCurriculum-v0' __init__.py:42 */
    BP_BLOCKS_TOUCH_VARIATION = 6,          /* 'BlocksTouchVariation-v0'        __init__.py:49 */
    BP_NUM_ENV_IDS = 7
} bp_env_id;

#define BP_MAX_BLOCKS 4
#define BP_MAX_DIMG 36
#define BP_MAX_DIMO 87
#define BP_MAX_EPISODE_STEPS 50 /* max_episode_steps, __init__.py:10 */
#define BP_NUM_STATS 8

/* indices into the stats vector (float64[BP_NUM_STATS]); this is the vector the
 * single NCCL all-reduce carries (replaces mpi_moments, train.py:21-26) */
enum {
    BP_STAT_EPISODES = 0,   /* episodes that reached T = 50 */
    BP_STAT_SUCCESSES = 1,  /* of those, is_success at the last step (rollout.py:163-167) */
    BP_STAT_STEPS = 2,
    BP_STAT_INVALID = 3,    /* non-finite action components (replaces the NaN restart, rollout.py:139-142) */
    BP_STAT_REWARD_SUM = 4,
    BP_STAT_WORKER_STEPS = 5 /* diagnostic: env-steps that took the full-physics worker path of the tiled kernel */
};

/* Canonical per-env state record used by bp_get_state / bp_set_state (the
 * bit-exact replay harness).  244 bytes, little endian, no padding. */
typedef struct bp_env_state {
    float grip_pos[3];
    float grip_vel[3];
    float finger_q[2];
    float finger_qv[2];
    float blk_pos[BP_MAX_BLOCKS][3];
    float blk_cs[BP_MAX_BLOCKS][2];   /* yaw as (cos, sin) */
    float blk_vel[BP_MAX_BLOCKS][3];
    float blk_w[BP_MAX_BLOCKS];
    int8_t ag[BP_MAX_DIMG];           /* touch matrix self.achieved_goal, fetch_env.py:78 */
    int32_t num_objs;                 /* fetch_env.py:75 */
    int32_t has_succeeded;            /* fetch_env.py:80 */
    int32_t t;                        /* TimeLimit elapsed steps */
    uint32_t episode;                 /* reset() calls so far */
    uint32_t draws[2];                /* Philox draw counters: [0] self.np_random, [1] global np.random */
} bp_env_state;

typedef struct bp_handle bp_handle;

/* ---- introspection ---- */
int bp_abi_version(void);
const char* bp_last_error(void);
/* dims per env id (observation_space, robot_env.py:40-44) */
int bp_env_dims(int env_id, int* dimo, int* dimg, int* nblocks);
/* name <-> id of the ids registered in __init__.py:6-53 */
int bp_env_id_from_name(const char* name);
const char* bp_env_name(int env_id);

/* ---- lifetime: replaces gym.make(env_name) x num_envs (rollout.py:33) ---- */
/* Allocates the device state for `num_envs` envs of `env_id` on CUDA device
 * `device`.  `env_index_offset` is the global index of this handle's env 0
 * (multi-GPU sharding: results do not depend on how envs are split). */
int bp_create(int env_id, int64_t num_envs, int device, uint64_t env_index_offset, bp_handle** out);
int bp_destroy(bp_handle* h);
int64_t bp_num_envs(const bp_handle* h);

/* RolloutStudent.seed (rollout.py:206-210) + RobotEnv.seed (robot_env.py:53-55):
 * env i gets seed + 1000 * (env_index_offset + i). */
int bp_seed(bp_handle* h, uint64_t seed, void* stream);

/* RobotEnv.reset (robot_env.py:71-82) for every env whose d_mask byte is
 * non-zero (d_mask == NULL: all).  Outputs (any may be NULL): d_obs
 * [B][dimo], d_ag [B][dimg], d_g [B][dimg]; rows of unselected envs are not
 * written. */
int bp_reset(bp_handle* h, const uint8_t* d_mask, float* d_obs, float* d_ag, float* d_g, void* stream);

/* RobotEnv.step (robot_env.py:57-69) under TimeLimit, K fused steps per launch.
 *   d_actions  [K][B][4] float32, or NULL: draw them in-kernel from Philox
 *              stream 2 (the replay harness); d_actions_out (nullable) receives them.
 *   d_obs      [K][B][dimo]   d_ag [K][B][dimg]   d_reward [K][B]
 *   d_success  [K][B] float32 info['is_success'] (latched, fetch_env.py:275-281)
 *   d_done     [K][B] uint8   TimeLimit done (nullable)
 *   auto_reset != 0: an env that finishes step T = 50 is reset inside the kernel
 *              after its outputs are written; its fresh reset observation goes to
 *              d_reset_obs [B][dimo] / d_reset_ag [B][dimg] (nullable).
 * Any output pointer may be NULL (that output is skipped).
 * Alignment: every tensor must be 16-byte aligned; with d_obs and d_ag 32-byte
 * aligned (any cudaMalloc / torch allocation) the plain fused step takes its
 * fastest instantiation (256-bit row stores), otherwise the general one. */
int bp_step(bp_handle* h, const float* d_actions, int K, float* d_obs, float* d_ag, float* d_reward,
            float* d_success, uint8_t* d_done, int auto_reset, float* d_reset_obs, float* d_reset_ag,
            float* d_actions_out, void* stream);

/* Same call with HOST buffers (pinned or pageable): actions are copied to the
 * device, outputs copied back, chunked over envs and double-buffered on two
 * internal streams.  Synchronous.  This is the end-to-end path a CPU trainer
 * (rollout.py:121-131) uses.  `stream` is the stream the caller queued its
 * earlier calls on this handle on (bp_seed / bp_reset / bp_set_state / bp_step):
 * the internal streams wait for it before the first copy. */
int bp_step_host(bp_handle* h, const float* h_actions, int K, float* h_obs, float* h_ag,
                 float* h_reward, float* h_success, int auto_reset, void* stream);

/* RolloutStudent.generate_rollouts (rollout.py:75-172) for open-loop actions, with the per-env loop
 * (rollout.py:121-131) and convert_episode_to_batch_major (util.py:118-128) fused into the step
 * kernel: resets every env (plus set_test() when `test`, rollout.py:52-55), runs T = 50 steps on
 * d_actions [T][B][4] (NULL: in-kernel Philox actions) and writes the episode batch-major:
 *   d_o [B][T+1][dimo], d_ag [B][T+1][dimg]   (slot 0 = the reset observation)
 *   d_g [B][T][dimg], d_u [B][T][4], d_success [B][T] (info_is_success), d_reward [B][T]
 * d_g, d_u, d_success, d_reward may be NULL.  No auto-reset: the episode ends at T. */
int bp_rollout(bp_handle* h, const float* d_actions, int test, float* d_o, float* d_ag, float* d_g, float* d_u,
               float* d_success, float* d_reward, void* stream);

/* The same collector CLOSED-LOOP: RolloutStudent asks the policy for u_t = policy.get_actions(o_t, ag_t, g) at every
 * one of the T steps (rollout.py:91-97), so the actions cannot be known up front.
 *   bp_rollout_begin  = reset_all_rollouts (rollout.py:48-64; + set_test() when `test`): writes slot 0 of
 *                       d_o [B][T+1][dimo] / d_ag [B][T+1][dimg] and the goal rows d_g0 [B][dimg] (nullable);
 *   bp_rollout_step t = one env step of every env on d_actions [B][4] (the policy's output for slot t): writes slot
 *                       t + 1 of d_o / d_ag and slot t of d_g [B][T][dimg], d_u [B][T][4], d_success [B][T],
 *                       d_reward [B][T] (each nullable) -- the same batch-major episode bp_rollout produces.
 * One launch per step on `stream`, no host synchronisation: the T-step loop (policy included) can be captured in a
 * CUDA graph.  t must run 0, 1, ..., T - 1 after a bp_rollout_begin. */
int bp_rollout_begin(bp_handle* h, int test, float* d_o, float* d_ag, float* d_g0, void* stream);
int bp_rollout_step(bp_handle* h, int t, const float* d_actions, float* d_o, float* d_ag, float* d_g, float* d_u,
                    float* d_success, float* d_reward, void* stream);

/* BlocksTouchEnv.set_test / BlocksTouchChooseEnv.set_test / Variation.set_test
 * (fetch_env.py:365-368, 443-446, 641-644); BP_ERR_NOT_IMPLEMENTED for
 * GripperTouch / ToppleTower (fetch_env.py:100-101). */
int bp_set_test(bp_handle* h, float* d_obs, float* d_ag, float* d_g, void* stream);

/* increase_difficulty (fetch_env.py:351-358, 419-432, 623-630): *max_reached
 * receives the python return value; BP_ERR_NOT_IMPLEMENTED where the reference raises NotImplementedError
 * (GripperTouch, ToppleTower: fetch_env.py:93-94), BP_ERR_NO_ATTRIBUTE where it raises AttributeError
 * (BlocksTouchChoose-v0 without curriculum: obj_range_step is never set, fetch_env.py:413-415,420). */
int bp_increase_difficulty(bp_handle* h, int* max_reached);
int bp_get_difficulty(const bp_handle* h, int* difficulty);           /* fetch_env.py:96-97 */
/* "challenge" (BlocksTouchChoose ids only; BP_ERR_INVALID_ARG otherwise): the `challenge` constructor argument of
 * BlocksTouchChooseEnv (fetch_env.py:403,416) -- every spawn then uses max_obj_range, keeps the wrong block within 0.04 of
 * the pair's centre and the pair at least 0.15 apart (fetch_env.py:452-463).  No tasks.py class sets it.
 * Measurement knobs (no reference counterpart): "step_kernel" (-1 default, 0 async, 2 simple, 4 split);
 * "force_full_physics" != 0: every env-step runs the complete
 * BlockPhys step (no quiet path) -- results are identical, only slower: the floor bench.py reports. */
int bp_set_option(bp_handle* h, const char* name, int value);
/* direct access to the curriculum knobs (obj_range, wrong_obj_range, max_obj_range) */
int bp_get_ranges(const bp_handle* h, double* obj_range, double* wrong_obj_range, double* max_obj_range);
int bp_set_ranges(bp_handle* h, double obj_range, double wrong_obj_range);

/* replay harness: canonical state records, [B] bp_env_state on the device */
int bp_get_state(bp_handle* h, bp_env_state* d_out, void* stream);
int bp_set_state(bp_handle* h, const bp_env_state* d_in, void* stream);

/* statistics: d_stats float64[BP_NUM_STATS] lives in the handle; bp_stats_ptr
 * returns its device address so the caller can all-reduce it in place. */
int bp_stats_ptr(bp_handle* h, double** d_stats);
int bp_stats_reset(bp_handle* h, void* stream);

/* BlocksEnv.compute_reward (fetch_env.py:135-143), batched: d_ag, d_g [n][dimg]
 * float32 -> d_r [n] float32 (-0.0 success, -1.0 otherwise).  Called through
 * reward_fun (config.py:110-111). */
int bp_compute_reward(const float* d_ag, const float* d_g, int64_t n, int dimg, float* d_r, void* stream);

/* HER relabel + reward: baselines.her.her._sample_her_transitions [upstream],
 * call sites config.py:121, ddpg.py:106,214-215.
 *   d_ep_ag [B][T+1][dimg], d_ep_g [B][T][dimg]: the episode store
 *   n transitions; transition i uses Philox counter (index_offset + i), stream 3
 *   outputs (nullable): d_ep_idx, d_t, d_future_t int32 [n] (future_t = -1 when
 *   not relabelled), d_ag2 [n][dimg], d_g [n][dimg], d_r [n]. */
int bp_her_relabel(const float* d_ep_ag, const float* d_ep_g, int32_t B, int32_t T, int32_t dimg,
                   int64_t n, float future_p, uint64_t seed, int64_t index_offset, int32_t* d_ep_idx,
                   int32_t* d_t, int32_t* d_future_t, float* d_ag2, float* d_g, float* d_r, void* stream);

/* ---- the callers' data path either side of the env ---- */

/* ReplayBuffer.sample + _sample_her_transitions [upstream baselines.her.replay_buffer / her.her]
 * as configured by config.py:107-123 and called from ddpg.py:106 (buffer), :168-171 (store_episode,
 * normaliser statistics) and :214-222 (sample_batch), with DDPG._preprocess_og's clip
 * (ddpg.py:111-120, clip_obs = 200 config.py:35, relative_goals = False config.py:37) fused in.
 *   episode store (device): d_ep_o [B][T+1][dimo], d_ep_u [B][T][dimu], d_ep_g [B][T][dimg],
 *     d_ep_ag [B][T+1][dimg], d_ep_succ [B][T] (info_is_success); d_ep_o / d_ep_u / d_ep_succ may
 *     be NULL when the matching outputs are NULL.
 *   Transition i draws Philox block (index_offset + i) of stream 3 under `seed` -- the same draw
 *   bp_her_relabel uses: episode = randint(B), t = randint(T), relabel when uniform < future_p with
 *   future_t = t + 1 + int(uniform * (T - t)).
 *   outputs (any may be NULL): d_ep_idx, d_t, d_future_t int32 [n] (future_t = -1: not relabelled);
 *     d_o = clip(o[e, t]), d_o2 = clip(o[e, t + 1]) [n][dimo]; d_u [n][dimu]; d_g = clip(relabelled
 *     goal) [n][dimg] (g_2 of sample_batch is the same tensor); d_ag = ag[e, t], d_ag2 = ag[e, t + 1]
 *     [n][dimg]; d_r [n] = compute_reward(ag_2, g); d_succ [n].  clip_obs <= 0: no clip (what
 *     buffer.sample returns before _preprocess_og).
 *   d_stats (nullable) float64 [2 * dimo + 1]: sum and sum of squares per column of the d_o rows and
 *     the row count are ADDED to it -- the quantities Normalizer.update accumulates at ddpg.py:185
 *     (for the Variation env the caller drops column 0, ddpg.py:180-181). */
int bp_her_sample(const float* d_ep_o, const float* d_ep_u, const float* d_ep_g, const float* d_ep_ag,
                  const float* d_ep_succ, int32_t B, int32_t T, int32_t dimo, int32_t dimu, int32_t dimg,
                  int64_t n, float future_p, float clip_obs, uint64_t seed, int64_t index_offset,
                  int32_t* d_ep_idx, int32_t* d_t, int32_t* d_future_t, float* d_o, float* d_o2, float* d_u,
                  float* d_g, float* d_ag, float* d_ag2, float* d_r, float* d_succ, double* d_stats,
                  void* stream);

/* Normalizer.update(v) [upstream baselines.her.normalizer], call site ddpg.py:185: adds
 * sum_r clip(x[r][col0 + c]), its square and the row count n to d_acc float64 [2 * dim + 1].
 * x has n rows of stride ld floats (col0 = 1, dim = ld - 1 is the Variation rule of ddpg.py:180-181). */
int bp_moments(const float* d_x, int64_t n, int32_t dim, int32_t ld, int32_t col0, float clip, double* d_acc,
               void* stream);

/* policy_gradient/rollout.py:255-258: d_G[b][t] = r[b][t] + sum_{j>=1} gamma_pow[j] * r[b][t + j],
 * accumulated in increasing j in float64 (gamma_pow[j] = gamma ** j, T doubles on the device;
 * gamma = 1 - 1/T, policy_gradient/config.py:84).  d_r [B][T] float32, d_G [B][T] float64. */
int bp_discounted_returns(const float* d_r, int64_t B, int32_t T, const double* d_gamma_pow, double* d_G,
                          void* stream);

/* RolloutStudent.trim, batched branch (policy_gradient/rollout.py:139-171): cut padded rows down to
 * the num_objs objects an expert policy was trained on.  d_o [n][dimo_in], d_g / d_ag [n][dimg_in] ->
 * d_o_out [n][dimo_out], d_g_out / d_ag_out [n][num_objs^2] (outputs nullable).  variation != 0 applies
 * the BlocksTouchVariation rule (:151-167: drop the block count, keep the GREEN / BLUE blocks' 15 base
 * features), else o[:, :dimo_out] (:169). */
int bp_trim(const float* d_o, const float* d_g, const float* d_ag, int64_t n, int32_t dimo_in, int32_t dimg_in,
            int32_t dimo_out, int32_t num_objs, int32_t variation, float* d_o_out, float* d_g_out, float* d_ag_out,
            void* stream);

#ifdef __cplusplus
}
#endif
#endif
