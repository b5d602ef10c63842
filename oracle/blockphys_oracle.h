/*
 * blockphys_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the gym_blocks environment hot path of
 * matthew9671/BlockPuzzle-gym, used only as the checker for the CUDA path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference).
 * Nothing in blockpuzzle_gym_b200/ may include, link or call this file.
 *
 * PARITY STATUS.  The reference-OWNED logic (action map, touch matrix, obs layout,
 * goal, reward, success latch, spawn samplers, curriculum, TimeLimit) is restated
 * line by line below AND PINNED to the reference's own code: oracle/refharness runs
 * the unmodified /root/reference/gym_blocks under stub gym / mujoco_py packages with
 * this file's BlockPhys model behind a fake MjSim and Philox behind the two numpy
 * RNGs; tests/test_ref_pin.py compares that run with this restatement (integer state,
 * draw counters and the binary32 sim state bit-exact; float64 observations within
 * 1e-6) live when /root/reference or oracle/_ref is present and through the committed
 * fixtures tests/golden/ref_*.npz otherwise.
 * The sim.step() slot itself stays "own spec": the reference delegates rigid-body
 * dynamics to MuJoCo (robot_env.py:60), absent from /root/reference and from this
 * image; it is filled by the "BlockPhys v2" model of DESIGN.md, which this file
 * implements normatively.
 *
 * Dynamics arithmetic is IEEE-754 binary32, the spawn samplers binary64 (as the
 * reference's numpy code), round-to-nearest-even, every operation individually
 * rounded except the explicit fma()s (compile with -ffp-contract=off, no -ffast-math).
 */
#ifndef BLOCKPHYS_ORACLE_H
#define BLOCKPHYS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- env ids, in the registration order of gym_blocks/__init__.py:6-53 ---- */
enum {
    BPO_GRIPPER_TOUCH = 0,            /* GripperTouch-v0                 __init__.py:7  */
    BPO_BLOCKS_TOUCH = 1,             /* BlocksTouch-v0                  __init__.py:14 */
    BPO_TOPPLE_TOWER = 2,             /* ToppleTower-v0                  __init__.py:21 */
    BPO_BLOCKS_TOUCH_CURRICULUM = 3,  /* BlocksTouchCurriculum-v0        __init__.py:28 */
    BPO_BLOCKS_TOUCH_CHOOSE = 4,      /* BlocksTouchChoose-v0            __init__.py:35 */
    BPO_BLOCKS_TOUCH_CHOOSE_CURRICULUM = 5, /* BlocksTouchChooseCurriculum-v0 __init__.py:42 */
    BPO_BLOCKS_TOUCH_VARIATION = 6,   /* BlocksTouchVariation-v0         __init__.py:49 */
    BPO_NUM_ENV_IDS = 7
};

#define BPO_MAX_BLOCKS 4
#define BPO_MAX_OBJS 6
#define BPO_MAX_DIMG 36
#define BPO_MAX_DIMO 87
#define BPO_MAX_EPISODE_STEPS 50 /* __init__.py:10 */

/* colours, fetch_env.py:11-15 */
enum { BPO_GREY = 0, BPO_RED = 1, BPO_GREEN = 2, BPO_BLUE = 3, BPO_NUM_COLORS = 4 };

/* ---- the dynamics slot (replaces MjSim; robot_env.py:24-25,60) ---- */
typedef struct {
    float pos[3];
    float c, s;    /* yaw as a unit complex number (cos, sin) */
    float vel[3];
    float w;       /* yaw rate */
} bpo_block;

typedef struct {
    float g[3];    /* grip site position (robot0:grip) */
    float gv[3];   /* grip site linear velocity */
    float q[2];    /* finger joint positions  [r, l]  (robot.xml:86-95) */
    float qv[2];   /* finger joint velocities [r, l] */
    float m[3];    /* mocap target, set by set_action */
    float ctrl[2]; /* finger position-actuator targets, set by set_action */
    bpo_block blk[BPO_MAX_BLOCKS];
    int32_t nblocks;
    int32_t block_gripper;
    uint32_t contacts; /* contact pairs found in the most recent substep, bit = pair_index(o1,o2) */
} bpo_sim;

/* ---- canonical per-env state record (same layout as bp_env_state in
 *      include/blockpuzzle_b200.h; tests compare the two byte for byte) ---- */
typedef struct {
    float grip_pos[3];
    float grip_vel[3];
    float finger_q[2];
    float finger_qv[2];
    float blk_pos[BPO_MAX_BLOCKS][3];
    float blk_cs[BPO_MAX_BLOCKS][2];
    float blk_vel[BPO_MAX_BLOCKS][3];
    float blk_w[BPO_MAX_BLOCKS];
    int8_t ag[BPO_MAX_DIMG];   /* touch matrix, row-major num_objs x num_objs, values -1/0/1 */
    int32_t num_objs;
    int32_t has_succeeded;
    int32_t t;                 /* TimeLimit elapsed steps */
    uint32_t episode;          /* number of reset() calls so far */
    uint32_t draws[2];         /* draw counters of stream 0 (env np_random) and 1 (global np.random) */
} bpo_env_state;

/* ---- one environment ---- */
typedef struct {
    int32_t env_id;
    int32_t nblocks_max;       /* blocks in the XML (tasks.py) */
    int32_t dimo, dimg;
    bpo_sim sim;
    int8_t ag[BPO_MAX_DIMG];   /* self.achieved_goal, fetch_env.py:78 */
    int8_t goal[BPO_MAX_DIMG]; /* self.goal */
    int32_t colors[BPO_MAX_OBJS];
    int32_t num_objs;          /* self.num_objs, fetch_env.py:75 */
    int32_t has_succeeded;     /* fetch_env.py:80 */
    int32_t t;                 /* gym TimeLimit._elapsed_steps [upstream] */
    /* curriculum knobs, fetch_env.py:340-348,404-415,561-563 (python floats = double) */
    double obj_range, obj_range_step, max_obj_range, wrong_obj_range, wrong_obj_range_step;
    int32_t has_curriculum_step; /* 0 for non-curriculum Choose: no obj_range_step attribute */
    int32_t difficulty;
    /* Philox replay of self.np_random (stream 0) and the global np.random (stream 1) */
    uint64_t seed;
    uint32_t episode;
    uint32_t draws[2];
    uint32_t invalid_actions;  /* count of non-finite action components seen */
    int32_t challenge;         /* BlocksTouchChooseEnv(challenge=True), fetch_env.py:403,416 */
} bpo_env;

/* geometry of the env ids */
int bpo_env_dimo(int env_id);
int bpo_env_dimg(int env_id);
int bpo_env_nblocks(int env_id);

/* Philox4x32-10, counter (c0,c1,c2,c3), key (k0,k1) */
void bpo_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t k0, uint32_t k1, uint32_t out[4]);
float bpo_u01(uint32_t x);          /* (x>>8) * 2^-24  in [0,1) */
float bpo_u01_open(uint32_t x);     /* ((x>>9)+0.5) * 2^-23 in (0,1) */
float bpo_log(float x);             /* spec'd log for x in (0,1) */
void bpo_sincos2pi(float u, float* s, float* c);
float bpo_atan2(float s, float c);  /* spec'd atan2 */
void bpo_normal2(uint32_t w0, uint32_t w1, float* z0, float* z1); /* Box-Muller */

/* sim-level API (what mujoco_py offers the reference) */
void bpo_sim_init(bpo_sim* sim, int env_id);       /* initial_state, robot_env.py:35 */
void bpo_sim_set_action(bpo_sim* sim, const float a[4]); /* fetch_env.py:170-185 (after clip) */
/* the same targets from what upstream utils.mocap_set_action / ctrl_set_action write (float64 mocap_pos, ctrl) */
void bpo_sim_set_targets(bpo_sim* sim, const double mocap_pos[3], const double ctrl[2]);
void bpo_sim_step(bpo_sim* sim);                   /* 20 substeps, robot_env.py:60 */
int bpo_pair_index(int o1, int o2);

/* env-level API */
void bpo_env_init(bpo_env* env, int env_id);
void bpo_env_seed(bpo_env* env, uint64_t seed);    /* robot_env.py:53-55 */
void bpo_env_reset(bpo_env* env, float* obs, float* ag, float* g);  /* robot_env.py:71-82 */
/* returns done (TimeLimit); reward and success via pointers. robot_env.py:57-69 */
int bpo_env_step(bpo_env* env, const float action[4], float* obs, float* ag, float* g,
                 float* reward, int* is_success);
int bpo_env_set_test(bpo_env* env, float* obs, float* ag, float* g);   /* 0 ok, -1 NotImplemented */
int bpo_env_increase_difficulty(bpo_env* env);     /* 1 max reached, 0 not, -1 NotImplementedError, -2 AttributeError */
int bpo_env_get_difficulty(const bpo_env* env);
double bpo_env_get_obj_range(const bpo_env* env);
void bpo_env_set_ranges(bpo_env* env, double obj_range, double wrong_obj_range); /* test hook, mirrors bp_set_ranges */
int bpo_env_set_challenge(bpo_env* env, int challenge); /* BlocksTouchChooseEnv(challenge=...), fetch_env.py:403,416; -1 for other ids */
void bpo_env_get_obs(const bpo_env* env, float* obs, float* ag, float* g);
void bpo_env_get_state(const bpo_env* env, bpo_env_state* out);
void bpo_env_set_state(bpo_env* env, const bpo_env_state* in);
void bpo_env_random_action(const bpo_env* env, float a[4]); /* stream 2, counter (t, episode-1) */

/* fetch_env.py:135-143, batched over n rows */
void bpo_compute_reward(const float* ag, const float* g, int64_t n, int dimg, float* r);

/* HER relabel (baselines.her.her._sample_her_transitions [upstream, recalled]),
 * Philox stream 3.  ag: [B][T+1][dimg], g: [B][T][dimg]. */
void bpo_her_relabel(const float* ep_ag, const float* ep_g, int32_t B, int32_t T, int32_t dimg,
                     int64_t n, float future_p, uint64_t seed, int64_t index_offset,
                     int32_t* ep_idx, int32_t* t_idx, int32_t* fut_t, float* ag2_out,
                     float* g_out, float* r_out);

/* vectorised helpers: loop (optionally OpenMP) over an array of envs */
void bpo_vec_init(bpo_env* envs, int64_t n, int env_id, uint64_t seed, uint64_t env_index_offset);
void bpo_vec_reset(bpo_env* envs, int64_t n, float* obs, float* ag, float* g);
void bpo_vec_step(bpo_env* envs, int64_t n, const float* actions, int auto_reset, float* obs,
                  float* ag, float* reward, float* success, float* reset_obs, float* reset_ag,
                  int32_t* stats /* [4]: episodes, successes at T, steps, invalid */);
void bpo_vec_step_random(bpo_env* envs, int64_t n, int steps, int nthreads, int64_t* stats);
int64_t bpo_sizeof_env(void);
int64_t bpo_sizeof_state(void);

#ifdef __cplusplus
}
#endif
#endif
