/*
 * blockphys_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 * See blockphys_oracle.h for the parity status ("parity unpinned" for the
 * MuJoCo slot) and DESIGN.md section 3 for the BlockPhys v2 specification.
 *
 * Reference files restated here (paths relative to /root/reference/gym_blocks):
 *   envs/robot_env.py:53-82      seed / step / reset order of operations
 *   envs/fetch_env.py:19-37      geometry constants, out_of_table, one_hot_color
 *   envs/fetch_env.py:106-124    geom -> object id map, symmetry check
 *   envs/fetch_env.py:135-143    compute_reward
 *   envs/fetch_env.py:148-167    _step_callback (touch matrix update)
 *   envs/fetch_env.py:170-185    _set_action
 *   envs/fetch_env.py:187-228    _get_obs           (:567-621 Variation)
 *   envs/fetch_env.py:247-281    _reset_sim, _sample_goal, _is_success
 *   envs/fetch_env.py:328-336, 370-399, 448-517, 646-764, 777-787  spawn samplers
 *   envs/fetch_env.py:340-358, 404-432, 623-630   curriculum
 *   envs/tasks.py:4-144          per-id constructor constants
 *   __init__.py:6-53             ids, max_episode_steps = 50
 */
#include "blockphys_oracle.h"
#include "blockphys_tables.h"

#include <math.h>
#include <string.h>

/* ======================================================================
 * BlockPhys constants
 * ====================================================================== */
#define NSUB 20          /* tasks.py:16 n_substeps */
#define H 0.002f         /* 2blocks.xml:4 timestep */
#define INV_H 500.0f
#define DT 0.04f         /* fetch_env.py:190 dt = nsubsteps * timestep */
#define GH 0.01962f      /* g*h = 9.81*0.002 */
#define FR 0.01962f      /* mu*g*h, mu = 1 (MuJoCo default sliding friction) */
#define FRW 0.9f         /* torsional friction, rad/s per substep */
#define KW 2500.0f       /* weld tracking stiffness: critically damped, time constant 0.02 s (shared.xml:49 solref) */
#define BW 100.0f
#define KF 288.46155f    /* finger actuator kp / (armature+mass) = 30000/104 (2blocks.xml:40, shared.xml:62, robot.xml:85) */
#define BF 9.615385f     /* finger joint damping / 104 (shared.xml:62) */
#define QMAX 0.05f       /* finger joint range (robot.xml:86) */
#define CTRL_MAX 0.2f    /* actuator ctrlrange (2blocks.xml:40) */
#define TBL_X 1.3f       /* table centre: body pos 0.25 0.35 0.23 + slides 1.05 0.4 0 (2blocks.xml:18, tasks.py:28-30) */
#define TBL_Y 0.75f
#define TBL_HX 0.25f     /* table half extents (2blocks.xml:19) */
#define TBL_HY 0.35f
#define HB 0.025f        /* cube half size (2blocks.xml:26) */
#define TWO_HB 0.05f
#define Z_REST 0.485f    /* table top 0.46 + HB */
#define Z_FLOOR 0.025f   /* floor plane z = 0 + HB */
#define FX 0.0135f       /* finger box half extents in the world frame (robot.xml:87 rotated by the fixed quat 1 0 1 0) */
#define FY 0.007f
#define FZ 0.0385f
#define FY0 0.0079f      /* finger centre |y| offset at q = 0: 0.0159 - 0.008 (robot.xml:84-87) */
#define FZ_OFF 0.02f     /* finger centre is 0.02 above the grip site (robot.xml:95) */
#define GZ_MIN 0.4785f   /* grip z at which the finger bottoms touch the table top: 0.46 + FZ - FZ_OFF */
#define MARGIN 0.001f    /* contact detection margin (shared.xml:58 geom margin) */
#define DEPEN 0.0005f    /* in-plane contacts never separate two bodies faster than DEPEN/h = 0.25 m/s */
#define VMAX 5.0f        /* safety clamps on derived velocities */
#define WMAX 60.0f
#define IINV 2400.0f     /* inverse yaw inertia / inverse mass of a cube, 6/a^2, a = 0.05 */
#define POS_SCALE 0.05f  /* fetch_env.py:175 */
#define WS_XLO 1.0f      /* arm reach box for the mocap target (BlockPhys) */
#define WS_XHI 1.6f
#define WS_YLO 0.35f
#define WS_YHI 1.15f
#define WS_ZHI 0.9f
#define GRIP0_X 1.3419f  /* initial_gripper_xpos, fetch_env.py:300-301 (pinned constant) */
#define GRIP0_Y 0.7491f
#define GRIP0_Z 0.5347f

/* spawn geometry, fetch_env.py:19-32: python floats (binary64), written as the reference's own
 * expressions so that the compiler folds them to the same doubles (TABLE_H is 0.32499999999999996,
 * MIN_BLOCK_DIST 0.07500000000000001); tests/test_ref_pin.py compares their bits with the module's. */
#define BLOCK_SIZE 0.05
#define MIN_BLOCK_DIST (1.5 * BLOCK_SIZE)
#define TABLE_X (1.05 + 0.25)
#define TABLE_Y (0.40 + 0.35)
#define TABLE_W (0.25 - BLOCK_SIZE / 2)
#define TABLE_H (0.35 - BLOCK_SIZE / 2)
#define MAX_SPAWN_ATTEMPTS 10000 /* BlockPhys cap; the reference loops without bound */

static const int k_nblocks[BPO_NUM_ENV_IDS] = {1, 2, 4, 2, 3, 3, 4};
static const int k_dimo[BPO_NUM_ENV_IDS] = {25, 40, 70, 40, 55, 55, 87};
static const int k_dimg[BPO_NUM_ENV_IDS] = {9, 16, 36, 16, 25, 25, 36};
/* block_gripper per id, tasks.py:15,36,56,77,98,120,142 */
static const int k_block_gripper[BPO_NUM_ENV_IDS] = {0, 0, 0, 1, 1, 1, 1};

int bpo_env_dimo(int id) { return k_dimo[id]; }
int bpo_env_dimg(int id) { return k_dimg[id]; }
int bpo_env_nblocks(int id) { return k_nblocks[id]; }
int64_t bpo_sizeof_env(void) { return (int64_t)sizeof(bpo_env); }
int64_t bpo_sizeof_state(void) { return (int64_t)sizeof(bpo_env_state); }

/* ======================================================================
 * Philox4x32-10 and the spec'd elementary functions
 * ====================================================================== */
void bpo_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

float bpo_u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }
float bpo_u01_open(uint32_t x) { return ((float)(x >> 9) + 0.5f) * 1.1920928955078125e-07f; }

static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* log(x), x in (0, 1]: x = m * 2^e, m in [sqrt(1/2), sqrt(2)); log(m) by the
 * classic s = f/(2+f) series with 4 terms.  Only determinism matters (it feeds
 * the Box-Muller radius, which cancels in the direction normalisation). */
float bpo_log(float x) {
    uint32_t ix = as_uint(x);
    ix += 0x3f800000u - 0x3f3504f3u;
    int32_t e = (int32_t)(ix >> 23) - 127;
    ix = (ix & 0x007fffffu) + 0x3f3504f3u;
    float f = as_float(ix) - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float w = z * z;
    float t1 = w * (0.40000972152f + w * 0.24279078841f);
    float t2 = z * (0.66666662693f + w * 0.28498786688f);
    float R = t2 + t1;
    float hfsq = 0.5f * f * f;
    float dk = (float)e;
    return ((s * (hfsq + R) + dk * 9.0580006145e-06f) - hfsq + f) + dk * 6.9313812256e-01f;
}

/* sin and cos of 2*pi*u, u in [0,1): quadrant reduction + Taylor on [-pi/4, pi/4] */
void bpo_sincos2pi(float u, float* sn, float* cs) {
    float t = u * 4.0f;
    int k = (int)(t + 0.5f);
    float f = t - (float)k;
    float x = f * 1.57079637f;
    float x2 = x * x;
    float sp = x + x * x2 * (-0.16666667f + x2 * (0.0083333338f + x2 * (-0.00019841270f)));
    float cp = 1.0f + x2 * (-0.5f + x2 * (0.041666668f + x2 * (-0.0013888889f + x2 * 2.4801588e-05f)));
    switch (k & 3) {
        case 0: *sn = sp; *cs = cp; break;
        case 1: *sn = cp; *cs = -sp; break;
        case 2: *sn = -sp; *cs = -cp; break;
        default: *sn = -cp; *cs = sp; break;
    }
}

/* atan2(s, c): octant reduction, then tan(pi/8) split and a 7-term odd Taylor series */
float bpo_atan2(float s, float c) {
    float as = fabsf(s), ac = fabsf(c);
    float mx = as > ac ? as : ac;
    float mn = as > ac ? ac : as;
    if (mx == 0.0f) return 0.0f;
    float a = mn / mx;
    float off = 0.0f;
    if (a > 0.41421357f) {
        a = (a - 1.0f) / (a + 1.0f);
        off = 0.78539819f;
    }
    float a2 = a * a;
    float p = 0.076923080f;              /* 1/13 */
    p = -0.090909094f + a2 * p;          /* -1/11 */
    p = 0.11111111f + a2 * p;            /* 1/9 */
    p = -0.14285715f + a2 * p;           /* -1/7 */
    p = 0.2f + a2 * p;                   /* 1/5 */
    p = -0.33333334f + a2 * p;           /* -1/3 */
    float r = off + (a + a * a2 * p);
    if (as > ac) r = 1.57079637f - r;
    if (c < 0.0f) r = 3.14159274f - r;
    if (s < 0.0f) r = -r;
    return r;
}

/* np.random.normal(size=2) replacement: Box-Muller on two Philox words */
void bpo_normal2(uint32_t w0, uint32_t w1, float* z0, float* z1) {
    float u1 = bpo_u01_open(w0);
    float u2 = bpo_u01(w1);
    float r = sqrtf(-2.0f * bpo_log(u1));
    float sn, cs;
    bpo_sincos2pi(u2, &sn, &cs);
    *z0 = r * cs;
    *z1 = r * sn;
}

/* ======================================================================
 * BlockPhys v2: the sim.step() slot
 * ====================================================================== */
int bpo_pair_index(int o1, int o2) {
    if (o1 > o2) { int t = o1; o1 = o2; o2 = t; }
    return o1 * (2 * BPO_MAX_OBJS - o1 - 1) / 2 + (o2 - o1 - 1);
}

static void sim_contact(bpo_sim* sim, int o1, int o2) {
    sim->contacts |= 1u << bpo_pair_index(o1, o2);
}

/* initial_state (robot_env.py:35): what _env_setup (fetch_env.py:283-302) leaves
 * after its 10 settling steps -- gripper at initial_gripper_xpos with closed
 * fingers, cubes at rest on the table top (ToppleTower: stacked). */
void bpo_sim_init(bpo_sim* sim, int env_id) {
    memset(sim, 0, sizeof(*sim));
    sim->g[0] = GRIP0_X; sim->g[1] = GRIP0_Y; sim->g[2] = GRIP0_Z;
    sim->m[0] = GRIP0_X; sim->m[1] = GRIP0_Y; sim->m[2] = GRIP0_Z;
    sim->nblocks = k_nblocks[env_id];
    sim->block_gripper = k_block_gripper[env_id];
    /* initial xy from tasks.py initial_qpos (always overwritten by the spawn samplers) */
    static const float xy0[BPO_NUM_ENV_IDS][4][2] = {
        {{1.25f, 0.55f}},
        {{1.25f, 0.55f}, {1.25f, 0.85f}},
        {{1.25f, 0.6f}, {1.25f, 0.6f}, {1.25f, 0.6f}, {1.25f, 0.6f}},
        {{1.25f, 0.55f}, {1.25f, 0.58f}},
        {{1.25f, 0.55f}, {1.25f, 0.6f}, {1.25f, 0.65f}},
        {{1.275f, 0.6f}, {1.3f, 0.75f}, {1.075f, 0.425f}},
        {{1.275f, 0.6f}, {1.3f, 0.75f}, {1.075f, 0.425f}, {1.3f, 0.425f}},
    };
    float z = Z_REST;
    for (int i = 0; i < BPO_MAX_BLOCKS; ++i) {
        bpo_block* b = &sim->blk[i];
        b->c = 1.0f;
        if (i < sim->nblocks) {
            b->pos[0] = xy0[env_id][i][0];
            b->pos[1] = xy0[env_id][i][1];
            b->pos[2] = (env_id == BPO_TOPPLE_TOWER) ? z : Z_REST;
            z = z + TWO_HB;
        }
    }
}

static float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
/* clamp between two NONZERO bounds as min(max(x, lo), hi) (IEEE minNum / maxNum: a NaN becomes lo).  Same value as
 * clampf for every non-NaN x -- no signed-zero case since neither bound is zero; two FMNMX instead of four
 * compare / select instructions in the CUDA library (BlockPhys v2: velocity caps and the mocap reach box). */
static float clampnz(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
/* BlockPhys v1.1+: the fused multiply-adds of the spec are written explicitly (single rounding) */
#define F(a, b, c) fmaf((a), (b), (c))

/* What MuJoCo is handed by gym.envs.robotics.utils.mocap_set_action / ctrl_set_action [upstream]:
 * mocap_pos and the two finger-actuator controls as float64.  BlockPhys narrows them to binary32,
 * keeps the mocap target inside the arm-reach box and the controls inside ctrlrange (2blocks.xml:40). */
void bpo_sim_set_targets(bpo_sim* sim, const double mocap_pos[3], const double ctrl[2]) {
    sim->m[0] = clampnz((float)mocap_pos[0], WS_XLO, WS_XHI);
    sim->m[1] = clampnz((float)mocap_pos[1], WS_YLO, WS_YHI);
    sim->m[2] = clampnz((float)mocap_pos[2], GZ_MIN, WS_ZHI);
    sim->ctrl[0] = clampf((float)ctrl[0], 0.0f, CTRL_MAX);
    sim->ctrl[1] = clampf((float)ctrl[1], 0.0f, CTRL_MAX);
}

/* fetch_env.py:170-185 + gym.envs.robotics.utils.ctrl_set_action / mocap_set_action
 * [upstream, recalled]: pos_ctrl = a[:3] * 0.05 in the action's float32 (:175), the mocap is
 * snapped to the welded body and moved by pos_ctrl; finger position-actuator target = qpos + a[3].
 * v1.3: the product and the sum are rounded separately, which is what the reference's float32
 * multiply followed by the float64 mocap_pos + pos_delta narrows to (v1.1-1.2 fused them). */
void bpo_sim_set_action(bpo_sim* sim, const float a[4]) {
    sim->m[0] = clampnz(sim->g[0] + a[0] * POS_SCALE, WS_XLO, WS_XHI);
    sim->m[1] = clampnz(sim->g[1] + a[1] * POS_SCALE, WS_YLO, WS_YHI);
    sim->m[2] = clampnz(sim->g[2] + a[2] * POS_SCALE, GZ_MIN, WS_ZHI);
    float ga = sim->block_gripper ? 0.0f : a[3]; /* fetch_env.py:179-180 */
    sim->ctrl[0] = clampf(sim->q[0] + ga, 0.0f, CTRL_MAX);
    sim->ctrl[1] = clampf(sim->q[1] + ga, 0.0f, CTRL_MAX);
}

/* first-order rotation of the unit complex number (c, s) by dth, then one Newton step of the inverse
 * square root at 1 as renormalisation (BlockPhys v1.2): |(c2, s2)|^2 = n2 = 1 + dth^2 up to the previous
 * step's residue, r = 1.5 - 0.5 n2 = n2^(-1/2) + O((n2 - 1)^2).  The norm stays within 3/8 dth^4 of 1
 * (< 8e-5 at the 60 rad/s cap) and the iteration is self-correcting; v1.1 used sqrtf and a reciprocal here. */
static void rot_apply(bpo_block* b, float dth) {
    float c2 = F(-b->s, dth, b->c);
    float s2 = F(b->c, dth, b->s);
    float n2 = F(c2, c2, s2 * s2);
    float r = F(-0.5f, n2, 1.5f);
    b->c = c2 * r;
    b->s = s2 * r;
}

typedef struct { float x, y, c, s, hx, hy; } rect2;

/* 2-D separating-axis test of two rectangles.  Axes: 0 = A.u, 1 = A.v, 2 = B.u, 3 = B.v. */
typedef struct {
    float ov[4];    /* overlap along each axis */
    float proj[4];  /* d . axis, d = B - A */
    float rtA[4];   /* half extent of A along the tangent of axis k */
    float rtB[4];   /* half extent of B along the tangent of axis k */
} sat2;

static void sat2_eval(const rect2* A, const rect2* B, sat2* o) {
    float cr = F(A->c, B->c, A->s * B->s);
    float sr = F(A->c, B->s, -(A->s * B->c));
    float C = fabsf(cr), S = fabsf(sr);
    float dx = B->x - A->x, dy = B->y - A->y;
    float RBu = F(B->hx, C, B->hy * S); /* B along A.u */
    float RBv = F(B->hx, S, B->hy * C); /* B along A.v */
    float RAu = F(A->hx, C, A->hy * S); /* A along B.u */
    float RAv = F(A->hx, S, A->hy * C); /* A along B.v */
    o->proj[0] = F(dx, A->c, dy * A->s);
    o->proj[1] = F(dy, A->c, -(dx * A->s));
    o->proj[2] = F(dx, B->c, dy * B->s);
    o->proj[3] = F(dy, B->c, -(dx * B->s));
    o->ov[0] = (A->hx + RBu) - fabsf(o->proj[0]);
    o->ov[1] = (A->hy + RBv) - fabsf(o->proj[1]);
    o->ov[2] = (RAu + B->hx) - fabsf(o->proj[2]);
    o->ov[3] = (RAv + B->hy) - fabsf(o->proj[3]);
    o->rtA[0] = A->hy; o->rtB[0] = RBv;
    o->rtA[1] = A->hx; o->rtB[1] = RBu;
    o->rtA[2] = RAv;   o->rtB[2] = B->hy;
    o->rtA[3] = RAu;   o->rtB[3] = B->hx;
}

static int sat2_argmin(const sat2* o) {
    int k = 0;
    for (int i = 1; i < 4; ++i)
        if (o->ov[i] < o->ov[k]) k = i;
    return k;
}

/* Contact geometry for in-plane axis k: unit normal n (A -> B) and the torque
 * arms rnA, rnB (z of r x n) of the contact centre, taken as the midpoint of the
 * overlap of the two tangent-projection intervals. */
static void sat2_contact(const rect2* A, const rect2* B, const sat2* o, int k,
                         float* nx, float* ny, float* rnA, float* rnB) {
    float sg = o->proj[k] >= 0.0f ? 1.0f : -1.0f;
    float ex, ey;
    switch (k) {
        case 0: ex = A->c; ey = A->s; break;
        case 1: ex = -A->s; ey = A->c; break;
        case 2: ex = B->c; ey = B->s; break;
        default: ex = -B->s; ey = B->c; break;
    }
    *nx = sg * ex;
    *ny = sg * ey;
    float dt = o->proj[k ^ 1];
    float lo = -o->rtA[k];
    float lo2 = dt - o->rtB[k];
    if (lo2 > lo) lo = lo2;
    float hi = o->rtA[k];
    float hi2 = dt + o->rtB[k];
    if (hi2 < hi) hi = hi2;
    float mid = 0.5f * (lo + hi);
    float chi = (k & 1) ? sg : -sg; /* sigma * (T x E) */
    *rnA = chi * mid;
    *rnB = chi * (mid - dt);
}

typedef struct {
    float old[3];
    float dth;
    int supported;
} blk_tmp;

/* per-substep scratch: cube start-of-substep data and the gripper/finger start positions */
typedef struct {
    blk_tmp b[BPO_MAX_BLOCKS];
    float g_old[3];
    float q_old[2];
    float closed[2];
} sub_tmp;

/* BlockPhys v2: the env-step's propagator base.  The weld and the finger actuators are linear, so until a contact
 * acts on a channel its state after n substeps is the n-th power of the substep map (blockphys_tables.h) applied to
 * the state the env-step started from: d0 / e0 = position - target, v0 / w0 = velocity.  x and y are never acted on.
 * The z channel leaves this "free" mode when a finger lands on a cube, a finger channel when its closing is undone
 * (collide_finger_block); from then on that channel advances by the substep recurrence of v1. */
typedef struct {
    float d0[3], v0[3]; /* gripper */
    float e0[2], w0[2]; /* fingers */
    int zfree, qfree[2];
    int n;              /* current substep, 1..NSUB */
} step_base;

static void finger_rect(const bpo_sim* sim, int f, rect2* r, float* z) {
    float sgn = f == 0 ? 1.0f : -1.0f;
    r->x = sim->g[0];
    r->y = sim->g[1] + sgn * (FY0 + sim->q[f]);
    r->c = 1.0f; r->s = 0.0f;
    r->hx = FX; r->hy = FY;
    *z = sim->g[2] + FZ_OFF;
}

static void collide_finger_block(bpo_sim* sim, int f, int bi, sub_tmp* st, step_base* sb) {
    float* closed = st->closed;
    blk_tmp* tmp = st->b;
    bpo_block* b = &sim->blk[bi];
    rect2 A, B;
    float az;
    finger_rect(sim, f, &A, &az);
    B.x = b->pos[0]; B.y = b->pos[1]; B.c = b->c; B.s = b->s; B.hx = HB; B.hy = HB;
    float dz = b->pos[2] - az;
    float ovz = (FZ + HB) - fabsf(dz);
    sat2 o;
    sat2_eval(&A, &B, &o);
    int k = sat2_argmin(&o);
    float minxy = o.ov[k];
    float minov = ovz < minxy ? ovz : minxy;
    if (!(minov > -MARGIN)) return;
    sim_contact(sim, 0, bi + 2); /* any "finger" geom -> object 0, fetch_env.py:111-112 */
    if (!(minov > 0.0f)) return;
    if (ovz <= minxy) {
        if (dz >= 0.0f) { /* block on top of the finger */
            b->pos[2] = az + (FZ + HB);
            tmp[bi].supported = 1;
        } else {          /* finger rests on the block: the gripper yields upward */
            sim->g[2] = (b->pos[2] + (FZ + HB)) - FZ_OFF;
            if (sb->zfree) { /* v2: the z channel leaves the propagator with the velocity it has there now */
                sim->gv[2] = F(BPO_GC[sb->n], sb->d0[2], BPO_GD[sb->n] * sb->v0[2]);
                sb->zfree = 0;
            }
            if (sim->gv[2] < 0.0f) sim->gv[2] = 0.0f;
        }
        return;
    }
    float delta = minxy;
    if (k == 1) {
        /* contact along the finger's closing axis on its inner face: first undo
         * this substep's closing motion of that finger (closing never pushes) */
        int inner = (f == 0) ? (o.proj[1] < 0.0f) : (o.proj[1] > 0.0f);
        if (inner) {
            float yield = delta < closed[f] ? delta : closed[f];
            float room = QMAX - sim->q[f];
            if (yield > room) yield = room;
            if (yield > 0.0f) {
                sim->q[f] = sim->q[f] + yield;
                sim->qv[f] = 0.0f;
                sb->qfree[f] = 0; /* v2: this finger advances by the recurrence for the rest of the env-step */
                closed[f] = closed[f] - yield;
                delta = delta - yield;
            }
            if (!(delta > 0.0f)) return;
        }
    }
    float nx, ny, rnA, rnB;
    sat2_contact(&A, &B, &o, k, &nx, &ny, &rnA, &rnB);
    (void)rnA;
    /* separation already under way this substep along n (cube minus finger displacement) */
    float sgn = f == 0 ? 1.0f : -1.0f;
    float fdx = sim->g[0] - st->g_old[0];
    float fdy = (sim->g[1] - st->g_old[1]) + sgn * (sim->q[f] - st->q_old[f]);
    float rel = F((b->pos[0] - tmp[bi].old[0]) - fdx, nx, ((b->pos[1] - tmp[bi].old[1]) - fdy) * ny);
    float cap = DEPEN - rel;
    float lam = delta < cap ? delta : cap;
    if (!(lam > 0.0f)) return;
    float D = F(IINV, rnB * rnB, 1.0f);
    float l = lam / D;
    b->pos[0] = F(nx, l, b->pos[0]);
    b->pos[1] = F(ny, l, b->pos[1]);
    float dth = (IINV * rnB) * l;
    if (dth != 0.0f) {
        rot_apply(b, dth);
        tmp[bi].dth = tmp[bi].dth + dth;
    }
}

static void collide_block_block(bpo_sim* sim, int i, int j, blk_tmp* tmp) {
    bpo_block* a = &sim->blk[i];
    bpo_block* b = &sim->blk[j];
    rect2 A, B;
    A.x = a->pos[0]; A.y = a->pos[1]; A.c = a->c; A.s = a->s; A.hx = HB; A.hy = HB;
    B.x = b->pos[0]; B.y = b->pos[1]; B.c = b->c; B.s = b->s; B.hx = HB; B.hy = HB;
    float dz = b->pos[2] - a->pos[2];
    float ovz = TWO_HB - fabsf(dz);
    sat2 o;
    sat2_eval(&A, &B, &o);
    int k = sat2_argmin(&o);
    float minxy = o.ov[k];
    float minov = ovz < minxy ? ovz : minxy;
    if (!(minov > -MARGIN)) return;
    sim_contact(sim, i + 2, j + 2); /* "objectK" -> K + 2, fetch_env.py:115-116 */
    if (!(minov > 0.0f)) return;
    int pin = 0; /* 1: A is pinned (B slides off it), 2: B is pinned */
    if (ovz <= minxy) {
        /* vertical contact: the upper cube is carried iff its centre is over the lower footprint */
        if (dz >= 0.0f) {
            if (fabsf(o.proj[0]) <= HB && fabsf(o.proj[1]) <= HB) {
                b->pos[2] = a->pos[2] + TWO_HB;
                tmp[j].supported = 1;
                return;
            }
            pin = 1;
        } else {
            if (fabsf(o.proj[2]) <= HB && fabsf(o.proj[3]) <= HB) {
                a->pos[2] = b->pos[2] + TWO_HB;
                tmp[i].supported = 1;
                return;
            }
            pin = 2;
        }
        /* overhanging: the upper cube slides off the (pinned) lower one through the in-plane axis */
    }
    float nx, ny, rnA, rnB;
    sat2_contact(&A, &B, &o, k, &nx, &ny, &rnA, &rnB);
    float rel = F((b->pos[0] - tmp[j].old[0]) - (a->pos[0] - tmp[i].old[0]), nx,
                  ((b->pos[1] - tmp[j].old[1]) - (a->pos[1] - tmp[i].old[1])) * ny);
    float cap = DEPEN - rel;
    float lam = minxy < cap ? minxy : cap;
    if (!(lam > 0.0f)) return;
    float wA = pin == 1 ? 0.0f : 1.0f;
    float wB = pin == 2 ? 0.0f : 1.0f;
    float D = F(IINV, F(wA, rnA * rnA, wB * (rnB * rnB)), wA + wB);
    float l = lam / D;
    float lA = wA * l, lB = wB * l;
    a->pos[0] = F(-nx, lA, a->pos[0]);
    a->pos[1] = F(-ny, lA, a->pos[1]);
    b->pos[0] = F(nx, lB, b->pos[0]);
    b->pos[1] = F(ny, lB, b->pos[1]);
    float dthA = -((IINV * rnA) * lA);
    float dthB = (IINV * rnB) * lB;
    if (dthA != 0.0f) { rot_apply(a, dthA); tmp[i].dth = tmp[i].dth + dthA; }
    if (dthB != 0.0f) { rot_apply(b, dthB); tmp[j].dth = tmp[j].dth + dthB; }
}

static int over_table(float x, float y) {
    return fabsf(x - TBL_X) <= TBL_HX && fabsf(y - TBL_Y) <= TBL_HY;
}

static void sim_substep(bpo_sim* sim, step_base* sb) {
    sub_tmp st;
    blk_tmp* tmp = st.b;
    float* closed = st.closed;
    const int n = sb->n;
    closed[0] = closed[1] = 0.0f;
    st.g_old[0] = sim->g[0]; st.g_old[1] = sim->g[1]; st.g_old[2] = sim->g[2];
    st.q_old[0] = sim->q[0]; st.q_old[1] = sim->q[1];
    /* 1. gripper: critically damped tracking of the mocap target (weld, shared.xml:48-50).
     * v2: x and y are the propagator applied to the env-step's start state (their velocities are only needed at the
     * end of the env-step, bpo_sim_step); z likewise while it is free, projected onto z >= GZ_MIN (the finger
     * bottoms on the table: a position projection inside the env-step, the velocity rule is applied at its end). */
    for (int k = 0; k < 2; ++k) sim->g[k] = F(BPO_GA[n], sb->d0[k], F(BPO_GB[n], sb->v0[k], sim->m[k]));
    if (sb->zfree) {
        float zv = F(BPO_GA[n], sb->d0[2], F(BPO_GB[n], sb->v0[2], sim->m[2]));
        sim->g[2] = zv < GZ_MIN ? GZ_MIN : zv;
    } else {
        float acc = F(KW, sim->m[2] - sim->g[2], -(BW * sim->gv[2]));
        sim->gv[2] = F(acc, H, sim->gv[2]);
        sim->g[2] = F(sim->gv[2], H, sim->g[2]);
        if (sim->g[2] < GZ_MIN) {
            sim->g[2] = GZ_MIN;
            if (sim->gv[2] < 0.0f) sim->gv[2] = 0.0f;
        }
    }
    /* 2. fingers: position actuators (2blocks.xml:39-42); block_gripper pins them at 0 (fetch_env.py:149-152).
     * v2: free fingers follow the propagator, projected onto the joint range [0, QMAX]. */
    if (!sim->block_gripper) {
        for (int f = 0; f < 2; ++f) {
            float q_old = sim->q[f];
            if (sb->qfree[f]) {
                float qv = F(BPO_FA[n], sb->e0[f], F(BPO_FB[n], sb->w0[f], sim->ctrl[f]));
                sim->q[f] = clampf(qv, 0.0f, QMAX);
            } else {
                float acc = F(KF, sim->ctrl[f] - sim->q[f], -(BF * sim->qv[f]));
                sim->qv[f] = F(acc, H, sim->qv[f]);
                sim->q[f] = F(sim->qv[f], H, sim->q[f]);
                if (sim->q[f] < 0.0f) { sim->q[f] = 0.0f; if (sim->qv[f] < 0.0f) sim->qv[f] = 0.0f; }
                if (sim->q[f] > QMAX) { sim->q[f] = QMAX; if (sim->qv[f] > 0.0f) sim->qv[f] = 0.0f; }
            }
            float cl = q_old - sim->q[f];
            closed[f] = cl > 0.0f ? cl : 0.0f;
        }
    }
    /* 3. predict cubes */
    for (int i = 0; i < sim->nblocks; ++i) {
        bpo_block* b = &sim->blk[i];
        tmp[i].old[0] = b->pos[0]; tmp[i].old[1] = b->pos[1]; tmp[i].old[2] = b->pos[2];
        tmp[i].dth = 0.0f;
        tmp[i].supported = 0;
        b->vel[2] = b->vel[2] - GH;
        b->pos[0] = F(b->vel[0], H, b->pos[0]);
        b->pos[1] = F(b->vel[1], H, b->pos[1]);
        b->pos[2] = F(b->vel[2], H, b->pos[2]);
        if (b->w != 0.0f) {
            float dth = b->w * H;
            rot_apply(b, dth);
            tmp[i].dth = dth;
        }
    }
    sim->contacts = 0;
    /* 4a. table / floor support */
    for (int i = 0; i < sim->nblocks; ++i) {
        bpo_block* b = &sim->blk[i];
        if (over_table(b->pos[0], b->pos[1])) {
            if (b->pos[2] - Z_REST < MARGIN) sim_contact(sim, 1, i + 2); /* "table" -> 1, fetch_env.py:113-114 */
            if (b->pos[2] < Z_REST) { b->pos[2] = Z_REST; tmp[i].supported = 1; }
        } else if (b->pos[2] < Z_FLOOR) {
            b->pos[2] = Z_FLOOR; /* floor0 geom maps to None: no touch entry */
            tmp[i].supported = 1;
        }
    }
    /* 4b. fingers vs cubes */
    for (int i = 0; i < sim->nblocks; ++i)
        for (int f = 0; f < 2; ++f)
            collide_finger_block(sim, f, i, &st, sb);
    /* 4c. cube pairs */
    for (int i = 0; i < sim->nblocks; ++i)
        for (int j = i + 1; j < sim->nblocks; ++j)
            collide_block_block(sim, i, j, tmp);
    /* 4d. fingers vs table */
    if (over_table(sim->g[0], sim->g[1]) && sim->g[2] - GZ_MIN < MARGIN) sim_contact(sim, 0, 1);
    /* 5. velocities from the position change, then Coulomb friction on supported cubes */
    for (int i = 0; i < sim->nblocks; ++i) {
        bpo_block* b = &sim->blk[i];
        b->vel[0] = (b->pos[0] - tmp[i].old[0]) * INV_H;
        b->vel[1] = (b->pos[1] - tmp[i].old[1]) * INV_H;
        b->vel[2] = (b->pos[2] - tmp[i].old[2]) * INV_H;
        b->w = tmp[i].dth * INV_H;
        for (int k = 0; k < 3; ++k) b->vel[k] = clampnz(b->vel[k], -VMAX, VMAX);
        b->w = clampnz(b->w, -WMAX, WMAX);
        if (tmp[i].supported) {
            float sp2 = F(b->vel[0], b->vel[0], b->vel[1] * b->vel[1]);
            if (sp2 <= FR * FR) {
                b->vel[0] = 0.0f; b->vel[1] = 0.0f;
            } else {
                float sp = sqrtf(sp2);
                float kf = (sp - FR) / sp;
                b->vel[0] = b->vel[0] * kf;
                b->vel[1] = b->vel[1] * kf;
            }
            if (fabsf(b->w) <= FRW) b->w = 0.0f;
            else b->w = b->w > 0.0f ? b->w - FRW : b->w + FRW;
        }
    }
}

/* sim.step() (robot_env.py:60): NSUB substeps towards the targets set_action / set_targets left in m, ctrl */
void bpo_sim_step(bpo_sim* sim) {
    step_base sb;
    for (int k = 0; k < 3; ++k) { sb.d0[k] = sim->g[k] - sim->m[k]; sb.v0[k] = sim->gv[k]; }
    for (int f = 0; f < 2; ++f) { sb.e0[f] = sim->q[f] - sim->ctrl[f]; sb.w0[f] = sim->qv[f]; }
    sb.zfree = 1; sb.qfree[0] = sb.qfree[1] = 1;
    for (sb.n = 1; sb.n <= NSUB; ++sb.n) sim_substep(sim, &sb);
    /* end of the env-step: velocities of the channels that stayed free, with the limit rules of v1 applied once */
    for (int k = 0; k < 2; ++k) sim->gv[k] = F(BPO_GC[NSUB], sb.d0[k], BPO_GD[NSUB] * sb.v0[k]);
    if (sb.zfree) {
        float zv = F(BPO_GA[NSUB], sb.d0[2], F(BPO_GB[NSUB], sb.v0[2], sim->m[2]));
        sim->gv[2] = F(BPO_GC[NSUB], sb.d0[2], BPO_GD[NSUB] * sb.v0[2]);
        if (zv < GZ_MIN && sim->gv[2] < 0.0f) sim->gv[2] = 0.0f;
    }
    if (!sim->block_gripper) {
        for (int f = 0; f < 2; ++f) {
            if (!sb.qfree[f]) continue;
            float qv = F(BPO_FA[NSUB], sb.e0[f], F(BPO_FB[NSUB], sb.w0[f], sim->ctrl[f]));
            sim->qv[f] = F(BPO_FC[NSUB], sb.e0[f], BPO_FD[NSUB] * sb.w0[f]);
            if (qv < 0.0f && sim->qv[f] < 0.0f) sim->qv[f] = 0.0f;
            if (qv > QMAX && sim->qv[f] > 0.0f) sim->qv[f] = 0.0f;
        }
    }
}

/* ======================================================================
 * Environment logic (reference-owned; restated line by line)
 * ====================================================================== */

/* _sample_colors: fetch_env.py:323-326, 360-363, 434-441, 632-639, 772-775 */
static void sample_colors(int env_id, int32_t* C) {
    static const int32_t tbl[BPO_NUM_ENV_IDS][BPO_MAX_OBJS] = {
        {BPO_BLUE, BPO_GREY, BPO_GREEN, 0, 0, 0},
        {BPO_GREY, BPO_GREY, BPO_GREEN, BPO_BLUE, 0, 0},
        {BPO_RED, BPO_GREEN, BPO_GREY, BPO_GREY, BPO_GREY, BPO_BLUE},
        {BPO_GREY, BPO_GREY, BPO_GREEN, BPO_BLUE, 0, 0},
        {BPO_GREY, BPO_GREY, BPO_GREEN, BPO_BLUE, BPO_GREY, 0},
        {BPO_GREY, BPO_GREY, BPO_GREEN, BPO_BLUE, BPO_GREY, 0},
        {BPO_GREY, BPO_GREY, BPO_GREEN, BPO_BLUE, BPO_GREY, BPO_GREY},
    };
    memcpy(C, tbl[env_id], sizeof(tbl[0]));
}

/* _sample_goal: fetch_env.py:260-273 (Variation always uses 6 objects, :682-695) */
static void sample_goal(bpo_env* env) {
    int N = env->env_id == BPO_BLOCKS_TOUCH_VARIATION ? BPO_MAX_OBJS : env->num_objs;
    const int32_t* C = env->colors;
    memset(env->goal, 0, sizeof(env->goal));
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            int8_t v = 0;
            if ((C[i] == BPO_RED && C[j] == BPO_BLUE) || (C[i] == BPO_BLUE && C[j] == BPO_RED)) v = -1;
            else if ((C[i] == BPO_GREEN && C[j] == BPO_BLUE) || (C[i] == BPO_BLUE && C[j] == BPO_GREEN)) v = 1;
            env->goal[i * N + j] = v;
        }
}

/* compute_reward: fetch_env.py:135-143 */
void bpo_compute_reward(const float* ag, const float* g, int64_t n, int dimg, float* r) {
    for (int64_t i = 0; i < n; ++i) {
        float d = 0.0f;
        int c = 0;
        for (int k = 0; k < dimg; ++k) {
            d = d + ag[i * dimg + k] * g[i * dimg + k];
            if (g[i * dimg + k] != 0.0f) ++c;
        }
        r[i] = -((d != (float)c) ? 1.0f : 0.0f);
    }
}

static float env_reward(const bpo_env* env) {
    /* the matrix stride of self.achieved_goal: num_objs, except Variation (6x6 always, :663) */
    int N = env->env_id == BPO_BLOCKS_TOUCH_VARIATION ? BPO_MAX_OBJS : env->num_objs;
    float ag[BPO_MAX_DIMG], g[BPO_MAX_DIMG], r;
    for (int k = 0; k < N * N; ++k) { ag[k] = (float)env->ag[k]; g[k] = (float)env->goal[k]; }
    bpo_compute_reward(ag, g, 1, N * N, &r);
    return r;
}

/* ---- Philox replay of the two RNG streams ---- */
static void env_draw(bpo_env* env, int stream, uint32_t out[4]) {
    uint32_t ep = env->episode; /* reset() bumps episode after its draws; see bpo_env_reset */
    bpo_philox4x32(env->draws[stream], ep, (uint32_t)stream, 0u,
                   (uint32_t)env->seed, (uint32_t)(env->seed >> 32), out);
    env->draws[stream] += 1;
}

/* The spawn samplers follow the reference's arithmetic operation by operation in binary64 (v1.3):
 * the env's python/numpy code computes object_xpos in float64 from initial_gripper_xpos (the float64
 * image of the sim's binary32 grip position), the RandomState draws and python-float ranges, and only
 * set_joint_qpos narrows the result into the sim's binary32 qpos.  Rejection tests therefore see the
 * same float64 values as the reference's, and the draw counters cannot drift.
 * RandomState.uniform(low, high) is low + (high - low) * u, u = 24-bit Philox fraction (exact). */
static void rs_uniform2(bpo_env* env, int stream, double lo, double hi, double* a, double* b) {
    uint32_t w[4];
    env_draw(env, stream, w);
    *a = lo + (hi - lo) * (double)bpo_u01(w[0]);
    *b = lo + (hi - lo) * (double)bpo_u01(w[1]);
}
/* uniform(lo, hi) scalar */
static double rs_uniform1(bpo_env* env, int stream, double lo, double hi) {
    uint32_t w[4];
    env_draw(env, stream, w);
    return lo + (hi - lo) * (double)bpo_u01(w[0]);
}
/* np.random.normal(size=2): the spec'd binary32 Box-Muller pair, widened */
static void rs_normal2(bpo_env* env, int stream, double* a, double* b) {
    uint32_t w[4];
    float z0, z1;
    env_draw(env, stream, w);
    bpo_normal2(w[0], w[1], &z0, &z1);
    *a = (double)z0; *b = (double)z1;
}
/* np_random.randint(n) */
static int rs_randint(bpo_env* env, int stream, int n) {
    uint32_t w[4];
    env_draw(env, stream, w);
    return (int)(((uint64_t)w[0] * (uint64_t)n) >> 32);
}

static int out_of_table(double x, double y) { /* fetch_env.py:30-32 */
    return fabs(x - TABLE_X) > TABLE_W || fabs(y - TABLE_Y) > TABLE_H;
}
/* np.linalg.norm of a float64 2-vector = sqrt(x.dot(x)); numpy's dot accumulates the second product
 * with a fused multiply-add (checked against numpy 2.3 in tests/test_ref_pin.py) */
static double norm2(double x, double y) { return sqrt(fma(y, y, x * x)); }

static void set_block_xy(bpo_env* env, int i, double x, double y) { /* object_qpos[:2] = object_xpos; set_joint_qpos */
    env->sim.blk[i].pos[0] = (float)x;
    env->sim.blk[i].pos[1] = (float)y;
}

/* direction = normal(2)/|.|; mag = uniform(lo, hi); xy = base + direction*mag
 * (fetch_env.py:390-393, 490-493, 507-510, 734-737); draws from the GLOBAL np.random -> stream 1 */
static void sample_around(bpo_env* env, double bx, double by, double lo, double hi, double* x, double* y) {
    double d0, d1;
    rs_normal2(env, 1, &d0, &d1);
    double n = norm2(d0, d1);
    d0 = d0 / n; d1 = d1 / n;
    double mag = rs_uniform1(env, 1, lo, hi);
    *x = bx + d0 * mag;
    *y = by + d1 * mag;
}

/* GripperTouchEnv._randomize_objects fetch_env.py:328-336; ToppleTowerEnv :777-787 */
static void randomize_gripper_touch(bpo_env* env, int nset) {
    const double g0x = (double)GRIP0_X, g0y = (double)GRIP0_Y; /* initial_gripper_xpos[:2] */
    double r = env->obj_range;
    double x = g0x, y = g0y;
    int it = 0;
    while (norm2(x - g0x, y - g0y) < 0.1 && it++ < MAX_SPAWN_ATTEMPTS) {
        double u0, u1;
        rs_uniform2(env, 0, -r, r, &u0, &u1);
        x = g0x + u0;
        y = g0y + u1;
    }
    for (int i = 0; i < nset; ++i) set_block_xy(env, i, x, y);
}

/* BlocksTouchEnv._randomize_objects fetch_env.py:370-399 */
static void randomize_blocks_touch(bpo_env* env, int test) {
    const double g0x = (double)GRIP0_X, g0y = (double)GRIP0_Y;
    double r = test ? env->max_obj_range : env->obj_range;
    double u0, u1;
    rs_uniform2(env, 0, -r / 2, r / 2, &u0, &u1);
    double x0 = g0x + u0, y0 = g0y + u1;
    set_block_xy(env, 0, x0, y0);
    double x, y;
    int it = 0;
    do {
        sample_around(env, x0, y0, MIN_BLOCK_DIST, r, &x, &y);
    } while (out_of_table(x, y) && ++it < MAX_SPAWN_ATTEMPTS);
    set_block_xy(env, 1, x, y);
}

/* blue block: loop until on the table (fetch_env.py:475-480, 719-724) */
static void sample_blue(bpo_env* env, double r, double* x, double* y) {
    const double g0x = (double)GRIP0_X, g0y = (double)GRIP0_Y;
    int it = 0;
    do {
        double u0, u1;
        rs_uniform2(env, 0, -r / 2, r / 2, &u0, &u1);
        *x = g0x + u0;
        *y = g0y + u1;
    } while (out_of_table(*x, *y) && ++it < MAX_SPAWN_ATTEMPTS);
}

/* BlocksTouchChooseEnv._randomize_objects fetch_env.py:448-517; `challenge` (:403,416,452-463) is a constructor
 * argument no tasks.py class sets -- bpo_env_set_challenge stands for constructing the env with challenge=True */
static void randomize_choose(bpo_env* env, int test) {
    double r, wrong_r;
    if (test || env->challenge) { r = env->max_obj_range; wrong_r = 0.0; }   /* :452-454 */
    else { r = env->obj_range; wrong_r = env->wrong_obj_range; }
    double min_r = env->challenge ? 0.15 : MIN_BLOCK_DIST;                     /* :458-463 */
    double max_wrong_r = env->challenge ? 0.04 : env->max_obj_range;
    int blue = 1, green = 0, wrong = 2; /* colours [GREEN, BLUE, GREY], :465-473 */
    double bx, by, gx, gy, wx, wy;
    sample_blue(env, r, &bx, &by);
    set_block_xy(env, blue, bx, by);
    int it = 0;
    do {
        sample_around(env, bx, by, min_r, r, &gx, &gy);
    } while (out_of_table(gx, gy) && ++it < MAX_SPAWN_ATTEMPTS);
    set_block_xy(env, green, gx, gy);
    double cx = (bx + gx) / 2.0, cy = (by + gy) / 2.0;
    it = 0;
    int again;
    do {
        sample_around(env, cx, cy, wrong_r, max_wrong_r, &wx, &wy);
        again = out_of_table(wx, wy) || norm2(wx - bx, wy - by) < MIN_BLOCK_DIST ||
                norm2(wx - gx, wy - gy) < MIN_BLOCK_DIST;
    } while (again && ++it < MAX_SPAWN_ATTEMPTS);
    set_block_xy(env, wrong, wx, wy);
}

/* BlocksTouchVariationEnv._randomize_objects fetch_env.py:697-764 */
static void randomize_variation(bpo_env* env, int test) {
    int num_blocks = env->num_objs - 2;
    double r = test ? env->max_obj_range : env->obj_range;
    int blue = 1, green = 0; /* colours [GREEN, BLUE, GREY, GREY], :710-716 */
    double px[BPO_MAX_BLOCKS], py[BPO_MAX_BLOCKS];
    int np = 0;
    double bx, by, gx, gy;
    sample_blue(env, r, &bx, &by);
    set_block_xy(env, blue, bx, by);
    int it = 0;
    do {
        sample_around(env, bx, by, MIN_BLOCK_DIST, r, &gx, &gy);
    } while (out_of_table(gx, gy) && ++it < MAX_SPAWN_ATTEMPTS);
    set_block_xy(env, green, gx, gy);
    px[np] = bx; py[np] = by; ++np;
    px[np] = gx; py[np] = gy; ++np;
    for (int i = 0; i < num_blocks; ++i) {
        if (i == blue || i == green) continue;
        double x, y;
        int again;
        it = 0;
        do {
            /* _sample_from_table fetch_env.py:88-90: two scalar draws from self.np_random */
            double ux = rs_uniform1(env, 0, -TABLE_W, TABLE_W);
            double uy = rs_uniform1(env, 0, -TABLE_H, TABLE_H);
            x = TABLE_X + ux;
            y = TABLE_Y + uy;
            again = 0;
            int hit = 0;
            for (int p = 0; p < np; ++p)
                if (norm2(x - px[p], y - py[p]) < MIN_BLOCK_DIST) { hit = 1; break; }
            if (hit) again = 1;
            else again = out_of_table(x, y); /* the for-else, :753-758 */
        } while (again && ++it < MAX_SPAWN_ATTEMPTS);
        set_block_xy(env, i, x, y);
        px[np] = x; py[np] = y; ++np;
    }
}

static void randomize_objects(bpo_env* env, int test) {
    switch (env->env_id) {
        case BPO_GRIPPER_TOUCH: randomize_gripper_touch(env, 1); break;
        case BPO_TOPPLE_TOWER: randomize_gripper_touch(env, 4); break;
        case BPO_BLOCKS_TOUCH:
        case BPO_BLOCKS_TOUCH_CURRICULUM: randomize_blocks_touch(env, test); break;
        case BPO_BLOCKS_TOUCH_CHOOSE:
        case BPO_BLOCKS_TOUCH_CHOOSE_CURRICULUM: randomize_choose(env, test); break;
        default: randomize_variation(env, test); break;
    }
}

void bpo_env_init(bpo_env* env, int env_id) {
    memset(env, 0, sizeof(*env));
    env->env_id = env_id;
    env->nblocks_max = k_nblocks[env_id];
    env->dimo = k_dimo[env_id];
    env->dimg = k_dimg[env_id];
    env->num_objs = env->nblocks_max + 2; /* fetch_env.py:75 */
    /* tasks.py:18 obj_range=0.15, then the subclass overrides */
    env->obj_range = 0.15;
    env->max_obj_range = 0.15;
    switch (env_id) {
        case BPO_BLOCKS_TOUCH: /* fetch_env.py:346-348 */
            env->max_obj_range = env->obj_range; env->obj_range_step = 0; env->has_curriculum_step = 1; break;
        case BPO_BLOCKS_TOUCH_CURRICULUM: /* :342-345 */
        case BPO_BLOCKS_TOUCH_VARIATION:  /* :561-563 */
            env->obj_range = 0.08; env->obj_range_step = 0.025; env->max_obj_range = 0.2;
            env->has_curriculum_step = 1; break;
        case BPO_BLOCKS_TOUCH_CHOOSE: /* :413-415 */
            env->wrong_obj_range = 0; env->max_obj_range = 0.2; env->has_curriculum_step = 0; break;
        case BPO_BLOCKS_TOUCH_CHOOSE_CURRICULUM: /* :407-412 */
            env->obj_range = 0.08; env->obj_range_step = 0.025; env->wrong_obj_range = 0.2;
            env->wrong_obj_range_step = 0.02; env->max_obj_range = 0.3; env->has_curriculum_step = 1; break;
        default: break;
    }
    sample_colors(env_id, env->colors);
    for (int k = 0; k < BPO_MAX_DIMG; ++k) env->ag[k] = -1; /* fetch_env.py:78 */
    bpo_sim_init(&env->sim, env_id);
    sample_goal(env); /* robot_env.py:37 */
}

void bpo_env_seed(bpo_env* env, uint64_t seed) { /* robot_env.py:53-55 */
    env->seed = seed;
    env->episode = 0;
    env->draws[0] = env->draws[1] = 0;
}

/* _get_obs fetch_env.py:187-228; Variation :567-621 */
void bpo_env_get_obs(const bpo_env* env, float* obs, float* ag, float* g) {
    const bpo_sim* s = &env->sim;
    int var = env->env_id == BPO_BLOCKS_TOUCH_VARIATION;
    int num_blocks = env->num_objs - 2;
    float gvp[3] = {s->gv[0] * DT, s->gv[1] * DT, s->gv[2] * DT}; /* grip_velp * dt, :191 */
    int o = 0;
    if (var) obs[o++] = (float)num_blocks; /* :580 */
    obs[o++] = s->g[0]; obs[o++] = s->g[1]; obs[o++] = s->g[2];
    obs[o++] = s->q[0]; obs[o++] = s->q[1];             /* robot_qpos[-2:], :196 */
    obs[o++] = gvp[0]; obs[o++] = gvp[1]; obs[o++] = gvp[2];
    obs[o++] = s->qv[0] * DT; obs[o++] = s->qv[1] * DT; /* :197 */
    for (int i = 0; i < num_blocks; ++i) {
        const bpo_block* b = &s->blk[i];
        obs[o++] = b->pos[0]; obs[o++] = b->pos[1]; obs[o++] = b->pos[2];
        obs[o++] = b->pos[0] - s->g[0]; obs[o++] = b->pos[1] - s->g[1]; obs[o++] = b->pos[2] - s->g[2];
        obs[o++] = -0.0f; obs[o++] = 0.0f; obs[o++] = bpo_atan2(b->s, b->c); /* mat2euler of a pure yaw: roll = -arctan2(0, 1) = -0.0 */
        obs[o++] = b->vel[0] * DT - gvp[0]; obs[o++] = b->vel[1] * DT - gvp[1]; obs[o++] = b->vel[2] * DT - gvp[2];
        obs[o++] = 0.0f; obs[o++] = 0.0f; obs[o++] = b->w * DT;
        if (var) { /* one_hot_color, :599 */
            int c = env->colors[i + 2];
            for (int k = 0; k < BPO_NUM_COLORS; ++k) obs[o++] = (k == c) ? 1.0f : 0.0f;
        }
    }
    while (o < env->dimo) obs[o++] = 0.0f; /* padding, :606-607 */
    for (int k = 0; k < env->dimg; ++k) {
        ag[k] = (float)env->ag[k];
        g[k] = (float)env->goal[k];
    }
}

/* _reset_sim fetch_env.py:247-255 (Variation :646-679) + RobotEnv.reset robot_env.py:71-82 */
void bpo_env_reset(bpo_env* env, float* obs, float* ag, float* g) {
    env->draws[0] = env->draws[1] = 0;
    if (env->env_id == BPO_BLOCKS_TOUCH_VARIATION) {
        int num_grey = rs_randint(env, 0, 3);              /* :647 */
        env->num_objs = 4 + num_grey;                      /* :649 */
        bpo_sim_init(&env->sim, env->env_id);              /* :670 set_state(initial_states[num_grey]) */
        env->sim.nblocks = 2 + num_grey;
        for (int k = 0; k < BPO_MAX_DIMG; ++k) env->ag[k] = -1; /* :663 */
    } else {
        bpo_sim_init(&env->sim, env->env_id);              /* :248 */
    }
    sample_colors(env->env_id, env->colors);               /* :250 */
    randomize_objects(env, 0);                             /* :252 */
    env->has_succeeded = 0;                                /* :254 */
    sample_goal(env);                                      /* robot_env.py:80 */
    env->t = 0;                                            /* TimeLimit.reset [upstream] */
    env->episode += 1;
    if (obs) bpo_env_get_obs(env, obs, ag, g);             /* robot_env.py:81 */
}

/* set_test fetch_env.py:365-368, 443-446 (re-spawn at max range, stale obs: appendix A4);
 * Variation :641-644 only returns obs; GripperTouch/ToppleTower raise NotImplementedError (:100-101) */
int bpo_env_set_test(bpo_env* env, float* obs, float* ag, float* g) {
    if (env->env_id == BPO_GRIPPER_TOUCH || env->env_id == BPO_TOPPLE_TOWER) return -1;
    if (env->env_id == BPO_BLOCKS_TOUCH_VARIATION) {
        if (obs) bpo_env_get_obs(env, obs, ag, g);
        return 0;
    }
    /* the obs returned is computed from the site positions cached by the last
     * forward()/step(), i.e. BEFORE the new qpos is visible */
    if (obs) bpo_env_get_obs(env, obs, ag, g);
    env->episode -= 1; /* draws continue inside the current episode's counter space */
    randomize_objects(env, 1);
    env->episode += 1;
    sample_goal(env);
    if (g) for (int k = 0; k < env->dimg; ++k) g[k] = (float)env->goal[k];
    return 0;
}

/* increase_difficulty fetch_env.py:351-358, 419-432, 623-630 */
int bpo_env_increase_difficulty(bpo_env* env) {
    switch (env->env_id) {
        case BPO_BLOCKS_TOUCH:
        case BPO_BLOCKS_TOUCH_CURRICULUM:
        case BPO_BLOCKS_TOUCH_VARIATION:
            env->obj_range += env->obj_range_step;
            if (env->obj_range > env->max_obj_range) { env->obj_range = env->max_obj_range; return 1; }
            env->difficulty += 1;
            return 0;
        case BPO_BLOCKS_TOUCH_CHOOSE:
        case BPO_BLOCKS_TOUCH_CHOOSE_CURRICULUM:
            if (!env->has_curriculum_step) return -2; /* AttributeError: no obj_range_step (:413-415,420) */
            env->obj_range += env->obj_range_step;
            env->wrong_obj_range -= env->wrong_obj_range_step;
            if (env->obj_range > env->max_obj_range) {
                env->obj_range = env->max_obj_range;
                if (env->wrong_obj_range < 0) { env->wrong_obj_range = 0; return 1; }
            } else {
                if (env->wrong_obj_range < 0) env->wrong_obj_range = 0;
            }
            env->difficulty += 1;
            return 0;
        default: return -1; /* NotImplementedError, fetch_env.py:93-94 */
    }
}

int bpo_env_get_difficulty(const bpo_env* env) { return env->difficulty; } /* fetch_env.py:96-97 */
double bpo_env_get_obj_range(const bpo_env* env) { return env->obj_range; }
/* test hook (mirrors bp_set_ranges): direct write of the curriculum knobs */
int bpo_env_set_challenge(bpo_env* env, int challenge) { /* fetch_env.py:403,416 */
    if (env->env_id != BPO_BLOCKS_TOUCH_CHOOSE && env->env_id != BPO_BLOCKS_TOUCH_CHOOSE_CURRICULUM) return -1;
    env->challenge = challenge != 0;
    return 0;
}
void bpo_env_set_ranges(bpo_env* env, double obj_range, double wrong_obj_range) { env->obj_range = obj_range; env->wrong_obj_range = wrong_obj_range; }

/* _step_callback fetch_env.py:148-167 */
static void step_callback(bpo_env* env) {
    int N = env->num_objs;
    int stride = env->env_id == BPO_BLOCKS_TOUCH_VARIATION ? BPO_MAX_OBJS : N;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j)
            if (env->ag[i * stride + j] == 1) env->ag[i * stride + j] = 0; /* :154-157 */
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j)
            if (env->sim.contacts & (1u << bpo_pair_index(i, j))) {       /* :159-167 */
                env->ag[i * stride + j] = 1;
                env->ag[j * stride + i] = 1;
            }
}

/* RobotEnv.step robot_env.py:57-69 under gym TimeLimit (max_episode_steps=50) */
int bpo_env_step(bpo_env* env, const float action[4], float* obs, float* ag, float* g,
                 float* reward, int* is_success) {
    float a[4];
    for (int k = 0; k < 4; ++k) {
        float x = action[k];
        if (!(x == x)) { x = 0.0f; env->invalid_actions += 1; } /* NaN: flagged, not raised */
        a[k] = x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x); /* np.clip, :58 */
    }
    bpo_sim_set_action(&env->sim, a);  /* :59 */
    bpo_sim_step(&env->sim);           /* :60 */
    step_callback(env);                /* :61 */
    if (obs) bpo_env_get_obs(env, obs, ag, g); /* :62 */
    float r = env_reward(env);
    if (r == 0.0f) env->has_succeeded = 1; /* _is_success latch, fetch_env.py:275-281 */
    if (is_success) *is_success = env->has_succeeded;
    if (reward) *reward = r;           /* :68 */
    env->t += 1;
    return env->t >= BPO_MAX_EPISODE_STEPS;
}

void bpo_env_random_action(const bpo_env* env, float a[4]) {
    uint32_t w[4];
    bpo_philox4x32((uint32_t)env->t, env->episode - 1u, 2u, 0u, (uint32_t)env->seed,
                   (uint32_t)(env->seed >> 32), w);
    for (int k = 0; k < 4; ++k) a[k] = 2.0f * bpo_u01(w[k]) - 1.0f;
}

void bpo_env_get_state(const bpo_env* env, bpo_env_state* out) {
    memset(out, 0, sizeof(*out));
    const bpo_sim* s = &env->sim;
    memcpy(out->grip_pos, s->g, 12);
    memcpy(out->grip_vel, s->gv, 12);
    memcpy(out->finger_q, s->q, 8);
    memcpy(out->finger_qv, s->qv, 8);
    for (int i = 0; i < s->nblocks; ++i) { /* cubes that are not in the scene stay zero */
        memcpy(out->blk_pos[i], s->blk[i].pos, 12);
        out->blk_cs[i][0] = s->blk[i].c; out->blk_cs[i][1] = s->blk[i].s;
        memcpy(out->blk_vel[i], s->blk[i].vel, 12);
        out->blk_w[i] = s->blk[i].w;
    }
    memcpy(out->ag, env->ag, BPO_MAX_DIMG);
    out->num_objs = env->num_objs;
    out->has_succeeded = env->has_succeeded;
    out->t = env->t;
    out->episode = env->episode;
    out->draws[0] = env->draws[0]; out->draws[1] = env->draws[1];
}

void bpo_env_set_state(bpo_env* env, const bpo_env_state* in) {
    bpo_sim* s = &env->sim;
    memcpy(s->g, in->grip_pos, 12);
    memcpy(s->gv, in->grip_vel, 12);
    memcpy(s->q, in->finger_q, 8);
    memcpy(s->qv, in->finger_qv, 8);
    for (int i = 0; i < BPO_MAX_BLOCKS; ++i) {
        memcpy(s->blk[i].pos, in->blk_pos[i], 12);
        s->blk[i].c = in->blk_cs[i][0]; s->blk[i].s = in->blk_cs[i][1];
        memcpy(s->blk[i].vel, in->blk_vel[i], 12);
        s->blk[i].w = in->blk_w[i];
    }
    memcpy(env->ag, in->ag, BPO_MAX_DIMG);
    env->num_objs = in->num_objs;
    s->nblocks = in->num_objs - 2;
    env->has_succeeded = in->has_succeeded;
    env->t = in->t;
    env->episode = in->episode;
    env->draws[0] = in->draws[0]; env->draws[1] = in->draws[1];
}

/* ======================================================================
 * HER relabel: baselines.her.her._sample_her_transitions [upstream, recalled];
 * call sites config.py:107-123, ddpg.py:106,214-215.
 * ====================================================================== */
void bpo_her_relabel(const float* ep_ag, const float* ep_g, int32_t B, int32_t T, int32_t dimg,
                     int64_t n, float future_p, uint64_t seed, int64_t index_offset,
                     int32_t* ep_idx, int32_t* t_idx, int32_t* fut_t, float* ag2_out,
                     float* g_out, float* r_out) {
    for (int64_t i = 0; i < n; ++i) {
        uint64_t gi = (uint64_t)(i + index_offset);
        uint32_t w[4];
        bpo_philox4x32((uint32_t)gi, (uint32_t)(gi >> 32), 3u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
        int32_t e = (int32_t)(((uint64_t)w[0] * (uint64_t)B) >> 32); /* episode_idxs = randint(0, B) */
        int32_t t = (int32_t)(((uint64_t)w[1] * (uint64_t)T) >> 32); /* t_samples = randint(T) */
        int her = bpo_u01(w[2]) < future_p;                         /* uniform(size) < future_p */
        int32_t off = (int32_t)(bpo_u01(w[3]) * (float)(T - t));    /* (uniform * (T - t)).astype(int) */
        int32_t ft = t + 1 + off;
        const float* ag2 = ep_ag + ((int64_t)e * (T + 1) + (t + 1)) * dimg; /* ag_2 = ag[:, 1:] */
        const float* gsrc = her ? ep_ag + ((int64_t)e * (T + 1) + ft) * dimg
                                : ep_g + ((int64_t)e * T + t) * dimg;
        for (int k = 0; k < dimg; ++k) {
            g_out[i * dimg + k] = gsrc[k];
            if (ag2_out) ag2_out[i * dimg + k] = ag2[k];
        }
        bpo_compute_reward(ag2, gsrc, 1, dimg, &r_out[i]);
        if (ep_idx) ep_idx[i] = e;
        if (t_idx) t_idx[i] = t;
        if (fut_t) fut_t[i] = her ? ft : -1;
    }
}

/* ======================================================================
 * vectorised helpers
 * ====================================================================== */
void bpo_vec_init(bpo_env* envs, int64_t n, int env_id, uint64_t seed, uint64_t env_index_offset) {
    for (int64_t i = 0; i < n; ++i) {
        bpo_env_init(&envs[i], env_id);
        /* rollout.py:206-210: env idx is seeded with seed + 1000*idx */
        bpo_env_seed(&envs[i], seed + 1000ull * (env_index_offset + (uint64_t)i));
    }
}

void bpo_vec_reset(bpo_env* envs, int64_t n, float* obs, float* ag, float* g) {
    for (int64_t i = 0; i < n; ++i) {
        int dimo = envs[i].dimo, dimg = envs[i].dimg;
        bpo_env_reset(&envs[i], obs ? obs + i * dimo : 0, ag ? ag + i * dimg : 0, g ? g + i * dimg : 0);
    }
}

void bpo_vec_step(bpo_env* envs, int64_t n, const float* actions, int auto_reset, float* obs,
                  float* ag, float* reward, float* success, float* reset_obs, float* reset_ag,
                  int32_t* stats) {
    float gbuf[BPO_MAX_DIMG];
    for (int64_t i = 0; i < n; ++i) {
        bpo_env* e = &envs[i];
        int dimo = e->dimo, dimg = e->dimg;
        float r;
        int succ;
        uint32_t inv0 = e->invalid_actions;
        int done = bpo_env_step(e, actions + i * 4, obs + i * dimo, ag + i * dimg, gbuf, &r, &succ);
        reward[i] = r;
        success[i] = (float)succ;
        if (stats) {
            stats[2] += 1;
            stats[3] += (int32_t)(e->invalid_actions - inv0);
            if (done) { stats[0] += 1; stats[1] += succ; }
        }
        if (done && auto_reset) {
            float tmp_o[BPO_MAX_DIMO], tmp_a[BPO_MAX_DIMG];
            bpo_env_reset(e, reset_obs ? reset_obs + i * dimo : tmp_o, reset_ag ? reset_ag + i * dimg : tmp_a, gbuf);
        }
    }
}

/* CPU throughput loop: every env runs `steps` steps with its own Philox actions,
 * auto-resetting at T = 50.  stats: [episodes, successes, steps].  Single-threaded:
 * callers parallelise by slicing the env array over host threads (ctypes drops the GIL). */
void bpo_vec_step_random(bpo_env* envs, int64_t n, int steps, int nthreads, int64_t* stats) {
    int64_t ep = 0, su = 0, st = 0;
    for (int64_t i = 0; i < n; ++i) {
        bpo_env* e = &envs[i];
        float obs[BPO_MAX_DIMO], ag[BPO_MAX_DIMG], g[BPO_MAX_DIMG], a[4], r;
        int succ;
        if (e->episode == 0) bpo_env_reset(e, obs, ag, g);
        for (int s = 0; s < steps; ++s) {
            bpo_env_random_action(e, a);
            int done = bpo_env_step(e, a, obs, ag, g, &r, &succ);
            st += 1;
            if (done) { ep += 1; su += succ; bpo_env_reset(e, obs, ag, g); }
        }
    }
    (void)nthreads;
    if (stats) { stats[0] += ep; stats[1] += su; stats[2] += st; }
}
