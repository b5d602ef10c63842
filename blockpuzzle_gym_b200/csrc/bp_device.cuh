// bp_device.cuh -- device-side implementation of the gym_blocks env hot path for sm_100a.
//
// One thread advances one env: the gripper in registers, the cubes in a private shared-memory
// column; the kernels in bp_kernels.cu decide how envs are tiled over warps.
// Arithmetic follows the BlockPhys v2 specification of DESIGN.md section 3 exactly (fp32,
// every operation individually rounded: compile with -fmad=false), so integer
// state is bit-exact and float state bit-identical against the CPU oracle.
//
// Reference functions implemented (paths relative to /root/reference/gym_blocks):
//   envs/robot_env.py:57-82      step / reset
//   envs/fetch_env.py:135-143    compute_reward
//   envs/fetch_env.py:148-167    _step_callback (touch matrix)
//   envs/fetch_env.py:170-185    _set_action
//   envs/fetch_env.py:187-228    _get_obs (:567-621 Variation)
//   envs/fetch_env.py:247-281    _reset_sim, _sample_goal, _is_success
//   envs/fetch_env.py:328-336, 370-399, 448-517, 646-764, 777-787  spawn samplers
#pragma once
#include <cstdint>

#include "bp_tables.cuh"

namespace bp {

// ---------------------------------------------------------------- constants
constexpr int kNSub = 20;            // tasks.py:16 n_substeps
constexpr float kH = 0.002f;         // 2blocks.xml:4
constexpr float kInvH = 500.0f;
constexpr float kDt = 0.04f;         // fetch_env.py:190
constexpr float kGH = 0.01962f;
constexpr float kFr = 0.01962f;
constexpr float kFrW = 0.9f;
constexpr float kKW = 2500.0f;
constexpr float kBW = 100.0f;
constexpr float kKF = 288.46155f;
constexpr float kBF = 9.615385f;
constexpr float kQMax = 0.05f;
constexpr float kCtrlMax = 0.2f;
constexpr float kTblX = 1.3f, kTblY = 0.75f, kTblHX = 0.25f, kTblHY = 0.35f;
constexpr float kHB = 0.025f, kTwoHB = 0.05f;
constexpr float kZRest = 0.485f, kZFloor = 0.025f;
constexpr float kFX = 0.0135f, kFY = 0.007f, kFZ = 0.0385f, kFY0 = 0.0079f, kFZOff = 0.02f;
constexpr float kGZMin = 0.4785f;
constexpr float kMargin = 0.001f;
constexpr float kDepen = 0.0005f;
constexpr float kVMax = 5.0f, kWMax = 60.0f;
constexpr float kIInv = 2400.0f;
constexpr float kPosScale = 0.05f;   // fetch_env.py:175
constexpr float kWsXLo = 1.0f, kWsXHi = 1.6f, kWsYLo = 0.35f, kWsYHi = 1.15f, kWsZHi = 0.9f;
constexpr float kGrip0X = 1.3419f, kGrip0Y = 0.7491f, kGrip0Z = 0.5347f;
// spawn geometry, fetch_env.py:19-32: python floats (binary64), written as the reference's own expressions so that
// they fold to the same doubles (kTableH = 0.32499999999999996, kMinBlockDist = 0.07500000000000001)
constexpr double kBlockSize = 0.05;
constexpr double kMinBlockDist = 1.5 * kBlockSize;
constexpr double kTableX = 1.05 + 0.25, kTableY = 0.40 + 0.35;
constexpr double kTableW = 0.25 - kBlockSize / 2, kTableH = 0.35 - kBlockSize / 2;
constexpr int kMaxSpawnAttempts = 10000;
constexpr int kT = 50;               // __init__.py:10
constexpr int kMaxObjs = 6;

__host__ __device__ constexpr int pair_index(int o1, int o2) {  // o1 < o2
    return o1 * (2 * kMaxObjs - o1 - 1) / 2 + (o2 - o1 - 1);
}
__host__ __device__ constexpr uint32_t pair_bit(int o1, int o2) { return 1u << pair_index(o1, o2); }

// ---------------------------------------------------------------- per-env-id configuration
// SURVEY.md section 8 table; colours from fetch_env.py:323-326,360-363,434-441,632-639,772-775
// reduce to the goal masks P (pairs that must touch) and M (pairs that must never have touched).
template <int ID> struct Cfg;
template <> struct Cfg<0> { static constexpr int NB = 1, DIMO = 25, DIMG = 9;  static constexpr bool BG = false, VAR = false; static constexpr uint32_t P = pair_bit(0, 2), M = 0; };
template <> struct Cfg<1> { static constexpr int NB = 2, DIMO = 40, DIMG = 16; static constexpr bool BG = false, VAR = false; static constexpr uint32_t P = pair_bit(2, 3), M = 0; };
template <> struct Cfg<2> { static constexpr int NB = 4, DIMO = 70, DIMG = 36; static constexpr bool BG = false, VAR = false; static constexpr uint32_t P = pair_bit(1, 5), M = pair_bit(0, 5); };
template <> struct Cfg<3> { static constexpr int NB = 2, DIMO = 40, DIMG = 16; static constexpr bool BG = true,  VAR = false; static constexpr uint32_t P = pair_bit(2, 3), M = 0; };
template <> struct Cfg<4> { static constexpr int NB = 3, DIMO = 55, DIMG = 25; static constexpr bool BG = true,  VAR = false; static constexpr uint32_t P = pair_bit(2, 3), M = 0; };
template <> struct Cfg<5> { static constexpr int NB = 3, DIMO = 55, DIMG = 25; static constexpr bool BG = true,  VAR = false; static constexpr uint32_t P = pair_bit(2, 3), M = 0; };
template <> struct Cfg<6> { static constexpr int NB = 4, DIMO = 87, DIMG = 36; static constexpr bool BG = true,  VAR = true;  static constexpr uint32_t P = pair_bit(2, 3), M = 0; };

// curriculum knobs handed to the kernels (host keeps the python doubles, fetch_env.py:340-348,404-415,561-563)
struct Ranges {
    double obj_range, max_obj_range, wrong_obj_range;
    int challenge;   // BlocksTouchChooseEnv(challenge=True), fetch_env.py:403,416,452-463 (bp_set_option "challenge")
};

// ---------------------------------------------------------------- Philox + elementary functions
struct U4 { uint32_t x, y, z, w; };

// out of line (like bp_log / bp_sincos2pi below): the spawn samplers call these from many sites, and the step
// kernel's speed depends on how much code its resident warps push through the instruction cache
static __device__ __noinline__ U4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }
__device__ __forceinline__ float u01_open(uint32_t x) { return ((float)(x >> 9) + 0.5f) * 1.1920928955078125e-07f; }

static __device__ __noinline__ float bp_log(float x) {
    uint32_t ix = __float_as_uint(x);
    ix += 0x3f800000u - 0x3f3504f3u;
    int e = (int)(ix >> 23) - 127;
    ix = (ix & 0x007fffffu) + 0x3f3504f3u;
    float f = __uint_as_float(ix) - 1.0f;
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

static __device__ __noinline__ void bp_sincos2pi(float u, float& sn, float& cs) {
    float t = u * 4.0f;
    int k = (int)(t + 0.5f);
    float f = t - (float)k;
    float x = f * 1.57079637f;
    float x2 = x * x;
    float sp = x + x * x2 * (-0.16666667f + x2 * (0.0083333338f + x2 * (-0.00019841270f)));
    float cp = 1.0f + x2 * (-0.5f + x2 * (0.041666668f + x2 * (-0.0013888889f + x2 * 2.4801588e-05f)));
    switch (k & 3) {
        case 0: sn = sp; cs = cp; break;
        case 1: sn = cp; cs = -sp; break;
        case 2: sn = -sp; cs = -cp; break;
        default: sn = -cp; cs = sp; break;
    }
}

__device__ __forceinline__ float bp_atan2(float s, float c) {
    float as = fabsf(s), ac = fabsf(c);
    float mx = as > ac ? as : ac;
    float mn = as > ac ? ac : as;
    if (mx == 0.0f) return 0.0f;
    // 0 / mx is +0 exactly; dividing mx / mx instead keeps the zero numerator (a never-turned cube: s = 0) away from the
    // IEEE division's slow path, which the hardware check sends every zero operand through (same result bits)
    float a = (mn == 0.0f ? mx : mn) / mx;
    a = mn == 0.0f ? 0.0f : a;
    float off = 0.0f;
    if (a > 0.41421357f) {
        a = (a - 1.0f) / (a + 1.0f);
        off = 0.78539819f;
    }
    float a2 = a * a;
    float p = 0.076923080f;
    p = -0.090909094f + a2 * p;
    p = 0.11111111f + a2 * p;
    p = -0.14285715f + a2 * p;
    p = 0.2f + a2 * p;
    p = -0.33333334f + a2 * p;
    float r = off + (a + a * a2 * p);
    if (as > ac) r = 1.57079637f - r;
    if (c < 0.0f) r = 3.14159274f - r;
    if (s < 0.0f) r = -r;
    return r;
}

__device__ __forceinline__ void bp_normal2(uint32_t w0, uint32_t w1, float& z0, float& z1) {
    float u1 = u01_open(w0);
    float u2 = u01(w1);
    float r = sqrtf(-2.0f * bp_log(u1));
    float sn, cs;
    bp_sincos2pi(u2, sn, cs);
    z0 = r * cs;
    z1 = r * sn;
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
// clamp between two NONZERO bounds (velocity caps, the mocap reach box): min(max(x, lo), hi) is the same value as clampf
// for every non-NaN x (no signed-zero case), a NaN becomes lo on both sides of the parity check, and it is two FMNMX
// instead of four compare / select instructions
__device__ __forceinline__ float clampnz(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
// BlockPhys v1.1: the fused multiply-adds of the spec are written explicitly (the file is compiled
// with -fmad=false, so nothing else is ever contracted)
#define F(a, b, c) __fmaf_rn((a), (b), (c))

// ---------------------------------------------------------------- env state in registers
template <int NB>
struct Env {
    // dynamics slot (replaces MjSim)
    float g[3], gv[3];
    float q[2], qv[2];
    float px[NB], py[NB], pz[NB], c[NB], s[NB], vx[NB], vy[NB], vz[NB], w[NB];
    // env logic
    uint32_t touch_now, touch_ever;  // touch matrix as pair masks: 1 <-> now, 0 <-> ever & ~now, -1 <-> ~ever
    uint32_t contacts;               // contact pairs of the most recent substep
    int nb;                          // blocks present (num_objs - 2)
    int t;
    int succ;
    uint32_t episode, draws0, draws1;
    uint32_t key0, key1;             // Philox key = this env's seed
    uint32_t priv;                   // kernel-private: bit 31 = cubes on a fixed point, bits 0-14 = their contact pairs
};

template <int NB>
__device__ __forceinline__ void sim_init(Env<NB>& e, bool tower) {
    e.g[0] = kGrip0X; e.g[1] = kGrip0Y; e.g[2] = kGrip0Z;
    e.gv[0] = e.gv[1] = e.gv[2] = 0.0f;
    e.q[0] = e.q[1] = 0.0f; e.qv[0] = e.qv[1] = 0.0f;
    float z = kZRest;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        e.px[i] = 0.0f; e.py[i] = 0.0f;
        e.pz[i] = (i < e.nb) ? (tower ? z : kZRest) : 0.0f;
        z = z + kTwoHB;
        e.c[i] = 1.0f; e.s[i] = 0.0f;
        e.vx[i] = e.vy[i] = e.vz[i] = 0.0f; e.w[i] = 0.0f;
    }
    e.contacts = 0;
}

// ---------------------------------------------------------------- BlockPhys v2: sim.step()
// The hot physics is written as compact loops over one env's cubes held in a PRIVATE shared-memory
// column (field f of cube b at p[(9*b+f)*STRIDE]; per-substep scratch behind it), the gripper in
// registers.  Loops are deliberately not unrolled: the whole step must stay inside the instruction
// cache (a fully unrolled version is >100 KB of SASS and stalls on instruction fetch).
struct Blk { float x, y, z, c, s, vx, vy, vz, w; };
struct Grip { float g[3], gv[3], q[2], qv[2]; };
struct GripSub { float gox, goy, qo[2], closed[2]; };
// BlockPhys v2: the env-step's propagator base.  The weld and the finger actuators are linear, so until a contact acts
// on a channel its state after n substeps is the n-th power of the substep map (bp_tables.cuh) applied to the state the
// env-step started from (d0 / e0 = position - target, v0 / w0 = velocity).  x and y are never acted on; z leaves the
// propagator when a finger lands on a cube (det bit 0), a finger when its closing is undone (bits 1 / 2) -- from then
// on that channel advances by the substep recurrence.
struct GripStep { float d0[3], v0[3], e0[2], w0[2]; uint32_t det; };

template <int NB, int STRIDE, int SSTRIDE = STRIDE>
struct Col {
    float* p;   // cube fields: field f of cube b at p[(9*b+f)*STRIDE]
    float* s;   // per-substep scratch: s[(4*b+f)*SSTRIDE] (default: right behind the fields)
    __device__ __forceinline__ Col(float* fields) : p(fields), s(fields + 9 * NB * STRIDE) {}
    __device__ __forceinline__ Col(float* fields, float* scratch) : p(fields), s(scratch) {}
    __device__ __forceinline__ Blk load(int b) const {
        const float* q = p + 9 * b * STRIDE;
        return Blk{q[0], q[STRIDE], q[2 * STRIDE], q[3 * STRIDE], q[4 * STRIDE], q[5 * STRIDE], q[6 * STRIDE], q[7 * STRIDE], q[8 * STRIDE]};
    }
    __device__ __forceinline__ void store(int b, const Blk& k) const {
        float* q = p + 9 * b * STRIDE;
        q[0] = k.x; q[STRIDE] = k.y; q[2 * STRIDE] = k.z; q[3 * STRIDE] = k.c; q[4 * STRIDE] = k.s;
        q[5 * STRIDE] = k.vx; q[6 * STRIDE] = k.vy; q[7 * STRIDE] = k.vz; q[8 * STRIDE] = k.w;
    }
    __device__ __forceinline__ void store_pose(int b, const Blk& k) const {
        float* q = p + 9 * b * STRIDE;
        q[0] = k.x; q[STRIDE] = k.y; q[2 * STRIDE] = k.z; q[3 * STRIDE] = k.c; q[4 * STRIDE] = k.s;
    }
    // scratch: start-of-substep position (0..2) and accumulated yaw change (3)
    __device__ __forceinline__ float& scr(int b, int f) const { return s[(4 * b + f) * SSTRIDE]; }
    static constexpr int kFields = 13 * NB;
    static constexpr int kScratch = 4 * NB;
};

__device__ __forceinline__ void rot_apply(float& c, float& s, float dth) {
    float c2 = F(-s, dth, c);
    float s2 = F(c, dth, s);
    float n2 = F(c2, c2, s2 * s2);
    float r = F(-0.5f, n2, 1.5f);
    c = c2 * r;
    s = s2 * r;
}

struct Rect { float x, y, c, s, hx, hy; };
struct Sat { float ov[4], proj[4], rtA[4], rtB[4]; };

// Separating-axis overlaps of two rectangles; axes 0 = A.u, 1 = A.v, 2 = B.u, 3 = B.v.  Evaluated
// axis by axis and abandoned (returns false) as soon as one overlap is <= -margin or `ovz` already
// is: the pair then cannot be in contact, exactly as min(ovz, ov[0..3]) > -margin would decide.
// AA: A is axis-aligned (c = 1, s = 0), which makes cr = B.c, sr = B.s, proj[0] = dx, proj[1] = dy exactly.
template <bool AA>
__device__ __forceinline__ bool sat_eval(const Rect& A, const Rect& B, Sat& o, const float ovz) {
    float cr = AA ? B.c : F(A.c, B.c, A.s * B.s);
    float sr = AA ? B.s : F(A.c, B.s, -(A.s * B.c));
    float C = fabsf(cr), S = fabsf(sr);
    float dx = B.x - A.x, dy = B.y - A.y;
    float RBu = F(B.hx, C, B.hy * S);
    float RBv = F(B.hx, S, B.hy * C);
    o.proj[0] = AA ? dx : F(dx, A.c, dy * A.s);
    o.proj[1] = AA ? dy : F(dy, A.c, -(dx * A.s));
    o.ov[0] = (A.hx + RBu) - fabsf(o.proj[0]);
    o.ov[1] = (A.hy + RBv) - fabsf(o.proj[1]);
    // one branch for the three cheap axes (z and A's two): a single reconvergence point per pair
    if (!(ovz > -kMargin) || !(o.ov[0] > -kMargin) || !(o.ov[1] > -kMargin)) return false;
    float RAu = F(A.hx, C, A.hy * S);
    float RAv = F(A.hx, S, A.hy * C);
    o.proj[2] = F(dx, B.c, dy * B.s);
    o.proj[3] = F(dy, B.c, -(dx * B.s));
    o.ov[2] = (RAu + B.hx) - fabsf(o.proj[2]);
    o.ov[3] = (RAv + B.hy) - fabsf(o.proj[3]);
    o.rtA[0] = A.hy; o.rtB[0] = RBv;
    o.rtA[1] = A.hx; o.rtB[1] = RBu;
    o.rtA[2] = RAv;  o.rtB[2] = B.hy;
    o.rtA[3] = RAu;  o.rtB[3] = B.hx;
    return true;
}

// argmin with first-wins ties; returns the value through `mn`
__device__ __forceinline__ int sat_argmin(const Sat& o, float& mn) {
    int k = 0;
    mn = o.ov[0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (o.ov[i] < mn) { mn = o.ov[i]; k = i; }
    return k;
}

__device__ __forceinline__ void sat_contact(const Rect& A, const Rect& B, const Sat& o, int k,
                                            float& nx, float& ny, float& rnA, float& rnB) {
    // select by k without dynamic register indexing
    float pk = k == 0 ? o.proj[0] : k == 1 ? o.proj[1] : k == 2 ? o.proj[2] : o.proj[3];
    float pt = k == 0 ? o.proj[1] : k == 1 ? o.proj[0] : k == 2 ? o.proj[3] : o.proj[2];
    float ra = k == 0 ? o.rtA[0] : k == 1 ? o.rtA[1] : k == 2 ? o.rtA[2] : o.rtA[3];
    float rb = k == 0 ? o.rtB[0] : k == 1 ? o.rtB[1] : k == 2 ? o.rtB[2] : o.rtB[3];
    float ex = k == 0 ? A.c : k == 1 ? -A.s : k == 2 ? B.c : -B.s;
    float ey = k == 0 ? A.s : k == 1 ? A.c : k == 2 ? B.s : B.c;
    float sg = pk >= 0.0f ? 1.0f : -1.0f;
    nx = sg * ex;
    ny = sg * ey;
    float lo = -ra;
    float lo2 = pt - rb;
    if (lo2 > lo) lo = lo2;
    float hi = ra;
    float hi2 = pt + rb;
    if (hi2 < hi) hi = hi2;
    float mid = 0.5f * (lo + hi);
    float chi = (k & 1) ? sg : -sg;
    rnA = chi * mid;
    rnB = chi * (mid - pt);
}

__device__ __forceinline__ bool over_table(float x, float y) {
    return fabsf(x - kTblX) <= kTblHX && fabsf(y - kTblY) <= kTblHY;
}

// finger f (0: +y, 1: -y) against cube b (index bi); ox/oy: the cube's start-of-substep position
// Returns false when the pair is not penetrating (nothing but `contacts` was touched), true when the response ran
// and the cube, the gripper height or the finger opening may have changed.
__device__ __forceinline__ bool collide_finger_block(Grip& e, GripSub& st, GripStep& sb, const int n, const int f, Blk& b, const float ox, const float oy,
                                                     float& dth_acc, bool& rotated, uint32_t& sup, uint32_t& contacts, const int bi) {
    const float sgn = f == 0 ? 1.0f : -1.0f;
    const float qf = f == 0 ? e.q[0] : e.q[1];
    Rect A{e.g[0], e.g[1] + sgn * (kFY0 + qf), 1.0f, 0.0f, kFX, kFY};
    float az = e.g[2] + kFZOff;
    Rect B{b.x, b.y, b.c, b.s, kHB, kHB};
    float dz = b.z - az;
    float ovz = (kFZ + kHB) - fabsf(dz);
    Sat o;
    if (!sat_eval<true>(A, B, o, ovz)) return false;
    float minxy;
    int k = sat_argmin(o, minxy);
    float minov = ovz < minxy ? ovz : minxy;
    if (!(minov > -kMargin)) return false;
    contacts |= pair_bit(0, 2) << bi;  // pair (0, bi+2): any "finger" geom -> object 0 (fetch_env.py:111-112)
    if (!(minov > 0.0f)) return false;
    if (ovz <= minxy) {
        if (dz >= 0.0f) {
            b.z = az + (kFZ + kHB);
            sup |= 1u << bi;
        } else {
            e.g[2] = (b.z + (kFZ + kHB)) - kFZOff;
            if (!(sb.det & 1u)) {   // z leaves the propagator with the velocity it has there at this substep
                e.gv[2] = F(kGC[n], sb.d0[2], kGD[n] * sb.v0[2]);
                sb.det |= 1u;
            }
            if (e.gv[2] < 0.0f) e.gv[2] = 0.0f;
        }
        return true;
    }
    float delta = minxy;
    float q_now = qf;
    if (k == 1) {
        bool inner = (f == 0) ? (o.proj[1] < 0.0f) : (o.proj[1] > 0.0f);
        if (inner) {
            const float closed = f == 0 ? st.closed[0] : st.closed[1];
            float yield = delta < closed ? delta : closed;
            float room = kQMax - qf;
            if (yield > room) yield = room;
            if (yield > 0.0f) {
                q_now = qf + yield;
                if (f == 0) { e.q[0] = q_now; e.qv[0] = 0.0f; st.closed[0] = closed - yield; }
                else { e.q[1] = q_now; e.qv[1] = 0.0f; st.closed[1] = closed - yield; }
                sb.det |= 2u << f;   // this finger advances by the recurrence for the rest of the env-step
                delta = delta - yield;
            }
            if (!(delta > 0.0f)) return true;
        }
    }
    float nx, ny, rnA, rnB;
    sat_contact(A, B, o, k, nx, ny, rnA, rnB);
    const float qo = f == 0 ? st.qo[0] : st.qo[1];
    float fdx = e.g[0] - st.gox;
    float fdy = (e.g[1] - st.goy) + sgn * (q_now - qo);
    float rel = F((b.x - ox) - fdx, nx, ((b.y - oy) - fdy) * ny);
    float cap = kDepen - rel;
    float lam = delta < cap ? delta : cap;
    if (!(lam > 0.0f)) return true;
    float D = F(kIInv, rnB * rnB, 1.0f);
    float l = lam / D;
    b.x = F(nx, l, b.x);
    b.y = F(ny, l, b.y);
    float dth = (kIInv * rnB) * l;
    if (dth != 0.0f) {
        rot_apply(b.c, b.s, dth);
        dth_acc = dth_acc + dth;
        rotated = true;
    }
    return true;
}

// cubes i < j; aox.. are their start-of-substep positions
__device__ __forceinline__ bool collide_block_block(Blk& a, const float aox, const float aoy, float& adth,
                                                    Blk& b, const float box, const float boy, float& bdth,
                                                    bool& rotated, uint32_t& sup, uint32_t& contacts, const int i, const int j) {
    Rect A{a.x, a.y, a.c, a.s, kHB, kHB};
    Rect B{b.x, b.y, b.c, b.s, kHB, kHB};
    float dz = b.z - a.z;
    float ovz = kTwoHB - fabsf(dz);
    Sat o;
    if (!sat_eval<false>(A, B, o, ovz)) return false;
    float minxy;
    int k = sat_argmin(o, minxy);
    float minov = ovz < minxy ? ovz : minxy;
    if (!(minov > -kMargin)) return false;
    contacts |= 1u << pair_index(i + 2, j + 2);  // "objectK" -> K + 2 (fetch_env.py:115-116)
    if (!(minov > 0.0f)) return false;
    int pin = 0;
    if (ovz <= minxy) {
        if (dz >= 0.0f) {
            if (fabsf(o.proj[0]) <= kHB && fabsf(o.proj[1]) <= kHB) {
                b.z = a.z + kTwoHB;
                sup |= 1u << j;
                return true;
            }
            pin = 1;
        } else {
            if (fabsf(o.proj[2]) <= kHB && fabsf(o.proj[3]) <= kHB) {
                a.z = b.z + kTwoHB;
                sup |= 1u << i;
                return true;
            }
            pin = 2;
        }
    }
    float nx, ny, rnA, rnB;
    sat_contact(A, B, o, k, nx, ny, rnA, rnB);
    float rel = F((b.x - box) - (a.x - aox), nx, ((b.y - boy) - (a.y - aoy)) * ny);
    float cap = kDepen - rel;
    float lam = minxy < cap ? minxy : cap;
    if (!(lam > 0.0f)) return true;
    float wA = pin == 1 ? 0.0f : 1.0f;
    float wB = pin == 2 ? 0.0f : 1.0f;
    float D = F(kIInv, F(wA, rnA * rnA, wB * (rnB * rnB)), wA + wB);
    float l = lam / D;
    float lA = wA * l, lB = wB * l;
    a.x = F(-nx, lA, a.x);
    a.y = F(-ny, lA, a.y);
    b.x = F(nx, lB, b.x);
    b.y = F(ny, lB, b.y);
    float dthA = -((kIInv * rnA) * lA);
    float dthB = (kIInv * rnB) * lB;
    if (dthA != 0.0f) { rot_apply(a.c, a.s, dthA); adth = adth + dthA; rotated = true; }
    if (dthB != 0.0f) { rot_apply(b.c, b.s, dthB); bdth = bdth + dthB; rotated = true; }
    return true;
}

// the gripper part of substep n (1..20) of an env-step (steps 1-2 of the spec, BlockPhys v2)
template <bool BG>
__device__ __forceinline__ void grip_step_begin(const Grip& e, const float m[3], const float ctrl[2], GripStep& sb) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { sb.d0[d] = e.g[d] - m[d]; sb.v0[d] = e.gv[d]; }
#pragma unroll
    for (int f = 0; f < 2; ++f) { sb.e0[f] = BG ? 0.0f : e.q[f] - ctrl[f]; sb.w0[f] = BG ? 0.0f : e.qv[f]; }
    sb.det = 0;
}

template <bool BG>
__device__ __forceinline__ void substep_gripper(Grip& e, GripSub& st, const float m[3], const float ctrl[2], const GripStep& sb, const int n) {
    st.closed[0] = st.closed[1] = 0.0f;
    st.gox = e.g[0]; st.goy = e.g[1];
    st.qo[0] = e.q[0]; st.qo[1] = e.q[1];
    const float ga = kGA[n], gb = kGB[n];
    e.g[0] = F(ga, sb.d0[0], F(gb, sb.v0[0], m[0]));
    e.g[1] = F(ga, sb.d0[1], F(gb, sb.v0[1], m[1]));
    if (!(sb.det & 1u)) {   // free: the propagator, projected onto z >= kGZMin (finger bottoms on the table)
        const float zv = F(ga, sb.d0[2], F(gb, sb.v0[2], m[2]));
        e.g[2] = zv < kGZMin ? kGZMin : zv;
    } else {
        const float az = F(kKW, m[2] - e.g[2], -(kBW * e.gv[2]));
        e.gv[2] = F(az, kH, e.gv[2]);
        e.g[2] = F(e.gv[2], kH, e.g[2]);
        const bool low = e.g[2] < kGZMin;
        e.gv[2] = (low && e.gv[2] < 0.0f) ? 0.0f : e.gv[2];
        e.g[2] = low ? kGZMin : e.g[2];
    }
    if (!BG) {
        const float fa = kFA[n], fb = kFB[n];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            const float q_old = e.q[f];
            float q;
            if (!(sb.det & (2u << f))) {   // free: the propagator, projected onto the joint range
                q = clampf(F(fa, sb.e0[f], F(fb, sb.w0[f], ctrl[f])), 0.0f, kQMax);
            } else {
                const float acc = F(kKF, ctrl[f] - e.q[f], -(kBF * e.qv[f]));
                float qv = F(acc, kH, e.qv[f]);
                q = F(qv, kH, e.q[f]);
                const bool low = q < 0.0f;
                qv = (low && qv < 0.0f) ? 0.0f : qv;
                q = low ? 0.0f : q;
                const bool high = q > kQMax;
                qv = (high && qv > 0.0f) ? 0.0f : qv;
                q = high ? kQMax : q;
                e.qv[f] = qv;
            }
            e.q[f] = q;
            st.closed[f] = fmaxf(q_old - q, 0.0f);
        }
    }
}

// end of the env-step: the velocities of the channels that stayed free, with the limit rules applied once
template <bool BG>
__device__ __forceinline__ void grip_step_end(Grip& e, const float m[3], const float ctrl[2], const GripStep& sb) {
    e.gv[0] = F(kGC20, sb.d0[0], kGD20 * sb.v0[0]);
    e.gv[1] = F(kGC20, sb.d0[1], kGD20 * sb.v0[1]);
    if (!(sb.det & 1u)) {
        const float zv = F(kGA20, sb.d0[2], F(kGB20, sb.v0[2], m[2]));
        const float v = F(kGC20, sb.d0[2], kGD20 * sb.v0[2]);
        e.gv[2] = (zv < kGZMin && v < 0.0f) ? 0.0f : v;
    }
    if (!BG) {
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            if (!(sb.det & (2u << f))) {
                const float qf = F(kFA20, sb.e0[f], F(kFB20, sb.w0[f], ctrl[f]));
                float v = F(kFC20, sb.e0[f], kFD20 * sb.w0[f]);
                v = (qf < 0.0f && v < 0.0f) ? 0.0f : v;
                v = (qf > kQMax && v > 0.0f) ? 0.0f : v;
                e.qv[f] = v;
            }
        }
    }
}

// The whole env-step of the gripper when nothing acts on it: the propagator's 20th power (bit for bit what
// substep_gripper x 20 + grip_step_end leave when no contact event occurs), plus a box [lo, hi] and a finger opening
// bound qmax that contain the gripper at EVERY substep n = 1..20: the state at substep n is GA[n] d0 + GB[n] v0 past
// the target with GA[n] in [kGAMin, kGAMax] and GB[n] in [kGBMin, kGBMax] (all positive), so the interval product
// bounds it; likewise the fingers.  (The bound is not part of the model: it only decides which path computes the step.)
template <bool BG>
__device__ __forceinline__ void quiet_gripper_step(Grip& e, const float m[3], const float ctrl[2], float lo[3], float hi[3], float& qmax) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float d0 = e.g[d] - m[d], v0 = e.gv[d];
        const float a1 = kGAMin * d0, a2 = kGAMax * d0, b1 = kGBMin * v0, b2 = kGBMax * v0;
        lo[d] = m[d] + (fminf(a1, a2) + fminf(b1, b2));
        hi[d] = m[d] + (fmaxf(a1, a2) + fmaxf(b1, b2));
        const float gn = F(kGA20, d0, F(kGB20, v0, m[d]));
        const float vn = F(kGC20, d0, kGD20 * v0);
        if (d < 2) {
            e.g[d] = gn; e.gv[d] = vn;
        } else {
            const bool low = gn < kGZMin;
            e.g[2] = low ? kGZMin : gn;
            e.gv[2] = (low && vn < 0.0f) ? 0.0f : vn;
            lo[2] = fmaxf(lo[2], kGZMin); hi[2] = fmaxf(hi[2], kGZMin);
        }
    }
    qmax = fmaxf(e.q[0], e.q[1]);
    if (!BG) {
        qmax = 0.0f;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            const float e0 = e.q[f] - ctrl[f], w0 = e.qv[f];
            const float top = ctrl[f] + (fmaxf(kFAMin * e0, kFAMax * e0) + fmaxf(kFBMin * w0, kFBMax * w0));
            qmax = fmaxf(qmax, clampf(top, 0.0f, kQMax));
            const float qf = F(kFA20, e0, F(kFB20, w0, ctrl[f]));
            float v = F(kFC20, e0, kFD20 * w0);
            v = (qf < 0.0f && v < 0.0f) ? 0.0f : v;
            v = (qf > kQMax && v > 0.0f) ? 0.0f : v;
            e.q[f] = clampf(qf, 0.0f, kQMax);
            e.qv[f] = v;
        }
    }
}

// the cube part of one substep (steps 3-5 of the spec).  Returns true when the substep left every
// cube bit-for-bit unchanged (all at rest before and after, nothing moved or rotated): a fixed point.
template <int NB, int STRIDE, int SS>
__device__ __forceinline__ bool substep_cubes(Grip& e, GripSub& st, GripStep& sb, const int n, const Col<NB, STRIDE, SS> col, const int nb, uint32_t& contacts) {
    uint32_t sup = 0;
    bool rest = true, rotated = false;
    contacts = 0;
    // 3. predict + 4a. table / floor support (both only depend on the cube itself)
#pragma unroll 1
    for (int i = 0; i < nb; ++i) {
        Blk b = col.load(i);
        rest = rest && b.vx == 0.0f && b.vy == 0.0f && b.vz == 0.0f && b.w == 0.0f;
        col.scr(i, 0) = b.x; col.scr(i, 1) = b.y; col.scr(i, 2) = b.z;
        float dth = 0.0f;
        b.vz = b.vz - kGH;
        b.x = F(b.vx, kH, b.x);
        b.y = F(b.vy, kH, b.y);
        b.z = F(b.vz, kH, b.z);
        if (b.w != 0.0f) {
            dth = b.w * kH;
            rot_apply(b.c, b.s, dth);
            rotated = true;
        }
        col.scr(i, 3) = dth;
        {   // written as selects: no divergent branches
            const bool ot = over_table(b.x, b.y);
            const bool touch_t = ot && (b.z - kZRest < kMargin);   // "table" -> 1 (fetch_env.py:113-114)
            const bool on_t = ot && b.z < kZRest;
            const bool on_f = !ot && b.z < kZFloor;                // floor0 maps to None: no touch entry
            contacts |= touch_t ? (pair_bit(1, 2) << i) : 0u;
            b.z = on_t ? kZRest : (on_f ? kZFloor : b.z);
            sup |= (on_t || on_f) ? (1u << i) : 0u;
        }
        col.store_pose(i, b);
    }
    // 4b. fingers vs cubes
#pragma unroll 1
    for (int i = 0; i < nb; ++i) {
        Blk b = col.load(i);
        const float ox = col.scr(i, 0), oy = col.scr(i, 1);
        float dth = col.scr(i, 3);
        bool moved = false;   // collide_* return false when they left the cubes untouched: nothing to write back then
#pragma unroll 1
        for (int f = 0; f < 2; ++f) moved = collide_finger_block(e, st, sb, n, f, b, ox, oy, dth, rotated, sup, contacts, i) || moved;
        if (moved) {
            col.scr(i, 3) = dth;
            col.store_pose(i, b);
        }
    }
    // 4c. cube pairs
#pragma unroll 1
    for (int i = 0; i + 1 < nb; ++i) {
#pragma unroll 1
        for (int j = i + 1; j < nb; ++j) {
            Blk a = col.load(i), b = col.load(j);
            float adth = col.scr(i, 3), bdth = col.scr(j, 3);
            if (collide_block_block(a, col.scr(i, 0), col.scr(i, 1), adth, b, col.scr(j, 0), col.scr(j, 1), bdth, rotated, sup, contacts, i, j)) {
                col.scr(i, 3) = adth; col.scr(j, 3) = bdth;
                col.store_pose(i, a); col.store_pose(j, b);
            }
        }
    }
    // 4d. fingers vs table
    if (over_table(e.g[0], e.g[1]) && e.g[2] - kGZMin < kMargin) contacts |= pair_bit(0, 1);
    // 5. velocities from the position change, then Coulomb friction on supported cubes
    bool same = rest && !rotated;
#pragma unroll 1
    for (int i = 0; i < nb; ++i) {
        Blk b = col.load(i);
        const float ox = col.scr(i, 0), oy = col.scr(i, 1), oz = col.scr(i, 2);
        same = same && __float_as_uint(b.x) == __float_as_uint(ox) && __float_as_uint(b.y) == __float_as_uint(oy) &&
               __float_as_uint(b.z) == __float_as_uint(oz);
        b.vx = clampnz((b.x - ox) * kInvH, -kVMax, kVMax);
        b.vy = clampnz((b.y - oy) * kInvH, -kVMax, kVMax);
        b.vz = clampnz((b.z - oz) * kInvH, -kVMax, kVMax);
        b.w = clampnz(col.scr(i, 3) * kInvH, -kWMax, kWMax);
        {
            const bool supd = (sup >> i & 1u) != 0u;
            const float sp2 = F(b.vx, b.vx, b.vy * b.vy);
            if (supd && sp2 > kFr * kFr) {      // sliding: the only branch (sqrt + divide)
                float sp = sqrtf(sp2);
                float kf = (sp - kFr) / sp;
                b.vx = b.vx * kf;
                b.vy = b.vy * kf;
            } else if (supd) {
                b.vx = 0.0f; b.vy = 0.0f;
            }
            const float wdec = b.w > 0.0f ? b.w - kFrW : b.w + kFrW;
            b.w = supd ? ((fabsf(b.w) <= kFrW) ? 0.0f : wdec) : b.w;
        }
        same = same && b.vx == 0.0f && b.vy == 0.0f && b.vz == 0.0f && b.w == 0.0f;
        col.store(i, b);
    }
    return same;
}

// mocap target and finger actuator targets from the clipped action (fetch_env.py:170-185)
template <bool BG>
__device__ __forceinline__ void action_targets(const Grip& e, const float a[4], float m[3], float ctrl[2]) {
    // v1.3: product and sum rounded separately -- the reference's float32 `pos_ctrl *= 0.05` (fetch_env.py:175)
    // followed by upstream's float64 mocap_pos + pos_delta narrows to exactly this
    m[0] = clampnz(e.g[0] + a[0] * kPosScale, kWsXLo, kWsXHi);
    m[1] = clampnz(e.g[1] + a[1] * kPosScale, kWsYLo, kWsYHi);
    m[2] = clampnz(e.g[2] + a[2] * kPosScale, kGZMin, kWsZHi);
    float ga = BG ? 0.0f : a[3];
    ctrl[0] = clampf(e.q[0] + ga, 0.0f, kCtrlMax);
    ctrl[1] = clampf(e.q[1] + ga, 0.0f, kCtrlMax);
}

// _set_action (fetch_env.py:170-185, after the clip of robot_env.py:58) + sim.step() (robot_env.py:60).
// Returns whether the cubes ended the step on an exact fixed point; `contacts` = pairs of the last substep.
template <int NB, int STRIDE, bool BG, int SS>
__device__ __forceinline__ bool sim_step_col(Grip& g, const float a[4], const Col<NB, STRIDE, SS> col, const int nb, uint32_t& contacts) {
    float m[3], ctrl[2];
    action_targets<BG>(g, a, m, ctrl);
    GripStep sb;
    grip_step_begin<BG>(g, m, ctrl, sb);
    bool still = false;
#pragma unroll 1
    for (int n = 1; n <= kNSub; ++n) {
        GripSub st;
        substep_gripper<BG>(g, st, m, ctrl, sb, n);
        still = substep_cubes<NB, STRIDE, SS>(g, st, sb, n, col, nb, contacts);
    }
    grip_step_end<BG>(g, m, ctrl, sb);
    return still;
}

// ---------------------------------------------------------------- the same step with the cubes in registers
// sim_step_col above walks the cubes through shared-memory loops and evaluates every (finger, cube) and (cube, cube)
// slot of every substep in turn: the contact-response tails then run once per slot with the one or two lanes of the
// warp that need them (ncu, round 1: 19 % of the kernel's issue slots at <= 4 active lanes).  Here the cubes of one
// env live in registers for the whole env-step, the per-cube stages are unrolled, and the contact stage is
//   (1) branch-free candidate masks: stage 1 of the separating-axis test (z and the first rectangle's two axes -- the
//       same expressions sat_eval evaluates first) for every slot at once;
//   (2) while (mask): pop the lowest slot, run the UNCHANGED collide_* function on it.
// A slot outside the mask would have returned from sat_eval's first test without touching anything, so skipping it
// is exact; slots are still visited in the oracle's Gauss-Seidel order (fingers cube-major, then pairs i-major), and
// whenever a response may have moved something the masks of the later slots are recomputed from the current state.
// Arithmetic per slot is identical: results stay bit-for-bit those of sim_step_col and of the oracle.
template <int NB>
struct CubeRegs {
    float x[NB], y[NB], z[NB], c[NB], s[NB], vx[NB], vy[NB], vz[NB], w[NB];   // state
    float ox[NB], oy[NB], oz[NB], dth[NB];                                    // start-of-substep position, yaw change
};

template <int NB>
__device__ __forceinline__ float pick(const float (&a)[NB], const int i) {
    float r = a[0];
#pragma unroll
    for (int k = 1; k < NB; ++k) r = (i == k) ? a[k] : r;
    return r;
}
template <int NB>
__device__ __forceinline__ void put(float (&a)[NB], const int i, const float v) {
#pragma unroll
    for (int k = 0; k < NB; ++k) a[k] = (i == k) ? v : a[k];
}

// bit 2i + f: finger f passes stage 1 of sat_eval<true> against cube i (see collide_finger_block)
template <int NB, bool VAR>
__device__ __forceinline__ uint32_t finger_candidates(const Grip& e, const CubeRegs<NB>& q, const int nb) {
    const float az = e.g[2] + kFZOff;
    const float ay0 = e.g[1] + (kFY0 + e.q[0]);
    const float ay1 = e.g[1] - (kFY0 + e.q[1]);
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const float C = fabsf(q.c[i]), S = fabsf(q.s[i]);
        const float RBu = F(kHB, C, kHB * S);
        const float RBv = F(kHB, S, kHB * C);
        const float ov0 = (kFX + RBu) - fabsf(q.x[i] - e.g[0]);
        const float ovz = (kFZ + kHB) - fabsf(q.z[i] - az);
        const bool common = (!VAR || i < nb) && ovz > -kMargin && ov0 > -kMargin;
        const float ov1a = (kFY + RBv) - fabsf(q.y[i] - ay0);
        const float ov1b = (kFY + RBv) - fabsf(q.y[i] - ay1);
        m |= (common && ov1a > -kMargin) ? (1u << (2 * i)) : 0u;
        m |= (common && ov1b > -kMargin) ? (2u << (2 * i)) : 0u;
    }
    return m;
}

// bit p: pair p = (i, j), i-major, passes stage 1 of sat_eval<false> (see collide_block_block)
template <int NB, bool VAR>
__device__ __forceinline__ uint32_t pair_candidates(const CubeRegs<NB>& q, const int nb) {
    uint32_t m = 0;
    int p = 0;
#pragma unroll
    for (int i = 0; i + 1 < NB; ++i) {
#pragma unroll
        for (int j = i + 1; j < NB; ++j) {
            const float cr = F(q.c[i], q.c[j], q.s[i] * q.s[j]);
            const float sr = F(q.c[i], q.s[j], -(q.s[i] * q.c[j]));
            const float C = fabsf(cr), S = fabsf(sr);
            const float dx = q.x[j] - q.x[i], dy = q.y[j] - q.y[i];
            const float RBu = F(kHB, C, kHB * S);
            const float RBv = F(kHB, S, kHB * C);
            const float p0 = F(dx, q.c[i], dy * q.s[i]);
            const float p1 = F(dy, q.c[i], -(dx * q.s[i]));
            const float ov0 = (kHB + RBu) - fabsf(p0);
            const float ov1 = (kHB + RBv) - fabsf(p1);
            const float ovz = kTwoHB - fabsf(q.z[j] - q.z[i]);
            const bool ok = (!VAR || j < nb) && ovz > -kMargin && ov0 > -kMargin && ov1 > -kMargin;
            m |= ok ? (1u << p) : 0u;
            ++p;
        }
    }
    return m;
}

template <int NB, bool VAR>
__device__ __forceinline__ bool substep_cubes_reg(Grip& e, GripSub& st, GripStep& sb, const int n, CubeRegs<NB>& q, const int nb, uint32_t& contacts) {
    uint32_t sup = 0;
    bool rest = true, rotated = false;
    contacts = 0;
    // 3. predict + 4a. table / floor support
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        if (!VAR || i < nb) {
            rest = rest && q.vx[i] == 0.0f && q.vy[i] == 0.0f && q.vz[i] == 0.0f && q.w[i] == 0.0f;
            q.ox[i] = q.x[i]; q.oy[i] = q.y[i]; q.oz[i] = q.z[i];
            float dth = 0.0f;
            q.vz[i] = q.vz[i] - kGH;
            q.x[i] = F(q.vx[i], kH, q.x[i]);
            q.y[i] = F(q.vy[i], kH, q.y[i]);
            q.z[i] = F(q.vz[i], kH, q.z[i]);
            if (q.w[i] != 0.0f) {
                dth = q.w[i] * kH;
                rot_apply(q.c[i], q.s[i], dth);
                rotated = true;
            }
            q.dth[i] = dth;
            const bool ot = over_table(q.x[i], q.y[i]);
            const bool touch_t = ot && (q.z[i] - kZRest < kMargin);
            const bool on_t = ot && q.z[i] < kZRest;
            const bool on_f = !ot && q.z[i] < kZFloor;
            contacts |= touch_t ? (pair_bit(1, 2) << i) : 0u;
            q.z[i] = on_t ? kZRest : (on_f ? kZFloor : q.z[i]);
            sup |= (on_t || on_f) ? (1u << i) : 0u;
        }
    }
    // 4b. fingers vs cubes
    uint32_t fm = finger_candidates<NB, VAR>(e, q, nb);
    while (fm) {
        const int slot = __ffs((int)fm) - 1;
        fm &= fm - 1u;
        const int i = slot >> 1, f = slot & 1;
        Blk b{pick<NB>(q.x, i), pick<NB>(q.y, i), pick<NB>(q.z, i), pick<NB>(q.c, i), pick<NB>(q.s, i), 0.0f, 0.0f, 0.0f, 0.0f};
        float dth = pick<NB>(q.dth, i);
        if (collide_finger_block(e, st, sb, n, f, b, pick<NB>(q.ox, i), pick<NB>(q.oy, i), dth, rotated, sup, contacts, i)) {
            put<NB>(q.x, i, b.x); put<NB>(q.y, i, b.y); put<NB>(q.z, i, b.z); put<NB>(q.c, i, b.c); put<NB>(q.s, i, b.s);
            put<NB>(q.dth, i, dth);
            fm = finger_candidates<NB, VAR>(e, q, nb) & ~((2u << slot) - 1u);
        }
    }
    // 4c. cube pairs
    if (NB > 1) {
        uint32_t pm = pair_candidates<NB, VAR>(q, nb);
        while (pm) {
            const int slot = __ffs((int)pm) - 1;
            pm &= pm - 1u;
            int i = 0, j = 1, p = 0;
#pragma unroll
            for (int a = 0; a + 1 < NB; ++a) {
#pragma unroll
                for (int b = a + 1; b < NB; ++b) {
                    if (slot == p) { i = a; j = b; }
                    ++p;
                }
            }
            Blk a{pick<NB>(q.x, i), pick<NB>(q.y, i), pick<NB>(q.z, i), pick<NB>(q.c, i), pick<NB>(q.s, i), 0.0f, 0.0f, 0.0f, 0.0f};
            Blk b{pick<NB>(q.x, j), pick<NB>(q.y, j), pick<NB>(q.z, j), pick<NB>(q.c, j), pick<NB>(q.s, j), 0.0f, 0.0f, 0.0f, 0.0f};
            float adth = pick<NB>(q.dth, i), bdth = pick<NB>(q.dth, j);
            if (collide_block_block(a, pick<NB>(q.ox, i), pick<NB>(q.oy, i), adth, b, pick<NB>(q.ox, j), pick<NB>(q.oy, j), bdth,
                                    rotated, sup, contacts, i, j)) {
                put<NB>(q.x, i, a.x); put<NB>(q.y, i, a.y); put<NB>(q.z, i, a.z); put<NB>(q.c, i, a.c); put<NB>(q.s, i, a.s);
                put<NB>(q.dth, i, adth);
                put<NB>(q.x, j, b.x); put<NB>(q.y, j, b.y); put<NB>(q.z, j, b.z); put<NB>(q.c, j, b.c); put<NB>(q.s, j, b.s);
                put<NB>(q.dth, j, bdth);
                pm = pair_candidates<NB, VAR>(q, nb) & ~((2u << slot) - 1u);
            }
        }
    }
    // 4d. fingers vs table
    if (over_table(e.g[0], e.g[1]) && e.g[2] - kGZMin < kMargin) contacts |= pair_bit(0, 1);
    // 5. velocities from the position change, then Coulomb friction on supported cubes
    bool same = rest && !rotated;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        if (!VAR || i < nb) {
            same = same && __float_as_uint(q.x[i]) == __float_as_uint(q.ox[i]) && __float_as_uint(q.y[i]) == __float_as_uint(q.oy[i]) &&
                   __float_as_uint(q.z[i]) == __float_as_uint(q.oz[i]);
            float vx = clampnz((q.x[i] - q.ox[i]) * kInvH, -kVMax, kVMax);
            float vy = clampnz((q.y[i] - q.oy[i]) * kInvH, -kVMax, kVMax);
            const float vz = clampnz((q.z[i] - q.oz[i]) * kInvH, -kVMax, kVMax);
            float w = clampnz(q.dth[i] * kInvH, -kWMax, kWMax);
            const bool supd = (sup >> i & 1u) != 0u;
            const float sp2 = F(vx, vx, vy * vy);
            if (supd && sp2 > kFr * kFr) {
                const float sp = sqrtf(sp2);
                const float kf = (sp - kFr) / sp;
                vx = vx * kf;
                vy = vy * kf;
            } else if (supd) {
                vx = 0.0f; vy = 0.0f;
            }
            const float wdec = w > 0.0f ? w - kFrW : w + kFrW;
            w = supd ? ((fabsf(w) <= kFrW) ? 0.0f : wdec) : w;
            same = same && vx == 0.0f && vy == 0.0f && vz == 0.0f && w == 0.0f;
            q.vx[i] = vx; q.vy[i] = vy; q.vz[i] = vz; q.w[i] = w;
        }
    }
    return same;
}

// _set_action + sim.step() on a register-resident env (full-physics pass of the async step kernel)
template <int NB, bool BG, bool VAR>
__device__ __forceinline__ bool sim_step_reg(Grip& g, const float a[4], CubeRegs<NB>& q, const int nb, uint32_t& contacts) {
    float m[3], ctrl[2];
    action_targets<BG>(g, a, m, ctrl);
    GripStep sb;
    grip_step_begin<BG>(g, m, ctrl, sb);
    bool still = false;
#pragma unroll 1
    for (int n = 1; n <= kNSub; ++n) {
        GripSub st;
        substep_gripper<BG>(g, st, m, ctrl, sb, n);
        still = substep_cubes_reg<NB, VAR>(g, st, sb, n, q, nb, contacts);
    }
    grip_step_end<BG>(g, m, ctrl, sb);
    return still;
}

// contact bits that involve the gripper (object 0)
__host__ __device__ constexpr uint32_t gripper_pair_mask() {
    return pair_bit(0, 1) | pair_bit(0, 2) | pair_bit(0, 3) | pair_bit(0, 4) | pair_bit(0, 5);
}

// Quiet-step test of the tiled kernel.  The gripper trajectory of this step (computed without
// cubes) stayed inside [lo, hi] per axis with finger opening <= qmax.  If every cube is separated
// from that swept finger volume by more than the contact margin along world x, y or z -- the same
// three axes (k = 0, 1 and z) collide_finger_block tests -- no substep can report a finger contact,
// so the cubes (already on a fixed point) are untouched by this step.  kQuietEps absorbs rounding.
constexpr float kQuietEps = 1e-5f;
__device__ __forceinline__ bool cube_out_of_reach(float bx, float by, float bz, float c, float s,
                                                  const float lo[3], const float hi[3], float qmax) {
    const float RB = kHB * fabsf(c) + kHB * fabsf(s);
    const float pad = kMargin + kQuietEps;
    const float sepx = fmaxf(bx - hi[0], lo[0] - bx);
    const float sepy = fmaxf(by - hi[1], lo[1] - by);
    const float sepz = fmaxf(bz - (hi[2] + kFZOff), (lo[2] + kFZOff) - bz);
    return sepx >= (kFX + RB) + pad || sepy >= ((kFY0 + qmax) + kFY + RB) + pad || sepz >= (kFZ + kHB) + pad;
}

// ---------------------------------------------------------------- RNG replay
template <int NB>
__device__ __forceinline__ U4 env_draw(Env<NB>& e, int stream, uint32_t ep) {
    uint32_t& d = stream == 0 ? e.draws0 : e.draws1;
    U4 r = philox4x32(d, ep, (uint32_t)stream, 0u, e.key0, e.key1);
    d += 1;
    return r;
}

// The spawn samplers follow the reference's arithmetic operation by operation in binary64 (BlockPhys v1.3): the
// env's numpy code computes object_xpos in float64 from initial_gripper_xpos (the float64 image of the sim's
// binary32 grip position), the RandomState draws and python-float ranges; only set_joint_qpos narrows the result
// into the sim's binary32 qpos.  Rejection tests see the same float64 values as the reference's, so the draw
// counters cannot drift.  (Resets are one in 50 env-steps and run a warp at a time: the FP64 cost is invisible.)
__device__ __forceinline__ bool out_of_table(double x, double y) {  // fetch_env.py:30-32
    return fabs(x - kTableX) > kTableW || fabs(y - kTableY) > kTableH;
}
// np.linalg.norm of a float64 pair = sqrt(x.dot(x)): numpy's dot accumulates the second product with an fma
__device__ __forceinline__ double norm2(double x, double y) { return sqrt(__fma_rn(y, y, x * x)); }
// RandomState.uniform(lo, hi) = lo + (hi - lo) * u on a 24-bit Philox fraction
__device__ __forceinline__ double uni(double lo, double hi, uint32_t w) { return lo + (hi - lo) * (double)u01(w); }

// direction = normal(2)/|.|, mag = uniform(lo, hi) from the global np.random (stream 1)
template <int NB>
__device__ __forceinline__ void sample_around(Env<NB>& e, uint32_t ep, double bx, double by, double lo, double hi, double& x, double& y) {
    U4 r = env_draw(e, 1, ep);
    float z0, z1;
    bp_normal2(r.x, r.y, z0, z1);
    double d0 = (double)z0, d1 = (double)z1;
    double n = norm2(d0, d1);
    d0 = d0 / n; d1 = d1 / n;
    U4 r2 = env_draw(e, 1, ep);
    double mag = uni(lo, hi, r2.x);
    x = bx + d0 * mag;
    y = by + d1 * mag;
}

template <int NB>
__device__ __forceinline__ void sample_blue(Env<NB>& e, uint32_t ep, double r, double& x, double& y) {  // fetch_env.py:475-480
    int it = 0;
    do {
        U4 w = env_draw(e, 0, ep);
        x = (double)kGrip0X + uni(-r / 2, r / 2, w.x);
        y = (double)kGrip0Y + uni(-r / 2, r / 2, w.y);
    } while (out_of_table(x, y) && ++it < kMaxSpawnAttempts);
}

// _randomize_objects for every env id; `ep` is the Philox episode counter to draw under
template <int ID>
__device__ __forceinline__ void randomize_objects(Env<Cfg<ID>::NB>& e, uint32_t ep, bool test, const Ranges& rg) {
    constexpr int NB = Cfg<ID>::NB;
    const double g0x = (double)kGrip0X, g0y = (double)kGrip0Y;   // initial_gripper_xpos[:2]
    if (ID == 0 || ID == 2) {  // GripperTouch fetch_env.py:328-336, ToppleTower :777-787
        double r = rg.obj_range;
        double x = g0x, y = g0y;
        int it = 0;
        while (norm2(x - g0x, y - g0y) < 0.1 && it++ < kMaxSpawnAttempts) {
            U4 w = env_draw(e, 0, ep);
            x = g0x + uni(-r, r, w.x);
            y = g0y + uni(-r, r, w.y);
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) { e.px[i] = (float)x; e.py[i] = (float)y; }
    } else if (ID == 1 || ID == 3) {  // BlocksTouch fetch_env.py:370-399
        double r = test ? rg.max_obj_range : rg.obj_range;
        U4 w = env_draw(e, 0, ep);
        double x0 = g0x + uni(-r / 2, r / 2, w.x);
        double y0 = g0y + uni(-r / 2, r / 2, w.y);
        e.px[0] = (float)x0; e.py[0] = (float)y0;
        double x, y;
        int it = 0;
        do {
            sample_around(e, ep, x0, y0, kMinBlockDist, r, x, y);
        } while (out_of_table(x, y) && ++it < kMaxSpawnAttempts);
        if (NB > 1) { e.px[NB > 1 ? 1 : 0] = (float)x; e.py[NB > 1 ? 1 : 0] = (float)y; }
    } else if (ID == 4 || ID == 5) {  // BlocksTouchChoose fetch_env.py:448-517
        double r, wrong_r;
        if (test || rg.challenge) { r = rg.max_obj_range; wrong_r = 0.0; }   // :452-454
        else { r = rg.obj_range; wrong_r = rg.wrong_obj_range; }
        const double min_r = rg.challenge ? 0.15 : kMinBlockDist;             // :458-463
        const double max_wrong_r = rg.challenge ? 0.04 : rg.max_obj_range;
        double bx, by, gx, gy, wx, wy;
        sample_blue(e, ep, r, bx, by);
        int it = 0;
        do {
            sample_around(e, ep, bx, by, min_r, r, gx, gy);
        } while (out_of_table(gx, gy) && ++it < kMaxSpawnAttempts);
        double cx = (bx + gx) / 2.0, cy = (by + gy) / 2.0;
        it = 0;
        bool again;
        do {
            sample_around(e, ep, cx, cy, wrong_r, max_wrong_r, wx, wy);
            again = out_of_table(wx, wy) || norm2(wx - bx, wy - by) < kMinBlockDist || norm2(wx - gx, wy - gy) < kMinBlockDist;
        } while (again && ++it < kMaxSpawnAttempts);
        // colours [GREEN, BLUE, GREY]: green = 0, blue = 1, wrong = 2 (fetch_env.py:465-473)
        constexpr int iG = 0, iB = NB > 1 ? 1 : 0, iW = NB > 2 ? 2 : 0;
        e.px[iB] = (float)bx; e.py[iB] = (float)by;
        e.px[iG] = (float)gx; e.py[iG] = (float)gy;
        e.px[iW] = (float)wx; e.py[iW] = (float)wy;
    } else {  // BlocksTouchVariation fetch_env.py:697-764
        double r = test ? rg.max_obj_range : rg.obj_range;
        double ppx[4], ppy[4];
        double bx, by, gx, gy;
        sample_blue(e, ep, r, bx, by);
        int it = 0;
        do {
            sample_around(e, ep, bx, by, kMinBlockDist, r, gx, gy);
        } while (out_of_table(gx, gy) && ++it < kMaxSpawnAttempts);
        constexpr int iG = 0, iB = NB > 1 ? 1 : 0;
        e.px[iB] = (float)bx; e.py[iB] = (float)by;
        e.px[iG] = (float)gx; e.py[iG] = (float)gy;
        ppx[0] = bx; ppy[0] = by; ppx[1] = gx; ppy[1] = gy;
#pragma unroll
        for (int i = 2; i < NB; ++i) {
            if (i < e.nb) {
                double x, y;
                bool again;
                it = 0;
                do {
                    U4 wa = env_draw(e, 0, ep);  // _sample_from_table fetch_env.py:88-90
                    U4 wb = env_draw(e, 0, ep);
                    x = kTableX + uni(-kTableW, kTableW, wa.x);
                    y = kTableY + uni(-kTableH, kTableH, wb.x);
                    bool hit = false;
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        if (p < i && !hit && norm2(x - ppx[p], y - ppy[p]) < kMinBlockDist) hit = true;
                    again = hit ? true : out_of_table(x, y);
                } while (again && ++it < kMaxSpawnAttempts);
                e.px[i] = (float)x; e.py[i] = (float)y;
                ppx[i] = x; ppy[i] = y;
            }
        }
    }
}

// Kernel-private flags right after a spawn.  Spawned cubes rest on the table top (or form the
// ToppleTower stack) with zero velocity and yaw: an exact fixed point of substep_cubes as long as no
// two cubes are within contact range, which is verified here (yaw = 0: the SAT reduces to |dx|, |dy|).
// Returns bit 31 (cubes static) | the cube-only contact pairs such a substep reports, or 0 when unsure.
template <int ID>
__device__ __forceinline__ uint32_t spawn_priv(const Env<Cfg<ID>::NB>& e) {
    constexpr int NB = Cfg<ID>::NB;
    if (ID == 2)  // tower: table-cube0 and the three stacked pairs (identical xy)
        return 0x80000000u | pair_bit(1, 2) | pair_bit(2, 3) | pair_bit(3, 4) | pair_bit(4, 5);
    bool ok = true;
    uint32_t contacts = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        if (i < e.nb) {
            ok = ok && over_table(e.px[i], e.py[i]);
            contacts |= pair_bit(1, 2) << i;
#pragma unroll
            for (int j = i + 1; j < NB; ++j)
                if (j < e.nb) ok = ok && (fabsf(e.px[j] - e.px[i]) > 0.0511f || fabsf(e.py[j] - e.py[i]) > 0.0511f);
        }
    }
    return ok ? (0x80000000u | contacts) : 0u;
}

// RobotEnv.reset (robot_env.py:71-82) + _reset_sim (fetch_env.py:247-255, Variation :646-679) + TimeLimit reset
template <int ID>
__device__ __forceinline__ void env_reset(Env<Cfg<ID>::NB>& e, const Ranges& rg) {
    using C = Cfg<ID>;
    e.draws0 = 0; e.draws1 = 0;
    uint32_t ep = e.episode;
    if (C::VAR) {
        U4 w = env_draw(e, 0, ep);
        int num_grey = (int)__umulhi(w.x, 3u);  // np_random.randint(3), :647
        e.nb = 2 + num_grey;
        e.touch_now = 0; e.touch_ever = 0;      // achieved_goal = -1, :663
    }
    sim_init<C::NB>(e, ID == 2);
    randomize_objects<ID>(e, ep, false, rg);
    e.succ = 0;
    e.t = 0;
    e.episode = ep + 1;
    e.priv = spawn_priv<ID>(e);
}

// -(d != c).astype(float32): -1.0 or -0.0 (appendix A7).  Stored as an integer bit pattern: nvcc
// folds `cond ? -1.0f : -0.0f` (even through __uint_as_float) into an int->float convert that
// loses the sign of zero.
__device__ __forceinline__ void store_reward(float* dst, bool fail) {
    *reinterpret_cast<uint32_t*>(dst) = fail ? 0xbf800000u : 0x80000000u;
}

// reward of the env's own touch matrix against its goal: fetch_env.py:135-143 on
// ag in {-1,0,1}: d = sum(ag*g), c = count_nonzero(g); both symmetric entries counted.
template <int ID>
__device__ __forceinline__ bool env_reward_fail(uint32_t now, uint32_t ever) {
    using C = Cfg<ID>;
    int sp = __popc(now & C::P) - __popc(~ever & C::P);   // sum of ag over the +1 goal pairs
    int sm = __popc(now & C::M) - __popc(~ever & C::M);   // sum of ag over the -1 goal pairs
    int d = 2 * (sp - sm);
    int c = 2 * (__popc(C::P) + __popc(C::M));
    return d != c;
}

// value of touch-matrix entry (i, j) as the reference stores it
__device__ __forceinline__ float ag_value(uint32_t now, uint32_t ever, int i, int j) {
    if (i == j) return -1.0f;  // never set: no self contacts (fetch_env.py:78)
    int lo = i < j ? i : j, hi = i < j ? j : i;
    uint32_t b = 1u << pair_index(lo, hi);
    return (now & b) ? 1.0f : ((ever & b) ? 0.0f : -1.0f);
}

// _get_obs (fetch_env.py:187-228; Variation :567-621) into a row of DIMO floats with stride `os`
template <int ID, typename Store>
__device__ __forceinline__ void env_write_obs(const Env<Cfg<ID>::NB>& e, Store&& put, const float* yaw = nullptr) {
    using C = Cfg<ID>;
    constexpr int NB = C::NB;
    float gvp0 = e.gv[0] * kDt, gvp1 = e.gv[1] * kDt, gvp2 = e.gv[2] * kDt;
    int o = 0;
    if (C::VAR) put(o++, (float)e.nb);
    put(o++, e.g[0]); put(o++, e.g[1]); put(o++, e.g[2]);
    put(o++, e.q[0]); put(o++, e.q[1]);
    put(o++, gvp0); put(o++, gvp1); put(o++, gvp2);
    put(o++, e.qv[0] * kDt); put(o++, e.qv[1] * kDt);
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        constexpr int per = C::VAR ? 19 : 15;
        const bool live = !C::VAR || i < e.nb;   // only the Variation env has fewer live cubes than slots
        put(o + 0, live ? e.px[i] : 0.0f);
        put(o + 1, live ? e.py[i] : 0.0f);
        put(o + 2, live ? e.pz[i] : 0.0f);
        put(o + 3, live ? e.px[i] - e.g[0] : 0.0f);
        put(o + 4, live ? e.py[i] - e.g[1] : 0.0f);
        put(o + 5, live ? e.pz[i] - e.g[2] : 0.0f);
        put(o + 6, live ? __uint_as_float(0x80000000u) : 0.0f);   // roll: mat2euler gives -arctan2(0, 1) = -0.0 (fetch_env.py:205)
        put(o + 7, 0.0f);
        put(o + 8, live ? (yaw ? yaw[i] : bp_atan2(e.s[i], e.c[i])) : 0.0f);   // yaw: the caller's cached bp_atan2(s, c), if it keeps one
        put(o + 9, live ? e.vx[i] * kDt - gvp0 : 0.0f);
        put(o + 10, live ? e.vy[i] * kDt - gvp1 : 0.0f);
        put(o + 11, live ? e.vz[i] * kDt - gvp2 : 0.0f);
        put(o + 12, 0.0f);
        put(o + 13, 0.0f);
        put(o + 14, live ? e.w[i] * kDt : 0.0f);
        if (C::VAR) {
            // one_hot_color of [GREEN, BLUE, GREY, GREY] (fetch_env.py:599,632-639): GREY=0, GREEN=2, BLUE=3
            const int col = i == 0 ? 2 : (i == 1 ? 3 : 0);
            put(o + 15, live && col == 0 ? 1.0f : 0.0f);
            put(o + 16, 0.0f);
            put(o + 17, live && col == 2 ? 1.0f : 0.0f);
            put(o + 18, live && col == 3 ? 1.0f : 0.0f);
        }
        o += per;
    }
}

template <int ID, typename Store>
__device__ __forceinline__ void env_write_ag(uint32_t now, uint32_t ever, Store&& put) {
    using C = Cfg<ID>;
    constexpr int N = C::VAR ? kMaxObjs : C::NB + 2;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) put(i * N + j, ag_value(now, ever, i, j));
}

template <int ID, typename Store>
__device__ __forceinline__ void env_write_goal(Store&& put) {
    using C = Cfg<ID>;
    constexpr int N = C::VAR ? kMaxObjs : C::NB + 2;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            float v = 0.0f;
            if (i != j) {
                const int lo = i < j ? i : j, hi = i < j ? j : i;
                const uint32_t b = 1u << pair_index(lo, hi);
                v = (C::P & b) ? 1.0f : ((C::M & b) ? -1.0f : 0.0f);
            }
            put(i * N + j, v);
        }
}

// np.clip(action, -1, 1) (robot_env.py:58); NaN components are zeroed and counted, not raised
__device__ __forceinline__ void clip_action(float a[4], int& invalid) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float x = a[k];
        if (!(x == x)) { x = 0.0f; invalid += 1; }
        a[k] = x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x);
    }
}

// what RobotEnv.step does after sim.step(): _step_callback (fetch_env.py:148-167: 1 -> 0 downgrade,
// then contacts -> 1), reward (fetch_env.py:135-143), _is_success latch (fetch_env.py:275-281), TimeLimit count.
// Returns true when the reward is -1 (false: -0.0).
template <int ID>
__device__ __forceinline__ bool env_post_step(uint32_t contacts, uint32_t& touch_now, uint32_t& touch_ever, int& succ, int& t) {
    touch_now = contacts;
    touch_ever |= contacts;
    const bool fail = env_reward_fail<ID>(touch_now, touch_ever);
    if (!fail) succ = 1;
    t += 1;
    return fail;
}

}  // namespace bp
