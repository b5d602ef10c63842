// bp_kernels.cu -- CUDA kernels (sm_100a) + the C-ABI of include/blockpuzzle_b200.h.
//
// Device state layout: field-major SoA, state[f * B + e] (32-bit words), so that
// a warp of 32 consecutive envs reads/writes every field with one coalesced
// 128-byte transaction.  Fields: see load_env/store_env.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/blockpuzzle_b200.h"
#include "bp_device.cuh"
#include "bp_host.h"

namespace bp {

// ---------------------------------------------------------------- state <-> registers
template <int NB> __host__ __device__ constexpr int num_fields() { return 18 + 9 * NB; }

template <int NB>
__device__ __forceinline__ void load_env(const uint32_t* __restrict__ st, int64_t B, int64_t i, Env<NB>& e) {
    const uint32_t* p = st + i;
    int f = 0;
    auto ldf = [&]() { float v = __uint_as_float(p[(int64_t)f * B]); ++f; return v; };
    auto ldu = [&]() { uint32_t v = p[(int64_t)f * B]; ++f; return v; };
    e.g[0] = ldf(); e.g[1] = ldf(); e.g[2] = ldf();
    e.gv[0] = ldf(); e.gv[1] = ldf(); e.gv[2] = ldf();
    e.q[0] = ldf(); e.q[1] = ldf(); e.qv[0] = ldf(); e.qv[1] = ldf();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        e.px[b] = ldf(); e.py[b] = ldf(); e.pz[b] = ldf();
        e.c[b] = ldf(); e.s[b] = ldf();
        e.vx[b] = ldf(); e.vy[b] = ldf(); e.vz[b] = ldf(); e.w[b] = ldf();
    }
    uint32_t touch = ldu();
    e.touch_now = touch & 0xffffu; e.touch_ever = touch >> 16;
    uint32_t flags = ldu();
    e.t = (int)(flags & 0xffu); e.succ = (int)((flags >> 8) & 1u); e.nb = (int)((flags >> 9) & 7u);
    e.priv = ldu();
    e.episode = ldu(); e.draws0 = ldu(); e.draws1 = ldu();
    e.key0 = ldu(); e.key1 = ldu();
    e.contacts = 0;
}

template <int NB>
__device__ __forceinline__ void store_env(uint32_t* __restrict__ st, int64_t B, int64_t i, const Env<NB>& e) {
    uint32_t* p = st + i;
    int f = 0;
    auto stf = [&](float v) { p[(int64_t)f * B] = __float_as_uint(v); ++f; };
    auto stu = [&](uint32_t v) { p[(int64_t)f * B] = v; ++f; };
    stf(e.g[0]); stf(e.g[1]); stf(e.g[2]);
    stf(e.gv[0]); stf(e.gv[1]); stf(e.gv[2]);
    stf(e.q[0]); stf(e.q[1]); stf(e.qv[0]); stf(e.qv[1]);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        stf(e.px[b]); stf(e.py[b]); stf(e.pz[b]);
        stf(e.c[b]); stf(e.s[b]);
        stf(e.vx[b]); stf(e.vy[b]); stf(e.vz[b]); stf(e.w[b]);
    }
    stu(e.touch_now | (e.touch_ever << 16));
    stu((uint32_t)(e.t < 255 ? e.t : 255) | ((uint32_t)e.succ << 8) | ((uint32_t)e.nb << 9));   // t saturates: any t >= T behaves alike (TimeLimit keeps done = True)
    stu(e.priv);
    stu(e.episode); stu(e.draws0); stu(e.draws1);
    stu(e.key0); stu(e.key1);
}

// ---------------------------------------------------------------- kernels
// construction: what BlocksEnv.__init__ leaves (fetch_env.py:75-86, robot_env.py:33-37)
template <int ID>
__global__ void init_kernel(uint32_t* st, int64_t B) {
    using C = Cfg<ID>;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    Env<C::NB> e;
    e.nb = C::NB;
    sim_init<C::NB>(e, ID == 2);
    e.touch_now = 0; e.touch_ever = 0;  // achieved_goal = -1 everywhere, fetch_env.py:78
    e.t = 0; e.succ = 0; e.episode = 0; e.draws0 = 0; e.draws1 = 0; e.key0 = 0; e.key1 = 0; e.priv = 0;
    store_env<C::NB>(st, B, i, e);
}

// RolloutStudent.seed: env idx gets seed + 1000*idx (rollout.py:206-210)
__global__ void seed_kernel(uint32_t* st, int64_t B, int nf, uint64_t seed, uint64_t env_offset) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    uint64_t s = seed + 1000ull * (env_offset + (uint64_t)i);
    st[(int64_t)(nf - 5) * B + i] = 0;  // episode
    st[(int64_t)(nf - 4) * B + i] = 0;  // draws0
    st[(int64_t)(nf - 3) * B + i] = 0;  // draws1
    st[(int64_t)(nf - 2) * B + i] = (uint32_t)s;
    st[(int64_t)(nf - 1) * B + i] = (uint32_t)(s >> 32);
}

template <int ID>
__device__ __forceinline__ void write_row_obs(const Env<Cfg<ID>::NB>& e, float* row) {
    env_write_obs<ID>(e, [&](int k, float v) { row[k] = v; });
}
template <int ID>
__device__ __forceinline__ void write_row_ag(const Env<Cfg<ID>::NB>& e, float* row) {
    env_write_ag<ID>(e.touch_now, e.touch_ever, [&](int k, float v) { row[k] = v; });
}
template <int ID>
__device__ __forceinline__ void write_row_goal(float* row) {
    env_write_goal<ID>([&](int k, float v) { row[k] = v; });
}

template <int ID>
__global__ void reset_kernel(uint32_t* st, int64_t B, const uint8_t* mask, Ranges rg, float* obs, float* ag, float* g, int64_t oag_stride) {
    using C = Cfg<ID>;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (mask && !mask[i]) return;
    Env<C::NB> e;
    load_env<C::NB>(st, B, i, e);
    env_reset<ID>(e, rg);
    store_env<C::NB>(st, B, i, e);
    if (obs) write_row_obs<ID>(e, obs + i * oag_stride * C::DIMO);  // oag_stride rows between envs (1, or T + 1 for episode tensors)
    if (ag) write_row_ag<ID>(e, ag + i * oag_stride * C::DIMG);
    if (g) write_row_goal<ID>(g + i * C::DIMG);
}

// set_test: fetch_env.py:365-368, 443-446 (stale obs, appendix A4); Variation :641-644
template <int ID>
__global__ void set_test_kernel(uint32_t* st, int64_t B, Ranges rg, float* obs, float* ag, float* g, int64_t oag_stride) {
    using C = Cfg<ID>;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    Env<C::NB> e;
    load_env<C::NB>(st, B, i, e);
    if (obs) write_row_obs<ID>(e, obs + i * oag_stride * C::DIMO);
    if (ag) write_row_ag<ID>(e, ag + i * oag_stride * C::DIMG);
    if (g) write_row_goal<ID>(g + i * C::DIMG);
    if (!C::VAR) {
        randomize_objects<ID>(e, e.episode - 1u, true, rg);
        e.priv = 0;
        store_env<C::NB>(st, B, i, e);
    }
}

struct StepArgs {
    const float* actions;  // [K][B][4] or null
    float* obs;            // [K][B][DIMO]
    float* ag;             // [K][B][DIMG]
    float* reward;         // [K][B]
    float* success;        // [K][B]
    uint8_t* done;         // [K][B]
    float* reset_obs;      // [B][DIMO]
    float* reset_ag;       // [B][DIMG]
    float* actions_out;    // [K][B][4]
    double* stats;
    int64_t B;             // envs in this launch
    int64_t stateB;        // stride of the state arrays (envs in the handle)
    int64_t env0;          // first env of this launch inside the handle
    int K;
    int auto_reset;
    Ranges rg;
    // output addressing.  layout 0 (time-major): row (k0 + k) * B + li for every tensor.
    // layout 1 (batch-major episode, async kernel only): per-step tensors [B][Ktot] -> row li * Ktot + k0 + k;
    // obs / ag are [B][Ktot + 1] with the reset observation in slot 0 -> row li * (Ktot + 1) + k0 + k + 1.
    int layout;
    int k0;
    int act_k0;            // row of `actions` that holds step k0's actions (== k0, or 0 for the per-step tensor of bp_rollout_step)
    int Ktot;
    float* goal_out;       // layout 1: desired_goal rows [B][Ktot][DIMG] (nullable)
    int tune;              // async kernel: queue thresholds (see launch_step) | kTuneForceFull
};

__device__ __forceinline__ int64_t step_row(const StepArgs& p, int k, int64_t li) {
    return p.layout ? li * p.Ktot + (p.k0 + k) : (int64_t)(p.k0 + k) * p.B + li;
}
__device__ __forceinline__ int64_t obs_row(const StepArgs& p, int k, int64_t li) {
    return p.layout ? li * (p.Ktot + 1) + (p.k0 + k) + 1 : (int64_t)(p.k0 + k) * p.B + li;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Env<NB> (registers) <-> Grip + private cube column
template <int NB, int STRIDE>
__device__ __forceinline__ void env_to_col(const Env<NB>& e, Grip& g, const Col<NB, STRIDE> col) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { g.g[d] = e.g[d]; g.gv[d] = e.gv[d]; }
    g.q[0] = e.q[0]; g.q[1] = e.q[1]; g.qv[0] = e.qv[0]; g.qv[1] = e.qv[1];
#pragma unroll
    for (int b = 0; b < NB; ++b) col.store(b, Blk{e.px[b], e.py[b], e.pz[b], e.c[b], e.s[b], e.vx[b], e.vy[b], e.vz[b], e.w[b]});
}
template <int NB, int STRIDE>
__device__ __forceinline__ void col_to_env(Env<NB>& e, const Grip& g, const Col<NB, STRIDE> col) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { e.g[d] = g.g[d]; e.gv[d] = g.gv[d]; }
    e.q[0] = g.q[0]; e.q[1] = g.q[1]; e.qv[0] = g.qv[0]; e.qv[1] = g.qv[1];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const Blk k = col.load(b);
        e.px[b] = k.x; e.py[b] = k.y; e.pz[b] = k.z; e.c[b] = k.c; e.s[b] = k.s; e.vx[b] = k.vx; e.vy[b] = k.vy; e.vz[b] = k.vz; e.w[b] = k.w;
    }
}

// Reference kernel: one thread per env, every step runs the full physics (no quiet path),
// direct stores.  Selected with BP_STEP_KERNEL=simple; the tests use it to cross-check the
// tiled kernel's quiet path bit for bit.
template <int ID>
__global__ void __launch_bounds__(128) step_kernel_simple(uint32_t* st, StepArgs p) {
    using C = Cfg<ID>;
    constexpr int NB = C::NB;
    extern __shared__ __align__(16) uint32_t smem[];
    const Col<NB, 128> col(reinterpret_cast<float*>(smem) + threadIdx.x);
    int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // index inside this launch
    const bool live = li < p.B;
    float n_ep = 0.f, n_su = 0.f, n_st = 0.f, n_inv = 0.f, r_sum = 0.f;
    if (live) {
        const int64_t i = p.env0 + li;
        Env<NB> e;
        load_env<NB>(st, p.stateB, i, e);
        for (int k = 0; k < p.K; ++k) {
            const int64_t row = (int64_t)k * p.B + li;
            float4 a4;
            if (p.actions) {
                a4 = reinterpret_cast<const float4*>(p.actions)[row];
            } else {
                U4 w = philox4x32((uint32_t)e.t, e.episode - 1u, 2u, 0u, e.key0, e.key1);
                a4 = make_float4(2.0f * u01(w.x) - 1.0f, 2.0f * u01(w.y) - 1.0f, 2.0f * u01(w.z) - 1.0f, 2.0f * u01(w.w) - 1.0f);
            }
            if (p.actions_out) reinterpret_cast<float4*>(p.actions_out)[row] = a4;
            int inv = 0;
            float a[4] = {a4.x, a4.y, a4.z, a4.w};
            clip_action(a, inv);
            Grip g;
            env_to_col<NB, 128>(e, g, col);
            sim_step_col<NB, 128, C::BG>(g, a, col, e.nb, e.contacts);
            col_to_env<NB, 128>(e, g, col);
            e.priv = 0;
            const bool fail = env_post_step<ID>(e.contacts, e.touch_now, e.touch_ever, e.succ, e.t);
            if (p.obs) write_row_obs<ID>(e, p.obs + row * C::DIMO);
            if (p.ag) write_row_ag<ID>(e, p.ag + row * C::DIMG);
            if (p.reward) store_reward(p.reward + row, fail);
            if (p.success) p.success[row] = (float)e.succ;
            const bool done = e.t >= kT;
            if (p.done) p.done[row] = done ? 1 : 0;
            n_st += 1.f; n_inv += (float)inv; r_sum += fail ? -1.f : 0.f;
            if (done) {
                n_ep += 1.f; n_su += (float)e.succ;
                if (p.auto_reset) {
                    env_reset<ID>(e, p.rg);
                    if (p.reset_obs) write_row_obs<ID>(e, p.reset_obs + li * C::DIMO);
                    if (p.reset_ag) write_row_ag<ID>(e, p.reset_ag + li * C::DIMG);
                }
            }
        }
        store_env<NB>(st, p.stateB, i, e);
    }
    n_ep = warp_sum(n_ep); n_su = warp_sum(n_su); n_st = warp_sum(n_st); n_inv = warp_sum(n_inv); r_sum = warp_sum(r_sum);
    if ((threadIdx.x & 31) == 0 && p.stats) {
        if (n_ep != 0.f) atomicAdd(p.stats + BP_STAT_EPISODES, (double)n_ep);
        if (n_su != 0.f) atomicAdd(p.stats + BP_STAT_SUCCESSES, (double)n_su);
        if (n_st != 0.f) atomicAdd(p.stats + BP_STAT_STEPS, (double)n_st);
        if (n_inv != 0.f) atomicAdd(p.stats + BP_STAT_INVALID, (double)n_inv);
        if (r_sum != 0.f) atomicAdd(p.stats + BP_STAT_REWARD_SUM, (double)r_sum);
    }
}


}  // namespace bp

#include "bp_async.cuh"
#include "bp_split.cuh"

namespace bp {

template <int ID>
__global__ void get_state_kernel(const uint32_t* st, int64_t B, bp_env_state* out) {
    using C = Cfg<ID>;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    Env<C::NB> e;
    load_env<C::NB>(st, B, i, e);
    bp_env_state s;
    memset(&s, 0, sizeof(s));
    for (int k = 0; k < 3; ++k) { s.grip_pos[k] = e.g[k]; s.grip_vel[k] = e.gv[k]; }
    for (int k = 0; k < 2; ++k) { s.finger_q[k] = e.q[k]; s.finger_qv[k] = e.qv[k]; }
#pragma unroll
    for (int b = 0; b < C::NB; ++b) {
        if (b < e.nb) {
            s.blk_pos[b][0] = e.px[b]; s.blk_pos[b][1] = e.py[b]; s.blk_pos[b][2] = e.pz[b];
            s.blk_cs[b][0] = e.c[b]; s.blk_cs[b][1] = e.s[b];
            s.blk_vel[b][0] = e.vx[b]; s.blk_vel[b][1] = e.vy[b]; s.blk_vel[b][2] = e.vz[b];
            s.blk_w[b] = e.w[b];
        }
    }
    constexpr int N = C::VAR ? kMaxObjs : C::NB + 2;
    for (int k = 0; k < BP_MAX_DIMG; ++k) s.ag[k] = -1;  // rows beyond N*N keep the constructor's -1
    for (int a = 0; a < N; ++a)
        for (int b = 0; b < N; ++b) s.ag[a * N + b] = (int8_t)ag_value(e.touch_now, e.touch_ever, a, b);
    s.num_objs = e.nb + 2;
    s.has_succeeded = e.succ;
    s.t = e.t;
    s.episode = e.episode;
    s.draws[0] = e.draws0; s.draws[1] = e.draws1;
    out[i] = s;
}

template <int ID>
__global__ void set_state_kernel(uint32_t* st, int64_t B, const bp_env_state* in) {
    using C = Cfg<ID>;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    Env<C::NB> e;
    load_env<C::NB>(st, B, i, e);  // keeps the Philox key
    const bp_env_state& s = in[i];
    for (int k = 0; k < 3; ++k) { e.g[k] = s.grip_pos[k]; e.gv[k] = s.grip_vel[k]; }
    for (int k = 0; k < 2; ++k) { e.q[k] = s.finger_q[k]; e.qv[k] = s.finger_qv[k]; }
#pragma unroll
    for (int b = 0; b < C::NB; ++b) {
        e.px[b] = s.blk_pos[b][0]; e.py[b] = s.blk_pos[b][1]; e.pz[b] = s.blk_pos[b][2];
        e.c[b] = s.blk_cs[b][0]; e.s[b] = s.blk_cs[b][1];
        e.vx[b] = s.blk_vel[b][0]; e.vy[b] = s.blk_vel[b][1]; e.vz[b] = s.blk_vel[b][2];
        e.w[b] = s.blk_w[b];
    }
    constexpr int N = C::VAR ? kMaxObjs : C::NB + 2;
    e.touch_now = 0; e.touch_ever = 0;
    for (int a = 0; a < N; ++a)
        for (int b = a + 1; b < N; ++b) {
            int v = s.ag[a * N + b];
            uint32_t bit = 1u << pair_index(a, b);
            if (v == 1) { e.touch_now |= bit; e.touch_ever |= bit; }
            else if (v == 0) e.touch_ever |= bit;
        }
    e.nb = s.num_objs - 2;
    e.succ = s.has_succeeded;
    e.t = s.t;
    e.episode = s.episode;
    e.draws0 = s.draws[0]; e.draws1 = s.draws[1];
    e.priv = 0;
    store_env<C::NB>(st, B, i, e);
}

// compute_reward for dimg = 4 * CPR with CPR a power of two (16: BlocksTouch): CPR consecutive lanes share a row, each
// loads ONE float4 of ag and g (a warp reads 512 contiguous bytes per tensor and instruction), and the dot product
// is accumulated in the reference's left-to-right order by handing the running sum from lane to lane.
template <int CPR>
__global__ void __launch_bounds__(256) compute_reward_coop_kernel(const float4* __restrict__ ag, const float4* __restrict__ g, int64_t n,
                                                                  float* __restrict__ r) {
    const int64_t total = n * CPR;
    const int chunk = (int)(threadIdx.x & (CPR - 1));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i - chunk < total; i += (int64_t)gridDim.x * blockDim.x) {
        const bool live = i < total;   // rows never straddle the end: total is a multiple of CPR
        const float4 x = live ? __ldcs(ag + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 y = live ? __ldcs(g + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        int c = (y.x != 0.0f) + (y.y != 0.0f) + (y.z != 0.0f) + (y.w != 0.0f);
        float d = 0.0f;
#pragma unroll
        for (int j = 0; j < CPR; ++j) {
            const float prev = __shfl_up_sync(0xffffffffu, d, 1, CPR);
            if (chunk == j) {
                d = j == 0 ? 0.0f : prev;
                d = d + x.x * y.x; d = d + x.y * y.y; d = d + x.z * y.z; d = d + x.w * y.w;
            }
        }
#pragma unroll
        for (int o = CPR / 2; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o, CPR);
        if (live && chunk == CPR - 1) store_reward(r + i / CPR, d != (float)c);
    }
}

// BlocksEnv.compute_reward (fetch_env.py:135-143): one thread per row, 128-bit loads when dimg % 4 == 0
__global__ void compute_reward_kernel(const float* __restrict__ ag, const float* __restrict__ g, int64_t n, int dimg, float* __restrict__ r) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = ag + i * dimg;
    const float* b = g + i * dimg;
    float d = 0.0f;
    int c = 0;
    if ((dimg & 3) == 0) {
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        for (int k = 0; k < dimg / 4; ++k) {
            float4 x = __ldg(a4 + k), y = __ldg(b4 + k);
            d = d + x.x * y.x; d = d + x.y * y.y; d = d + x.z * y.z; d = d + x.w * y.w;
            c += (y.x != 0.0f) + (y.y != 0.0f) + (y.z != 0.0f) + (y.w != 0.0f);
        }
    } else {
        for (int k = 0; k < dimg; ++k) {
            float x = __ldg(a + k), y = __ldg(b + k);
            d = d + x * y;
            c += (y != 0.0f);
        }
    }
    store_reward(r + i, d != (float)c);
}


}  // namespace bp

// ======================================================================= C-ABI
using namespace bp;

static thread_local std::string g_err;
int bp_fail(int code, const std::string& msg) { g_err = msg; return code; }
static int fail(int code, const std::string& msg) { return bp_fail(code, msg); }
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(BP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));      \
    } while (0)

// Entry points run on the handle's device and leave the caller's current device as they found it (torch reads it
// through cudaGetDevice: a library that switches it silently redirects the caller's later allocations).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) err = cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(h)                                                                          \
    DeviceGuard _dg((h)->device);                                                             \
    if (_dg.err != cudaSuccess) return fail(BP_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(_dg.err))

static const char* kNames[BP_NUM_ENV_IDS] = {
    "GripperTouch-v0", "BlocksTouch-v0", "ToppleTower-v0", "BlocksTouchCurriculum-v0",
    "BlocksTouchChoose-v0", "BlocksTouchChooseCurriculum-v0", "BlocksTouchVariation-v0"};
static const int kNB[BP_NUM_ENV_IDS] = {1, 2, 4, 2, 3, 3, 4};
static const int kDimO[BP_NUM_ENV_IDS] = {25, 40, 70, 40, 55, 55, 87};
static const int kDimG[BP_NUM_ENV_IDS] = {9, 16, 36, 16, 25, 25, 36};

struct bp_handle {
    int env_id = 0;
    int device = 0;
    int64_t B = 0;
    uint64_t env_offset = 0;
    int nf = 0;
    uint32_t* d_state = nullptr;
    double* d_stats = nullptr;
    // curriculum knobs: python floats of fetch_env.py:340-348, 404-415, 561-563
    double obj_range = 0.15, obj_range_step = 0, max_obj_range = 0.15;
    double wrong_obj_range = 0, wrong_obj_range_step = 0;
    bool has_step = false, has_curriculum = false;
    int difficulty = 0;
    // bp_step_host staging
    cudaStream_t hs[2] = {nullptr, nullptr};
    float* d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    cudaEvent_t ev = nullptr;       // orders the staging streams after the caller's stream
    int challenge = 0;              // bp_set_option("challenge"): BlocksTouchChooseEnv(challenge=True), fetch_env.py:403,416
    int force_full = 0;             // bp_set_option("force_full_physics"): every env-step takes the full-physics pass
    int step_kernel = -1;           // bp_set_option("step_kernel"): -1 default (BP_STEP_KERNEL, else async), 0 async, 2 simple, 4 split
    // step-synchronous path (bp_split.cuh): per-step work lists and per-block statistics slots
    int32_t* d_lists = nullptr;     // [2][B]
    int32_t* d_counts = nullptr;    // [4]
    double* d_part = nullptr;
    int part_slots = 0;
};

static Ranges ranges_of(const bp_handle* h) {
    return Ranges{h->obj_range, h->max_obj_range, h->wrong_obj_range, h->challenge};
}

template <typename F>
static int dispatch(int env_id, F&& f) {
    switch (env_id) {
        case 0: return f(std::integral_constant<int, 0>());
        case 1: return f(std::integral_constant<int, 1>());
        case 2: return f(std::integral_constant<int, 2>());
        case 3: return f(std::integral_constant<int, 3>());
        case 4: return f(std::integral_constant<int, 4>());
        case 5: return f(std::integral_constant<int, 5>());
        case 6: return f(std::integral_constant<int, 6>());
        default: return fail(BP_ERR_INVALID_ARG, "unknown env id");
    }
}

static inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// BP_STEP_KERNEL = async (default) | split (bp_split.cuh) | simple (full physics for every env-step: the cross-check of the quiet path and
// the scheduler)
static int step_kernel_choice() {
    static const int v = [] {
        const char* e = getenv("BP_STEP_KERNEL");
        if (e && strcmp(e, "simple") == 0) return 2;
        if (e && strcmp(e, "async") == 0) return 0;
        if (e && strcmp(e, "split") == 0) return 4;
        return 0;   // default: the slab-resident kernel (step_kernel_async); split measured 4.09e9 vs 5.59e9 env-steps/s (BlockPhys v2)
    }();
    return v;
}

// cudaFuncSetAttribute is per device: remember which devices a kernel has been configured on
struct AttrOnce {
    unsigned long long done = 0;   // bit d: set on device d (devices >= 64 are configured on every launch)
    bool need(int dev) const { return dev < 0 || dev >= 64 || !((done >> dev) & 1ull); }
    void mark(int dev) { if (dev >= 0 && dev < 64) done |= 1ull << dev; }
};

template <class K>
static int set_smem_attr(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(BP_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    return BP_OK;
}

static constexpr int kSplitFullBlocks = 148 * 8, kSplitResetBlocks = 148;

// the step-synchronous path: K x (quiet kernel, full-physics kernel, reset kernel) on the handle's global state
static int launch_step_split(bp_handle* h, StepArgs& a, cudaStream_t s) {
    const int q_blocks = (int)nblk(a.B, kSplitQuietThreads);
    const int need_slots = (int)nblk(h->B, kSplitQuietThreads) + kSplitFullBlocks;
    if (!h->d_lists) {
        CU(cudaMalloc(&h->d_lists, sizeof(int32_t) * 2 * (size_t)h->B));
        CU(cudaMalloc(&h->d_counts, sizeof(int32_t) * 4));
        CU(cudaMalloc(&h->d_part, sizeof(double) * 8 * (size_t)need_slots));
        CU(cudaMemsetAsync(h->d_part, 0, sizeof(double) * 8 * (size_t)need_slots, s));
        h->part_slots = need_slots;
    }
    CU(cudaMemsetAsync(h->d_counts, 0, sizeof(int32_t) * 4, s));
    SplitBufs sb{h->d_lists, h->d_lists + h->B, h->d_counts, h->d_part, q_blocks};
    a.tune = h->force_full ? kTuneForceFull : 0;
    int rc = dispatch(h->env_id, [&](auto id) {
        constexpr int ID = decltype(id)::value;
        static const int skip = [] { const char* e = getenv("BP_SPLIT_SKIP"); return e ? atoi(e) : 0; }();   // timing experiments only
        for (int k = 0; k < a.K; ++k) {
            const int set = k & 1;
            if (!(skip & 1)) split_quiet_kernel<ID><<<q_blocks, kSplitQuietThreads, 0, s>>>(h->d_state, a, k, sb, set);
            if (!(skip & 2)) split_full_kernel<ID><<<kSplitFullBlocks, kSplitFullThreads, 0, s>>>(h->d_state, a, k, sb, set);
            if (!(skip & 4)) split_reset_kernel<ID><<<kSplitResetBlocks, 128, 0, s>>>(h->d_state, a, sb, set);
        }
        return (int)BP_OK;
    });
    if (rc != BP_OK) return rc;
    split_stats_reduce_kernel<<<1, 256, 0, s>>>(h->d_part, q_blocks + kSplitFullBlocks, a.stats);
    CU(cudaGetLastError());
    return BP_OK;
}

static int launch_step(bp_handle* h, StepArgs& a, cudaStream_t s) {
    // rows are read and written with 128-bit accesses: a tensor that is not 16-byte aligned would be a sticky device fault
    if ((reinterpret_cast<uintptr_t>(a.actions) | reinterpret_cast<uintptr_t>(a.obs) | reinterpret_cast<uintptr_t>(a.ag) |
         reinterpret_cast<uintptr_t>(a.goal_out) | reinterpret_cast<uintptr_t>(a.reset_obs) | reinterpret_cast<uintptr_t>(a.reset_ag) |
         reinterpret_cast<uintptr_t>(a.actions_out)) & 15)
        return fail(BP_ERR_INVALID_ARG, "step tensors must be 16-byte aligned");
    if ((h->step_kernel >= 0 ? h->step_kernel : step_kernel_choice()) == 4) return launch_step_split(h, a, s);
    static const size_t pad = [] { const char* e = getenv("BP_SMEM_PAD"); return e ? (size_t)atoi(e) : (size_t)0; }();  // occupancy experiments
    int rc = dispatch(h->env_id, [&](auto id) {
        constexpr int ID = decltype(id)::value;
        const int choice = h->step_kernel >= 0 ? h->step_kernel : step_kernel_choice();
        if (a.layout != 0 && choice != 0) return fail(BP_ERR_INVALID_ARG, "the batch-major episode layout needs the async step kernel");
        if (choice == 2) {
            constexpr size_t kSmem = sizeof(float) * Col<Cfg<ID>::NB, 128>::kFields * 128;
            step_kernel_simple<ID><<<nblk(a.B, 128), 128, kSmem, s>>>(h->d_state, a);
        } else {
            // envs per lane (tuning).  Measured at the 18 KB slab, resident warps in brackets: E = 2 [20] 2.22e9, 3 [15] 3.16e9,
            // 4 [12] 3.28e9, 5 [10] 3.23e9, 6 [8] 3.08e9 env-steps/s
            // per id (E = 4 / 3 / 2): GripperTouch 3.78 / 3.97 / 3.83e9, ToppleTower 1.32 / 1.34 / 1.22e9, Variation 1.42 / 1.46 / 1.41e9
            // (E is a build-time constant: -DBP_ASYNC_E=3 for sweeps.  BlockPhys v2 kernel, E = 3 [15 warps] / 4 [12]: 4.65 / 4.87e9)
            // the lean instantiation serves the plain fused step (see step_kernel_async)
            const bool lean = a.layout == 0 && a.actions && !a.actions_out && !a.done && !a.goal_out && !a.reset_obs && !a.reset_ag &&
                              a.B * (int64_t)kMaxFused < (int64_t)1 << 31 &&   // 32-bit row indices inside the lean kernel
                              ((reinterpret_cast<uintptr_t>(a.obs) | reinterpret_cast<uintptr_t>(a.ag)) & 31) == 0;   // ... which writes 32-byte-multiple rows with unchecked 256-bit stores
            auto go = [&](auto ec, auto lc) -> int {
                constexpr int E = decltype(ec)::value;
                constexpr bool LEAN = decltype(lc)::value;
                using A = Async<ID, E>;
                static AttrOnce once;
                if (once.need(h->device)) {
                    int r = set_smem_attr(step_kernel_async<ID, E, LEAN>, A::SMEM + pad);
                    if (r != BP_OK) return r;
                    once.mark(h->device);
                }
                // the kernel keeps the reward / success bits of at most kMaxFused steps on chip: split longer K
                // reset-pass threshold | full-physics-pass threshold << 8 | fill rule (tuning knobs)
                constexpr int kPassMin = ID == 0 ? 3 : ((ID == 4 || ID == 5) ? 16 : 28);   // see the sweep quoted below
                static const int tune = [pm = kPassMin] {
                    const char* r = getenv("BP_RESET_MIN"); const char* q = getenv("BP_PASS_MIN");
                    const char* fr = getenv("BP_FILL_RULE");   // margin of the fill-comparison pass trigger, -1: off
                    const int frv = fr ? atoi(fr) : 8;
                    // per-id pass threshold (sweep of round 2, tools/sweep_ids_thresholds*.sh): GripperTouch-v0's passes are cheap (one cube, two finger
                    // slots) and rare (3.6 % of env-steps), so waiting for a full warp only stretches the chain of the warp's hardest env:
                    // 28 / 16 / 8 / 4 / 3 / 2 / 1 -> 4.38 / 4.49 / 5.08 / 7.32 / 8.04 / 7.75 / 6.55e9; the Choose ids (3 cubes): 28 / 20 / 16 / 12 / 8 ->
                    // 1.88 / 1.96 / 1.99 / 1.93 / 1.74e9; ToppleTower / Variation: flat from 16 to 28
                    return (r ? atoi(r) : 32) | ((q ? atoi(q) : pm) << 8) | (frv >= 0 ? (1 << 16) | (frv << 17) : 0);   // measured (round 2, lean kernel): reset 4 / 32 -> 4.33 / 4.48e9 at pass 24; pass 20 / 24 / 28 -> 4.44 / 4.48 / 4.46e9 (every reset pass streams ~12 KB of cold code through the instruction cache); BlockPhys v2 kernel: pass 16 / 20 / 24 / 28 / 32 -> 4.48 / 4.74 / 4.86 / 4.92 / 4.84e9
                }();
                for (int k0 = 0; k0 < a.K; k0 += kMaxFused) {
                    StepArgs c = a;
                    c.tune = tune | (h->force_full ? kTuneForceFull : 0);
                    c.K = (a.K - k0) < kMaxFused ? (a.K - k0) : kMaxFused;
                    c.k0 = a.k0 + k0;
                    c.act_k0 = a.act_k0 + k0;
                    if (LEAN) {   // the lean kernel addresses rows from the launch's first step: advance the tensors, k0 = act_k0 = 0
                        const int64_t r0 = (int64_t)c.k0 * a.B;
                        c.actions = a.actions + (int64_t)c.act_k0 * a.B * 4;
                        if (c.obs) c.obs = a.obs + r0 * Cfg<ID>::DIMO;
                        if (c.ag) c.ag = a.ag + r0 * Cfg<ID>::DIMG;
                        if (c.reward) c.reward = a.reward + r0;
                        if (c.success) c.success = a.success + r0;
                        c.k0 = 0; c.act_k0 = 0;
                    }
                    step_kernel_async<ID, E, LEAN><<<nblk(a.B, A::CS), 32, A::SMEM + pad, s>>>(h->d_state, c);
                }
                return (int)BP_OK;
            };
            using T_ = std::true_type; using F_ = std::false_type;
            auto go_e = [&](auto ec) -> int { return lean ? go(ec, T_()) : go(ec, F_()); };
#ifndef BP_E_ID0
#define BP_E_ID0 3
#endif
#ifndef BP_E_CHOOSE
#define BP_E_CHOOSE kAsyncE
#endif
            int r = go_e(std::integral_constant<int, (ID == 0 ? BP_E_ID0 : ((ID == 2 || ID == 6) ? 3 : ((ID == 4 || ID == 5) ? BP_E_CHOOSE : kAsyncE)))>());
            if (r != BP_OK) return r;
        }
        return (int)BP_OK;
    });
    if (rc != BP_OK) return rc;
    CU(cudaGetLastError());
    return BP_OK;
}

extern "C" {

int bp_abi_version(void) { return BP_ABI_VERSION; }
const char* bp_last_error(void) { return g_err.c_str(); }

int bp_env_dims(int env_id, int* dimo, int* dimg, int* nblocks) {
    if (env_id < 0 || env_id >= BP_NUM_ENV_IDS) return fail(BP_ERR_INVALID_ARG, "unknown env id");
    if (dimo) *dimo = kDimO[env_id];
    if (dimg) *dimg = kDimG[env_id];
    if (nblocks) *nblocks = kNB[env_id];
    return BP_OK;
}

int bp_env_id_from_name(const char* name) {
    if (!name) return fail(BP_ERR_INVALID_ARG, "null name");
    for (int i = 0; i < BP_NUM_ENV_IDS; ++i)
        if (strcmp(name, kNames[i]) == 0) return i;
    return fail(BP_ERR_INVALID_ARG, std::string("no registered env id ") + name);
}

const char* bp_env_name(int env_id) {
    if (env_id < 0 || env_id >= BP_NUM_ENV_IDS) return nullptr;
    return kNames[env_id];
}

int bp_create(int env_id, int64_t num_envs, int device, uint64_t env_index_offset, bp_handle** out) {
    if (!out) return fail(BP_ERR_INVALID_ARG, "out is null");
    if (env_id < 0 || env_id >= BP_NUM_ENV_IDS) return fail(BP_ERR_INVALID_ARG, "unknown env id");
    if (num_envs <= 0) return fail(BP_ERR_INVALID_ARG, "num_envs must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(BP_ERR_NO_DEVICE, "no CUDA device: blockpuzzle_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(BP_ERR_INVALID_ARG, "bad device index");
    DeviceGuard dg(device);
    if (dg.err != cudaSuccess) return fail(BP_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(dg.err));
    bp_handle* h = new bp_handle();
    h->env_id = env_id; h->device = device; h->B = num_envs; h->env_offset = env_index_offset;
    h->nf = 18 + 9 * kNB[env_id];
    switch (env_id) {  // fetch_env.py:340-348, 404-415, 561-563 on top of tasks.py obj_range=0.15
        case BP_BLOCKS_TOUCH: h->max_obj_range = h->obj_range; h->obj_range_step = 0; h->has_step = true; h->has_curriculum = true; break;
        case BP_BLOCKS_TOUCH_CURRICULUM:
        case BP_BLOCKS_TOUCH_VARIATION:
            h->obj_range = 0.08; h->obj_range_step = 0.025; h->max_obj_range = 0.2; h->has_step = true; h->has_curriculum = true; break;
        case BP_BLOCKS_TOUCH_CHOOSE: h->wrong_obj_range = 0; h->max_obj_range = 0.2; h->has_step = false; h->has_curriculum = true; break;
        case BP_BLOCKS_TOUCH_CHOOSE_CURRICULUM:
            h->obj_range = 0.08; h->obj_range_step = 0.025; h->wrong_obj_range = 0.2; h->wrong_obj_range_step = 0.02;
            h->max_obj_range = 0.3; h->has_step = true; h->has_curriculum = true; break;
        default: break;
    }
    cudaError_t e1 = cudaMalloc(&h->d_state, sizeof(uint32_t) * (size_t)h->nf * (size_t)num_envs);
    cudaError_t e2 = cudaMalloc(&h->d_stats, sizeof(double) * BP_NUM_STATS);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        cudaFree(h->d_state); cudaFree(h->d_stats); delete h;
        return fail(BP_ERR_CUDA, "cudaMalloc of env state failed");
    }
    cudaError_t e0 = cudaMemset(h->d_stats, 0, sizeof(double) * BP_NUM_STATS);
    if (e0 != cudaSuccess) { bp_destroy(h); return fail(BP_ERR_CUDA, std::string("cudaMemset of the statistics vector: ") + cudaGetErrorString(e0)); }
    int rc = dispatch(env_id, [&](auto id) {
        init_kernel<decltype(id)::value><<<nblk(num_envs, 128), 128>>>(h->d_state, num_envs);
        return BP_OK;
    });
    if (rc != BP_OK) { bp_destroy(h); return rc; }
    seed_kernel<<<nblk(num_envs, 128), 128>>>(h->d_state, num_envs, h->nf, 0, env_index_offset);
    cudaError_t e3 = cudaDeviceSynchronize();
    if (e3 != cudaSuccess) { bp_destroy(h); return fail(BP_ERR_CUDA, std::string("init: ") + cudaGetErrorString(e3)); }
    *out = h;
    return BP_OK;
}

int bp_destroy(bp_handle* h) {
    if (!h) return BP_OK;
    DeviceGuard dg(h->device);
    cudaFree(h->d_state);
    cudaFree(h->d_stats);
    for (int i = 0; i < 2; ++i) {
        if (h->d_stage[i]) cudaFree(h->d_stage[i]);
        if (h->hs[i]) cudaStreamDestroy(h->hs[i]);
    }
    if (h->ev) cudaEventDestroy(h->ev);
    cudaFree(h->d_lists); cudaFree(h->d_counts); cudaFree(h->d_part);
    delete h;
    return BP_OK;
}

int64_t bp_num_envs(const bp_handle* h) { return h ? h->B : 0; }

int bp_seed(bp_handle* h, uint64_t seed, void* stream) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    ON_DEVICE(h);
    seed_kernel<<<nblk(h->B, 128), 128, 0, (cudaStream_t)stream>>>(h->d_state, h->B, h->nf, seed, h->env_offset);
    CU(cudaGetLastError());
    return BP_OK;
}

int bp_reset(bp_handle* h, const uint8_t* d_mask, float* d_obs, float* d_ag, float* d_g, void* stream) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    ON_DEVICE(h);
    Ranges rg = ranges_of(h);
    int rc = dispatch(h->env_id, [&](auto id) {
        reset_kernel<decltype(id)::value><<<nblk(h->B, 128), 128, 0, (cudaStream_t)stream>>>(h->d_state, h->B, d_mask, rg, d_obs, d_ag, d_g, 1);
        return BP_OK;
    });
    if (rc != BP_OK) return rc;
    CU(cudaGetLastError());
    return BP_OK;
}

int bp_step(bp_handle* h, const float* d_actions, int K, float* d_obs, float* d_ag, float* d_reward,
            float* d_success, uint8_t* d_done, int auto_reset, float* d_reset_obs, float* d_reset_ag,
            float* d_actions_out, void* stream) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    if (K <= 0 || K > 65535) return fail(BP_ERR_INVALID_ARG, "K must be in 1..65535");
    ON_DEVICE(h);
    StepArgs a{};
    a.actions = d_actions; a.obs = d_obs; a.ag = d_ag; a.reward = d_reward; a.success = d_success; a.done = d_done;
    a.reset_obs = d_reset_obs; a.reset_ag = d_reset_ag; a.actions_out = d_actions_out; a.stats = h->d_stats;
    a.B = h->B; a.stateB = h->B; a.env0 = 0; a.K = K; a.auto_reset = auto_reset; a.rg = ranges_of(h);
    a.layout = 0; a.k0 = 0; a.act_k0 = 0; a.Ktot = K; a.goal_out = nullptr;
    return launch_step(h, a, (cudaStream_t)stream);
}

// reset_all_rollouts (rollout.py:48-64): reset(), then set_test() for test rollouts; slot 0 of the episode's o / ag
static int rollout_reset(bp_handle* h, int test, float* d_o, float* d_ag, float* d_g0, cudaStream_t s) {
    const int T = BP_MAX_EPISODE_STEPS;
    if (test && (h->env_id == BP_GRIPPER_TOUCH || h->env_id == BP_TOPPLE_TOWER))
        return fail(BP_ERR_NOT_IMPLEMENTED, "set_test raises NotImplementedError for this env (fetch_env.py:100-101)");
    Ranges rg = ranges_of(h);
    int rc = dispatch(h->env_id, [&](auto id) {
        constexpr int ID = decltype(id)::value;
        reset_kernel<ID><<<nblk(h->B, 128), 128, 0, s>>>(h->d_state, h->B, nullptr, rg, d_o, d_ag, d_g0, T + 1);
        return (int)BP_OK;
    });
    if (rc != BP_OK) return rc;
    CU(cudaGetLastError());
    if (test) {
        rc = dispatch(h->env_id, [&](auto id) {
            constexpr int ID = decltype(id)::value;
            set_test_kernel<ID><<<nblk(h->B, 128), 128, 0, s>>>(h->d_state, h->B, rg, d_o, d_ag, d_g0, T + 1);
            return (int)BP_OK;
        });
        if (rc != BP_OK) return rc;
        CU(cudaGetLastError());
    }
    return BP_OK;
}

int bp_rollout(bp_handle* h, const float* d_actions, int test, float* d_o, float* d_ag, float* d_g, float* d_u,
               float* d_success, float* d_reward, void* stream) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    if (!d_o || !d_ag) return fail(BP_ERR_INVALID_ARG, "bp_rollout needs the o and ag episode tensors");
    if (step_kernel_choice() == 2 || h->step_kernel == 2) return fail(BP_ERR_INVALID_ARG, "bp_rollout needs the async or the split step kernel");
    ON_DEVICE(h);
    const int T = BP_MAX_EPISODE_STEPS;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = rollout_reset(h, test, d_o, d_ag, nullptr, s);
    if (rc != BP_OK) return rc;
    StepArgs a{};
    a.actions = d_actions; a.obs = d_o; a.ag = d_ag; a.reward = d_reward; a.success = d_success; a.actions_out = d_u;
    a.goal_out = d_g; a.stats = h->d_stats;
    a.B = h->B; a.stateB = h->B; a.env0 = 0; a.K = T; a.auto_reset = 0; a.rg = ranges_of(h);
    a.layout = 1; a.k0 = 0; a.act_k0 = 0; a.Ktot = T;
    return launch_step(h, a, s);
}

int bp_rollout_begin(bp_handle* h, int test, float* d_o, float* d_ag, float* d_g0, void* stream) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    if (!d_o || !d_ag) return fail(BP_ERR_INVALID_ARG, "bp_rollout_begin needs the o and ag episode tensors");
    ON_DEVICE(h);
    return rollout_reset(h, test, d_o, d_ag, d_g0, (cudaStream_t)stream);
}

int bp_rollout_step(bp_handle* h, int t, const float* d_actions, float* d_o, float* d_ag, float* d_g, float* d_u,
                    float* d_success, float* d_reward, void* stream) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    if (t < 0 || t >= BP_MAX_EPISODE_STEPS) return fail(BP_ERR_INVALID_ARG, "t must be in 0..T-1");
    if (!d_actions || !d_o || !d_ag) return fail(BP_ERR_INVALID_ARG, "bp_rollout_step needs actions [B][4] and the o / ag episode tensors");
    if (step_kernel_choice() == 2 || h->step_kernel == 2) return fail(BP_ERR_INVALID_ARG, "bp_rollout_step needs the async or the split step kernel");
    ON_DEVICE(h);
    StepArgs a{};
    a.actions = d_actions; a.obs = d_o; a.ag = d_ag; a.reward = d_reward; a.success = d_success; a.actions_out = d_u;
    a.goal_out = d_g; a.stats = h->d_stats;
    a.B = h->B; a.stateB = h->B; a.env0 = 0; a.K = 1; a.auto_reset = 0; a.rg = ranges_of(h);
    a.layout = 1; a.k0 = t; a.act_k0 = 0; a.Ktot = BP_MAX_EPISODE_STEPS;
    return launch_step(h, a, (cudaStream_t)stream);
}

static inline size_t align4(size_t nfloats) { return (nfloats + 7) & ~(size_t)7; }   // sub-buffers start on 32 bytes (128- / 256-bit row stores)

int bp_step_host(bp_handle* h, const float* h_actions, int K, float* h_obs, float* h_ag, float* h_reward,
                 float* h_success, int auto_reset, void* stream) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    if (K <= 0 || K > 65535 || !h_actions) return fail(BP_ERR_INVALID_ARG, "bp_step_host needs actions and K in 1..65535");
    ON_DEVICE(h);
    const int dimo = kDimO[h->env_id], dimg = kDimG[h->env_id];
    // chunk of envs: per env and step 4 (action) + dimo + dimg + 2 floats
    const size_t per_env = (size_t)K * (size_t)(4 + dimo + dimg + 2) * sizeof(float);
    int64_t chunk = (int64_t)((size_t)(192u << 20) / per_env);
    if (chunk > h->B) chunk = h->B;
    chunk &= ~(int64_t)127;
    if (chunk < 128) chunk = h->B < 128 ? h->B : 128;
    const size_t need = per_env * (size_t)chunk + 5 * 32;   // + the alignment padding of the five sub-buffers
    if (h->stage_bytes < need) {
        for (int i = 0; i < 2; ++i) {
            if (h->d_stage[i]) cudaFree(h->d_stage[i]);
            h->d_stage[i] = nullptr;
            h->stage_bytes = 0;
            CU(cudaMalloc(&h->d_stage[i], need));
            if (!h->hs[i]) CU(cudaStreamCreateWithFlags(&h->hs[i], cudaStreamNonBlocking));
        }
        h->stage_bytes = need;
    }
    if (!h->ev) CU(cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming));
    // the staging streams start after whatever the caller has queued on its own stream (bp_seed / bp_reset /
    // bp_set_state / bp_step run there): without this the first step kernel could overtake a pending reset
    CU(cudaEventRecord(h->ev, (cudaStream_t)stream));
    CU(cudaStreamWaitEvent(h->hs[0], h->ev, 0));
    CU(cudaStreamWaitEvent(h->hs[1], h->ev, 0));
    // on any error both staging streams are drained before returning: no copy into the caller's arrays is left in flight
#define CUH(call)                                                                             \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            cudaStreamSynchronize(h->hs[0]); cudaStreamSynchronize(h->hs[1]);                 \
            return fail(BP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));      \
        }                                                                                     \
    } while (0)
    // the host arrays are [K][B][dim]; a chunk [K][n][dim] is strided -> 2-D copies
    int buf = 0;
    for (int64_t e0 = 0; e0 < h->B; e0 += chunk, buf ^= 1) {
        const int64_t n = (h->B - e0) < chunk ? (h->B - e0) : chunk;
        cudaStream_t s = h->hs[buf];
        float* base = h->d_stage[buf];
        float* d_act = base;
        float* d_obs = d_act + align4((size_t)K * n * 4);
        float* d_ag = d_obs + align4((size_t)K * n * dimo);
        float* d_r = d_ag + align4((size_t)K * n * dimg);
        float* d_s = d_r + align4((size_t)K * n);
        CUH(cudaMemcpy2DAsync(d_act, (size_t)n * 4 * 4, h_actions + e0 * 4, (size_t)h->B * 4 * 4, (size_t)n * 4 * 4, K, cudaMemcpyHostToDevice, s));
        StepArgs a{};
        a.actions = d_act; a.obs = h_obs ? d_obs : nullptr; a.ag = h_ag ? d_ag : nullptr;
        a.reward = h_reward ? d_r : nullptr; a.success = h_success ? d_s : nullptr;
        a.stats = h->d_stats; a.B = n; a.stateB = h->B; a.env0 = e0; a.K = K; a.auto_reset = auto_reset; a.rg = ranges_of(h);
        a.layout = 0; a.k0 = 0; a.act_k0 = 0; a.Ktot = K; a.goal_out = nullptr;
        int rc = launch_step(h, a, s);
        if (rc != BP_OK) { cudaStreamSynchronize(h->hs[0]); cudaStreamSynchronize(h->hs[1]); return rc; }
        if (h_obs) CUH(cudaMemcpy2DAsync(h_obs + e0 * dimo, (size_t)h->B * dimo * 4, d_obs, (size_t)n * dimo * 4, (size_t)n * dimo * 4, K, cudaMemcpyDeviceToHost, s));
        if (h_ag) CUH(cudaMemcpy2DAsync(h_ag + e0 * dimg, (size_t)h->B * dimg * 4, d_ag, (size_t)n * dimg * 4, (size_t)n * dimg * 4, K, cudaMemcpyDeviceToHost, s));
        if (h_reward) CUH(cudaMemcpy2DAsync(h_reward + e0, (size_t)h->B * 4, d_r, (size_t)n * 4, (size_t)n * 4, K, cudaMemcpyDeviceToHost, s));
        if (h_success) CUH(cudaMemcpy2DAsync(h_success + e0, (size_t)h->B * 4, d_s, (size_t)n * 4, (size_t)n * 4, K, cudaMemcpyDeviceToHost, s));
    }
#undef CUH
    cudaError_t s0 = cudaStreamSynchronize(h->hs[0]);
    cudaError_t s1 = cudaStreamSynchronize(h->hs[1]);
    if (s0 != cudaSuccess || s1 != cudaSuccess)
        return fail(BP_ERR_CUDA, std::string("bp_step_host: ") + cudaGetErrorString(s0 != cudaSuccess ? s0 : s1));
    return BP_OK;
}

int bp_set_test(bp_handle* h, float* d_obs, float* d_ag, float* d_g, void* stream) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    if (h->env_id == BP_GRIPPER_TOUCH || h->env_id == BP_TOPPLE_TOWER)
        return fail(BP_ERR_NOT_IMPLEMENTED, "set_test raises NotImplementedError for this env (fetch_env.py:100-101)");
    ON_DEVICE(h);
    Ranges rg = ranges_of(h);
    int rc = dispatch(h->env_id, [&](auto id) {
        set_test_kernel<decltype(id)::value><<<nblk(h->B, 128), 128, 0, (cudaStream_t)stream>>>(h->d_state, h->B, rg, d_obs, d_ag, d_g, 1);
        return BP_OK;
    });
    if (rc != BP_OK) return rc;
    CU(cudaGetLastError());
    return BP_OK;
}

int bp_increase_difficulty(bp_handle* h, int* max_reached) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    int ret = 0;
    switch (h->env_id) {
        case BP_BLOCKS_TOUCH:
        case BP_BLOCKS_TOUCH_CURRICULUM:
        case BP_BLOCKS_TOUCH_VARIATION:  // fetch_env.py:351-358, 623-630
            h->obj_range += h->obj_range_step;
            if (h->obj_range > h->max_obj_range) { h->obj_range = h->max_obj_range; ret = 1; }
            else h->difficulty += 1;
            break;
        case BP_BLOCKS_TOUCH_CHOOSE:
        case BP_BLOCKS_TOUCH_CHOOSE_CURRICULUM:  // fetch_env.py:419-432
            if (!h->has_step) return fail(BP_ERR_NO_ATTRIBUTE, "'BlocksTouchChooseEnv' object has no attribute 'obj_range_step' (fetch_env.py:413-415,420)");
            h->obj_range += h->obj_range_step;
            h->wrong_obj_range -= h->wrong_obj_range_step;
            if (h->obj_range > h->max_obj_range) {
                h->obj_range = h->max_obj_range;
                if (h->wrong_obj_range < 0) { h->wrong_obj_range = 0; ret = 1; break; }
            } else if (h->wrong_obj_range < 0) {
                h->wrong_obj_range = 0;
            }
            h->difficulty += 1;
            break;
        default:
            return fail(BP_ERR_NOT_IMPLEMENTED, "increase_difficulty raises NotImplementedError (fetch_env.py:93-94)");
    }
    if (max_reached) *max_reached = ret;
    return BP_OK;
}

int bp_set_option(bp_handle* h, const char* name, int value) {
    if (!h || !name) return fail(BP_ERR_INVALID_ARG, "null argument");
    if (strcmp(name, "force_full_physics") == 0) { h->force_full = value != 0; return BP_OK; }
    if (strcmp(name, "challenge") == 0) {   // the `challenge` constructor argument of BlocksTouchChooseEnv (fetch_env.py:403,416)
        if (h->env_id != 4 && h->env_id != 5) return fail(BP_ERR_INVALID_ARG, "challenge is an argument of BlocksTouchChooseEnv only");
        h->challenge = value != 0;
        return BP_OK;
    }
    if (strcmp(name, "step_kernel") == 0) {   // -1 default, 0 async (slab-resident, K steps fused in one kernel), 2 simple, 4 split
        if (value != -1 && value != 0 && value != 2 && value != 4) return fail(BP_ERR_INVALID_ARG, "step_kernel must be -1, 0, 2 or 4");
        h->step_kernel = value;
        return BP_OK;
    }
    return fail(BP_ERR_INVALID_ARG, std::string("unknown option ") + name);
}

int bp_get_difficulty(const bp_handle* h, int* difficulty) {
    if (!h || !difficulty) return fail(BP_ERR_INVALID_ARG, "null argument");
    *difficulty = h->difficulty;
    return BP_OK;
}

int bp_get_ranges(const bp_handle* h, double* obj_range, double* wrong_obj_range, double* max_obj_range) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    if (obj_range) *obj_range = h->obj_range;
    if (wrong_obj_range) *wrong_obj_range = h->wrong_obj_range;
    if (max_obj_range) *max_obj_range = h->max_obj_range;
    return BP_OK;
}

int bp_set_ranges(bp_handle* h, double obj_range, double wrong_obj_range) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    h->obj_range = obj_range;
    h->wrong_obj_range = wrong_obj_range;
    return BP_OK;
}

int bp_get_state(bp_handle* h, bp_env_state* d_out, void* stream) {
    if (!h || !d_out) return fail(BP_ERR_INVALID_ARG, "null argument");
    ON_DEVICE(h);
    int rc = dispatch(h->env_id, [&](auto id) {
        get_state_kernel<decltype(id)::value><<<nblk(h->B, 128), 128, 0, (cudaStream_t)stream>>>(h->d_state, h->B, d_out);
        return BP_OK;
    });
    if (rc != BP_OK) return rc;
    CU(cudaGetLastError());
    return BP_OK;
}

int bp_set_state(bp_handle* h, const bp_env_state* d_in, void* stream) {
    if (!h || !d_in) return fail(BP_ERR_INVALID_ARG, "null argument");
    ON_DEVICE(h);
    int rc = dispatch(h->env_id, [&](auto id) {
        set_state_kernel<decltype(id)::value><<<nblk(h->B, 128), 128, 0, (cudaStream_t)stream>>>(h->d_state, h->B, d_in);
        return BP_OK;
    });
    if (rc != BP_OK) return rc;
    CU(cudaGetLastError());
    return BP_OK;
}

int bp_stats_ptr(bp_handle* h, double** d_stats) {
    if (!h || !d_stats) return fail(BP_ERR_INVALID_ARG, "null argument");
    *d_stats = h->d_stats;
    return BP_OK;
}

int bp_stats_reset(bp_handle* h, void* stream) {
    if (!h) return fail(BP_ERR_INVALID_ARG, "null handle");
    ON_DEVICE(h);
    CU(cudaMemsetAsync(h->d_stats, 0, sizeof(double) * BP_NUM_STATS, (cudaStream_t)stream));
    return BP_OK;
}

int bp_compute_reward(const float* d_ag, const float* d_g, int64_t n, int dimg, float* d_r, void* stream) {
    if (n < 0 || dimg <= 0) return fail(BP_ERR_INVALID_ARG, "bad n or dimg");
    if (n == 0) return BP_OK;
    if (!d_ag || !d_g || !d_r) return fail(BP_ERR_INVALID_ARG, "null pointer");
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_ag) | reinterpret_cast<uintptr_t>(d_g)) & 15) == 0;
    const int cpr = dimg / 4;
    if (aligned && (dimg & 3) == 0 && cpr <= 8 && (cpr & (cpr - 1)) == 0) {
        int64_t blocks = (n * cpr + 255) / 256;
        if (blocks > 148 * 32) blocks = 148 * 32;   // grid-stride: a few waves of full blocks
        const float4* a4 = reinterpret_cast<const float4*>(d_ag);
        const float4* g4 = reinterpret_cast<const float4*>(d_g);
        cudaStream_t s = (cudaStream_t)stream;
        if (cpr == 1) compute_reward_coop_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(a4, g4, n, d_r);
        else if (cpr == 2) compute_reward_coop_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(a4, g4, n, d_r);
        else if (cpr == 4) compute_reward_coop_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(a4, g4, n, d_r);
        else compute_reward_coop_kernel<8><<<(unsigned)blocks, 256, 0, s>>>(a4, g4, n, d_r);
    } else {
        compute_reward_kernel<<<nblk(n, 256), 256, 0, (cudaStream_t)stream>>>(d_ag, d_g, n, dimg, d_r);
    }
    CU(cudaGetLastError());
    return BP_OK;
}

int bp_her_relabel(const float* d_ep_ag, const float* d_ep_g, int32_t B, int32_t T, int32_t dimg, int64_t n,
                   float future_p, uint64_t seed, int64_t index_offset, int32_t* d_ep_idx, int32_t* d_t,
                   int32_t* d_future_t, float* d_ag2, float* d_g, float* d_r, void* stream) {
    if (B <= 0 || T <= 0 || dimg <= 0 || n < 0) return fail(BP_ERR_INVALID_ARG, "bad sizes");
    if (n == 0) return BP_OK;
    if (!d_ep_ag || !d_ep_g) return fail(BP_ERR_INVALID_ARG, "null episode store");
    if (dimg > 256) return fail(BP_ERR_INVALID_ARG, "dimg > 256 is not supported (the registered ids have dimg <= 36)");
    // rows of 1 / 2 / 4 / 8 float4 chunks (dimg = 16: the BlocksTouch ids): the lane-cooperative kernel
    const int rc = bp_her_relabel_coop(d_ep_ag, d_ep_g, B, T, dimg, n, future_p, seed, index_offset, d_ep_idx, d_t, d_future_t, d_ag2, d_g, d_r, stream);
    if (rc != BP_ERR_NOT_IMPLEMENTED) return rc;
    // otherwise the goal / reward subset of the transition sampler: same draw, same outputs (block-cooperative row gathers)
    return bp_her_sample(nullptr, nullptr, d_ep_g, d_ep_ag, nullptr, B, T, 1, 0, dimg, n, future_p, 0.0f, seed, index_offset,
                         d_ep_idx, d_t, d_future_t, nullptr, nullptr, nullptr, d_g, nullptr, d_ag2, d_r, nullptr, nullptr, stream);
}

}  // extern "C"
