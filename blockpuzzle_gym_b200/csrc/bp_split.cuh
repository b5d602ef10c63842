// bp_split.cuh -- the step-synchronous form of the step path (included by bp_kernels.cu).
//
// step_kernel_async keeps a slab of envs on chip for all K fused steps and schedules quiet env-steps and
// full-physics passes inside one warp.  Round 2's profiles showed what bounds it: instruction delivery (its hot code
// -- scheduler + quiet path + pass + finalize -- just exceeds the SM's instruction cache; the GPC-level instruction
// cache runs at 85-93 % of its request rate), 12-13 warps per SM (the slab's shared memory), passes filled to ~22 of 32
// lanes, and a chain of ~50 serial passes per slab.  None of that is inherent in the work.  Here every env step is
// three small kernels over the handle's global state (field-major, L2 / HBM resident):
//   split_quiet_kernel   thread per env: clip, the quiet path (gripper-only integration + swept-volume test) and, for
//                        the envs it settles, everything RobotEnv.step does after sim.step(); the others are appended to
//                        the step's full-physics list (warp-aggregated atomics);
//   split_full_kernel    thread per list entry -- every warp has 32 busy lanes -- the complete BlockPhys step in
//                        registers (sim_step_reg) + the same post-step code;
//   split_reset_kernel   thread per env that finished an episode (auto-reset): RobotEnv.reset.
// Each kernel's code fits the instruction cache on its own, occupancy is set by registers alone, and there is no
// scheduler.  The price is the env state's round trip through L2 / HBM every step (about +160 B per env-step of
// traffic on top of the 252.5 algorithmic bytes).  Env-steps are independent and each env's steps stay ordered (stream
// order of the launches), so results are bit-identical to step_kernel_async, step_kernel_simple and the oracle
// (test_all_step_kernels_agree; the whole GPU suite passes with BP_STEP_KERNEL=split).
//
// MEASURED (round 2, 1 Mi BlocksTouch-v0 envs, K = 64, same box): 4.21e9 env-steps/s against 4.50e9 for
// step_kernel_async, so this path is NOT the default (BP_STEP_KERNEL=split or bp_set_option("step_kernel", 4) selects
// it).  ncu per step: quiet kernel 100-125 us (57 M warp-instructions at 29.8 of 32 lanes, issue 44-54 %, 52 % of HBM
// bandwidth), full-physics kernel 80-230 us (27-176 M warp-instructions at only 15-20 of 32 lanes, issue 71-74 %),
// reset kernel 4 us.  What it settles: fully packed 32-lane warps do NOT make the full physics cheaper -- the contact
// responses diverge inside the warp (vertical / lateral / yield branches, 1-3 candidate slots per lane), and a wider
// warp executes the union of more lanes' paths: 65 instructions per lane and substep here against 42 in the async
// kernel's 22-lane passes.  Total instructions per env-step come out the same (155); the slab kernel then wins on
// memory traffic and launch gaps.
#pragma once

namespace bp {

struct SplitBufs {
    int32_t* list_full;    // [B] launch-local env indices that need the full physics this step
    int32_t* list_reset;   // [B] envs that finished an episode this step
    int32_t* counts;       // [2 sets][2]: full, reset (the sets alternate between steps; split_reset_kernel zeroes the next one)
    double* part;          // [slots][8] per-block statistics partial sums (no same-address atomics in the step kernels)
    int q_blocks;          // blocks of split_quiet_kernel = first slot of split_full_kernel
};

constexpr int kSplitQuietThreads = 256, kSplitFullThreads = 128;

// the action of env-step k of env li (state already in `e`): the caller's tensor or Philox stream 2
template <int NB>
__device__ __forceinline__ float4 split_action(const StepArgs& p, const Env<NB>& e, int64_t li, int k) {
    float4 a4;
    if (p.actions) {
        a4 = __ldg(reinterpret_cast<const float4*>(p.actions) + ((int64_t)(p.act_k0 + k) * p.B + li));
    } else {
        U4 w = philox4x32((uint32_t)e.t, e.episode - 1u, 2u, 0u, e.key0, e.key1);
        a4 = make_float4(2.0f * u01(w.x) - 1.0f, 2.0f * u01(w.y) - 1.0f, 2.0f * u01(w.z) - 1.0f, 2.0f * u01(w.w) - 1.0f);
    }
    if (p.actions_out) reinterpret_cast<float4*>(p.actions_out)[step_row(p, k, li)] = a4;
    return a4;
}

// everything RobotEnv.step does after sim.step() (robot_env.py:61-69 under TimeLimit): touch matrix, reward, latch,
// outputs.  Returns bit 0: reward == -1, bit 1: done, bit 2: success at done.
template <int ID>
__device__ __forceinline__ uint32_t split_post_step(const StepArgs& p, Env<Cfg<ID>::NB>& e, uint32_t contacts, int64_t li, int k) {
    using C = Cfg<ID>;
    const bool fail = env_post_step<ID>(contacts, e.touch_now, e.touch_ever, e.succ, e.t);
    const bool done = e.t >= kT;
    const int64_t row = step_row(p, k, li), orow = obs_row(p, k, li);
    if (p.obs) store_row<C::DIMO>(p.obs + orow * C::DIMO, [&](auto&& put) { env_write_obs<ID>(e, put); });
    if (p.ag) store_row<C::DIMG>(p.ag + orow * C::DIMG, [&](auto&& put) { env_write_ag<ID>(e.touch_now, e.touch_ever, put); });
    if (p.goal_out) store_row<C::DIMG>(p.goal_out + row * C::DIMG, [&](auto&& put) { env_write_goal<ID>(put); });
    if (p.reward) store_reward(p.reward + row, fail);
    if (p.success) p.success[row] = (float)e.succ;
    if (p.done) p.done[row] = done ? 1 : 0;
    return (fail ? 1u : 0u) | (done ? 2u : 0u) | ((done && e.succ) ? 4u : 0u);
}

// append the calling lanes' values to a global list: one atomicAdd per warp
__device__ __forceinline__ void split_append(bool want, int32_t value, int32_t* list, int32_t* count) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const int lane = (int)(threadIdx.x & 31);
    int base = 0;
    if (lane == __ffs((int)m) - 1) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs((int)m) - 1);
    if (want) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// block-wide sum of six per-thread statistics into this block's slot (plain read-modify-write: the slot is this
// block's alone, and successive launches are ordered by the stream)
template <int THREADS>
__device__ __forceinline__ void split_stats(float (&v)[6], double* slot) {
    __shared__ float s_red[THREADS / 32][6];
#pragma unroll
    for (int j = 0; j < 6; ++j) v[j] = warp_sum(v[j]);
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 6; ++j) s_red[warp][j] = v[j];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) t += s_red[w][threadIdx.x];
        if (t != 0.f) slot[threadIdx.x] += (double)t;
    }
}

// stats order inside a slot: episodes, successes, steps, invalid, reward_sum, worker_steps
template <int ID>
__global__ void __launch_bounds__(kSplitQuietThreads) split_quiet_kernel(uint32_t* __restrict__ st, const __grid_constant__ StepArgs p,
                                                                        const int k, const SplitBufs sb, const int set) {
    using C = Cfg<ID>;
    constexpr int NB = C::NB;
    const int64_t li = (int64_t)blockIdx.x * kSplitQuietThreads + threadIdx.x;
    const bool live = li < p.B;
    float sv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bool want_full = false, want_reset = false;
    if (live) {
        const int64_t gi = p.env0 + li;
        Env<NB> e;
        load_env<NB>(st, p.stateB, gi, e);
        const float4 a4 = split_action<NB>(p, e, li, k);
        float a[4] = {a4.x, a4.y, a4.z, a4.w};
        int inv = 0;
        clip_action(a, inv);
        bool quiet = false;
        Grip g2;
        if ((e.priv >> 31) && !(p.tune & kTuneForceFull)) {
#pragma unroll
            for (int d = 0; d < 3; ++d) { g2.g[d] = e.g[d]; g2.gv[d] = e.gv[d]; }
            g2.q[0] = e.q[0]; g2.q[1] = e.q[1]; g2.qv[0] = e.qv[0]; g2.qv[1] = e.qv[1];
            float m[3], ctrl[2];
            action_targets<C::BG>(g2, a, m, ctrl);
            float lo[3], hi[3], qmax;
            quiet_gripper_step<C::BG>(g2, m, ctrl, lo, hi, qmax);
            quiet = true;
#pragma unroll
            for (int b = 0; b < NB; ++b)
                if (b < e.nb) quiet = quiet && cube_out_of_reach(e.px[b], e.py[b], e.pz[b], e.c[b], e.s[b], lo, hi, qmax);
        }
        if (quiet) {
#pragma unroll
            for (int d = 0; d < 3; ++d) { e.g[d] = g2.g[d]; e.gv[d] = g2.gv[d]; }
            e.q[0] = g2.q[0]; e.q[1] = g2.q[1]; e.qv[0] = g2.qv[0]; e.qv[1] = g2.qv[1];
            uint32_t contacts = e.priv & 0x7fffu;
            if (over_table(e.g[0], e.g[1]) && e.g[2] - kGZMin < kMargin) contacts |= pair_bit(0, 1);
            const uint32_t code = split_post_step<ID>(p, e, contacts, li, k);
            // only the gripper and the two bookkeeping words changed
            uint32_t* q = st + gi;
#pragma unroll
            for (int d = 0; d < 3; ++d) { q[(int64_t)d * p.stateB] = __float_as_uint(e.g[d]); q[(int64_t)(3 + d) * p.stateB] = __float_as_uint(e.gv[d]); }
            q[(int64_t)6 * p.stateB] = __float_as_uint(e.q[0]); q[(int64_t)7 * p.stateB] = __float_as_uint(e.q[1]);
            q[(int64_t)8 * p.stateB] = __float_as_uint(e.qv[0]); q[(int64_t)9 * p.stateB] = __float_as_uint(e.qv[1]);
            q[(int64_t)(10 + 9 * NB) * p.stateB] = e.touch_now | (e.touch_ever << 16);
            q[(int64_t)(11 + 9 * NB) * p.stateB] = (uint32_t)(e.t < 255 ? e.t : 255) | ((uint32_t)e.succ << 8) | ((uint32_t)e.nb << 9);
            sv[2] += 1.f; sv[3] += (float)inv; sv[4] -= (float)(code & 1u); sv[0] += (float)((code >> 1) & 1u); sv[1] += (float)((code >> 2) & 1u);
            want_reset = p.auto_reset && (code & 2u);
        } else {
            want_full = true;
        }
    }
    split_append(want_full, (int32_t)li, sb.list_full, sb.counts + 2 * set);
    split_append(want_reset, (int32_t)li, sb.list_reset, sb.counts + 2 * set + 1);
    split_stats<kSplitQuietThreads>(sv, sb.part + (int64_t)blockIdx.x * 8);
}

template <int ID>
__global__ void __launch_bounds__(kSplitFullThreads) split_full_kernel(uint32_t* __restrict__ st, const __grid_constant__ StepArgs p,
                                                                      const int k, const SplitBufs sb, const int set) {
    using C = Cfg<ID>;
    constexpr int NB = C::NB;
    const int n = sb.counts[2 * set];
    float sv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int stride = (int)gridDim.x * kSplitFullThreads;
    // whole warps iterate together (split_append uses full-mask warp intrinsics)
    for (int base = (int)blockIdx.x * kSplitFullThreads + (int)(threadIdx.x & ~31u); base < n; base += stride) {
        const int idx = base + (int)(threadIdx.x & 31);
        bool want_reset = false;
        int32_t li32 = 0;
        if (idx < n) {
            li32 = sb.list_full[idx];
            const int64_t li = li32, gi = p.env0 + li;
            Env<NB> e;
            load_env<NB>(st, p.stateB, gi, e);
            const float4 a4 = split_action<NB>(p, e, li, k);
            float a[4] = {a4.x, a4.y, a4.z, a4.w};
            int inv = 0;
            clip_action(a, inv);
            Grip g;
#pragma unroll
            for (int d = 0; d < 3; ++d) { g.g[d] = e.g[d]; g.gv[d] = e.gv[d]; }
            g.q[0] = e.q[0]; g.q[1] = e.q[1]; g.qv[0] = e.qv[0]; g.qv[1] = e.qv[1];
            CubeRegs<NB> q;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                q.x[b] = e.px[b]; q.y[b] = e.py[b]; q.z[b] = e.pz[b]; q.c[b] = e.c[b]; q.s[b] = e.s[b];
                q.vx[b] = e.vx[b]; q.vy[b] = e.vy[b]; q.vz[b] = e.vz[b]; q.w[b] = e.w[b];
            }
            uint32_t contacts = 0;
            const bool is_static = sim_step_reg<NB, C::BG, C::VAR>(g, a, q, e.nb, contacts);
#pragma unroll
            for (int d = 0; d < 3; ++d) { e.g[d] = g.g[d]; e.gv[d] = g.gv[d]; }
            e.q[0] = g.q[0]; e.q[1] = g.q[1]; e.qv[0] = g.qv[0]; e.qv[1] = g.qv[1];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                e.px[b] = q.x[b]; e.py[b] = q.y[b]; e.pz[b] = q.z[b]; e.c[b] = q.c[b]; e.s[b] = q.s[b];
                e.vx[b] = q.vx[b]; e.vy[b] = q.vy[b]; e.vz[b] = q.vz[b]; e.w[b] = q.w[b];
            }
            e.priv = (is_static ? 0x80000000u : 0u) | (contacts & ~gripper_pair_mask());
            const uint32_t code = split_post_step<ID>(p, e, contacts, li, k);
            // gripper, cubes and the three bookkeeping words (episode / draw counters / key are untouched)
            uint32_t* w = st + gi;
#pragma unroll
            for (int d = 0; d < 3; ++d) { w[(int64_t)d * p.stateB] = __float_as_uint(e.g[d]); w[(int64_t)(3 + d) * p.stateB] = __float_as_uint(e.gv[d]); }
            w[(int64_t)6 * p.stateB] = __float_as_uint(e.q[0]); w[(int64_t)7 * p.stateB] = __float_as_uint(e.q[1]);
            w[(int64_t)8 * p.stateB] = __float_as_uint(e.qv[0]); w[(int64_t)9 * p.stateB] = __float_as_uint(e.qv[1]);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                uint32_t* c = w + (int64_t)(10 + 9 * b) * p.stateB;
                c[0] = __float_as_uint(e.px[b]); c[p.stateB] = __float_as_uint(e.py[b]); c[2 * p.stateB] = __float_as_uint(e.pz[b]);
                c[3 * p.stateB] = __float_as_uint(e.c[b]); c[4 * p.stateB] = __float_as_uint(e.s[b]);
                c[5 * p.stateB] = __float_as_uint(e.vx[b]); c[6 * p.stateB] = __float_as_uint(e.vy[b]); c[7 * p.stateB] = __float_as_uint(e.vz[b]);
                c[8 * p.stateB] = __float_as_uint(e.w[b]);
            }
            w[(int64_t)(10 + 9 * NB) * p.stateB] = e.touch_now | (e.touch_ever << 16);
            w[(int64_t)(11 + 9 * NB) * p.stateB] = (uint32_t)(e.t < 255 ? e.t : 255) | ((uint32_t)e.succ << 8) | ((uint32_t)e.nb << 9);
            w[(int64_t)(12 + 9 * NB) * p.stateB] = e.priv;
            sv[2] += 1.f; sv[3] += (float)inv; sv[5] += 1.f;
            sv[4] -= (float)(code & 1u); sv[0] += (float)((code >> 1) & 1u); sv[1] += (float)((code >> 2) & 1u);
            want_reset = p.auto_reset && (code & 2u);
        }
        split_append(want_reset, li32, sb.list_reset, sb.counts + 2 * set + 1);
    }
    split_stats<kSplitFullThreads>(sv, sb.part + (int64_t)(sb.q_blocks + (int)blockIdx.x) * 8);
}

template <int ID>
__global__ void __launch_bounds__(128) split_reset_kernel(uint32_t* __restrict__ st, const __grid_constant__ StepArgs p, const SplitBufs sb, const int set) {
    using C = Cfg<ID>;
    constexpr int NB = C::NB;
    const int n = sb.counts[2 * set + 1];
    if (blockIdx.x == 0 && threadIdx.x == 0) { sb.counts[2 * (set ^ 1)] = 0; sb.counts[2 * (set ^ 1) + 1] = 0; }   // the next step's lists
    for (int idx = (int)blockIdx.x * 128 + (int)threadIdx.x; idx < n; idx += (int)gridDim.x * 128) {
        const int64_t li = sb.list_reset[idx], gi = p.env0 + li;
        Env<NB> e;
        load_env<NB>(st, p.stateB, gi, e);
        env_reset<ID>(e, p.rg);
        store_env<NB>(st, p.stateB, gi, e);
        if (p.reset_obs) write_row_obs<ID>(e, p.reset_obs + li * C::DIMO);
        if (p.reset_ag) write_row_ag<ID>(e, p.reset_ag + li * C::DIMG);
    }
}

// slots -> the handle's statistics vector (once per bp_step call), and the slots are cleared
__global__ void __launch_bounds__(256) split_stats_reduce_kernel(double* part, int slots, double* stats) {
    __shared__ double s_red[8][6];
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int s = (int)threadIdx.x; s < slots; s += 256) {
#pragma unroll
        for (int j = 0; j < 6; ++j) { v[j] += part[(int64_t)s * 8 + j]; part[(int64_t)s * 8 + j] = 0.0; }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int j = 0; j < 6; ++j) s_red[threadIdx.x >> 5][j] = v[j];
    }
    __syncthreads();
    if (threadIdx.x < 6 && stats) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
        // slot order -> BP_STAT_* order: episodes, successes, steps, invalid, reward_sum, worker_steps
        if (t != 0.0) atomicAdd(stats + threadIdx.x, t);
    }
}

}  // namespace bp
