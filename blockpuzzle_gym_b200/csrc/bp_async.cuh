// bp_async.cuh -- the warp-autonomous step kernel (included by bp_kernels.cu).
//
// Why: with one CTA-wide barrier per phase (step_kernel_tiled) every env of a tile waits, at every
// step, for the serial full-physics pass of the tile's few active envs; measured throughput was
// exactly proportional to the number of resident CTAs (latency-bound, 45 % issue utilisation).
// Here each WARP is an independent scheduler over its own 32*E envs (one single-warp CTA, no block
// barriers at all):
//   * every lane owns E envs; in each iteration it picks its least-advanced ready env-step and tries
//     the quiet path (gripper-only integration + swept-volume test, see cube_out_of_reach);
//   * env-steps that need the full physics are appended to the warp's pending list and the lane moves
//     on to another of its envs -- quiet envs never wait for active ones;
//   * whenever 32 env-steps are pending (or nothing else is runnable) the whole warp runs ONE
//     full-physics pass with all lanes busy (lane i takes pending item i, whichever lane owns it).
// Env-steps of different envs are independent and each env's steps stay in order, so the result is
// schedule-independent and bit-identical to step_kernel_simple and to the oracle.  All env state lives
// in the warp's private shared-memory slab (SoA, stride 32*E); outputs are written row by row straight
// from registers (rows are multiples of 32-byte sectors for the 16-byte-multiple row sizes).
#pragma once

namespace bp {

// steps per launch of the async kernel (the host splits longer K): reward and success of every
// (step, env) are kept as bit masks in shared memory and written as whole coalesced rows at the end.
// Writing them per env-step (4 bytes each, at scattered times) made every 32-byte sector a partial
// write: ncu showed +64 B/env-step of DRAM fills and +24 B/env-step of write-backs.
constexpr int kMaxFused = 64;

template <int ID, int E>
struct Async {
    using C = Cfg<ID>;
    static constexpr int NB = C::NB;
    static constexpr int CS = 32 * E;                       // envs per warp = column stride of the slab
    static constexpr int W_GRP = 10 * CS, W_MSC = 3 * CS;
    // ids with three or four cubes keep round 1's pass (cubes stepped in place in the slab through shared-memory
    // loops, 4 words of per-lane scratch per cube): unrolled over 4 cubes, 8 finger slots and 6 pairs the register
    // form is 74-91 KB of SASS and measured 13-23 % slower there (ToppleTower 1.32 -> 1.01e9, Variation 1.41 -> 1.22e9)
    static constexpr bool REG_PASS = NB <= 2;
    // yaw cache (ids with the register pass): the yaw observation of every cube, bp_atan2(s, c), kept as a tenth column
    // per cube behind the nine state columns and refreshed only when the cube may have turned (slab load, reset, full-
    // physics pass).  _get_obs evaluated it for every cube at every env-step: 5.6 % of the kernel's instructions incl. the
    // division slow path, and 4 KB of the hot code.  One more KB of slab: 12 x (18 KB + 1 KB reserved) is exactly the SM's 228 KB.
    static constexpr bool YAW = REG_PASS;
    static constexpr int W_CUB = (9 * NB + (YAW ? NB : 0)) * CS;
    static constexpr int W_COL = REG_PASS ? 0 : Col<NB, CS, 32>::kScratch * 32;
    static constexpr int W_PEND = CS;                        // two uint16 rings of CS entries: pending full-physics env-steps, pending resets
    static constexpr int W_BITS = kMaxFused * E;             // reward bits of every (step, env) of the launch
    static constexpr size_t SMEM = sizeof(uint32_t) * (size_t)(W_CUB + W_GRP + W_MSC + W_COL + W_PEND + W_BITS);
};

#ifndef BP_ASYNC_E
#define BP_ASYNC_E 4
#endif
constexpr int kAsyncE = BP_ASYNC_E;  // envs per lane (build-time; -DBP_ASYNC_E for sweeps)
constexpr int kTuneForceFull = 1 << 23;   // StepArgs::tune: skip the quiet path (measurement of the all-full-physics floor)

// The slab's three bookkeeping words per env: [0] touch masks (now | ever << 16), [1] kernel-private flags
// (bit 31: cubes static, bits 0-14: their contact pairs) | env flags << 16 (t | succ << 8 | nb << 9),
// [2] status: steps completed in this launch | pending flag.
constexpr uint32_t kPendingBit = 1u << 16;
constexpr uint32_t kPrivMask = 0x80007fffu;
__device__ __forceinline__ uint32_t pf_flags(uint32_t pf) { return (pf >> 16) & 0xfffu; }
__device__ __forceinline__ uint32_t pf_make(uint32_t priv, uint32_t flags) { return (priv & kPrivMask) | (flags << 16); }

// refresh the yaw cache of one slab column (see Async::YAW); one out-of-line copy serves all call sites
template <int NB, int CS>
__device__ __noinline__ void yaw_refresh(float* cub) {
#pragma unroll 1
    for (int b = 0; b < NB; ++b) cub[(9 * NB + b) * CS] = bp_atan2(cub[(9 * b + 4) * CS], cub[(9 * b + 3) * CS]);
}

// RobotEnv.reset for one env of a slab (rare and large: kept out of line)
template <int ID, int CS, bool LEAN>
__device__ __noinline__ void reset_env_slab(uint32_t* __restrict__ st, const StepArgs& p, int64_t li, float* cub, float* grp, uint32_t* msc) {
    using C = Cfg<ID>;
    constexpr int NB = C::NB;
    constexpr int NF = num_fields<NB>();
    const int64_t gi = p.env0 + li;
    Env<NB> e;
    const uint32_t touch = msc[0], flags = pf_flags(msc[CS]);
    e.nb = (int)((flags >> 9) & 7u);
    e.touch_now = touch & 0xffffu; e.touch_ever = touch >> 16;
    e.episode = st[(int64_t)(NF - 5) * p.stateB + gi];
    e.key0 = st[(int64_t)(NF - 2) * p.stateB + gi]; e.key1 = st[(int64_t)(NF - 1) * p.stateB + gi];
    env_reset<ID>(e, p.rg);
    st[(int64_t)(NF - 5) * p.stateB + gi] = e.episode;
    st[(int64_t)(NF - 4) * p.stateB + gi] = e.draws0;
    st[(int64_t)(NF - 3) * p.stateB + gi] = e.draws1;
#pragma unroll
    for (int d = 0; d < 3; ++d) { grp[d * CS] = e.g[d]; grp[(3 + d) * CS] = e.gv[d]; }
    grp[6 * CS] = e.q[0]; grp[7 * CS] = e.q[1]; grp[8 * CS] = e.qv[0]; grp[9 * CS] = e.qv[1];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float* bb = cub + (9 * b) * CS;
        bb[0] = e.px[b]; bb[CS] = e.py[b]; bb[2 * CS] = e.pz[b]; bb[3 * CS] = e.c[b]; bb[4 * CS] = e.s[b];
        bb[5 * CS] = e.vx[b]; bb[6 * CS] = e.vy[b]; bb[7 * CS] = e.vz[b]; bb[8 * CS] = e.w[b];
    }
    if (Async<ID, CS / 32>::YAW) yaw_refresh<NB, CS>(cub);
    msc[0] = e.touch_now | (e.touch_ever << 16);
    msc[CS] = pf_make(e.priv, (uint32_t)e.t | ((uint32_t)e.succ << 8) | ((uint32_t)e.nb << 9));
    if (!LEAN) {
        if (p.reset_obs) write_row_obs<ID>(e, p.reset_obs + li * C::DIMO);
        if (p.reset_ag) write_row_ag<ID>(e, p.reset_ag + li * C::DIMG);
    }
}

// write a row of W floats generated by `gen(put)` to global memory (128-bit stores when rows are 16-byte multiples)
// A32: the caller guarantees a 32-byte aligned tensor (the lean kernel: launch_step checks it), so rows of whole sectors
// need no run-time alignment test and no 128-bit fallback in the hot code.
template <int W, bool A32 = false, class Gen>
__device__ __forceinline__ void store_row(float* __restrict__ dst, Gen&& gen) {
    if constexpr (W % 4 == 0) {
        float row[W];
        gen([&](int c, float v) { row[c] = v; });
        float4* d4 = reinterpret_cast<float4*>(dst);
        // streaming stores (evict-first): the 15.6 GB of output rows of a launch must not push the action rows the
        // neighbouring envs still need out of L2
#ifndef BP_NO_STG256
        if constexpr (W % 8 == 0) {
            // rows of whole 32-byte sectors (40 / 16 floats): one 256-bit store per sector (sm_100: STG.E.256) when the
            // caller's tensor is 32-byte aligned -- half the store instructions, and no sector is written in two halves
            if (A32 || (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
#pragma unroll
                for (int j = 0; j < W / 8; ++j)
                    asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * j), "f"(row[8 * j]), "f"(row[8 * j + 1]),
                                 "f"(row[8 * j + 2]), "f"(row[8 * j + 3]), "f"(row[8 * j + 4]), "f"(row[8 * j + 5]), "f"(row[8 * j + 6]), "f"(row[8 * j + 7]) : "memory");
                return;
            }
        }
#endif
#pragma unroll
        for (int j = 0; j < W / 4; ++j) __stcs(d4 + j, make_float4(row[4 * j], row[4 * j + 1], row[4 * j + 2], row[4 * j + 3]));
    } else {
#ifdef BP_SCALAR_ROWS
        gen([&](int c, float v) { dst[c] = v; });
#else
        // Rows that are not 16-byte multiples (25 / 9, 55 / 25, 70, 87 floats) start at any 4-byte (70: 8-byte) alignment, a different one in
        // every lane.  Written float by float they were W separate 4-byte requests per lane to the L2 (ncu, GripperTouch-v0:
        // 28 written sectors per env-step for 5 sectors' worth of bytes, mio_throttle + short scoreboard the top stalls; Choose:
        // the L2 at 57 % of its peak).  Here every lane writes `h` floats up to its next 16-byte boundary, then 128-bit stores
        // from the row shifted left by `h` in registers (two select stages), then the tail: W / 4 + 10 store instructions
        // (up to nine of them predicated scalars: three head, six tail slots) and 2 W selects instead of W stores.
        float row[W + 3];
        gen([&](int c, float v) { row[c] = v; });
        row[W] = 0.0f; row[W + 1] = 0.0f; row[W + 2] = 0.0f;
        const unsigned h = (4u - ((unsigned)(reinterpret_cast<uintptr_t>(dst) >> 2) & 3u)) & 3u;
        if (h > 0) __stcs(dst, row[0]);
        if (h > 1) __stcs(dst + 1, row[1]);
        if (h > 2) __stcs(dst + 2, row[2]);
        float t[W + 2], u[W + 4];
#pragma unroll
        for (int i = 0; i < W + 2; ++i) t[i] = (h & 1u) ? row[i + 1] : row[i];
#pragma unroll
        for (int i = 0; i < W; ++i) u[i] = (h & 2u) ? t[i + 2] : t[i];     // u[i] = row[i + h]
        u[W] = 0.0f; u[W + 1] = 0.0f; u[W + 2] = 0.0f; u[W + 3] = 0.0f;
        constexpr int NV = (W - 3) / 4;                                    // 128-bit stores every alignment has
        float* d = dst + h;
        float4* d4 = reinterpret_cast<float4*>(d);
#pragma unroll
        for (int j = 0; j < NV; ++j) __stcs(d4 + j, make_float4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]));
        const int rem = W - 4 * NV - (int)h;                               // 0 .. 6 floats left
        constexpr int T0 = 4 * NV;
        if (W - T0 >= 4 && rem >= 4) __stcs(d4 + NV, make_float4(u[T0], u[T0 + 1], u[T0 + 2], u[T0 + 3]));
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            if (rem < 4 && l < rem) __stcs(d + T0 + l, u[T0 + l]);
            if (W - T0 > 4 + l) { if (rem >= 4 && l < rem - 4) __stcs(d + T0 + 4 + l, u[T0 + 4 + l]); }
        }
#endif
    }
}

// Everything RobotEnv.step does after sim.step() for one env-step whose post-physics state is in the slab
// columns (cub, grp, msc; stride CS): _step_callback, reward, latch, TimeLimit, outputs.  The output row
// pointers arrive BY VALUE (any may be null): reading them through a `const StepArgs&` inside this
// out-of-line function compiled to a chain of generic loads from the parameter window, each one a
// long-scoreboard stall (ncu: half of this function's stall samples).  The auto-reset is NOT done here:
// the caller queues finished envs and resets them a warp at a time (see step_kernel_async).
// Returns bit 0: reward == -1, bit 1: episode done, bit 2: success at done.
template <int ID, int CS, bool LEAN>
__device__ __noinline__ uint32_t finalize_step(uint32_t contacts, float* cub, float* grp, uint32_t* msc, uint32_t* bits_word, uint32_t bit,
                                               uint8_t* done_out, const float4* next_action, float* goal_row, float* obs_row_p, float* ag_row_p) {
    using C = Cfg<ID>;
    constexpr int NB = C::NB;
    const uint32_t touch = msc[0], pf = msc[CS], flags = pf_flags(pf);
    uint32_t touch_now = touch & 0xffffu, touch_ever = touch >> 16;
    int t = (int)(flags & 0xffu), succ = (int)((flags >> 8) & 1u);
    const int nb = (int)((flags >> 9) & 7u);
    const bool fail = env_post_step<ID>(contacts, touch_now, touch_ever, succ, t);
    const bool done = t >= kT;
    msc[0] = touch_now | (touch_ever << 16);
    msc[CS] = pf_make(pf, (uint32_t)(t < 255 ? t : 255) | ((uint32_t)succ << 8) | ((uint32_t)nb << 9));   // t saturates in its 8 bits (auto_reset = 0 callers may step past T)
    if (fail) atomicOr(bits_word, bit);          // reward bit of (k, env): flushed as whole rows at the end (the latched
                                                 // success rows are rebuilt from these bits there, see slab_flush)
    if (!LEAN && done_out) *done_out = done ? 1 : 0;
    if (next_action) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(next_action));  // the next action of this env: pull it into L2 now, and keep it: the env sharing its sector needs the other half later
    if (!LEAN && goal_row) store_row<C::DIMG>(goal_row, [&](auto&& put) { env_write_goal<ID>(put); });
    if (obs_row_p) {
        Env<NB> e;
#pragma unroll
        for (int d = 0; d < 3; ++d) { e.g[d] = grp[d * CS]; e.gv[d] = grp[(3 + d) * CS]; }
        e.q[0] = grp[6 * CS]; e.q[1] = grp[7 * CS]; e.qv[0] = grp[8 * CS]; e.qv[1] = grp[9 * CS];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float* bb = cub + (9 * b) * CS;
            e.px[b] = bb[0]; e.py[b] = bb[CS]; e.pz[b] = bb[2 * CS]; e.c[b] = bb[3 * CS]; e.s[b] = bb[4 * CS];
            e.vx[b] = bb[5 * CS]; e.vy[b] = bb[6 * CS]; e.vz[b] = bb[7 * CS]; e.w[b] = bb[8 * CS];
        }
        e.nb = nb;
        constexpr bool YAW = Async<ID, CS / 32>::YAW;
        float yw[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) yw[b] = YAW ? cub[(9 * NB + b) * CS] : 0.0f;
        store_row<C::DIMO, LEAN>(obs_row_p, [&](auto&& put) { env_write_obs<ID>(e, put, YAW ? yw : nullptr); });
    }
    if (ag_row_p) store_row<C::DIMG, LEAN>(ag_row_p, [&](auto&& put) { env_write_ag<ID>(touch_now, touch_ever, put); });
    return (fail ? 1u : 0u) | (done ? 2u : 0u) | ((done && succ) ? 4u : 0u);
}

// finalize_step on env-step (k, li) of slab column w with the output rows of `p` (read here, in the
// kernel, from the constant bank)
template <int ID, int E, bool LEAN>
__device__ __forceinline__ uint32_t finalize_at(const StepArgs& p, int64_t li, int k, int w, uint32_t contacts,
                                                float* s_cub, float* s_grp, uint32_t* s_msc, uint32_t* s_bits) {
    using C = Cfg<ID>;
    constexpr int CS = 32 * E;
    if (LEAN) {   // time-major rows, no done / goal rows: one row index serves every tensor.  launch_step hands this
                  // instantiation pointers already advanced to the launch's first step (k0 = act_k0 = 0) and only launches whose
                  // K * B fits 31 bits, so the row index is 32-bit arithmetic and every address one widening multiply-add
        const uint32_t row = (uint32_t)k * (uint32_t)p.B + (uint32_t)li;
        return finalize_step<ID, CS, true>(contacts, s_cub + w, s_grp + w, s_msc + w, s_bits + (k * E + (w >> 5)), 1u << (w & 31), nullptr,
                                     (k + 1 < p.K) ? reinterpret_cast<const float4*>(p.actions) + (row + (uint32_t)p.B) : nullptr,
                                     nullptr, p.obs ? p.obs + (uint64_t)row * C::DIMO : nullptr, p.ag ? p.ag + (uint64_t)row * C::DIMG : nullptr);
    }
    const int64_t row = step_row(p, k, li), orow = obs_row(p, k, li);
    return finalize_step<ID, CS, false>(contacts, s_cub + w, s_grp + w, s_msc + w, s_bits + (k * E + (w >> 5)), 1u << (w & 31),
                                 p.done ? p.done + row : nullptr,
                                 (p.actions && k + 1 < p.K) ? reinterpret_cast<const float4*>(p.actions) + ((int64_t)(p.act_k0 + k + 1) * p.B + li) : nullptr,
                                 p.goal_out ? p.goal_out + row * C::DIMG : nullptr,
                                 p.obs ? p.obs + orow * C::DIMO : nullptr,
                                 p.ag ? p.ag + orow * C::DIMG : nullptr);
}

// the action of env-step (k, li): from the caller's tensor or from Philox stream 2 (replay harness)
template <int ID, bool LEAN>
__device__ __forceinline__ float4 fetch_action(const uint32_t* __restrict__ st, const StepArgs& p, int64_t li, int k, int t) {
    constexpr int NF = num_fields<Cfg<ID>::NB>();
    if (LEAN) return __ldg(reinterpret_cast<const float4*>(p.actions) + ((uint32_t)k * (uint32_t)p.B + (uint32_t)li));   // k0 = 0, 32-bit rows: see finalize_at
    float4 a4;
    if (p.actions) {  // the action tensor is always time-major [Ktot][B][4]
        a4 = __ldg(reinterpret_cast<const float4*>(p.actions) + ((int64_t)(p.act_k0 + k) * p.B + li));
    } else {
        const int64_t gi = p.env0 + li;
        const uint32_t ep = st[(int64_t)(NF - 5) * p.stateB + gi];
        const uint32_t k0 = st[(int64_t)(NF - 2) * p.stateB + gi], k1 = st[(int64_t)(NF - 1) * p.stateB + gi];
        U4 w = philox4x32((uint32_t)t, ep - 1u, 2u, 0u, k0, k1);
        a4 = make_float4(2.0f * u01(w.x) - 1.0f, 2.0f * u01(w.y) - 1.0f, 2.0f * u01(w.z) - 1.0f, 2.0f * u01(w.w) - 1.0f);
    }
    if (p.actions_out) reinterpret_cast<float4*>(p.actions_out)[step_row(p, k, li)] = a4;
    return a4;
}


// per-warp statistics accumulators (reduced once at the end of the kernel)
struct WarpStats { float n_ep = 0.f, n_su = 0.f, n_st = 0.f, n_inv = 0.f, r_sum = 0.f, n_wrk = 0.f; };

// The quiet path of env-step (ksel, column w): clip the action, integrate the gripper alone for the 20
// substeps while tracking the swept finger volume, and -- if the cubes sit on a fixed point that volume
// cannot reach -- commit the step (finalize_step).  Returns 0 when the step needs the full physics (nothing
// has been modified in that case), else 1 | finalize code << 1.
template <int ID, int E, bool LEAN>
__device__ __forceinline__ uint32_t try_quiet_step(uint32_t* __restrict__ st, const StepArgs& p, int64_t li, int w, int ksel,
                                               float* s_cub, float* s_grp, uint32_t* s_msc, uint32_t* s_bits, WarpStats& ws) {
    using C = Cfg<ID>;
    constexpr int NB = C::NB, CS = 32 * E;
    const uint32_t priv = s_msc[CS + w], flags = pf_flags(priv);
    const float4 a4 = fetch_action<ID, LEAN>(st, p, li, ksel, (int)(flags & 0xffu));
    float a[4] = {a4.x, a4.y, a4.z, a4.w};
    int inv = 0;
    clip_action(a, inv);
    if (!(priv >> 31) || (p.tune & kTuneForceFull)) return 0u;
    Grip g2;
#pragma unroll
    for (int d = 0; d < 3; ++d) { g2.g[d] = s_grp[d * CS + w]; g2.gv[d] = s_grp[(3 + d) * CS + w]; }
    g2.q[0] = s_grp[6 * CS + w]; g2.q[1] = s_grp[7 * CS + w]; g2.qv[0] = s_grp[8 * CS + w]; g2.qv[1] = s_grp[9 * CS + w];
    float m[3], ctrl[2];
    action_targets<C::BG>(g2, a, m, ctrl);
    float lo[3], hi[3], qmax;
    quiet_gripper_step<C::BG>(g2, m, ctrl, lo, hi, qmax);
    const int nb = (int)((flags >> 9) & 7u);
    bool quiet = true;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (!C::VAR || b < nb) {
            const float* bb = s_cub + (9 * b) * CS + w;
            quiet = quiet && cube_out_of_reach(bb[0], bb[CS], bb[2 * CS], bb[3 * CS], bb[4 * CS], lo, hi, qmax);
        }
    }
    if (!quiet) return 0u;
#pragma unroll
    for (int d = 0; d < 3; ++d) { s_grp[d * CS + w] = g2.g[d]; s_grp[(3 + d) * CS + w] = g2.gv[d]; }
    s_grp[6 * CS + w] = g2.q[0]; s_grp[7 * CS + w] = g2.q[1]; s_grp[8 * CS + w] = g2.qv[0]; s_grp[9 * CS + w] = g2.qv[1];
    uint32_t contacts = priv & 0x7fffu;
    if (over_table(g2.g[0], g2.g[1]) && g2.g[2] - kGZMin < kMargin) contacts |= pair_bit(0, 1);
    const uint32_t code = finalize_at<ID, E, LEAN>(p, li, ksel, w, contacts, s_cub, s_grp, s_msc, s_bits);
    ws.n_st += 1.f; ws.n_inv += (float)inv;
    ws.r_sum -= (float)(code & 1u); ws.n_ep += (float)((code >> 1) & 1u); ws.n_su += (float)((code >> 2) & 1u);
    return 1u | (code << 1);
}

// One full-physics env-step of slab column pw at step k: the env's cubes are loaded into registers, stepped there
// (sim_step_reg) and written back to the slab column.
template <int ID, int E, bool LEAN>
__device__ __forceinline__ uint32_t full_step_item(uint32_t* __restrict__ st, const StepArgs& p, int64_t li, int pw, int k,
                                               float* s_scr, float* s_cub, float* s_grp, uint32_t* s_msc,
                                               uint32_t* s_bits, WarpStats& ws) {
    using C = Cfg<ID>;
    constexpr int NB = C::NB, CS = 32 * E;
    const uint32_t flags = pf_flags(s_msc[CS + pw]);
    const float4 a4 = fetch_action<ID, LEAN>(st, p, li, k, (int)(flags & 0xffu));
    float a[4] = {a4.x, a4.y, a4.z, a4.w};
    int inv = 0;
    clip_action(a, inv);
    Grip g;
#pragma unroll
    for (int d = 0; d < 3; ++d) { g.g[d] = s_grp[d * CS + pw]; g.gv[d] = s_grp[(3 + d) * CS + pw]; }
    g.q[0] = s_grp[6 * CS + pw]; g.q[1] = s_grp[7 * CS + pw]; g.qv[0] = s_grp[8 * CS + pw]; g.qv[1] = s_grp[9 * CS + pw];
    uint32_t contacts = 0;
    bool is_static;
    if constexpr (Async<ID, E>::REG_PASS) {
        CubeRegs<NB> q;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float* bb = s_cub + (9 * b) * CS + pw;
            q.x[b] = bb[0]; q.y[b] = bb[CS]; q.z[b] = bb[2 * CS]; q.c[b] = bb[3 * CS]; q.s[b] = bb[4 * CS];
            q.vx[b] = bb[5 * CS]; q.vy[b] = bb[6 * CS]; q.vz[b] = bb[7 * CS]; q.w[b] = bb[8 * CS];
        }
        is_static = sim_step_reg<NB, C::BG, C::VAR>(g, a, q, (int)((flags >> 9) & 7u), contacts);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float* bb = s_cub + (9 * b) * CS + pw;
            bb[0] = q.x[b]; bb[CS] = q.y[b]; bb[2 * CS] = q.z[b]; bb[3 * CS] = q.c[b]; bb[4 * CS] = q.s[b];
            bb[5 * CS] = q.vx[b]; bb[6 * CS] = q.vy[b]; bb[7 * CS] = q.vz[b]; bb[8 * CS] = q.w[b];
        }
        if (Async<ID, E>::YAW) yaw_refresh<NB, CS>(s_cub + pw);
    } else {
        is_static = sim_step_col<NB, CS, C::BG>(g, a, Col<NB, CS, 32>(s_cub + pw, s_scr), (int)((flags >> 9) & 7u), contacts);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) { s_grp[d * CS + pw] = g.g[d]; s_grp[(3 + d) * CS + pw] = g.gv[d]; }
    s_grp[6 * CS + pw] = g.q[0]; s_grp[7 * CS + pw] = g.q[1]; s_grp[8 * CS + pw] = g.qv[0]; s_grp[9 * CS + pw] = g.qv[1];
    s_msc[CS + pw] = pf_make((is_static ? 0x80000000u : 0u) | (contacts & ~gripper_pair_mask()), flags);
    const uint32_t code = finalize_at<ID, E, LEAN>(p, li, k, pw, contacts, s_cub, s_grp, s_msc, s_bits);
    ws.n_st += 1.f; ws.n_inv += (float)inv; ws.n_wrk += 1.f;
    ws.r_sum -= (float)(code & 1u); ws.n_ep += (float)((code >> 1) & 1u); ws.n_su += (float)((code >> 2) & 1u);
    return code;
}

__device__ __forceinline__ void add_warp_stats(const StepArgs& p, WarpStats ws, int lane) {
    ws.n_ep = warp_sum(ws.n_ep); ws.n_su = warp_sum(ws.n_su); ws.n_st = warp_sum(ws.n_st);
    ws.n_inv = warp_sum(ws.n_inv); ws.r_sum = warp_sum(ws.r_sum); ws.n_wrk = warp_sum(ws.n_wrk);
    if (lane == 0 && p.stats) {
        if (ws.n_wrk != 0.f) atomicAdd(p.stats + BP_STAT_WORKER_STEPS, (double)ws.n_wrk);
        if (ws.n_ep != 0.f) atomicAdd(p.stats + BP_STAT_EPISODES, (double)ws.n_ep);
        if (ws.n_su != 0.f) atomicAdd(p.stats + BP_STAT_SUCCESSES, (double)ws.n_su);
        if (ws.n_st != 0.f) atomicAdd(p.stats + BP_STAT_STEPS, (double)ws.n_st);
        if (ws.n_inv != 0.f) atomicAdd(p.stats + BP_STAT_INVALID, (double)ws.n_inv);
        if (ws.r_sum != 0.f) atomicAdd(p.stats + BP_STAT_REWARD_SUM, (double)ws.r_sum);
    }
}

// ---- pieces shared by the one-warp (async) and two-warp (duo) kernels -------------------------------

// load the slab columns e = e0, e0 + estep, ... of this warp's lanes from the handle's field-major state.
// Returns, packed 8 bits per e, the (t | succ << 7) each of the lane's envs starts the launch with: slab_flush
// rebuilds the success rows from them.  Rolled loops on purpose: this code runs once per CTA, and unrolled
// it was 26 KB of SASS streaming through the instruction cache the resident warps' hot loops live in.
template <int ID, int E>
__device__ __forceinline__ uint32_t slab_load(const uint32_t* __restrict__ st, const StepArgs& p, int64_t wbase, int lane, int e0, int estep,
                                              float* s_cub, float* s_grp, uint32_t* s_msc) {
    static_assert(E <= 4, "the start flags of a lane's envs are packed into one 32-bit word");
    constexpr int NB = Cfg<ID>::NB, CS = 32 * E;
    uint32_t f0 = 0u;
#pragma unroll 1
    for (int e = e0; e < E; e += estep) {
        const int w = e * 32 + lane;
        const int64_t li = wbase + w;
        const bool live = li < p.B;
        const uint32_t* q = st + p.env0 + (live ? li : 0);
        // five / nine independent loads in flight per lane (fully rolled, every load waited for the one before: 144
        // serial DRAM round trips per lane at the start of each CTA)
#pragma unroll 1
        for (int d0 = 0; d0 < 10; d0 += 5) {
            uint32_t v[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) v[j] = q[(int64_t)(d0 + j) * p.stateB];
#pragma unroll
            for (int j = 0; j < 5; ++j) s_grp[(d0 + j) * CS + w] = __uint_as_float(v[j]);
        }
#pragma unroll 1
        for (int d0 = 0; d0 < 9 * NB; d0 += 9) {
            uint32_t v[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) v[j] = q[(int64_t)(10 + d0 + j) * p.stateB];
#pragma unroll
            for (int j = 0; j < 9; ++j) s_cub[(d0 + j) * CS + w] = __uint_as_float(v[j]);
        }
        if (Async<ID, E>::YAW) yaw_refresh<NB, CS>(s_cub + w);
        int f = 10 + 9 * NB;
        s_msc[w] = q[(int64_t)f * p.stateB]; ++f;                        // touch
        const uint32_t flags = q[(int64_t)f * p.stateB]; ++f;
        const uint32_t priv = q[(int64_t)f * p.stateB]; ++f;
        s_msc[CS + w] = pf_make(priv, flags);
        s_msc[2 * CS + w] = live ? 0u : (uint32_t)p.K;                   // status: steps done (dead columns start finished)
        const uint32_t t0 = (flags & 0xffu) < 127u ? (flags & 0xffu) : 127u;   // any t >= T behaves alike: saturate into 7 bits
        f0 |= (t0 | (((flags >> 8) & 1u) << 7)) << (8 * e);
    }
    return f0;
}

// reward / success rows of the whole launch (coalesced 128-byte stores) + the slab written back.
// Only the reward bits are kept on chip; info['is_success'] is the latch of fetch_env.py:275-281 -- set by
// the first -0.0 reward of an episode, cleared by the reset after step T -- so its rows follow from the
// reward bits and the (t, latch) the launch started with (f0, see slab_load).
template <int ID, int E, bool LEAN>
__device__ __forceinline__ void slab_flush(uint32_t* __restrict__ st, const StepArgs& p, int64_t wbase, int lane, int e0, int estep,
                                           const float* s_cub, const float* s_grp, const uint32_t* s_msc, const uint32_t* s_bits,
                                           const uint32_t f0) {
    constexpr int NB = Cfg<ID>::NB, CS = 32 * E;
#pragma unroll 1
    for (int e = e0; e < E; e += estep) {
        const int w = e * 32 + lane;
        const int64_t li = wbase + w;
        if (li >= p.B) continue;
        if (p.reward || p.success) {
            int t = (int)((f0 >> (8 * e)) & 0x7fu);
            uint32_t latch = (f0 >> (8 * e + 7)) & 1u;
#pragma unroll 1
            for (int k = 0; k < p.K; ++k) {
                const uint32_t fail = (s_bits[k * E + e] >> lane) & 1u;
                latch |= fail ^ 1u;
                const int64_t row = LEAN ? (int64_t)((uint32_t)k * (uint32_t)p.B + (uint32_t)li) : step_row(p, k, li);
                if (p.reward) store_reward(p.reward + row, fail);
                if (p.success) p.success[row] = (float)latch;
                t += 1;
                if (p.auto_reset && t >= kT) { t = 0; latch = 0u; }   // TimeLimit + reset (_reset_sim clears the latch, :254)
            }
        }
        uint32_t* q = st + p.env0 + li;
#pragma unroll 1
        for (int d = 0; d < 10; ++d) q[(int64_t)d * p.stateB] = __float_as_uint(s_grp[d * CS + w]);
#pragma unroll 1
        for (int d = 0; d < 9 * NB; ++d) q[(int64_t)(10 + d) * p.stateB] = __float_as_uint(s_cub[d * CS + w]);
        int f = 10 + 9 * NB;
        const uint32_t pf = s_msc[CS + w];
        q[(int64_t)f * p.stateB] = s_msc[w]; ++f;
        q[(int64_t)f * p.stateB] = pf_flags(pf); ++f;
        q[(int64_t)f * p.stateB] = pf & kPrivMask; ++f;
    }
}

// this lane's least-advanced runnable env (status words are read volatile: in the duo kernel the
// worker warp clears pending flags concurrently).  Returns the slab column or -1; ksel = its step index.
template <int E>
__device__ __forceinline__ int pick_ready(const uint32_t* s_status, int lane, int K, int& ksel) {
    int sel = -1;
    ksel = 0x7fffffff;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const uint32_t s = *reinterpret_cast<const volatile uint32_t*>(s_status + e * 32 + lane);
        const int k = (int)(s & 0xffffu);
        if (!(s & kPendingBit) && k < K && k < ksel) { ksel = k; sel = e; }
    }
    return sel < 0 ? -1 : sel * 32 + lane;
}

// `p` is __grid_constant__: the out-of-line helpers take it by reference, which would otherwise force a
// per-thread local-memory copy (ncu: 158 M local-load sectors per launch, most of them missing the
// shrunken L1 and a third of those reaching DRAM).
// LEAN: the instantiation for the plain fused step (time-major rows, caller-provided actions, no done / goal /
// actions_out / reset-observation outputs).  It exists for CODE SIZE: ncu shows the kernel bound by instruction
// delivery (gcc__cache_requests_type_instruction at 80-93 % of peak, sm__icc hit rate 84-92 %) -- the hot code of
// the scheduler, the quiet path, the pass and finalize_step together just exceeds the SM's instruction cache, so
// every instruction the common case does not need is left out of it.
template <int ID, int E, bool LEAN>
__global__ void __launch_bounds__(32, E == 4 ? 12 : (E == 3 ? 15 : 20)) step_kernel_async(uint32_t* __restrict__ st, const __grid_constant__ StepArgs p) {
    using A = Async<ID, E>;
    using C = Cfg<ID>;
    constexpr int NB = C::NB, CS = A::CS;
    extern __shared__ __align__(16) uint32_t smem[];
    float* s_cub = reinterpret_cast<float*>(smem);                    // [9*NB][CS]
    float* s_grp = s_cub + A::W_CUB;                                  // [10][CS]
    uint32_t* s_msc = reinterpret_cast<uint32_t*>(s_grp + A::W_GRP);  // [3][CS]: touch, priv | flags, status
    float* s_col = reinterpret_cast<float*>(s_msc + A::W_MSC);        // [4*NB][32] per-lane substep scratch (ids with > 2 cubes only)
    uint16_t* s_pend = reinterpret_cast<uint16_t*>(s_col + A::W_COL); // [CS] ring: env-steps waiting for a full-physics pass
    uint16_t* s_rset = s_pend + CS;                                    // [CS] ring: envs waiting for their reset
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_col + A::W_COL) + A::W_PEND;  // [K][E]: reward-fail masks over 32 envs

    const int lane = threadIdx.x;
    const int64_t wbase = (int64_t)blockIdx.x * CS;  // launch-local index of the warp's env 0
    WarpStats ws;

    // ---- load the slab: env (e, lane) is launch-local env wbase + e*32 + lane, slab column e*32 + lane
    const uint32_t f0 = slab_load<ID, E>(st, p, wbase, lane, 0, 1, s_cub, s_grp, s_msc);
    for (int i = lane; i < A::W_BITS; i += 32) s_bits[i] = 0u;
    __syncwarp();

    // two rings (an env is in at most one of them): no compaction, a pass / reset pass takes from the head
    int pend_head = 0, npend = 0, rset_head = 0, nreset = 0;
    float n_iter = 0.f, n_pass = 0.f;
    const unsigned lt = (1u << lane) - 1u;
    const int reset_min = p.tune & 0xff, pass_min = (p.tune >> 8) & 0xff;   // queue lengths that trigger a reset / full-physics pass
    auto ring = [](int i) { return i >= CS ? i - CS : i; };
    while (true) {
        n_iter += 1.f;
        // ------------------------------------------------ pick this lane's least-advanced ready env
        // (letting idle lanes steal runnable envs of other lanes was tried and changes nothing: the
        // iteration count is set by the pass chain of the warp's hardest env, not by lane starvation)
        int ksel;
        const int w = pick_ready<E>(s_msc + 2 * CS, lane, p.K, ksel);
        const bool have = w >= 0;
        const unsigned any_ready = __ballot_sync(0xffffffffu, have);
        if (!any_ready && npend == 0 && nreset == 0) break;
        bool want_full = false, want_reset = false;
        if (have) {
            const uint32_t q = try_quiet_step<ID, E, LEAN>(st, p, wbase + w, w, ksel, s_cub, s_grp, s_msc, s_bits, ws);
            if (q) {
                want_reset = p.auto_reset && ((q >> 2) & 1u);  // episode done: queue the reset
                s_msc[2 * CS + w] = (uint32_t)(ksel + 1) | (want_reset ? kPendingBit : 0u);
            } else {
                want_full = true;
            }
        }
        // ------------------------------------------------ queue full-physics env-steps and resets
        const unsigned fullm = __ballot_sync(0xffffffffu, want_full), resetm = __ballot_sync(0xffffffffu, want_reset);
        if (want_full) {
            s_pend[ring(ring(pend_head + npend) + __popc(fullm & lt))] = (uint16_t)w;
            s_msc[2 * CS + w] = (uint32_t)ksel | kPendingBit;
        }
        if (want_reset) s_rset[ring(ring(rset_head + nreset) + __popc(resetm & lt))] = (uint16_t)w;
        npend += __popc(fullm);
        nreset += __popc(resetm);
        __syncwarp();
        // ------------------------------------------------ full-physics pass with all lanes busy
        // (keeping an env in the pass for 2/3/4 consecutive steps was measured: 2.39 -> 2.16 / 1.93 / 1.76e9
        // env-steps/s -- longer, emptier passes -- so a pass advances every item by exactly one step)
        // a pass also fires when it would be fuller than the scheduler iterations currently are, by a margin
        // (p.tune bit 16, margin in bits 17-22): measured +1.8 % at margin 8 (0: +1.6 %, 16: +0.6 %)
        if (npend >= pass_min || (npend > 0 && (!any_ready || (((p.tune >> 16) & 1) && npend >= __popc(any_ready) + ((p.tune >> 17) & 63))))) {
            const int take = npend < 32 ? npend : 32;
            n_pass += 1.f;
            bool rs = false;
            int pw = 0;
            if (lane < take) {
                pw = s_pend[ring(pend_head + lane)];
                const int k = (int)(s_msc[2 * CS + pw] & 0xffffu);
                const uint32_t code = full_step_item<ID, E, LEAN>(st, p, wbase + pw, pw, k, s_col + lane, s_cub, s_grp, s_msc, s_bits, ws);
                rs = p.auto_reset && (code & 2u);
                s_msc[2 * CS + pw] = (uint32_t)(k + 1) | (rs ? kPendingBit : 0u);  // clears the pending flag unless a reset is due
            }
            const unsigned rm = __ballot_sync(0xffffffffu, rs);
            pend_head = ring(pend_head + take);
            npend -= take;
            if (rs) s_rset[ring(ring(rset_head + nreset) + __popc(rm & lt))] = (uint16_t)pw;
            nreset += __popc(rm);
            __syncwarp();
        }
        // ------------------------------------------------ reset pass: RobotEnv.reset for up to 32 finished envs at once
        // (resetting inside finalize_step ran the Philox / Box-Muller / rejection-sampling code with 1-3
        // active lanes: 5 % of the kernel's time for one reset per 50 env-steps)
        if (nreset >= reset_min || (nreset > 0 && !any_ready)) {
            const int take = nreset < 32 ? nreset : 32;
            if (lane < take) {
                const int rw = s_rset[ring(rset_head + lane)];
                reset_env_slab<ID, CS, LEAN>(st, p, wbase + rw, s_cub + rw, s_grp + rw, s_msc + rw);
                s_msc[2 * CS + rw] &= 0xffffu;
            }
            rset_head = ring(rset_head + take);
            nreset -= take;
            __syncwarp();
        }
    }

    __syncwarp();
    slab_flush<ID, E, LEAN>(st, p, wbase, lane, 0, 1, s_cub, s_grp, s_msc, s_bits, f0);
    add_warp_stats(p, ws, lane);
    if (lane == 0 && p.stats) {
        atomicAdd(p.stats + 6, (double)n_iter);   // diagnostics: scheduler iterations and full-physics passes per warp
        atomicAdd(p.stats + 7, (double)n_pass);
    }
}


}  // namespace bp
