// bp_replay.cu -- the callers' data path around the env hot path:
//
//   bp_her_sample          ReplayBuffer.sample + _sample_her_transitions [upstream baselines.her] as wired by
//                          config.py:107-123 and used by ddpg.py:106,168-171,214-222, with _preprocess_og's clip
//                          (ddpg.py:111-120) and the observation-normaliser sums (ddpg.py:166-190) fused in
//   bp_moments             Normalizer.update [upstream baselines.her.normalizer], call site ddpg.py:185
//   bp_discounted_returns  the return accumulation of policy_gradient/rollout.py:255-258
//   bp_trim                RolloutStudent.trim, policy_gradient/rollout.py:105-171 (batched branch)
//
// All four are HBM-bound streaming / gather kernels: rows are moved with 128-bit accesses when the row
// length allows, a block's threads walk whole output rows in address order (coalesced stores), and the
// reductions stay in registers / shared memory until one atomic per block and column.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <string>

#include "../../include/blockpuzzle_b200.h"
#include "bp_device.cuh"
#include "bp_host.h"

namespace bp {

constexpr int kSampleThreads = 256;  // = transitions per block
constexpr int kPartials = 2048;       // per-block table of partial column sums (doubles)

// np.clip(x, -c, c) (ddpg.py:118-119): NaN passes through like numpy's; c <= 0 disables the clip
__device__ __forceinline__ float clip_sym(float x, float c) { return c > 0.0f ? (x < -c ? -c : (x > c ? c : x)) : x; }

// Copy rows src[s_off[r]] -> dst[r] for the block's `rows` transitions, V floats per access.  Thread ->
// (row slot, chunk) is fixed, so a thread always serves the same columns: consecutive threads touch
// consecutive addresses of a row and consecutive rows of the output.  With STATS the thread keeps the sum
// and sum of squares of its V columns and leaves them in s_acc[row slot][2 * dim] (shared, double); the
// caller adds the row slots up (kPartials doubles cover every dim <= 256: slots * 2 * dim <= 512 * V).
template <int V, bool STATS>
__device__ __forceinline__ void gather_rows(const float* __restrict__ src, const int64_t* s_off, float* __restrict__ dst, int dim, int rows,
                                            float clip, double* s_acc) {
    const int chunks = dim / V;
    const int rpi = kSampleThreads / chunks;  // rows per iteration (dim <= 256 floats)
    const int c = (int)threadIdx.x % chunks, rr = (int)threadIdx.x / chunks;
    double sum[V], sq[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { sum[j] = 0.0; sq[j] = 0.0; }
    if (rr < rpi) {
        // U independent row loads are issued before the first one is consumed (a rolled / exit-checked loop leaves one
        // load in flight per thread: the gathers are bound by memory latency, as moments_kernel was); rows past the end
        // are predicated off
        constexpr int U = 4;
        for (int r = rr; r < rows; r += rpi * U) {
            float v[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int ru = r + u * rpi;
                if constexpr (V == 4) {
                    const float4 x = ru < rows ? __ldg(reinterpret_cast<const float4*>(src + s_off[ru]) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[u][0] = x.x; v[u][1] = x.y; v[u][2] = x.z; v[u][3] = x.w;
                } else {
                    v[u][0] = ru < rows ? __ldg(src + s_off[ru] + c) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int ru = r + u * rpi;
                if (ru < rows) {
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        v[u][j] = clip_sym(v[u][j], clip);
                        if (STATS) { const double d = (double)v[u][j]; sum[j] += d; sq[j] += d * d; }
                    }
                    if constexpr (V == 4) reinterpret_cast<float4*>(dst + (int64_t)ru * dim)[c] = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
                    else dst[(int64_t)ru * dim + c] = v[u][0];
                }
            }
        }
        if (STATS) {  // this thread's partial sums: slot [rr][2 * dim] of the block's table (no atomics)
#pragma unroll
            for (int j = 0; j < V; ++j) {
                s_acc[rr * 2 * dim + c * V + j] = sum[j];
                s_acc[rr * 2 * dim + dim + c * V + j] = sq[j];
            }
        }
    }
}

template <bool STATS>
__device__ __forceinline__ void gather_any(const float* src, const int64_t* s_off, float* dst, int dim, int rows, float clip, double* s_acc) {
    if (!dst) return;
    // 128-bit path: rows and the tensors' bases must keep 16-byte alignment
    if ((dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0)
        gather_rows<4, STATS>(src, s_off, dst, dim, rows, clip, s_acc);
    else
        gather_rows<1, STATS>(src, s_off, dst, dim, rows, clip, s_acc);
}

struct HerSampleArgs {
    const float *ep_o, *ep_u, *ep_g, *ep_ag, *ep_succ;  // [B][T+1][dimo], [B][T][dimu], [B][T][dimg], [B][T+1][dimg], [B][T]
    int B, T, dimo, dimu, dimg;
    int64_t n;
    float future_p, clip_obs;
    uint32_t k0, k1;
    int64_t index_offset;
    int32_t *ep_idx, *t_idx, *fut_t;
    float *o, *o2, *u, *g, *ag, *ag2, *r, *succ;
    double* stats;  // [2 * dimo + 1]: sum, sum of squares of the (clipped) o rows, row count
};

// One block = 256 transitions.  Phase 1 (thread per transition): the Philox draw -> (episode, t, her,
// future_t) exactly as bpo_her_relabel / her_relabel_kernel, the reward of (ag_2, relabelled g), the
// index outputs.  Phase 2 (block-cooperative): the row gathers.
__global__ void __launch_bounds__(kSampleThreads) her_sample_kernel(const __grid_constant__ HerSampleArgs a) {
    __shared__ int64_t s_o[kSampleThreads], s_u[kSampleThreads], s_ag[kSampleThreads], s_g[kSampleThreads];
    __shared__ double s_acc[kPartials];
    const int64_t base = (int64_t)blockIdx.x * kSampleThreads;
    const int rows = (int)((a.n - base) < kSampleThreads ? (a.n - base) : kSampleThreads);
    const int tid = (int)threadIdx.x;
    const bool stats = a.stats != nullptr && a.o != nullptr;
    if (tid < rows) {
        const int64_t i = base + tid;
        const uint64_t gi = (uint64_t)(i + a.index_offset);
        const U4 w = philox4x32((uint32_t)gi, (uint32_t)(gi >> 32), 3u, 0u, a.k0, a.k1);
        const int e = (int)__umulhi(w.x, (uint32_t)a.B);           // episode_idxs = randint(0, B)
        const int t = (int)__umulhi(w.y, (uint32_t)a.T);           // t_samples = randint(T)
        const bool her = u01(w.z) < a.future_p;                    // uniform(size) < future_p
        const int off = (int)(u01(w.w) * (float)(a.T - t));        // (uniform * (T - t)).astype(int)
        const int ft = t + 1 + off;
        const int64_t ag_off = ((int64_t)e * (a.T + 1) + t) * a.dimg;
        const int64_t g_off = her ? ((int64_t)e * (a.T + 1) + ft) * a.dimg : ((int64_t)e * a.T + t) * a.dimg;
        s_o[tid] = ((int64_t)e * (a.T + 1) + t) * a.dimo;
        s_u[tid] = ((int64_t)e * a.T + t) * a.dimu;
        s_ag[tid] = ag_off;
        s_g[tid] = her ? -(g_off + 1) : g_off;                     // negative: the row lives in ep_ag
        // reward_fun(ag_2, g, info) = compute_reward (config.py:110-111 -> fetch_env.py:135-143)
        const float* ag2 = a.ep_ag + ag_off + a.dimg;
        const float* gs = (her ? a.ep_ag : a.ep_g) + g_off;
        float d = 0.0f;
        int c = 0;
        if ((a.dimg & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.ep_ag) | reinterpret_cast<uintptr_t>(a.ep_g)) & 15) == 0) {
            for (int k = 0; k < a.dimg / 4; ++k) {   // same left-to-right sum, 128-bit loads
                const float4 x = __ldg(reinterpret_cast<const float4*>(ag2) + k), y = __ldg(reinterpret_cast<const float4*>(gs) + k);
                d = d + x.x * y.x; d = d + x.y * y.y; d = d + x.z * y.z; d = d + x.w * y.w;
                c += (y.x != 0.0f) + (y.y != 0.0f) + (y.z != 0.0f) + (y.w != 0.0f);
            }
        } else {
            for (int k = 0; k < a.dimg; ++k) {
                const float x = __ldg(ag2 + k), y = __ldg(gs + k);
                d = d + x * y;
                c += (y != 0.0f);
            }
        }
        if (a.r) store_reward(a.r + i, d != (float)c);
        if (a.ep_idx) a.ep_idx[i] = e;
        if (a.t_idx) a.t_idx[i] = t;
        if (a.fut_t) a.fut_t[i] = her ? ft : -1;
        if (a.succ && a.ep_succ) a.succ[i] = __ldg(a.ep_succ + (int64_t)e * a.T + t);
    }
    __syncthreads();
    float* const o = a.o ? a.o + base * a.dimo : nullptr;
    float* const o2 = a.o2 ? a.o2 + base * a.dimo : nullptr;
    float* const u = a.u ? a.u + base * a.dimu : nullptr;
    float* const ag = a.ag ? a.ag + base * a.dimg : nullptr;
    float* const ag2 = a.ag2 ? a.ag2 + base * a.dimg : nullptr;
    float* const g = a.g ? a.g + base * a.dimg : nullptr;
    if (stats) gather_any<true>(a.ep_o, s_o, o, a.dimo, rows, a.clip_obs, s_acc);
    else gather_any<false>(a.ep_o, s_o, o, a.dimo, rows, a.clip_obs, nullptr);
    gather_any<false>(a.ep_o + a.dimo, s_o, o2, a.dimo, rows, a.clip_obs, nullptr);          // o_2 = o[:, 1:]
    if (a.ep_u) gather_any<false>(a.ep_u, s_u, u, a.dimu, rows, 0.0f, nullptr);
    gather_any<false>(a.ep_ag, s_ag, ag, a.dimg, rows, 0.0f, nullptr);
    gather_any<false>(a.ep_ag + a.dimg, s_ag, ag2, a.dimg, rows, 0.0f, nullptr);            // ag_2 = ag[:, 1:]
    if (g) {
        // the goal row comes from ep_ag (relabelled) or ep_g: resolve per row, then one gather over a common base
        __syncthreads();
        const float* lo = a.ep_ag < a.ep_g ? a.ep_ag : a.ep_g;
        if (tid < rows) {
            const int64_t v = s_g[tid];
            s_g[tid] = v < 0 ? (a.ep_ag - lo) + (-v - 1) : (a.ep_g - lo) + v;
        }
        __syncthreads();
        const bool v4 = (a.dimg & 3) == 0 &&
                        ((reinterpret_cast<uintptr_t>(a.ep_ag) | reinterpret_cast<uintptr_t>(a.ep_g) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
        if (v4) gather_rows<4, false>(lo, s_g, g, a.dimg, rows, a.clip_obs, nullptr);
        else gather_rows<1, false>(lo, s_g, g, a.dimg, rows, a.clip_obs, nullptr);
    }
    if (stats) {
        __syncthreads();
        const bool v4 = (a.dimo & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.ep_o) | reinterpret_cast<uintptr_t>(o)) & 15) == 0;
        const int slots_all = kSampleThreads / (v4 ? a.dimo / 4 : a.dimo);
        const int slots = slots_all < rows ? slots_all : rows;         // row slots that saw at least one row
        for (int j = tid; j < 2 * a.dimo; j += kSampleThreads) {
            double t = 0.0;
            for (int k = 0; k < slots; ++k) t += s_acc[k * 2 * a.dimo + j];
            atomicAdd(a.stats + j, t);
        }
        if (tid == 0) atomicAdd(a.stats + 2 * a.dimo, (double)rows);
    }
}

// bp_her_relabel for rows of CPR float4 chunks (CPR = 1, 2, 4, 8; dimg = 16 -> 4): CPR lanes per transition.  Every lane of a
// row group makes the same Philox draw (redundant, but shuffle-free and the kernel is bound by the random row reads), loads
// ONE float4 of ag_2 = ag[t + 1] and of the goal row (the future ag when relabelled, else g[t]) -- a warp reads 32 / CPR
// random 64-byte rows per tensor and instruction as whole sector pairs -- and stores its chunk of g' (and ag_2) coalesced.
// The reward's dot product is accumulated in the reference's left-to-right order by handing the running sum from lane
// to lane, exactly as compute_reward_coop_kernel does.  Same draws, same outputs as her_sample_kernel (the tests compare
// both with the oracle); measured 54 -> see profiles/README.md: the thread-per-transition phase + block-cooperative
// re-gather of the generic sampler was the cost, not the random reads (tools/gather_ceiling.cu: 35 us for this pattern).
template <int CPR>
__global__ void __launch_bounds__(256) her_relabel_coop_kernel(const float4* __restrict__ ep_ag, const float4* __restrict__ ep_g, const int B,
                                                               const int T, const int64_t n, const float future_p, const uint32_t k0,
                                                               const uint32_t k1, const int64_t index_offset, int32_t* __restrict__ ep_idx,
                                                               int32_t* __restrict__ t_idx, int32_t* __restrict__ fut_t,
                                                               float4* __restrict__ ag2_out, float4* __restrict__ g_out, float* __restrict__ r) {
    const int64_t tid = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t i = tid / CPR;
    const int chunk = (int)(threadIdx.x & (CPR - 1));
    const bool live = i < n;   // a row group never straddles the end: the grid covers n * CPR rounded up to whole blocks
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
    int e = 0, t = 0, ft = 0;
    bool her = false;
    if (live) {
        const uint64_t gi = (uint64_t)(i + index_offset);
        const U4 w = philox4x32((uint32_t)gi, (uint32_t)(gi >> 32), 3u, 0u, k0, k1);
        e = (int)__umulhi(w.x, (uint32_t)B);                       // episode_idxs = randint(0, B)
        t = (int)__umulhi(w.y, (uint32_t)T);                       // t_samples = randint(T)
        her = u01(w.z) < future_p;                                 // uniform(size) < future_p
        ft = t + 1 + (int)(u01(w.w) * (float)(T - t));             // t + 1 + (uniform * (T - t)).astype(int)
        const int64_t ag_row = (int64_t)e * (T + 1) + t;
        x = __ldg(ep_ag + (ag_row + 1) * CPR + chunk);             // ag_2 = ag[:, 1:]
        y = her ? __ldg(ep_ag + ((int64_t)e * (T + 1) + ft) * CPR + chunk) : __ldg(ep_g + ((int64_t)e * T + t) * CPR + chunk);
    }
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
    if (!live) return;
    if (g_out) __stcs(g_out + i * CPR + chunk, y);
    if (ag2_out) __stcs(ag2_out + i * CPR + chunk, x);
    if (chunk == CPR - 1 && r) store_reward(r + i, d != (float)c);
    if (chunk == 0) {
        if (ep_idx) ep_idx[i] = e;
        if (t_idx) t_idx[i] = t;
        if (fut_t) fut_t[i] = her ? ft : -1;
    }
}

// Normalizer.update(v): local_sum += v.sum(0), local_sumsq += (v ** 2).sum(0), local_count += v.shape[0]
// x [n][dim] (row stride ld, first column col0: the Variation rule o[:, 1:] of ddpg.py:180-181) ->
// acc [2 * dim + 1] double, added atomically.  A block of 256 threads walks a slab of rows_per_block rows;
// a thread always serves the same V columns (V = 4: 128-bit loads when rows keep 16-byte alignment), so its
// partial sums stay in registers; one shared and one global atomic per column and block at the end.
template <int V>
__global__ void __launch_bounds__(256, 4) moments_kernel(const float* __restrict__ x, int64_t n, int dim, int ld, int col0, float clip,
                                                      int rows_per_block, double* acc) {
    const int chunks = dim / V;
    const int rpi = 256 / chunks;                       // row slots per block iteration (dim <= 256)
    const int c = (int)threadIdx.x % chunks, rr = (int)threadIdx.x / chunks;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
    double s[V], q[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { s[j] = 0.0; q[j] = 0.0; }
    if (rr < rpi) {
        // U independent 128-bit loads are issued before the first float64 operation consumes one: the kernel is bound by
        // memory latency (ncu, round 2: 28 long-scoreboard stall cycles per issued instruction at 32 warps / SM with the
        // rolled loop), so the bytes in flight per thread set its bandwidth.  Rows past the end contribute exact zeros.
        constexpr int U = V == 4 ? 8 : 4;
        for (int64_t r = r0 + rr; r < r1; r += (int64_t)rpi * U) {
            float v[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t ru = r + (int64_t)u * rpi;
                if constexpr (V == 4) {
                    const float4 t = ru < r1 ? __ldg(reinterpret_cast<const float4*>(x + ru * ld + col0) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
                } else {
                    v[u][0] = ru < r1 ? __ldg(x + ru * ld + col0 + c) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const double d = (double)clip_sym(v[u][j], clip);
                    s[j] += d;
                    q[j] += d * d;
                }
            }
        }
    }
    __shared__ double s_part[kPartials];                // [row slot][2 * dim]
    if (rr < rpi) {
#pragma unroll
        for (int j = 0; j < V; ++j) { s_part[rr * 2 * dim + c * V + j] = s[j]; s_part[rr * 2 * dim + dim + c * V + j] = q[j]; }
    }
    __syncthreads();
    for (int j = (int)threadIdx.x; j < 2 * dim; j += 256) {
        double t = 0.0;
        for (int k = 0; k < rpi; ++k) t += s_part[k * 2 * dim + j];
        atomicAdd(acc + j, t);
    }
    if (threadIdx.x == 0 && r1 > r0) atomicAdd(acc + 2 * dim, (double)(r1 - r0));
}

// policy_gradient/rollout.py:255-258: after step t, returns[t_] += gamma ** (t - t_) * r_t for t_ < t, with
// returns[t] = r_t appended first.  So G[t_] = r[t_] + sum_{j >= 1} pw[j] * r[t_ + j], accumulated in that
// order in float64 (pw[j] = gamma ** j as the host's Python float power): T - 1 - t_ dependent multiply-adds per output,
// an O(T^2) triangle per episode that is kept for bit-exactness.  The kernel is bound by shared-memory wavefronts and
// float64 issue, not by HBM, so the mapping minimises wavefronts per multiply-add:
//   * a block owns 32 episodes, LANE = EPISODE: the rewards are staged transposed ([t][episode], float32), so the reward
//     operand of a warp is one conflict-free wavefront and the power pw[j] is the same address for all lanes (a broadcast);
//   * a warp produces the outputs t_ and T - 1 - t_ of its 32 episodes back to back: T - 1 multiply-adds per pair, the same
//     for every pair, so the warps of a block finish together (the host picks the warp count that divides the pair count);
//   * no lane ever idles inside a chain (all lanes of a warp share t_), and the outputs leave through a shared-memory
//     stage as whole coalesced rows.
// History (1 Mi episodes x T = 50): thread per output, rewards [episode][t] 0.79 ms; thread per output PAIR with float64
// rewards 0.92 ms (4 wavefronts per multiply-add: slower); this form: see profiles/README.md.  T <= kRetMaxT.
constexpr int kRetMaxT = 128;
__global__ void __launch_bounds__(256) discounted_returns_kernel(const float* __restrict__ r, int64_t B, int T, const double* __restrict__ pw,
                                                                 double* __restrict__ G) {
    extern __shared__ double s_ret[];                     // [kRetMaxT] powers | [32][T] output stage | [T][32] rewards (float)
    double* s_pw = s_ret;
    double* s_out = s_ret + kRetMaxT;
    float* s_r = reinterpret_cast<float*>(s_out + 32 * T);
    const int64_t b0 = (int64_t)blockIdx.x * 32;
    const int nb = (int)((B - b0) < 32 ? (B - b0) : 32);
    const int nthr = (int)blockDim.x, lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5), nwarp = nthr >> 5;
    for (int i = (int)threadIdx.x; i < 32 * T; i += nthr) {   // coalesced read of [episode][t], transposed store
        const int e = i / T, t = i - e * T;
        s_r[t * 32 + e] = e < nb ? __ldg(r + b0 * T + i) : 0.0f;
    }
    for (int i = (int)threadIdx.x; i < T; i += nthr) s_pw[i] = pw[i];
    __syncthreads();
    const int P = (T + 1) / 2;                            // output pairs (an odd T pairs its middle output with itself)
    for (int p = warp; p < P; p += nwarp) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int t0 = half == 0 ? p : T - 1 - p;
            double acc = (double)s_r[t0 * 32 + lane];
#pragma unroll 4
            for (int j = 1; j < T - t0; ++j) acc = acc + __dmul_rn(s_pw[j], (double)s_r[(t0 + j) * 32 + lane]);
            s_out[lane * T + t0] = acc;
        }
    }
    __syncthreads();
    for (int i = (int)threadIdx.x; i < nb * T; i += nthr) G[b0 * T + i] = s_out[i];
}

// RolloutStudent.trim (policy_gradient/rollout.py:139-171): cut a batch of padded observations / touch
// matrices down to the `num_objs` objects the expert policy was trained on.
//   g_, ag_: entries i of the max_objs x max_objs matrix with i // max_objs < num_objs and i % max_objs < num_objs
//   o_ (Variation, :151-167): columns 1..10, then the 15 base features of every block whose colour one-hot
//      (argmax of the 4 trailing features, get_color :24-26) is GREEN (2) or BLUE (3), in block order
//   o_ (otherwise, :169): the first dimo_out columns
// A block serves kTrimRows rows: the colour of every (row, block slot) is decided once, the kept slots are
// listed per row in shared memory, then the block's threads write the three outputs in address order.
constexpr int kTrimRows = 64, kTrimMaxBlocks = 8;
__global__ void __launch_bounds__(256) trim_kernel(const float* __restrict__ o, const float* __restrict__ g, const float* __restrict__ ag, int64_t n,
                                                   int dimo_in, int dimg_in, int dimo_out, int num_objs, int max_objs, int variation,
                                                   float* __restrict__ o_out, float* __restrict__ g_out, float* __restrict__ ag_out) {
    __shared__ uint8_t s_keep[kTrimRows][kTrimMaxBlocks];   // [row][k] = k-th kept block slot (0xff: none)
    const int64_t row0 = (int64_t)blockIdx.x * kTrimRows;
    const int rows = (int)((n - row0) < kTrimRows ? (n - row0) : kTrimRows);
    const int tid = (int)threadIdx.x;
    if (variation) {
        const int max_blocks = (dimo_in - 11) / 19;
        if (tid < rows) {
            const float* src = o + (row0 + tid) * dimo_in;
            int kept = 0;
            for (int j = 0; j < max_blocks; ++j) {
                const float* oh = src + 11 + j * 19 + 15;
                int am = 0;                                  // np.argmax: the first maximum wins
                float best = oh[0];
                for (int k = 1; k < 4; ++k)
                    if (oh[k] > best) { best = oh[k]; am = k; }
                if (am == 2 || am == 3) s_keep[tid][kept++] = (uint8_t)j;
            }
            for (; kept < kTrimMaxBlocks; ++kept) s_keep[tid][kept] = 0xff;
        }
        __syncthreads();
    }
    if (o_out) {
        const int total = rows * dimo_out;
        for (int i = tid; i < total; i += 256) {
            const int rr = i / dimo_out, c = i - rr * dimo_out;
            const float* src = o + (row0 + rr) * dimo_in;
            float v;
            if (!variation) {
                v = src[c];
            } else if (c < 10) {                             // ENV_FEATURES after the leading num_blocks scalar
                v = src[1 + c];
            } else {
                const int slot = s_keep[rr][(c - 10) / 15];
                v = slot == 0xff ? 0.0f : src[11 + slot * 19 + (c - 10) % 15];
            }
            o_out[row0 * dimo_out + i] = v;
        }
    }
    const int dimg_out = num_objs * num_objs;
    const int total = rows * dimg_out;
    for (int i = tid; i < total; i += 256) {
        const int rr = i / dimg_out, c = i - rr * dimg_out;
        const int64_t src = (row0 + rr) * dimg_in + (c / num_objs) * max_objs + (c % num_objs);
        if (g_out) g_out[row0 * dimg_out + i] = g[src];
        if (ag_out) ag_out[row0 * dimg_out + i] = ag[src];
    }
}

}  // namespace bp

using namespace bp;

// bp_her_relabel's fast path (see her_relabel_coop_kernel); returns BP_ERR_NOT_IMPLEMENTED when the shape does not qualify
int bp_her_relabel_coop(const float* d_ep_ag, const float* d_ep_g, int32_t B, int32_t T, int32_t dimg, int64_t n, float future_p,
                        uint64_t seed, int64_t index_offset, int32_t* d_ep_idx, int32_t* d_t, int32_t* d_future_t, float* d_ag2,
                        float* d_g, float* d_r, void* stream) {
    const int cpr = dimg / 4;
    const uintptr_t al = reinterpret_cast<uintptr_t>(d_ep_ag) | reinterpret_cast<uintptr_t>(d_ep_g) | reinterpret_cast<uintptr_t>(d_ag2) |
                         reinterpret_cast<uintptr_t>(d_g);
    if ((dimg & 3) != 0 || (cpr != 1 && cpr != 2 && cpr != 4 && cpr != 8) || (al & 15) != 0) return BP_ERR_NOT_IMPLEMENTED;
    const unsigned blocks = (unsigned)((n * cpr + 255) / 256);
    const float4 *a4 = reinterpret_cast<const float4*>(d_ep_ag), *g4 = reinterpret_cast<const float4*>(d_ep_g);
    float4 *o2 = reinterpret_cast<float4*>(d_ag2), *og = reinterpret_cast<float4*>(d_g);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    cudaStream_t s = (cudaStream_t)stream;
    if (cpr == 1) her_relabel_coop_kernel<1><<<blocks, 256, 0, s>>>(a4, g4, B, T, n, future_p, k0, k1, index_offset, d_ep_idx, d_t, d_future_t, o2, og, d_r);
    else if (cpr == 2) her_relabel_coop_kernel<2><<<blocks, 256, 0, s>>>(a4, g4, B, T, n, future_p, k0, k1, index_offset, d_ep_idx, d_t, d_future_t, o2, og, d_r);
    else if (cpr == 4) her_relabel_coop_kernel<4><<<blocks, 256, 0, s>>>(a4, g4, B, T, n, future_p, k0, k1, index_offset, d_ep_idx, d_t, d_future_t, o2, og, d_r);
    else her_relabel_coop_kernel<8><<<blocks, 256, 0, s>>>(a4, g4, B, T, n, future_p, k0, k1, index_offset, d_ep_idx, d_t, d_future_t, o2, og, d_r);
    BP_CU(cudaGetLastError());
    return BP_OK;
}

extern "C" {

int bp_her_sample(const float* d_ep_o, const float* d_ep_u, const float* d_ep_g, const float* d_ep_ag, const float* d_ep_succ,
                  int32_t B, int32_t T, int32_t dimo, int32_t dimu, int32_t dimg, int64_t n, float future_p, float clip_obs,
                  uint64_t seed, int64_t index_offset, int32_t* d_ep_idx, int32_t* d_t, int32_t* d_future_t, float* d_o,
                  float* d_o2, float* d_u, float* d_g, float* d_ag, float* d_ag2, float* d_r, float* d_succ, double* d_stats,
                  void* stream) {
    if (B <= 0 || T <= 0 || dimg <= 0 || dimo <= 0 || n < 0) return bp_fail(BP_ERR_INVALID_ARG, "bad sizes");
    if (dimo > BP_MAX_DIMO || dimg > 256 || dimu < 0 || dimu > 256) return bp_fail(BP_ERR_INVALID_ARG, "row too long");
    if (n == 0) return BP_OK;
    if (!d_ep_ag || !d_ep_g) return bp_fail(BP_ERR_INVALID_ARG, "null episode store");
    if ((d_o || d_o2) && !d_ep_o) return bp_fail(BP_ERR_INVALID_ARG, "o requested without an observation store");
    if (d_u && (!d_ep_u || dimu == 0)) return bp_fail(BP_ERR_INVALID_ARG, "u requested without an action store");
    HerSampleArgs a{};
    a.ep_o = d_ep_o; a.ep_u = d_ep_u; a.ep_g = d_ep_g; a.ep_ag = d_ep_ag; a.ep_succ = d_ep_succ;
    a.B = B; a.T = T; a.dimo = dimo; a.dimu = dimu > 0 ? dimu : 1; a.dimg = dimg; a.n = n;
    a.future_p = future_p; a.clip_obs = clip_obs;
    a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32); a.index_offset = index_offset;
    a.ep_idx = d_ep_idx; a.t_idx = d_t; a.fut_t = d_future_t;
    a.o = d_o; a.o2 = d_o2; a.u = d_u; a.g = d_g; a.ag = d_ag; a.ag2 = d_ag2; a.r = d_r; a.succ = d_succ; a.stats = d_stats;
    const unsigned blocks = (unsigned)((n + kSampleThreads - 1) / kSampleThreads);
    her_sample_kernel<<<blocks, kSampleThreads, 0, (cudaStream_t)stream>>>(a);
    BP_CU(cudaGetLastError());
    return BP_OK;
}

int bp_moments(const float* d_x, int64_t n, int32_t dim, int32_t ld, int32_t col0, float clip, double* d_acc, void* stream) {
    if (n < 0 || dim <= 0 || dim > 256 || ld < dim + col0 || col0 < 0) return bp_fail(BP_ERR_INVALID_ARG, "bad sizes");
    if (n == 0) return BP_OK;
    if (!d_x || !d_acc) return bp_fail(BP_ERR_INVALID_ARG, "null pointer");
    // blocks per SM: every block ends with 2 * dim same-address float64 atomics (~7 ns each, serialised), so few fat blocks win:
    // 4 / 8 / 16 / 32 / 64 per SM -> 61.5 / 67.6 / 71.6 / 77.8 / 133.7 us for [1 Mi, 40]
    static const int mult = [] { const char* e = getenv("BP_MOMENTS_BLOCKS_PER_SM"); return e ? atoi(e) : 4; }();
    int64_t rows_per_block = (n + 148 * mult - 1) / (148 * mult);
    if (rows_per_block < 64) rows_per_block = 64;
    const unsigned blocks = (unsigned)((n + rows_per_block - 1) / rows_per_block);
    const bool v4 = (dim & 3) == 0 && (ld & 3) == 0 && (col0 & 3) == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0;
    if (v4) moments_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_x, n, dim, ld, col0, clip, (int)rows_per_block, d_acc);
    else moments_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_x, n, dim, ld, col0, clip, (int)rows_per_block, d_acc);
    BP_CU(cudaGetLastError());
    return BP_OK;
}

int bp_discounted_returns(const float* d_r, int64_t B, int32_t T, const double* d_gamma_pow, double* d_G, void* stream) {
    if (B < 0 || T <= 0 || T > kRetMaxT) return bp_fail(BP_ERR_INVALID_ARG, "bad sizes (T must be in 1..128)");
    if (B == 0) return BP_OK;
    if (!d_r || !d_gamma_pow || !d_G) return bp_fail(BP_ERR_INVALID_ARG, "null pointer");
    // warps per block: the count in 4..8 that wastes the fewest pair slots (T = 50: 25 pairs -> 5 warps x 5 pairs)
    const int P = (T + 1) / 2;
    int nw = 8, best = 1 << 30;
    for (int w = 8; w >= 4; --w) {
        const int waste = ((P + w - 1) / w) * w - P;
        if (waste < best) { best = waste; nw = w; }
    }
    const size_t smem = sizeof(double) * ((size_t)kRetMaxT + (size_t)32 * T) + sizeof(float) * (size_t)32 * T;
    if (smem > 48 * 1024)   // T > 119: above the default dynamic shared-memory limit (the attribute is per device: set it on every such call)
        BP_CU(cudaFuncSetAttribute(discounted_returns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    discounted_returns_kernel<<<(unsigned)((B + 31) / 32), 32 * nw, smem, (cudaStream_t)stream>>>(d_r, B, T, d_gamma_pow, d_G);
    BP_CU(cudaGetLastError());
    return BP_OK;
}

int bp_trim(const float* d_o, const float* d_g, const float* d_ag, int64_t n, int32_t dimo_in, int32_t dimg_in, int32_t dimo_out,
            int32_t num_objs, int32_t variation, float* d_o_out, float* d_g_out, float* d_ag_out, void* stream) {
    if (n < 0 || dimo_in <= 0 || dimg_in <= 0 || dimo_out <= 0 || num_objs <= 0) return bp_fail(BP_ERR_INVALID_ARG, "bad sizes");
    if (num_objs * num_objs > dimg_in || dimo_out > dimo_in) return bp_fail(BP_ERR_INVALID_ARG, "cannot trim to a larger shape");
    if (variation && (dimo_out != 10 + 15 * (num_objs - 2) || (dimo_in - 11) % 19 != 0))
        return bp_fail(BP_ERR_INVALID_ARG, "Variation trim needs dimo_out = 10 + 15 * (num_objs - 2) and dimo_in = 11 + 19 * max_blocks");
    if (n == 0) return BP_OK;
    if (!d_o || !d_g || !d_ag) return bp_fail(BP_ERR_INVALID_ARG, "null pointer");
    int max_objs = 1;
    while (max_objs * max_objs < dimg_in) ++max_objs;        // (int)(len(g) ** 0.5), rollout.py:143
    if (variation && (dimo_in - 11) / 19 > kTrimMaxBlocks) return bp_fail(BP_ERR_INVALID_ARG, "too many block slots");
    trim_kernel<<<(unsigned)((n + kTrimRows - 1) / kTrimRows), 256, 0, (cudaStream_t)stream>>>(d_o, d_g, d_ag, n, dimo_in, dimg_in, dimo_out, num_objs,
                                                                                               max_objs, variation, d_o_out, d_g_out, d_ag_out);
    BP_CU(cudaGetLastError());
    return BP_OK;
}

}  // extern "C"
