"""CPU ORACLE (test infrastructure, NOT the product) for the callers' data path either side of the env
hot path -- SURVEY.md section 8(f) rows 2-4.  numpy restatements, each citing what it follows
(paths relative to /root/reference/gym_blocks).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
legs may import this module; nothing under blockpuzzle_gym_b200/ does.

PARITY STATUS: "parity unpinned" for `sample_her_transitions`, `ReplayBufferOracle` and
`NormalizerOracle`: their bodies live in OpenAI baselines (baselines.her.her / replay_buffer /
normalizer, master of about April 2018 -- the reference pins no version, setup.py:6, and baselines is
neither vendored nor installed here).  They restate the published algorithm and are anchored on the
reference's own call sites (config.py:107-123, ddpg.py:100-120, 158-190, 214-222).  The reference's
random draws (np.random.randint / uniform) are replaced by one Philox4x32-10 block per transition, stream 3,
exactly as oracle/blockphys_oracle.c: bpo_her_relabel does, so the device sampler can be replayed.
`discounted_returns` and `trim` follow in-tree code (policy_gradient/rollout.py) line by line and are PINNED to it:
tests/test_ref_callers_cpu.py runs the reference's own policy-gradient RolloutStudent (its return accumulation) and
calls its `trim` from source, bit for bit; the episode-batch layout is pinned to the reference's RolloutStudent +
convert_episode_to_batch_major run unmodified, and `compute_reward` to BlocksEnv.compute_reward called directly.
"""
import numpy as np

M32 = 0xFFFFFFFF


def philox4x32_vec(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint32 arrays (Random123; the same block function as bpo_philox4x32)."""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) & M32 for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint64(k0 & M32)
    k1 = np.uint64(k1 & M32)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        h0, l0 = p0 >> np.uint64(32), p0 & M32
        h1, l1 = p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = (h1 ^ c1 ^ k0) & M32, l1, (h0 ^ c3 ^ k1) & M32, l0
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def u01(x):
    return (np.asarray(x, np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(5.9604644775390625e-08)


def compute_reward(achieved_goal, goal, info=None):
    """BlocksEnv.compute_reward, fetch_env.py:135-143."""
    achieved_goal = np.asarray(achieved_goal, np.float32)
    goal = np.asarray(goal, np.float32)
    d = np.sum(achieved_goal * goal, axis=-1)                  # :141
    c = np.count_nonzero(goal, axis=-1)                        # :142
    return -(d != c).astype(np.float32)                        # :143


def preprocess_og(o, ag, g, clip_obs=200.0):
    """DDPG._preprocess_og (ddpg.py:111-120) with relative_goals = False (config.py:37)."""
    o = np.clip(o, -clip_obs, clip_obs)                        # :118
    g = np.clip(g, -clip_obs, clip_obs)                        # :119
    return o, g


def make_sample_her_transitions(replay_strategy, replay_k, reward_fun, seed=0):
    """baselines.her.her.make_sample_her_transitions [upstream], constructed at config.py:121 with
    replay_strategy / replay_k of config.py:49-50 and reward_fun of config.py:110-111."""
    if replay_strategy == 'future':
        future_p = 1 - (1. / (1 + replay_k))
    else:  # 'replay_strategy' == 'none'
        future_p = 0

    def _sample_her_transitions(episode_batch, batch_size_in_transitions, index_offset=0):
        T = episode_batch['u'].shape[1]
        rollout_batch_size = episode_batch['u'].shape[0]
        batch_size = batch_size_in_transitions

        # one Philox block per transition replaces upstream's four np.random draws
        gi = np.arange(batch_size, dtype=np.uint64) + np.uint64(index_offset)
        w0, w1, w2, w3 = philox4x32_vec(gi & M32, gi >> np.uint64(32), 3, 0, seed & M32, (seed >> 32) & M32)
        # Select which episodes and time steps to use.
        episode_idxs = ((w0.astype(np.uint64) * np.uint64(rollout_batch_size)) >> np.uint64(32)).astype(np.int64)  # randint(0, B, n)
        t_samples = ((w1.astype(np.uint64) * np.uint64(T)) >> np.uint64(32)).astype(np.int64)                      # randint(T, size=n)
        transitions = {key: episode_batch[key][episode_idxs, t_samples].copy() for key in episode_batch.keys()}

        # Select future time indexes proportional with probability future_p.
        her_indexes = np.where(u01(w2) < np.float32(future_p))                                                    # uniform(size=n) < future_p
        future_offset = u01(w3) * (T - t_samples).astype(np.float32)                                              # uniform(size=n) * (T - t)
        future_offset = future_offset.astype(np.int64)
        future_t = (t_samples + 1 + future_offset)[her_indexes]

        # Replace goal with achieved goal but only for the previously-selected HER transitions.
        future_ag = episode_batch['ag'][episode_idxs[her_indexes], future_t]
        transitions['g'][her_indexes] = future_ag

        # Reconstruct info dictionary for reward computation.
        info = {}
        for key, value in transitions.items():
            if key.startswith('info_'):
                info[key.replace('info_', '')] = value

        # Re-compute reward since we may have substituted the goal.
        reward_params = {k: transitions[k] for k in ['ag_2', 'g']}
        reward_params['info'] = info
        transitions['r'] = reward_fun(**reward_params)

        transitions = {k: transitions[k].reshape(batch_size, *transitions[k].shape[1:]) for k in transitions.keys()}
        assert transitions['u'].shape[0] == batch_size_in_transitions
        # replay bookkeeping (not part of upstream's dict): which draw produced each row
        fut = np.full(batch_size, -1, np.int32)
        fut[her_indexes] = future_t
        transitions['_ep_idx'] = episode_idxs.astype(np.int32)
        transitions['_t'] = t_samples.astype(np.int32)
        transitions['_future_t'] = fut
        return transitions

    return _sample_her_transitions


class ReplayBufferOracle:
    """baselines.her.replay_buffer.ReplayBuffer [upstream], constructed at ddpg.py:100-106."""

    def __init__(self, buffer_shapes, size_in_transitions, T, sample_transitions, rng=None):
        self.buffer_shapes = buffer_shapes
        self.size = size_in_transitions // T
        self.T = T
        self.sample_transitions = sample_transitions
        self.buffers = {key: np.empty([self.size, *shape], np.float32) for key, shape in buffer_shapes.items()}
        self.current_size = 0
        self.n_transitions_stored = 0
        self.rng = rng or np.random.RandomState(0)

    @property
    def full(self):
        return self.current_size == self.size

    def sample(self, batch_size, **kw):
        buffers = {}
        assert self.current_size > 0
        for key in self.buffers.keys():
            buffers[key] = self.buffers[key][:self.current_size]
        buffers['o_2'] = buffers['o'][:, 1:, :]
        buffers['ag_2'] = buffers['ag'][:, 1:, :]
        transitions = self.sample_transitions(buffers, batch_size, **kw)
        for key in (['r', 'o_2', 'ag_2'] + list(self.buffers.keys())):
            assert key in transitions, "key %s missing from transitions" % key
        return transitions

    def store_episode(self, episode_batch):
        batch_sizes = [len(episode_batch[key]) for key in episode_batch.keys()]
        assert np.all(np.array(batch_sizes) == batch_sizes[0])
        batch_size = batch_sizes[0]
        idxs = self._get_storage_idx(batch_size)
        for key in self.buffers.keys():
            self.buffers[key][idxs] = episode_batch[key]
        self.n_transitions_stored += batch_size * self.T

    def get_current_episode_size(self):
        return self.current_size

    def get_current_size(self):
        return self.current_size * self.T

    def get_transitions_stored(self):
        return self.n_transitions_stored

    def clear_buffer(self):
        self.current_size = 0

    def _get_storage_idx(self, inc=None):
        inc = inc or 1   # size increment
        assert inc <= self.size, "Batch committed to replay is too large!"
        # go consecutively until you hit the end, and then go randomly.
        if self.current_size + inc <= self.size:
            idx = np.arange(self.current_size, self.current_size + inc)
        elif self.current_size < self.size:
            overflow = inc - (self.size - self.current_size)
            idx_a = np.arange(self.current_size, self.size)
            idx_b = self.rng.randint(0, self.current_size, overflow)
            idx = np.concatenate([idx_a, idx_b])
        else:
            idx = self.rng.randint(0, self.size, inc)
        # update replay size
        self.current_size = min(self.size, self.current_size + inc)
        if inc == 1:
            idx = idx[0]
        return idx


class NormalizerOracle:
    """baselines.her.normalizer.Normalizer [upstream] on one process (the MPI average of recompute_stats is
    the identity at world size 1); used at ddpg.py:185-188 for o_stats.  float64 accumulators here: the
    upstream float32 sums depend on numpy's pairwise summation order, so GPU parity is tolerance-based."""

    def __init__(self, size, eps=1e-2, default_clip_range=np.inf):
        self.size = size
        self.eps = eps
        self.default_clip_range = default_clip_range
        self.local_sum = np.zeros(size, np.float64)
        self.local_sumsq = np.zeros(size, np.float64)
        self.local_count = np.zeros(1, np.float64)
        self.total_sum = np.zeros(size, np.float64)
        self.total_sumsq = np.zeros(size, np.float64)
        self.total_count = np.ones(1, np.float64)
        self.mean = np.zeros(size, np.float32)
        self.std = np.ones(size, np.float32)

    def update(self, v):
        v = np.asarray(v, np.float64).reshape(-1, self.size)
        self.local_sum += v.sum(axis=0)
        self.local_sumsq += (np.square(v)).sum(axis=0)
        self.local_count[0] += v.shape[0]

    def recompute_stats(self):
        local_count, local_sum, local_sumsq = self.local_count.copy(), self.local_sum.copy(), self.local_sumsq.copy()
        self.local_count[...] = 0
        self.local_sum[...] = 0
        self.local_sumsq[...] = 0
        self.total_sum += local_sum
        self.total_sumsq += local_sumsq
        self.total_count += local_count
        mean = self.total_sum / self.total_count
        self.mean = mean.astype(np.float32)
        self.std = np.sqrt(np.maximum(np.square(self.eps), self.total_sumsq / self.total_count - np.square(mean))).astype(np.float32)

    def normalize(self, v, clip_range=None):
        if clip_range is None:
            clip_range = self.default_clip_range
        return np.clip((np.asarray(v, np.float32) - self.mean) / self.std, -clip_range, clip_range)


def store_episode_stats(episode_batch, sample_transitions, o_stats, env_name, clip_obs=200.0, index_offset=0):
    """The update_stats branch of DDPG.store_episode, ddpg.py:166-190."""
    episode_batch = dict(episode_batch)
    episode_batch['o_2'] = episode_batch['o'][:, 1:, :]                              # :168
    episode_batch['ag_2'] = episode_batch['ag'][:, 1:, :]                            # :169
    num_normalizing_transitions = episode_batch['u'].shape[0] * episode_batch['u'].shape[1]   # transitions_in_episode_batch :170
    transitions = sample_transitions(episode_batch, num_normalizing_transitions, index_offset=index_offset)   # :171
    o, o_2, g, ag = transitions['o'], transitions['o_2'], transitions['g'], transitions['ag']
    transitions['o'], transitions['g'] = preprocess_og(o, ag, g, clip_obs)          # :174
    if 'Variation' in env_name:                                                      # :180-181
        o = transitions['o'][:, 1:]
    else:
        o = transitions['o']
    o_stats.update(o)                                                                # :185
    o_stats.recompute_stats()                                                        # :188
    return transitions


def discounted_returns(r, gamma):
    """policy_gradient/rollout.py:255-258 for a whole episode batch: r [T][B] (time-major, as the loop sees
    it) -> G [T][B] float64, the `returns` list stacked."""
    r = np.asarray(r)
    T = r.shape[0]
    returns = []
    for t in range(T):
        r_new = np.zeros(r.shape[1])                      # :217 (float64)
        r_new[:] = r[t]                                   # :236 r_new[i] = r
        returns.append(r_new.copy())                      # :255
        for t_ in range(t):                               # :256
            r_new = r_new.copy()
            returns[t_] += gamma ** (t - t_) * r_new      # :258
    return np.stack(returns)


# constants of policy_gradient/rollout.py:12-22
COLOR_FEATURES = 4
ENV_FEATURES = 10
BLOCK_BASE_FEATURES = 15
BLOCK_FEATURES = BLOCK_BASE_FEATURES + COLOR_FEATURES
GREY, RED, GREEN, BLUE = 0, 1, 2, 3


def get_color(one_hot):                                    # policy_gradient/rollout.py:24-26
    assert len(one_hot) == 4
    return np.argmax(one_hot)


def trim(o, g, ag, dimo, dimg, env_name, num_objs=4):
    """RolloutStudent.trim, batched branch (policy_gradient/rollout.py:105-108 and :139-171), restated with
    index lists: same selection, same order."""
    o, g, ag = np.asarray(o), np.asarray(g), np.asarray(ag)
    if o.shape[-1] == dimo:                                            # :107-108 nothing to trim
        return o, g, ag
    max_objs = int(g.shape[1] ** 0.5)                                  # :143
    keep = [i for i in range(g.shape[1]) if i // max_objs < num_objs and i % max_objs < num_objs]   # :144-146
    assert len(keep) == dimg                                           # :149
    g_, ag_ = g[:, keep], ag[:, keep]
    if 'Variation' in env_name:                                        # :152-167
        rows = []
        for row in o:
            parts = [row[1:ENV_FEATURES + 1]]                          # drop the leading block count
            for j in range(max_objs - 2):                              # every (padded) block slot
                start = ENV_FEATURES + 1 + j * BLOCK_FEATURES
                if get_color(row[start + BLOCK_BASE_FEATURES:start + BLOCK_FEATURES]) in (GREEN, BLUE):
                    parts.append(row[start:start + BLOCK_BASE_FEATURES])
            rows.append(np.concatenate(parts))
        o_ = np.stack(rows)
    else:
        o_ = o[:, :dimo]                                               # :169
    assert o_.shape[1] == dimo                                         # :170
    return o_, g_, ag_
