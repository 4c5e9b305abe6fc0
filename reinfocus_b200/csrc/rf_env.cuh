// Env-step glue on the device: everything reference environments/vector_environment.py:75-164
// does around the render + focus measure of a step, for the compositions the example envs
// use (reference examples/custom_environments.py) - transformer, enders, Delta / Normalized
// observers, rewarders, the initializer's generator and the same-step auto-reset - so that
// states, observations and rewards never visit the host.
//
// The arithmetic follows the NumPy expressions of the reference classes operation by
// operation, dtype promotion included (float32 unless noted; Python floats are weak under
// NumPy 2, so they enter float32 expressions as float32). Every float operation is an
// explicit round-to-nearest intrinsic: nothing here may be contracted into an FMA.
#pragma once

#include <cstdint>

namespace rf {

enum { kEnvDiscreteMove = 0, kEnvContinuousJump = 1, kEnvContinuousMove = 2, kEnvDiscreteJump = 3 };
enum { kEnvActionsInt32 = 0, kEnvActionsInt64 = 1, kEnvActionsFloat32 = 2 };
// ender / rewarder expression trees, flattened to postfix programs
enum { kEnderTimeLimit = 0, kEnderDiverging = 1, kEnderOnTarget = 2, kEnderStopped = 3, kEnderEndless = 4,
       kEnderAnd = 5, kEnderOr = 6 };
enum { kRewardDelta = 0, kRewardDistance = 1, kRewardObservation = 2, kRewardOnTarget = 3, kRewardStopped = 4,
       kRewardAdd = 5, kRewardMul = 6 };
enum { kObsElement = 0, kObsFocus = 1, kObsDelta = 2, kObsNormalized = 3 };
constexpr int kEnvMaxMoves = 32;
constexpr int kEnvMaxNodes = 24;     // nodes of the ender / rewarder programs
constexpr int kEnvMaxObsNodes = 16;  // nodes of the observer program
constexpr int kEnvMaxObsDim = 16;    // columns of an observation
constexpr int kEnvMaxObsStack = 32;  // float32 values alive while the observer program runs
constexpr int kEnvMaxRanges = 4;
constexpr int kEnvMaxWindow = 16;  // StoppedEnder: early_end_steps + 1 positions
constexpr int kEnvPreThreads = 1024;

struct EnderNode {
    int kind;
    int i0, i1;    // state indices it looks at
    int steps;     // TimeLimit: max_steps; Diverging / OnTarget / Stopped: early_end_steps
    float value;   // Diverging: threshold; OnTarget: radius; Stopped: span
    int slot;      // first of its per-env state rows in EnvArrays::node_state
};

struct RewardNode {
    int kind;
    int i0, i1;        // state indices (Observation: the observation column)
    int slot;          // Delta / Stopped: the row with the previous value of state[i0]
    float f0, f1;      // Delta: reward, scale; Distance: span, float32(high - low); Stopped: threshold
    double d0, d1;     // Distance: low; OnTarget: off, delta; Stopped: reward
};

// One node of the observer tree, in postfix order (reference state_observer.py:103-517).
// ELEMENT pushes a state element, FOCUS the focus value; DELTA / NORMALIZED consume the last
// `arg` vectors side by side (the hstack of WrapperObserver.wrapped_observations).
struct ObsNode {
    int kind;
    int arg;     // ELEMENT: state index; DELTA / NORMALIZED: number of child observers
    int flag;    // DELTA: include_original
    int offset;  // DELTA: first of its old-observation columns; NORMALIZED: first mid / scale column
};

struct EnvParams {
    int n;
    int transformer;             // kEnvDiscreteMove ... kEnvDiscreteJump
    int n_moves;
    double moves[kEnvMaxMoves];  // DiscreteMoveTransformer keeps its action set in float64
    float limit_lo, limit_hi;
    float jump_span;             // float32(limits[1] - limits[0]), the difference taken in Python
    float jump_threshold;        // ContinuousJump / ContinuousMove: moves shorter than this are ignored
    float move_speed;            // ContinuousMoveTransformer
    float jumps[kEnvMaxMoves];   // DiscreteJumpTransformer keeps its action set in float32
    int n_enders, n_rewards;
    EnderNode enders[kEnvMaxNodes];
    RewardNode rewards[kEnvMaxNodes];
    // observer: any tree of IndexedElementObservers and one FocusObserver under DeltaObservers
    // and NormalizedObservers, as a postfix program
    int n_obs;
    ObsNode obs_nodes[kEnvMaxObsNodes];
    int obs_dim;      // columns the root leaves
    int delta_width;  // old-observation columns per env, over all DeltaObservers
    float obs_mid[kEnvMaxObsStack], obs_scale[kEnvMaxObsStack];  // NormalizedObservers, by their offsets
    // RangedInitializer: per state element up to kEnvMaxRanges (low, high - low) pairs; with
    // more than one, Generator.choice picks the pair before Generator.uniform draws in it
    int init_options[2];
    double init_low[2][kEnvMaxRanges], init_range[2][kEnvMaxRanges];
};

struct EnvArrays {
    float *states;        // [n, 2]  target, focus plane
    float *new_states;    // [n, 2]  first states of restarted episodes, by reset rank
    int *reset_rank;      // [n]     position among this step's restarted envs, -1 if none
    float *old_obs;       // [n, delta_width]  DeltaObservers: their children's previous values
    uint32_t *node_state; // [rows, n] per-env state of the ender / rewarder nodes (int or float bits)
    uint64_t *generator;  // [6]     PCG64DXSM state hi, lo, increment hi, lo, has_uint32, uinteger
    int *counters;        // [2]     number of restarted envs, invalid-action flag
};

// ------------------------------------------------------------------ numpy.random.PCG64DXSM

using u128 = unsigned __int128;
constexpr uint64_t kPcgCheapMultiplier = 0xda942042e4dd58b5ull;

__device__ inline u128 pcg_make(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | lo; }

// state after `delta` steps of state = state * M + inc (Brown's O(log delta) LCG skip)
__device__ inline u128 pcg_advance(u128 state, u128 inc, uint64_t delta) {
    u128 acc_mult = 1, acc_plus = 0, cur_mult = kPcgCheapMultiplier, cur_plus = inc;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    return acc_mult * state + acc_plus;
}

// Generator.random's double: DXSM output of the pre-step state, top 53 bits * 2^-53
__device__ inline double pcg_next_double(u128 &state, u128 inc) {
    uint64_t hi = (uint64_t)(state >> 64);
    const uint64_t lo = (uint64_t)state | 1;
    hi ^= hi >> 32;
    hi *= kPcgCheapMultiplier;
    hi ^= hi >> 48;
    hi *= lo;
    state = state * kPcgCheapMultiplier + inc;
    return __dmul_rn((double)(hi >> 11), 1.0 / 9007199254740992.0);
}

__device__ inline uint64_t pcg_next64(u128 &state, u128 inc) {
    uint64_t hi = (uint64_t)(state >> 64);
    const uint64_t lo = (uint64_t)state | 1;
    hi ^= hi >> 32;
    hi *= kPcgCheapMultiplier;
    hi ^= hi >> 48;
    hi *= lo;
    state = state * kPcgCheapMultiplier + inc;
    return hi;
}

// numpy's buffered 32-bit draw (pcg64_cm_next32): the low half of a fresh 64-bit draw, then
// its high half
struct PcgBuffer {
    uint32_t has, value;
};
__device__ inline uint32_t pcg_next32(u128 &state, u128 inc, PcgBuffer &buffer) {
    if (buffer.has) {
        buffer.has = 0;
        return buffer.value;
    }
    const uint64_t v = pcg_next64(state, inc);
    buffer.has = 1;
    buffer.value = (uint32_t)(v >> 32);
    return (uint32_t)v;
}

// Generator.integers(0, n) as Generator.choice uses it: Lemire's bounded draw on 32 bits
// (numpy/random/src/distributions/distributions.c buffered_bounded_lemire_uint32); n = 1
// consumes nothing
__device__ inline uint32_t pcg_choice(u128 &state, u128 inc, PcgBuffer &buffer, uint32_t n) {
    if (n <= 1) return 0;
    const uint32_t rng = n - 1, exclusive = n;
    uint64_t m = (uint64_t)pcg_next32(state, inc, buffer) * exclusive;
    uint32_t leftover = (uint32_t)m;
    if (leftover < exclusive) {
        const uint32_t threshold = (0xffffffffu - rng) % exclusive;
        while (leftover < threshold) {
            m = (uint64_t)pcg_next32(state, inc, buffer) * exclusive;
            leftover = (uint32_t)m;
        }
    }
    return (uint32_t)(m >> 32);
}

// ------------------------------------------------------------------------------- helpers

// numpy.clip = minimum(maximum(x, lo), hi)
__device__ inline float clip_f32(float x, float lo, float hi) {
    x = x < lo ? lo : x;
    return x > hi ? hi : x;
}

__device__ inline float env_gap(float target, float plane) { return fabsf(__fsub_rn(target, plane)); }

// observer.observe / observer.reset of one env: the observer tree as a postfix program over
// a stack of float32 vectors that lie side by side in `out`. A wrapper takes its children's
// vectors as one (hstack; the float64 focus value is cast to float32 there). DeltaObserver:
// changes since the previous step - zeros when the episode just started (`fresh`, reference
// state_observer.py:283-290) - after the values themselves with include_original;
// NormalizedObserver: clip((values - mid) / scale, -1, 1). Leaves the root's vector in
// out[0 .. obs_dim).
__device__ inline void observe(const EnvParams &p, const EnvArrays &a, int env, const float *state, double focus,
                               bool fresh, float *out) {
    int widths[kEnvMaxObsNodes];
    int top = 0, used = 0;
    for (int k = 0; k < p.n_obs; ++k) {
        const ObsNode &node = p.obs_nodes[k];
        if (node.kind == kObsElement || node.kind == kObsFocus) {
            out[used++] = node.kind == kObsFocus ? __double2float_rn(focus) : state[node.arg];
            widths[top++] = 1;
            continue;
        }
        int width = 0;
        for (int j = 0; j < node.arg; ++j) width += widths[--top];
        float *values = out + used - width;
        if (node.kind == kObsDelta) {
            float *previous = a.old_obs + (size_t)p.delta_width * env + node.offset;
            float *changes = node.flag ? values + width : values;
            for (int c = 0; c < width; ++c) {
                const float current = values[c];
                changes[c] = fresh ? 0.0f : __fsub_rn(current, previous[c]);
                previous[c] = current;
            }
            if (node.flag) {
                used += width;
                width *= 2;
            }
        } else {  // kObsNormalized
            for (int c = 0; c < width; ++c)
                values[c] = clip_f32(__fdiv_rn(__fsub_rn(values[c], p.obs_mid[node.offset + c]),
                                               p.obs_scale[node.offset + c]), -1.0f, 1.0f);
        }
        widths[top++] = width;
    }
}

// per-env state rows of the strategy nodes
__device__ inline int &node_int(const EnvArrays &a, int n, int row, int env) {
    return reinterpret_cast<int *>(a.node_state)[(size_t)row * n + env];
}
__device__ inline float &node_float(const EnvArrays &a, int n, int row, int env) {
    return reinterpret_cast<float *>(a.node_state)[(size_t)row * n + env];
}

// ender.step(states) for every node, then is_truncated() of the tree (nothing ever
// terminates: reference episode_ender.py:92-98). Postfix evaluation on a small stack.
__device__ inline bool enders_step(const EnvParams &p, const EnvArrays &a, int env, const float *state) {
    bool stack[kEnvMaxNodes];
    int top = 0;
    for (int k = 0; k < p.n_enders; ++k) {
        const EnderNode &node = p.enders[k];
        switch (node.kind) {
            case kEnderTimeLimit: {  // episode_ender.py:590-656
                const int steps = ++node_int(a, p.n, node.slot, env);
                stack[top++] = steps >= node.steps;
                break;
            }
            case kEnderDiverging: {  // :112-207: the gap grew by more than the threshold
                const float gap = env_gap(state[node.i0], state[node.i1]);
                int &count = node_int(a, p.n, node.slot, env);
                float &last = node_float(a, p.n, node.slot + 1, env);
                if (gap > __fadd_rn(last, node.value)) ++count;
                last = gap;
                stack[top++] = count >= node.steps;
                break;
            }
            case kEnderOnTarget: {  // :273-369: consecutive steps within the radius
                int &count = node_int(a, p.n, node.slot, env);
                count = env_gap(state[node.i0], state[node.i1]) < node.value ? count + 1 : 0;
                stack[top++] = count >= node.steps;
                break;
            }
            case kEnderStopped: {  // :466-587: span of the last steps + 1 positions
                const int window = node.steps + 1;
                int &filled = node_int(a, p.n, node.slot, env);
                for (int j = 0; j + 1 < window; ++j)
                    node_float(a, p.n, node.slot + 1 + j, env) = node_float(a, p.n, node.slot + 2 + j, env);
                node_float(a, p.n, node.slot + window, env) = state[node.i0];
                filled = min(filled + 1, window);
                float lo = state[node.i0], hi = lo;
                for (int j = 0; j < window; ++j) {
                    const float v = node_float(a, p.n, node.slot + 1 + j, env);
                    lo = fminf(lo, v);  // rows not filled yet hold NaN: ignored like nanmin / nanmax
                    hi = fmaxf(hi, v);
                }
                stack[top++] = filled >= window && fabsf(__fsub_rn(hi, lo)) < node.value;
                break;
            }
            case kEnderEndless:
                stack[top++] = false;
                break;
            case kEnderAnd:
                --top;
                stack[top - 1] = stack[top - 1] & stack[top];
                break;
            default:  // kEnderOr
                --top;
                stack[top - 1] = stack[top - 1] | stack[top];
                break;
        }
    }
    return stack[0];
}

// ender.reset(new_state, done) for one restarted env
__device__ inline void enders_reset(const EnvParams &p, const EnvArrays &a, int env, const float *state) {
    for (int k = 0; k < p.n_enders; ++k) {
        const EnderNode &node = p.enders[k];
        if (node.kind == kEnderTimeLimit || node.kind == kEnderOnTarget) {
            node_int(a, p.n, node.slot, env) = 0;
        } else if (node.kind == kEnderDiverging) {
            node_int(a, p.n, node.slot, env) = 0;
            node_float(a, p.n, node.slot + 1, env) = env_gap(state[node.i0], state[node.i1]);
        } else if (node.kind == kEnderStopped) {
            const int window = node.steps + 1;
            for (int j = 0; j + 1 < window; ++j) node_float(a, p.n, node.slot + 1 + j, env) = __int_as_float(0x7fc00000);
            node_float(a, p.n, node.slot + window, env) = state[node.i0];
            node_int(a, p.n, node.slot, env) = 1;
        }
    }
}

// rewarder.reward(states, observations): postfix evaluation with NumPy's result types - a
// value is float32 until a float64 operand (bool * Python float) joins it
struct RewardValue {
    double value;
    bool wide;
};

__device__ inline RewardValue rewards_eval(const EnvParams &p, const EnvArrays &a, int env, const float *state,
                                            const float *obs) {
    RewardValue stack[kEnvMaxNodes];
    int top = 0;
    for (int k = 0; k < p.n_rewards; ++k) {
        const RewardNode &node = p.rewards[k];
        switch (node.kind) {
            case kRewardDelta: {  // episode_rewarder.py:86-156: |x - previous| * reward / scale
                float &previous = node_float(a, p.n, node.slot, env);
                const float moved = fabsf(__fsub_rn(state[node.i0], previous));
                previous = state[node.i0];
                stack[top++] = {(double)__fdiv_rn(__fmul_rn(moved, node.f0), node.f1), false};
                break;
            }
            case kRewardDistance: {  // :159-207: (1 - gap / span) * (high - low) + low
                const float gap = env_gap(state[node.i0], state[node.i1]);
                const float unit = __fsub_rn(1.0f, __fdiv_rn(gap, node.f0));
                stack[top++] = {(double)__fadd_rn(__fmul_rn(unit, node.f1), (float)node.d0), false};
                break;
            }
            case kRewardObservation:  // :210-238
                stack[top++] = {(double)obs[node.i0], false};
                break;
            case kRewardOnTarget:  // :241-292: bool * Python float -> float64
                stack[top++] = {__dadd_rn(env_gap(state[node.i0], state[node.i1]) < node.f0 ? node.d1 : 0.0, node.d0),
                                true};
                break;
            case kRewardStopped: {  // :361-429
                float &previous = node_float(a, p.n, node.slot, env);
                const float moved = fabsf(__fsub_rn(state[node.i0], previous));
                previous = state[node.i0];
                stack[top++] = {moved < node.f0 ? node.d0 : 0.0, true};
                break;
            }
            default: {  // kRewardAdd / kRewardMul
                --top;
                RewardValue &l = stack[top - 1];
                const RewardValue r = stack[top];
                if (l.wide || r.wide) {
                    l.value = node.kind == kRewardAdd ? __dadd_rn(l.value, r.value) : __dmul_rn(l.value, r.value);
                    l.wide = true;
                } else {
                    l.value = node.kind == kRewardAdd ? (double)__fadd_rn((float)l.value, (float)r.value)
                                                      : (double)__fmul_rn((float)l.value, (float)r.value);
                }
                break;
            }
        }
    }
    return stack[0];
}

// rewarder.reset(new_state, new_observations, done) for one restarted env
__device__ inline void rewards_reset(const EnvParams &p, const EnvArrays &a, int env, const float *state) {
    for (int k = 0; k < p.n_rewards; ++k) {
        const RewardNode &node = p.rewards[k];
        if (node.kind == kRewardDelta || node.kind == kRewardStopped)
            node_float(a, p.n, node.slot, env) = state[node.i0];
    }
}

// --------------------------------------------------------------------------- step, part 1
//
// One block walks the envs in chunks (an ordered scan is needed: restarted envs draw their
// first states in env order from one generator, reference vector_environment.py:139 ->
// state_initializer.py:63-69). Per env: transformer, ender step, is_truncated; restarted
// envs get their rank and their new state. reset_all: every env restarts (env.reset()).
__global__ void __launch_bounds__(kEnvPreThreads)
env_pre_kernel(EnvParams p, EnvArrays a, const void *actions, int action_kind, int reset_all) {
    __shared__ int warp_totals[kEnvPreThreads / 32];
    __shared__ int chunk_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u128 gen_state = pcg_make(a.generator[0], a.generator[1]);
    const u128 gen_inc = pcg_make(a.generator[2], a.generator[3]);
    int base = 0;
    bool invalid = false;
    // several ranges for some element: Generator.choice consumes a data-dependent number of
    // buffered 32-bit draws before each uniform, so the restarts draw one after another
    const bool choosing = p.init_options[0] > 1 || p.init_options[1] > 1;

    for (int chunk = 0; chunk < p.n; chunk += kEnvPreThreads) {
        const int i = chunk + tid;
        bool done = false;
        if (i < p.n) {
            if (reset_all) {
                done = true;
            } else {
                float target = a.states[2 * i], plane = a.states[2 * i + 1];
                if (p.transformer == kEnvDiscreteMove) {
                    // DiscreteMoveTransformer: float32 state += float64 move, one rounding
                    long long action = action_kind == kEnvActionsInt64
                                           ? ((const long long *)actions)[i]
                                           : (long long)((const int *)actions)[i];
                    if (action < 0) action += p.n_moves;  // NumPy index wrap-around
                    if (action < 0 || action >= p.n_moves) {
                        invalid = true;
                        action = 0;
                    }
                    plane = __double2float_rn(__dadd_rn((double)plane, p.moves[action]));
                } else if (p.transformer == kEnvDiscreteJump) {
                    // DiscreteJumpTransformer: the action picks the position (float32 set)
                    long long action = action_kind == kEnvActionsInt64
                                           ? ((const long long *)actions)[i]
                                           : (long long)((const int *)actions)[i];
                    if (action < 0) action += p.n_moves;
                    if (action < 0 || action >= p.n_moves) {
                        invalid = true;
                        action = 0;
                    }
                    plane = p.jumps[action];
                } else if (p.transformer == kEnvContinuousMove) {
                    // ContinuousMoveTransformer: clip(a, -1, 1) * speed, ignored when shorter
                    // than the stop threshold: state += (|move| > threshold) * move
                    const float move = __fmul_rn(clip_f32(((const float *)actions)[i], -1.0f, 1.0f), p.move_speed);
                    plane = __fadd_rn(plane, __fmul_rn(fabsf(move) > p.jump_threshold ? 1.0f : 0.0f, move));
                } else {
                    // ContinuousJumpTransformer: ((a + 1) / 2) * (hi - lo) + lo, ignored when
                    // closer than the stop threshold
                    const float act = ((const float *)actions)[i];
                    const float fraction = __fdiv_rn(__fadd_rn(act, 1.0f), 2.0f);
                    const float destination =
                        __fadd_rn(__fmul_rn(fraction, p.jump_span), p.limit_lo);
                    if (fabsf(__fsub_rn(plane, destination)) > p.jump_threshold) plane = destination;
                }
                if (p.transformer != kEnvContinuousJump) {  // the jump transformer does not clip
                    target = clip_f32(target, p.limit_lo, p.limit_hi);
                    plane = clip_f32(plane, p.limit_lo, p.limit_hi);
                }
                a.states[2 * i] = target;
                a.states[2 * i + 1] = plane;
                const float state[2] = {target, plane};
                done = enders_step(p, a, i, state);
            }
        }
        // ordered rank of the restarted envs
        const unsigned ballot = __ballot_sync(0xffffffffu, done);
        if (lane == 0) warp_totals[warp] = __popc(ballot);
        __syncthreads();
        if (warp == 0) {
            int v = warp_totals[lane];
#pragma unroll
            for (int offset = 1; offset < 32; offset <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, v, offset);
                if (lane >= offset) v += up;
            }
            warp_totals[lane] = v;  // inclusive
            if (lane == 31) chunk_total = v;
        }
        __syncthreads();
        const int rank = base + (warp > 0 ? warp_totals[warp - 1] : 0) +
                         __popc(ballot & ((1u << lane) - 1));
        if (i < p.n) a.reset_rank[i] = done ? rank : -1;
        if (done && !choosing) {
            // RangedInitializer, one range per element: uniform(low, high, size=(k, 2)) =
            // low + range * double, in env-major order, rounded to float32 by astype - two
            // draws per restart, so restart r starts 2 r draws into the stream
            u128 state = pcg_advance(gen_state, gen_inc, 2ull * (uint64_t)rank);
#pragma unroll
            for (int element = 0; element < 2; ++element) {
                const double unit = pcg_next_double(state, gen_inc);
                a.new_states[2 * rank + element] = __double2float_rn(
                    __dadd_rn(p.init_low[element][0], __dmul_rn(p.init_range[element][0], unit)));
            }
        }
        base += chunk_total;
        __syncthreads();
    }
    if (invalid) atomicExch(&a.counters[1], 1);
    if (tid == 0) {
        u128 state = gen_state;
        if (choosing) {
            PcgBuffer buffer{(uint32_t)a.generator[4], (uint32_t)a.generator[5]};
            for (int r = 0; r < base; ++r) {
                for (int element = 0; element < 2; ++element) {
                    const uint32_t k = pcg_choice(state, gen_inc, buffer, (uint32_t)p.init_options[element]);
                    const double unit = pcg_next_double(state, gen_inc);
                    a.new_states[2 * r + element] = __double2float_rn(
                        __dadd_rn(p.init_low[element][k], __dmul_rn(p.init_range[element][k], unit)));
                }
            }
            a.generator[4] = buffer.has;
            a.generator[5] = buffer.value;
        } else {
            state = pcg_advance(gen_state, gen_inc, 2ull * (uint64_t)base);
        }
        a.generator[0] = (uint64_t)(state >> 64);
        a.generator[1] = (uint64_t)state;
        a.counters[0] = base;
    }
}

// --------------------------------------------------------------------------- step, part 2
//
// focus_main[i]: focus value of env i's scene after the transformer; focus_reset[r]: focus
// value of the first scene of the r-th restarted env (rendered as batch position r).
// Observations, rewards (from the pre-reset observations, reference
// vector_environment.py:128-130), then the restart bookkeeping of every strategy.
__global__ void env_post_kernel(EnvParams p, EnvArrays a, const double *focus_main,
                                const double *focus_reset, float *obs, double *rewards,
                                uint8_t *truncated, int reset_all) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const int rank = a.reset_rank[i];
    float out[kEnvMaxObsStack];
    if (!reset_all) {
        const float state[2] = {a.states[2 * i], a.states[2 * i + 1]};
        observe(p, a, i, state, focus_main[i], false, out);
        rewards[i] = rewards_eval(p, a, i, state, out).value;
        truncated[i] = rank >= 0;
    }
    if (rank >= 0) {
        // same-step auto-reset (reference vector_environment.py:137-151)
        const float state[2] = {a.new_states[2 * rank], a.new_states[2 * rank + 1]};
        a.states[2 * i] = state[0];
        a.states[2 * i + 1] = state[1];
        enders_reset(p, a, i, state);
        observe(p, a, i, state, focus_reset[rank], true, out);
        rewards_reset(p, a, i, state);
    }
    for (int c = 0; c < p.obs_dim; ++c) obs[(size_t)p.obs_dim * i + c] = out[c];
}

// ------------------------------------------------------------------- scene packing (a1, a2)

struct ScenePacking {
    float world_tan;                       // float32(tan(radians(r_size / 2)))
    float half_width, half_height;         // float32(hw), float32(hh)
    float full_width, full_height;         // float32(2 * hw), float32(2 * hh)
    float origin[3], u[3], v[3], w[3];
};

// FastWorlds._make_device_data (reference graphics/world.py:107-123) and
// FastCameras._make_device_data (reference graphics/camera.py:144-179) from device-resident
// target / focus-plane positions, same float32 operations in the same order
__global__ void pack_scene_kernel(int n, const float *targets, const float *planes, int stride,
                                  ScenePacking k, float *world, float *cam_dyn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float target = targets[(size_t)i * stride], f = planes[(size_t)i * stride];
    world[2 * i] = __fmul_rn(target, k.world_tan);
    world[2 * i + 1] = -target;
    const float wide = __fmul_rn(k.half_width, f), tall = __fmul_rn(k.half_height, f);
    const float full_wide = __fmul_rn(k.full_width, f), full_tall = __fmul_rn(k.full_height, f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float total = __fadd_rn(__fadd_rn(__fmul_rn(k.u[c], wide), __fmul_rn(k.v[c], tall)),
                                      __fmul_rn(k.w[c], f));
        cam_dyn[9 * i + c] = __fsub_rn(k.origin[c], total);
        cam_dyn[9 * i + 3 + c] = __fmul_rn(k.u[c], full_wide);
        cam_dyn[9 * i + 6 + c] = __fmul_rn(k.v[c], full_tall);
    }
}

}  // namespace rf
